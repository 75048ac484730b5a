"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for Biopython <= 1.77.

Biopython is not installed in the build container and cannot be installed
(no network; the reference needs `Bio.Alphabet`, removed in 1.78).  This
package restates, from the published behaviour of Biopython 1.66-1.77, the
handful of classes/functions the reference's hot path touches
(/root/reference/rnascan/rnascan.py:37-40,174,191-193,244-248,263-264,451;
/root/reference/rnascan/BioAddons/motifs/matrix.py:9-15,54,73) so that the
reference's *own* Python files can be imported and executed in this container
to produce golden vectors (tests/golden/make_golden.py).

Nothing in the product package may import this.
"""
__version__ = "1.77-shim"
