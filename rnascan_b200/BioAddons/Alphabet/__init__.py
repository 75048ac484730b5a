"""Same module path as the reference (rnascan/BioAddons/Alphabet/__init__.py:21-24)."""
from ...seq import SecondaryStructure


class ContextualSecondaryStructure(SecondaryStructure):
    """Alphabet of RNA structural contexts.  The ORDER of `letters` fixes the key order of
    structure backgrounds and PSSMs (and therefore of --bgonly output)."""

    letters = "EHTBLRM"
