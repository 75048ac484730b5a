"""Dot-bracket -> structural context strings (the input alphabet of the one-hot structure scan).

Python face of ``rs_host_annotate_structures``; replaces the reference's stand-alone C++ tool
``scripts/parse_secondary_structure.cpp`` (compiled by hand per README.md:46 and called once per
folded fragment by average_structure.py:78-90).
"""
import numpy as np

from . import _lib


def parse_many(structures):
    """Annotations (str) for a list of dot-bracket strings.  Raises ValueError naming the first
    structure that is unbalanced or holds characters other than ``( ) .``"""
    structures = [s if isinstance(s, str) else s.decode("ascii") for s in structures]
    lengths = np.fromiter((len(s) for s in structures), dtype=np.int64, count=len(structures))
    offsets = np.zeros(len(structures), np.int64)
    if len(structures) > 1:
        np.cumsum(lengths[:-1], out=offsets[1:])
    text = np.frombuffer("".join(structures).encode("latin-1", "replace"), dtype=np.uint8)
    out = np.zeros(max(len(text), 1), dtype=np.uint8)
    status = np.zeros(max(len(structures), 1), dtype=np.int32)
    rc = _lib.lib.rs_host_annotate_structures(text.ctypes.data if len(text) else 0, offsets.ctypes.data,
                                              lengths.ctypes.data, len(structures), out.ctypes.data,
                                              status.ctypes.data)
    if rc != 0:
        bad = int(np.nonzero(status[:len(structures)])[0][0]) if len(structures) else -1
        raise ValueError("structure %d is not a balanced dot-bracket string" % bad)
    raw = out.tobytes().decode("ascii")
    return [raw[o:o + n] for o, n in zip(offsets.tolist(), lengths.tolist())]


def parse(structure):
    """Annotation of one dot-bracket string, e.g. ``..((...))..`` -> ``EELLHHHRREE``."""
    return parse_many([structure])[0]


def parse_file(infile, outfile):
    """The reference tool's file interface (parse_secondary_structure.cpp:234-259): every line that
    contains a '.' is treated as ``<structure> [anything]`` and its annotation is written out."""
    lines = []
    with open(infile) as fh:
        for line in fh:
            if "." in line:
                lines.append(line.split()[0])
    with open(outfile, "w") as fh:
        for a in parse_many(lines):
            fh.write(a + "\n")
