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


# ----------------------------------------------------------------------------- averaged profiles
PROFILE_ALPHABET = ["B", "E", "H", "L", "M", "R", "T"]


def struct_pfm_from_aligned(sequences):
    """Per-column letter counts of equal-length annotated strings with '-' for gaps
    (average_structure.py:28-42): ``{letter: [count per column]}`` for B,E,H,L,M,R,T.  A
    character outside the alphabet raises KeyError like the reference's dict lookup."""
    length = len(sequences[0])
    if any(len(s) != length for s in sequences):
        raise IndexError("list index out of range")          # what the reference's counts[char][index] raises
    text = np.frombuffer("".join(sequences).encode("latin-1", "replace"), dtype=np.uint8)
    grid = text.reshape(len(sequences), length) if length else np.zeros((len(sequences), 0), np.uint8)
    known = np.isin(grid, np.frombuffer(("".join(PROFILE_ALPHABET) + "-").encode(), dtype=np.uint8))
    if not known.all():
        r, c = np.argwhere(~known)[0]
        raise KeyError(chr(int(grid[r, c])))
    return {letter: (grid == ord(letter)).sum(axis=0).astype(np.int64).tolist() for letter in PROFILE_ALPHABET}


def profile_from_aligned(sequences):
    """norm_pfm(struct_pfm_from_aligned(...)) (average_structure.py:95-99): the L x 7 averaged
    structure profile the scan consumes, as ``{letter: [fraction per column]}``."""
    from . import pfmutil
    return pfmutil.norm_pfm(struct_pfm_from_aligned(sequences))


def profile_from_fragments(length, fragments):
    """Averaged profile of a sequence of `length` nt from folded fragments: `fragments` is a list of
    ``(start, dot_bracket)`` (start may be negative as in average_structure.py:47-58: the fragment then
    begins at 0).  Annotates every centroid (C++), aligns the annotations with '-' gaps
    (average_structure.py:89-93) and averages the columns."""
    starts = [max(0, int(s)) for s, _ in fragments]
    ann = parse_many([d for _, d in fragments])
    aligned = []
    for s, a in zip(starts, ann):
        aligned.append(("-" * s + a + "-" * max(0, length - s - len(a)))[:max(length, s + len(a))])
    return profile_from_aligned(aligned)
