"""Synthetic config-4-shaped CLI inputs as FILES: a FASTA, a directory of structure.<id>.txt averaged profiles
(pfmutil.format_pfm layout), sequence / structure PFMs and background files.

    python tools/cli_dataset.py <out_dir> [n_symbols] [n_records]     # prints the rnascan argv, one item per line
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnascan_b200 import synth  # noqa: E402


def write_pfm(path, rows, letters):
    with open(path, "w") as fh:
        fh.write("PO\t" + "\t".join(letters) + "\n")
        for i, r in enumerate(rows):
            fh.write("%d\t%s\n" % (i + 1, "\t".join(repr(float(v)) for v in r)))


def make(out, n=600_000, n_records=200, seed=4):
    os.makedirs(os.path.join(out, "profiles"), exist_ok=True)
    rng = np.random.default_rng(seed)
    lengths = synth.record_lengths(n, n_records, rng)
    codes, off = synth.rna_codes(lengths, rng)
    rows = synth.profile_rows(len(codes), rng, lengths=lengths).astype(np.float64)
    text = synth.to_text(codes, "rna").decode().split("\n")[:-1]
    with open(os.path.join(out, "seqs.fa"), "w") as fh:
        for k, r in enumerate(text):
            fh.write(">rec%d synthetic record %d\n" % (k, k))
            fh.write("\n".join(r[a:a + 60] for a in range(0, len(r), 60)) + "\n")
    for k in range(len(lengths)):
        block = rows[off[k]:off[k] + lengths[k]]
        with open(os.path.join(out, "profiles", "structure.rec%d.txt" % k), "w") as fh:
            fh.write("PO\tB\tE\tH\tL\tM\tR\tT\n")
            fh.write("".join("%d\t%s\n" % (i, "\t".join(repr(float(v)) for v in r)) for i, r in enumerate(block)))
    write_pfm(os.path.join(out, "seq.pfm"), synth.pfm_rows(7, 4, np.random.default_rng(102)) + 0.01, "ACGU")
    write_pfm(os.path.join(out, "struct.pfm"), synth.pfm_rows(7, 7, np.random.default_rng(103)) + 0.01, "BEHLMRT")
    with open(os.path.join(out, "bg_struct.txt"), "w") as fh:
        fh.write(repr({c: synth.SS_P[c] for c in "EHTBLRM"}))
    return ["-p", os.path.join(out, "seq.pfm"), "-q", os.path.join(out, "struct.pfm"), "-B",
            os.path.join(out, "bg_struct.txt"), os.path.join(out, "seqs.fa"), os.path.join(out, "profiles")]


if __name__ == "__main__":
    argv = make(sys.argv[1], int(float(sys.argv[2])) if len(sys.argv) > 2 else 600_000,
                int(sys.argv[3]) if len(sys.argv) > 3 else 200)
    print("\n".join(argv))
