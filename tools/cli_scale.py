"""Inputs for timing the CLI itself at scale (single process and under torchrun):
    python tools/cli_scale.py <dir> [fasta_symbols] [profile_rows]
writes <dir>/seqs.fa (one line per record), <dir>/profiles/rnascan_b200.pack (a pack standing for the structure.<id>.txt
files of the same records), seq.pfm, struct.pfm, bg_struct.txt, and prints the three argument lists used by
tools/gpu_cli_scale.sh:  RNA (sequence scan, computed background), SS (averaged profiles from the pack), RNASS."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rnascan_b200 import synth, device as dev, pack  # noqa: E402
from cli_dataset import write_pfm  # noqa: E402


def main(out, n_fasta, n_rows):
    os.makedirs(os.path.join(out, "profiles"), exist_ok=True)
    rng = np.random.default_rng(7)
    t0 = time.time()
    lengths = synth.record_lengths(n_fasta, max(1, n_fasta // 3334), rng)
    codes, off = synth.rna_codes(lengths, rng)
    text = synth.to_text(codes, "rna")                       # records separated by newlines
    with open(os.path.join(out, "seqs.fa"), "wb") as fh:
        view = memoryview(text)
        for k in range(len(lengths)):
            fh.write(b">rec%d synthetic record %d\n" % (k, k))
            fh.write(view[off[k]:off[k] + lengths[k] + 1])
    # profiles for the first records, up to n_rows rows, as a pack only
    k = max(1, int(np.searchsorted(np.cumsum(lengths + 1), n_rows, side="right")))
    m = int(off[k - 1] + lengths[k - 1] + 1)
    rows = np.empty((m, 7), np.float64)
    step = 1 << 22
    for a in range(0, m, step):
        b = min(m, a + step)
        g = rng.standard_gamma(0.2, size=(b - a, 7))
        rows[a:b] = g / np.maximum(g.sum(axis=1, keepdims=True), 1e-300)
    rows[off[:k] + lengths[:k]] = 0.0
    hp = dev.HostProfile(rows)
    sep = np.zeros(m, np.uint8)
    sep[off[:k] + lengths[:k]] = 0xFF
    assert hp.make_q8(sep) and hp.make_q4(sep)
    pack.write(os.path.join(out, "profiles"), None, rows, lengths[:k], hp.stats(), hp.q8, hp.q8_scale, q4=hp.q4,
               names=["structure.rec%d.txt" % i for i in range(k)])
    write_pfm(os.path.join(out, "seq.pfm"), synth.pfm_rows(7, 4, np.random.default_rng(102)) + 0.01, "ACGU")
    write_pfm(os.path.join(out, "struct.pfm"), synth.pfm_rows(7, 7, np.random.default_rng(103)) + 0.01, "BEHLMRT")
    with open(os.path.join(out, "bg_struct.txt"), "w") as fh:
        fh.write(repr({c: synth.SS_P[c] for c in "EHTBLRM"}))
    sys.stderr.write("generated %d symbols / %d records, %d profile rows / %d profiles in %.1f s\n"
                     % (len(codes), len(lengths), m, k, time.time() - t0))


if __name__ == "__main__":
    main(sys.argv[1], int(float(sys.argv[2])) if len(sys.argv) > 2 else 200_000_000,
         int(float(sys.argv[3])) if len(sys.argv) > 3 else 50_000_000)
