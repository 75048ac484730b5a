import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rnascan_b200 import device, synth, _lib
from rnascan_b200.device import lib
rng = np.random.default_rng(855)
lengths = synth.record_lengths(150_000, 100, rng)
codes, off = synth.rna_codes(lengths, rng, n_frac=0.005)
rows = synth.profile_rows(len(codes), rng, lengths=lengths)
st, pf = device.SymbolStream(codes, off, lengths), device.ProfileStream(rows)
bg = [synth.SS_P[c] for c in "BEHLMRT"]
for M in (40, 256, 257, 300):
    widths = rng.integers(7, 13, size=M)
    tq = [synth.pssm_table(synth.pfm_rows(int(w), 7, rng), background=bg) for w in widths]
    ts = [synth.pssm_table(synth.pfm_rows(int(w), 4, rng)) for w in widths]
    out = {}
    for path in (1, 2):
        try:
            m, p, sq, sc, b = device.scan_batched(st, pf, ts, tq, 0.5, capacity=1 << 16, path=path)
            out[path] = (m, p, sq, sc, b)
            print(M, "path", path, "hits", len(p), "took", lib.rs_last_batched_path(), flush=True)
        except Exception as e:
            print(M, "path", path, "ERR", e, flush=True)
    if 1 in out and 2 in out:
        a, c = out[1], out[2]
        same = all(np.array_equal(x, y) for x, y in zip(a, c))
        print("  identical:", same)
        if not same:
            print("  bases diff at", np.nonzero(a[4] != c[4])[0][:10], len(a[1]), len(c[1]))
            k = min(len(a[1]), len(c[1]))
            d = np.nonzero((a[1][:k] != c[1][:k]) | (a[0][:k] != c[0][:k]))[0]
            print("  first diffs", d[:5], a[0][d[:5]], a[1][d[:5]], c[0][d[:5]], c[1][d[:5]])
