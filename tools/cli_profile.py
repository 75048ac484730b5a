"""Scratch: where does the CLI's wall time go?  cProfile of `rnascan -p W7.pfm seqs.fa` (config-2 shape, m = 6)
and of the combined mode on a directory of averaged profiles.   python tools/cli_profile.py [n_symbols]"""
import contextlib, cProfile, io, os, pstats, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rnascan_b200 import synth, rnascan as ms
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(2)
lengths = synth.record_lengths(n, max(1, n // 3334), rng)
d = tempfile.mkdtemp()
codes, _ = synth.rna_codes(lengths, rng)
text = synth.to_text(codes, "rna").decode().split("\n")[:-1]
fa = os.path.join(d, "seq.fa")
with open(fa, "w") as fh:
    for k, r in enumerate(text):
        fh.write(">rec%d synthetic record %d\n" % (k, k))
        fh.write("\n".join(r[a:a + 60] for a in range(0, len(r), 60)) + "\n")
pfm = os.path.join(d, "w7.pfm")
rows = synth.pfm_rows(7, 4, np.random.default_rng(102))
with open(pfm, "w") as fh:
    fh.write("PO\tA\tC\tG\tU\n")
    for i, r in enumerate(rows):
        fh.write("%d\t%s\n" % (i + 1, "\t".join(repr(float(v)) for v in r)))
for name, argv in (("C2 shape: -p W=7 -C 0.01 (m = 6)", ["-p", pfm, "-C", "0.01", fa]),):
    for rep in range(2):
        ms._BATCH_CACHE.clear()
        out = open(os.path.join(d, "hits.tab"), "w")
        prof = cProfile.Profile()
        t0 = time.time()
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
            prof.enable(); ms.main(argv); prof.disable()
        out.close()
        dt = time.time() - t0
        rows_out = sum(1 for _ in open(os.path.join(d, "hits.tab"))) - 1
        print("%-34s run %d: %6.2f s  %d symbols  %d rows  (%.1f Mnt/s)" % (name, rep, dt, n, rows_out, n / dt / 1e6), flush=True)
    s = io.StringIO()
    pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(28)
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[4:45]))
