"""Scratch: wall time of the CLI on synthetic files of config-2/3 shapes (sequence scan at m=6, structure at -inf)."""
import contextlib, io, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rnascan_b200 import synth, rnascan as ms
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(2)
lengths = synth.record_lengths(n, max(1, n // 3334), rng)
d = tempfile.mkdtemp()
def write(path, codes, kind):
    text = synth.to_text(codes, kind).decode().split("\n")[:-1]
    with open(path, "w") as fh:
        for k, r in enumerate(text):
            fh.write(">rec%d synthetic record %d\n" % (k, k))
            fh.write("\n".join(r[a:a + 60] for a in range(0, len(r), 60)) + "\n")
codes, _ = synth.rna_codes(lengths, rng); write(os.path.join(d, "seq.fa"), codes, "rna")
codes, _ = synth.struct_codes(lengths, rng); write(os.path.join(d, "ss.fa"), codes, "struct")
inp = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "inputs")
for name, argv in (("C2-like  -p W=4 m=2.0", ["-p", os.path.join(inp, "test_seq_pfm.txt"), "-m", "2.0", os.path.join(d, "seq.fa")]),
                   ("C3-like  -q W=4 -m -inf", ["-q", os.path.join(inp, "test_struct_pfm.txt"), "-C", "0.01", "-m", " -inf", os.path.join(d, "ss.fa")])):
    ms._BATCH_CACHE.clear()
    out = open(os.path.join(d, "hits.tab"), "w")
    t0 = time.time()
    with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
        ms.main(argv)
    out.close()
    dt = time.time() - t0
    size = os.path.getsize(os.path.join(d, "hits.tab"))
    rows = sum(1 for _ in open(os.path.join(d, "hits.tab"))) - 1
    print("%-26s %6.2f s  %d symbols  %d rows  %.0f MB out  (%.1f Mnt/s, %.2f M rows/s)" %
          (name, dt, n, rows, size / 1e6, n / dt / 1e6, rows / dt / 1e6), flush=True)
