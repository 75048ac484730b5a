"""Scratch: where does scan_main(<directory with a pack>) spend its time?  (GPU box)
    python tools/api_profile.py [rows]"""
import argparse, cProfile, io, os, pstats, shutil, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rnascan_b200 import synth, device as dev, pack, rnascan as ms
from rnascan_b200.BioAddons.Alphabet import ContextualSecondaryStructure
from rnascan_b200.BioAddons.motifs import matrix

m = int(float(sys.argv[1])) if len(sys.argv) > 1 else 32_000_000
rng = np.random.default_rng(4)
lengths = synth.record_lengths(m, max(1, m // 3334), rng)
off, total = synth.layout(lengths)
rows = rng.dirichlet(0.3 * np.ones(7), size=total)
rows[off + lengths] = 0.0
hp = dev.HostProfile(rows)
sep = np.zeros(total, np.uint8); sep[off + lengths] = 0xFF
assert hp.make_q8(sep) and hp.make_q4(sep)
tmp = tempfile.mkdtemp(prefix="apiprof_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
pack.write(tmp, None, rows, lengths, hp.stats(), hp.q8, hp.q8_scale, q4=hp.q4, names=["structure.rec%d.txt" % i for i in range(len(lengths))])
import bench
tq = bench.make_tables_fn("c4")(np.zeros(8, np.int64))[1]              # the bench's structure table
alphabet = ContextualSecondaryStructure()
pssm = {"m": matrix.ExtendedPositionSpecificScoringMatrix(alphabet, {l: tq[:, "BEHLMRT".index(l)].tolist() for l in alphabet.letters})}
ns = argparse.Namespace(minscore=float(sys.argv[2]) if len(sys.argv) > 2 else 6.0, debug=False, pack=False)
for rep in range(5):
    t0 = time.perf_counter(); f = ms.scan_main(tmp, pssm, alphabet, None, ns); print("call %d: %.1f ms, %d hits" % (rep, (time.perf_counter() - t0) * 1e3, len(f)), flush=True)
pk = pack.read(tmp); print("page-locked sections:", {k: (v is not None) for k, v in pk._locked.items()}, "asked", pk._asked)
prof = cProfile.Profile(); prof.enable(); ms.scan_main(tmp, pssm, alphabet, None, ns); prof.disable()
s = io.StringIO(); pstats.Stats(prof, stream=s).sort_stats("cumulative").print_stats(40)
print("\n".join(l[:160] for l in s.getvalue().splitlines()[4:40]))
s = io.StringIO(); pstats.Stats(prof, stream=s).sort_stats("tottime").print_stats(25)
print("\n".join(l[:160] for l in s.getvalue().splitlines()[4:40]))
shutil.rmtree(tmp, ignore_errors=True)
