"""Scratch timing of each kernel on device-resident synthetic data (not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rnascan_b200 import device, synth, _lib
from rnascan_b200.device import lib, check, _ptr, _stream

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 64_000_000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 7
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(4)
npad = device.padded_count(n)
codes = torch.randint(0, 4, (npad,), device=dev, dtype=torch.uint8, generator=g)
sep = torch.randint(0, n, (n // 3300,), device=dev, generator=g)
codes[sep] = 0xFF
codes[n:] = 0xFF
scodes = torch.randint(0, 7, (npad,), device=dev, dtype=torch.uint8, generator=g)
scodes[sep] = 0xFF; scodes[n:] = 0xFF
gam = torch._standard_gamma(torch.full((npad, 7), 0.2, device=dev), generator=g)
prof = (gam / gam.sum(1, keepdim=True).clamp_min(1e-30)).float().contiguous()
del gam
rng = np.random.default_rng(102)
ts = synth.pssm_table(synth.pfm_rows(W, 4, rng))
tq = synth.pssm_table(synth.pfm_rows(W, 7, rng))

def timeit(name, fn, bytes_per_pos, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-28s %8.3f ms  %8.1f Gpos/s  %8.1f GB/s (%.1f%% of 6547)" % (
        name, ms, n / ms / 1e6, n * bytes_per_pos / ms / 1e6, n * bytes_per_pos / ms / 1e6 / 65.475), flush=True)

class S: pass
st = S(); st.codes = codes; st.n = n
ss = S(); ss.codes = scodes; ss.n = n
pf = device.ProfileStream.from_device(prof, n)
counts = torch.zeros(8, dtype=torch.int64, device=dev)
timeit("hist (3 planes)", lambda: check(lib.rs_hist(_ptr(codes), n, _ptr(counts), _stream())), 1)
timeit("hist_rna (2 planes)", lambda: check(lib.rs_hist_rna(_ptr(codes), n, _ptr(counts), _stream())), 1)
hb = device.HitBuffers(n, n // 64, dev)
outf = torch.empty(n, dtype=torch.float32, device=dev)
outd = torch.empty(n, dtype=torch.float64, device=dev)
for thr in (6.0, 2.0):
    def f():
        check(lib.rs_scan_seq(_ptr(codes), n, ts.ctypes.data, W, thr, hb.capacity, _ptr(hb.pos), _ptr(hb.seq),
                              _ptr(hb.counters), _ptr(hb.work), hb.work_bytes, _stream()))
    timeit("scan_seq thr=%g" % thr, f, 1)
    print("   hits", hb.counters.cpu().tolist())
ss_hb = device.HitBuffers(n, n // 16, dev, want_seq=False)
for thr in (3.0, 0.0):
    def g():
        check(lib.rs_scan_struct_onehot(_ptr(scodes), n, tq.ctypes.data, W, thr, ss_hb.capacity, _ptr(ss_hb.pos), _ptr(ss_hb.struct), _ptr(ss_hb.counters), _ptr(ss_hb.work), ss_hb.work_bytes, _stream()))
    timeit("scan_struct thr=%g" % thr, g, 1)
    print("   hits", ss_hb.counters.cpu().tolist())
timeit("dense_seq", lambda: check(lib.rs_scores_dense_seq(_ptr(codes), n, ts.ctypes.data, W, _ptr(outf), _stream())), 5)
timeit("dense_struct(f64 out)", lambda: check(lib.rs_scores_dense_struct(_ptr(scodes), n, tq.ctypes.data, W, _ptr(outd), _stream())), 9)
absmax = pf.absrow_max()
print("absrow_max", absmax)
for mode, name in ((_lib.RS_MODE_AND, "fused AND"), (_lib.RS_MODE_STRUCT, "fused STRUCT")):
    for thr in (6.0, 2.0):
        def f():
            check(lib.rs_scan_fused(_ptr(codes), _ptr(prof), _lib.RS_F32, n, ts.ctypes.data, tq.ctypes.data, W, thr,
                                    absmax, mode, hb.capacity, _ptr(hb.pos), _ptr(hb.seq), _ptr(hb.struct),
                                    _ptr(hb.counters), _ptr(hb.work), hb.work_bytes, _stream()))
        timeit("%s thr=%g" % (name, thr), f, 29)
        print("   hits/rescored", hb.counters.cpu().tolist())
timeit("dense_profile(f64 out)", lambda: check(lib.rs_scores_dense_profile(_ptr(prof), _lib.RS_F32, n, _ptr(codes), tq.ctypes.data, W, _ptr(outd), _stream())), 37, reps=3)
