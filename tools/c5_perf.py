"""Scratch: time the tensor-core batched kernel alone at several thresholds (candidate rates)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from rnascan_b200 import device as dev, _lib
from rnascan_b200.device import lib, check, _ptr
n_target = int(float(sys.argv[1])) if len(sys.argv) > 1 else 32_000_000
device = torch.device("cuda", 0)
shard = bench.make_device_shard(n_target, 4000, "c5", device)
n, codes, prof = shard["n"], shard["codes"], shard["prof"]
tables = bench.make_tables_fn("c5")
counts = torch.zeros(8, dtype=torch.int64, device=device)
check(lib.rs_hist(_ptr(codes), n, _ptr(counts), 0)); torch.cuda.synchronize()
ss, qs = tables(counts.cpu().numpy())
M, stride = qs.shape[0], qs.shape[1]
widths = tables.widths
cap = max(1 << 16, n // 64)
hb = dev.HitBuffers(n, cap, device)
wb = int(lib.rs_scan_batched_workspace_bytes(n, M, stride, cap)); work = torch.empty(wb, dtype=torch.uint8, device=device)
motif = torch.empty(cap, dtype=torch.int32, device=device)
c2 = torch.zeros(2 * M, dtype=torch.int64, device=device); bases = torch.zeros(M + 1, dtype=torch.int64, device=device)
absmax = dev.ProfileStream.from_device(prof, n).absrow_max()
check(lib.rs_set_batched_path(2))
import ctypes
lib.rs_debug_set_tc_seqmask.argtypes = [ctypes.c_int]
for mask_on, thr in ((0, 6.0), (1, 6.0), (1, 4.0), (1, 2.0), (1, 50.0)):
    lib.rs_debug_set_tc_seqmask(mask_on)
    for rep in range(3):
        check(lib.rs_prof_begin(4))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.rs_scan_batched(_ptr(codes), _ptr(prof), _lib.RS_F32, n, M, widths.ctypes.data, ss.ctypes.data,
                                  qs.ctypes.data, stride, thr, absmax, _lib.RS_MODE_AND, cap, _ptr(motif), _ptr(hb.pos),
                                  _ptr(hb.seq), _ptr(hb.struct), _ptr(c2), _ptr(bases), _ptr(work), wb, 0))
        e1.record(); torch.cuda.synchronize()
        kms = np.zeros(4, np.float32); nrec = np.zeros(1, np.int32)
        check(lib.rs_prof_end(kms.ctypes.data, 4, nrec.ctypes.data))
    resc = int(c2[1::2].sum().item()); hits = int(bases[-1].item())
    print("mask %d  thr %5.1f  total %.2f ms  tc kernel %.2f ms  (%.1f Gpos/s kernel)  candidates %d  hits %d" %
          (mask_on, thr, e0.elapsed_time(e1), kms[0], n / kms[0] / 1e6, resc, hits), flush=True)
