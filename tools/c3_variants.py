"""Scratch (GPU box): dense structure kernel variants -- float64 / thousandths output, prefix table on / off."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from rnascan_b200 import device as dev
from rnascan_b200.device import lib, _ptr, check
n = 125_000_000
shard = bench.make_device_shard(n, 4000, "c3", torch.device("cuda"))
codes, n = shard["codes"], shard["n"]
tq = bench.make_tables_fn("c3")(np.array([1, 2, 3, 4, 5, 6, 7, 0]) * 1000)[1]
out64 = torch.empty(n, dtype=torch.float64, device="cuda")
out32 = torch.empty(n, dtype=torch.int32, device="cuda")
s = torch.cuda.current_stream().cuda_stream
for W in (7, 12):
    t = np.ascontiguousarray(np.vstack([tq] * 2)[:W])
    for pre in (0,):
        for name, fn, out in (("f64", lib.rs_scores_dense_struct, out64), ("milli", lib.rs_scores_dense_struct_milli, out32)):
            for _ in range(3):
                check(fn(_ptr(codes), n, t.ctypes.data, W, _ptr(out), s))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                check(fn(_ptr(codes), n, t.ctypes.data, W, _ptr(out), s))
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print("W=%d prefix=%d out=%-5s %.3f ms  %.0f Gpos/s" % (W, pre, name, ms, n / ms / 1e6), flush=True)
