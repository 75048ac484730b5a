"""Scratch: stage hand-off timeline of batched_tc_kernel.  Needs batched_tc.cu compiled with -DRS_TC_TIMELINE:
   cd rnascan_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
       -DRS_TC_TIMELINE -c batched_tc.cu -o /tmp/tl.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static \
       -o ../librnascan_b200.so $(ls *.o | grep -v batched_tc.o) /tmp/tl.o -lpthread;  python tools/tc_timeline.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, bench
from rnascan_b200 import _lib
class A: pass
args = A(); args.n_per_gpu = 125_000_000; args.warmup = 3; args.no_e2e = True; args.c5_path = "tensor"; args.serial_bg = False
bench.measure(args, "c5", 1, (0, 1, 0, torch.device("cuda", 0)), full=False)
tl = np.zeros(96 * 8, np.int64)
_lib.lib.rs_debug_tc_timeline.argtypes = [ctypes.c_void_p]
assert _lib.lib.rs_debug_tc_timeline(tl.ctypes.data) == 0
tl = tl.reshape(96, 8)
t0 = tl[0, 0]
names = ["operands ready", "acc stage free", "MMAs issued", "epi: acc full", "epi: chunks done", "conv: raw landed", "conv: converted"]
print("tile  " + "  ".join("%16s" % n for n in names))
for it in range(40, 60):
    print("%4d  " % it + "  ".join("%16d" % (tl[it, k] - t0) for k in range(7)))
d = np.diff(tl[32:96, 2])
print("period (MMAs issued, tiles 32..95): mean %.0f clk" % d.mean())
print("wait for operands   (stamp0 - prev stamp2): %.0f" % (tl[33:96, 0] - tl[32:95, 2]).mean())
print("wait for acc stage  (stamp1 - stamp0)     : %.0f" % (tl[32:96, 1] - tl[32:96, 0]).mean())
print("issue 6 MMAs+commit (stamp2 - stamp1)     : %.0f" % (tl[32:96, 2] - tl[32:96, 1]).mean())
print("MMA issued -> epilogue sees acc full (3-2): %.0f" % (tl[32:96, 3] - tl[32:96, 2]).mean())
print("epilogue chunks (4-3)                     : %.0f" % (tl[32:96, 4] - tl[32:96, 3]).mean())
print("epilogue done(t) -> acc free seen for t+2 : %.0f" % (tl[34:96, 1] - tl[32:94, 4]).mean())
print("converter: raw landed -> converted (6-5)  : %.0f" % (tl[32:96, 6] - tl[32:96, 5]).mean())
print("converted(t) -> operands ready seen (0-6) : %.0f" % (tl[32:96, 0] - tl[32:96, 6]).mean())
