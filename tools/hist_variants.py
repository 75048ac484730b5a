"""Scratch: time rs_hist_rna cold (L2 flushed, clean) for the variant selected by RS_X_HIST_* (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from rnascan_b200 import device as dev
from rnascan_b200.device import lib, _ptr, check
n = 125_000_000
codes = torch.randint(0, 4, (dev.padded_count(n),), dtype=torch.uint8, device="cuda")
counts = torch.zeros(8, dtype=torch.int64, device="cuda")
scrub = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
ts = []
for it in range(40):
    scrub.fill_(1); scrub[: 256 << 20].view(torch.int64).sum()
    counts.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); check(lib.rs_hist_rna(_ptr(codes), n, _ptr(counts), st)); b.record()
    torch.cuda.synchronize()
    if it >= 5: ts.append(a.elapsed_time(b) * 1e3)
want = torch.bincount(codes[:n].long(), minlength=4).cpu().numpy()
assert np.array_equal(counts.cpu().numpy()[:4], want)
print("per_sm=%s unroll=%s: median %.1f us, min %.1f us" % (os.environ.get("RS_X_HIST_PER_SM", "max"), os.environ.get("RS_X_HIST_UNROLL", "4"), np.median(ts), min(ts)))
