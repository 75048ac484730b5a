"""Scratch: cross-kernel check of the thresholded sequence / structure scans at a size where the look-back kernels
(order_kernel, kmer_finish_kernel) run more CTAs than fit on the device at once.   python tools/big_check.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from rnascan_b200 import device as dev, synth
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 600_000_000
device = torch.device("cuda", 0)
for wl, kind, A, W, thr in (("c2", "rna", 4, 7, 6.0), ("c3", "struct", 7, 7, 7.5), ("c2", "rna", 4, 12, 5.0)):
    shard = bench.make_device_shard(n, 77, wl, device)
    st = dev.SymbolStream.__new__(dev.SymbolStream)
    st.kind, st.n, st.codes = kind, shard["n"], shard["codes"]
    st.offsets, st.lengths, st._host = shard["offsets"], shard["lengths"], None
    rng = np.random.default_rng(5 + W)
    tab = synth.pssm_table(synth.pfm_rows(W, A, rng))
    t0 = time.time()
    if A == 4:
        pos, sc = dev.scan_seq(st, tab, thr, capacity=st.n // 64)
        dense = dev.dense_seq(st, tab)
        keep = dense.double() > thr
    else:
        pos, sc = dev.scan_struct_onehot(st, tab, thr, capacity=st.n // 64)
        dense = dev.dense_struct(st, tab)
        keep = dense > thr
    want = torch.nonzero(keep).flatten()
    ok_pos = torch.equal(torch.from_numpy(pos).to(device), want)
    ok_sc = torch.equal(torch.from_numpy(sc).to(device), dense[want])
    inc = bool(np.all(np.diff(pos) > 0))
    print("%s W=%d n=%d: %d hits, positions equal %s, scores equal %s, increasing %s  (%.1f s)"
          % (kind, W, st.n, len(pos), ok_pos, ok_sc, inc, time.time() - t0), flush=True)
    assert ok_pos and ok_sc and inc
    del shard, st, dense, keep, want
    torch.cuda.empty_cache()
print("big check ok")
