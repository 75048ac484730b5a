"""Parity of the overlapped background pipelines under a REAL multi-rank all-reduce (NCCL):
    torchrun --nproc-per-node 2 tools/ranks_check_bg.py
Every rank scans its own synthetic shard with device.BackgroundFusedScan / BackgroundOneHotScan (collective on the side
stream, one SM reserved, shard-local provisional start) and compares with the serial path: all-reduce first, host tables
from the GLOBAL counts, one-call scans.  Also checks that every rank ends up with the same global counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from rnascan_b200 import device as dev, _lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
n_target = 40_000_000
ok = True
total_hits = 0
for wl in ("c4", "c2"):
    shard = bench.make_device_shard(n_target, 9000 + rank, wl, device)
    n, codes, prof = shard["n"], shard["codes"], shard["prof"]
    tables = bench.make_tables_fn(wl)
    # serial path: global counts first
    counts = torch.zeros(8, dtype=torch.int64, device=device)
    dev.check(dev.lib.rs_hist_rna(codes.data_ptr(), n, counts.data_ptr(), torch.cuda.current_stream().cuda_stream))
    dist.all_reduce(counts)
    gcounts = counts.cpu().numpy()
    ts, tq = tables(gcounts)
    st = dev.SymbolStream.__new__(dev.SymbolStream)
    st.kind, st.n, st.codes, st.offsets, st.lengths, st._host = "rna", n, codes, shard["offsets"], shard["lengths"], None
    for thr in ((6.0, 2.0) if wl == "c4" else (6.0,)):
        if wl == "c4":
            pf = dev.ProfileStream.from_device(prof, n)
            want = dev.scan_fused(st, pf, ts, tq, thr)
            job = dev.BackgroundFusedScan(n, device, capacity=max(1 << 16, n // 16))
            job.launch(codes, prof, _lib.RS_F32, bench.W_MOTIF, tq, lambda c: tables(c)[0], thr, pf.absrow_max(),
                       all_reduce=dist.all_reduce)
            got = job.results()
            same = all(np.array_equal(a.view(np.uint8), b.view(np.uint8)) for a, b in zip(got, want))
            extra = ""
        else:
            want = dev.scan_seq(st, ts, thr, capacity=n // 64)
            job = dev.BackgroundOneHotScan(n, "rna", device, capacity=n // 64)
            job.launch(codes, tables.seq_prob, lambda c: tables(c)[0], thr, all_reduce=dist.all_reduce)
            pos, sc, rejected = job.results()
            same = np.array_equal(pos, want[0]) and np.array_equal(sc.view(np.uint32), want[1].view(np.uint32))
            extra = " (%d provisional candidates rejected, rescanned=%s)" % (rejected, job.rescanned)
        same = same and np.array_equal(job.counts_host.numpy()[:8], gcounts)
        print("rank %d %s thr %.1f: %d hits, identical to the serial path: %s%s" % (rank, wl, thr, len(want[0]), same, extra), flush=True)
        ok = ok and same
        total_hits += len(want[0])
    del shard, codes, prof
    torch.cuda.empty_cache()
flag = torch.tensor([1 if (ok and total_hits > 1000) else 0], device=device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("ranks check ok" if int(flag.item()) else "RANKS CHECK FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) else 1)
