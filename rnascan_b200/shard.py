"""Multi-GPU sharding of the scan: one process per GPU (torchrun), records partitioned by
total length, ONE collective -- the all-reduce of the integer background counts.

The reference's only parallelism is ``multiprocessing.Pool`` over records
(rnascan.py:388-395); a window depends on its own W symbols only, so ranks never exchange
sequence data.  A record longer than the balance quantum is split with a (W-1)-symbol
overlap; each piece OWNS a half-open range of window starts, so no hit is lost or reported
twice, and background counts are taken over the owned symbols only.
"""
import os

import numpy as np


def world():
    """(rank, world_size) of this process: torch.distributed if initialised, else the
    torchrun environment, else (0, 1)."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    """Join the torchrun rendezvous if there is one (idempotent).  NCCL on a GPU box (one GPU
    per rank, LOCAL_RANK picks it), gloo otherwise."""
    import torch
    import torch.distributed as dist
    size = int(os.environ.get("WORLD_SIZE", 1))
    if size <= 1 or dist.is_initialized():
        return world()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", 0))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group(backend)
    global _JOINED_HERE
    _JOINED_HERE = True
    return world()


_JOINED_HERE = False


def initialized():
    try:
        import torch.distributed as dist
        return bool(dist.is_available() and dist.is_initialized())
    except ImportError:
        return False


def finalize():
    """Leave the process group that init() created."""
    global _JOINED_HERE
    if not _JOINED_HERE:
        return
    _JOINED_HERE = False
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            dist.barrier()
            dist.destroy_process_group()
    except Exception:
        pass


def allreduce_counts(counts):
    """Element-wise sum of an int64 count vector over all ranks (exact, order-independent:
    bit-identical to counting the whole input in one process).  Identity without a group."""
    counts = np.asarray(counts, dtype=np.int64)
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return counts
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return counts
    t = torch.from_numpy(counts.copy())
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def gather_objects(obj, dst=0):
    """List of every rank's `obj` on rank `dst` (None elsewhere); [obj] without a group."""
    try:
        import torch.distributed as dist
    except ImportError:
        return [obj]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out if dist.get_rank() == dst else None


def gather_arrays(arrays, dst=0):
    """Every rank's list of numpy arrays on rank `dst` as [[rank 0's arrays], [rank 1's], ...] (None elsewhere).
    The payload travels as ONE byte tensor per rank through all_gather (padded to the longest rank): no
    pickling of the data, only the shapes/dtypes go through the object channel."""
    arrays = [np.ascontiguousarray(a) for a in arrays]
    try:
        import torch
        import torch.distributed as dist
    except ImportError:
        return [arrays]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [arrays]
    size, rank = dist.get_world_size(), dist.get_rank()
    meta = [(a.shape, a.dtype.str) for a in arrays]
    nbytes = int(sum(a.nbytes for a in arrays))
    metas = [None] * size
    dist.all_gather_object(metas, (meta, nbytes))
    longest = max(1, max(m[1] for m in metas))
    on_gpu = dist.get_backend() == "nccl"
    buf = np.zeros(longest, np.uint8)
    off = 0
    for a in arrays:
        buf[off:off + a.nbytes] = a.view(np.uint8).reshape(-1)
        off += a.nbytes
    mine = torch.from_numpy(buf)
    if on_gpu:
        mine = mine.cuda()
    parts = [torch.empty_like(mine) for _ in range(size)]
    dist.all_gather(parts, mine)
    if rank != dst:
        return None
    out = []
    for (meta_r, _), part in zip(metas, parts):
        raw = part.cpu().numpy()
        off, items = 0, []
        for shape, dt in meta_r:
            n = int(np.prod(shape)) * np.dtype(dt).itemsize
            items.append(raw[off:off + n].view(np.dtype(dt)).reshape(shape).copy())
            off += n
        out.append(items)
    return out


def broadcast_object(obj, src=0):
    """Rank `src`'s `obj` on every rank (identity without a group)."""
    try:
        import torch.distributed as dist
    except ImportError:
        return obj
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return obj
    box = [obj if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def plan_shards(lengths, n_ranks, W):
    """Partition records over ranks by total length.

    Returns, per rank, a list of pieces ``(record, start, stop, own_stop)``: the rank holds
    symbols [start, stop) of that record, owns the window starts in [start, own_stop) and
    counts the symbols [start, own_stop') for the background, where own_stop' = own_stop for
    a split piece and stop for the last piece of a record.  Pieces are in record order, so
    concatenating the ranks' hit lists in rank order gives the reference's output order.
    """
    lengths = [int(v) for v in lengths]
    n_ranks = max(1, int(n_ranks))
    W = max(1, int(W))
    total = sum(lengths)
    plan = [[] for _ in range(n_ranks)]
    if total == 0:
        for r, _ in enumerate(lengths):
            plan[0].append((r, 0, 0, 0))
        return plan
    quantum = -(-total // n_ranks)
    rank, load = 0, 0
    for r, L in enumerate(lengths):
        a = 0
        while True:
            room = quantum - load
            last_rank = rank == n_ranks - 1
            rest = L - a
            if last_rank or rest <= room or rest - room < W or room < W:
                # the whole remainder goes to this rank (never leave a tail shorter than a window)
                plan[rank].append((r, a, L, L))
                load += rest
                if load >= quantum and not last_rank:
                    rank, load = rank + 1, 0
                break
            b = a + room                                  # split: next rank starts owning at b
            plan[rank].append((r, a, min(L, b + W - 1), b))
            rank, load = rank + 1, 0
            a = b
    return plan


def owned_symbols(piece):
    """Number of leading symbols of a piece that count towards the background."""
    _, start, stop, own_stop = piece
    return (stop if own_stop >= stop else own_stop) - start
