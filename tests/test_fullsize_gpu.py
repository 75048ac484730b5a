"""Parity at BASELINE.json's full sizes (100-125 M symbols per GPU), where the CPU oracle is too slow
to be the checker: size-independent properties of the path, each comparing two INDEPENDENT
implementations of the same quantity or two decompositions of the same input.

* sortedness / uniqueness of hit positions;
* the k-mer decision-table scan == thresholding the dense scores of a different kernel;
* the fused AND scan == intersection of the sequence scan and the structure-only scan;
* shard invariance: scanning two overlapping halves and keeping owned starts == scanning the whole
  (the multi-GPU decomposition, shard.plan_shards);
* background counts: sum over shards == whole, and == a torch.bincount of the same bytes;
* the tensor-core batched path == the per-motif CUDA-core loop;
* host-resident rows: the quantised and the float32 filter + gather + resolve pipelines == the device-resident
  fused scan (three forms of the same input, different kernels, different precision of the filter);
* every-position structure scores: int32 thousandths == Python round() of the float64 kernel's output on a sample;
* idempotence: a second run returns bit-identical arrays;
* a bounded random sample of positions is re-scored by the CPU oracle (bit-exact).
"""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_FULL = 100_000_000          # BASELINE configs 2/3: 100 Mnt


@pytest.fixture(scope="module")
def env():
    import bench
    from rnascan_b200 import device as dev
    dev.require_cuda()
    device = torch.device("cuda", 0)
    shard = bench.make_device_shard(N_FULL, 4242, "c4", device)
    n = shard["n"]

    class Stream(object):
        pass
    st = Stream()
    st.codes, st.n, st.kind = shard["codes"], n, "rna"
    st.offsets, st.lengths = shard["offsets"], shard["lengths"]
    pf = dev.ProfileStream.from_device(shard["prof"], n)
    counts = dev.histogram(st).cpu().numpy()
    ts, tq = bench.make_tables_fn("c4")(counts)
    return {"dev": dev, "st": st, "pf": pf, "ts": ts, "tq": tq, "counts": counts, "shard": shard, "bench": bench}


def test_counts_full_size(env):
    dev, st = env["dev"], env["st"]
    want = torch.bincount(st.codes[:st.n].to(torch.int64), minlength=256)[:8].cpu().numpy()
    assert np.array_equal(env["counts"], want)
    # generic 3-plane kernel and the sum over two shards agree
    st.kind = None
    assert np.array_equal(dev.histogram(st).cpu().numpy(), want)
    st.kind = "rna"
    half = (st.n // 2) // 256 * 256

    class Part(object):
        pass
    a, b = Part(), Part()
    a.codes, a.n, a.kind = st.codes, half, "rna"
    b.codes, b.n, b.kind = st.codes[half:], st.n - half, "rna"
    assert np.array_equal(dev.histogram(a).cpu().numpy() + dev.histogram(b).cpu().numpy(), want)


def test_sequence_scan_full_size(env):
    dev, st, ts = env["dev"], env["st"], env["ts"]
    pos, sc = dev.scan_seq(st, ts, 6.0)
    assert len(pos) > 100_000 and np.all(np.diff(pos) > 0)
    # an independent kernel: dense scores (dense_w_kernel) thresholded on the device
    dense = dev.dense_seq(st, ts)
    keep = torch.nonzero(dense.double() > 6.0).flatten()
    assert np.array_equal(pos, keep.cpu().numpy())
    assert np.array_equal(sc.view(np.uint32), dense[keep].cpu().numpy().view(np.uint32))
    # idempotence
    pos2, sc2 = dev.scan_seq(st, ts, 6.0)
    assert np.array_equal(pos, pos2) and np.array_equal(sc.view(np.uint32), sc2.view(np.uint32))
    env["seq_hits"] = (pos, sc)


def test_shard_invariance_full_size(env):
    """Two ranks' view of the same stream: pieces overlap by W-1 symbols, each keeps the starts it owns."""
    dev, st, ts = env["dev"], env["st"], env["ts"]
    W = ts.shape[0]
    pos, sc = env.get("seq_hits") or dev.scan_seq(st, ts, 6.0)
    cut = (st.n // 2) // 256 * 256 + 256          # 16-byte aligned piece start

    class Part(object):
        pass
    a, b = Part(), Part()
    a.codes, a.n = st.codes, cut + W - 1          # owns starts [0, cut)
    b.codes, b.n = st.codes[cut:], st.n - cut     # owns starts [cut, n)
    pa, sa = dev.scan_seq(a, ts, 6.0)
    pb, sb = dev.scan_seq(b, ts, 6.0)
    keep = pa < cut
    got_pos = np.concatenate([pa[keep], pb + cut])
    got_sc = np.concatenate([sa[keep], sb])
    assert np.array_equal(got_pos, pos)
    assert np.array_equal(got_sc.view(np.uint32), sc.view(np.uint32))


def test_fused_scan_is_the_intersection_full_size(env):
    dev, st, pf, ts, tq = env["dev"], env["st"], env["pf"], env["ts"], env["tq"]
    thr = 2.0                                       # low enough for thousands of joint hits
    pos, sq, sb = dev.scan_fused(st, pf, ts, tq, thr)
    assert len(pos) > 1000 and np.all(np.diff(pos) > 0)
    ps, ss = dev.scan_seq(st, ts, thr)
    pq, _, sq_only = dev.scan_fused(st, pf, None, tq, thr)
    both, ia, ib = np.intersect1d(ps, pq, assume_unique=True, return_indices=True)
    assert np.array_equal(pos, both)
    assert np.array_equal(sq.view(np.uint32), ss[ia].view(np.uint32))
    assert np.array_equal(sb.view(np.uint64), sq_only[ib].view(np.uint64))
    env["fused"] = (pos, sq, sb, thr)


def test_sampled_positions_against_the_cpu_oracle(env, oracle):
    """Bit-exact re-score of hits and of random windows by the oracle on the bytes/rows they cover."""
    from rnascan_b200 import synth
    dev, st, pf, ts, tq = env["dev"], env["st"], env["pf"], env["ts"], env["tq"]
    pos, sq, sb, thr = env.get("fused") or (dev.scan_fused(st, pf, ts, tq, 2.0) + (2.0,))
    W = ts.shape[0]
    rng = np.random.default_rng(1)
    pick = rng.choice(len(pos), size=min(300, len(pos)), replace=False)
    for k in pick.tolist():
        p = int(pos[k])
        codes = st.codes[p:p + W].cpu().numpy()
        rows = pf.rows[p:p + W].cpu().numpy()
        a = oracle.seq_scores(synth.to_text(codes, "rna"), ts)
        b = oracle.profile_scores(rows, tq)
        assert a.view(np.uint32)[0] == sq[k:k + 1].view(np.uint32)[0]
        assert b.view(np.uint64)[0] == sb[k:k + 1].view(np.uint64)[0]
        assert float(a[0]) > thr and b[0] > thr
    # random windows through the dense kernels
    starts = rng.integers(0, st.n - 4096, size=40)
    dense = dev.dense_seq(st, ts)
    for s0 in starts.tolist():
        codes = st.codes[s0:s0 + 2048 + W - 1].cpu().numpy()
        want = oracle.seq_scores(synth.to_text(codes, "rna"), ts)
        got = dense[s0:s0 + 2048].cpu().numpy()
        nan = np.isnan(want)
        assert np.array_equal(nan, np.isnan(got))
        assert np.array_equal(want[~nan].view(np.uint32), got[~nan].view(np.uint32))


@pytest.mark.parametrize("thr", [6.0, 2.5])
def test_host_resident_pipelines_equal_the_device_resident_scan_full_size(env, thr):
    """100 M rows in HOST memory through device.HostProfileScanner -- 8-byte and 4-byte quantised rows (background
    counted in the filter pass; 4-byte: sequence table applied to the packed symbols on the device) and float32 rows
    -- against rs_scan_fused on the resident streams: identical positions and scores,
    identical counts."""
    dev, st, pf, tq, bench = env["dev"], env["st"], env["pf"], env["tq"], env["bench"]
    n = st.n
    tables = bench.make_tables_fn("c4")
    want = dev.scan_fused(st, pf, env["ts"], tq, thr)
    rows = np.empty((n, 7), np.float32)
    bounce = torch.empty((1 << 23, 7), dtype=torch.float32).pin_memory()
    for a in range(0, n, 1 << 23):
        b = min(n, a + (1 << 23))
        bounce[:b - a].copy_(pf.rows[a:b])
        rows[a:b] = bounce[:b - a].numpy()
    codes = st.codes[:n].cpu().numpy()
    hp = dev.HostProfile(rows)
    assert hp.make_q8(codes)
    seen = []

    def seq_fn(counts8):
        seen.append(np.array(counts8, np.int64))
        return tables(counts8)[0]

    assert hp.make_q4(codes)
    for form, src, cd in (("q8", hp.q8, None), ("q4", hp.q4, None), ("f32", rows, codes)):
        sc = dev.HostProfileScanner(n, tq.shape[0], form)
        got = sc.run(cd, src, rows, tq, seq_fn, thr, hp.absrow_max(), q8_scale=hp.q8_scale if form != "f32" else 1.0)
        assert np.array_equal(seen[-1], env["counts"]), form
        assert np.array_equal(got[0], want[0]), form
        assert np.array_equal(got[1].view(np.uint32), want[1].view(np.uint32)), form
        assert np.array_equal(got[2].view(np.uint64), want[2].view(np.uint64)), form
        assert sc.n_candidates >= len(want[0])
        assert np.all(np.diff(got[0]) > 0)
    assert len(want[0]) > 0


def test_thousandths_equal_python_round_of_the_float64_kernel_full_size(env):
    """C3 shape: 100 M one-hot structure windows; int32 thousandths vs the float64 kernel + CPython round() on a
    random sample of 200 k positions, sentinels where the float64 score is NaN."""
    import math
    from rnascan_b200 import _lib
    dev, bench = env["dev"], env["bench"]
    shard = bench.make_device_shard(N_FULL, 777, "c3", torch.device("cuda", 0))

    class Stream(object):
        pass
    st = Stream()
    st.codes, st.n, st.kind = shard["codes"], shard["n"], "struct"
    counts = dev.histogram(st).cpu().numpy()
    tq = bench.make_tables_fn("c3")(counts)[1]
    milli = dev.dense_struct_milli(st, tq)
    dense = dev.dense_struct(st, tq)
    idx = torch.randint(0, milli.numel(), (200_000,), device=milli.device)
    m, d = milli[idx].cpu().numpy(), dense[idx].cpu().numpy()
    for k, x in zip(m.tolist(), d.tolist()):
        if x != x:
            assert k == _lib.RS_MILLI_NAN
        else:
            r = round(x, 3)
            assert k == (_lib.RS_MILLI_NEG0 if (r == 0 and math.copysign(1.0, r) < 0) else int(round(r * 1000)))
    # and the NaN pattern agrees everywhere
    assert torch.equal(milli == _lib.RS_MILLI_NAN, dense != dense)


def test_batched_paths_agree_full_size(env):
    dev, st, pf, bench = env["dev"], env["st"], env["pf"], env["bench"]
    tables = bench.make_tables_fn("c5")
    ss, qs = tables(env["counts"])
    M = 64                                          # a quarter of config 5's motifs keeps the CUDA-core loop short
    tsl = [ss[m, :tables.widths[m]] for m in range(M)]
    tql = [qs[m, :tables.widths[m]] for m in range(M)]
    a = dev.scan_batched(st, pf, tsl, tql, 6.0, path=1)
    b = dev.scan_batched(st, pf, tsl, tql, 6.0, path=2)
    assert len(a[1]) > 100
    for x, y in zip(a, b):
        assert np.array_equal(np.asarray(x).view(np.uint8), np.asarray(y).view(np.uint8))
    motif, pos = a[0], a[1]
    key = motif.astype(np.int64) * (1 << 40) + pos
    assert np.all(np.diff(key) > 0)                 # grouped by motif, sorted by position


@pytest.mark.parametrize("wl,kind,A,W,thr", [("c2", "rna", 4, 7, 6.0), ("c3", "struct", 7, 7, 7.5), ("c2", "rna", 4, 12, 5.0)])
def test_look_back_kernels_beyond_one_resident_wave(wl, kind, A, W, thr):
    """300 M symbols: the single-pass finish kernel runs 327 CTAs of 1024 threads (148 fit on the device at once), so
    its ticket + look-back chain is exercised with CTAs that start only after earlier ones have retired.
    Thresholded scan == thresholding the dense scores of a different kernel: same positions, same scores, sorted."""
    import bench
    from rnascan_b200 import device as dev, synth
    device = torch.device("cuda", 0)
    torch.cuda.empty_cache()
    shard = bench.make_device_shard(300_000_000, 77, wl, device)

    class Stream(object):
        pass
    st = Stream()
    st.codes, st.n, st.kind = shard["codes"], shard["n"], kind
    st.offsets, st.lengths = shard["offsets"], shard["lengths"]
    tab = synth.pssm_table(synth.pfm_rows(W, A, np.random.default_rng(5 + W)))
    if A == 4:
        pos, sc = dev.scan_seq(st, tab, thr, capacity=st.n // 64)
        dense = dev.dense_seq(st, tab)
        want = torch.nonzero(dense.double() > thr).flatten()
    else:
        pos, sc = dev.scan_struct_onehot(st, tab, thr, capacity=st.n // 64)
        dense = dev.dense_struct(st, tab)
        want = torch.nonzero(dense > thr).flatten()
    assert len(pos) > 100_000
    assert torch.equal(torch.from_numpy(pos).to(device), want)
    assert torch.equal(torch.from_numpy(sc).to(device), dense[want])
    assert np.all(np.diff(pos) > 0)
    del shard, dense, want
    torch.cuda.empty_cache()
