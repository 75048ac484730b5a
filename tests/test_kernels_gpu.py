"""GPU parity: every C-ABI scan entry point against the CPU oracle on the same seeded inputs.

Bit-exact for positions, counts and one-hot scores (integer / sequential-double work);
the averaged-profile scores are compared bit-exactly too, because the device re-scores
every reported window in fp64 in the reference's operation order (rnascan.py:302-307).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def dev():
    from rnascan_b200 import device
    device.require_cuda()
    return device


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint32 if a.dtype == np.float32 else np.uint64)


def assert_same_float(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape and a.dtype == b.dtype
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b)
    assert np.array_equal(_bits(a[~nan_a]), _bits(b[~nan_b]))


def make_stream(dev, total, n_records, seed, kind="rna", n_frac=0.01):
    from rnascan_b200 import synth
    rng = np.random.default_rng(seed)
    lengths = synth.record_lengths(total, n_records, rng)
    if kind == "rna":
        codes, offsets = synth.rna_codes(lengths, rng, n_frac=n_frac)
    else:
        codes, offsets = synth.struct_codes(lengths, rng)
        # sprinkle lower-case (scored, not counted) and unknown symbols
        k = rng.integers(0, len(codes), size=max(1, len(codes) // 200))
        keep = codes[k] != 0xFF
        codes[k[keep]] |= 8
        k = rng.integers(0, len(codes), size=max(1, len(codes) // 500))
        keep = codes[k] != 0xFF
        codes[k[keep]] = 0x0F
    return dev.SymbolStream(codes, offsets, lengths), codes, lengths


def pick_threshold(scores, q):
    """A threshold given as a quantile of the finite oracle scores ('q0.99'), or a number."""
    if isinstance(q, str):
        s = np.asarray(scores, np.float64)
        s = s[np.isfinite(s) & (s > -1e300)]
        assert len(s) > 100, "test input has (almost) no finite scores"
        return float(np.quantile(s, float(q[1:])))
    return float(q)


def window_has_sep(codes, W):
    sep = (codes == 0xFF).astype(np.int32)
    c = np.concatenate([[0], np.cumsum(sep)])
    n = len(codes) - W + 1
    return (c[W:W + n] - c[:n]) > 0


# ----------------------------------------------------------------------------- histogram
@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 4095, 100003, 3_000_000])
def test_hist_exact(dev, n):
    rng = np.random.default_rng(n + 1)
    codes = rng.choice(np.array([0, 1, 2, 3, 4, 5, 6, 8, 9, 0x0C, 0x0F, 0xFF], np.uint8), size=n)
    st = dev.SymbolStream(codes)
    got = dev.histogram(st).cpu().numpy()
    want = np.array([(codes == k).sum() for k in range(8)], np.int64)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n", [0, 5, 16, 63, 64, 65, 1000, 65537, 5_000_003])
def test_hist_rna_exact(dev, n):
    rng = np.random.default_rng(n + 7)
    codes = rng.choice(np.array([0, 1, 2, 3, 0x0C, 0xFF], np.uint8), size=n, p=[.27, .22, .22, .27, .01, .01])
    st = dev.SymbolStream(codes, kind="rna")
    got = dev.histogram(st).cpu().numpy()
    want = np.array([(codes == k).sum() for k in range(8)], np.int64)
    assert np.array_equal(got, want)
    assert np.array_equal(dev.histogram(dev.SymbolStream(codes)).cpu().numpy(), want)     # generic kernel agrees


# ----------------------------------------------------------------------------- sequence
@pytest.mark.parametrize("W", [1, 2, 3, 4, 5, 7, 8, 9, 12, 15, 16, 17, 18, 33, 64])
def test_dense_seq_bit_exact(dev, oracle, W):
    from rnascan_b200 import synth
    st, codes, _ = make_stream(dev, 200_000, 40, seed=W)
    rng = np.random.default_rng(100 + W)
    tab = synth.pssm_table(synth.pfm_rows(W, 4, rng), pseudocount=0.01)
    got = dev.dense_seq(st, tab).cpu().numpy()
    want = oracle.seq_scores(synth.to_text(codes, "rna"), tab)
    assert_same_float(got, want)


def test_dense_seq_nonfinite_table(dev, oracle):
    from rnascan_b200 import synth
    st, codes, _ = make_stream(dev, 50_000, 10, seed=5)
    rng = np.random.default_rng(6)
    pfm = synth.pfm_rows(9, 4, rng)
    pfm[pfm < 0.05] = 0.0
    tab = synth.pssm_table(pfm, pseudocount=0.0)            # holds -inf
    assert np.isinf(tab).any()
    got = dev.dense_seq(st, tab).cpu().numpy()
    want = oracle.seq_scores(synth.to_text(codes, "rna"), tab)
    assert_same_float(got, want)


@pytest.mark.parametrize("n", [0, 3, 7, 8, 4095, 4096, 4097, 4102, 4103, 8192 + 6])
def test_scan_seq_tile_edges(dev, oracle, n):
    from rnascan_b200 import synth
    rng = np.random.default_rng(n)
    codes = rng.integers(0, 4, size=n).astype(np.uint8)
    st = dev.SymbolStream(codes)
    tab = synth.pssm_table(synth.pfm_rows(7, 4, rng))
    pos, sc = dev.scan_seq(st, tab, 0.0)
    want = oracle.seq_scores(synth.to_text(codes, "rna"), tab)
    wpos = oracle.search_hits(want, 0.0)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, want[wpos])


@pytest.mark.parametrize("thr", [6.0, 0.0, -3.5, float("-inf"), 1e9])
def test_scan_seq_hits(dev, oracle, thr):
    from rnascan_b200 import synth
    st, codes, _ = make_stream(dev, 1_000_000, 300, seed=2)
    rng = np.random.default_rng(102)
    tab = synth.pssm_table(synth.pfm_rows(7, 4, rng))
    pos, sc = dev.scan_seq(st, tab, thr, capacity=64)       # tiny capacity -> exercises regrow
    want = oracle.seq_scores(synth.to_text(codes, "rna"), tab)
    wpos = oracle.search_hits(want, thr)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, want[wpos])
    assert np.all(np.diff(pos) > 0)


# ----------------------------------------------------------------------------- structure one-hot
@pytest.mark.parametrize("W", [1, 2, 3, 4, 6, 7, 8, 11, 13, 16, 17, 18, 64])
def test_dense_struct_bit_exact(dev, oracle, W):
    from rnascan_b200 import synth
    st, codes, _ = make_stream(dev, 150_000, 30, seed=W, kind="struct")
    rng = np.random.default_rng(200 + W)
    tab = synth.pssm_table(synth.pfm_rows(W, 7, rng))
    got = dev.dense_struct(st, tab).cpu().numpy()
    want = oracle.alpha_scores(synth.to_text(codes, "struct"), tab, "BEHLMRT")
    assert_same_float(got, want)


@pytest.mark.parametrize("thr", [3.0, float("-inf")])
def test_scan_struct_hits(dev, oracle, thr):
    from rnascan_b200 import synth
    st, codes, _ = make_stream(dev, 700_000, 200, seed=3, kind="struct")
    rng = np.random.default_rng(103)
    tab = synth.pssm_table(synth.pfm_rows(7, 7, rng))
    pos, sc = dev.scan_struct_onehot(st, tab, thr)
    want = oracle.alpha_scores(synth.to_text(codes, "struct"), tab, "BEHLMRT")
    wpos = oracle.search_hits(want, thr)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, want[wpos])


def test_scan_pair_onehot(dev, oracle):
    from rnascan_b200 import synth
    rng = np.random.default_rng(77)
    lengths = synth.record_lengths(400_000, 100, rng)
    ca, off = synth.rna_codes(lengths, rng, n_frac=0.01)
    cb, _ = synth.struct_codes(lengths, rng)
    sa, sb = dev.SymbolStream(ca, off, lengths), dev.SymbolStream(cb, off, lengths)
    ta = synth.pssm_table(synth.pfm_rows(6, 4, rng))
    tb = synth.pssm_table(synth.pfm_rows(6, 7, rng))
    a = oracle.seq_scores(synth.to_text(ca, "rna"), ta)
    b = oracle.alpha_scores(synth.to_text(cb, "struct"), tb, "BEHLMRT")
    for thr in (0.0, float("-inf")):
        pos, qa, qb = dev.scan_pair_onehot(sa, sb, ta, tb, thr)
        with np.errstate(invalid="ignore"):
            wpos = np.nonzero((a.astype(np.float64) > thr) & (b > thr))[0]
        assert np.array_equal(pos, wpos)
        assert_same_float(qa, a[wpos])
        assert_same_float(qb, b[wpos])


# ----------------------------------------------------------------------------- averaged profiles
def _profile_case(dev, total, n_records, seed, dtype, W, zero_frac=0.0, pseudocount=0.01):
    from rnascan_b200 import synth
    rng = np.random.default_rng(seed)
    lengths = synth.record_lengths(total, n_records, rng)
    codes, off = synth.rna_codes(lengths, rng, n_frac=0.005)
    rows = synth.profile_rows(len(codes), rng, lengths=lengths)
    pfm = synth.pfm_rows(W, 7, rng)
    if zero_frac:
        # -inf log-odds (zero-probability letters) meet exact zeros in the profile, as in the
        # reference's example data: 0 * -inf = NaN -> 0, w * -inf -> -DBL_MAX
        pfm[pfm < zero_frac] = 0.0
        rows = rows.astype(np.float64)
        rows[rows < 0.08] = 0.0
        rows /= np.maximum(rows.sum(axis=1, keepdims=True), 1e-300)
    rows = rows.astype(np.float32).astype(dtype)
    tq = synth.pssm_table(pfm, background=[synth.SS_P[c] for c in "BEHLMRT"], pseudocount=pseudocount)
    ts = synth.pssm_table(synth.pfm_rows(W, 4, rng))
    return (dev.SymbolStream(codes, off, lengths), dev.ProfileStream(rows), codes, rows, ts, tq)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("W,zero", [(7, 0.0), (18, 0.05), (30, 0.0)])
def test_dense_profile_bit_exact(dev, oracle, dtype, W, zero):
    st, pf, codes, rows, ts, tq = _profile_case(dev, 120_000, 25, 11 + W, dtype, W, zero,
                                                pseudocount=0.0 if zero else 0.01)
    got = dev.dense_profile(pf, tq, st).cpu().numpy()
    with np.errstate(all="ignore"):
        want = oracle.profile_scores(rows, tq)
    want[window_has_sep(codes, W)] = np.nan
    assert_same_float(got, want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("W,zero,thr", [(7, 0.0, "q0.999"), (7, 0.0, "q0.5"), (12, 0.05, "q0.99"),
                                        (18, 0.0, "q0.9999"), (24, 0.0, "q0.99"), (25, 0.0, "q0.99"),
                                        (7, 0.0, 6.0), (7, 0.05, float("-inf"))])
def test_scan_fused_struct_mode(dev, oracle, dtype, W, zero, thr):
    st, pf, codes, rows, ts, tq = _profile_case(dev, 600_000, 150, 21 + W, dtype, W, zero,
                                                pseudocount=0.0 if zero else 0.01)
    with np.errstate(all="ignore"):
        want = oracle.profile_scores(rows, tq)
    want[window_has_sep(codes, W)] = np.nan
    thr = pick_threshold(want, thr)
    pos, sq, sc, resc = dev.scan_fused(st, pf, None, tq, thr, return_stats=True)
    wpos = oracle.search_hits(want, thr)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, want[wpos])
    assert sq is None


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("W,thr", [(7, 0.0), (7, 1.0), (10, -2.0), (7, float("-inf"))])
def test_scan_fused_and_mode(dev, oracle, dtype, W, thr):
    from rnascan_b200 import synth
    st, pf, codes, rows, ts, tq = _profile_case(dev, 800_000, 200, 31 + W, dtype, W)
    with np.errstate(all="ignore"):
        b = oracle.profile_scores(rows, tq)
    a = oracle.seq_scores(synth.to_text(codes, "rna"), ts)
    pos, sq, sc = dev.scan_fused(st, pf, ts, tq, thr)
    with np.errstate(invalid="ignore"):
        wpos = np.nonzero((a.astype(np.float64) > thr) & (b > thr))[0]
    assert len(wpos) > 0
    assert np.array_equal(pos, wpos)
    assert_same_float(sq, a[wpos])
    assert_same_float(sc, b[wpos])


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("W,thr,cap", [(7, 0.0, None), (7, 1.0, 64), (10, -2.0, None), (18, 0.5, 1000), (7, 6.0, None)])
def test_scan_fused_bg_matches_oracle_and_one_pass(dev, oracle, dtype, W, thr, cap):
    """histogram overlapped with the structure-only candidate scan + rs_refine_hits_seq ==
    histogram -> tables -> one-pass RS_MODE_AND scan == oracle (computed background)."""
    from rnascan_b200 import synth
    st, pf, codes, rows, _, tq = _profile_case(dev, 800_000, 200, 131 + W, dtype, W)
    prob = synth.pfm_rows(W, 4, np.random.default_rng(7 + W))
    seen = []

    def seq_table(counts8):
        seen.append(np.array(counts8[:4], np.int64))
        bg = (np.asarray(counts8[:4], np.float64) + 1) / (float(np.sum(counts8[:4])) + 4)
        return synth.pssm_table(prob, background=list(bg / bg.sum()))

    pos, sq, sc, counts = dev.scan_fused_bg(st, pf, tq, seq_table, thr, capacity=cap)
    want_counts = np.array([(codes == k).sum() for k in range(4)], np.int64)
    assert np.array_equal(counts[:4], want_counts) and np.array_equal(seen[-1], want_counts)
    ts = seq_table(want_counts)
    with np.errstate(all="ignore"):
        b = oracle.profile_scores(rows, tq)
    a = oracle.seq_scores(synth.to_text(codes, "rna"), ts)
    with np.errstate(invalid="ignore"):
        wpos = np.nonzero((a.astype(np.float64) > thr) & (b > thr))[0]
    assert np.array_equal(pos, wpos)
    assert_same_float(sq, a[wpos])
    assert_same_float(sc, b[wpos])
    p1, q1, s1 = dev.scan_fused(st, pf, ts, tq, thr)
    assert np.array_equal(pos, p1) and np.array_equal(_bits(sq), _bits(q1)) and np.array_equal(_bits(sc), _bits(s1))
    if cap:
        assert len(seen) > 2, "a tiny candidate buffer must have forced a regrow + second launch"


@pytest.mark.parametrize("W,thr,cap", [(7, 0.0, None), (7, 6.0, None), (18, 0.5, 1000), (1, 0.2, None)])
def test_scan_fused_bg_counting_inside_the_scan_kernel(dev, oracle, monkeypatch, W, thr, cap):
    """Very long streams let the scan kernel count the letters itself (rs_scan_fused_candidates_counting) instead of
    running the histogram beside it; forced here on a small stream: counts exact, hits identical to the oracle."""
    from rnascan_b200 import synth
    monkeypatch.setattr(dev.BackgroundFusedScan, "COUNT_IN_KERNEL_FROM", 0)
    st, pf, codes, rows, _, tq = _profile_case(dev, 700_003, 180, 531 + W, np.float32, W)
    prob = synth.pfm_rows(W, 4, np.random.default_rng(17 + W))
    seen = []

    def seq_table(counts8):
        seen.append(np.array(counts8[:8], np.int64))
        bg = (np.asarray(counts8[:4], np.float64) + 1) / (float(np.sum(counts8[:4])) + 4)
        return synth.pssm_table(prob, background=list(bg / bg.sum()))

    pos, sq, sc, counts = dev.scan_fused_bg(st, pf, tq, seq_table, thr, capacity=cap)
    want_counts = np.array([(codes == k).sum() for k in range(4)], np.int64)
    assert np.array_equal(counts[:4], want_counts) and not counts[4:].any() and np.array_equal(seen[-1][:4], want_counts)
    ts = seq_table(want_counts)
    with np.errstate(all="ignore"):
        b = oracle.profile_scores(rows, tq)
    a = oracle.seq_scores(synth.to_text(codes, "rna"), ts)
    with np.errstate(invalid="ignore"):
        wpos = np.nonzero((a.astype(np.float64) > thr) & (b > thr))[0]
    assert np.array_equal(pos, wpos)
    assert_same_float(sq, a[wpos])
    assert_same_float(sc, b[wpos])


@pytest.mark.parametrize("W,thr", [(7, 0.0), (12, -1.0), (30, -5.0)])
def test_refine_hits_seq_equals_and_scan(dev, oracle, W, thr):
    """structure-only scan, then rs_refine_hits_seq on its ordered hits == the one-pass AND scan."""
    st, pf, codes, rows, ts, tq = _profile_case(dev, 500_000, 120, 171 + W, np.float32, W)
    p0, _, s0 = dev.scan_fused(st, pf, None, tq, thr)
    pos, sq, sc = dev.refine_hits_seq(st, p0, s0, ts, thr)
    p1, q1, s1 = dev.scan_fused(st, pf, ts, tq, thr)
    assert len(p1) > 0 and len(p0) > len(p1)
    assert np.array_equal(pos, p1) and np.array_equal(_bits(sq), _bits(q1)) and np.array_equal(_bits(sc), _bits(s1))
    # no candidates at all / candidates without structure scores
    e = dev.refine_hits_seq(st, np.zeros(0, np.int64), None, ts, thr)
    assert len(e[0]) == 0 and e[2] is None
    pos2, sq2, none = dev.refine_hits_seq(st, p0, None, ts, thr)
    assert none is None and np.array_equal(pos2, p1) and np.array_equal(_bits(sq2), _bits(q1))


def test_refine_hits_seq_rejects_aliased_counters(dev):
    from rnascan_b200.device import lib, _ptr, HitBuffers
    st = dev.SymbolStream(np.zeros(4096, np.uint8))
    hb = HitBuffers(st.n, 128, st.codes.device)
    t = np.zeros((7, 4))
    rc = lib.rs_refine_hits_seq(_ptr(st.codes), st.n, t.ctypes.data, 7, 0.0, _ptr(hb.counters), hb.capacity,
                                _ptr(hb.pos), _ptr(hb.seq), _ptr(hb.struct), _ptr(hb.counters), _ptr(hb.work),
                                hb.work_bytes, 0)
    assert rc == 1


def test_fused_filter_is_used_and_rare(dev):
    """The fp32 filter must leave only a small fraction of windows for exact re-scoring."""
    st, pf, codes, rows, ts, tq = _profile_case(dev, 2_000_000, 500, 99, np.float32, 7)
    pos, sq, sc, resc = dev.scan_fused(st, pf, ts, tq, 6.0, return_stats=True)
    assert resc < 0.02 * len(codes)


# ----------------------------------------------------------------------------- provisional tables / two-call one-hot scans
def _bg_table_fn(prob, A):
    from rnascan_b200 import synth

    def fn(counts8):
        c = np.asarray(counts8[:A], np.float64)
        bg = (c + 1) / (float(c.sum()) + A)
        return synth.pssm_table(prob, background=list(bg / bg.sum()), pseudocount=0.0)
    return fn


@pytest.mark.parametrize("kind,W,thr,zero", [("rna", 7, "q0.999", False), ("rna", 3, "q0.9", False), ("rna", 8, 6.0, False),
                                             ("rna", 12, "q0.999", False), ("rna", 7, "q0.99", True),
                                             ("struct", 7, "q0.999", False), ("struct", 16, "q0.99", False),
                                             ("struct", 4, "q0.9", True)])
def test_scan_onehot_bg_equals_serial_path(dev, oracle, kind, W, thr, zero):
    """provisional device table -> candidates -> exact finish == histogram -> host table -> one-call scan."""
    from rnascan_b200 import synth
    from rnascan_b200.device import lib, _ptr
    A = 4 if kind == "rna" else 7
    st0, codes, lengths = make_stream(dev, 900_000, 250, seed=300 + W, kind=kind)
    st = dev.SymbolStream(codes, st0.offsets, lengths, kind=kind)
    rng = np.random.default_rng(400 + W)
    pfm = synth.pfm_rows(W, A, rng) + (0.0 if zero else 0.01)
    if zero:
        pfm[pfm < 0.03] = 0.0
    prob = pfm / pfm.sum(axis=1, keepdims=True)
    fn = _bg_table_fn(prob, A)
    counts = dev.histogram(st).cpu().numpy()
    table = fn(counts)
    # the provisional table is within its own margin of the exact one
    buf = torch.zeros(W * A + 1, dtype=torch.float64, device="cuda")
    cdev = torch.from_numpy(counts).cuda()
    assert lib.rs_provisional_table(_ptr(cdev), np.ascontiguousarray(prob).ctypes.data, W, A, _ptr(buf), 0) == 0
    got = buf.cpu().numpy()
    pt, margin = got[:-1].reshape(W, A), got[-1]
    assert np.array_equal(np.isinf(pt), np.isinf(table))
    fin = np.isfinite(table)
    assert 0 < margin < 1e-7 and np.abs(pt[fin] - table[fin]).max() * W < margin / 16
    if kind == "rna":
        want = oracle.seq_scores(synth.to_text(codes, "rna"), table)
    else:
        want = oracle.alpha_scores(synth.to_text(codes, "struct"), table, "BEHLMRT")
    thr = pick_threshold(want, thr)
    wpos = oracle.search_hits(want, thr)
    assert len(wpos) > 0
    pos, sc, cnt, false_cand = dev.scan_onehot_bg(st, prob, fn, thr, capacity=64)      # tiny capacity -> regrow
    assert np.array_equal(cnt, counts)
    # candidates the exact table rejects: exactly the windows that TIE with the threshold (strict >), plus
    # at most those within the margin of it
    w64 = np.asarray(want, np.float64)
    with np.errstate(invalid="ignore"):
        ties = int(np.sum(w64 == thr))
        near = int(np.sum(np.abs(w64 - thr) <= 1e-6))
    # (a float32 tie is not borderline: the double sum sits inside the float's rounding interval)
    assert (ties if kind == "struct" else 0) <= false_cand <= near, (ties, false_cand, near, thr)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, want[wpos])
    # extra slack makes many candidates that the exact table rejects: same hits, in order
    pos2, sc2, _, false2 = dev.scan_onehot_bg(st, prob, fn, thr, extra_margin=0.75)
    assert false2 > 0
    assert np.array_equal(pos2, wpos)
    assert_same_float(sc2, want[wpos])


@pytest.mark.parametrize("other,expect_rescan", [([270_000, 220_000, 220_000, 290_000], False),
                                                 ([900_000, 10_000, 10_000, 80_000], True)])
def test_scan_onehot_bg_sharded_starts_from_local_counts(dev, oracle, other, expect_rescan):
    """Sharded runs decide from the shard's own counts + a verified slack while the all-reduce is in flight;
    a shard whose composition is too far from the global one is re-scanned.  Either way the hits are those
    of the exact table built from the GLOBAL counts."""
    from rnascan_b200 import synth
    W = 7
    st0, codes, lengths = make_stream(dev, 900_000, 250, seed=511, kind="rna")
    st = dev.SymbolStream(codes, st0.offsets, lengths, kind="rna")
    pfm = synth.pfm_rows(W, 4, np.random.default_rng(512)) + 0.01
    prob = pfm / pfm.sum(axis=1, keepdims=True)
    fn = _bg_table_fn(prob, 4)
    extra = torch.tensor(other + [0, 0, 0, 0], dtype=torch.int64, device="cuda")     # "the other shard"
    job = dev.BackgroundOneHotScan(st.n, "rna", st.codes.device, capacity=st.n)
    table = job.launch(st.codes, prob, fn, 2.0, all_reduce=lambda t: t.add_(extra))
    pos, sc, _ = job.results()
    assert job.rescanned == expect_rescan
    local = np.array([(codes == k).sum() for k in range(4)], np.int64)
    assert np.array_equal(job.counts_host.numpy()[:4], local + np.array(other))
    assert np.array_equal(table, fn(np.concatenate([local + np.array(other), np.zeros(4, np.int64)])))
    want = oracle.seq_scores(synth.to_text(codes, "rna"), table)
    wpos = oracle.search_hits(want, 2.0)
    assert len(wpos) > 100
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, want[wpos])


@pytest.mark.parametrize("kind,W", [("rna", 7), ("rna", 11), ("struct", 5)])
def test_onehot_counts_reach_the_host_without_a_copy(dev, oracle, kind, W):
    """One device: the decision pass's first kernel stores the counts into pinned host memory and zeroes the next
    launch's counters (rs_scan_onehot_begin_notify).  Launch after launch -- and mixed with the sharded form on the
    same object -- the counts the host sees are those of the data, never an accumulation; shorter-than-motif input
    still reports its counts."""
    from rnascan_b200 import synth
    from rnascan_b200.device import lib, _ptr
    A = 4 if kind == "rna" else 7
    st0, codes, lengths = make_stream(dev, 300_000, 200, seed=620 + W, kind=kind)
    st = dev.SymbolStream(codes, st0.offsets, lengths, kind=kind)
    pfm = synth.pfm_rows(W, A, np.random.default_rng(621)) + 0.01
    prob = pfm / pfm.sum(axis=1, keepdims=True)
    fn = _bg_table_fn(prob, A)
    local = np.array([(codes == k).sum() for k in range(8)], np.int64)
    local[A:] = 0
    table = fn(local)
    if kind == "rna":
        want = oracle.seq_scores(synth.to_text(codes, "rna"), table)
    else:
        want = oracle.alpha_scores(synth.to_text(codes, "struct"), table, "BEHLMRT")
    wpos = oracle.search_hits(want, 1.0)
    job = dev.BackgroundOneHotScan(st.n, kind, st.codes.device, capacity=st.n)
    for sharded in (False, False, True, False, False):
        job.launch(st.codes, prob, fn, 1.0, all_reduce=(lambda t: t) if sharded else None)
        pos, sc, _ = job.results()
        assert np.array_equal(job.counts_host.numpy()[:8], local), sharded
        assert np.array_equal(pos, wpos)
        assert_same_float(sc, want[wpos])
    # n < W: nothing is scanned, the host is still told
    note = torch.zeros(8, dtype=torch.int64).pin_memory()
    cdev = torch.tensor([5, 6, 7, 8, 0, 0, 0, 0], dtype=torch.int64, device="cuda")
    clear = torch.ones(8, dtype=torch.int64, device="cuda")
    short = torch.zeros(256, dtype=torch.uint8, device="cuda")
    assert lib.rs_scan_onehot_begin_notify(A, _ptr(short), W - 1, _ptr(cdev), np.ascontiguousarray(prob).ctypes.data,
                                           W, 1.0, 0.0, 16, None, 0, note.data_ptr(), 41, _ptr(clear), 0) == 0
    torch.cuda.synchronize()
    assert note.tolist() == [(41 << 48) | v for v in (5, 6, 7, 8, 0, 0, 0, 0)] and clear.sum().item() == 0
    assert lib.rs_scan_onehot_begin_notify(A, _ptr(short), 200, _ptr(cdev), np.ascontiguousarray(prob).ctypes.data,
                                           W, 1.0, 0.0, 16, None, 0, note.data_ptr(), 42, _ptr(cdev), 0) != 0
    assert lib.rs_scan_onehot_begin_notify(A, _ptr(short), 200, _ptr(cdev), np.ascontiguousarray(prob).ctypes.data,
                                           W, 1.0, 0.0, 16, None, 0, note.data_ptr(), 65536, _ptr(clear), 0) != 0


def test_scan_onehot_begin_rejects_wide_motifs(dev):
    st = dev.SymbolStream(np.zeros(5000, np.uint8), kind="rna")
    with pytest.raises(ValueError):
        dev.scan_onehot_bg(st, np.full((17, 4), 0.25), lambda c: np.zeros((17, 4)), 1.0)


# ----------------------------------------------------------------------------- error behaviour
def test_bad_arguments(dev):
    st = dev.SymbolStream(np.zeros(100, np.uint8))
    with pytest.raises(ValueError):
        dev.dense_seq(st, np.zeros((7, 5)))
    with pytest.raises(ValueError):
        dev.dense_seq(st, np.zeros((65, 4)))
    with pytest.raises(ValueError):
        dev.scan_seq(st, np.zeros((7, 4)), float("nan"))


# ----------------------------------------------------------------------------- structure scores as printed (thousandths)
@pytest.mark.parametrize("W", [1, 4, 5, 7, 12, 16])
@pytest.mark.parametrize("kind", ["logodds", "dyadic", "tiny"])
def test_dense_struct_milli_is_python_round3(dev, oracle, W, kind):
    """rs_scores_dense_struct_milli == int(round(score, 3) * 1000) with Python's round (correctly rounded
    decimal, ties to even only on exact ties) of the oracle's float64 score; dyadic tables make exact ties
    (x * 1000 = k + 0.5) common, tiny ones exercise -0.0; NaN / -inf windows get their sentinels."""
    import math
    from rnascan_b200 import synth, _lib
    st, codes, lengths = make_stream(dev, 150_000, 40, 900 + W, kind="struct")
    rng = np.random.default_rng(W)
    if kind == "logodds":
        table = synth.pssm_table(synth.pfm_rows(W, 7, rng), background=[synth.SS_P[c] for c in "BEHLMRT"])
        table[0, 3] = -np.inf
    elif kind == "dyadic":
        table = rng.integers(-40, 41, size=(W, 7)).astype(np.float64) / 16.0 + 0.0625 * (rng.random((W, 7)) < 0.5) / 2
    else:
        table = (rng.random((W, 7)) - 0.5) * 4e-4
    got = dev.dense_struct_milli(st, table).cpu().numpy()
    want_sc = oracle.alpha_scores(synth.to_text(codes, "struct"), table, "BEHLMRT")
    want_sc[window_has_sep(codes, W)] = np.nan
    want = np.empty(len(want_sc), np.int64)
    for k, x in enumerate(want_sc.tolist()):
        if x != x:
            want[k] = _lib.RS_MILLI_NAN
        elif x == -math.inf:
            want[k] = _lib.RS_MILLI_NINF
        else:
            r = round(x, 3)
            want[k] = _lib.RS_MILLI_NEG0 if (r == 0 and math.copysign(1.0, r) < 0) else int(round(r * 1000))
    assert np.array_equal(got.astype(np.int64), want)
    if kind == "dyadic":
        frac = np.abs(want_sc[np.isfinite(want_sc)] * 1000) % 1
        assert (frac == 0.5).sum() > 100                       # exact ties really occurred
    if kind == "tiny":
        assert (want == _lib.RS_MILLI_NEG0).sum() > 100
    # the every-position scan built on it: finite windows only, same positions as the float64 scan
    pos, milli = dev.scan_struct_every_position(st, table)
    p64, s64 = dev.scan_struct_onehot(st, table, float("-inf"))
    assert np.array_equal(pos, p64)
    assert np.array_equal(milli.astype(np.int64), want[pos])


# ----------------------------------------------------------------------------- batched many-PFM scan
@pytest.mark.parametrize("with_seq", [True, False])
def test_scan_batched_float64_rows_through_the_float32_shadow(dev, oracle, with_seq):
    """rs_scan_batched_shadow: float64 rows (what the CLI parses; not float32-representable) are filtered on
    the tensor cores through their float32 shadow and re-scored from the float64 rows -- per motif exactly
    what the oracle gives on the float64 rows, and the path taken is the tensor-core one."""
    from rnascan_b200 import synth
    from rnascan_b200.device import lib
    rng = np.random.default_rng(77)
    M = 40
    lengths = synth.record_lengths(200_000, 50, rng)
    codes, off = synth.rna_codes(lengths, rng, n_frac=0.005)
    g = rng.standard_gamma(0.3, size=(len(codes), 7))
    rows = g / np.maximum(g.sum(axis=1, keepdims=True), 1e-300)
    rows[off + lengths] = 0.0
    rows = np.ascontiguousarray(rows, np.float64)
    assert not np.array_equal(rows, rows.astype(np.float32).astype(np.float64))
    st, pf = dev.SymbolStream(codes, off, lengths), dev.ProfileStream(rows)
    bg = [synth.SS_P[c] for c in "BEHLMRT"]
    widths = rng.integers(7, 13, size=M)
    tq = [synth.pssm_table(synth.pfm_rows(int(w), 7, rng), background=bg) for w in widths]
    ts = [synth.pssm_table(synth.pfm_rows(int(w), 4, rng)) for w in widths]
    text = synth.to_text(codes, "rna")
    thr = 0.5 if with_seq else 4.0
    motif, pos, sq, sc, bases = dev.scan_batched(st, pf, ts if with_seq else None, tq, thr, capacity=1 << 20)
    assert int(lib.rs_last_batched_path()) == 2
    total = 0
    for m in range(M):
        W = int(widths[m])
        with np.errstate(all="ignore"):
            b = oracle.profile_scores(rows, tq[m])
        b[window_has_sep(codes, W)] = np.nan
        keep = b > thr
        if with_seq:
            a = oracle.seq_scores(text, ts[m])
            with np.errstate(invalid="ignore"):
                keep &= a.astype(np.float64) > thr
        want = np.nonzero(keep)[0]
        lo, hi = int(bases[m]), int(bases[m + 1])
        assert np.array_equal(pos[lo:hi], want), m
        assert_same_float(sc[lo:hi], b[want])
        if with_seq:
            assert_same_float(sq[lo:hi], a[want])
        total += len(want)
    assert total > 0 and bases[-1] == total


def test_scan_batched_tensor_filter_at_bf16_rounding_midpoints(dev, oracle):
    """The tensor-core filter rounds profile values to bf16 (8 significant bits: up to 2^-8 relative).  Planted
    windows put, in every row, three quarters of the mass on three channels that all carry the row's maximum
    weight, each value just BELOW the midpoint of two bf16 neighbours (0.25 + 2^-10): the filter sees 0.25 and
    loses 0.75 * 2^-8 * R * S -- more than a 2^-9 guard band.  With the threshold a hair below the planted
    windows' exact scores they must still be found, bit-identical to the per-motif scan."""
    from rnascan_b200 import synth
    rng = np.random.default_rng(2024)
    M, W = 40, 7
    lengths = synth.record_lengths(80_000, 20, rng)
    codes, off = synth.rna_codes(lengths, rng, n_frac=0.0)
    rows = synth.profile_rows(len(codes), rng, lengths=lengths).astype(np.float32)
    q = np.nextafter(np.float32(0.25 + 2.0 ** -10), np.float32(0))
    rest = np.float32(1.0) - np.float32(3) * q
    tq, planted = [], []
    for m in range(M):
        t = rng.integers(1, 9, size=(W, 7)).astype(np.float64) / 8.0          # bf16-representable, <= 1.0
        p0 = int(off[m % len(off)] + 10 + 20 * (m // len(off)))
        for j in range(W):
            ch = rng.permutation(7)
            t[j, ch[:3]] = 4.875                                               # the row maximum, three times (the bf16
                                                                               # grid of the threshold bias does not
                                                                               # hide a 2^-9 guard band at this value)
            t[j, ch[3]] = 0.0
            rows[p0 + j] = 0.0
            rows[p0 + j, ch[:3]] = q
            rows[p0 + j, ch[3]] = rest
        tq.append(t)
        planted.append(p0)
    rows = np.ascontiguousarray(rows)
    st, pf = dev.SymbolStream(codes, off, lengths), dev.ProfileStream(rows)
    exact = []
    for m in range(M):
        with np.errstate(all="ignore"):
            exact.append(oracle.profile_scores(rows, tq[m]))
        exact[m][window_has_sep(codes, W)] = np.nan
    scores = np.array([exact[m][planted[m]] for m in range(M)])
    assert scores.max() - scores.min() < 1e-6                                  # all planted windows tie (nearly)
    thr = float(np.nextafter(scores.min(), -np.inf))
    motif, pos, sq, sc, bases = dev.scan_batched(st, pf, None, tq, thr, capacity=1 << 20, path=2)
    for m in range(M):
        want = np.nonzero(exact[m] > thr)[0]
        lo_, hi_ = int(bases[m]), int(bases[m + 1])
        assert planted[m] in set(want.tolist())
        assert np.array_equal(pos[lo_:hi_], want), m
        assert_same_float(sc[lo_:hi_], exact[m][want])


@pytest.mark.parametrize("path,M,zero", [(1, 9, False), (2, 9, False), (2, 40, True), (2, 300, False)],
                         ids=["cuda-core", "tensor-core", "tensor-core-inf-tables", "tensor-core-two-groups"])
@pytest.mark.parametrize("with_seq", [True, False])
def test_scan_batched_matches_per_motif_oracle(dev, oracle, with_seq, path, M, zero):
    """rs_scan_batched, CUDA-core loop (path 1) and tcgen05 GEMM filter + exact re-score (path 2):
    both must reproduce, per motif, exactly what the oracle gives for that motif alone."""
    from rnascan_b200 import synth
    rng = np.random.default_rng(555 + M)
    lengths = synth.record_lengths(100_000 if zero else (400_000 if M < 100 else 150_000), 100, rng)
    codes, off = synth.rna_codes(lengths, rng, n_frac=0.005)
    rows = synth.profile_rows(len(codes), rng, lengths=lengths)
    if zero:
        rows = rows.astype(np.float64)
        rows[rows < 0.08] = 0.0
        rows = (rows / np.maximum(rows.sum(axis=1, keepdims=True), 1e-300)).astype(np.float32)
    st, pf = dev.SymbolStream(codes, off, lengths), dev.ProfileStream(rows)
    bg = [synth.SS_P[c] for c in "BEHLMRT"]
    widths = rng.integers(7, 13, size=M)
    widths[0], widths[-1] = 12, 1 if not zero else 7
    tq = []
    for w in widths:
        pfm = synth.pfm_rows(int(w), 7, rng)
        if zero:
            pfm[pfm < 0.05] = 0.0
        tq.append(synth.pssm_table(pfm, background=bg, pseudocount=0.0 if zero else 0.01))
    ts = [synth.pssm_table(synth.pfm_rows(int(w), 4, rng)) for w in widths]
    text = synth.to_text(codes, "rna")
    thr = -3.0 if zero else (0.5 if with_seq else 3.0)
    # -inf table rows only give a loose upper bound to the tensor-core filter: allow many candidates
    motif, pos, sq, sc, bases = dev.scan_batched(st, pf, ts if with_seq else None, tq, thr,
                                                 capacity=(1 << 20) if (zero or M > 256) else 128, path=path)
    assert bases[0] == 0 and bases[-1] == len(pos) and np.all(np.diff(bases) >= 0)
    total = 0
    for m in range(M):
        W = int(widths[m])
        with np.errstate(all="ignore"):
            b = oracle.profile_scores(rows, tq[m])
        b[window_has_sep(codes, W)] = np.nan
        keep = b > thr
        if with_seq:
            a = oracle.seq_scores(text, ts[m])
            with np.errstate(invalid="ignore"):
                keep &= a.astype(np.float64) > thr
        want = np.nonzero(keep)[0]
        lo, hi = int(bases[m]), int(bases[m + 1])
        assert np.array_equal(pos[lo:hi], want), m
        assert np.all(motif[lo:hi] == m)
        assert_same_float(sc[lo:hi], b[want])
        if with_seq:
            assert_same_float(sq[lo:hi], a[want])
        total += len(want)
    assert total == len(pos) and total > 0
    assert dev.lib.rs_last_batched_path() == path


# ----------------------------------------------------------------------------- k-mer decision-table scan (W <= 8)
@pytest.mark.parametrize("W", [1, 2, 3, 4, 5, 6, 7, 8, 9, 12])
@pytest.mark.parametrize("q", ["q0.5", "q0.99", "q0.9999"])
def test_scan_seq_all_widths(dev, oracle, W, q):
    from rnascan_b200 import synth
    st, codes, _ = make_stream(dev, 1_500_000, 400, seed=40 + W, n_frac=0.02)
    rng = np.random.default_rng(140 + W)
    pfm = synth.pfm_rows(W, 4, rng)
    if W % 3 == 0:
        pfm[pfm < 0.03] = 0.0                                   # -inf entries
    tab = synth.pssm_table(pfm, pseudocount=0.0 if W % 3 == 0 else 0.01)
    want = oracle.seq_scores(synth.to_text(codes, "rna"), tab)
    thr = pick_threshold(want, q)
    pos, sc = dev.scan_seq(st, tab, thr)
    wpos = oracle.search_hits(want, thr)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, want[wpos])


@pytest.mark.parametrize("n", [1, 7, 8, 27, 28, 29, 35, 36, 7167, 7168, 7169, 7175, 14336 + 3, 3 * 7168])
def test_scan_seq_kmer_tile_edges(dev, oracle, n):
    from rnascan_b200 import synth
    rng = np.random.default_rng(n)
    codes = rng.integers(0, 4, size=n).astype(np.uint8)
    if n > 40:
        codes[rng.integers(0, n, size=3)] = 0x0C
        codes[n // 2] = 0xFF
    st = dev.SymbolStream(codes)
    for W in (4, 7, 8):
        tab = synth.pssm_table(synth.pfm_rows(W, 4, rng))
        want = oracle.seq_scores(synth.to_text(codes, "rna"), tab)
        for thr in (-50.0, 0.0):
            pos, sc = dev.scan_seq(st, tab, thr)
            wpos = oracle.search_hits(want, thr)
            assert np.array_equal(pos, wpos), (W, thr)
            assert_same_float(sc, want[wpos])


def test_scan_seq_threshold_equal_to_a_score(dev, oracle):
    """Strict `>`: a window whose float32 score equals the threshold is not a hit."""
    from rnascan_b200 import synth
    st, codes, _ = make_stream(dev, 300_000, 80, seed=9)
    rng = np.random.default_rng(19)
    tab = synth.pssm_table(synth.pfm_rows(7, 4, rng))
    want = oracle.seq_scores(synth.to_text(codes, "rna"), tab)
    finite = want[np.isfinite(want)]
    thr = float(np.sort(finite)[-50])                           # exactly a score that occurs
    pos, sc = dev.scan_seq(st, tab, thr)
    wpos = oracle.search_hits(want, thr)
    assert np.array_equal(pos, wpos) and not np.any(sc.astype(np.float64) == thr)


@pytest.mark.parametrize("W", [1, 3, 4, 7, 8, 11, 16, 17, 30])
@pytest.mark.parametrize("q", ["q0.5", "q0.999"])
def test_scan_struct_all_widths(dev, oracle, W, q):
    """Structure threshold scan: ballot-mask kernel for W <= 16, generic kernel beyond."""
    from rnascan_b200 import synth
    st, codes, _ = make_stream(dev, 900_000, 250, seed=60 + W, kind="struct")
    rng = np.random.default_rng(160 + W)
    pfm = synth.pfm_rows(W, 7, rng)
    if W % 4 == 0:
        pfm[pfm < (0.02 if W <= 4 else 0.0005)] = 0.0          # some -inf log-odds, most windows still finite
    tab = synth.pssm_table(pfm, background=[synth.SS_P[c] for c in "BEHLMRT"], pseudocount=0.0 if W % 4 == 0 else 0.01)
    want = oracle.alpha_scores(synth.to_text(codes, "struct"), tab, "BEHLMRT")
    thr = pick_threshold(want, q)
    pos, sc = dev.scan_struct_onehot(st, tab, thr, capacity=1000)
    wpos = oracle.search_hits(want, thr)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, want[wpos])


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1023, 1024, 1025, 8191, 8192, 8193, 8192 + 1030, 3 * 8192])
def test_scan_struct_tile_edges(dev, oracle, n):
    from rnascan_b200 import synth
    rng = np.random.default_rng(n)
    codes = rng.integers(0, 7, size=n).astype(np.uint8)
    if n > 40:
        codes[rng.integers(0, n, size=3)] = 0x0F
        codes[n // 2] = 0xFF
        codes[rng.integers(0, n, size=5)] |= 8
        codes[n // 2] = 0xFF
    st = dev.SymbolStream(codes)
    for W in (5, 7, 16):
        tab = synth.pssm_table(synth.pfm_rows(W, 7, rng))
        want = oracle.alpha_scores(synth.to_text(codes, "struct"), tab, "BEHLMRT")
        for thr in (-80.0, 0.0):
            pos, sc = dev.scan_struct_onehot(st, tab, thr)
            wpos = oracle.search_hits(want, thr)
            assert np.array_equal(pos, wpos), (W, thr)
            assert_same_float(sc, want[wpos])
