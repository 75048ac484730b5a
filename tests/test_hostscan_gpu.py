"""GPU parity of the filter + gather + resolve scans (exact rows stay on the host; a float32 shadow or an
8-byte quantised form is filtered on the device): rs_filter_profile / rs_host_gather_windows /
rs_resolve_candidates through device.HostProfileScanner, against the CPU oracle (rnascan.py:302-307 and
_pwm.c:34-68 arithmetic) -- bit-exact positions and scores for every filter form."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from test_kernels_gpu import assert_same_float, pick_threshold, window_has_sep, _bits  # noqa: E402


@pytest.fixture(scope="module")
def dev():
    from rnascan_b200 import device
    device.require_cuda()
    return device


def make_case(total, n_records, seed, W, dtype=np.float64, zero_frac=0.0, pseudocount=0.01, n_frac=0.01):
    """codes, rows (dtype; float64 rows are NOT float32-representable, like parsed text), tables."""
    from rnascan_b200 import synth
    rng = np.random.default_rng(seed)
    lengths = synth.record_lengths(total, n_records, rng)
    codes, off = synth.rna_codes(lengths, rng, n_frac=n_frac)
    g = rng.standard_gamma(0.3, size=(len(codes), 7))
    g /= np.maximum(g.sum(axis=1, keepdims=True), 1e-300)
    # smooth along the stream so that windows score high now and then
    c = np.cumsum(np.vstack([np.zeros((1, 7)), g]), axis=0)
    n = len(codes)
    lo, hi = np.maximum(np.arange(n) - 2, 0), np.minimum(np.arange(n) + 3, n)
    rows = c[hi] - c[lo]
    rows /= np.maximum(rows.sum(axis=1, keepdims=True), 1e-300)
    pfm = synth.pfm_rows(W, 7, rng)
    if zero_frac:
        pfm[pfm < zero_frac] = 0.0
        rows[rows < 0.08] = 0.0
        rows /= np.maximum(rows.sum(axis=1, keepdims=True), 1e-300)
    rows[off + lengths] = 0.0
    rows = np.ascontiguousarray(rows.astype(dtype))
    tq = synth.pssm_table(pfm, background=[synth.SS_P[c] for c in "BEHLMRT"], pseudocount=pseudocount)
    ts = synth.pssm_table(synth.pfm_rows(W, 4, rng))
    return codes, off, lengths, rows, ts, tq


def oracle_hits(oracle, codes, rows, ts, tq, thr, W):
    from rnascan_b200 import synth
    with np.errstate(all="ignore"):
        b = oracle.profile_scores(rows, tq)
    sep = window_has_sep(codes, W)
    with np.errstate(invalid="ignore"):
        keep = (b > thr) & ~sep
        a = None
        if ts is not None:
            a = oracle.seq_scores(synth.to_text(codes, "rna"), ts)
            keep &= a.astype(np.float64) > thr
    pos = np.nonzero(keep)[0]
    return pos, (a[pos] if a is not None else None), b[pos]


def run_form(dev, form, codes, rows, seq, tq, thr, chunk_rows=1 << 23, cand_per_row=1.0 / 256):
    hp = dev.HostProfile(rows)
    W = tq.shape[0]
    if form == "q8":
        assert hp.make_q8(codes)
        src, scale, cd = hp.q8, hp.q8_scale, None
    elif form == "q4":
        assert hp.make_q4(codes)
        src, scale, cd = hp.q4, hp.q8_scale, None
    else:
        src, scale, cd = rows, 1.0, codes
    sc = dev.HostProfileScanner(len(codes), W, form, chunk_rows=chunk_rows, cand_per_row=cand_per_row)
    out = sc.run(cd, src, rows, tq, seq, thr, hp.absrow_max(), q8_scale=scale)
    return out, sc


FORMS64 = ["shadow", "q8", "q4"]


@pytest.mark.parametrize("form", FORMS64)
@pytest.mark.parametrize("W,zero,thr", [(7, 0.0, "q0.999"), (7, 0.0, "q0.9"), (1, 0.0, "q0.99"), (12, 0.05, "q0.99"),
                                        (18, 0.0, "q0.9999"), (24, 0.0, "q0.99"), (7, 0.0, 6.0)])
def test_struct_mode_matches_oracle(dev, oracle, form, W, zero, thr):
    codes, off, lengths, rows, ts, tq = make_case(500_000, 120, 300 + W, W, np.float64, zero,
                                                  pseudocount=0.0 if zero else 0.01)
    with np.errstate(all="ignore"):
        thr = pick_threshold(oracle.profile_scores(rows, tq), thr)
    wpos, _, wsc = oracle_hits(oracle, codes, rows, None, tq, thr, W)
    (pos, sq, sc), scanner = run_form(dev, form, codes, rows, None, tq, thr, chunk_rows=65536)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, wsc)
    assert sq is None
    assert scanner.n_candidates >= len(wpos)


@pytest.mark.parametrize("form", ["f32", "shadow", "q8", "q4"])
@pytest.mark.parametrize("W,thr", [(7, 0.0), (7, 1.5), (10, -2.0), (16, -4.0)])
def test_and_mode_matches_oracle(dev, oracle, form, W, thr):
    dtype = np.float32 if form == "f32" else np.float64
    codes, off, lengths, rows, ts, tq = make_case(600_000, 150, 400 + W, W, dtype)
    wpos, wsq, wsc = oracle_hits(oracle, codes, rows, ts, tq, thr, W)
    assert len(wpos) > 0
    (pos, sq, sc), scanner = run_form(dev, form, codes, rows, ts, tq, thr, chunk_rows=100_096)
    assert np.array_equal(pos, wpos)
    assert_same_float(sq, wsq)
    assert_same_float(sc, wsc)


@pytest.mark.parametrize("form", ["f32", "shadow", "q8", "q4"])
def test_computed_background_counts_in_the_same_pass(dev, oracle, form):
    """seq given as a callable of the counts: the device counts the symbols (inside the quantised filter
    kernel / rs_hist_rna) and the sequence table is applied in the resolve step."""
    from rnascan_b200 import synth
    W, thr = 7, 0.5
    dtype = np.float32 if form == "f32" else np.float64
    codes, off, lengths, rows, _, tq = make_case(700_000, 170, 77, W, dtype)
    prob = synth.pfm_rows(W, 4, np.random.default_rng(5))
    seen = []

    def seq_table(counts8):
        seen.append(np.array(counts8[:8], np.int64))
        bg = (np.asarray(counts8[:4], np.float64) + 1) / (float(np.sum(counts8[:4])) + 4)
        return synth.pssm_table(prob, background=list(bg / bg.sum()))

    (pos, sq, sc), scanner = run_form(dev, form, codes, rows, seq_table, tq, thr, chunk_rows=50_176)
    want_counts = np.array([(codes == k).sum() for k in range(4)], np.int64)
    assert np.array_equal(seen[-1][:4], want_counts) and not seen[-1][4:].any()
    wpos, wsq, wsc = oracle_hits(oracle, codes, rows, seq_table(want_counts), tq, thr, W)
    assert len(wpos) > 0
    assert np.array_equal(pos, wpos)
    assert_same_float(sq, wsq)
    assert_same_float(sc, wsc)


@pytest.mark.parametrize("form", FORMS64)
def test_candidate_buffer_regrows(dev, oracle, form):
    W, thr = 7, -3.0                                   # lets a large share of the windows through
    codes, off, lengths, rows, ts, tq = make_case(200_000, 40, 9, W)
    wpos, _, wsc = oracle_hits(oracle, codes, rows, None, tq, thr, W)
    assert len(wpos) > 20_000
    (pos, sq, sc), scanner = run_form(dev, form, codes, rows, None, tq, thr, chunk_rows=131072, cand_per_row=1e-4)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, wsc)
    assert scanner.cap > 4096


@pytest.mark.parametrize("form", FORMS64)
def test_thresholds_at_and_just_below_exact_scores(dev, oracle, form):
    """strict `>`: a window whose exact score EQUALS the threshold is not a hit, one a hair above is --
    neither may be lost or invented by the reduced-precision filter."""
    W = 7
    codes, off, lengths, rows, ts, tq = make_case(150_000, 30, 1234, W)
    with np.errstate(all="ignore"):
        b = oracle.profile_scores(rows, tq)
    b[window_has_sep(codes, W)] = np.nan
    top = np.argsort(np.nan_to_num(b, nan=-1e300))[-40:]
    for k in top[::8]:
        for thr in (float(b[k]), float(np.nextafter(b[k], -np.inf))):
            wpos, _, wsc = oracle_hits(oracle, codes, rows, None, tq, thr, W)
            (pos, _, sc), _ = run_form(dev, form, codes, rows, None, tq, thr)
            assert np.array_equal(pos, wpos)
            assert_same_float(sc, wsc)
            assert (k in set(pos.tolist())) == (b[k] > thr)


def test_shadow_rows_at_float32_rounding_midpoints(dev, oracle):
    """float64 rows exactly half way between two float32 values: the shadow's rounding error is at its
    maximum (2^-24 relative); thresholds a hair below exact scores must still find their windows."""
    W = 7
    codes, off, lengths, rows, ts, tq = make_case(120_000, 20, 4321, W)
    f = rows.astype(np.float32)
    up = np.nextafter(f, np.float32(2.0))
    mid = (f.astype(np.float64) + up.astype(np.float64)) / 2           # exactly representable in float64
    mid[rows == 0] = 0.0
    rows = np.ascontiguousarray(mid)
    with np.errstate(all="ignore"):
        b = oracle.profile_scores(rows, tq)
    b[window_has_sep(codes, W)] = np.nan
    top = np.argsort(np.nan_to_num(b, nan=-1e300))[-30:]
    for k in top[::6]:
        thr = float(np.nextafter(b[k], -np.inf))
        wpos, _, wsc = oracle_hits(oracle, codes, rows, None, tq, thr, W)
        (pos, _, sc), _ = run_form(dev, "shadow", codes, rows, None, tq, thr)
        assert k in set(pos.tolist())
        assert np.array_equal(pos, wpos)
        assert_same_float(sc, wsc)


@pytest.mark.parametrize("n", [0, 3, 7, 8, 1151, 1152, 1153, 2310, 9999])
@pytest.mark.parametrize("form", ["shadow", "q8", "q4"])
def test_tiny_and_tile_edge_lengths(dev, oracle, form, n):
    W = 7
    rng = np.random.default_rng(n + 5)
    rows = rng.dirichlet(0.3 * np.ones(7), size=n) if n else np.zeros((0, 7))
    rows = np.ascontiguousarray(rows, np.float64)
    codes = rng.integers(0, 4, size=n).astype(np.uint8)
    from rnascan_b200 import synth
    tq = synth.pssm_table(synth.pfm_rows(W, 7, rng), background=[synth.SS_P[c] for c in "BEHLMRT"])
    ts = synth.pssm_table(synth.pfm_rows(W, 4, rng))
    thr = -1.0
    hp = dev.HostProfile(rows)
    if form == "q8" and n:
        assert hp.make_q8(codes)
    if form == "q4" and n:
        assert hp.make_q4(codes)
    pos, sq, sc = dev.scan_profile_host(codes, hp, ts, tq, thr, chunk_rows=1024)
    if n >= W:
        wpos, wsq, wsc = oracle_hits(oracle, codes, rows, ts, tq, thr, W)
    else:
        wpos, wsq, wsc = np.zeros(0, np.int64), np.zeros(0, np.float32), np.zeros(0, np.float64)
    assert np.array_equal(pos, wpos)
    assert_same_float(sq, wsq)
    assert_same_float(sc, wsc)


def test_scan_profile_host_falls_back_to_the_exact_kernel(dev, oracle):
    """-m -inf, negative rows, W > 24: the fp32 filter does not apply; results still equal the oracle."""
    W = 7
    codes, off, lengths, rows, ts, tq = make_case(60_000, 12, 55, W)
    for thr, r, t in ((float("-inf"), rows, tq), (0.0, rows - 0.01, tq)):
        pos, sq, sc = dev.scan_profile_host(codes, dev.HostProfile(r), None, t, thr)
        with np.errstate(all="ignore"):
            b = oracle.profile_scores(r, t)
        b[window_has_sep(codes, W)] = np.nan
        wpos = oracle.search_hits(b, thr)
        assert np.array_equal(pos, wpos)
        assert_same_float(sc, b[wpos])
    codes, off, lengths, rows, ts, tq = make_case(60_000, 12, 56, 30)
    pos, sq, sc = dev.scan_profile_host(codes, dev.HostProfile(rows), ts, tq, -6.0)
    wpos, wsq, wsc = oracle_hits(oracle, codes, rows, ts, tq, -6.0, 30)
    assert np.array_equal(pos, wpos)
    assert_same_float(sc, wsc)


def test_quantised_form_declines_rows_it_cannot_hold(dev):
    rows = np.random.default_rng(1).dirichlet(np.ones(7), size=100)
    hp = dev.HostProfile(rows - 0.2)
    assert not hp.make_q8(None) and hp.q8 is None
    bad = rows.copy()
    bad[5, 3] = np.nan
    assert not dev.HostProfile(bad).make_q8(None)
    ok = dev.HostProfile(rows * 3.0)                   # any non-negative range: scale = max value
    assert ok.make_q8(None) and abs(ok.q8_scale - (rows * 3.0).max()) < 1e-12
    q = ok.q8[:, :7].astype(np.float64) * ok.q8_scale / 255
    assert np.abs(q - rows * 3.0).max() <= ok.q8_scale / 510 * (1 + 1e-9)


def test_filter_entry_point_validates_arguments(dev):
    from rnascan_b200.device import lib, _lib
    t = np.zeros((7, 7))
    d = torch.zeros(4096 * 8, dtype=torch.uint8, device="cuda")
    c = torch.zeros(2, dtype=torch.int64, device="cuda")
    args = lambda fmt, W, thr, scale=1.0: lib.rs_filter_profile(d.data_ptr(), d.data_ptr(), fmt, scale, 1000, 0,
                                                                t.ctypes.data, W, thr, 1.0, 0, 0, 0, 0, 0, 0,
                                                                c.data_ptr(), d.data_ptr(), d.numel(), 0)
    assert args(7, 7, 0.0) == _lib.RS_ERR_INVALID                     # unknown row format
    assert args(_lib.RS_ROWS_F32, 25, 0.0) == _lib.RS_ERR_INVALID     # W beyond the filter kernels
    assert args(_lib.RS_ROWS_F32, 7, float("nan")) == _lib.RS_ERR_INVALID
    assert args(_lib.RS_ROWS_F32, 7, float("-inf")) == _lib.RS_ERR_INVALID
    assert args(_lib.RS_ROWS_Q8, 7, 0.0, scale=0.0) == _lib.RS_ERR_INVALID
    assert args(_lib.RS_ROWS_Q8, 7, 0.0) == _lib.RS_OK
    assert args(_lib.RS_ROWS_Q4, 7, 0.0) == _lib.RS_OK
    # candidate symbols are a feature of the 4-bit rows only
    assert lib.rs_filter_profile(d.data_ptr(), d.data_ptr(), _lib.RS_ROWS_Q8, 1.0, 1000, 0, t.ctypes.data, 7, 0.0, 1.0, 0,
                                 0, 0, 0, 0, d.data_ptr(), c.data_ptr(), d.data_ptr(), d.numel(), 0) == _lib.RS_ERR_INVALID


def test_host_fused_scanner_matches_oracle_and_regrows(dev, oracle):
    """Round 1's float32 host pipeline (device.HostFusedScanner, kept as bench.py's e2e_f32 leg): chunks with
    a W-1 overlap, background from the device counts, and hit buffers that grow when a chunk overflows."""
    from rnascan_b200 import synth
    from rnascan_b200 import _lib
    W, thr = 7, -3.0
    codes, off, lengths, rows, _, tq = make_case(300_000, 60, 808, W, np.float32)
    prob = synth.pfm_rows(W, 4, np.random.default_rng(11))

    def tables(counts8):
        bg = (np.asarray(counts8[:4], np.float64) + 1) / (float(np.sum(counts8[:4])) + 4)
        return synth.pssm_table(prob, background=list(bg / bg.sum())), tq

    n = len(codes)
    h_codes = torch.from_numpy(codes).pin_memory()
    h_prof = torch.from_numpy(rows).pin_memory()
    pipe = dev.HostFusedScanner(n, W, chunk_rows=65536, hits_per_row=1e-5)
    cap0 = pipe.hb[0].capacity
    amax = float(np.abs(rows).sum(axis=1).max())
    pos, sq, sc = pipe.run(h_codes, h_prof, tables, thr, absrow_max=amax)
    counts = np.array([(codes == k).sum() for k in range(4)] + [0] * 4, np.int64)
    wpos, wsq, wsc = oracle_hits(oracle, codes, rows, tables(counts)[0], tq, thr, W)
    assert np.array_equal(pos, wpos)
    assert_same_float(sq, wsq)
    assert_same_float(sc, wsc)
    # structure only: far more hits than a chunk's first buffer holds
    pos, sq, sc = pipe.run(h_codes, h_prof, lambda c: (None, tq), thr, mode=_lib.RS_MODE_STRUCT, absrow_max=amax)
    wpos, _, wsc = oracle_hits(oracle, codes, rows, None, tq, thr, W)
    assert len(wpos) > 4 * cap0 and pipe.hb[0].capacity > cap0
    assert np.array_equal(pos, wpos) and sq is None
    assert_same_float(sc, wsc)


def test_four_bit_form_floors_and_carries_symbols(dev, oracle):
    """RS_ROWS_Q4: every stored value is a floor (never above the exact one, less than one step below), the symbol
    sits in the top nibble; with the sequence table arriving after the filter pass the candidates are thinned on
    the device (rs_refine_candidates_packed) before the host gathers anything."""
    from rnascan_b200 import synth
    W, thr = 7, 1.0
    codes, off, lengths, rows, _, tq = make_case(400_000, 90, 4040, W)
    hp = dev.HostProfile(rows)
    assert hp.make_q4(codes)
    words = hp.q4.view(np.uint32).ravel()
    step = hp.q8_scale / 15
    for c in range(7):
        stored = ((words >> (4 * c)) & 0xF).astype(np.float64) * step
        assert (stored <= rows[:, c]).all() and (rows[:, c] - stored < step * (1 + 1e-9)).all()
    assert np.array_equal((words >> 28).astype(np.uint8), codes & 0xF)
    prob = synth.pfm_rows(W, 4, np.random.default_rng(9))

    def seq_table(counts8):
        bg = (np.asarray(counts8[:4], np.float64) + 1) / (float(np.sum(counts8[:4])) + 4)
        return synth.pssm_table(prob, background=list(bg / bg.sum()))

    sc = dev.HostProfileScanner(len(codes), W, "q4", chunk_rows=65536)
    pos, sq, st = sc.run(None, hp.q4, rows, tq, seq_table, thr, hp.absrow_max(), q8_scale=hp.q8_scale)
    counts = np.array([(codes == k).sum() for k in range(4)] + [0] * 4, np.int64)
    wpos, wsq, wsc = oracle_hits(oracle, codes, rows, seq_table(counts), tq, thr, W)
    assert len(wpos) > 0
    assert np.array_equal(pos, wpos)
    assert_same_float(sq, wsq)
    assert_same_float(st, wsc)
    # the structure-only candidates were many, the gathered ones few
    assert sc.n_struct_candidates > 20 * max(sc.n_candidates, 1) and sc.n_candidates >= len(wpos)
    assert dev.q4_guard(tq, hp.q8_scale) > 0.5
