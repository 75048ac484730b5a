"""Synthetic workloads of the shapes BASELINE.json names.

Everything is produced directly as symbol codes / float32 profile rows (the layouts of
include/rnascan_b200.h) so that 10^8..10^9 symbols can be generated in seconds; `to_text`
maps codes back to letters for the CPU oracle.  numpy only (host); bench.py has a torch
variant of `profile_rows` for device-side generation of the large profile streams.
"""
import numpy as np

RS_SEP, RS_RNA_OTHER = 0xFF, 0x0C          # include/rnascan_b200.h (kept literal: importing this module must
                                          # not load the CUDA library -- bench.py's reference arm uses it)

RNA_P = (0.27, 0.22, 0.22, 0.29)                       # p(A, C, G, U)
# stationary structure-context distribution = example/3p_UTR_background_structural_context.txt
SS_P = {"B": 0.0163181097311479, "E": 0.272087789050946, "H": 0.153012079123538, "L": 0.204624685341275,
        "M": 0.0196001330531237, "R": 0.196989713257981, "T": 0.137367490441988}


def record_lengths(total, n_records, rng):
    """lognormal(ln 2500, 0.9) clipped to [50, 50000], rescaled so the lengths sum to `total`."""
    if n_records <= 0:
        return np.zeros(0, np.int64)
    raw = np.clip(rng.lognormal(np.log(2500.0), 0.9, size=n_records), 50, 50000)
    ln = np.maximum(1, np.floor(raw * (total / raw.sum()))).astype(np.int64)
    ln[-1] += total - int(ln.sum())
    if ln[-1] < 1:                                     # tiny totals: fall back to equal split
        ln = np.full(n_records, total // n_records, np.int64)
        ln[-1] += total - int(ln.sum())
    return ln


def layout(lengths):
    """offsets of each record in the stream (one separator after every record) and stream size."""
    lengths = np.asarray(lengths, np.int64)
    offsets = np.zeros(len(lengths), np.int64)
    if len(lengths) > 1:
        np.cumsum(lengths[:-1] + 1, out=offsets[1:])
    return offsets, int(lengths.sum() + len(lengths))


def rna_codes(lengths, rng, n_frac=0.001):
    """iid bases with p = RNA_P; runs of N (length U[1,50]) covering ~n_frac of the bases;
    0xFF separator after every record."""
    offsets, n = layout(lengths)
    cdf = np.cumsum(RNA_P)
    codes = np.searchsorted(cdf, rng.random(n, dtype=np.float32), side="right").astype(np.uint8)
    np.minimum(codes, 3, out=codes)
    n_runs = int(n * n_frac / 25.5)
    if n_runs:
        starts = rng.integers(0, n, size=n_runs)
        runlen = rng.integers(1, 51, size=n_runs)
        for s, r in zip(starts.tolist(), runlen.tolist()):
            codes[s:s + r] = RS_RNA_OTHER
    codes[offsets + np.asarray(lengths, np.int64)] = RS_SEP
    return codes, offsets


def struct_codes(lengths, rng, stay=0.8):
    """7-state first-order Markov chain with self-transition `stay` whose stationary
    distribution is SS_P, as codes in B,E,H,L,M,R,T order."""
    offsets, n = layout(lengths)
    p = np.array([SS_P[c] for c in "BEHLMRT"])
    p /= p.sum()
    change = rng.random(n, dtype=np.float32) >= stay
    change[0] = True
    draws = np.searchsorted(np.cumsum(p), rng.random(int(change.sum())), side="right").astype(np.uint8)
    np.minimum(draws, 6, out=draws)
    idx = np.cumsum(change) - 1
    codes = draws[idx]
    codes[offsets + np.asarray(lengths, np.int64)] = RS_SEP
    return codes, offsets


def profile_rows(n, rng, alpha=0.2, box=5, lengths=None):
    """(n, 7) float32 rows: Dirichlet(alpha) smoothed with a length-`box` box filter along
    the stream and re-normalised; separator rows (if `lengths` is given) are zero."""
    g = rng.standard_gamma(alpha, size=(n, 7)).astype(np.float64)
    g /= np.maximum(g.sum(axis=1, keepdims=True), 1e-300)
    c = np.cumsum(np.vstack([np.zeros((1, 7)), g]), axis=0)
    lo = np.maximum(np.arange(n) - box // 2, 0)
    hi = np.minimum(np.arange(n) + box // 2 + 1, n)
    sm = c[hi] - c[lo]
    sm /= np.maximum(sm.sum(axis=1, keepdims=True), 1e-300)
    out = sm.astype(np.float32)
    if lengths is not None:
        offsets, _ = layout(lengths)
        out[offsets + np.asarray(lengths, np.int64)] = 0.0
    return out


def pfm_rows(W, A, rng, alpha=0.3):
    """W Dirichlet(alpha) rows over an A-letter alphabet (a synthetic PFM)."""
    return rng.dirichlet(alpha * np.ones(A), size=W)


def pssm_table(pfm, background=None, pseudocount=0.01):
    """normalize(pseudocount) + log_odds(background) on a (W, A) array, same arithmetic as
    rnascan_b200.motifs (column order preserved)."""
    import math
    W, A = pfm.shape
    bg = np.full(A, 1.0) if background is None else np.asarray(background, np.float64)
    bg = bg / bg.sum()
    out = np.empty((W, A), np.float64)
    for i in range(W):
        row = [float(pseudocount) + float(v) for v in pfm[i]]
        tot = sum(row)
        for a in range(A):
            p = row[a] / tot
            out[i, a] = math.log(p / bg[a], 2) if p > 0 else float("-inf")
    return out


_RNA_TEXT = np.full(256, ord("N"), np.uint8)
_RNA_TEXT[:4] = np.frombuffer(b"ACGU", np.uint8)
_RNA_TEXT[RS_SEP] = ord("\n")
_SS_TEXT = np.full(256, ord("X"), np.uint8)
_SS_TEXT[:7] = np.frombuffer(b"BEHLMRT", np.uint8)
_SS_TEXT[8:15] = np.frombuffer(b"behlmrt", np.uint8)
_SS_TEXT[RS_SEP] = ord("\n")


def to_text(codes, kind):
    """codes -> bytes (separators become newlines, invalid symbols N / X)."""
    lut = _RNA_TEXT if kind == "rna" else _SS_TEXT
    return lut[np.asarray(codes, np.uint8)].tobytes()
