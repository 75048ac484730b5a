"""TEST INFRASTRUCTURE ONLY -- CPU oracle for rnascan's motif-scoring hot path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product (``rnascan_b200``) never
does and has no CPU fallback.

What is restated here, and from where (reference = /root/reference, v0.10.2):

* ``seq_scores``            _pwm.c:34-68                (via oracle/liboracle.so)
* ``alpha_scores``          BioAddons/motifs/matrix.py:25-43
* ``profile_scores``        rnascan.py:302-307
* ``search_hits``           Biopython<=1.77 PSSM.search called at rnascan.py:263
                            (one window per call, strict ``>``; NaN never passes)
* ``scan_rows``             rnascan.py:258-275          (1-based Start, inclusive End,
                            fragment, round(score, 3) in the score's own dtype)
* ``averaged_rows``         rnascan.py:293-315
* ``background``            rnascan.py:440-465 (+ preprocess_seq :177-204)
* ``pfm_to_pssm``           rnascan.py:238-252 + Biopython normalize()/log_odds()
* ``combine_rows``          rnascan.py:416-434

Parity status: sequence scoring is pinned against the reference's compiled ``_pwm.c``
(oracle/_ref) and the whole chain is pinned against golden vectors produced by running
the reference's own Python over ``oracle/bio_shim`` (tests/golden/make_golden.py).
Biopython's normalize/log_odds/search are NOT in /root/reference; they are restated from
Biopython 1.66-1.77 behaviour -> "parity unpinned" for those three.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RNA_LETTERS = "GAUC"          # IUPACUnambiguousRNA.letters (Biopython)
SS_LETTERS = "EHTBLRM"        # BioAddons/Alphabet/__init__.py:24
PROFILE_CHANNELS = "BEHLMRT"  # profile file column order (pfmutil.py:62-69)


def build(force=False):
    """Compile oracle/liboracle.so (and oracle/_ref when the reference is mounted)."""
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "pwm_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/rnascan/BioAddons/motifs/_pwm.c"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build())
        i64, vp, ci = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
        L.orc_seq_scores.argtypes = [ctypes.c_char_p, i64, vp, i64, vp]
        L.orc_seq_scores.restype = i64
        L.orc_seq_scores_mt.argtypes = [ctypes.c_char_p, i64, vp, i64, vp]
        L.orc_seq_scores_mt.restype = i64
        L.orc_alpha_scores.argtypes = [ctypes.c_char_p, i64, vp, i64, ctypes.c_char_p, ci, vp]
        L.orc_alpha_scores.restype = i64
        L.orc_profile_scores.argtypes = [vp, i64, vp, i64, ci, vp]
        L.orc_profile_scores.restype = i64
        L.orc_profile_scores_f32.argtypes = [vp, i64, vp, i64, ci, vp]
        L.orc_profile_scores_f32.restype = i64
        L.orc_count_letters.argtypes = [ctypes.c_char_p, i64, ctypes.c_char_p, ci, vp]
        L.orc_count_letters.restype = None
        _LIB = L
    return _LIB


def _bytes(s):
    return s if isinstance(s, (bytes, bytearray)) else str(s).encode("latin-1")


# --------------------------------------------------------------------------- scoring
def seq_scores(seq, table_acgu, threads=False):
    """float32[n-m+1]; table_acgu is (m,4) float64 in A,C,G,U column order
    (matrix.py:57-59 sorts the letters)."""
    M = np.ascontiguousarray(table_acgu, dtype=np.float64)
    b = _bytes(seq)
    n, m = len(b), M.shape[0]
    out = np.empty(max(0, n - m + 1), dtype=np.float32)
    fn = lib().orc_seq_scores_mt if threads else lib().orc_seq_scores
    fn(b, n, M.ctypes.data, m, out.ctypes.data)
    return out


def alpha_scores(seq, table, letters):
    """float64[n-m+1]; table is (m, len(letters)) float64, columns in `letters` order."""
    M = np.ascontiguousarray(table, dtype=np.float64)
    b = _bytes(seq)
    n, m = len(b), M.shape[0]
    out = np.empty(max(0, n - m + 1), dtype=np.float64)
    lib().orc_alpha_scores(b, n, M.ctypes.data, m, _bytes(letters), len(letters),
                           out.ctypes.data)
    return out


def profile_scores(profile, table):
    """float64[L-N+1]; profile (L,C) float64 or float32, table (N,C) float64, label-aligned
    channels."""
    M = np.ascontiguousarray(table, dtype=np.float64)
    P = np.ascontiguousarray(profile)
    L_, C = P.shape
    N = M.shape[0]
    assert M.shape[1] == C
    out = np.empty(max(0, L_ - N + 1), dtype=np.float64)
    if P.dtype == np.float32:
        rc = lib().orc_profile_scores_f32(P.ctypes.data, L_, M.ctypes.data, N, C, out.ctypes.data)
    else:
        P = np.ascontiguousarray(P, dtype=np.float64)
        rc = lib().orc_profile_scores(P.ctypes.data, L_, M.ctypes.data, N, C, out.ctypes.data)
    if rc < 0:
        raise ValueError("profile_scores: exactly 7 channels are supported")
    return out


def profile_scores_py(profile, table):
    """Pure-numpy-per-window restatement of rnascan.py:302-307 for tiny inputs; used to
    check the C loop.  The reference's operands are rows of single-block pandas frames =
    STRIDED views, which sends np.dot to BLAS ddot's non-unit-stride loop; Fortran-ordered
    arrays reproduce that here."""
    P = np.asfortranarray(np.asarray(profile, dtype=np.float64))
    M = np.asfortranarray(np.asarray(table, dtype=np.float64))
    N = M.shape[0]
    out = []
    for i in range(0, P.shape[0] - N + 1):
        score = 0
        for j in range(N):
            with np.errstate(invalid="ignore"):
                score += np.nan_to_num(np.dot(P[i + j, :], M[j, :]))
        out.append(score)
    return np.array(out, dtype=np.float64)


def search_hits(scores, threshold):
    """Positions (0-based) with score > threshold, Biopython<=1.77 search semantics:
    NaN and -inf windows never pass, even for threshold=-inf.  float32 scores are
    widened to double for the comparison (SURVEY.md note N1)."""
    s = np.asarray(scores)
    with np.errstate(invalid="ignore"):
        keep = s.astype(np.float64) > float(threshold)
    return np.nonzero(keep)[0]


def round3(score):
    """round(score, 3) exactly as the reference gets it (rnascan.py:273): numpy.float32
    scores round in float32 arithmetic, Python floats in float64."""
    return round(score, 3)


# --------------------------------------------------------------------------- sequences
def preprocess_rna(text):
    """rnascan.py:186-193 for FASTA input scanned with the RNA alphabet:
    transcribe (T->U, t->u) then upper()."""
    return text.replace("T", "U").replace("t", "u").upper()


def parse_fasta(path):
    """(id, description, sequence) per record; Biopython SimpleFastaParser semantics."""
    recs = []
    title, chunks = None, []
    with open(path) as fh:
        for line in fh:
            if line.startswith(">"):
                if title is not None:
                    recs.append((title, "".join(chunks)))
                title, chunks = line[1:].rstrip(), []
            elif title is not None:
                chunks.append(line.rstrip())
    if title is not None:
        recs.append((title, "".join(chunks)))
    out = []
    for title, seq in recs:
        words = title.split(None, 1)
        out.append((words[0] if words else "", title,
                    seq.replace(" ", "").replace("\r", "")))
    return out


def background(seqs, letters):
    """rnascan.py:440-465: p = (count+1)/(sum(count)+len(letters)); dict in `letters`
    order.  `seqs` are already preprocessed strings."""
    counts = np.zeros(len(letters), dtype=np.int64)
    for s in seqs:
        b = _bytes(s)
        lib().orc_count_letters(b, len(b), _bytes(letters), len(letters), counts.ctypes.data)
    total = len(letters) + int(counts.sum())
    return {l: (float(int(c)) + 1) / total for l, c in zip(letters, counts)}


# --------------------------------------------------------------------------- PFM -> PSSM
def read_pfm_table(path):
    """rnascan.py:242-243: tab-separated, first column dropped, header letters."""
    with open(path) as fh:
        rows = [ln.rstrip("\n").split("\t") for ln in fh if ln.strip()]
    header = rows[0][1:]
    cols = {h: [] for h in header}
    for r in rows[1:]:
        for h, v in zip(header, r[1:]):
            cols[h].append(float(v))
    return cols


def pfm_to_pssm(counts, letters, pseudocount=0, background=None):
    """Biopython normalize(pseudocount) then log_odds(background) (called at
    rnascan.py:245,248).  Returns {letter: [log-odds per position]} in `letters` order."""
    W = len(counts[letters[0]])
    vals = {l: [float(pseudocount) + counts[l][i] for i in range(W)] for l in letters}
    for i in range(W):
        total = sum(float(vals[l][i]) for l in letters)
        for l in letters:
            vals[l][i] /= total
    if background is None:
        bg = dict.fromkeys(sorted(letters), 1.0)
    else:
        bg = dict(background)
    tot = sum(bg.values())
    for l in letters:
        bg[l] /= tot
    out = {l: [] for l in letters}
    for i in range(W):
        for l in letters:
            b, p = bg[l], vals[l][i]
            if b > 0:
                out[l].append(math.log(p / b, 2) if p > 0 else float("-inf"))
            else:
                out[l].append(float("inf") if p > 0 else float("nan"))
    return out


def table(pssm, order):
    """(W, len(order)) float64 array with columns in `order`."""
    return np.array([pssm[l] for l in order], dtype=np.float64).T.copy()


# --------------------------------------------------------------------------- hit rows
def scan_rows(motif_id, seq, scores, threshold):
    """rnascan.py:258-275 -> [motif_id, Start(1-based), End, fragment, round(score,3)]."""
    W = len(seq) - len(scores) + 1 if len(scores) else 0
    rows = []
    for pos in search_hits(scores, threshold):
        pos = int(pos)
        sc = scores[pos]
        sc = round3(sc if isinstance(sc, np.float32) else float(sc))
        rows.append([motif_id, pos + 1, pos + W, seq[pos:pos + W], sc])
    return rows


def averaged_rows(motif_id, scores, N, threshold):
    """rnascan.py:309-314 -> [motif_id, i+1, i+N, '.', score] (unrounded float64)."""
    return [[motif_id, int(i) + 1, int(i) + N, ".", float(scores[i])]
            for i in search_hits(scores, threshold)]


def combine_rows(seq_rows, struct_rows):
    """rnascan.py:416-434: inner join on (Sequence_ID, Start, End); rows are
    (seq_id, start, end, seq_score, struct_score) tuples in; sum appended out."""
    idx = {}
    for r in struct_rows:
        idx.setdefault((r[0], r[1], r[2]), []).append(r)
    out = []
    for r in seq_rows:
        for s in idx.get((r[0], r[1], r[2]), []):
            out.append((r[0], r[1], r[2], r[3], s[3], float(r[3]) + float(s[3])))
    return out
