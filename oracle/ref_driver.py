"""TEST / BENCH INFRASTRUCTURE ONLY -- drives the reference's CPU path the way the reference does.

Used by ``bench.py --impl reference`` (and nothing in the product).  /root/reference does not
exist on the GPU box, so the reference's Python cannot be imported there; what travels is
the reference's own compiled kernel ``oracle/_ref/_pwm.so`` (built from
/root/reference/rnascan/BioAddons/motifs/_pwm.c by oracle/Makefile).  This module restates
the Python call pattern around it:

* sequence   rnascan.py:263 -> Biopython<=1.77 ``search``: for every position slice the window,
             ``calculate(window)`` (matrix.py:68-81) which REBUILDS the m x 4 log-odds list
             (matrix.py:57-59) and calls ``_pwm.calculate`` once per window, keep ``score > m``;
* structure  rnascan.py:302-307: two ``DataFrame.iloc`` row extractions + ``np.dot`` +
             ``np.nan_to_num`` per (window, motif row);
* one-hot    matrix.py:25-43: pure-Python dict lookups per (window, motif row);
* fan-out    rnascan.py:388-395: ``multiprocessing.Pool(cores).map`` over records.

Plain ``str`` slices are used where the reference slices ``Bio.Seq`` objects, which makes this
driver slightly FASTER than the real thing.
"""
import importlib.machinery
import importlib.util
import multiprocessing
import os
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PWM = None
_KIND = None


def _pwm():
    """The reference's compiled _pwm extension if present, else a per-window ctypes call of
    the C port (same arithmetic, oracle/pwm_oracle.c)."""
    global _PWM, _KIND
    if _PWM is not None:
        return _PWM
    path = os.path.join(_HERE, "_ref", "_pwm.so")
    try:
        loader = importlib.machinery.ExtensionFileLoader("_pwm", path)
        spec = importlib.util.spec_from_loader("_pwm", loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        mod.calculate("ACGU", np.zeros((4, 4)))
        _PWM, _KIND = mod.calculate, "reference"
    except Exception:
        from . import oracle as orc

        def calculate(sequence, matrix):
            return orc.seq_scores(sequence, np.asarray(matrix, dtype=np.float64))
        _PWM, _KIND = calculate, "port"
    return _PWM


def kind():
    _pwm()
    return _KIND


def describe(workload):
    k = kind()
    seq = ("one %s _pwm.calculate call per window (matrix.py:57-60 list rebuild included)"
           % ("reference-compiled" if k == "reference" else "C-port"))
    avg = "pandas .iloc + np.dot + nan_to_num per (window, row) as rnascan.py:302-307"
    one = "pure-Python dict lookups per (window, row) as matrix.py:25-43"
    return {"c2": seq, "c3": one, "c4": seq + " AND " + avg,
            "c5": "for each of the 256 motif pairs: " + seq + " AND " + avg}[workload]


# ----------------------------------------------------------------------------- per-record work
def _search_seq(text, pssm, m, threshold):
    """Biopython<=1.77 search() + matrix.py calculate()/_calculate() for a nucleotide PSSM."""
    calc = _pwm()
    letters = "ACGU"
    hits = 0
    for position in range(0, len(text) - m + 1):
        s = text[position:position + m]
        logodds = [[pssm[letter][i] for letter in letters] for i in range(m)]
        scores = calc(s, logodds)
        score = scores[0] if len(scores) == 1 else scores
        if score > threshold:
            hits += 1
    return hits


def _search_onehot(text, pssm, m, threshold):
    """search() + _py_calculate (matrix.py:25-43) for the structure alphabet."""
    hits = 0
    text = text.upper()
    for position in range(0, len(text) - m + 1):
        s = text[position:position + m]
        score = 0.0
        for pos in range(m):
            try:
                score += pssm[s[pos]][pos]
            except KeyError:
                score = float("nan")
                break
        if score > threshold:
            hits += 1
    return hits


def _scan_averaged(rows, pssm_frame, threshold):
    """rnascan.py:296-314 on an in-memory profile (no file parsing charged)."""
    import pandas as pd
    struct = pd.DataFrame(rows, columns=list("BEHLMRT"))
    pm = pssm_frame
    N = len(pm.index)
    hits = 0
    for i in range(0, len(struct.index) - N + 1):
        score = 0
        for j in range(0, N):
            score += np.nan_to_num(np.dot(struct.iloc[i + j, :], pm.iloc[j, :]))
        if score > threshold:
            hits += 1
    return hits


def _record_task(task):
    workload, text, rows, seq_pssm, str_pssm, m, threshold = task
    hits = 0
    if workload == "c5":                         # one rnascan run per motif pair
        import pandas as pd
        for sp, qp in zip(seq_pssm, str_pssm):
            w = len(sp["A"])
            hits += _search_seq(text, sp, w, threshold)
            hits += _scan_averaged(rows, pd.DataFrame({c: qp[c] for c in "BEHLMRT"}), threshold)
        return hits
    if workload in ("c2", "c4"):
        hits += _search_seq(text, seq_pssm, m, threshold)
    if workload == "c3":
        hits += _search_onehot(text, str_pssm, m, threshold)
    if workload == "c4":
        import pandas as pd
        hits += _scan_averaged(rows, pd.DataFrame({c: str_pssm[c] for c in "BEHLMRT"}), threshold)
    return hits


# ----------------------------------------------------------------------------- steps
def _pssm_dicts(workload, tables, codes):
    counts = np.array([(codes == k).sum() for k in range(8)], np.int64)
    if workload == "c5":
        tsl, tql = tables.lists(counts)
        return ([{l: t[:, k].tolist() for k, l in enumerate("ACGU")} for t in tsl],
                [{l: t[:, k].tolist() for k, l in enumerate("BEHLMRT")} for t in tql])
    ts, tq = tables(counts)
    seq_pssm = None if ts is None else {l: ts[:, k].tolist() for k, l in enumerate("ACGU")}
    str_pssm = None if tq is None else {l: tq[:, k].tolist() for k, l in enumerate("BEHLMRT")}
    return seq_pssm, str_pssm


def _tasks(workload, lengths, offsets, codes, rows, tables, m, threshold):
    from rnascan_b200 import synth
    seq_pssm, str_pssm = _pssm_dicts(workload, tables, codes)
    text = synth.to_text(codes, "struct" if workload == "c3" else "rna").decode("latin-1")
    out = []
    for off, ln in zip(offsets.tolist(), lengths.tolist()):
        r = None if rows is None else np.asarray(rows[off:off + ln], dtype=np.float64)
        out.append((workload, text[off:off + ln], r, seq_pssm, str_pssm, m, threshold))
    return out


def calibrate(workload, tables, m, threshold, windows=300):
    """scored positions per second of ONE core on a small record."""
    from rnascan_b200 import synth
    if workload == "c5":
        windows = 12
    rng = np.random.default_rng(1)
    lengths = np.array([windows + m - 1], np.int64)
    if workload == "c3":
        codes, offsets = synth.struct_codes(lengths, rng)
        rows = None
    else:
        codes, offsets = synth.rna_codes(lengths, rng, n_frac=0.0)
        rows = synth.profile_rows(len(codes), rng, lengths=lengths) if workload in ("c4", "c5") else None
    tasks = _tasks(workload, lengths, offsets, codes, rows, tables, m, threshold)
    _record_task(tasks[0])                       # imports, first-call costs
    t0 = time.perf_counter()
    _record_task(tasks[0])
    return windows / max(time.perf_counter() - t0, 1e-9)


def run_step(workload, lengths, offsets, codes, rows, tables, m, threshold, cores):
    tasks = _tasks(workload, lengths, offsets, codes, rows, tables, m, threshold)
    if cores <= 1:
        return sum(_record_task(t) for t in tasks)
    total = 0
    pool = multiprocessing.Pool(cores)
    try:
        for a in range(0, len(tasks), 2000):     # rnascan.py:389 batches of 2000 records
            total += sum(pool.map(_record_task, tasks[a:a + 2000]))
    finally:
        pool.close()
        pool.join()
    return total
