"""PFM -> log-odds preprocessing without Biopython.

The reference builds its PSSM with three Biopython (<= 1.77) calls
(rnascan.py:244-248): ``motifs.Motif(alphabet, counts)``, ``.counts.normalize(pc)`` and
``.log_odds(background)``.  Biopython is not vendored by the reference; the behaviour is
restated here from its published semantics:

  normalize(pc)      every cell gets +pc, then each position is divided by the sum of its
                     cells taken in ``alphabet.letters`` order
  log_odds(bg)       bg=None means 1.0 for every letter; the background is divided by the
                     sum of its values (dict order), then each cell is
                     ``math.log(p / b, 2)``; p == 0 gives -inf; b == 0 gives +inf (p > 0)
                     or NaN
All arithmetic is Python float (IEEE double), executed in the same order.
"""
import math


def normalize_counts(counts, letters, pseudocount=0):
    """{letter: [count per position]} -> {letter: (probability per position)}."""
    width = None
    cells = {}
    for letter in letters:
        column = list(counts[letter])          # KeyError: PFM header lacks this letter
        if width is None:
            width = len(column)
        elif len(column) != width:
            raise Exception("data has inconsistent lengths")
        cells[letter] = column
    pc = None if pseudocount is None else pseudocount
    if isinstance(pc, dict):
        add = {letter: float(pc[letter]) for letter in letters}
    else:
        add = {letter: 0.0 if pc is None else float(pc) for letter in letters}
    for letter in letters:
        cells[letter] = [add[letter] + v for v in cells[letter]]
    for i in range(width or 0):
        total = sum(float(cells[letter][i]) for letter in letters)
        for letter in letters:
            cells[letter][i] /= total
    return {letter: tuple(cells[letter]) for letter in letters}


def log_odds(probabilities, letters, background=None):
    """{letter: probabilities} -> {letter: [log2(p / b)]} in `letters` order."""
    if background is None:
        bg = dict.fromkeys(sorted(letters), 1.0)
    else:
        bg = dict(background)
    total = sum(bg.values())
    for letter in letters:
        bg[letter] /= total
    width = len(probabilities[letters[0]]) if letters else 0
    out = {letter: [] for letter in letters}
    for i in range(width):
        for letter in letters:
            b = bg[letter]
            p = probabilities[letter][i]
            if b > 0:
                value = math.log(p / b, 2) if p > 0 else float("-inf")
            else:
                value = float("inf") if p > 0 else float("nan")
            out[letter].append(value)
    return out


class Motif(object):
    """Counts container mirroring ``Bio.motifs.Motif(alphabet=..., counts=...)``."""

    def __init__(self, alphabet=None, counts=None):
        if counts is None:
            raise ValueError("counts are required")
        self.alphabet = alphabet
        self.counts = {letter: list(counts[letter]) for letter in alphabet.letters}
        lengths = {len(v) for v in self.counts.values()}
        if len(lengths) > 1:
            raise Exception("data has inconsistent lengths")
        self.length = lengths.pop() if lengths else 0

    def pssm(self, pseudocount=0, background=None):
        letters = self.alphabet.letters
        return log_odds(normalize_counts(self.counts, letters, pseudocount), letters, background)


def log_odds_table(prob_rows, background_row):
    """Array version of `log_odds` for many motifs per call (the batched scan): `prob_rows` is a
    (W, A) array of probabilities, `background_row` an (A,) array already normalised to sum 1.
    Computed by the library's host helper in the same arithmetic as math.log(p / b, 2), so the
    result is bit-identical to `log_odds` (tests/test_host_cpu.py checks it)."""
    import numpy as np
    from . import _lib
    prob = np.ascontiguousarray(prob_rows, dtype=np.float64)
    bg = np.ascontiguousarray(background_row, dtype=np.float64)
    out = np.empty_like(prob)
    _lib.check(_lib.lib.rs_host_log_odds(prob.ctypes.data, bg.ctypes.data, prob.shape[0], prob.shape[1],
                                         out.ctypes.data))
    return out
