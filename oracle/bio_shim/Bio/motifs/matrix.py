"""Position matrices as in Biopython <= 1.77 (Bio/motifs/matrix.py):
dict keyed by letter in `alphabet.letters` order; normalize() adds the
pseudocount to every cell then divides each position by its total;
log_odds() re-normalises the background and takes math.log(p/b, 2);
search() scores one window per calculate() call and keeps score > threshold.
"""
import math

from ..Seq import Seq


class GenericPositionMatrix(dict):
    def __init__(self, alphabet, values):
        self.length = None
        for letter in alphabet.letters:
            if self.length is None:
                self.length = len(values[letter])
            elif self.length != len(values[letter]):
                raise Exception("data has inconsistent lengths")
            self[letter] = list(values[letter])
        self.alphabet = alphabet
        self._letters = sorted(self.alphabet.letters)

    @property
    def consensus(self):
        sequence = ""
        for i in range(self.length):
            maximum = float("-inf")
            for letter in self.alphabet.letters:
                count = self[letter][i]
                if count > maximum:
                    maximum = count
                    sequence_letter = letter
            sequence += sequence_letter
        return Seq(sequence, self.alphabet)


class FrequencyPositionMatrix(GenericPositionMatrix):
    def normalize(self, pseudocounts=None):
        counts = {}
        if pseudocounts is None:
            for letter in self.alphabet.letters:
                counts[letter] = [0.0] * self.length
        elif isinstance(pseudocounts, dict):
            for letter in self.alphabet.letters:
                counts[letter] = [float(pseudocounts[letter])] * self.length
        else:
            for letter in self.alphabet.letters:
                counts[letter] = [float(pseudocounts)] * self.length
        for i in range(self.length):
            for letter in self.alphabet.letters:
                counts[letter][i] += self[letter][i]
        return PositionWeightMatrix(self.alphabet, counts)


class PositionWeightMatrix(GenericPositionMatrix):
    def __init__(self, alphabet, counts):
        GenericPositionMatrix.__init__(self, alphabet, counts)
        for i in range(self.length):
            total = sum(float(self[letter][i]) for letter in alphabet.letters)
            for letter in alphabet.letters:
                self[letter][i] /= total
        for letter in alphabet.letters:
            self[letter] = tuple(self[letter])

    def log_odds(self, background=None):
        values = {}
        alphabet = self.alphabet
        if background is None:
            background = dict.fromkeys(self._letters, 1.0)
        else:
            background = dict(background)
        total = sum(background.values())
        for letter in alphabet.letters:
            background[letter] /= total
            values[letter] = []
        for i in range(self.length):
            for letter in alphabet.letters:
                b = background[letter]
                if b > 0:
                    p = self[letter][i]
                    if p > 0:
                        logodds = math.log(p / b, 2)
                    else:
                        logodds = float("-inf")
                else:
                    p = self[letter][i]
                    if p > 0:
                        logodds = float("inf")
                    else:
                        logodds = float("nan")
                values[letter].append(logodds)
        return PositionSpecificScoringMatrix(alphabet, values)


class PositionSpecificScoringMatrix(GenericPositionMatrix):
    def calculate(self, sequence):
        raise ValueError("shim: base-class calculate is DNA-only in Biopython; "
                         "rnascan overrides it")

    def search(self, sequence, threshold=0.0, both=True):
        sequence = sequence.upper()
        n = len(sequence)
        m = self.length
        if both:
            raise NotImplementedError("shim: both=False only (rnascan.py:263)")
        for position in range(0, n - m + 1):
            s = sequence[position:position + m]
            score = self.calculate(s)
            if score > threshold:
                yield (position, score)
