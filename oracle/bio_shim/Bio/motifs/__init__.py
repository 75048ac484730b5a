from . import matrix


class Motif(object):
    """Only the counts= constructor path (rnascan.py:244)."""

    def __init__(self, alphabet=None, instances=None, counts=None):
        if counts is None or instances is not None:
            raise NotImplementedError("shim: counts= only")
        self.alphabet = alphabet
        self.counts = matrix.FrequencyPositionMatrix(alphabet, counts)
        self.length = self.counts.length
        self.name = ""
