"""Seq as in Biopython <= 1.77: an immutable string with an alphabet."""
from . import Alphabet
from .Alphabet import IUPAC


class Seq(object):
    def __init__(self, data, alphabet=Alphabet.generic_alphabet):
        if not isinstance(data, str):
            raise TypeError("The sequence data given to a Seq object should "
                            "be a string (not another Seq object etc)")
        self._data = data
        self.alphabet = alphabet

    def __str__(self):
        return self._data

    def __repr__(self):
        return "Seq(%r, %r)" % (self._data, self.alphabet)

    def __len__(self):
        return len(self._data)

    def __eq__(self, other):
        return str(self) == str(other)

    def __hash__(self):
        return hash(str(self))

    def __getitem__(self, index):
        if isinstance(index, int):
            return self._data[index]
        return Seq(self._data[index], self.alphabet)

    def __iter__(self):
        return iter(self._data)

    def count(self, sub, start=0, end=None):
        # non-overlapping count, exactly str.count
        if end is None:
            end = len(self._data)
        return self._data.count(str(sub), start, end)

    def upper(self):
        return Seq(self._data.upper(), self.alphabet._upper())

    def lower(self):
        return Seq(self._data.lower(), self.alphabet._lower())

    def transcribe(self):
        base = Alphabet._get_base_alphabet(self.alphabet)
        if isinstance(base, Alphabet.ProteinAlphabet):
            raise ValueError("Proteins cannot be transcribed!")
        if isinstance(base, Alphabet.RNAAlphabet):
            raise ValueError("RNA cannot be transcribed!")
        if self.alphabet == IUPAC.unambiguous_dna:
            alphabet = IUPAC.unambiguous_rna
        elif self.alphabet == IUPAC.ambiguous_dna:
            alphabet = IUPAC.ambiguous_rna
        else:
            alphabet = Alphabet.generic_rna
        return Seq(self._data.replace("T", "U").replace("t", "u"), alphabet)
