"""Alphabet hierarchy as in Biopython <= 1.77 (only what rnascan touches)."""


class Alphabet(object):
    size = None
    letters = None

    def __repr__(self):
        return self.__class__.__name__ + "()"

    def contains(self, other):
        return isinstance(other, self.__class__)

    def _upper(self):
        return self

    def _lower(self):
        return self


generic_alphabet = Alphabet()


class SingleLetterAlphabet(Alphabet):
    size = 1
    letters = None


single_letter_alphabet = SingleLetterAlphabet()


class ProteinAlphabet(SingleLetterAlphabet):
    pass


class NucleotideAlphabet(SingleLetterAlphabet):
    pass


class DNAAlphabet(NucleotideAlphabet):
    pass


class RNAAlphabet(NucleotideAlphabet):
    pass


class SecondaryStructure(SingleLetterAlphabet):
    letters = "HSTC"


generic_protein = ProteinAlphabet()
generic_nucleotide = NucleotideAlphabet()
generic_dna = DNAAlphabet()
generic_rna = RNAAlphabet()


def _get_base_alphabet(alphabet):
    return alphabet


from . import IUPAC  # noqa: E402,F401
