"""IUPAC alphabets; letter ORDER matters (it is the dict order of
backgrounds and PSSMs in the reference: "GAUC")."""
from . import DNAAlphabet, RNAAlphabet, ProteinAlphabet


class IUPACAmbiguousDNA(DNAAlphabet):
    letters = "GATCRYWSMKHBVDN"


class IUPACUnambiguousDNA(IUPACAmbiguousDNA):
    letters = "GATC"


class IUPACAmbiguousRNA(RNAAlphabet):
    letters = "GAUCRYWSMKHBVDN"


class IUPACUnambiguousRNA(IUPACAmbiguousRNA):
    letters = "GAUC"


class IUPACProtein(ProteinAlphabet):
    letters = "ACDEFGHIKLMNPQRSTVWY"


ambiguous_dna = IUPACAmbiguousDNA()
unambiguous_dna = IUPACUnambiguousDNA()
ambiguous_rna = IUPACAmbiguousRNA()
unambiguous_rna = IUPACUnambiguousRNA()
protein = IUPACProtein()
