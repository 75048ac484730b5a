"""Bio-free stand-ins for the few Biopython (<= 1.77) types rnascan's API exposes.

The reference takes/returns ``Bio.Seq.Seq`` / ``Bio.SeqRecord.SeqRecord`` objects and
alphabet instances (rnascan.py:37-40).  Biopython >= 1.78 removed ``Bio.Alphabet``, so the
reference cannot be installed next to a current Biopython; this module provides the same
names with the same behaviour for what the scan path uses.  Objects from a real old
Biopython work too: everything here is duck-typed on ``.seq``, ``.id``, ``.description``,
``.alphabet`` and ``.letters``.
"""
import fileinput


class Alphabet(object):
    size = None
    letters = None

    def __repr__(self):
        return self.__class__.__name__ + "()"

    def _upper(self):
        return self

    def _lower(self):
        return self


class SingleLetterAlphabet(Alphabet):
    size = 1


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


class IUPACAmbiguousDNA(DNAAlphabet):
    letters = "GATCRYWSMKHBVDN"


class IUPACUnambiguousDNA(IUPACAmbiguousDNA):
    letters = "GATC"


class IUPACAmbiguousRNA(RNAAlphabet):
    letters = "GAUCRYWSMKHBVDN"


class IUPACUnambiguousRNA(IUPACAmbiguousRNA):
    letters = "GAUC"          # this order is the dict order of backgrounds and PSSMs


class _IUPACNamespace(object):
    """``IUPAC.IUPACUnambiguousRNA()`` spelling used by the reference and its tests."""
    IUPACAmbiguousDNA = IUPACAmbiguousDNA
    IUPACUnambiguousDNA = IUPACUnambiguousDNA
    IUPACAmbiguousRNA = IUPACAmbiguousRNA
    IUPACUnambiguousRNA = IUPACUnambiguousRNA
    ambiguous_dna = IUPACAmbiguousDNA()
    unambiguous_dna = IUPACUnambiguousDNA()
    ambiguous_rna = IUPACAmbiguousRNA()
    unambiguous_rna = IUPACUnambiguousRNA()


IUPAC = _IUPACNamespace()
generic_alphabet = Alphabet()
single_letter_alphabet = SingleLetterAlphabet()
generic_rna = RNAAlphabet()
generic_dna = DNAAlphabet()


def _mro_names(obj):
    return {c.__name__ for c in type(obj).__mro__}


def is_rna_alphabet(alphabet):
    return "RNAAlphabet" in _mro_names(alphabet)


def is_ambiguous_rna_alphabet(alphabet):
    """isinstance(alphabet, IUPAC.IUPACAmbiguousRNA) -- true for IUPACUnambiguousRNA too."""
    return "IUPACAmbiguousRNA" in _mro_names(alphabet)


def is_nucleotide_alphabet(alphabet):
    return "NucleotideAlphabet" in _mro_names(alphabet)


class Seq(object):
    """Immutable sequence string + alphabet."""

    def __init__(self, data, alphabet=generic_alphabet):
        if not isinstance(data, str):
            raise TypeError("The sequence data given to a Seq object should be a string")
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
        return hash(self._data)

    def __iter__(self):
        return iter(self._data)

    def __getitem__(self, index):
        if isinstance(index, int):
            return self._data[index]
        return Seq(self._data[index], self.alphabet)

    def count(self, sub, start=0, end=None):
        return self._data.count(str(sub), start, len(self._data) if end is None else end)

    def upper(self):
        return Seq(self._data.upper(), self.alphabet._upper())

    def lower(self):
        return Seq(self._data.lower(), self.alphabet._lower())

    def transcribe(self):
        names = _mro_names(self.alphabet)
        if "ProteinAlphabet" in names:
            raise ValueError("Proteins cannot be transcribed!")
        if "RNAAlphabet" in names:
            raise ValueError("RNA cannot be transcribed!")
        if type(self.alphabet) is IUPACUnambiguousDNA:
            alpha = IUPAC.unambiguous_rna
        elif type(self.alphabet) is IUPACAmbiguousDNA:
            alpha = IUPAC.ambiguous_rna
        else:
            alpha = generic_rna
        return Seq(self._data.replace("T", "U").replace("t", "u"), alpha)


class SeqRecord(object):
    def __init__(self, seq, id="<unknown id>", name="<unknown name>",
                 description="<unknown description>"):
        self.seq = seq
        self.id = id
        self.name = name
        self.description = description

    def __len__(self):
        return len(self.seq)

    def __repr__(self):
        return "SeqRecord(seq=%r, id=%r, description=%r)" % (self.seq, self.id, self.description)


def is_seqrecord(obj):
    return hasattr(obj, "seq") and hasattr(obj, "id") and hasattr(obj, "description")


def iter_fasta(handle):
    """(title, sequence) pairs with Biopython's FASTA conventions: text before the first
    '>' is skipped, the title is the header line without '>' (right-stripped), sequence
    lines are right-stripped and joined, blanks and carriage returns removed."""
    title, chunks = None, []
    for line in handle:
        if isinstance(line, bytes):
            line = line.decode("latin-1")
        if line[:1] == ">":
            if title is not None:
                yield title, "".join(chunks).replace(" ", "").replace("\r", "")
            title, chunks = line[1:].rstrip(), []
        elif title is not None:
            chunks.append(line.rstrip())
    if title is not None:
        yield title, "".join(chunks).replace(" ", "").replace("\r", "")


def parse_fasta(fasta_file):
    """Generator of SeqRecord (id = first word of the title, description = whole title).
    `fasta_file` is a path or list of paths; .gz/.bz2 are decompressed on the fly
    (rnascan.py:173 uses fileinput.hook_compressed)."""
    fin = fileinput.input(fasta_file, openhook=fileinput.hook_compressed)
    try:
        for title, sequence in iter_fasta(fin):
            words = title.split(None, 1)
            first = words[0] if words else ""
            yield SeqRecord(Seq(sequence, single_letter_alphabet), id=first, name=first,
                            description=title)
    finally:
        fin.close()
