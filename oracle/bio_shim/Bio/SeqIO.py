"""SeqIO.parse(handle, 'fasta') as in Biopython 1.7x (SimpleFastaParser)."""
from .Alphabet import single_letter_alphabet
from .Seq import Seq
from .SeqRecord import SeqRecord


def _simple_fasta(handle):
    line = handle.readline()
    while True:
        if line == "":
            return
        if line[0] == ">":
            break
        line = handle.readline()
    while True:
        title = line[1:].rstrip()
        chunks = []
        line = handle.readline()
        while line and line[0] != ">":
            chunks.append(line.rstrip())
            line = handle.readline()
        yield title, "".join(chunks).replace(" ", "").replace("\r", "")
        if not line:
            return


def parse(handle, format, alphabet=None):
    if format != "fasta":
        raise ValueError("shim supports only 'fasta'")
    if isinstance(handle, str):
        handle = open(handle)
    alpha = alphabet if alphabet is not None else single_letter_alphabet
    for title, sequence in _simple_fasta(handle):
        words = title.split(None, 1)
        first = words[0] if words else ""
        yield SeqRecord(Seq(sequence, alpha), id=first, name=first,
                        description=title)
