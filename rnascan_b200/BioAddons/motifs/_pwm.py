"""Drop-in for the reference's C extension ``rnascan.BioAddons.motifs._pwm``.

``calculate(sequence, matrix)`` keeps the signature, validation and error messages of
/root/reference/rnascan/BioAddons/motifs/_pwm.c:79-121 and returns a float32 array of
n-m+1 window scores -- computed by the sm_100a kernel behind ``rs_scores_dense_seq``.
"""
import numpy as np

from ... import device


def calculate(sequence, matrix):
    """calculate(sequence, pwm) -> array of score values (float32, NaN for windows that
    contain a letter other than A/C/G/T/U in either case)."""
    if isinstance(sequence, bytes):
        sequence = sequence.decode("latin-1")
    if not isinstance(sequence, str):
        raise TypeError("argument 1 must be str, not %s" % type(sequence).__name__)
    array = np.asarray(matrix)
    if array.dtype != np.float64:
        raise ValueError("position-weight matrix should contain floating-point values")
    if array.ndim != 2:
        raise ValueError("position-weight matrix has incorrect rank (%d expected 2)" % array.ndim)
    if array.shape[1] != 4:
        raise ValueError("position-weight matrix should have four columns (%d columns found)"
                         % array.shape[1])
    n, m = len(sequence), array.shape[0]
    if n - m + 1 <= 0:
        return np.empty(0, dtype=np.float32)
    stream = device.SymbolStream.from_texts([sequence], "rna")
    # the stream holds the record plus its separator; only the record's windows are returned
    return device.dense_seq(stream, array).cpu().numpy()[:n - m + 1].copy()
