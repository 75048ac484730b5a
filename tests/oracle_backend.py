"""TEST INFRASTRUCTURE: an oracle-backed stand-in for ``rnascan_b200.device`` so that the HOST
logic of the product (record packing, hit -> (record, start) mapping, frame assembly, CLI flow,
stderr messages, TSV text) can be exercised on a machine without a GPU.

``install(monkeypatch)`` swaps the device entry points for functions that compute the same
results with the CPU oracle on the same packed streams.  Nothing here is shipped or imported
by the product; the GPU tests (-m gpu) run the same cases through the real kernels.
"""
import numpy as np
import torch

from oracle import oracle as orc
from rnascan_b200 import device, synth, _lib


class FakeSymbolStream(object):
    def __init__(self, codes, offsets=None, lengths=None, device=None, kind=None):
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        self.kind = kind
        self.n = int(codes.shape[0])
        self.offsets = np.zeros(1, np.int64) if offsets is None else np.asarray(offsets, np.int64)
        self.lengths = np.array([self.n], np.int64) if lengths is None else np.asarray(lengths, np.int64)
        self._codes = codes
        self.codes = torch.from_numpy(codes.copy())

    @classmethod
    def from_texts(cls, texts, kind, device=None):
        codes, offsets, lengths = device_pack(texts, kind)
        return cls(codes, offsets, lengths, kind=kind)

    def host_codes(self):
        return self._codes

    locate = device.SymbolStream.locate


def device_pack(texts, kind):
    return device.pack_texts(texts, kind)


class FakeProfileStream(object):
    def __init__(self, rows, device=None):
        rows = np.ascontiguousarray(rows)
        if rows.dtype not in (np.float32, np.float64):
            rows = rows.astype(np.float64)
        self.n = rows.shape[0]
        self._rows = rows
        self.dtype = _lib.RS_F32 if rows.dtype == np.float32 else _lib.RS_F64

    def absrow_max(self):
        return float(np.abs(self._rows).sum(axis=1).max()) if self.n else 0.0


def _text(stream, kind):
    return synth.to_text(stream._codes, kind)


def _sep_mask(codes, W):
    sep = (codes == 0xFF).astype(np.int64)
    c = np.concatenate([[0], np.cumsum(sep)])
    n = len(codes) - W + 1
    return (c[W:W + n] - c[:n]) > 0 if n > 0 else np.zeros(0, bool)


def histogram(stream):
    c = stream._codes
    return torch.from_numpy(np.array([(c == k).sum() for k in range(8)], np.int64))


def dense_seq(stream, table):
    t = device._table(table, 4)
    return torch.from_numpy(orc.seq_scores(_text(stream, "rna"), t))


def dense_struct(stream, table):
    t = device._table(table, 7)
    return torch.from_numpy(orc.alpha_scores(_text(stream, "struct"), t, "BEHLMRT"))


def dense_profile(profile, table, stream=None):
    t = device._table(table, 7)
    with np.errstate(all="ignore"):
        sc = orc.profile_scores(profile._rows, t)
    if stream is not None and len(sc):
        sc[_sep_mask(stream._codes, t.shape[0])] = np.nan
    return torch.from_numpy(sc)


def _gt(scores, thr):
    with np.errstate(invalid="ignore"):
        return np.asarray(scores, np.float64) > float(thr)


def scan_seq(stream, table, threshold, capacity=None):
    if float(threshold) != float(threshold):
        raise ValueError("threshold is NaN")
    sc = dense_seq(stream, table).numpy()
    pos = np.nonzero(_gt(sc, threshold))[0].astype(np.int64)
    return pos, sc[pos]


def scan_struct_onehot(stream, table, threshold, capacity=None):
    sc = dense_struct(stream, table).numpy()
    pos = np.nonzero(_gt(sc, threshold))[0].astype(np.int64)
    return pos, sc[pos]


def scan_struct_every_position(stream, table):
    import math
    t = device._table(table, 7)
    if t.shape[0] > 16:
        return None
    sc = dense_struct(stream, t).numpy()
    pos = np.nonzero(np.isfinite(sc))[0].astype(np.int64)
    milli = np.empty(len(pos), np.int32)
    for k, x in enumerate(sc[pos].tolist()):
        r = round(x, 3)
        milli[k] = _lib.RS_MILLI_NEG0 if (r == 0 and math.copysign(1.0, r) < 0) else int(round(r * 1000))
    return pos, milli


def scan_pair_onehot(seq_stream, struct_stream, seq_table, struct_table, threshold, capacity=None):
    a = dense_seq(seq_stream, seq_table).numpy()
    b = dense_struct(struct_stream, struct_table).numpy()
    pos = np.nonzero(_gt(a, threshold) & _gt(b, threshold))[0].astype(np.int64)
    return pos, a[pos], b[pos]


def scan_fused(stream, profile, seq_table, struct_table, threshold, capacity=None, return_stats=False):
    b = dense_profile(profile, struct_table, stream).numpy()
    keep = _gt(b, threshold)
    a = None
    if seq_table is not None:
        a = dense_seq(stream, seq_table).numpy()
        keep &= _gt(a, threshold)
    pos = np.nonzero(keep)[0].astype(np.int64)
    out = (pos, a[pos] if a is not None else None, b[pos])
    return out + (0,) if return_stats else out


def scan_profile_host(codes, hp, seq, struct_table, threshold, all_reduce=None, chunk_rows=1 << 23, form=None,
                      return_scanner=False):
    form_arg = form
    if codes is None:
        codes = np.zeros(hp.n, np.uint8)
    stream = FakeSymbolStream(codes, kind="rna")
    if callable(seq):
        seq = seq(histogram(stream).numpy())
    out = scan_fused(stream, FakeProfileStream(hp.rows), seq, struct_table, threshold)
    if not return_scanner:
        return out
    from rnascan_b200 import device                 # the product's own choice of filter form (pure host logic)
    tq = np.asarray(struct_table, np.float64)
    applies = device.filter_applies(tq, threshold, hp.absrow_max())

    class Scanner(object):
        form = form_arg or device.pick_filter_form(hp, tq, threshold)
        h2d_bytes = 0
        n_candidates = len(out[0])
    return out + (Scanner if applies else None,)


def scan_onehot_bg(stream, prob, table_fn, threshold, all_reduce=None, capacity=None, extra_margin=0.0):
    counts = histogram(stream).numpy()
    table = table_fn(counts)
    fn = scan_seq if (stream.kind or "struct") == "rna" else scan_struct_onehot
    pos, sc = fn(stream, table, threshold)
    return pos, sc, counts, 0


def scan_batched(stream, profile, seq_tables, struct_tables, threshold, capacity=None, path=0):
    motif, pos, sq, st, bases = [], [], [], [], [0]
    for m, tq in enumerate(struct_tables):
        p_, a_, b_ = scan_fused(stream, profile, None if seq_tables is None else seq_tables[m], tq, threshold)
        motif.append(np.full(len(p_), m, np.int32)); pos.append(p_); st.append(b_)
        if seq_tables is not None:
            sq.append(a_)
        bases.append(bases[-1] + len(p_))
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)
    return (cat(motif, np.int32), cat(pos, np.int64), cat(sq, np.float32) if seq_tables is not None else None,
            cat(st, np.float64), np.array(bases, np.int64))


def install(monkeypatch):
    for name, fn in (("SymbolStream", FakeSymbolStream), ("ProfileStream", FakeProfileStream),
                     ("histogram", histogram), ("dense_seq", dense_seq), ("dense_struct", dense_struct),
                     ("dense_profile", dense_profile), ("scan_seq", scan_seq),
                     ("scan_struct_onehot", scan_struct_onehot), ("scan_pair_onehot", scan_pair_onehot),
                     ("scan_struct_every_position", scan_struct_every_position),
                     ("scan_fused", scan_fused),
                     ("scan_profile_host", scan_profile_host), ("scan_onehot_bg", scan_onehot_bg),
                     ("scan_batched", scan_batched)):
        monkeypatch.setattr(device, name, fn)
    from rnascan_b200 import rnascan as ms
    ms._BATCH_CACHE.clear()
