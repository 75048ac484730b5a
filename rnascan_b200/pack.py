"""Binary pack of a directory of averaged-structure profiles.

The reference re-parses every ``structure.<id>.txt`` with pandas on each run
(/root/reference/rnascan/rnascan.py:296-297, :351).  With the scan itself at hundreds of Gpos/s the text
is the whole cost of a run, so a scan may leave ``<dir>/rnascan_b200.pack`` behind (``rnascan --pack``)
and later scans map it instead of parsing:

    bytes 0..7    magic  b"RSB200P1"
    bytes 8..15   little-endian uint64: length H of the JSON header that follows
    JSON header   {"version", "n_rows", "files": [[basename, size, mtime_ns, rows, row_offset], ...],
                   "stats": [max_abs_row_sum, n_nonfinite, n_negative, max_abs_value],
                   "q8_scale": float | null, "off_q8", "off_q4", "off_rows"}
    off_q8        uint8  [n_rows][8]   quantised filter rows (include/rnascan_b200.h, RS_ROWS_Q8); byte 7 is
                                       0, or 0xFF on the separator row that follows every profile; absent
                                       (q8_scale null) when the rows do not fit the form
    off_q4        uint8  [n_rows][4]   the 4-byte form (RS_ROWS_Q4: floored 4-bit channels, top nibble 0 or 0xF),
                                       used when its wider guard band still leaves the threshold selective
    off_rows      float64[n_rows][7]   the exact rows, bit-identical to what pandas parses, B,E,H,L,M,R,T

Both sections start at multiples of 4096 and are read through ``numpy.memmap``: a scan copies all of the
8-byte rows into pinned staging buffers on host threads (they are what travels to the device) and touches
only the candidate windows of the 56-byte ones.
A pack is used only if it lists exactly the directory's current files (same order, size and mtime).
"""
import json
import os
import struct

import numpy as np

MAGIC = b"RSB200P1"
NAME = "rnascan_b200.pack"
ALIGN = 4096
VERSION = 2


class ProfilePack(object):
    def __init__(self, path, header, q8, rows, q4=None):
        self.path, self.header = path, header
        self.q8, self.q4, self.rows = q8, q4, rows
        files = header["files"]
        self.names = [f[0] for f in files]
        self.lengths = np.array([f[3] for f in files], np.int64)
        self.offsets = np.array([f[4] for f in files], np.int64)
        self.q8_scale = header.get("q8_scale")
        self.stats = tuple(header["stats"])
        self.n_rows = int(header["n_rows"])
        self._asked = {}          # section name -> how often a page-locked copy was asked for
        self._locked = {}         # section name -> pinned torch tensor holding a copy of the section, or None

    def page_locked(self, name):
        """The quantised section `name` ("q4" | "q8") as a PINNED torch uint8 tensor (a copy held in page-locked host
        memory, from which the device copies without staging), or None.  The copy is made on the SECOND request: a
        one-shot CLI run never pays for it, a process that scans the same pack again and again does once and then
        ships chunks at link speed.  Sections above RNASCAN_PIN_CACHE_BYTES (default 8 GiB) are never copied;
        callers then stage chunk by chunk as before.  (Page-locking the mapping itself with cudaHostRegister was
        tried first: the driver refuses read-only file mappings here.)"""
        if name in self._locked:
            return self._locked[name]
        self._asked[name] = self._asked.get(name, 0) + 1
        arr = getattr(self, name, None)
        if arr is None or self._asked[name] < 2:
            return None
        tensor = None
        if arr.nbytes <= int(os.environ.get("RNASCAN_PIN_CACHE_BYTES", str(8 << 30))):
            import torch
            from . import _lib
            from .device import HOST_THREADS
            try:
                tensor = torch.empty(arr.shape, dtype=torch.uint8, pin_memory=True)
                src = np.asarray(arr)
                _lib.check(_lib.lib.rs_host_copy(tensor.data_ptr(), src.ctypes.data, src.nbytes, HOST_THREADS))
            except RuntimeError:                                # no page-locked memory to be had
                tensor = None
        self._locked[name] = tensor
        return tensor

    def release(self):
        self._locked = {}


def pack_path(directory):
    return os.path.join(directory, NAME)


def _roundup(x):
    return (x + ALIGN - 1) // ALIGN * ALIGN


def _file_entries(files, lengths, offsets):
    out = []
    for path, rows, off in zip(files, lengths.tolist(), offsets.tolist()):
        if path is None or not os.path.exists(path):
            out.append([os.path.basename(path) if path else "", -1, -1, int(rows), int(off)])
        else:
            st = os.stat(path)
            out.append([os.path.basename(path), int(st.st_size), int(st.st_mtime_ns), int(rows), int(off)])
    return out


def write(directory, files, packed_rows, lengths, stats, q8=None, q8_scale=None, names=None, q4=None):
    """Write the pack of `files` (paths, in scan order; or `names` alone for a pack that stands for the
    text files) whose rows are `packed_rows` (sum L + len(files), 7) float64 with a zero separator row
    after each profile.  Written to a temporary name and renamed, so readers never see a partial pack."""
    lengths = np.asarray(lengths, np.int64)
    offsets = np.zeros(len(lengths), np.int64)
    if len(lengths) > 1:
        np.cumsum(lengths[:-1] + 1, out=offsets[1:])
    n_rows = int(packed_rows.shape[0])
    if names is not None:
        entries = [[str(nm), -1, -1, int(r), int(o)] for nm, r, o in zip(names, lengths.tolist(), offsets.tolist())]
    else:
        entries = _file_entries(files, lengths, offsets)
    header = {"version": VERSION, "n_rows": n_rows, "files": entries, "stats": [float(v) for v in stats],
              "q8_scale": None if q8 is None else float(q8_scale), "off_q8": 0, "off_q4": 0, "off_rows": 0}
    if q8 is None:
        q4 = None
    # two passes: the offsets are part of the header whose length they depend on
    for _ in range(3):
        blob = json.dumps(header).encode("utf-8")
        off_q8 = _roundup(16 + len(blob))
        off_q4 = _roundup(off_q8 + (n_rows * 8 if q8 is not None else 0))
        off_rows = _roundup(off_q4 + (n_rows * 4 if q4 is not None else 0))
        if header["off_q8"] == off_q8 and header["off_rows"] == off_rows and header["off_q4"] == (off_q4 if q4 is not None else 0):
            break
        header["off_q8"], header["off_rows"] = off_q8, off_rows
        header["off_q4"] = off_q4 if q4 is not None else 0
    blob = json.dumps(header).encode("utf-8")
    path = pack_path(directory)
    tmp = path + ".tmp.%d" % os.getpid()
    with open(tmp, "wb") as fh:
        fh.write(MAGIC)
        fh.write(struct.pack("<Q", len(blob)))
        fh.write(blob)
        if q8 is not None:
            fh.seek(header["off_q8"])
            np.ascontiguousarray(q8, np.uint8).tofile(fh)
        if q4 is not None:
            fh.seek(header["off_q4"])
            np.ascontiguousarray(q4, np.uint8).tofile(fh)
        fh.seek(header["off_rows"])
        np.ascontiguousarray(packed_rows, np.float64).tofile(fh)
    os.replace(tmp, path)
    return path


# One mapping per pack file and process: a new mapping means a page fault per 4 KB page on the first touch (16 ms per
# 256 MB section, more than the copy to the device takes) and an munmap of gigabytes when it is dropped.  Keyed by the
# file's identity (write() replaces the file, so a refreshed pack is another inode); two entries at most.
_CACHE = {}


def read(directory):
    """The pack of `directory` (memory-mapped) or None when there is none / it is not readable."""
    path = os.path.realpath(pack_path(directory))
    try:
        st = os.stat(path)
    except OSError:
        _drop(path)
        return None
    key = (st.st_ino, st.st_size, st.st_mtime_ns)
    hit = _CACHE.get(path)
    if hit is not None and hit[0] == key:
        return hit[1]
    _drop(path)
    pk = _map(path)
    if pk is not None:
        while len(_CACHE) >= 2:
            _drop(next(iter(_CACHE)))
        _CACHE[path] = (key, pk)
    return pk


def _drop(path):
    hit = _CACHE.pop(path, None)
    if hit is not None:
        hit[1].release()


def _map(path):
    try:
        with open(path, "rb") as fh:
            if fh.read(8) != MAGIC:
                return None
            (hlen,) = struct.unpack("<Q", fh.read(8))
            if hlen > (1 << 31):
                return None
            header = json.loads(fh.read(hlen).decode("utf-8"))
        if header.get("version") != VERSION:
            return None
        n_rows = int(header["n_rows"])
        size = os.path.getsize(path)
        if header["off_rows"] + n_rows * 56 > size:
            return None
        rows = np.memmap(path, dtype=np.float64, mode="r", offset=header["off_rows"], shape=(n_rows, 7)) \
            if n_rows else np.zeros((0, 7), np.float64)
        q8 = None
        if header.get("q8_scale") is not None and n_rows:
            q8 = np.memmap(path, dtype=np.uint8, mode="r", offset=header["off_q8"], shape=(n_rows, 8))
        q4 = None
        if q8 is not None and header.get("off_q4"):
            q4 = np.memmap(path, dtype=np.uint8, mode="r", offset=header["off_q4"], shape=(n_rows, 4))
        return ProfilePack(path, header, q8, rows, q4)
    except (OSError, ValueError, KeyError, struct.error):
        return None


def matches(pack, files):
    """Does the pack describe exactly these files (same order, sizes and modification times)?"""
    entries = pack.header["files"]
    if len(entries) != len(files):
        return False
    for (name, size, mtime, _, _), path in zip(entries, files):
        if name != os.path.basename(path):
            return False
        try:
            st = os.stat(path)
        except OSError:
            return False
        if size != st.st_size or mtime != st.st_mtime_ns:
            return False
    return True
