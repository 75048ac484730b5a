"""rnascan_b200 -- B200-native sliding-window motif scoring, drop-in for rnascan's scan path.

Layout (only what the hot path needs):
  csrc/                 hand-written sm_100a CUDA kernels + the C ABI (include/rnascan_b200.h)
  _lib.py               ctypes binding (no CPU fallback)
  device.py             device-buffer plumbing (PyTorch owns memory/streams, nothing else)
  seq.py, motifs.py     Bio-free Seq/SeqRecord/alphabets and PFM -> log-odds preprocessing
  rnascan.py            the reference's scan API + CLI (same names, flags, output format)
  BioAddons/, pfmutil.py  same module paths as the reference
  shard.py              multi-GPU sharding (one process per GPU, one int64 all-reduce)
"""
from .version import __version__  # noqa: F401
