"""bbmap_b200 — B200 (sm_100a) kernels for BBMap's seed-and-extend mapping core, behind the reference's plug-in API.

Only what the hot path needs lives here: csrc/ (CUDA kernels + the C ABI of libbbmapcuda.so), a ctypes loader, the
host-side mirror of the reference's aligner interface, and the synthetic workload generators shared by tests and bench.
"""
from .lib import BbmError, SO_PATH  # noqa: F401
