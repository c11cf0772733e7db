"""Host-side mirror of the reference's read ingest contract for the CUDA path: Read.validate
(current/stream/Read.java:81-215) and the once-per-read minus-strand copy (AbstractMapThread.java:492-503,
AminoAcid.reverseComplementBases, current/dna/AminoAcid.java:203-211), batched."""
import ctypes as C

import numpy as np

from . import lib as _lib

FIX_JUNK, U_TO_T, TO_UPPER_CASE, LOWER_CASE_TO_N = 1, 2, 4, 8          # Read.FIX_JUNK / U_TO_T / TO_UPPER_CASE / LOWER_CASE_TO_N
READ_JUNK = 1


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def validate_batch(ctx, bases, quality, read_off, flags=0):
    """ctx: a bbm_ctx handle (e.g. MultiStateAligner11tsCUDA.h).  Returns (bases, quality, basesM, read_flags); the inputs are
    left untouched.  quality = phred values (ASCII offset removed) or None for FASTA input."""
    L = _lib.load()
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError("no CUDA device visible: validate_batch has no CPU fallback")
    b = np.ascontiguousarray(bases).view(np.int8).copy()
    q = None if quality is None else np.ascontiguousarray(quality).view(np.int8).copy()
    ro = np.ascontiguousarray(read_off, np.int64)
    n = len(ro) - 1
    bm = np.zeros(len(b), np.int8); fl = np.zeros(n, np.int32)
    _lib.check(L.bbm_ingest_batch_host(ctx, _p(b), _p(q), _p(ro), n, int(flags), _p(bm), _p(fl)), "bbm_ingest_batch_host")
    return b, q, bm, fl
