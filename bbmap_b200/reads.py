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


def break_reads(bases, quality, read_off, names, name_off, max_len, min_len=0, paired=False):
    """ReformatReads.breakReads as AbstractMapThread.run applies it under `maxlen` / `minlen` (current/jgi/ReformatReads.java:1179-1219,
    current/align2/AbstractMapThread.java:441-443): reads longer than max_len are cut into max_len pieces named "<name>_<n>", reads shorter than
    min_len are dropped.  Host code (bbm_break_reads).  Returns dict(bases, quality, read_off, names, name_off, src, piece_start)."""
    L = _lib.load()
    b = np.ascontiguousarray(bases).view(np.int8)
    q = None if quality is None else np.ascontiguousarray(quality).view(np.int8)
    ro = np.ascontiguousarray(read_off, np.int64); nm = np.ascontiguousarray(names).view(np.int8); no = np.ascontiguousarray(name_off, np.int64)
    n = len(ro) - 1
    cnt, nb, nn = C.c_int64(), C.c_int64(), C.c_int64()
    rc = L.bbm_break_reads(_p(b), _p(q), _p(ro), n, _p(nm), _p(no), int(bool(paired)), int(max_len), int(min_len), C.byref(cnt), C.byref(nb), C.byref(nn),
                           None, None, None, None, None, None, None)
    if rc:
        raise _lib.BbmError("bbm_break_reads: %d %s" % (rc, (L.bbm_wire_last_error() or b"").decode()))
    out = dict(bases=np.zeros(nb.value, np.int8), quality=None if q is None else np.zeros(nb.value, np.int8), read_off=np.zeros(cnt.value + 1, np.int64),
               names=np.zeros(nn.value, np.int8), name_off=np.zeros(cnt.value + 1, np.int64), src=np.zeros(cnt.value, np.int64),
               piece_start=np.zeros(cnt.value, np.int32))
    ob = out["bases"] if nb.value else np.zeros(1, np.int8)           # a non-NULL pointer selects the fill pass even when nothing is left
    rc = L.bbm_break_reads(_p(b), _p(q), _p(ro), n, _p(nm), _p(no), int(bool(paired)), int(max_len), int(min_len), C.byref(cnt), C.byref(nb), C.byref(nn),
                           _p(ob), _p(out["quality"]), _p(out["read_off"]), _p(out["names"] if nn.value else np.zeros(1, np.int8)), _p(out["name_off"]),
                           _p(out["src"] if cnt.value else np.zeros(1, np.int64)), _p(out["piece_start"] if cnt.value else np.zeros(1, np.int32)))
    if rc:
        raise _lib.BbmError("bbm_break_reads: %d %s" % (rc, (L.bbm_wire_last_error() or b"").decode()))
    return out
