"""Host-side mirror of the reference's BandedAligner plug-in for the CUDA path.

Same operations as current/align2/BandedAlignerJNI.java:40-46 (alignForward / alignForwardRC / alignReverse /
alignReverseRC, returning edits and {lastQueryLoc,lastRefLoc,lastRow,lastEdits,lastOffset}), batched.  Parity target is the
JNI C (jni/BandedAlignerJNI.c), which differs from BandedAlignerConcrete.java.
"""
import ctypes as C

import numpy as np

from . import lib as _lib
from .workloads import BAND_TASK_DTYPE, BAND_OUT_DTYPE


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class BandedAlignerCUDA:
    def __init__(self, device=0, ctx=None):
        self.L = _lib.load()
        if self.L.bbm_device_count() <= 0:
            raise _lib.BbmError("no CUDA device visible: BandedAlignerCUDA has no CPU fallback")
        self._own = ctx is None
        if ctx is None:
            h = C.c_void_p()
            _lib.check(self.L.bbm_init(device, C.byref(h)), "bbm_init")
            ctx = h
        self.h = ctx

    def close(self):
        if self._own and getattr(self, "h", None):
            self.L.bbm_destroy(self.h)
        self.h = None

    def align_batch(self, queries, refs, tasks):
        q = np.ascontiguousarray(queries).view(np.int8); r = np.ascontiguousarray(refs).view(np.int8)
        tasks = np.ascontiguousarray(tasks, BAND_TASK_DTYPE)
        outs = np.zeros(len(tasks), BAND_OUT_DTYPE)
        _lib.check(self.L.bbm_banded_batch_host(self.h, _p(q), q.size, _p(r), r.size, _p(tasks), _p(outs), len(tasks)), "bbm_banded_batch_host")
        return outs

    def align_batch_dev(self, d_q, d_r, d_tasks, d_outs, n, stream=None):
        ms = C.c_float(0)
        _lib.check(self.L.bbm_banded_batch_dev(self.h, d_q, d_r, d_tasks, d_outs, n, stream, C.byref(ms)), "bbm_banded_batch_dev")
        return ms.value
