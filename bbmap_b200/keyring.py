"""Host-side mirror of the reference's seeding stage (KeyRing / QualityTools as driven by AbstractMapThread.quickMap,
current/align2/AbstractMapThread.java:643-733) for the CUDA path: batched, one call per read batch."""
import ctypes as C

import numpy as np

from . import lib as _lib

SEED_CFG_DTYPE = np.dtype([("keylen", "<i4"), ("maxDesiredKeys", "<i4"), ("baseKeyHitScore", "<i4"), ("minApproxHitsToKeep", "<i4"),
                           ("keyDensity", "<f4"), ("maxKeyDensity", "<f4"), ("minKeyDensity", "<f4"), ("pad_", "<f4")], align=True)


def default_cfg():
    """BBMap.setDefaults (current/align2/BBMap.java:45-65); BASE_KEY_HIT_SCORE = 100*k (AbstractIndex.java:17)."""
    c = np.zeros(1, SEED_CFG_DTYPE)
    c[0] = (13, 15, 1300, 1, 1.9, 3.0, 1.5, 0.0)
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class KeyRingCUDA:
    def __init__(self, device=0, ctx=None):
        self.L = _lib.load()
        if self.L.bbm_device_count() <= 0:
            raise _lib.BbmError("no CUDA device visible: KeyRingCUDA has no CPU fallback")
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

    def seed_batch(self, bases, quality, read_off, cfg=None, maxKeys=96, minus=True):
        """Returns dict(nkeys, offsets, keys, keyScores, baseScores[, offsetsM, keysM])."""
        cfg = default_cfg() if cfg is None else cfg
        bases = np.ascontiguousarray(bases).view(np.int8)
        quality = None if quality is None else np.ascontiguousarray(quality).view(np.int8)
        read_off = np.ascontiguousarray(read_off, np.int64)
        n = len(read_off) - 1
        out = dict(nkeys=np.zeros(n, np.int32), offsets=np.zeros((n, maxKeys), np.int32), keys=np.zeros((n, maxKeys), np.int32),
                   keyScores=np.zeros((n, maxKeys), np.int32), baseScores=np.zeros(len(bases), np.int8))
        if minus:
            out["offsetsM"] = np.zeros((n, maxKeys), np.int32); out["keysM"] = np.zeros((n, maxKeys), np.int32)
        _lib.check(self.L.bbm_seed_batch_host(self.h, _p(bases), _p(quality), _p(read_off), n, _p(cfg), maxKeys, _p(out["nkeys"]),
                                             _p(out["offsets"]), _p(out["keys"]), _p(out["keyScores"]), _p(out["baseScores"]),
                                             _p(out.get("offsetsM")), _p(out.get("keysM"))), "bbm_seed_batch_host")
        return out
