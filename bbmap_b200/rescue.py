"""Host-side mirror of the reference's ungapped scans around candidate sites, batched for the CUDA path:
findTipDeletions(SiteScore, bases, maxImperfectScore, lookRight, lookLeft) (current/align2/AbstractMapThread.java:1107-1141, with
findTipDeletionsRight/Left :2178-2294) and quickRescue (:2303-2405, plus SiteScore.setPerfect / isInBounds of the returned site)."""
import ctypes as C

import numpy as np

from . import lib as _lib

TIPDEL_TASK_DTYPE = np.dtype([("read_off", "<i8"), ("ref_off", "<i8"), ("read_len", "<i4"), ("ref_len", "<i4"), ("min_index", "<i4"),
                              ("start", "<i4"), ("stop", "<i4"), ("slow_score", "<i4"), ("max_imperfect", "<i4"), ("flags", "<i4")], align=True)
TIPDEL_OUT_DTYPE = np.dtype([("start", "<i4"), ("stop", "<i4"), ("right", "<i4"), ("left", "<i4")], align=True)
TIPDEL_CFG_DTYPE = np.dtype([("search_range", "<i4"), ("max_tiplen", "<i4"), ("align_columns", "<i4"), ("slow_rescue_padding", "<i4")], align=True)
RESCUE_TASK_DTYPE = np.dtype([("read_off", "<i8"), ("ref_off", "<i8"), ("read_len", "<i4"), ("ref_len", "<i4"), ("min_index", "<i4"),
                              ("max_index", "<i4"), ("loc", "<i4"), ("search_dist", "<i4"), ("ideal_start", "<i4"), ("max_mismatches", "<i4"),
                              ("flags", "<i4"), ("pad_", "<i4")], align=True)
RESCUE_OUT_DTYPE = np.dtype([("start", "<i4"), ("stop", "<i4"), ("mismatches", "<i4"), ("max_contig", "<i4"), ("score", "<i4"),
                             ("perfect", "<i4"), ("in_bounds", "<i4"), ("pad_", "<i4")], align=True)
RESCUE_CFG_DTYPE = np.dtype([("points_match", "<i4"), ("points_match2", "<i4"), ("use_affine", "<i4"), ("base_hit_score", "<i4")], align=True)
assert TIPDEL_TASK_DTYPE.itemsize == 48 and TIPDEL_OUT_DTYPE.itemsize == 16 and RESCUE_TASK_DTYPE.itemsize == 56 and RESCUE_OUT_DTYPE.itemsize == 32
LOOK_RIGHT, LOOK_LEFT = 1, 2
SEARCH_RIGHT = 1


def tipdel_cfg(search_range=100, max_tiplen=8, align_columns=3000, slow_rescue_padding=8):
    """TIP_SEARCH_DIST=100, SLOW_RESCUE_PADDING=4+SLOW_ALIGN_PADDING=8 (BBMap.java:57-59); TIP_DELETION_MAX_TIPLEN=8
    (AbstractMapThread.java:2989); ALIGN_COLUMNS=3000 (BBMapThread.java)."""
    c = np.zeros(1, TIPDEL_CFG_DTYPE)
    c[0] = (search_range, max_tiplen, align_columns, slow_rescue_padding)
    return c


def rescue_cfg(points_match=70, points_match2=100, use_affine=1, base_hit_score=100):
    """POINTS_MATCH/POINTS_MATCH2 of MultiStateAligner11ts; USE_AFFINE_SCORE=true, BASE_HIT_SCORE=100 (AbstractIndex.java:17)."""
    c = np.zeros(1, RESCUE_CFG_DTYPE)
    c[0] = (points_match, points_match2, use_affine, base_hit_score)
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _need_device(L, what):
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError(f"no CUDA device visible: {what} has no CPU fallback")


def findTipDeletions(ctx, reads, d_ref, tasks, cfg=None):
    """Batch of findTipDeletions(ss, bases, maxImperfectScore, lookRight, lookLeft).  ctx: bbm_ctx handle; d_ref: device pointer of the
    packed chromosome arrays.  Returns TIPDEL_OUT_DTYPE[n] (changed <=> right>0 or left>0)."""
    L = _lib.load(); _need_device(L, "findTipDeletions")
    cfg = tipdel_cfg() if cfg is None else cfg
    reads = np.ascontiguousarray(reads).view(np.int8); tasks = np.ascontiguousarray(tasks, TIPDEL_TASK_DTYPE)
    outs = np.zeros(len(tasks), TIPDEL_OUT_DTYPE)
    _lib.check(L.bbm_tipdel_batch_host(ctx, _p(reads), reads.size, d_ref, _p(tasks), len(tasks), _p(cfg), _p(outs)), "bbm_tipdel_batch_host")
    return outs


def quickRescue(ctx, reads, d_ref, tasks, cfg=None):
    """Batch of quickRescue(bases, chrom, strand, loc, searchDist, searchRight, idealStart, maxAllowedMismatches, ...).
    Returns RESCUE_OUT_DTYPE[n]; start == -1 where the reference returns null."""
    L = _lib.load(); _need_device(L, "quickRescue")
    cfg = rescue_cfg() if cfg is None else cfg
    reads = np.ascontiguousarray(reads).view(np.int8); tasks = np.ascontiguousarray(tasks, RESCUE_TASK_DTYPE)
    outs = np.zeros(len(tasks), RESCUE_OUT_DTYPE)
    _lib.check(L.bbm_rescue_batch_host(ctx, _p(reads), reads.size, d_ref, _p(tasks), len(tasks), _p(cfg), _p(outs)), "bbm_rescue_batch_host")
    return outs
