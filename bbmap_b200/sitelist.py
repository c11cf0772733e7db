"""Host-side mirror of the per-read site-list policies of the unpaired mapping loop (part of SURVEY 8f.1), batched for the CUDA path:
Collections.sort + BBMapThread.trimList (current/align2/BBMapThread.java:428-431, 140-249), AbstractMapThread.scoreNoIndels(Read, ...)
(current/align2/AbstractMapThread.java:762-855) and the post-alignment list handling (mergeDuplicateSites, Read.setPerfectFlag, clearzone /
ambiguity, removeLowQualitySitesUnpaired; BBMapThread.java:478-553)."""
import ctypes as C

import numpy as np

from . import lib as _lib

SS_DTYPE = np.dtype([("chrom", "<i4"), ("start", "<i4"), ("stop", "<i4"), ("hits", "<i4"), ("score", "<i4"), ("quick_score", "<i4"),
                     ("slow_score", "<i4"), ("paired_score", "<i4"), ("strand", "i1"), ("perfect", "i1"), ("semiperfect", "i1"), ("rescued", "i1"),
                     ("ngaps", "<i4"), ("gaps", "<i4", (9,)), ("has_match", "<i4")], align=True)
POLICY_CFG_DTYPE = np.dtype([("trim_list", "<i4"), ("min_trim_sites_to_retain", "<i4"), ("max_trim_sites_to_retain", "<i4"), ("quick_match_strings", "<i4"),
                             ("clearzone1", "<i4"), ("clearzone1b", "<i4"), ("clearzone1c", "<i4"), ("clearzonep", "<i4"), ("clearzone3", "<i4"),
                             ("clearzone1e", "<i4"), ("clearzone_limit1e", "<i4"), ("print_secondary", "<i4"), ("min_align_ratio", "<f4"),
                             ("cz1b_scale", "<f4"), ("cz1b_flat", "<f4"), ("cz1c_scale", "<f4"), ("cz1c_flat", "<f4"), ("pad_", "<i4", (3,))], align=True)
READ_OUT_DTYPE = np.dtype([("near_perfect", "<i4"), ("flags", "<i4"), ("clearzone", "<i4"), ("best_sites", "<i4")], align=True)
assert SS_DTYPE.itemsize == 80 and POLICY_CFG_DTYPE.itemsize == 80 and READ_OUT_DTYPE.itemsize == 16
SL_TRIM, SL_NOINDEL, SL_FINAL = 1, 2, 3
F_MAPPED, F_PERFECT, F_AMBIGUOUS = 1, 2, 4


def policy_cfg(**kw):
    """BBMap defaults: TRIM_LIST=true (AbstractMapper.java:2678), MIN_TRIM_SITES_TO_RETAIN_SINGLE=3 (BBMapThread.java:62),
    MAX_TRIM_SITES_TO_RETAIN=800 (AbstractMapThread.java:3004), CLEARZONE1/1b/1c/P/3 = (2.0, 2.6, 4.6, 1.6, 8.0) x POINTS_MATCH2
    (BBMapThread.java:38-42,114-118), CLEARZONE1e = 2*100-70+127+1 = 258 (AbstractMapThread.java:142), CLEARZONE_LIMIT1e=40,
    cutoffs 0.97/12x100 and 0.92/26x100 (BBMapThread.java:52-57), MINIMUM_ALIGNMENT_SCORE_RATIO=0.56 (BBMap.java:50),
    QUICK_MATCH_STRINGS=false (AbstractMapper.java:2731)."""
    c = np.zeros(1, POLICY_CFG_DTYPE)
    d = dict(trim_list=1, min_trim_sites_to_retain=3, max_trim_sites_to_retain=800, quick_match_strings=0, clearzone1=200, clearzone1b=260,
             clearzone1c=460, clearzonep=160, clearzone3=800, clearzone1e=258, clearzone_limit1e=40, print_secondary=0, min_align_ratio=0.56,
             cz1b_scale=0.97, cz1b_flat=1200.0, cz1c_scale=0.92, cz1c_flat=2600.0)
    d.update(kw)
    for k, v in d.items():
        c[k] = v
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def site_lists(ctx, op, lists, nss, read_off, cfg=None, basesP=None, basesM=None, d_ref=None, chrom_off=None):
    """Apply one policy to every read's list.  lists: SS_DTYPE[nreads, cap]; nss: int32[nreads].  Returns (lists, nss, READ_OUT_DTYPE[nreads])
    (copies; the inputs are not modified)."""
    L = _lib.load()
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError("no CUDA device visible: site_lists has no CPU fallback")
    cfg = policy_cfg() if cfg is None else cfg
    lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32).copy()
    n, cap = lists.shape
    ro = np.ascontiguousarray(read_off, np.int64)
    out = np.zeros(n, READ_OUT_DTYPE)
    bp = None if basesP is None else np.ascontiguousarray(basesP).view(np.int8)
    bm = None if basesM is None else np.ascontiguousarray(basesM).view(np.int8)
    co = None if chrom_off is None else np.ascontiguousarray(chrom_off, np.int64)
    _lib.check(L.bbm_sitelist_batch_host(ctx, op, _p(lists), _p(nss), n, cap, _p(ro), _p(bp), _p(bm), d_ref, _p(co), 0 if co is None else len(co) - 1,
                                         _p(cfg), _p(out)), "bbm_sitelist_batch_host")
    return lists, nss, out


SLOW_CFG_DTYPE = np.dtype([("paired", "<i4"), ("min_ratio", "<f4"), ("min_ratio_pre_rescue", "<f4"), ("clearzone1e", "<i4"), ("clearzone3", "<i4"),
                           ("slow_align_padding", "<i4"), ("extra_padding", "<i4"), ("expected_len_limit", "<i4")], align=True)
assert SLOW_CFG_DTYPE.itemsize == 32
SLOW_GAPPED, SLOW_ALIGNER_ERROR = 1, 2


def slow_cfg(**kw):
    """scoreSlow constants: MINIMUM_ALIGNMENT_SCORE_RATIO 0.56 (BBMap.java:50; the pre-rescue ratio only matters for pairs), CLEARZONE1e 258,
    CLEARZONE3 800, SLOW_ALIGN_PADDING 4 (BBMap.java:57), EXTRA_PADDING 10 (AbstractMapThread.java:2953), EXPECTED_LEN_LIMIT =
    (ALIGN_COLUMNS*17)/20-2*(SLOW_ALIGN_PADDING+10) = 2522 (AbstractMapThread.java:92)."""
    c = np.zeros(1, SLOW_CFG_DTYPE)
    d = dict(paired=0, min_ratio=0.56, min_ratio_pre_rescue=0.56, clearzone1e=258, clearzone3=800, slow_align_padding=4, extra_padding=10, expected_len_limit=2522)
    d.update(kw)
    for k, v in d.items():
        c[k] = v
    return c


def scoreSlow(ctx, lists, nss, read_off, basesP, basesM, d_ref, chrom_off, run, cfg=None):
    """BBMapThread.scoreSlow for every read with run[r] != 0, in rounds over the device aligner.  Returns (lists, status, alignments)."""
    L = _lib.load()
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError("no CUDA device visible: scoreSlow has no CPU fallback")
    cfg = slow_cfg() if cfg is None else cfg
    lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32)
    n, cap = lists.shape
    ro = np.ascontiguousarray(read_off, np.int64); co = np.ascontiguousarray(chrom_off, np.int64)
    bp = np.ascontiguousarray(basesP).view(np.int8); bm = np.ascontiguousarray(basesM).view(np.int8)
    rn = np.ascontiguousarray(run, np.int32); status = np.zeros(n, np.int32); na = C.c_int64(0)
    _lib.check(L.bbm_scoreslow_host(ctx, _p(lists), _p(nss), n, cap, _p(ro), _p(bp), _p(bm), d_ref, _p(co), len(co) - 1, _p(rn), _p(cfg), _p(status),
                                    C.byref(na)), "bbm_scoreslow_host")
    return lists, status, na.value


def findTipDeletions(ctx, lists, nss, read_off, basesP, basesM, quality, d_ref, chrom_off, cfg=None, chrom_min_index=None, device=0):
    """AbstractMapThread.findTipDeletions(Read, ...) on every list (staged through torch tensors; the entry point itself is device-resident).
    Returns (lists, READ_OUT_DTYPE[n])."""
    import torch
    from .rescue import tipdel_cfg
    L = _lib.load()
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError("no CUDA device visible: findTipDeletions has no CPU fallback")
    cfg = tipdel_cfg() if cfg is None else cfg
    lists = np.ascontiguousarray(lists, SS_DTYPE); n, cap = lists.shape
    dev = torch.device("cuda", device)
    up = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else np.asarray(a, dt)).view(np.uint8).reshape(-1).copy()).to(dev)
    q = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    d_l = up(lists); d_n = up(nss, np.int32); d_o = up(read_off, np.int64); d_p = up(np.concatenate([np.ascontiguousarray(basesP).view(np.uint8), np.zeros(16, np.uint8)]))
    d_m = up(np.concatenate([np.ascontiguousarray(basesM).view(np.uint8), np.zeros(16, np.uint8)]))
    d_q = None if quality is None else up(np.concatenate([np.ascontiguousarray(quality).view(np.uint8), np.zeros(16, np.uint8)]))
    d_c = up(chrom_off, np.int64); d_mi = None if chrom_min_index is None else up(chrom_min_index, np.int32)
    d_out = torch.zeros(max(n, 1) * READ_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    _lib.check(L.bbm_sitelist_tipdel_dev(ctx, q(d_l), q(d_n), n, cap, q(d_o), q(d_p), q(d_m), q(d_q), d_ref, q(d_c), q(d_mi), _p(cfg), q(d_out), None, None),
               "bbm_sitelist_tipdel_dev")
    torch.cuda.synchronize()
    return (np.frombuffer(d_l.cpu().numpy().tobytes(), SS_DTYPE).reshape(n, cap).copy(),
            np.frombuffer(d_out.cpu().numpy().tobytes(), READ_OUT_DTYPE)[:n].copy())


def removeOutOfBounds(ctx, lists, nss, read_off, chrom_max_index, scaf=None, inter_scaffold_padding=300, sam_out=1, expected_len_limit=2522, device=0):
    """AbstractMapThread.removeOutOfBounds on every list (staged through torch tensors).  scaf = (scaf_off, scaf_loc, scaf_len) as for
    bbmap_b200.sam.sam_batch, or None.  Returns (lists, nss, READ_OUT_DTYPE[n])."""
    import torch
    L = _lib.load()
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError("no CUDA device visible: removeOutOfBounds has no CPU fallback")
    lists = np.ascontiguousarray(lists, SS_DTYPE); n, cap = lists.shape
    dev = torch.device("cuda", device)
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dt)).view(np.uint8).reshape(-1).copy()).to(dev)
    q = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    d_l = torch.from_numpy(lists.view(np.uint8).reshape(-1).copy()).to(dev); d_n = up(nss, np.int32); d_o = up(read_off, np.int64); d_m = up(chrom_max_index, np.int32)
    d_so = None if scaf is None else up(scaf[0], np.int32); d_sl = None if scaf is None else up(scaf[1], np.int32)
    d_out = torch.zeros(max(n, 1) * READ_OUT_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    _lib.check(L.bbm_sitelist_bounds_dev(ctx, q(d_l), q(d_n), n, cap, q(d_o), q(d_m), q(d_so), q(d_sl), inter_scaffold_padding, sam_out, expected_len_limit,
                                         q(d_out), None), "bbm_sitelist_bounds_dev")
    torch.cuda.synchronize()
    return (np.frombuffer(d_l.cpu().numpy().tobytes(), SS_DTYPE).reshape(n, cap).copy(), np.frombuffer(d_n.cpu().numpy().tobytes(), np.int32)[:n].copy(),
            np.frombuffer(d_out.cpu().numpy().tobytes(), READ_OUT_DTYPE)[:n].copy())


def _staged(device):
    import torch
    dev = torch.device("cuda", device)
    up = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a if dt is None else np.asarray(a, dt)).view(np.uint8).reshape(-1).copy()).to(dev)
    q = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    return torch, dev, up, q


def applyClearzone3(ctx, lists, nss, read_off, flags, cfg=None, ambiguous_toss=False, device=0):
    """The clearzone-3 block and the final score gate of processRead (BBMapThread.java:667-684, 698-700; AbstractMapThread.applyClearzone3
    :1820-1870) on every list (staged through torch tensors).  flags: READ_OUT_DTYPE[n] as SL_FINAL returned it.  Returns (lists, nss,
    READ_OUT_DTYPE[n]) with flags updated, near_perfect = r.mapScore and best_sites = the amount subtracted."""
    L = _lib.load()
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError("no CUDA device visible: applyClearzone3 has no CPU fallback")
    torch, dev, up, q = _staged(device)
    cfg = policy_cfg() if cfg is None else cfg
    lists = np.ascontiguousarray(lists, SS_DTYPE); n, cap = lists.shape
    d_l = up(lists); d_n = up(nss, np.int32); d_o = up(read_off, np.int64)
    d_io = up(np.ascontiguousarray(flags, READ_OUT_DTYPE)) if n else torch.zeros(16, dtype=torch.uint8, device=dev)
    _lib.check(L.bbm_sitelist_clearzone3_dev(ctx, q(d_l), q(d_n), n, cap, q(d_o), _p(cfg), int(bool(ambiguous_toss)), q(d_io), None), "bbm_sitelist_clearzone3_dev")
    torch.cuda.synchronize()
    return (np.frombuffer(d_l.cpu().numpy().tobytes(), SS_DTYPE).reshape(n, cap).copy(), np.frombuffer(d_n.cpu().numpy().tobytes(), np.int32)[:n].copy(),
            np.frombuffer(d_io.cpu().numpy().tobytes(), READ_OUT_DTYPE)[:n].copy())


def tipScorePenalty(ctx, lists, nss, read_off, bases, match, match_off, flags, tiplen=7, device=0):
    """calcTipScorePenalty(r, maxSwScore, 7) + applyScorePenalty (BBMapThread.java:706-709; AbstractMapThread.java:2499-2567, 2601-2609) on every
    read (staged through torch tensors).  Returns (lists, penalty int32[n], status int32[n])."""
    L = _lib.load()
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError("no CUDA device visible: tipScorePenalty has no CPU fallback")
    torch, dev, up, q = _staged(device)
    lists = np.ascontiguousarray(lists, SS_DTYPE); n, cap = lists.shape
    d_l = up(lists); d_n = up(nss, np.int32); d_o = up(read_off, np.int64); d_mo = up(match_off, np.int64)
    d_b = up(np.concatenate([np.ascontiguousarray(bases).view(np.uint8), np.zeros(16, np.uint8)]))
    d_m = up(np.concatenate([np.ascontiguousarray(match).view(np.uint8), np.zeros(16, np.uint8)]))
    d_f = up(np.ascontiguousarray(flags, READ_OUT_DTYPE)) if n else torch.zeros(16, dtype=torch.uint8, device=dev)
    d_p = torch.zeros(max(n, 1), dtype=torch.int32, device=dev); d_s = torch.zeros(max(n, 1), dtype=torch.int32, device=dev)
    _lib.check(L.bbm_sitelist_tip_penalty_dev(ctx, q(d_l), q(d_n), n, cap, q(d_o), q(d_b), q(d_m), q(d_mo), q(d_f), tiplen, q(d_p), q(d_s), None),
               "bbm_sitelist_tip_penalty_dev")
    torch.cuda.synchronize()
    return np.frombuffer(d_l.cpu().numpy().tobytes(), SS_DTYPE).reshape(n, cap).copy(), d_p.cpu().numpy()[:n].copy(), d_s.cpu().numpy()[:n].copy()
