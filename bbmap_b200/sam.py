"""Host-side mirror of the reference's SAM record formation for the CUDA path: the fields SamLine(Read, int) derives from a mapped
read (current/stream/SamLine.java:82-330) — FLAG (makeFlag :2134-2151), POS/PNEXT/TLEN in scaffold coordinates, MAPQ (toMapq
:1709-1723) and the CIGAR text (toCigar13/toCigar14 :600-750) — batched."""
import ctypes as C

import numpy as np

from . import lib as _lib

SAM_TASK_DTYPE = np.dtype([("match_off", "<i8"), ("match_len", "<i4"), ("chrom", "<i4"), ("start", "<i4"), ("stop", "<i4"), ("read_len", "<i4"),
                           ("score", "<i4"), ("mate", "<i4"), ("flags", "<i4"), ("pad_", "<i4")], align=True)
SAM_OUT_DTYPE = np.dtype([("flag", "<i4"), ("pos", "<i4"), ("mapq", "<i4"), ("scaffold", "<i4"), ("rnext", "<i4"), ("pnext", "<i4"), ("tlen", "<i4"),
                          ("cigar_len", "<i4")], align=True)
SAM_CFG_DTYPE = np.dtype([("version14", "<i4"), ("soft_clip", "<i4"), ("intron_limit", "<i4"), ("penalize_ambig", "<i4"),
                          ("inter_scaffold_padding", "<i4"), ("pad_", "<i4", (3,))], align=True)
assert SAM_TASK_DTYPE.itemsize == 48 and SAM_OUT_DTYPE.itemsize == 32 and SAM_CFG_DTYPE.itemsize == 32
RF_MAPPED, RF_MINUS, RF_PERFECT, RF_AMBIGUOUS, RF_SECONDARY, RF_DISCARDED, RF_PAIRED, RF_PAIRNUM1 = 1, 2, 4, 8, 16, 32, 64, 128


def default_cfg(version=1.4):
    """SamLine.VERSION / SOFT_CLIP / INTRON_LIMIT / PENALIZE_AMBIG defaults (SamLine.java:2424-2434); MID_PADDING 300."""
    c = np.zeros(1, SAM_CFG_DTYPE)
    c[0] = (1 if version > 1.3 else 0, 1, 2 ** 31 - 1, 1, 300, (0, 0, 0))
    return c


def scaffold_table(table, nchroms):
    """(scaf_off, scaf_loc, scaf_len) from bbmap_b200.index.pack_chromosomes' scaffold table [(chrom, start, length)]."""
    t = sorted(table)
    off = np.zeros(nchroms + 1, np.int32)
    for ch, _, _ in t:
        off[ch] += 1
    off = np.cumsum(off).astype(np.int32)
    return off, np.array([x[1] for x in t], np.int32), np.array([x[2] for x in t], np.int32)


def cigar_offsets(tasks):
    cap = 2 * tasks["match_len"].astype(np.int64) + 12
    off = np.zeros(len(tasks) + 1, np.int64)
    np.cumsum(cap, out=off[1:])
    return off


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def sam_batch(ctx, tasks, match_buf, scaf, cfg=None):
    """ctx: a bbm_ctx handle.  Returns (outs SAM_OUT_DTYPE[n], cigar_buf int8[], cigar_off int64[n+1])."""
    L = _lib.load()
    if L.bbm_device_count() <= 0:
        raise _lib.BbmError("no CUDA device visible: sam_batch has no CPU fallback")
    cfg = default_cfg() if cfg is None else cfg
    tasks = np.ascontiguousarray(tasks, SAM_TASK_DTYPE)
    mb = np.ascontiguousarray(match_buf).view(np.int8)
    so, sl, sn = (np.ascontiguousarray(x, np.int32) for x in scaf)
    coff = cigar_offsets(tasks)
    outs = np.zeros(len(tasks), SAM_OUT_DTYPE); cbuf = np.zeros(max(int(coff[-1]), 1), np.int8)
    _lib.check(L.bbm_sam_batch_host(ctx, _p(tasks), len(tasks), _p(mb), mb.size, _p(so), _p(sl), _p(sn), len(so) - 1, _p(cfg), _p(outs), _p(cbuf), _p(coff)),
               "bbm_sam_batch_host")
    return outs, cbuf, coff


def tasks_from_lists(lists, nss, read_off, flags, match_off=None):
    """Read.setFromTopSite for unpaired reads (current/stream/Read.java:1171-1190, 1213-1224; clearSite :1278-1286) in numpy — what
    bbm_sam_tasks_from_lists_dev builds on the device.  flags: sitelist.READ_OUT_DTYPE[n]."""
    n = len(nss)
    t = np.zeros(n, SAM_TASK_DTYPE)
    top = lists[:, 0]; f = flags["flags"]
    m = (np.asarray(nss) > 0) & ((f & 1) != 0)
    t["read_len"] = np.diff(np.asarray(read_off, np.int64)); t["mate"] = -1
    if match_off is not None:
        t["match_off"] = np.asarray(match_off, np.int64)[:-1]; t["match_len"] = np.where(m, np.diff(np.asarray(match_off, np.int64)), 0)
    for k in ("chrom", "start", "stop"):
        t[k] = np.where(m, top[k], -1)
    t["score"] = np.where(m, top["slow_score"], 0)
    t["flags"] = np.where(m, RF_MAPPED | np.where(top["strand"] != 0, RF_MINUS, 0) | np.where(top["perfect"] != 0, RF_PERFECT, 0), 0) | np.where((f & 4) != 0, RF_AMBIGUOUS, 0)
    return t
