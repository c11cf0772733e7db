"""Host-side mirror of the batched mapper (bbm_map_batch_*, include/bbmap_cuda.h): BBMapThread.processRead / processReadPair
(current/align2/BBMapThread.java:389-733, 943-1362) over a batch of reads, from FASTQ-shaped bytes to SAM records.  Record layouts and the
reference's default switches; the compute is in libbbmapcuda.so (no CPU fallback)."""
import numpy as np

MAP_CFG_DTYPE = np.dtype([("paired", "<i4"), ("min_ratio", "<f4"), ("min_ratio_paired", "<f4"), ("min_ratio_pre_rescue", "<f4"),
                          ("secondary_site_score_ratio", "<f4"), ("slow_align_padding", "<i4"), ("max_indel", "<i4"), ("ambiguous_toss", "<i4"),
                          ("penalize_ambig", "<i4"), ("average_pair_dist", "<i4"), ("max_pair_dist", "<i4"), ("max_rescue_dist", "<i4"),
                          ("max_rescue_mismatches", "<i4"), ("do_rescue", "<i4"), ("kill_bad_pairs", "<i4"), ("require_correct_strands", "<i4"),
                          ("same_strand_pairs", "<i4"), ("match_slot", "<i4"), ("pad_", "<i4", (2,))], align=True)
MAP_REC_DTYPE = np.dtype([("chrom", "<i4"), ("start", "<i4"), ("stop", "<i4"), ("strand", "<i4"), ("map_score", "<i4"), ("flags", "<i4"),
                          ("match_len", "<i4"), ("cz3_sub", "<i4"), ("tip_penalty", "<i4"), ("status", "<i4"), ("pad_", "<i4", (2,))], align=True)
assert MAP_CFG_DTYPE.itemsize == 80 and MAP_REC_DTYPE.itemsize == 48
MF_MAPPED, MF_PERFECT, MF_AMBIGUOUS, MF_PAIRED, MF_RESCUED, MF_DISCARDED = 1, 2, 4, 8, 16, 32


def map_cfg(**kw):
    """BBMap defaults: MINIMUM_ALIGNMENT_SCORE_RATIO 0.56 (BBMap.java:50); the paired / pre-rescue ratios derived from it as
    AbstractMapThread.java:106-107 does (float arithmetic); SECONDARY_SITE_SCORE_RATIO .95 (:2970); SLOW_ALIGN_PADDING 4 (BBMap.java:57);
    MAX_INDEL 16000 (BBIndex.java:3170); ambig=best (no toss); PENALIZE_AMBIG true (:2959); INITIAL_AVERAGE_PAIR_DIST 100 (:2948);
    MAX_PAIR_DIST 32000, MAX_RESCUE_DIST 1200, MAX_RESCUE_MISMATCHES 32 (:2975-2977); DO_RESCUE true; KILL_BAD_PAIRS false,
    REQUIRE_CORRECT_STRANDS_PAIRS true, SAME_STRAND_PAIRS false (AbstractMapper.java:2681-2683)."""
    c = np.zeros(1, MAP_CFG_DTYPE)
    r = np.float32(kw.get("min_ratio", 0.56))
    one = np.float32(1)
    d = dict(paired=0, min_ratio=r,
             min_ratio_paired=max(np.float32(r * np.float32(.80)), np.float32(one - np.float32(np.float32(one - r) * np.float32(1.4)))),
             min_ratio_pre_rescue=max(np.float32(r * np.float32(.60)), np.float32(one - np.float32(np.float32(one - r) * np.float32(1.8)))),
             secondary_site_score_ratio=.95, slow_align_padding=4, max_indel=16000, ambiguous_toss=0, penalize_ambig=1, average_pair_dist=100,
             max_pair_dist=32000, max_rescue_dist=1200, max_rescue_mismatches=32, do_rescue=1, kill_bad_pairs=0, require_correct_strands=1,
             same_strand_pairs=0, match_slot=0)
    d.update(kw)
    for k, v in d.items():
        c[k] = v
    return c


# ---------------------------------------------------------------------------------------------------------------------------------
import ctypes as C

from . import lib as _lib
from .index import BBIndexCUDA, pack_chromosomes
from .keyring import SEED_CFG_DTYPE, default_cfg as seed_default_cfg
from .rescue import TIPDEL_CFG_DTYPE, tipdel_cfg
from .sam import SAM_CFG_DTYPE, SAM_OUT_DTYPE, default_cfg as sam_default_cfg, scaffold_table
from .sitelist import POLICY_CFG_DTYPE, SLOW_CFG_DTYPE, policy_cfg, slow_cfg

MAPPER_CFG_DTYPE = np.dtype([("map", MAP_CFG_DTYPE), ("seed", SEED_CFG_DTYPE), ("policy", POLICY_CFG_DTYPE), ("slow", SLOW_CFG_DTYPE), ("tip", TIPDEL_CFG_DTYPE),
                             ("sam", SAM_CFG_DTYPE), ("ingest_flags", "<i4"), ("max_keys", "<i4"), ("max_sites", "<i4"), ("sam_text", "<i4")], align=True)
MAP_STATS_DTYPE = np.dtype([("reads", "<i8"), ("mapped", "<i8"), ("slow_alignments", "<i8"), ("realign_fills", "<i8"), ("site_overflow_reads", "<i8"),
                            ("status_reads", "<i8"), ("sam_bytes", "<i8"), ("max_sites_used", "<i4"), ("genmatch_rounds", "<i4"), ("ms_total", "<f4"),
                            ("ms_seed_search", "<f4"), ("ms_lists", "<f4"), ("ms_slow", "<f4"), ("ms_genmatch", "<f4"), ("ms_sam", "<f4"), ("ms_rescue", "<f4"), ("pad_", "<i4"), ("rescue_scans", "<i8"),
                            ("rescue_fills", "<i8"), ("mated_pairs", "<i8"), ("inner_length_sum", "<i8")], align=True)
assert MAPPER_CFG_DTYPE.itemsize == 80 + 32 + 80 + 32 + 16 + 32 + 16 and MAP_STATS_DTYPE.itemsize == 128
ST_MATCH_OVERFLOW, ST_TIP, ST_SLOTS, ST_ALIGNER, ST_SITE_OVERFLOW, ST_SLOW = 1, 2, 4, 8, 16, 32


def mapper_cfg(paired=False, max_sites=16, max_keys=32, sam_text=False, **map_kw):
    """The reference's default switches for a whole run (BBMap.setDefaults, AbstractMapper / AbstractMapThread statics)."""
    c = np.zeros(1, MAPPER_CFG_DTYPE)
    c["map"] = map_cfg(paired=int(paired), **map_kw)
    c["seed"] = seed_default_cfg(); c["policy"] = policy_cfg(); c["tip"] = tipdel_cfg(); c["sam"] = sam_default_cfg()
    c["slow"] = slow_cfg(paired=int(paired), min_ratio=c["map"]["min_ratio"][0], min_ratio_pre_rescue=c["map"]["min_ratio_pre_rescue"][0])
    c["ingest_flags"] = 0; c["max_keys"] = max_keys; c["max_sites"] = max_sites; c["sam_text"] = int(sam_text)
    return c


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class BBMapCUDA:
    """The mapper of one GPU: reference + index resident (replicated per GPU), read batches in, SAM records out.  Mirrors what align2.BBMap does
    with its mapping threads (current/align2/BBMap.java:424-491, AbstractMapThread.run :387-584) for the path named in BASELINE.json."""

    def __init__(self, scaffolds, names=None, keylen=13, device=0):
        self.L = _lib.load()
        if self.L.bbm_device_count() <= 0:
            raise _lib.BbmError("no CUDA device visible: BBMapCUDA has no CPU fallback")
        self.cb, self.co, self.table = pack_chromosomes(scaffolds)
        self.index = BBIndexCUDA(self.cb, self.co, keylen=keylen, device=device)
        self.h = self.index.h
        so, sl_, sn = scaffold_table(self.table, len(self.co) - 1)
        self.scaf = (np.ascontiguousarray(so, np.int32), np.ascontiguousarray(sl_, np.int32), np.ascontiguousarray(sn, np.int32))
        nb = no = None
        if names is not None:
            order = sorted(range(len(self.table)), key=lambda i: self.table[i])
            enc = [names[i].encode() for i in order]
            no = np.zeros(len(enc) + 1, np.int64); np.cumsum([len(e) for e in enc], out=no[1:])
            nb = np.frombuffer(b"".join(enc) + b"\0", np.int8).copy()
        self._names = (nb, no); self._device = device
        self.L.bbm_map_set_scaffolds.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.check(self.L.bbm_map_set_scaffolds(self.h, _p(self.scaf[0]), _p(self.scaf[1]), _p(self.scaf[2]), len(self.co) - 1, _p(nb), _p(no)), "bbm_map_set_scaffolds")
        self.L.bbm_map_batch_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        self.L.bbm_map_batch_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                             C.c_void_p, C.c_void_p]

    def clone(self):
        """Another context on the same device that shares this mapper's resident index and reference (bbm_index_share): one per batch in flight."""
        o = BBMapCUDA.__new__(BBMapCUDA)
        o.L = self.L; o.cb, o.co, o.table, o.scaf = self.cb, self.co, self.table, self.scaf; o.index = None; o._names = self._names; o._parent = self
        h = C.c_void_p()
        _lib.check(self.L.bbm_init(self._device, C.byref(h)), "bbm_init")
        o.h = h; o._device = self._device
        _lib.check(self.L.bbm_index_share(o.h, self.h), "bbm_index_share")
        nb, no = self._names
        _lib.check(self.L.bbm_map_set_scaffolds(o.h, _p(self.scaf[0]), _p(self.scaf[1]), _p(self.scaf[2]), len(self.co) - 1, _p(nb), _p(no)), "bbm_map_set_scaffolds")
        return o

    def close(self):
        if self.index is not None:
            self.index.close()
        elif getattr(self, "h", None):
            self.L.bbm_destroy(self.h)
        self.h = None

    def map_batch(self, bases, quality, read_off, cfg=None, names=None, name_off=None, match_stride=None, sam_cap=0):
        """bases/quality: concatenated bytes (quality = phred values or None), read_off int64[n+1].  Returns dict(recs, sam, match, match_stride, stats
        [, sam_text, sam_off])."""
        cfg = mapper_cfg() if cfg is None else cfg
        ro = np.ascontiguousarray(read_off, np.int64); n = len(ro) - 1
        b = np.ascontiguousarray(bases).view(np.int8); q = None if quality is None else np.ascontiguousarray(quality).view(np.int8)
        if match_stride is None:
            match_stride = 2 * int(np.diff(ro).max() if n else 1) + 128
        recs = np.zeros(n, MAP_REC_DTYPE); sam = np.zeros(n, SAM_OUT_DTYPE)
        match = np.zeros(n * match_stride + 16, np.int8) if match_stride > 0 else None          # match_stride=0: records and SAM fields only
        stats = np.zeros(1, MAP_STATS_DTYPE)
        text = np.zeros(max(sam_cap, 1), np.int8) if cfg["sam_text"][0] else None
        toff = np.zeros(n + 1, np.int64) if cfg["sam_text"][0] else None
        nm = None if names is None else np.ascontiguousarray(names).view(np.int8); no = None if name_off is None else np.ascontiguousarray(name_off, np.int64)
        _lib.check(self.L.bbm_map_batch_host(self.h, _p(b), _p(q), _p(ro), n, _p(nm), _p(no), _p(cfg), _p(recs), _p(sam), _p(match), match_stride, _p(text), sam_cap,
                                             _p(toff), _p(stats)), "bbm_map_batch_host")
        r = {"recs": recs, "sam": sam, "match": match, "match_stride": match_stride, "stats": stats[0]}
        if text is not None:
            r["sam_text"] = text; r["sam_off"] = toff
        return r
