"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the CPU oracle.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (bbmap_b200/) never does.

Two libraries:
  oracle/liborc.so          the "port": a C restatement of the reference's algorithm
  oracle/_ref/libbbref.so   the reference's own C (jni/*.c) compiled unmodified from
                            /root/reference against a stand-in jni.h; present only if it
                            was built in the dev container (it travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
TABLE_LEN = 604

# task / out records — byte-compatible with include/bbmap_cuda.h (bbm_msa_task / bbm_msa_out)
TASK_DTYPE = np.dtype([("read_off", "<i8"), ("ref_off", "<i8"), ("read_len", "<i4"), ("ref_len", "<i4"),
                       ("ref_start", "<i4"), ("ref_end", "<i4"), ("min_score", "<i4"), ("flags", "<i4")], align=True)
OUT_DTYPE = np.dtype([("result", "<i4", (5,)), ("path", "<i4"), ("iterations", "<i8"), ("score", "<i4", (8,)),
                      ("score_len", "<i4"), ("match_len", "<i4"), ("status", "<i4"), ("pad_", "<i4")], align=True)
assert TASK_DTYPE.itemsize == 40 and OUT_DTYPE.itemsize == 80

TF_RAW_LIMITED, TF_RAW_UNLIMITED, TF_CLAMP, TF_SCORE, TF_TRACEBACK = 1, 2, 4, 8, 16

BAND_TASK_DTYPE = np.dtype([("query_off", "<i8"), ("ref_off", "<i8"), ("query_len", "<i4"), ("ref_len", "<i4"), ("qstart", "<i4"),
                            ("rstart", "<i4"), ("max_edits", "<i4"), ("max_width", "<i4"), ("exact", "<i4"), ("dir", "<i4")], align=True)
BAND_OUT_DTYPE = np.dtype([("edits", "<i4"), ("rv", "<i4", (5,)), ("status", "<i4"), ("pad_", "<i4")], align=True)
assert BAND_TASK_DTYPE.itemsize == 48 and BAND_OUT_DTYPE.itemsize == 32
DIR_FORWARD, DIR_FORWARD_RC, DIR_REVERSE, DIR_REVERSE_RC = 0, 1, 2, 3


def build(quiet=True):
    """(Re)build liborc.so and, when /root/reference is mounted, _ref/libbbref.so."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self):
        so = os.environ.get("ORC_SO", os.path.join(HERE, "liborc.so"))
        if not os.path.exists(so):
            build()
        self.lib = C.CDLL(so)
        L = self.lib
        L.orc_msa_new.restype = C.c_void_p
        L.orc_msa_new.argtypes = [C.c_int, C.c_int]
        L.orc_msa_free.argtypes = [C.c_void_p]
        L.orc_msa_packed.restype = C.c_void_p
        L.orc_msa_packed.argtypes = [C.c_void_p]
        L.orc_msa_iterations.restype = C.c_int64
        L.orc_msa_iterations.argtypes = [C.c_void_p, C.c_int]
        L.orc_msa_set_band.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.orc_msa_set_backend.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_msa_set_shape.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_msa_last_path.argtypes = [C.c_void_p]
        L.orc_batch_run.restype = C.c_int64
        L.orc_set_reference_fns.argtypes = [C.c_void_p, C.c_void_p]
        self.ref = None
        refso = os.path.join(HERE, "_ref", "libbbref.so")
        if os.path.exists(refso):
            self.ref = C.CDLL(refso)
            fl = C.cast(self.ref.fillLimitedX, C.c_void_p).value
            fu = C.cast(self.ref.fillUnlimited, C.c_void_p).value
            L.orc_set_reference_fns(fl, fu)
            L.orc_banded_set_reference_fns.argtypes = [C.c_void_p] * 4
            L.orc_banded_set_reference_fns(*[C.cast(getattr(self.ref, n), C.c_void_p).value
                                             for n in ("alignForward", "alignForwardRC", "alignReverse", "alignReverseRC")])
        self.sub = np.zeros(TABLE_LEN, np.int32)
        self.ins = np.zeros(TABLE_LEN, np.int32)
        self.insC = np.zeros(TABLE_LEN, np.int32)
        L.orc_msa_tables(_p(self.sub), _p(self.ins), _p(self.insC), None, None, None)
        self.b2n = np.zeros(128, np.int8)
        L.orc_base_to_number(_p(self.b2n))

    @property
    def has_reference(self):
        return self.ref is not None

    # ---------------- raw fills on a caller-owned packed matrix ----------------
    def new_packed(self, maxRows, maxColumns):
        """A fresh `packed` initialised like the Java constructor (MSA11tsJNI.java:98-112)."""
        m = self.lib.orc_msa_new(maxRows, maxColumns)
        n = 3 * (maxRows + 1) * (maxColumns + 1)
        arr = np.ctypeslib.as_array(C.cast(self.lib.orc_msa_packed(m), C.POINTER(C.c_int32)), shape=(n,)).copy()
        self.lib.orc_msa_free(m)
        return arr

    def _fn(self, name, kind):
        if kind == "reference":
            if not self.ref:
                raise RuntimeError("oracle/_ref/libbbref.so not built")
            return getattr(self.ref, {"limited": "fillLimitedX", "unlimited": "fillUnlimited"}[name])
        return getattr(self.lib, {"limited": "orc_fill_limitedX", "unlimited": "orc_fill_unlimited"}[name])

    def fill_limited(self, read, ref, a, b, minScore, packed, maxRows, maxColumns, bandwidth=0, ratio=0.0, kind="port",
                     vl=None, hl=None):
        read = np.ascontiguousarray(read, np.int8); ref = np.ascontiguousarray(ref, np.int8)
        res = np.zeros(5, np.int32); it = np.zeros(1, np.int64)
        vl = np.zeros(maxRows + 1, np.int32) if vl is None else vl
        hl = np.zeros(maxColumns + 1, np.int32) if hl is None else hl
        f = self._fn("limited", kind); f.restype = None
        f(_p(read), _p(ref), C.c_int(len(read)), C.c_int(len(ref)), C.c_int(a), C.c_int(b), C.c_int(minScore), _p(res), _p(it),
          _p(packed), _p(self.sub), _p(self.ins), C.c_int(maxRows), C.c_int(maxColumns), C.c_int(bandwidth), C.c_float(ratio),
          _p(vl), _p(hl), _p(self.b2n), _p(self.insC))
        return res, int(it[0])

    def fill_unlimited(self, read, ref, a, b, packed, maxRows, maxColumns, kind="port"):
        read = np.ascontiguousarray(read, np.int8); ref = np.ascontiguousarray(ref, np.int8)
        res = np.zeros(4, np.int32); it = np.zeros(1, np.int64)
        f = self._fn("unlimited", kind); f.restype = None
        f(_p(read), _p(ref), C.c_int(len(read)), C.c_int(len(ref)), C.c_int(a), C.c_int(b), _p(res), _p(it),
          _p(packed), _p(self.sub), _p(self.ins), C.c_int(maxRows), C.c_int(maxColumns))
        return res, int(it[0])

    # ---------------- index build + analysis ----------------
    def index_build(self, chrom_bytes, chrom_off, keylen, chrombits=-1):
        """Returns (cfg structured array[1], [(starts, sites) per block], COUNTS, hist1001)."""
        from bbmap_b200.index import INDEX_CFG_DTYPE
        L = self.lib
        b = np.ascontiguousarray(chrom_bytes).view(np.int8); off = np.ascontiguousarray(chrom_off, np.int64)
        nch = len(off) - 1
        if chrombits < 0:
            chrombits = L.orc_auto_chrombits(_p(off), C.c_int(nch))
        ndef = int(np.isin(b.view(np.uint8), np.frombuffer(b"ACGTUacgtu", np.uint8)).sum())
        cfg = np.zeros(1, INDEX_CFG_DTYPE)
        L.orc_index_cfg_init(_p(cfg), C.c_int(keylen), C.c_int(chrombits), C.c_int64(ndef))
        cpb = int(cfg["chroms_per_block"][0]); low = cpb - 1
        ks = 1 << (2 * keylen)
        blocks = []
        chrom = 1
        L.orc_index_build_block.restype = C.c_int64
        while chrom <= nch:
            a = max(1, chrom & ~low); bmax = min(nch, (chrom & ~low) + cpb - 1)
            starts = np.zeros(ks + 1, np.int32)
            ptr = C.POINTER(C.c_int32)()
            n = L.orc_index_build_block(_p(b), _p(off), C.c_int(a), C.c_int(bmax), _p(cfg), _p(starts), C.byref(ptr))
            sites = np.ctypeslib.as_array(ptr, shape=(max(int(n), 1),)).copy()[: int(n)]
            blocks.append((starts, sites))
            chrom = bmax + 1
        counts = np.zeros(ks, np.int32); hist = np.zeros(1001, np.int32)
        sp = (C.c_void_p * len(blocks))(*[s.ctypes.data for s, _ in blocks])
        tp = (C.c_void_p * len(blocks))(*[(t if len(t) else np.zeros(1, np.int32)).ctypes.data for _, t in blocks])
        L.orc_index_analyze(C.c_int(len(blocks)), sp, tp, _p(cfg), _p(counts), _p(hist))
        return cfg, blocks, counts, hist

    # ---------------- index search (BBIndex.find) ----------------
    def search_batch(self, index, chrom_bytes, chrom_off, bases, baseScores, read_off, seeds, quit_after_two_perfects=True):
        """index = (cfg, blocks, counts, hist) from index_build; seeds = dict(nkeys, offsets, keyScores) from seed_batch.
        Returns a SEARCH_RESULT_DTYPE array."""
        from bbmap_b200.search import SEARCH_RESULT_DTYPE
        cfg, blocks, counts, hist = index
        L = self.lib

        class Blk(C.Structure):
            _fields_ = [("starts", C.c_void_p), ("sites", C.c_void_p)]

        class Idx(C.Structure):
            _fields_ = [("cfg", C.c_void_p), ("blocks", C.c_void_p), ("nblocks", C.c_int32), ("nchroms", C.c_int32),
                        ("counts", C.c_void_p), ("hist", C.c_void_p), ("chroms", C.c_void_p), ("chrom_off", C.c_void_p)]
        keep = [(np.ascontiguousarray(s_), np.ascontiguousarray(t_ if len(t_) else np.zeros(1, np.int32))) for s_, t_ in blocks]
        barr = (Blk * len(keep))(*[Blk(a.ctypes.data, b.ctypes.data) for a, b in keep])
        cb = np.ascontiguousarray(chrom_bytes).view(np.int8); co = np.ascontiguousarray(chrom_off, np.int64)
        X = Idx(cfg.ctypes.data, C.addressof(barr), len(keep), len(co) - 1, counts.ctypes.data, hist.ctypes.data, cb.ctypes.data, co.ctypes.data)
        bases = np.ascontiguousarray(bases).view(np.int8); bs = np.ascontiguousarray(baseScores).view(np.int8)
        ro = np.ascontiguousarray(read_off, np.int64)
        n = len(ro) - 1
        res = np.zeros(n, SEARCH_RESULT_DTYPE)
        nk = np.ascontiguousarray(seeds["nkeys"], np.int32); of = np.ascontiguousarray(seeds["offsets"], np.int32); ks = np.ascontiguousarray(seeds["keyScores"], np.int32)
        L.orc_search_batch.restype = None
        L.orc_search_batch(C.byref(X), _p(bases), _p(bs), _p(ro), C.c_int64(n), _p(nk), _p(of), _p(ks), C.c_int32(of.shape[1]),
                           C.c_int(1 if quit_after_two_perfects else 0), _p(res))
        return res

    # ---------------- gapped alignment (makeGref) ----------------
    def fill_and_score_limited_gapped(self, read, ref, refStart, refEnd, minScore, gaps, traceback=True, maxRows=601, maxColumns=3000):
        """MSA.fillAndScoreLimited(read, ref, refStart, refEnd, minScore, gaps) then msa.traceback(…, gapped) as
        BBMapThread.scoreSlow does (BBMapThread.java:306-356).  Returns (score8 list or None, match bytes or None, max4)."""
        L = self.lib
        L.orc_msa_new.restype = C.c_void_p
        m = C.c_void_p(L.orc_msa_new(C.c_int32(maxRows), C.c_int32(maxColumns)))
        try:
            read = np.ascontiguousarray(read).view(np.int8); ref = np.ascontiguousarray(ref).view(np.int8)
            g = None if gaps is None or len(gaps) == 0 else np.ascontiguousarray(gaps, np.int32).copy()
            max4 = np.zeros(4, np.int32); out8 = np.zeros(8, np.int32)
            n = L.orc_msa_fillAndScoreLimited(m, _p(read), C.c_int32(len(read)), _p(ref), C.c_int32(len(ref)), C.c_int32(refStart), C.c_int32(refEnd),
                                              C.c_int32(minScore), None if g is None else _p(g), C.c_int32(0 if g is None else len(g)), _p(max4), _p(out8))
            if n <= 0:
                return None, None, max4
            match = None
            if traceback:
                a, b = max(0, refStart), min(len(ref) - 1, refEnd)
                buf = np.zeros(len(read) + 3002 + 128 * 64, np.int8)
                k = L.orc_msa_traceback(m, _p(read), _p(ref), C.c_int32(a), C.c_int32(b), C.c_int32(int(max4[0])), C.c_int32(int(max4[1])),
                                        C.c_int32(int(max4[2])), C.c_int(0 if g is None else 1), _p(buf), C.c_int32(len(buf)))
                match = buf[:k].copy()
            return out8[:n].tolist(), match, max4
        finally:
            L.orc_msa_free(m)

    # ---------------- SAM record fields (SamLine) ----------------
    def sam_batch(self, tasks, match_buf, scaf, cfg):
        from bbmap_b200.sam import SAM_OUT_DTYPE, SAM_TASK_DTYPE, cigar_offsets
        tasks = np.ascontiguousarray(tasks, SAM_TASK_DTYPE); mb = np.ascontiguousarray(match_buf).view(np.int8)
        so, sl, sn = (np.ascontiguousarray(x, np.int32) for x in scaf)
        coff = cigar_offsets(tasks)
        outs = np.zeros(len(tasks), SAM_OUT_DTYPE); cbuf = np.zeros(max(int(coff[-1]), 1), np.int8)
        self.lib.orc_sam_batch.restype = None
        self.lib.orc_sam_batch(_p(tasks), C.c_int64(len(tasks)), _p(mb), _p(so), _p(sl), _p(sn), C.c_int32(len(so) - 1), _p(cfg), _p(outs), _p(cbuf), _p(coff))
        return outs, cbuf, coff

    # ---------------- tip-deletion search / mate rescue scans ----------------
    def tipdel_batch(self, reads, refs, tasks, cfg):
        from bbmap_b200.rescue import TIPDEL_OUT_DTYPE, TIPDEL_TASK_DTYPE
        reads = np.ascontiguousarray(reads).view(np.int8); refs = np.ascontiguousarray(refs).view(np.int8)
        tasks = np.ascontiguousarray(tasks, TIPDEL_TASK_DTYPE); outs = np.zeros(len(tasks), TIPDEL_OUT_DTYPE)
        self.lib.orc_tipdel_batch.restype = None
        self.lib.orc_tipdel_batch(_p(reads), _p(refs), _p(tasks), C.c_int64(len(tasks)), _p(cfg), _p(outs))
        return outs

    def rescue_batch(self, reads, refs, tasks, cfg):
        from bbmap_b200.rescue import RESCUE_OUT_DTYPE, RESCUE_TASK_DTYPE
        reads = np.ascontiguousarray(reads).view(np.int8); refs = np.ascontiguousarray(refs).view(np.int8)
        tasks = np.ascontiguousarray(tasks, RESCUE_TASK_DTYPE); outs = np.zeros(len(tasks), RESCUE_OUT_DTYPE)
        self.lib.orc_rescue_batch.restype = None
        self.lib.orc_rescue_batch(_p(reads), _p(refs), _p(tasks), C.c_int64(len(tasks)), _p(cfg), _p(outs))
        return outs

    def find_tip_deletions_right(self, bases, ref, min_index, original_stop, search_dist, tiplen):
        b = np.ascontiguousarray(bases).view(np.int8); r = np.ascontiguousarray(ref).view(np.int8)
        self.lib.orc_find_tip_deletions_right.restype = C.c_int
        return self.lib.orc_find_tip_deletions_right(_p(b), C.c_int(len(b)), _p(r), C.c_int(len(r)), C.c_int(min_index), C.c_int(original_stop),
                                                     C.c_int(search_dist), C.c_int(tiplen))

    def find_tip_deletions_left(self, bases, ref, min_index, original_start, search_dist, tiplen):
        b = np.ascontiguousarray(bases).view(np.int8); r = np.ascontiguousarray(ref).view(np.int8)
        self.lib.orc_find_tip_deletions_left.restype = C.c_int
        return self.lib.orc_find_tip_deletions_left(_p(b), C.c_int(len(b)), _p(r), C.c_int(len(r)), C.c_int(min_index), C.c_int(original_start),
                                                    C.c_int(search_dist), C.c_int(tiplen))

    # ---------------- per-read site-list policies ----------------
    def sitelist(self, op, lists, nss, read_off, cfg, basesP=None, basesM=None, refs=None, chrom_off=None):
        from bbmap_b200.sitelist import READ_OUT_DTYPE, SL_FINAL, SL_NOINDEL, SL_TRIM, SS_DTYPE
        lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32).copy()
        n, cap = lists.shape
        ro = np.ascontiguousarray(read_off, np.int64); rl = np.ascontiguousarray(np.diff(ro), np.int32)
        out = np.zeros(n, READ_OUT_DTYPE)
        if op == SL_TRIM:
            self.lib.orc_sitelist_trim.restype = None
            self.lib.orc_sitelist_trim(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(rl), _p(cfg), _p(out))
        elif op == SL_NOINDEL:
            bp = np.ascontiguousarray(basesP).view(np.int8); bm = np.ascontiguousarray(basesM).view(np.int8)
            rf = np.ascontiguousarray(refs).view(np.int8); co = np.ascontiguousarray(chrom_off, np.int64)
            self.lib.orc_sitelist_noindel.restype = None
            self.lib.orc_sitelist_noindel(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(bp), _p(bm), _p(ro), _p(rf), _p(co), _p(cfg), _p(out))
        else:
            assert op == SL_FINAL
            self.lib.orc_sitelist_final.restype = None
            self.lib.orc_sitelist_final(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(rl), _p(cfg), _p(out))
        return lists, nss, out

    def sitelist_tipdel(self, lists, nss, read_off, basesP, basesM, quality, refs, chrom_off, cfg, chrom_min_index=None):
        from bbmap_b200.sitelist import READ_OUT_DTYPE, SS_DTYPE
        lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32)
        n, cap = lists.shape
        ro = np.ascontiguousarray(read_off, np.int64); co = np.ascontiguousarray(chrom_off, np.int64)
        bp = np.ascontiguousarray(basesP).view(np.int8); bm = np.ascontiguousarray(basesM).view(np.int8); rf = np.ascontiguousarray(refs).view(np.int8)
        qq = None if quality is None else np.ascontiguousarray(quality).view(np.int8)
        mi = None if chrom_min_index is None else np.ascontiguousarray(chrom_min_index, np.int32)
        out = np.zeros(n, READ_OUT_DTYPE)
        self.lib.orc_sitelist_tipdel.restype = None
        self.lib.orc_sitelist_tipdel(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(bp), _p(bm), None if qq is None else _p(qq), _p(ro), _p(rf), _p(co),
                                     None if mi is None else _p(mi), _p(cfg), _p(out))
        return lists, out

    def sitelist_bounds(self, lists, nss, read_off, chrom_max_index, scaf=None, inter_scaffold_padding=300, sam_out=1, expected_len_limit=2522):
        from bbmap_b200.sitelist import READ_OUT_DTYPE, SS_DTYPE
        lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32).copy()
        n, cap = lists.shape
        rl = np.ascontiguousarray(np.diff(np.ascontiguousarray(read_off, np.int64)), np.int32)
        mi = np.ascontiguousarray(chrom_max_index, np.int32)
        so = None if scaf is None else np.ascontiguousarray(scaf[0], np.int32); sl_ = None if scaf is None else np.ascontiguousarray(scaf[1], np.int32)
        out = np.zeros(n, READ_OUT_DTYPE)
        self.lib.orc_sitelist_bounds.restype = None
        self.lib.orc_sitelist_bounds(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(rl), _p(mi), None if so is None else _p(so), None if sl_ is None else _p(sl_),
                                     C.c_int32(inter_scaffold_padding), C.c_int32(sam_out), C.c_int32(expected_len_limit), _p(out))
        return lists, nss, out

    def sitelist_clearzone3(self, lists, nss, read_off, flags, cfg, ambiguous_toss=False):
        from bbmap_b200.sitelist import READ_OUT_DTYPE, SS_DTYPE
        lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32).copy()
        n, cap = lists.shape
        rl = np.ascontiguousarray(np.diff(np.ascontiguousarray(read_off, np.int64)), np.int32)
        io = np.ascontiguousarray(flags, READ_OUT_DTYPE).copy()
        self.lib.orc_sitelist_clearzone3.restype = None
        self.lib.orc_sitelist_clearzone3(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(rl), _p(cfg), C.c_int32(int(bool(ambiguous_toss))), _p(io))
        return lists, nss, io

    def sitelist_tip_penalty(self, lists, nss, read_off, bases, match, match_off, flags, tiplen=7):
        from bbmap_b200.sitelist import READ_OUT_DTYPE, SS_DTYPE
        lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32)
        n, cap = lists.shape
        ro = np.ascontiguousarray(read_off, np.int64); mo = np.ascontiguousarray(match_off, np.int64)
        bb = np.concatenate([np.ascontiguousarray(bases).view(np.int8), np.zeros(16, np.int8)])
        mm = np.concatenate([np.ascontiguousarray(match).view(np.int8), np.zeros(16, np.int8)])
        ff = np.ascontiguousarray(flags, READ_OUT_DTYPE)
        pen = np.zeros(n, np.int32); st = np.zeros(n, np.int32)
        self.lib.orc_sitelist_tip_penalty.restype = None
        self.lib.orc_sitelist_tip_penalty(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(ro), _p(bb), _p(mm), _p(mo), _p(ff), C.c_int32(tiplen), _p(pen), _p(st))
        return lists, pen, st

    def score_slow(self, lists, nss, read_off, basesP, basesM, refs, chrom_off, run, cfg):
        from bbmap_b200.sitelist import SS_DTYPE
        lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32)
        n, cap = lists.shape
        ro = np.ascontiguousarray(read_off, np.int64); co = np.ascontiguousarray(chrom_off, np.int64)
        bp = np.ascontiguousarray(basesP).view(np.int8); bm = np.ascontiguousarray(basesM).view(np.int8); rf = np.ascontiguousarray(refs).view(np.int8)
        rn = np.ascontiguousarray(run, np.int32); status = np.zeros(n, np.int32)
        self.lib.orc_score_slow.restype = C.c_int64
        na = self.lib.orc_score_slow(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(bp), _p(bm), _p(ro), _p(rf), _p(co), _p(rn), _p(cfg), _p(status))
        return lists, status, na

    def map_finish_single(self, lists, nss, read_off, basesP, basesM, refs, chrom_off, flags, pcfg, mcfg, match_stride):
        """Tail of processRead (match string of the primary site ... tip penalty) for every read; see mapper_oracle.c."""
        from bbmap_b200.sitelist import READ_OUT_DTYPE, SS_DTYPE
        from bbmap_b200.mapper import MAP_REC_DTYPE
        lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32).copy()
        n, cap = lists.shape
        ro = np.ascontiguousarray(read_off, np.int64); co = np.ascontiguousarray(chrom_off, np.int64)
        bp = np.ascontiguousarray(basesP).view(np.int8); bm = np.ascontiguousarray(basesM).view(np.int8); rf = np.ascontiguousarray(refs).view(np.int8)
        ff = np.ascontiguousarray(flags, READ_OUT_DTYPE)
        recs = np.zeros(n, MAP_REC_DTYPE); match = np.zeros(n * match_stride + 16, np.int8)
        self.lib.orc_map_finish_single.restype = C.c_int64
        fills = self.lib.orc_map_finish_single(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(bp), _p(bm), _p(ro), _p(rf), _p(co), _p(pcfg), _p(mcfg),
                                               _p(ff), _p(recs), _p(match), C.c_int64(match_stride))
        return lists, nss, recs, match, fills

    def map_pairs(self, lists, nss, read_off, basesP, basesM, quality, refs, chrom_off, nkeys, pcfg, mcfg, scfg, tcfg, match_stride):
        """processReadPair after quickMap for every pair (reads 2i / 2i+1); see mapper_oracle.c."""
        from bbmap_b200.sitelist import SS_DTYPE
        from bbmap_b200.mapper import MAP_REC_DTYPE
        lists = np.ascontiguousarray(lists, SS_DTYPE).copy(); nss = np.ascontiguousarray(nss, np.int32).copy()
        n, cap = lists.shape
        ro = np.ascontiguousarray(read_off, np.int64); co = np.ascontiguousarray(chrom_off, np.int64)
        bp = np.ascontiguousarray(basesP).view(np.int8); bm = np.ascontiguousarray(basesM).view(np.int8); rf = np.ascontiguousarray(refs).view(np.int8)
        q = None if quality is None else np.ascontiguousarray(quality).view(np.int8)
        nk = np.ascontiguousarray(nkeys, np.int32)
        recs = np.zeros(n, MAP_REC_DTYPE); match = np.zeros(n * match_stride + 16, np.int8); stats = np.zeros(8, np.int64)
        self.lib.orc_map_pairs.restype = C.c_int64
        self.lib.orc_map_pairs(_p(lists), _p(nss), C.c_int64(n), C.c_int32(cap), _p(bp), _p(bm), None if q is None else _p(q), _p(ro), _p(rf), _p(co), _p(nk), _p(pcfg),
                               _p(mcfg), _p(scfg), _p(tcfg), _p(recs), _p(match), C.c_int64(match_stride), _p(stats))
        return lists, nss, recs, match, stats

    # ---------------- scoreNoIndels ----------------
    def noindel_batch(self, reads, refs, tasks, match_off=None):
        reads = np.ascontiguousarray(reads).view(np.int8); refs = np.ascontiguousarray(refs).view(np.int8)
        tasks = np.ascontiguousarray(tasks)
        scores = np.zeros(len(tasks), np.int32)
        mbuf = np.zeros(int(match_off[-1]) if match_off is not None else 1, np.int8)
        self.lib.orc_noindel_batch.restype = None
        self.lib.orc_noindel_batch(_p(reads), _p(refs), _p(tasks), _p(scores), _p(mbuf) if match_off is not None else None,
                                   _p(np.ascontiguousarray(match_off, np.int64)) if match_off is not None else None, C.c_int64(len(tasks)))
        return scores, mbuf

    # ---------------- read ingest (Read.validate + reverse complement) ----------------
    def ingest_batch(self, bases, quality, read_off, flags=0):
        """Returns (bases, quality, basesM, read_flags) after Read.validate / reverseComplementBases; inputs are not modified."""
        b = np.ascontiguousarray(bases).view(np.int8).copy()
        q = None if quality is None else np.ascontiguousarray(quality).view(np.int8).copy()
        ro = np.ascontiguousarray(read_off, np.int64); n = len(ro) - 1
        bm = np.zeros(len(b), np.int8); fl = np.zeros(n, np.int32)
        self.lib.orc_ingest_batch.restype = None
        self.lib.orc_ingest_batch(_p(b), None if q is None else _p(q), _p(ro), C.c_int64(n), C.c_int(flags), _p(bm), _p(fl))
        return b, q, bm, fl

    # ---------------- KeyRing seeding ----------------
    def seed_batch(self, bases, quality, read_off, cfg, maxKeys=96):
        bases = np.ascontiguousarray(bases).view(np.int8)
        quality = None if quality is None else np.ascontiguousarray(quality).view(np.int8)
        read_off = np.ascontiguousarray(read_off, np.int64)
        n = len(read_off) - 1
        out = dict(nkeys=np.zeros(n, np.int32), offsets=np.zeros((n, maxKeys), np.int32), keys=np.zeros((n, maxKeys), np.int32),
                   keyScores=np.zeros((n, maxKeys), np.int32), baseScores=np.zeros(len(bases), np.int8))
        self.lib.orc_seed_batch.restype = None
        self.lib.orc_seed_batch(_p(bases), None if quality is None else _p(quality), _p(read_off), C.c_int64(n), _p(cfg), C.c_int32(maxKeys),
                                _p(out["nkeys"]), _p(out["offsets"]), _p(out["keys"]), _p(out["keyScores"]), _p(out["baseScores"]))
        k = int(cfg["keylen"][0])
        self.lib.orc_rcomp_key_fast.restype = C.c_int
        om = np.full((n, maxKeys), -1, np.int32); km = np.full((n, maxKeys), -1, np.int32)
        lens = np.diff(read_off)
        for r in range(n):
            m = int(out["nkeys"][r])
            for i in range(max(m, 0)):
                om[r, i] = lens[r] - (out["offsets"][r, m - 1 - i] + k)
                km[r, i] = self.lib.orc_rcomp_key_fast(C.c_int(int(out["keys"][r, m - 1 - i])), C.c_int(k))
        out["offsetsM"] = om; out["keysM"] = km
        return out

    # ---------------- BandedAligner ----------------
    def banded(self, dir, query, ref, qstart, rstart, maxEdits, exact, maxWidth):
        q = np.ascontiguousarray(query, np.int8); r = np.ascontiguousarray(ref, np.int8)
        rv = np.zeros(5, np.int32)
        e = self.lib.orc_banded_align(C.c_int(dir), _p(q), _p(r), C.c_int(len(q)), C.c_int(len(r)), C.c_int(qstart), C.c_int(rstart),
                                      C.c_int(maxEdits), C.c_int(1 if exact else 0), C.c_int(maxWidth), _p(rv))
        return int(e), rv.tolist()

    def banded_batch(self, queries, refs, tasks, kind="port", threads=1):
        queries = np.ascontiguousarray(queries, np.int8); refs = np.ascontiguousarray(refs, np.int8)
        tasks = np.ascontiguousarray(tasks, BAND_TASK_DTYPE)
        outs = np.zeros(len(tasks), BAND_OUT_DTYPE)
        rc = self.lib.orc_banded_batch(_p(queries), _p(refs), _p(tasks), _p(outs), C.c_int64(len(tasks)),
                                       C.c_int(1 if kind == "reference" else 0), C.c_int(threads))
        if rc != 0:
            raise RuntimeError("reference banded backend unavailable")
        return outs

    # ---------------- batch driver ----------------
    def run_batch(self, reads, refs, tasks, match_off=None, bandwidth=0, ratio=0.0, maxRows=601, maxColumns=3000,
                  kind="port", threads=1):
        """Returns (outs, match_buf, cells). `tasks` is a TASK_DTYPE array."""
        reads = np.ascontiguousarray(reads, np.int8); refs = np.ascontiguousarray(refs, np.int8)
        tasks = np.ascontiguousarray(tasks, TASK_DTYPE)
        outs = np.zeros(len(tasks), OUT_DTYPE)
        if match_off is None:
            match_off = match_offsets(tasks)
        match_off = np.ascontiguousarray(match_off, np.int64)
        mbuf = np.zeros(max(int(match_off[-1]), 1), np.int8)
        cells = self.lib.orc_batch_run(_p(reads), _p(refs), _p(tasks), _p(outs), C.c_int64(len(tasks)), _p(mbuf), _p(match_off),
                                       C.c_int(bandwidth), C.c_float(ratio), C.c_int(maxRows), C.c_int(maxColumns),
                                       C.c_int(1 if kind == "reference" else 0), C.c_int(threads))
        if cells < 0:
            raise RuntimeError("reference fill backend unavailable")
        return outs, mbuf, int(cells)


def match_offsets(tasks, extra=0):
    """Per-task match-string slots: rows + columns (+extra) bytes each (Java allocates rows+cols-1, MSA11tsJNI.java:380)."""
    a = tasks["ref_start"].astype(np.int64); b = tasks["ref_end"].astype(np.int64)
    clamp = (tasks["flags"] & TF_CLAMP) != 0
    a = np.where(clamp, np.maximum(a, 0), a)
    b = np.where(clamp, np.minimum(b, tasks["ref_len"].astype(np.int64) - 1), b)
    cap = tasks["read_len"].astype(np.int64) + np.maximum(b - a + 1, 0) + extra
    cap = (cap + 3) & ~np.int64(3)
    off = np.zeros(len(tasks) + 1, np.int64)
    np.cumsum(cap, out=off[1:])
    return off


_ORACLE = None


def get():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
