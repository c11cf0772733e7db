"""Host-side mirror of the reference's index stage for the CUDA path: reference packing as FastaToChromArrays2 lays
chromosomes out (dna/FastaToChromArrays2.java:432-507,565-575), index build (IndexMaker4) and analysis
(BBIndex.analyzeIndex) on the device."""
import ctypes as C

import numpy as np

from . import lib as _lib

INDEX_CFG_DTYPE = np.dtype([("keylen", "<i4"), ("chrombits", "<i4"), ("shift_length", "<i4"), ("chroms_per_block", "<i4"),
                            ("max_hits_reduction2", "<i4"), ("maximum_max_hits_reduction", "<i4"), ("hit_reduction_div", "<i4"),
                            ("points_per_site", "<i4"), ("min_index_to_drop_long_hit_list", "<i4"), ("max_average_list_to_search", "<i4"),
                            ("max_average_list_to_search2", "<i4"), ("max_single_list_to_search", "<i4"),
                            ("max_shortest_list_to_search", "<i4"), ("max_usable_length", "<i4"), ("max_usable_length2", "<i4"), ("pad_", "<i4"),
                            ("fraction_to_exclude", "<f4"), ("padf_", "<f4", (3,))], align=True)
assert INDEX_CFG_DTYPE.itemsize == 80

START_PADDING, MID_PADDING, END_PADDING = 8000, 300, 8000        # dna/FastaToChromArrays2.java:569-571
MAX_LENGTH = (1 << 29) - 200000                                   # :575


def pack_chromosomes(scaffolds, max_length=MAX_LENGTH):
    """Lay scaffolds (uint8 arrays of upper-case ACGTN) out as chromosome arrays the way FastaToChromArrays2.makeNextChrom
    does: 8000 leading N, 300 N between merged scaffolds, trailing N until more than 8000 terminal N; a new chromosome starts
    when the next scaffold would not fit.  Returns (bytes uint8[], chrom_off int64[nchroms+1], scaffold table
    [(chrom (1-based), start, length)])."""
    chroms, table = [], []
    cur, nscaf = None, 0

    def finish(c):
        arr = np.concatenate(c)
        term = 0
        for b in arr[::-1]:
            if b == ord("N") and term < END_PADDING:
                term += 1
            else:
                break
        pad = 0
        while term <= END_PADDING and len(arr) + pad < max_length:
            pad += 1; term += 1
        return np.concatenate([arr, np.full(pad, ord("N"), np.uint8)])

    for s in scaffolds:
        s = np.ascontiguousarray(s, np.uint8)
        if cur is not None and len(s) + MID_PADDING + END_PADDING + (sum(len(x) for x in cur) - 1) > max_length:
            chroms.append(finish(cur)); cur = None
        if cur is None:
            cur = [np.full(START_PADDING, ord("N"), np.uint8)]; nscaf = 0
        if nscaf > 0:
            cur.append(np.full(MID_PADDING, ord("N"), np.uint8))
        table.append((len(chroms) + 1, sum(len(x) for x in cur), len(s)))
        cur.append(s); nscaf += 1
    if cur is not None:
        chroms.append(finish(cur))
    off = np.zeros(len(chroms) + 1, np.int64)
    np.cumsum([len(c) for c in chroms], out=off[1:])
    return np.concatenate(chroms), off, table


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class BBIndexCUDA:
    """Device-resident index of one packed reference (replicated per GPU)."""

    def __init__(self, chrom_bytes, chrom_off, keylen=13, chrombits=-1, device=0, ctx=None, load_from=None, build=1):
        """load_from: the reference's Data.ROOT_INDEX ("<path>/ref/index/") — read the blocks IndexMaker4 / `save` wrote there instead of building
        (IndexMaker4.java:135-139); the analysis is recomputed on the device either way."""
        self.L = _lib.load()
        if self.L.bbm_device_count() <= 0:
            raise _lib.BbmError("no CUDA device visible: BBIndexCUDA has no CPU fallback")
        self._own = ctx is None
        if ctx is None:
            h = C.c_void_p()
            _lib.check(self.L.bbm_init(device, C.byref(h)), "bbm_init")
            ctx = h
        self.h = ctx
        self.chrom_off = np.ascontiguousarray(chrom_off, np.int64)
        b = np.ascontiguousarray(chrom_bytes).view(np.int8)
        d = C.c_void_p()
        _lib.check(self.L.bbm_upload(self.h, _p(b), b.size, C.byref(d)), "bbm_upload")
        self.d_chroms = d
        self.cfg = np.zeros(1, INDEX_CFG_DTYPE)
        nb = C.c_int32(0)
        if load_from is None:
            _lib.check(self.L.bbm_index_build(self.h, d, _p(self.chrom_off), len(self.chrom_off) - 1, keylen, chrombits, _p(self.cfg), C.byref(nb)),
                       "bbm_index_build")
        else:
            import os
            _lib.check(self.L.bbm_index_load(self.h, d, _p(self.chrom_off), len(self.chrom_off) - 1, keylen, chrombits, os.fsencode(load_from), build, _p(self.cfg),
                                             C.byref(nb)), "bbm_index_load")
        self.nblocks = nb.value
        self.keylen = keylen

    def close(self):
        if self._own and getattr(self, "h", None):
            self.L.bbm_destroy(self.h)
        self.h = None

    def save(self, root_index, build=1):
        """Block.write for every block (IndexMaker4.java:197-200): <root_index><build>/chr<a>[-<b>]_index_k<k>_c<chrombits>_b<build>.block + ...block2.gz"""
        import os
        os.makedirs(os.path.join(root_index, str(build)), exist_ok=True)
        _lib.check(self.L.bbm_index_save(self.h, os.fsencode(root_index), build), "bbm_index_save")

    def download(self, block=0):
        n = C.c_int64(0)
        _lib.check(self.L.bbm_index_block_sites(self.h, block, C.byref(n)), "bbm_index_block_sites")
        ks = 1 << (2 * self.keylen)
        starts = np.zeros(ks + 1, np.int32); sites = np.zeros(max(n.value, 1), np.int32)
        counts = np.zeros(ks, np.int32); hist = np.zeros(1001, np.int32)
        _lib.check(self.L.bbm_index_download(self.h, block, _p(starts), _p(sites), _p(counts), _p(hist)), "bbm_index_download")
        return starts, sites[: n.value], counts, hist
