"""Host-side mirror of BBIndex.find (current/align2/BBIndex.java:403-639) for the CUDA path: batched index search that
turns seeds into candidate sites (SiteScore records)."""
import ctypes as C

import numpy as np

from . import lib as _lib

MAX_SITES, MAX_GAPS = 48, 10
SITE_DTYPE = np.dtype([("chrom", "<i4"), ("start", "<i4"), ("stop", "<i4"), ("hits", "<i4"), ("score", "<i4"), ("ngaps", "<i4"),
                       ("strand", "i1"), ("perfect", "i1"), ("semiperfect", "i1"), ("pad_", "i1"), ("gaps", "<i4", (MAX_GAPS - 1,))], align=True)
SEARCH_RESULT_DTYPE = np.dtype([("nsites", "<i4"), ("status", "<i4"), ("num_hits", "<i4"), ("max_score", "<i4"), ("max_quick_score", "<i4"),
                                ("pad_", "<i4"), ("best_scores", "<i4", (6,)), ("sites", SITE_DTYPE, (MAX_SITES,))], align=True)
assert SITE_DTYPE.itemsize == 64 and SEARCH_RESULT_DTYPE.itemsize == 48 + 64 * MAX_SITES

HEAD_DTYPE = np.dtype([("nsites", "<i4"), ("status", "<i4"), ("num_hits", "<i4"), ("max_score", "<i4"), ("max_quick_score", "<i4"),
                       ("pad_", "<i4"), ("best_scores", "<i4", (6,))], align=True)
assert HEAD_DTYPE.itemsize == 48


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def search_batch(index, bases, baseScores, read_off, seeds, max_sites=MAX_SITES, quit_after_two_perfects=True, shared=False, split=3):
    """BBIndex.find for a batch of reads against `index` (a bbmap_b200.index.BBIndexCUDA).  `seeds` is the dict returned by
    KeyRingCUDA.seed_batch (nkeys, offsets, keyScores).  Returns (heads HEAD_DTYPE[n], sites SITE_DTYPE[n, max_sites])."""
    L = index.L
    bases = np.ascontiguousarray(bases).view(np.int8); bs = np.ascontiguousarray(baseScores).view(np.int8)
    ro = np.ascontiguousarray(read_off, np.int64)
    n = len(ro) - 1
    nk = np.ascontiguousarray(seeds["nkeys"], np.int32); of = np.ascontiguousarray(seeds["offsets"], np.int32)
    ks = np.ascontiguousarray(seeds["keyScores"], np.int32)
    heads = np.zeros(n, HEAD_DTYPE); sites = np.zeros((n, max_sites), SITE_DTYPE)
    _lib.check(L.bbm_set_option(index.h, b"search_shared", 1 if shared else 0), "bbm_set_option")
    _lib.check(L.bbm_set_option(index.h, b"search_split", int(split)), "bbm_set_option")        # 0 one launch, 1 three thread-per-read phases, 2 warp-per-read prescan, 3 (default) warp-per-read prescan and walk
    _lib.check(L.bbm_search_batch_host(index.h, _p(bases), _p(bs), _p(ro), n, _p(nk), _p(of), _p(ks), of.shape[1],
                                       1 if quit_after_two_perfects else 0, _p(heads), _p(sites), max_sites), "bbm_search_batch_host")
    return heads, sites
