"""GPU parity (part of SURVEY §8 f1): the per-read site-list policies on the device vs the C restatement — every field of every surviving
site, list lengths and per-read outputs bit-exact — for trimList, scoreNoIndels(Read) and the post-alignment list handling, plus the
conversion of BBIndex.find's sites into SiteScore records."""
import ctypes as C

import numpy as np
import pytest

from bbmap_b200 import lib as _lib
from bbmap_b200 import sitelist as sl
from sitelist_cases import noindel_lists, random_lists

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def msa():
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    m = MultiStateAligner11tsCUDA(device=0)
    yield m
    m.close()


def _same(got, exp):
    gl, gn, go = got; el, en, eo = exp
    assert np.array_equal(gn, en)
    for f in eo.dtype.names:
        assert np.array_equal(go[f], eo[f]), f
    live = np.arange(el.shape[1])[None, :] < en[:, None]
    assert gl[live].tobytes() == el[live].tobytes()


@pytest.mark.parametrize("seed,cap", [(51, 40), (52, 64), (53, 5)])
def test_trim_parity(oracle, msa, seed, cap):
    lists, nss, ro = random_lists(nreads=4000, cap=cap, seed=seed)
    for cfg in (sl.policy_cfg(), sl.policy_cfg(min_trim_sites_to_retain=1), sl.policy_cfg(trim_list=0)):
        _same(sl.site_lists(msa.h, sl.SL_TRIM, lists, nss, ro, cfg), oracle.sitelist(sl.SL_TRIM, lists, nss, ro, cfg))


@pytest.mark.parametrize("seed", [61, 62])
def test_noindel_policy_parity(oracle, msa, seed):
    refs, co, P, M, ro, lists, nss = noindel_lists(nreads=3000, seed=seed)
    d_ref = msa.load_reference(refs)
    try:
        for cfg in (sl.policy_cfg(), sl.policy_cfg(print_secondary=1), sl.policy_cfg(quick_match_strings=0)):
            exp = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, cfg, P, M, refs, co)
            _same(sl.site_lists(msa.h, sl.SL_NOINDEL, lists, nss, ro, cfg, P, M, d_ref, co), exp)
    finally:
        msa.free(d_ref)
    assert (exp[2]["near_perfect"] < 0).any() and (exp[2]["near_perfect"] > 0).any()


@pytest.mark.parametrize("seed", [71, 72])
def test_final_policy_parity(oracle, msa, seed):
    lists, nss, ro = random_lists(nreads=5000, cap=24, seed=seed, after_alignment=True)
    for cfg in (sl.policy_cfg(), sl.policy_cfg(clearzone3=0), sl.policy_cfg(clearzone_limit1e=2, clearzone1e=300)):
        exp = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
        _same(sl.site_lists(msa.h, sl.SL_FINAL, lists, nss, ro, cfg), exp)
    assert (exp[2]["flags"] & sl.F_AMBIGUOUS).any()


def test_bad_arguments(msa):
    lists, nss, ro = random_lists(nreads=4, cap=5, seed=1)
    L = _lib.load()
    out = np.zeros(4, sl.READ_OUT_DTYPE); cfg = sl.policy_cfg()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.bbm_sitelist_batch_host(msa.h, 9, p(lists), p(nss), 4, 5, p(ro), None, None, None, None, 0, p(cfg), p(out)) == _lib.BBM_E_ARG
    assert L.bbm_sitelist_batch_host(msa.h, sl.SL_NOINDEL, p(lists), p(nss), 4, 5, p(ro), None, None, None, None, 0, p(cfg), p(out)) == _lib.BBM_E_ARG
    big = np.zeros((4, 65), sl.SS_DTYPE)
    assert L.bbm_sitelist_batch_host(msa.h, sl.SL_TRIM, p(big), p(nss), 4, 65, p(ro), None, None, None, None, 0, p(cfg), p(out)) == _lib.BBM_E_ARG


def test_from_search(msa):
    import torch
    from bbmap_b200.search import HEAD_DTYPE, SITE_DTYPE
    rng = np.random.default_rng(3)
    n, ms, cap = 500, 6, 4
    heads = np.zeros(n, HEAD_DTYPE); sites = np.zeros((n, ms), SITE_DTYPE)
    heads["nsites"] = rng.integers(0, ms + 1, size=n)
    for f in ("chrom", "start", "stop", "hits", "score", "ngaps"):
        sites[f] = rng.integers(0, 9, size=(n, ms)) if f == "ngaps" else rng.integers(1, 100000, size=(n, ms))
    sites["strand"] = rng.integers(0, 2, size=(n, ms)); sites["perfect"] = rng.integers(0, 2, size=(n, ms)); sites["semiperfect"] = sites["perfect"]
    sites["gaps"] = rng.integers(0, 1000, size=sites["gaps"].shape)
    dev = torch.device("cuda", 0)
    d_h = torch.from_numpy(heads.view(np.uint8)).to(dev); d_s = torch.from_numpy(sites.view(np.uint8).reshape(-1)).to(dev)
    d_l = torch.zeros(n * cap * sl.SS_DTYPE.itemsize, dtype=torch.uint8, device=dev); d_n = torch.zeros(n, dtype=torch.int32, device=dev)
    L = _lib.load()
    q = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(L.bbm_sitelist_from_search_dev(msa.h, q(d_h), q(d_s), n, ms, q(d_l), q(d_n), cap, None), "from_search")
    torch.cuda.synchronize()
    got = np.frombuffer(d_l.cpu().numpy().tobytes(), sl.SS_DTYPE).reshape(n, cap); gn = d_n.cpu().numpy()
    assert np.array_equal(gn, np.minimum(heads["nsites"], cap))
    for r in range(n):
        for i in range(gn[r]):
            a, b = got[r, i], sites[r, i]
            assert (a["chrom"], a["start"], a["stop"], a["hits"], a["score"], a["quick_score"], a["slow_score"], a["paired_score"]) == \
                   (b["chrom"], b["start"], b["stop"], b["hits"], b["score"], b["score"], 0, 0)
            assert (a["strand"], a["perfect"], a["semiperfect"], a["rescued"], a["ngaps"]) == (b["strand"], b["perfect"], b["semiperfect"], 0, b["ngaps"])
            assert np.array_equal(a["gaps"], b["gaps"])


@pytest.mark.parametrize("seed,with_quality", [(81, True), (86, False), (87, True)])
def test_find_tip_deletions_read_parity(oracle, msa, seed, with_quality):
    """findTipDeletions(Read, ...): quality gate, both tips, rescoring of the changed sites — on lists scored by scoreNoIndels."""
    from bbmap_b200 import rescue as rs
    from sitelist_cases import slow_cases
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=3000, seed=seed)
    lists, _, _ = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, sl.policy_cfg(), P, M, refs, co)
    rng = np.random.default_rng(seed)
    quality = None
    if with_quality:
        quality = np.full(len(P), 30, np.int8)
        for r in rng.choice(len(nss), size=600, replace=False):          # low-quality tips switch one or both searches off
            a, b = int(ro[r]), int(ro[r + 1])
            if rng.random() < 0.5:
                quality[a:a + 8] = rng.integers(0, 16, size=8)
            if rng.random() < 0.5:
                quality[b - 8:b] = rng.integers(0, 16, size=8)
    for cfg, mi in ((rs.tipdel_cfg(), None), (rs.tipdel_cfg(search_range=30), np.array([250], np.int32))):
        exp, eo = oracle.sitelist_tipdel(lists, nss, ro, P, M, quality, refs, co, cfg, mi)
        d_ref = msa.load_reference(refs)
        try:
            got, go = sl.findTipDeletions(msa.h, lists, nss, ro, P, M, quality, d_ref, co, cfg, mi)
        finally:
            msa.free(d_ref)
        for f in eo.dtype.names:
            assert np.array_equal(go[f], eo[f]), f
        live = np.arange(exp.shape[1])[None, :] < nss[:, None]
        assert got[live].tobytes() == exp[live].tobytes()
    assert eo["best_sites"].sum() > 20 and not eo["flags"].any()


@pytest.mark.parametrize("seed,sam_out,with_scaf", [(91, 1, True), (92, 0, True), (93, 1, False)])
def test_remove_out_of_bounds_parity(oracle, msa, seed, sam_out, with_scaf):
    rng = np.random.default_rng(seed)
    n, cap = 4000, 12
    lists, nss, ro = random_lists(nreads=n, cap=cap, seed=seed)
    maxidx = np.array([5200, 4900, 5600], np.int32)
    scaf_loc = [np.array([100, 1800, 3700]), np.array([50]), np.array([0, 900, 1700, 2500, 4000])]
    scaf = (np.cumsum([0] + [len(x) for x in scaf_loc]).astype(np.int32), np.concatenate(scaf_loc).astype(np.int32), None) if with_scaf else None
    for r in range(n):                      # sites hanging over the arrays, over-long sites, gapped over-long sites
        for i in range(nss[r]):
            u = rng.random()
            if u < 0.05:
                lists[r, i]["start"] = -int(rng.integers(1, 50)); lists[r, i]["stop"] = lists[r, i]["start"] + 120
            elif u < 0.10:
                lists[r, i]["stop"] = int(maxidx[lists[r, i]["chrom"] - 1]) + int(rng.integers(0, 3))
            elif u < 0.14:
                lists[r, i]["stop"] = lists[r, i]["start"] + int(rng.integers(2500, 2600))
    for cfg_limit in (2522, 180):
        exp = oracle.sitelist_bounds(lists, nss, ro, maxidx, scaf, sam_out=sam_out, expected_len_limit=cfg_limit)
        got = sl.removeOutOfBounds(msa.h, lists, nss, ro, maxidx, scaf, sam_out=sam_out, expected_len_limit=cfg_limit)
        assert np.array_equal(got[1], exp[1])
        for f in exp[2].dtype.names:
            assert np.array_equal(got[2][f], exp[2][f]), f
        live = np.arange(cap)[None, :] < exp[1][:, None]
        assert got[0][live].tobytes() == exp[0][live].tobytes()
    assert exp[2]["best_sites"].sum() > 500


@pytest.mark.parametrize("seed", [77, 78])
def test_clearzone3_parity(oracle, msa, seed):
    """processRead's clearzone-3 block + score gate (BBMapThread.java:667-684, 698-700): lists, lengths, flags, mapScore, amount subtracted."""
    lists, nss, ro = random_lists(nreads=6000, cap=16, seed=seed, after_alignment=True)
    for cfg, toss in ((sl.policy_cfg(), False), (sl.policy_cfg(), True), (sl.policy_cfg(clearzone3=150), False), (sl.policy_cfg(min_align_ratio=0.7), False)):
        l1, n1, fl = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
        for r in range(0, len(n1), 9):                                     # copies of the top site at the tail (removeDuplicateBestSites)
            k = int(n1[r])
            if 2 <= k < l1.shape[1] - 1:
                l1[r, k] = l1[r, 0]; l1[r, k + 1] = l1[r, 0]; l1[r, k + 1]["slow_score"] -= 7; n1[r] = k + 2
        exp = oracle.sitelist_clearzone3(l1, n1, ro, fl, cfg, ambiguous_toss=toss)
        assert (exp[1] < n1).sum() > 20
        _same(sl.applyClearzone3(msa.h, l1, n1, ro, fl, cfg, ambiguous_toss=toss), exp)
    l1, n1, fl = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, sl.policy_cfg())
    exp = oracle.sitelist_clearzone3(l1, n1, ro, fl, sl.policy_cfg())
    assert (exp[2]["best_sites"] > 0).sum() > 100 and ((exp[2]["flags"] & sl.F_AMBIGUOUS) > (fl["flags"] & sl.F_AMBIGUOUS)).any()
    assert (oracle.sitelist_clearzone3(l1, n1, ro, fl, sl.policy_cfg(), ambiguous_toss=True)[1] < n1).any()
    empty = sl.applyClearzone3(msa.h, l1[:0], n1[:0], ro[:1], fl[:0])
    assert empty[0].shape[0] == 0


@pytest.mark.parametrize("seed,tiplen", [(5, 7), (6, 7), (7, 4)])
def test_tip_penalty_parity(oracle, msa, seed, tiplen):
    """calcTipScorePenalty + applyScorePenalty (AbstractMapThread.java:2499-2567, 2601-2609): penalty, status and every score of the list."""
    from sitelist_cases import homopolymer_reads, random_match_strings
    lists, nss, ro = random_lists(nreads=6000, cap=8, seed=seed + 90, after_alignment=True)
    l1, n1, fl = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, sl.policy_cfg())
    bases = homopolymer_reads(ro, seed); match, mo = random_match_strings(ro, seed)
    el, ep, es = oracle.sitelist_tip_penalty(l1, n1, ro, bases, match, mo, fl, tiplen)
    gl, gp, gs = sl.tipScorePenalty(msa.h, l1, n1, ro, bases, match, mo, fl, tiplen)
    assert np.array_equal(gp, ep) and np.array_equal(gs, es)
    live = np.arange(el.shape[1])[None, :] < n1[:, None]
    assert gl[live].tobytes() == el[live].tobytes()
    assert (ep > 0).sum() > 1000 and (es == 1).any() and (es == 2).any()
