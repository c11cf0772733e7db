"""CPU: the per-read site-list policies of the unpaired loop — the C restatement the CUDA kernels are checked against
(oracle/sitelist_oracle.c) must equal a second restatement written from the Java text alone (tests/pysitelist.py: SiteScore objects in Python lists,
Collections.sort as a stable comparator sort, numpy float32 for Java float) on every surviving site in order, and on the per-read outcome
(mapped / perfect / ambiguous, clearzone, number of best sites, highest quick score)."""
import numpy as np
import pytest

from bbmap_b200 import sitelist as sl
from sitelist_cases import random_lists

import pysitelist as ps


def _to_sites(row, n):
    out = []
    for i in range(n):
        s = row[i]
        g = None if s["ngaps"] == 0 else [int(x) for x in s["gaps"][: s["ngaps"]]]
        out.append(ps.Site(int(s["chrom"]), int(s["strand"]), int(s["start"]), int(s["stop"]), int(s["hits"]), int(s["score"]), int(s["quick_score"]),
                           int(s["slow_score"]), int(s["paired_score"]), bool(s["perfect"]), bool(s["semiperfect"]), bool(s["rescued"]), g, tag=i))
    return out


def _same(sites, row, n, r):
    assert len(sites) == n, (r, len(sites), n)
    for i, s in enumerate(sites):
        t = row[i]
        got = (s.chrom, s.strand, s.start, s.stop, s.hits, s.score, s.quickScore, s.slowScore, s.pairedScore, int(s.perfect), int(s.semiperfect), int(s.rescued),
               [] if s.gaps is None else list(s.gaps))
        exp = (int(t["chrom"]), int(t["strand"]), int(t["start"]), int(t["stop"]), int(t["hits"]), int(t["score"]), int(t["quick_score"]), int(t["slow_score"]),
               int(t["paired_score"]), int(t["perfect"]), int(t["semiperfect"]), int(t["rescued"]), t["gaps"][: t["ngaps"]].tolist())
        assert got == exp, (r, i, got, exp)


@pytest.mark.parametrize("seed,kw", [(109, {}), (110, dict(min_trim_sites_to_retain=1)), (111, dict(max_trim_sites_to_retain=20)), (112, dict(trim_list=0))])
def test_trim_policy(oracle, seed, kw):
    lists, nss, ro = random_lists(nreads=1200, seed=seed)
    cfg = sl.policy_cfg(**kw)
    L2, n2, out = oracle.sitelist(sl.SL_TRIM, lists, nss, ro, cfg)
    trimmed = 0
    for r in range(len(nss)):
        sites = _to_sites(lists[r], int(nss[r]))
        hi = ps.trim_policy(sites, int(ro[r + 1] - ro[r]), cfg[0])
        _same(sites, L2[r], int(n2[r]), r)
        if hi is not None:
            assert hi == out["best_sites"][r], (r, hi, out[r])
        trimmed += len(sites) < nss[r]
    assert trimmed > 200 or kw.get("trim_list") == 0


@pytest.mark.parametrize("seed,kw", [(120, {}), (121, dict(clearzone3=100)), (122, dict(min_align_ratio=0.3, clearzone_limit1e=2))])
def test_final_policy(oracle, seed, kw):
    lists, nss, ro = random_lists(nreads=1500, seed=seed, after_alignment=True)
    # paired scores on some sites: mergeDuplicateSites' setSlowScore / setPairedScore interplay
    rng = np.random.default_rng(seed)
    for r in range(0, len(nss), 5):
        for i in range(int(nss[r])):
            if rng.random() < 0.5:
                lists[r, i]["paired_score"] = int(lists[r, i]["slow_score"]) + int(rng.integers(-50, 400))
    cfg = sl.policy_cfg(**kw)
    L2, n2, out = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
    seen = dict(mapped=0, ambiguous=0, perfect=0, merged=0, dropped=0)
    for r in range(len(nss)):
        sites = _to_sites(lists[r], int(nss[r]))
        res = ps.final_policy(sites, int(ro[r + 1] - ro[r]), cfg[0])
        _same(sites, L2[r], int(n2[r]), r)
        f = int(out["flags"][r])
        assert (bool(f & sl.F_MAPPED), bool(f & sl.F_AMBIGUOUS)) == (res["mapped"], res["ambiguous"]), (r, f, res)
        if res["mapped"]:
            assert bool(f & sl.F_PERFECT) == res["perfect"], (r, f, res)
        if res["clearzone"] is not None:
            assert (int(out["clearzone"][r]), int(out["best_sites"][r])) == (res["clearzone"], res["best_sites"]), (r, out[r], res)
        for k in ("mapped", "ambiguous", "perfect"):
            seen[k] += res[k]
        seen["dropped"] += nss[r] > 0 and not res["mapped"]
    assert seen["mapped"] > 300 and seen["ambiguous"] > 50 and seen["perfect"] > 50 and (seen["dropped"] > 50 or "min_align_ratio" in kw), seen


@pytest.mark.parametrize("seed,tiplen", [(3, 7), (4, 5)])
def test_tip_penalty(oracle, seed, tiplen):
    """calcTipScorePenalty + applyScorePenalty: the penalty of every read and every score of its list."""
    from sitelist_cases import homopolymer_reads, random_match_strings
    lists, nss, ro = random_lists(nreads=1500, cap=8, seed=seed + 90, after_alignment=True)
    l1, n1, fl = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, sl.policy_cfg())
    bases = homopolymer_reads(ro, seed); match, mo = random_match_strings(ro, seed)
    el, ep, es = oracle.sitelist_tip_penalty(l1, n1, ro, bases, match, mo, fl, tiplen)
    b8 = np.ascontiguousarray(bases).view(np.int8); m8 = np.ascontiguousarray(match).view(np.uint8)
    hit = 0
    for r in range(len(n1)):
        n = int(n1[r])
        sites = _to_sites(l1[r], n)
        mapped = bool(fl["flags"][r] & sl.F_MAPPED) and n > 0
        ms = m8[int(mo[r]): int(mo[r + 1])]
        mstr = bytes(ms) if len(ms) else None
        pen, st = ps.calc_tip_score_penalty(mapped, mstr, b8[int(ro[r]): int(ro[r + 1])].tolist(), sites[0].score if n else 0, tiplen)
        assert (pen, st) == (int(ep[r]), int(es[r])), (r, pen, st, ep[r], es[r], mstr)
        ps.apply_score_penalty(sites, pen)
        _same(sites, el[r], n, r)
        hit += pen > 0
    assert hit > 250


@pytest.mark.parametrize("seed,sam_out,with_scaf", [(91, 1, True), (92, 0, True), (93, 1, False)])
def test_remove_out_of_bounds(oracle, seed, sam_out, with_scaf):
    """removeOutOfBounds with GapTools.calcGrefLen / fixGaps and Data.isSingleScaffold: surviving sites, their stops and gap arrays."""
    import pysam_fields as psf
    rng = np.random.default_rng(seed)
    n, cap = 1500, 12
    lists, nss, ro = random_lists(nreads=n, cap=cap, seed=seed)
    maxidx = np.array([5200, 4900, 5600], np.int32)
    scaf_loc = [np.array([100, 1800, 3700]), np.array([50]), np.array([0, 900, 1700, 2500, 4000])]
    scaf = (np.cumsum([0] + [len(x) for x in scaf_loc]).astype(np.int32), np.concatenate(scaf_loc).astype(np.int32), None) if with_scaf else None
    for r in range(n):
        for i in range(nss[r]):
            u = rng.random()
            if u < 0.05:
                lists[r, i]["start"] = -int(rng.integers(1, 50)); lists[r, i]["stop"] = lists[r, i]["start"] + 120
            elif u < 0.10:
                lists[r, i]["stop"] = int(maxidx[lists[r, i]["chrom"] - 1]) + int(rng.integers(0, 3))
            elif u < 0.14:
                lists[r, i]["stop"] = lists[r, i]["start"] + int(rng.integers(2500, 2600))
    S = psf.Scaffolds([(ch + 1, int(x), 1) for ch, locs in enumerate(scaf_loc) for x in locs], 300) if with_scaf else None
    single = (lambda c, a, b: S.is_single(c, a, b)) if with_scaf else (lambda c, a, b: True)
    mi = {c + 1: int(maxidx[c]) for c in range(3)}
    removed_total = cut = 0
    for limit in (2522, 180):
        L2, n2, out = oracle.sitelist_bounds(lists, nss, ro, maxidx, scaf, sam_out=sam_out, expected_len_limit=limit)
        for r in range(n):
            sites = _to_sites(lists[r], int(nss[r]))
            before = [(s.start, s.stop) for s in sites]
            removed = ps.remove_out_of_bounds(sites, int(ro[r + 1] - ro[r]), mi, single, bool(sam_out), limit)
            for s in sites:
                if s.gaps is None:
                    s.gaps = None
            # a gap array that fixGaps dissolved is None here and ngaps == 0 on the C side
            _same(sites, L2[r], int(n2[r]), (r, limit))
            assert removed == int(out["best_sites"][r]), (r, removed, out[r])
            removed_total += removed
            cut += sum(1 for s in sites if (s.start, s.stop) not in before)
    assert removed_total > 300 and cut > 100, (removed_total, cut)
