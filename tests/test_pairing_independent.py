"""CPU: the pairing helpers of processReadPair — pairSiteScoresInitial, pairSiteScoresFinal, canPair of the C restatement
(oracle/mapper_oracle.c, through its test entry points) must equal a second restatement written from the Java text alone (tests/pypairing.py) on both
reads' lists: order, scores, paired scores, which sites the trims keep, and the number of perfect pairs."""
import ctypes as C

import numpy as np
import pytest

from bbmap_b200 import sitelist as sl
from bbmap_b200.mapper import map_cfg

import pypairing as pp
from test_sitelist_independent import _same, _to_sites


def _mate_lists(rng, n_pairs, cap=24, after_alignment=False):
    """Two site lists per pair: mates placed 0-700 bp apart on opposite strands (proper pairs), plus same-strand neighbours, sites on other chromosomes,
    far-away sites, duplicates of a position with other scores, perfect sites."""
    A = np.zeros((n_pairs, cap), sl.SS_DTYPE); B = np.zeros((n_pairs, cap), sl.SS_DTYPE)
    nA = np.zeros(n_pairs, np.int32); nB = np.zeros(n_pairs, np.int32)
    lens = np.zeros((n_pairs, 2), np.int32)
    for p in range(n_pairs):
        L1, L2 = (int(x) for x in rng.choice([50, 100, 150, 250], size=2))
        lens[p] = (L1, L2)
        na = int(rng.choice([0, 1, 2, 3, 5, 6, 9, 14, 20])); nb = int(rng.choice([0, 1, 2, 3, 5, 6, 9, 14, 20]))
        if rng.random() < 0.8:
            na, nb = max(na, 1), max(nb, 1)
        maxq1, maxq2 = 70 + 100 * (L1 - 1), 70 + 100 * (L2 - 1)
        for i in range(na):
            s = A[p, i]
            s["chrom"] = int(rng.integers(1, 4)); s["strand"] = int(rng.integers(0, 2)); s["start"] = int(rng.integers(0, 40000)); s["stop"] = s["start"] + L1 - 1 + int(rng.integers(0, 3))
            sc = maxq1 if rng.random() < 0.25 else int(maxq1 * rng.uniform(0.2, 1.0))
            s["score"] = sc; s["quick_score"] = sc; s["slow_score"] = sc if after_alignment else 0
            s["perfect"] = 1 if sc == maxq1 and rng.random() < 0.8 else 0; s["semiperfect"] = 1 if s["perfect"] or rng.random() < 0.1 else 0
            s["hits"] = int(rng.integers(1, 18))
        for i in range(nb):
            s = B[p, i]
            if na and rng.random() < 0.7:            # near one of the first read's sites
                m = A[p, int(rng.integers(0, na))]
                s["chrom"] = m["chrom"]; s["strand"] = (1 - m["strand"]) if rng.random() < 0.8 else m["strand"]
                d = int(rng.integers(-60, 700)) if rng.random() < 0.85 else int(rng.integers(30000, 36000))
                s["start"] = max(0, (int(m["stop"]) + d) if m["strand"] == 0 or rng.random() < 0.3 else (int(m["start"]) - d - L2))
            else:
                s["chrom"] = int(rng.integers(1, 4)); s["strand"] = int(rng.integers(0, 2)); s["start"] = int(rng.integers(0, 40000))
            s["stop"] = s["start"] + L2 - 1 + int(rng.integers(0, 3))
            sc = maxq2 if rng.random() < 0.25 else int(maxq2 * rng.uniform(0.2, 1.0))
            s["score"] = sc; s["quick_score"] = sc; s["slow_score"] = sc if after_alignment else 0
            s["perfect"] = 1 if sc == maxq2 and rng.random() < 0.8 else 0; s["semiperfect"] = 1 if s["perfect"] or rng.random() < 0.1 else 0
            s["hits"] = int(rng.integers(1, 18))
        if na >= 2 and na < cap and rng.random() < 0.3:
            A[p, na] = A[p, 0]; A[p, na]["score"] -= 37; na += 1
        nA[p] = na; nB[p] = nb
    return A, nA, B, nB, lens


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("which,seed,kw", [("initial", 301, {}), ("initial", 302, dict(average_pair_dist=350)), ("initial", 303, dict(require_correct_strands=0)),
                                           ("final", 304, {}), ("final", 305, dict(average_pair_dist=20, secondary_site_score_ratio=0.9)),
                                           ("final", 306, dict(same_strand_pairs=1))])
def test_pair_site_scores(oracle, which, seed, kw):
    rng = np.random.default_rng(seed)
    A, nA, B, nB, lens = _mate_lists(rng, 1200, after_alignment=(which == "final"))
    cfg = map_cfg(paired=1, **kw)
    max_trim = 800
    lib = oracle.lib
    lib.orc_test_pair_initial.restype = C.c_int
    lib.orc_test_pair_final.restype = None
    paired_sites = trimmed = perfect = 0
    for p in range(len(nA)):
        a = A[p].copy(); b = B[p].copy()
        na = np.array([nA[p]], np.int32); nb = np.array([nB[p]], np.int32)
        sa = _to_sites(A[p], int(nA[p])); sb = _to_sites(B[p], int(nB[p]))
        L1, L2 = int(lens[p, 0]), int(lens[p, 1])
        if which == "initial":
            exp = lib.orc_test_pair_initial(_p(a), _p(na), C.c_int(L1), _p(b), _p(nb), C.c_int(L2), _p(cfg), C.c_int(max_trim))
            got = pp.pair_site_scores_initial(sa, L1, sb, L2, cfg[0], max_trim)
            assert got == exp, (p, got, exp)
            perfect += got > 0
        else:
            lib.orc_test_pair_final(_p(a), _p(na), C.c_int(L1), _p(b), _p(nb), C.c_int(L2), _p(cfg), C.c_int(max_trim))
            pp.pair_site_scores_final(sa, L1, sb, L2, cfg[0], max_trim)
        _same(sa, a, int(na[0]), (p, "first read"))
        _same(sb, b, int(nb[0]), (p, "mate"))
        paired_sites += sum(1 for s in sa if s.pairedScore > 0)
        trimmed += len(sa) < nA[p] or len(sb) < nB[p]
    assert paired_sites > 800 and trimmed > 100 and (perfect > 20 or which == "final"), (paired_sites, trimmed, perfect)


def test_can_pair(oracle):
    rng = np.random.default_rng(310)
    A, nA, B, nB, lens = _mate_lists(rng, 600)
    lib = oracle.lib
    lib.orc_test_can_pair.restype = C.c_int
    yes = no = 0
    for kw in ({}, dict(require_correct_strands=0), dict(same_strand_pairs=1), dict(max_pair_dist=300)):
        cfg = map_cfg(paired=1, **kw)
        for p in range(len(nA)):
            sa = _to_sites(A[p], int(nA[p])); sb = _to_sites(B[p], int(nB[p]))
            for i in range(min(3, len(sa))):
                for j in range(min(4, len(sb))):
                    exp = lib.orc_test_can_pair(_p(A[p, i:i + 1]), _p(B[p, j:j + 1]), C.c_int(int(lens[p, 0])), C.c_int(int(lens[p, 1])), _p(cfg))
                    got = pp.can_pair(sa[i], sb[j], int(lens[p, 0]), int(lens[p, 1]), cfg[0])
                    assert bool(exp) == got, (p, i, j, kw)
                    yes += got; no += not got
    assert yes > 500 and no > 500


def test_remove_low_quality_sites_paired(oracle):
    rng = np.random.default_rng(320)
    lib = oracle.lib
    lib.orc_test_remove_low_quality_paired.restype = C.c_int
    cleared = trimmed = 0
    for _ in range(1500):
        L = int(rng.choice([50, 100, 150, 250])); maxq = 70 + 100 * (L - 1)
        n = int(rng.integers(1, 14))
        v = np.zeros(n, sl.SS_DTYPE)
        top = int(maxq * rng.uniform(0.2, 1.0))
        sc = np.sort((top * rng.uniform(0.1, 1.0, size=n)).astype(np.int32))[::-1]
        v["score"] = sc; v["slow_score"] = sc; v["chrom"] = 1; v["start"] = rng.integers(0, 9000, size=n); v["stop"] = v["start"] + L - 1
        v["paired_score"] = np.where(rng.random(n) < 0.5, sc + rng.integers(1, 400, size=n), 0)
        ms, mp = (0.56, 0.448) if rng.random() < 0.7 else (0.7, 0.3)
        sites = _to_sites(v, n)
        a = v.copy()
        k = lib.orc_test_remove_low_quality_paired(_p(a), C.c_int(n), C.c_int(maxq), C.c_float(ms), C.c_float(mp))
        removed = pp.remove_low_quality_sites_paired(sites, maxq, ms, mp)
        # the C restatement signals "list cleared" by returning 0 survivors only when the top site fails; it returns the survivor count otherwise
        if len(sites) == 0:
            assert k == 0, (k, removed)
            cleared += 1
        else:
            _same(sites, a, k, "rlqsp")
            trimmed += removed > 0
    assert cleared > 100 and trimmed > 300


def test_is_bad_pair(oracle):
    from bbmap_b200.mapper import MAP_REC_DTYPE
    rng = np.random.default_rng(321)
    lib = oracle.lib
    lib.orc_test_is_bad_pair.restype = C.c_int

    class R:
        pass
    bad = good = 0
    for kw in ({}, dict(require_correct_strands=0), dict(same_strand_pairs=1), dict(max_pair_dist=400)):
        cfg = map_cfg(paired=1, **kw)
        for _ in range(1500):
            recs = np.zeros(2, MAP_REC_DTYPE); objs = []
            for i in range(2):
                o = R()
                o.chrom = int(rng.integers(1, 3)) if rng.random() < 0.2 else 1
                o.start = int(rng.integers(0, 3000)) if rng.random() < 0.8 else int(rng.integers(0, 60000)); o.stop = o.start + int(rng.choice([49, 99, 149]))
                o.strand = int(rng.integers(0, 2)); o.mapped = rng.random() < 0.9; o.paired = rng.random() < 0.2
                recs[i]["chrom"] = o.chrom; recs[i]["start"] = o.start; recs[i]["stop"] = o.stop; recs[i]["strand"] = o.strand
                recs[i]["flags"] = (1 if o.mapped else 0) | (8 if o.paired else 0)
                objs.append(o)
            exp = lib.orc_test_is_bad_pair(_p(recs[0:1]), _p(recs[1:2]), _p(cfg))
            got = pp.is_bad_pair(objs[0], objs[1], bool(cfg["require_correct_strands"][0]), bool(cfg["same_strand_pairs"][0]), int(cfg["max_pair_dist"][0]))
            assert bool(exp) == got, (kw, recs)
            bad += got; good += not got
    assert bad > 800 and good > 800
