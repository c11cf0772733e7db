"""The C restatement of the per-read site-list policies (oracle/sitelist_oracle.c) against lists whose outcome is worked out by hand from
the reference's text (BBMapThread.java:140-249, 478-553; Tools.java:654-760, 913-1003).  Parity unpinned against Java (no JVM)."""
import numpy as np

from bbmap_b200 import sitelist as sl
from sitelist_cases import noindel_lists, random_lists


def _mk(rows, L=100, cap=8):
    """rows: list of dicts of SS fields."""
    lists = np.zeros((1, cap), sl.SS_DTYPE)
    for i, d in enumerate(rows):
        for k, v in d.items():
            lists[0, i][k] = v
        if "quick_score" not in d:
            lists[0, i]["quick_score"] = d.get("score", 0)
        if "stop" not in d:
            lists[0, i]["stop"] = lists[0, i]["start"] + L - 1
    return lists, np.array([len(rows)], np.int32), np.array([0, L], np.int64)


def test_trim_list_by_hand(oracle):
    cfg = sl.policy_cfg()
    # perfect top score: cutoff (int)(9970*.6f)=5982; at most size-3 = 1 removal, from the end; then the .94 pass cannot shrink 3 sites
    lists, nss, ro = _mk([dict(chrom=1, start=10, score=5000), dict(chrom=1, start=20, score=9970, perfect=1, semiperfect=1),
                          dict(chrom=1, start=30, score=4000), dict(chrom=1, start=40, score=9000)])
    L2, n2, out = oracle.sitelist(sl.SL_TRIM, lists, nss, ro, cfg)
    assert n2[0] == 3 and list(L2[0, :3]["score"]) == [9970, 9000, 5000] and out["best_sites"][0] == 9970
    # imperfect top: .6 -> cutoff 4800 removes 2000, 3000, 4000 (maxToRemove = 3); the list is then too short for the later passes
    lists, nss, ro = _mk([dict(chrom=1, start=10 * i, score=s) for i, s in enumerate([3000, 8000, 2000, 7000, 5000, 4000])])
    L2, n2, out = oracle.sitelist(sl.SL_TRIM, lists, nss, ro, cfg)
    assert n2[0] == 3 and list(L2[0, :3]["score"]) == [8000, 7000, 5000] and out["best_sites"][0] == 8000
    # semiperfect sites survive any cutoff; a single site and an empty list are untouched
    lists, nss, ro = _mk([dict(chrom=1, start=10 * i, score=s, semiperfect=sp) for i, (s, sp) in enumerate([(9000, 0), (100, 1), (90, 1), (80, 1), (70, 0)])])
    L2, n2, _ = oracle.sitelist(sl.SL_TRIM, lists, nss, ro, cfg)
    assert n2[0] == 4 and list(L2[0, :4]["score"]) == [9000, 100, 90, 80]
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=5)])
    assert oracle.sitelist(sl.SL_TRIM, lists, nss, ro, cfg)[1][0] == 1


def test_final_policy_by_hand(oracle):
    cfg = sl.policy_cfg()
    f32 = np.float32
    # perfect read, second site inside CLEARZONEP=160: ambiguous
    lists, nss, ro = _mk([dict(chrom=1, start=100, score=9900, slow_score=9900), dict(chrom=1, start=5, score=9970, slow_score=9970, perfect=1, semiperfect=1)])
    L2, n2, out = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
    assert n2[0] == 2 and out["flags"][0] == sl.F_MAPPED | sl.F_PERFECT | sl.F_AMBIGUOUS and out["clearzone"][0] == 160 and out["best_sites"][0] == 2
    assert L2[0, 0]["score"] == 9970
    # imperfect: clearzone interpolated between CLEARZONE1b and CLEARZONE1 -> 210; the runner-up at 9400 is outside it
    lists, nss, ro = _mk([dict(chrom=1, start=100, score=9400, slow_score=9400), dict(chrom=1, start=5, score=9700, slow_score=9700)])
    L2, n2, out = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
    lim = f32(9970) * f32(0.97) - f32(1200)
    cz = int((f32((9970 - 9700) * 260) + (f32(9700) - lim) * f32(200)) / (f32(9970) - lim))
    assert cz == 210 and out["clearzone"][0] == 210 and out["flags"][0] == sl.F_MAPPED and out["best_sites"][0] == 1 and n2[0] == 2
    # ... and at 9500 inside it
    lists[0, 0]["score"] = lists[0, 0]["slow_score"] = 9500
    assert oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)[2]["flags"][0] == sl.F_MAPPED | sl.F_AMBIGUOUS
    # same place twice: merged (scores maxed), then a lone site is never ambiguous
    lists, nss, ro = _mk([dict(chrom=2, start=50, strand=1, score=9000, slow_score=9000), dict(chrom=2, start=50, strand=1, score=8000, slow_score=8500)])
    L2, n2, out = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
    assert n2[0] == 1 and L2[0, 0]["score"] == 9000 and L2[0, 0]["slow_score"] == 9000 and out["flags"][0] == sl.F_MAPPED
    # below (int)(9970*.56f)=5583: unmapped; low third site (< 5583-800) dropped, the second is always kept (loop stops at index 2)
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=5500, slow_score=5500)])
    assert oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)[1][0] == 0
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=9000, slow_score=9000), dict(chrom=1, start=300, score=4000, slow_score=4000),
                          dict(chrom=1, start=600, score=3000, slow_score=3000)])
    L2, n2, out = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
    assert n2[0] == 2 and list(L2[0, :2]["score"]) == [9000, 4000]


def test_noindel_policy_properties(oracle):
    refs, co, P, M, ro, lists, nss = noindel_lists(nreads=800, seed=7)
    cfg = sl.policy_cfg(quick_match_strings=1)
    L2, n2, out = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, cfg, P, M, refs, co)
    assert np.array_equal(n2, nss)
    seen_shift = seen_match = 0
    for r in range(len(nss)):
        L = int(ro[r + 1] - ro[r]); maxq = 70 + 100 * (L - 1); v = L2[r, :n2[r]]
        assert all(v["score"][i] >= v["score"][i + 1] for i in range(len(v) - 1))              # sorted by compareTo
        assert (v["score"] == v["slow_score"]).all() and (v["score"] <= maxq).all()
        near = (v["slow_score"] >= maxq - 495).sum()
        assert abs(out["near_perfect"][r]) == near
        for s in v[v["slow_score"] >= maxq - 495]:
            assert s["stop"] - s["start"] + 1 == L and s["ngaps"] == 0
            assert bool(s["perfect"]) == (s["slow_score"] == maxq)
        seen_match += int(v["has_match"].sum())
    # the stop-anchored rescoring moved some long sites onto the read
    before = {(r, int(s["stop"])) for r in range(len(nss)) for s in lists[r, :nss[r]] if s["stop"] - s["start"] + 1 != ro[r + 1] - ro[r]}
    after = {(r, int(s["stop"])) for r in range(len(nss)) for s in L2[r, :n2[r]] if s["slow_score"] >= 70 + 100 * (ro[r + 1] - ro[r] - 1) - 495}
    assert len(before & after) > 20 and seen_match > 50


def test_random_lists_invariants(oracle):
    lists, nss, ro = random_lists(nreads=1500, seed=9)
    cfg = sl.policy_cfg()
    L2, n2, out = oracle.sitelist(sl.SL_TRIM, lists, nss, ro, cfg)
    assert (n2 <= nss).all() and (n2[nss >= 3] >= 3).all() and (n2[nss < 3] == nss[nss < 3]).all() and (n2 < nss).sum() > 300
    for r in np.nonzero(nss > 1)[0][:300]:
        assert out["best_sites"][r] == lists[r, :nss[r]]["score"].max() == L2[r, 0]["score"]
    lists, nss, ro = random_lists(nreads=1500, seed=10, after_alignment=True)
    L3, n3, out3 = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
    assert ((out3["flags"] & sl.F_MAPPED) != 0).sum() > 300 and ((out3["flags"] & sl.F_AMBIGUOUS) != 0).sum() > 50 and (n3 == 0).sum() > 100


def test_remove_out_of_bounds_by_hand(oracle):
    """chromosome 1: maxIndex 9999, scaffolds at 1000 (len 3000) and 4300 (len 5000), 300 N between them (Data.interScaffoldPadding)."""
    L = 100
    rows = [dict(chrom=1, start=-3, stop=96, score=1),            # hangs over the start: removed
            dict(chrom=1, start=9950, stop=10049, score=2),       # hangs over maxIndex: removed
            dict(chrom=1, start=1500, stop=1599, score=3),        # inside scaffold 0
            dict(chrom=1, start=3950, stop=4349, score=4),        # runs from scaffold 0 into scaffold 1: removed with SAM output
            dict(chrom=1, start=5000, stop=5000 + 2600, score=5), # over-long: cut to start + read length + 40
            dict(chrom=1, start=4300, stop=4399, score=6)]        # first base of scaffold 1
    lists, nss, ro = _mk(rows, L=L, cap=8)
    scaf = (np.array([0, 2], np.int32), np.array([1000, 4300], np.int32), np.array([3000, 5000], np.int32))
    L2, n2, out = oracle.sitelist_bounds(lists, nss, ro, [9999], scaf)
    assert n2[0] == 3 and out["best_sites"][0] == 3
    assert [(int(s["start"]), int(s["stop"])) for s in L2[0, :3]] == [(1500, 1599), (5000, 5140), (4300, 4399)]
    # without SAM output the scaffold test is skipped
    L3, n3, _ = oracle.sitelist_bounds(lists, nss, ro, [9999], scaf, sam_out=0)
    assert n3[0] == 4 and int(L3[0, 1]["start"]) == 3950


# ---- clearzone 3 / tip penalty: a second, independent restatement in numpy float32, straight from the Java text ----
_F = np.float32
_CZ3_MULTS = [_F(x) for x in (0, 1, .75, .5, .25, .125, .0625)]           # AbstractMapThread.java:2809


def _py_cz3_fraction(s1, s2, cz3, inv):                                    # :1893-1911
    dif = s1 - s2
    if dif >= cz3:
        return _F(0)
    f = _F(cz3 - dif) * inv
    f2 = f * f
    return f + _F(2) * f2 + _F(2) * f2 * f


def _py_set_slow(s, x):                                                    # SiteScore.setSlowScore, stream/SiteScore.java:962-983
    if x <= 0:
        s["paired_score"] = x
    elif s["paired_score"] > 0:
        s["paired_score"] = x + (s["paired_score"] - s["slow_score"]) if s["slow_score"] > 0 else x + 1
    s["slow_score"] = x


def _py_clearzone3(v, n, L, flags, cfg, toss=False):
    """BBMapThread.java:667-684, 698-700 + applyClearzone3 :1820-1870 for one read; returns (n, flags, mapScore, subi)."""
    maxSw = 70 + 100 * (L - 1); ratio = _F(cfg["min_align_ratio"][0]); CZ3 = int(cfg["clearzone3"][0])
    for i in range(n - 1, 0, -1):                                          # removeDuplicateBestSites, AbstractMapThread.java:1328-1349
        if all(v[0][k] == v[i][k] for k in ("chrom", "strand", "start", "stop")):
            n -= 1
        else:
            break
    if n == 0:
        flags &= ~1
    mapScore = int(v[0]["slow_score"]) if n else 0
    subi = 0
    if (CZ3 > cfg["clearzone1"][0] or CZ3 > cfg["clearzonep"][0]) and n > 0 and not flags & 4 and mapScore > 0:
        cz3v2 = _F(CZ3) * min(_F(1.25), _F(maxSw) / _F(mapScore))
        cz3, inv = int(cz3v2), _F(1) / cz3v2
        if flags & 1 and n >= 2:
            sub = _F(0)
            for i in range(1, min(7, n)):
                if i > 2 and v[i]["slow_score"] < v[i - 1]["slow_score"]:
                    break
                f = _py_cz3_fraction(mapScore, int(v[i]["slow_score"]), cz3, inv)
                if f <= 0:
                    break
                sub = sub + f * _CZ3_MULTS[i]
            if sub > 0:
                asym = _F(4) + _F(0.03) * _F(L)
                sub = sub * _F(1.8)
                sub2 = _F(cz3) * ((asym * sub) / (sub + asym))
                subi = int(sub2 + _F(0.5))
                if subi >= mapScore - 300:
                    subi = mapScore - 300
                if subi <= 0:
                    subi = 0
                else:
                    for i in range(n):
                        _py_set_slow(v[i], int(v[i]["slow_score"]) - subi); v[i]["score"] -= subi
        if subi > 0:
            mapScore -= subi
            if mapScore < int(_F(maxSw) * ratio):
                flags |= 4
    if flags & 4 and toss:
        n = 0; flags &= ~1; mapScore = 0
    if n == 0 or (not flags & 4 and _F(mapScore) < _F(maxSw) * ratio):
        n = 0; flags &= ~1; mapScore = 0
    return n, flags, mapScore, subi


def test_clearzone3_by_hand(oracle):
    cfg = sl.policy_cfg()
    fl = np.zeros(1, sl.READ_OUT_DTYPE); fl["flags"] = sl.F_MAPPED
    # cz3v2 = 800*min(1.25, 9970/9000) = 886.22 -> CLEARZONE3 886; the runner-up 100 below gives f = 786/886.22, sub = 1.8*(f+2f^2+2f^3) = 6.94,
    # sub2 = 886*(7*sub/(sub+7)) = 3087.6 -> every score drops by 3088; 5912 is still above (int)(9970*.56f) = 5583
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=9000, slow_score=9000), dict(chrom=1, start=500, score=8900, slow_score=8900, paired_score=9100)])
    L2, n2, out = oracle.sitelist_clearzone3(lists, nss, ro, fl, cfg)
    assert out["best_sites"][0] == 3088 and out["near_perfect"][0] == 5912 and out["flags"][0] == sl.F_MAPPED and n2[0] == 2
    assert list(L2[0, :2]["slow_score"]) == [5912, 5812] and list(L2[0, :2]["score"]) == [5912, 5812] and L2[0, 1]["paired_score"] == 5812 + 200
    # runner-up a full clearzone below, a single site, an ambiguous read: untouched
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=9000, slow_score=9000), dict(chrom=1, start=500, score=8100, slow_score=8100)])
    assert oracle.sitelist_clearzone3(lists, nss, ro, fl, cfg)[2]["best_sites"][0] == 0
    fa = fl.copy(); fa["flags"] = sl.F_MAPPED | sl.F_AMBIGUOUS
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=9000, slow_score=9000), dict(chrom=1, start=500, score=8990, slow_score=8990)])
    L2, n2, out = oracle.sitelist_clearzone3(lists, nss, ro, fa, cfg)
    assert out["best_sites"][0] == 0 and n2[0] == 2 and out["flags"][0] == sl.F_MAPPED | sl.F_AMBIGUOUS
    assert oracle.sitelist_clearzone3(lists, nss, ro, fa, cfg, ambiguous_toss=True)[1][0] == 0
    # a tie close to the ratio gate: the subtraction is capped at mapScore-300, the read turns ambiguous and stays mapped
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=5700, slow_score=5700), dict(chrom=1, start=500, score=5700, slow_score=5700),
                          dict(chrom=1, start=900, score=5690, slow_score=5690)])
    L2, n2, out = oracle.sitelist_clearzone3(lists, nss, ro, fl, cfg)
    exp = _py_clearzone3(lists[0].copy(), 3, 100, sl.F_MAPPED, cfg)
    assert (int(n2[0]), int(out["flags"][0]), int(out["near_perfect"][0]), int(out["best_sites"][0])) == exp
    assert out["flags"][0] == sl.F_MAPPED | sl.F_AMBIGUOUS and out["best_sites"][0] > 0 and out["near_perfect"][0] >= 300
    # CLEARZONE3 not above CLEARZONE1/P: only the ratio gate acts (5500 < 5583.2 -> unmapped)
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=5500, slow_score=5500)])
    L2, n2, out = oracle.sitelist_clearzone3(lists, nss, ro, fl, sl.policy_cfg(clearzone3=100))
    assert n2[0] == 0 and out["flags"][0] == 0 and out["near_perfect"][0] == 0


def test_clearzone3_against_numpy_restatement(oracle):
    cfg = sl.policy_cfg()
    lists, nss, ro = random_lists(nreads=1500, cap=12, seed=77, after_alignment=True)
    lists, nss, fl = oracle.sitelist(sl.SL_FINAL, lists, nss, ro, cfg)
    for r in range(0, len(nss), 9):                                        # copies of the top site at the tail, as a realignment can leave them
        k = int(nss[r])
        if 2 <= k < lists.shape[1] - 1:
            lists[r, k] = lists[r, 0]; lists[r, k + 1] = lists[r, 0]; lists[r, k + 1]["slow_score"] -= 7; nss[r] = k + 2
    L2, n2, out = oracle.sitelist_clearzone3(lists, nss, ro, fl, cfg)
    assert (n2 < nss).sum() > 20
    changed = 0
    for r in range(len(nss)):
        v = lists[r].copy()
        exp = _py_clearzone3(v, int(nss[r]), int(ro[r + 1] - ro[r]), int(fl["flags"][r]), cfg)
        assert (int(n2[r]), int(out["flags"][r]), int(out["near_perfect"][r]), int(out["best_sites"][r])) == exp, r
        if exp[0]:
            assert v[:exp[0]].tobytes() == L2[r, :exp[0]].tobytes(), r
        changed += exp[3] > 0
    assert changed > 50


def test_tip_penalty_by_hand(oracle):
    fl = np.zeros(1, sl.READ_OUT_DTYPE); fl["flags"] = sl.F_MAPPED
    lists, nss, ro = _mk([dict(chrom=1, start=1, score=9000, slow_score=9000), dict(chrom=1, start=500, score=8000, slow_score=8000)])
    bases = np.frombuffer((b"ACGT" * 25), np.int8)
    enc = lambda s: (np.frombuffer(s.encode(), np.int8), np.array([0, len(s)], np.int64))
    # one substitution at read position 3: 2*(7+2-3) = 12 points -> (int)((80*12/92)*.0022f*9970) = 228
    m, mo = enc("mmmSmmmm" + "m" * 92)
    L2, pen, st = oracle.sitelist_tip_penalty(lists, nss, ro, bases, m, mo, fl)
    assert pen[0] == 228 and st[0] == 0 and list(L2[0, :2]["score"]) == [8772, 7772] and list(L2[0, :2]["slow_score"]) == [8772, 7772]
    # all matches: nothing; a deletion run counts once (2*(9-2) = 14), an N costs half (9-0 = 9) at the far tip
    m, mo = enc("m" * 100)
    assert oracle.sitelist_tip_penalty(lists, nss, ro, bases, m, mo, fl)[1][0] == 0
    m, mo = enc("mmDDDmm" + "m" * 95 + "N")
    pts = 14 + 9
    exp = int((_F(80) * _F(pts)) / (_F(pts) + _F(80)) * _F(.0022) * _F(9970))
    assert oracle.sitelist_tip_penalty(lists, nss, ro, bases, m, mo, fl)[1][0] == exp
    # homopolymer tips: AAAAAC... adds 3 points at the left, ...GTTT adds 1 at the right
    hb = np.frombuffer(b"AAAAAC" + b"ACGT" * 22 + b"ACGTTT", np.int8)
    m, mo = enc("m" * 100)
    exp = int((_F(80) * _F(4)) / (_F(4) + _F(80)) * _F(.0022) * _F(9970))
    assert oracle.sitelist_tip_penalty(lists, nss, ro, hb, m, mo, fl)[1][0] == exp
    # capped at mapScore - maxScore/10; unmapped, no match string, short format, truncated string
    low, nl, _ = _mk([dict(chrom=1, start=1, score=1000, slow_score=1000)])
    m, mo = enc("SSSSSSSS" + "m" * 92)
    assert oracle.sitelist_tip_penalty(low, nl, ro, bases, m, mo, fl)[1][0] == 1000 - 997
    f0 = fl.copy(); f0["flags"] = 0
    assert oracle.sitelist_tip_penalty(lists, nss, ro, bases, m, mo, f0)[1][0] == 0
    assert oracle.sitelist_tip_penalty(lists, nss, ro, bases, m[:0], np.array([0, 0], np.int64), fl)[1][0] == 0
    m, mo = enc("m3S" + "m" * 90)
    assert oracle.sitelist_tip_penalty(lists, nss, ro, bases, m, mo, fl)[2][0] == 2
    m, mo = enc("mmmm")
    assert oracle.sitelist_tip_penalty(lists, nss, ro, bases, m, mo, fl)[2][0] == 1
