"""CPU: AbstractMapThread.genMatchString / genMatchStringForSite — the C restatement inside oracle/mapper_oracle.c (through its test entry point) must
equal a second restatement written from the Java text (tests/pygenmatch.py on tests/pyrealign.py; fills by the reference's own C) on the read's list afterwards
(order, every field), the top site's match string, the `paired` flag and the number of fills, for lists as scoreSlow leaves them (no gap arrays)."""
import ctypes as C

import numpy as np
import pytest

from bbmap_b200 import sitelist as sl
from bbmap_b200.mapper import map_cfg
from sitelist_cases import slow_cases

import pyclip
import pygenmatch
import pyrealign
from test_sitelist_independent import _same, _to_sites


def maxq_of(ro, r):
    return 70 + 100 * (int(ro[r + 1] - ro[r]) - 1)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("seed,kw,set_score,paired", [(805, {}, 1, 0), (806, dict(paired=1), 0, 1), (807, dict(paired=1), 1, 0)])
def test_gen_match_string(oracle, seed, kw, set_score, paired):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=240, seed=seed)
    pcfg = sl.policy_cfg()
    lists, _, _ = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, pcfg, P, M, refs, co)
    L2, status, _ = oracle.score_slow(lists, nss, ro, P, M, refs, co, np.ones(len(nss), np.int32), sl.slow_cfg())
    L3, n3, fl = oracle.sitelist(sl.SL_FINAL, L2, nss, ro, pcfg)                  # merged, sorted, gated: what processRead hands to genMatchString
    cfg = map_cfg(**kw)
    P8 = np.ascontiguousarray(P).view(np.int8); M8 = np.ascontiguousarray(M).view(np.int8); R8 = np.ascontiguousarray(refs).view(np.int8)
    lib = oracle.lib
    lib.orc_test_gen_match_string.restype = C.c_int
    realigners = {}
    rng = np.random.default_rng(seed)
    done = multi = resorted = 0
    for r in range(len(nss)):
        n = int(n3[r])
        if n == 0 or status[r] or (L3[r, :n]["ngaps"] > 0).any():
            continue
        chroms = set(int(x) for x in L3[r, :n]["chrom"])
        if len(chroms) != 1:
            continue
        ch = chroms.pop()
        rows = L3[r, :n].copy()
        if r % 3 == 0:                                       # a misplaced top site: its realignment changes scores, which is what the re-sort loop is for
            d = int(rng.integers(-20, 21)); rows[0]["start"] += d; rows[0]["stop"] += d
        if r % 3 == 1 and n > 1:  # a top site that claims more than its alignment gives: after realign_new the list is out of order
            better, worse = rows[0].copy(), rows[1].copy()
            if int(better["slow_score"]) > int(worse["slow_score"]) + 60 and not worse["perfect"] and not worse["semiperfect"]:
                worse["slow_score"] = worse["score"] = int(better["slow_score"]) + 50 if int(better["slow_score"]) + 50 < maxq_of(ro, r) else int(better["slow_score"])
                rows[0], rows[1] = worse, better
        ref8 = R8[int(co[ch - 1]): int(co[ch])]
        co1 = np.array([0, len(ref8)], np.int64)
        a, b = int(ro[r]), int(ro[r + 1]); L = b - a
        maxq = 70 + 100 * (L - 1)
        cap = n + 2
        buf = np.zeros(cap, sl.SS_DTYPE); buf[:n] = rows; buf["chrom"][:n] = 1
        nn = np.array([n], np.int32); pf = np.array([paired], np.int32)
        tm = np.zeros(L + 3400, np.int8); tl = np.array([-1], np.int32)
        bp, bm = P8[a:b].copy(), M8[a:b].copy()
        nf = lib.orc_test_gen_match_string(_p(buf), _p(nn), C.c_int(cap), _p(bp), _p(bm), C.c_int(L), _p(ref8), _p(co1), _p(cfg), C.c_int(maxq), C.c_int(set_score),
                                           _p(pf), _p(tm), _p(tl), C.c_int(len(tm)))
        if ch not in realigners:
            realigners[ch] = pyrealign.Realigner(oracle, ref8)
        R = realigners[ch]
        sites = [pyclip.ClipSite(s, None) for s in _to_sites(rows, n)]
        for cs in sites:
            cs.s.chrom = 1
        f0 = R.fills
        got_paired = pygenmatch.gen_match_string(R, sites, bp, bm, maxq, cfg[0], bool(set_score), bool(paired))
        _same([cs.s for cs in sites], buf, int(nn[0]), r)
        assert bytes(sites[0].match) == tm[: int(tl[0])].tobytes(), (r, bytes(sites[0].match), tm[: int(tl[0])].tobytes())
        assert int(got_paired) == int(pf[0]) and R.fills - f0 == nf, (r, got_paired, pf, R.fills - f0, nf)
        done += 1; multi += n > 1; resorted += [cs.s.start for cs in sites][:1] != [int(rows[0]["start"])]
    resorts = sum(getattr(R, "resorts", 0) for R in realigners.values())
    assert done > 120 and multi > 40 and resorts > 3, (done, multi, resorted, resorts)
