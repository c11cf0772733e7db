"""CPU: the tail of BBMapThread.processRead after scoreSlow (BBMapThread.java:478-709) as ONE sequence — final list policy, the genMatchString loop,
removeDuplicateBestSites, the clearzone-3 block and score gate, the tip penalty — assembled from the independent restatements (tests/pysitelist.py, pygenmatch.py /
pyrealign.py with fills by the reference's own C, the numpy clearzone 3 of tests/test_sitelist_oracle.py) and compared with the C chain (SL_FINAL +
orc_map_finish_single): the read record (locus, strand, mapScore, mapped / perfect / ambiguous, clearzone-3 subtraction, tip penalty) and the primary match string.
Reads whose sites carry gap arrays, and reads whose match string keeps X / Y / C symbols (toLocalAlignment, not restated twice), are left out."""
import functools

import numpy as np
import pytest

from bbmap_b200 import sitelist as sl
from bbmap_b200.mapper import map_cfg
from oracle import chain
from sitelist_cases import slow_cases

import pyclip
import pygenmatch
import pyrealign
import pysitelist as ps
from test_sitelist_independent import _to_sites
from test_sitelist_oracle import _py_clearzone3


def _finish(R, sites, fin, bp8, bm8, pcfg, mcfg):
    L = len(bp8)
    max_sw = 70 + 100 * (L - 1)
    cs = [pyclip.ClipSite(s, None) for s in sites]
    flags = (1 if cs else 0) | (2 if fin["perfect"] else 0) | (4 if fin["ambiguous"] else 0)
    if cs:
        first = True
        while True:
            if not first:
                cs.sort(key=functools.cmp_to_key(lambda a, b: ps.compare_to(a.s, b.s)))
            pygenmatch.gen_match_string(R, cs, bp8, bm8, max_sw, mcfg, True, False)
            cs[0].s.score = cs[0].s.slowScore
            first = False
            if not (len(cs) > 1 and cs[0].s.score < cs[1].s.score):
                break
        flags = (flags & ~2) | (2 if cs[0].s.perfect else 0)          # genMatchString: r.setPerfect(ss.perfect())
    n = len(cs)
    v = np.zeros(max(n, 1), sl.SS_DTYPE)
    for i, c in enumerate(cs):
        s = c.s
        v[i]["chrom"], v[i]["strand"], v[i]["start"], v[i]["stop"] = s.chrom, s.strand, s.start, s.stop
        v[i]["score"], v[i]["slow_score"], v[i]["paired_score"], v[i]["quick_score"] = s.score, s.slowScore, s.pairedScore, s.quickScore
    n2, flags, map_score, subi = _py_clearzone3(v, n, L, flags, pcfg)
    match = bytes(cs[0].match) if n2 else None
    pen = 0
    if n2:
        pen, st = ps.calc_tip_score_penalty(True, match, bp8.tolist(), map_score, 7)
        map_score -= pen
    top = cs[0].s if n2 else None
    return dict(mapped=bool(flags & 1) and n2 > 0, perfect=bool(flags & 2), ambiguous=bool(flags & 4), map_score=map_score if n2 else 0, cz3=subi, pen=pen,
                chrom=top.chrom if top else -1, start=top.start if top else -1, stop=top.stop if top else -1, strand=top.strand if top else 0, match=match)


@pytest.mark.parametrize("seed", [905, 906])
def test_process_read_tail(oracle, seed):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=260, seed=seed)
    pcfg = sl.policy_cfg(); mcfg = map_cfg()
    lists, _, _ = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, pcfg, P, M, refs, co)
    L2, status, _ = oracle.score_slow(lists, nss, ro, P, M, refs, co, np.ones(len(nss), np.int32), sl.slow_cfg())
    L3, n3, out = oracle.sitelist(sl.SL_FINAL, L2, nss, ro, pcfg)
    ms = chain.match_stride(int(np.diff(ro).max()), mcfg)
    L4, n4, recs, match, fills = oracle.map_finish_single(L3, n3, ro, P, M, refs, co, out, pcfg, mcfg, ms)
    P8 = np.ascontiguousarray(P).view(np.int8); M8 = np.ascontiguousarray(M).view(np.int8); R8 = np.ascontiguousarray(refs).view(np.int8)
    realigners = {}
    done = mapped = cz = pens = 0
    for r in range(len(nss)):
        n = int(nss[r])
        if n == 0 or status[r] or recs["status"][r] or (L2[r, :n]["ngaps"] > 0).any() or len(set(int(x) for x in L2[r, :n]["chrom"])) != 1:
            continue
        mlen = int(recs["match_len"][r])
        m_exp = match[r * ms: r * ms + mlen].tobytes() if mlen > 0 else None
        if m_exp is not None and any(c in m_exp for c in b"XYC"):
            continue
        ch = int(L2[r, 0]["chrom"])
        ref8 = R8[int(co[ch - 1]): int(co[ch])]
        if ch not in realigners:
            realigners[ch] = pyrealign.Realigner(oracle, ref8)
        a, b = int(ro[r]), int(ro[r + 1])
        sites = _to_sites(L2[r], n)
        fin = ps.final_policy(sites, b - a, pcfg[0])
        got = _finish(realigners[ch], sites, fin, P8[a:b].copy(), M8[a:b].copy(), pcfg, mcfg[0])
        e = recs[r]
        ef = int(e["flags"])
        assert got["mapped"] == bool(ef & 1), (r, got, e)
        if got["mapped"]:
            assert (got["chrom"], got["start"], got["stop"], got["strand"], got["map_score"], got["perfect"], got["ambiguous"], got["cz3"], got["pen"]) == \
                   (int(e["chrom"]), int(e["start"]), int(e["stop"]), int(e["strand"]), int(e["map_score"]), bool(ef & 2), bool(ef & 4), int(e["cz3_sub"]), int(e["tip_penalty"])), (r, got, e)
            assert got["match"] == m_exp, (r, got["match"], m_exp)
            mapped += 1; cz += got["cz3"] > 0; pens += got["pen"] > 0
        done += 1
    assert done > 120 and mapped > 80 and pens > 10, (done, mapped, cz, pens)
