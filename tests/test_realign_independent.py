"""CPU: TranslateColorspaceRead.realign_new (the primary site's match string) — the C restatement inside oracle/mapper_oracle.c must equal a second
restatement written from the Java text (tests/pyrealign.py; every fill by the reference's own C, every walk by tests/pywalk.py) on plain and gapped sites:
the new match string, start / stop, the three scores, the perfect bits and the number of fills requested."""
import ctypes as C

import numpy as np
import pytest

from bbmap_b200 import sitelist as sl
from sitelist_cases import slow_cases

import pyclip
import pyrealign
import pysitelist as ps


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("seed,ratio,recur", [(605, 0.56, 1), (606, 0.336, 1), (607, 0.56, 0)])
def test_realign_new(oracle, seed, ratio, recur):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=220, seed=seed)
    lists, _, _ = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, sl.policy_cfg(), P, M, refs, co)
    L2, status, _ = oracle.score_slow(lists, nss, ro, P, M, refs, co, np.ones(len(nss), np.int32), sl.slow_cfg())
    P8 = np.ascontiguousarray(P).view(np.int8); M8 = np.ascontiguousarray(M).view(np.int8); R8 = np.ascontiguousarray(refs).view(np.int8)
    lib = oracle.lib
    lib.orc_test_realign_new.restype = C.c_int
    realigners = {}
    rng = np.random.default_rng(seed)
    done = indel = refilled = gapped = 0
    for r in range(len(nss)):
        for i in range(min(int(nss[r]), 2)):
            rec = L2[r, i:i + 1].copy()
            if status[r]:
                continue
            ng = int(rec["ngaps"][0])
            if (r + i) % 2 and ng == 0:                                   # a misplaced or shrunken site: the alignment runs into the window edge, which is what the
                d = int(rng.integers(-28, 29))                # padding suggestions, the wider re-fills and the recursion are for
                rec["start"] += d; rec["stop"] += d - int(rng.integers(0, 12))
                if rec["stop"][0] <= rec["start"][0]:
                    rec["stop"] = rec["start"] + 5
            ch = int(rec["chrom"][0])
            ref8 = R8[int(co[ch - 1]): int(co[ch])]
            a, b = int(ro[r]), int(ro[r + 1])
            bases8 = (P8 if rec["strand"][0] == 0 else M8)[a:b].copy()
            L = b - a
            maxq = 70 + 100 * (L - 1)
            min_valid = int(np.float32(ratio) * np.float32(maxq)) - 258
            s = rec[0]
            site = ps.Site(ch, int(s["strand"]), int(s["start"]), int(s["stop"]), int(s["hits"]), int(s["score"]), int(s["quick_score"]), int(s["slow_score"]),
                           int(s["paired_score"]), bool(s["perfect"]), bool(s["semiperfect"]), bool(s["rescued"]), None if ng == 0 else [int(x) for x in s["gaps"][:ng]])
            cs = pyclip.ClipSite(site, None)
            co1 = np.array([0, len(ref8)], np.int64)
            rec["chrom"] = 1
            mbuf = np.zeros(L + 3200, np.int8); mlen = np.array([-1], np.int32)
            nf = lib.orc_test_realign_new(_p(rec), _p(mbuf), _p(mlen), C.c_int(len(mbuf)), _p(bases8), C.c_int(L), _p(ref8), _p(co1), C.c_int(4), C.c_int(recur),
                                          C.c_int(min_valid), C.c_int(0), C.c_int(0))
            if ch not in realigners:
                realigners[ch] = pyrealign.Realigner(oracle, ref8)
            R = realigners[ch]
            f0 = R.fills
            R.realign(cs, bases8, 4, recur, min_valid, False, False)
            e = rec[0]
            assert bytes(cs.match) == mbuf[: int(mlen[0])].tobytes(), (r, i, bytes(cs.match), mbuf[: int(mlen[0])].tobytes())
            assert (site.start, site.stop, site.score, site.slowScore, site.pairedScore, int(site.perfect), int(site.semiperfect)) == \
                   (int(e["start"]), int(e["stop"]), int(e["score"]), int(e["slow_score"]), int(e["paired_score"]), int(e["perfect"]), int(e["semiperfect"])), (r, i)
            assert R.fills - f0 == nf, (r, i, R.fills - f0, nf)
            assert ([] if site.gaps is None else list(site.gaps)) == e["gaps"][: int(e["ngaps"])].tolist(), (r, i, site.gaps, e["gaps"], e["ngaps"])
            gapped += ng > 0
            done += 1; indel += (b"D" in bytes(cs.match)) or (b"I" in bytes(cs.match)); refilled += nf > 1
    assert done > 150 and indel > 40 and refilled > 40 and gapped > 10, (done, indel, refilled, gapped)
