"""The C restatement of BBMapThread.scoreSlow (oracle/scoreslow_oracle.c) on a case built from SURVEY Appendix B's known-answer vector
(KAT3: a 3-base deletion scores 9402 over reference 100..202) and on seeded lists.  Parity unpinned against Java (no JVM)."""
import numpy as np

from bbmap_b200 import sitelist as sl
from kat import REF
from sitelist_cases import slow_cases


def test_score_slow_kat3(oracle):
    ref = np.frombuffer(REF.encode() if isinstance(REF, str) else bytes(REF), np.int8)
    read = np.concatenate([ref[100:150], ref[153:203]])
    rc = np.frombuffer(bytes(read)[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA")), np.int8)
    lists = np.zeros((2, 2), sl.SS_DTYPE)
    for r in range(2):
        lists[r, 0]["chrom"] = 1; lists[r, 0]["start"] = 100; lists[r, 0]["stop"] = 199
    nss = np.array([1, 1], np.int32); ro = np.array([0, 100, 200], np.int64); co = np.array([0, 400], np.int64)
    P = np.concatenate([read, read]); M = np.concatenate([rc, rc])
    cfg = sl.policy_cfg()
    lists, _, out = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, cfg, P, M, ref, co)
    assert out["near_perfect"][0] == 0 and lists[0, 0]["slow_score"] < 9970 - 495
    L2, status, na = oracle.score_slow(lists, nss, ro, P, M, ref, co, np.array([1, 0], np.int32), sl.slow_cfg())
    s = L2[0, 0]
    assert (s["slow_score"], s["score"], s["start"], s["stop"], s["perfect"], s["semiperfect"]) == (9402, 9402, 100, 202, 0, 0)
    assert na == 1 and (status == 0).all()
    assert L2[1, 0].tobytes() == lists[1, 0].tobytes()                   # run[r] == 0: untouched


def test_score_slow_properties(oracle):
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=400, seed=5)
    cfg = sl.policy_cfg()
    lists, _, _ = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, cfg, P, M, refs, co)
    L2, status, na = oracle.score_slow(lists, nss, ro, P, M, refs, co, run, sl.slow_cfg())
    assert na > 300 and not (status & sl.SLOW_GAPPED).any()
    gapped = [(r, i) for r in range(len(nss)) for i in range(nss[r]) if lists[r, i]["ngaps"] > 0 and run[r]]
    assert len(gapped) > 20
    for r, i in gapped:                                   # gapped sites go through the gapped reference; their gap array follows start/stop
        b = L2[r, i]
        assert b["ngaps"] == 0 or (b["gaps"][0] == b["start"] and b["gaps"][b["ngaps"] - 1] == b["stop"] and b["ngaps"] % 2 == 0)
    improved = moved = 0
    for r in range(len(nss)):
        Lr = int(ro[r + 1] - ro[r]); maxq = 70 + 100 * (Lr - 1)
        for i in range(nss[r]):
            a, b = lists[r, i], L2[r, i]
            if not run[r]:
                assert a.tobytes() == b.tobytes(); continue
            assert b["score"] == b["slow_score"] <= maxq and bool(b["perfect"]) == (b["slow_score"] == maxq)
            if a["stop"] - a["start"] == Lr - 1:
                assert b["slow_score"] >= a["slow_score"] - 120     # fillLimited runs with minScore-MIN_SCORE_ADJUST (MSA.java:868)
            improved += b["slow_score"] > a["slow_score"]; moved += (b["stop"] - b["start"]) != (a["stop"] - a["start"])
    assert improved > 100 and moved > 60
