"""TEST INFRASTRUCTURE — an independent restatement (Python, from the Java text alone) of the control flow of BBMapThread.scoreSlow for the default flag set
(current/align2/BBMapThread.java:252-386; QUICK_MATCH_STRINGS off), gapped sites included (the fill then runs on makeGref's reference, tests/pygapped.py; setStop /
setLimits re-fix the gap array, tests/pysitelist.py): which sites are re-aligned, with which window and minScore, the
"more padding" retry, setSlowScore / setLimits, the minMsaLimit ratchet, the perfect / semiperfect bits.  Every alignment is MSA.fillAndScoreLimited restated in
tests/pygapped.py (fill by the reference's own C, walk by tests/pywalk.py).  Shares no code with oracle/scoreslow_oracle.c."""
import numpy as np

import pyclip
import pygapped
import pysitelist as ps

F = np.float32


def score_slow(oracle, packed, sites, basesP, basesM, ref8, cfg, maxR=601, maxC=3000):
    """sites: pysitelist.Site objects of one read (edited in place); basesP / basesM: numpy int8; ref8: numpy int8 chromosome array.  Returns the number of
    alignments requested."""
    L = len(basesP)
    max_sw = 70 + (L - 1) * 100
    max_imperfect = max_sw + min(-472, -395 - 100)
    ratio = cfg["min_ratio_pre_rescue"] if cfg["paired"] else cfg["min_ratio"]
    min_msa_limit = -int(cfg["clearzone1e"]) + int(F(ratio) * F(max_sw))
    limit = int(cfg["expected_len_limit"])
    pad0, extra = int(cfg["slow_align_padding"]), int(cfg["extra_padding"])
    ref_list = None
    fills = 0

    def align(ss, bases, pad, minscore):
        nonlocal fills
        fills += 1
        sv, _, _ = pygapped.fill_and_score_limited(oracle, packed, maxR, maxC, bases, ref8, ss.start - pad, ss.stop + pad, minscore, ss.gaps)
        return sv

    for ss in sites:
        bases = basesP if ss.strand == 0 else basesM
        if ss.stop - ss.start != L - 1:
            ss.set_slow_score(0)
            ss.semiperfect = ss.perfect = False
        no_indel = ss.slowScore
        arr = None
        if no_indel < max_imperfect and not ss.semiperfect:
            expected_len = ps.calc_gref_len(ss.start, ss.stop, ss.gaps)
            if expected_len >= limit:
                ps.set_stop(ss, ss.start + min(L + 40, limit))
            minscore = max(no_indel, min_msa_limit)
            arr = align(ss, bases, pad0, minscore)
            if arr is not None and len(arr) > 6 and arr[3] + arr[4] + expected_len < limit:
                old = list(arr)
                ps.set_limits(ss, ss.start - arr[6], ss.stop + arr[7])
                arr = align(ss, bases, pad0 + extra, minscore)
                if arr is None or arr[0] < old[0]:
                    arr = old
        if arr is not None:
            ss.set_slow_score(arr[0])
            ps.set_limits(ss, arr[1], arr[2])
        ss.score = ss.slowScore
        min_msa_limit = max(min_msa_limit, ss.slowScore - int(cfg["clearzone3"]))
        ss.perfect = ss.slowScore == max_sw
        if ss.perfect:
            ss.semiperfect = True
        elif not ss.semiperfect:
            if ref_list is None:
                ref_list = ref8.tolist()
            pyclip.set_perfect(ss, bases.tolist(), ref_list)
    return fills
