"""CPU: score2 / traceback2 pinned by an INDEPENDENT walk over the matrix the reference's own C filled.

For every task the unmodified reference C (oracle/_ref/libbbref.so) fills `packed`; tests/pywalk.py — plain Python written from
MultiStateAligner11tsJNI.java:376-495,537-658 — walks it, and the result must equal what the C restatement
(oracle/msa_oracle.c via orc_batch_run) reports for the same task: the score vector (6 or 8 ints, padding suggestions included)
and the match string.  >= 10^4 random G4 alignments (true locus, planted indels up to 40 bp, unrelated loci, N bases)."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl
from oracle.oracle import TF_RAW_LIMITED, TF_RAW_UNLIMITED, TF_SCORE, TF_TRACEBACK, match_offsets

import pywalk

MAXR, MAXC = 260, 340


def _walk_all(oracle, genome, reads, tasks, limited):
    if limited:
        tasks["min_score"] = np.maximum(tasks["min_score"] - 120, 1)          # what the Java wrapper hands to fillLimitedX (MSA11tsJNI.java:144)
    outs, mbuf, _ = oracle.run_batch(reads, genome, tasks, maxRows=MAXR, maxColumns=MAXC, kind="port")
    moff = match_offsets(tasks)
    packed = oracle.new_packed(MAXR, MAXC)
    M = pywalk.Matrix(memoryview(packed).cast("B").cast("i"), MAXR, MAXC)
    g8 = genome.view(np.int8); gl = genome.tobytes()
    walked = padded = indel = 0
    for i, t in enumerate(tasks):
        r = reads[t["read_off"]: t["read_off"] + t["read_len"]]
        a, b = int(t["ref_start"]), int(t["ref_end"])
        rows, cols = int(t["read_len"]), b - a + 1
        if limited:
            res, _ = oracle.fill_limited(r.view(np.int8), g8, a, b, int(t["min_score"]), packed, MAXR, MAXC, kind="reference")
            assert res.tolist() == outs["result"][i].tolist()
            if res[4]:
                assert outs["score_len"][i] == 0 and outs["match_len"][i] == -1
                continue
        else:
            res, _ = oracle.fill_unlimited(r.view(np.int8), g8, a, b, packed, MAXR, MAXC, kind="reference")
            assert res.tolist() == outs["result"][i][:4].tolist()
        maxRow, maxCol, maxState = int(res[0]), int(res[1]), int(res[2])
        sv = pywalk.score2(M, rows, cols, a, b, maxRow, maxCol, maxState)
        assert len(sv) == outs["score_len"][i] and sv == outs["score"][i][:len(sv)].tolist(), (i, sv, outs["score"][i])
        ms = pywalk.traceback2(M, r.tobytes(), gl, cols, a, maxRow, maxCol, maxState)
        got = mbuf[moff[i]: moff[i] + outs["match_len"][i]].tobytes()
        assert ms == got, (i, ms, got)
        walked += 1; padded += len(sv) == 8; indel += (b"D" in ms) or (b"I" in ms)
    return walked, padded, indel


def test_limited_fills_walked_independently(oracle):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    genome = wl.random_genome(60000, seed=21)
    reads, tasks = wl.make_msa_tasks(genome, 10000, seed=31, flags=TF_RAW_LIMITED | TF_SCORE | TF_TRACEBACK, n_rate=0.002)
    walked, padded, indel = _walk_all(oracle, genome, reads, tasks, True)
    assert walked >= 8000 and indel >= 1000, (walked, padded, indel)


def test_loose_limits_walked_independently(oracle):
    """The paired pre-rescue ratio (0.336): wide ragged bands; windows without padding, so alignments run into the window edges (padLeft/padRight)."""
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    genome = wl.random_genome(60000, seed=22)
    reads, tasks = wl.make_msa_tasks(genome, 1500, seed=32, flags=TF_RAW_LIMITED | TF_SCORE | TF_TRACEBACK, ratio=0.336, tight=False, pad=0)
    walked, padded, indel = _walk_all(oracle, genome, reads, tasks, True)
    assert walked >= 1200 and padded >= 10, (walked, padded, indel)          # pad=0 windows: the 8-int vector with padding suggestions


def test_unlimited_fills_walked_independently(oracle):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    genome = wl.random_genome(60000, seed=23)
    reads, tasks = wl.make_msa_tasks(genome, 1500, seed=33, flags=TF_RAW_UNLIMITED | TF_SCORE | TF_TRACEBACK, lengths=(100, 150))
    walked, padded, indel = _walk_all(oracle, genome, reads, tasks, False)
    assert walked == 1500
