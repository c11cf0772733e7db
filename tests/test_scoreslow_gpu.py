"""GPU parity (part of SURVEY §8 f1): BBMapThread.scoreSlow in rounds on the device vs the sequential C restatement — every field of every
site, the per-read status bits and the number of alignments requested (first pass + padding retries)."""
import numpy as np
import pytest

from bbmap_b200 import sitelist as sl
from sitelist_cases import slow_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def msa():
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    m = MultiStateAligner11tsCUDA(device=0)
    yield m
    m.close()


@pytest.mark.parametrize("seed,kw", [(81, {}), (82, {}), (83, dict(extra_padding=0)), (84, dict(paired=1, min_ratio_pre_rescue=0.4, clearzone3=0))])
def test_score_slow_parity(oracle, msa, seed, kw):
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=1500, seed=seed)
    lists, _, _ = oracle.sitelist(sl.SL_NOINDEL, lists, nss, ro, sl.policy_cfg(), P, M, refs, co)
    cfg = sl.slow_cfg(**kw)
    exp, est, ena = oracle.score_slow(lists, nss, ro, P, M, refs, co, run, cfg)
    msa.set_option("strip_min_tasks", 0 if seed % 2 == 0 else 8192)      # both routings of the aligner: thread-per-alignment strips and warp-per-alignment tiles
    d_ref = msa.load_reference(refs)
    try:
        got, gst, gna = sl.scoreSlow(msa.h, lists, nss, ro, P, M, d_ref, co, run, cfg)
    finally:
        msa.free(d_ref)
    assert np.array_equal(gst, est) and gna == ena
    live = np.arange(exp.shape[1])[None, :] < nss[:, None]
    for f in exp.dtype.names:
        assert np.array_equal(got[f][live], exp[f][live]), f
    changed = (exp["slow_score"][live] != lists["slow_score"][live]).sum()
    assert changed > 500 and ena > 1500


def test_score_slow_nothing_to_do(msa):
    refs, co, P, M, ro, lists, nss, run = slow_cases(nreads=50, seed=85)
    d_ref = msa.load_reference(refs)
    try:
        got, st, na = sl.scoreSlow(msa.h, lists, nss, ro, P, M, d_ref, co, np.zeros(50, np.int32))
        got2, st2, na2 = sl.scoreSlow(msa.h, lists[:0], nss[:0], ro[:1], P, M, d_ref, co, run[:0])
    finally:
        msa.free(d_ref)
    assert na == 0 and got.tobytes() == lists.tobytes() and na2 == 0 and len(got2) == 0
