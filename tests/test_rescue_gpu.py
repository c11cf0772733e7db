"""GPU parity: findTipDeletions and quickRescue on the device vs the C restatement, every output field bit-exact, on
seeded cases that cover right/left/both tips, unrelated reads, sites at the ends of the chromosome array and next to N blocks,
already-extended sites, planted repeats (ties, perfect hits that shrink the scan), clipped ranges, short reads and empty batches."""
import numpy as np
import pytest

from bbmap_b200 import rescue as rs
from rescue_cases import rescue_cases, tipdel_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def msa():
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    m = MultiStateAligner11tsCUDA(device=0)
    yield m
    m.close()


@pytest.mark.parametrize("seed,search_range", [(31, 100), (32, 150), (33, 20)])
def test_tipdel_parity(oracle, msa, seed, search_range):
    g, reads, tasks = tipdel_cases(n=6000, seed=seed)
    cfg = rs.tipdel_cfg(search_range=search_range)
    exp = oracle.tipdel_batch(reads, g, tasks, cfg)
    d_ref = msa.load_reference(g)
    try:
        got = rs.findTipDeletions(msa.h, reads, d_ref, tasks, cfg)
    finally:
        msa.free(d_ref)
    for f in exp.dtype.names:
        assert np.array_equal(got[f], exp[f]), f
    lim = 300 if search_range >= 100 else 100
    assert (exp["right"] > 0).sum() > lim and (exp["left"] > 0).sum() > lim


@pytest.mark.parametrize("seed", [37, 38])
def test_rescue_parity(oracle, msa, seed):
    g, reads, tasks = rescue_cases(n=3000, seed=seed)
    cfg = rs.rescue_cfg()
    exp = oracle.rescue_batch(reads, g, tasks, cfg)
    d_ref = msa.load_reference(g)
    try:
        got = rs.quickRescue(msa.h, reads, d_ref, tasks, cfg)
        none = rs.quickRescue(msa.h, reads, d_ref, tasks[:0], cfg)
    finally:
        msa.free(d_ref)
    assert len(none) == 0
    for f in exp.dtype.names:
        assert np.array_equal(got[f], exp[f]), f
    assert (exp["start"] >= 0).sum() > 1000 and (exp["start"] < 0).sum() > 300 and (exp["perfect"] == 3).sum() > 200
    assert (exp["mismatches"] > 5).any()


def test_rescue_non_affine_score(oracle, msa):
    g, reads, tasks = rescue_cases(n=500, seed=39)
    cfg = rs.rescue_cfg(use_affine=0)
    exp = oracle.rescue_batch(reads, g, tasks, cfg)
    d_ref = msa.load_reference(g)
    try:
        got = rs.quickRescue(msa.h, reads, d_ref, tasks, cfg)
    finally:
        msa.free(d_ref)
    assert np.array_equal(got["score"], exp["score"]) and np.array_equal(got["start"], exp["start"])
