"""CPU: the C restatement of the index build + BBIndex.find (oracle/index_oracle.c, oracle/search_oracle.c) against what the
reference offers for this Java-only stage: the phiX truth-in-name fixture shipped with the reference
(tests/golden/phix.npz, made by tests/golden/make_phix_fixture.py) and structural invariants of the emitted sites."""
import os

import numpy as np
import pytest

from bbmap_b200.index import pack_chromosomes
from bbmap_b200.keyring import default_cfg

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phix.npz")


def phix():
    d = np.load(GOLD)
    cb, co, table = pack_chromosomes([d["genome"]])
    return d, cb, co, table


def top_site_correct(sites, ns, truth, thresh=20):
    """AbstractMapper 'correctthresh'-style check: strand equal and either end within `thresh` of the true locus."""
    if ns == 0:
        return False
    s = sites[:ns]; best = s[np.argmax(s["score"])]
    return bool(best["chrom"] == truth[0] and best["strand"] == truth[1] and
                (abs(int(best["start"]) - truth[2]) <= thresh or abs(int(best["stop"]) - truth[3]) <= thresh))


def test_phix_layout_matches_read_names():
    d, cb, co, table = phix()
    assert table == [(1, 8000, 5386)] and len(cb) == 8000 + 5386 + 8001       # 8000+5386+8001 bytes
    assert (d["r1_truth"][:, 2] - d["r1_truth"][:, 4] == 8000).all()         # chromosome coordinate = scaffold position + lead pad


@pytest.mark.parametrize("tag", ["r1", "r2"])
def test_phix_truth_in_name(oracle, tag):
    d, cb, co, table = phix()
    idx = oracle.index_build(cb, co, 13, -1)
    b, q, off, truth = d[tag + "_bases"], d[tag + "_qual"], d[tag + "_off"], d[tag + "_truth"]
    seeds = oracle.seed_batch(b, q, off, default_cfg(), 96)
    nk = seeds["nkeys"]
    assert (nk[nk > 0] <= 15).all() and (nk == 15).sum() >= 80                # 100 bp -> 15 keys
    for quit2 in (True, False):
        res = oracle.search_batch(idx, cb, co, b, seeds["baseScores"], off, seeds, quit_after_two_perfects=quit2)
        assert (res["status"] == 0).all()
        good = sum(top_site_correct(res["sites"][i], res["nsites"][i], truth[i]) for i in range(len(truth)))
        assert good >= 95, good
        for i in range(len(truth)):                                           # structural invariants of SiteScore
            s = res["sites"][i, :res["nsites"][i]]
            assert (s["stop"] >= s["start"]).all() and (s["start"] >= 0).all() and (s["stop"] < co[1]).all()
            assert (s["score"] <= res["max_score"][i]).all() and (s["hits"] >= 1).all() and (s["hits"] <= nk[i]).all()
            assert ((s["perfect"] == 0) | (s["semiperfect"] == 1)).all()
            assert ((s["perfect"] == 0) | (s["score"] == res["max_score"][i])).all()


def test_perfect_reads_score_max(oracle):
    """A read copied from the reference must come back as a perfect site at its origin with score == maxScore, on both strands."""
    from bbmap_b200 import workloads as wl
    rng = np.random.Generator(np.random.PCG64(5))
    g = wl.ACGT[rng.integers(0, 4, size=60000, dtype=np.uint8)]
    cb, co, table = pack_chromosomes([g])
    idx = oracle.index_build(cb, co, 13, -1)
    L, n = 150, 200
    pos = rng.integers(0, len(g) - L, size=n)
    reads = [g[p:p + L] if i % 2 == 0 else wl.revcomp(g[p:p + L]) for i, p in enumerate(pos)]
    bases = np.concatenate(reads); off = np.arange(n + 1, dtype=np.int64) * L
    qual = np.full(len(bases), 30, np.uint8)
    seeds = oracle.seed_batch(bases, qual, off, default_cfg(), 96)
    assert (seeds["nkeys"] == 18).all()                                        # 150 bp -> 18 keys
    res = oracle.search_batch(idx, cb, co, bases, seeds["baseScores"], off, seeds, quit_after_two_perfects=True)
    for i in range(n):
        assert res["nsites"][i] >= 1
        s = res["sites"][i, :res["nsites"][i]]; best = s[np.argmax(s["score"])]
        assert best["perfect"] == 1 and best["score"] == res["max_score"][i]
        assert best["start"] == 8000 + pos[i] and best["stop"] == 8000 + pos[i] + L - 1 and best["strand"] == i % 2
