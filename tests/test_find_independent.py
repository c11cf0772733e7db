"""CPU: BBIndex.find — the C restatement (oracle/search_oracle.c, which the CUDA kernel is tested against) must equal a second,
structurally independent restatement written from the Java text alone (tests/pyfind.py: a real binary QuadHeap of Quad objects, one heap pop at
a time, Java-order loops, numpy float32 for Java float) on every emitted SiteScore — chrom, strand, start, stop, hits, score, perfect,
semiperfect, the gap array — and on bestScores[6], maxScore, maxQuickScore and the number of hit lists kept by the key filter."""
import os

import numpy as np
import pytest

from bbmap_b200 import workloads as wl
from bbmap_b200.index import pack_chromosomes
from bbmap_b200.keyring import default_cfg

import pyfind

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phix.npz")


def _compare(oracle, cb, co, bases, qual, off, k=13, chrombits=-1, quit2=True, max_length=None, expect_sites=1):
    idx = oracle.index_build(cb, co, k, chrombits)
    seeds = oracle.seed_batch(bases, qual, off, default_cfg(), 96)
    res = oracle.search_batch(idx, cb, co, bases, seeds["baseScores"], off, seeds, quit_after_two_perfects=quit2)
    cfg, blocks, counts, hist = idx
    py = pyfind.BBIndexPy(cfg, blocks, counts, hist, cb, co, quit_after_two_perfects=quit2)
    nsites = gapped = multi = walked = 0
    for i in range(len(off) - 1):
        nk = int(seeds["nkeys"][i])
        if nk < 1:
            continue
        a, b = int(off[i]), int(off[i + 1])
        r = py.find(bases[a:b].tobytes(), seeds["baseScores"][a:b].view(np.int8), seeds["offsets"][i, :nk], seeds["keyScores"][i, :nk])
        e = res[i]
        if e["status"] & 8:                         # subsumption into a gapped site: both sides stop there by design
            assert r["gapfix"]
            continue
        assert e["status"] == 0 and not r["gapfix"], (i, e["status"])
        assert r["num_hits"] == e["num_hits"], (i, r["num_hits"], e["num_hits"])
        if r["best_scores"] is None:                # find() returned before the walks
            assert e["nsites"] == 0
            continue
        walked += 1
        assert r["max_score"] == e["max_score"] and r["max_quick_score"] == e["max_quick_score"], i
        assert list(r["best_scores"]) == e["best_scores"].tolist(), (i, r["best_scores"], e["best_scores"])
        assert len(r["sites"]) == e["nsites"], (i, len(r["sites"]), e["nsites"])
        for s, t in zip(r["sites"], e["sites"][: e["nsites"]]):
            got = (s.chrom, s.strand, s.start, s.stop, s.hits, s.score, int(s.perfect), int(s.semiperfect))
            exp = (int(t["chrom"]), int(t["strand"]), int(t["start"]), int(t["stop"]), int(t["hits"]), int(t["score"]), int(t["perfect"]), int(t["semiperfect"]))
            assert got == exp, (i, got, exp)
            g = [] if s.gaps is None else list(s.gaps)
            assert g == t["gaps"][: t["ngaps"]].tolist(), (i, g, t["gaps"], t["ngaps"])
            gapped += len(g) > 0
        nsites += len(r["sites"]); multi += len(r["sites"]) > 1
    assert walked >= 0.8 * (len(off) - 1) and nsites >= expect_sites * walked * 0.9
    return nsites, gapped, multi


@pytest.mark.parametrize("quit2", [True, False])
def test_phix_shipped_reads(oracle, quit2):
    """configs[0]: the 2x100 reads shipped with the reference (tests/golden/phix.npz), qualities as shipped."""
    d = np.load(GOLD)
    cb, co, table = pack_chromosomes([d["genome"]])
    for tag in ("r1", "r2"):
        _compare(oracle, cb, co, d[tag + "_bases"], d[tag + "_qual"], d[tag + "_off"], quit2=quit2)


def _reads(g, rng, n, L, sub=0.02, indel=0.3, ns=0.05):
    reads = []
    for i in range(n):
        p = int(rng.integers(0, len(g) - L - 40))
        r = g[p:p + L + 30].copy()
        if rng.random() < indel:                                   # one short insertion or deletion
            q = int(rng.integers(20, L - 20)); d = int(rng.integers(1, 12))
            r = np.concatenate([r[:q], r[q + d:]]) if rng.random() < 0.5 else np.concatenate([r[:q], wl.ACGT[rng.integers(0, 4, size=d, dtype=np.uint8)], r[q:]])
        r = r[:L].copy()
        m = rng.random(L) < sub
        r[m] = wl.ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        if rng.random() < ns:
            r[int(rng.integers(0, L))] = ord("N")
        reads.append(r if i % 2 == 0 else wl.revcomp(r))
    bases = np.concatenate(reads); off = np.arange(n + 1, dtype=np.int64) * L
    qual = rng.integers(2, 41, size=len(bases)).astype(np.uint8)
    qual[bases == ord("N")] = 0
    return bases, qual, off


@pytest.mark.parametrize("quit2", [True, False])
def test_planted_repeats_one_block(oracle, quit2):
    """Repeat families (ties, subsumption classes 1-3, two perfect sites -> QUIT_AFTER_TWO_PERFECTS), varied qualities, N, short indels."""
    rng = np.random.Generator(np.random.PCG64(11))
    g = wl.ACGT[rng.integers(0, 4, size=120000, dtype=np.uint8)]
    unit = g[5000:5400].copy()
    for c in range(6):
        p = 12000 + 15000 * c; g[p:p + 400] = unit
        if c % 2:
            g[p + 200] = wl.ACGT[(int(np.searchsorted(wl.ACGT, g[p + 200])) + 1) & 3]          # near-identical copies
    g[70000:70300] = g[70300:70600]                                                         # tandem duplication: overlapping candidate sites
    cb, co, table = pack_chromosomes([g])
    bases, qual, off = _reads(g, rng, 260, 150)
    # a third of the reads from inside the repeat families
    for i in range(0, 260, 3):
        p = 12000 + 15000 * int(rng.integers(0, 6)) + int(rng.integers(0, 250))
        r = g[p:p + 150]
        bases[off[i]:off[i + 1]] = r if i % 2 == 0 else wl.revcomp(r)
    nsites, gapped, multi = _compare(oracle, cb, co, bases, qual, off, quit2=quit2)
    assert multi >= 40, (nsites, gapped, multi)


def test_several_blocks_and_spliced_reads(oracle):
    """Five chromosomes in three index blocks (chrombits 1): bestScores carried from block to block and strand to strand; reads spliced over
    300-3000 bp introns so that makeGapArray runs; k=11 keeps the hit lists long."""
    rng = np.random.Generator(np.random.PCG64(12))
    scafs = [wl.ACGT[rng.integers(0, 4, size=30000, dtype=np.uint8)] for _ in range(5)]
    scafs[3][1000:1300] = scafs[0][2000:2300]                                                # the same 300-mer on chromosomes of different blocks
    scafs[4][5000:5300] = wl.revcomp(scafs[0][2000:2300])
    cb, co, table = pack_chromosomes(scafs, max_length=50000)
    assert len(co) - 1 == 5
    L = 150
    parts = []
    for s in (0, 2, 3, 4):
        b_, q_, o_ = _reads(scafs[s], rng, 40, L, sub=0.01)
        parts.append(b_)
    spl = []
    for i in range(40):                                                                      # spliced: 75 + intron + 75
        g = scafs[i % 5]; p = int(rng.integers(100, 20000)); gap = int(rng.integers(300, 3000))
        r = np.concatenate([g[p:p + 75], g[p + 75 + gap:p + 150 + gap]])
        spl.append(r if i % 2 == 0 else wl.revcomp(r))
    rep = [scafs[0][2000 + j:2150 + j] for j in range(0, 150, 15)]
    bases = np.concatenate(parts + spl + rep)
    n = len(bases) // L; off = np.arange(n + 1, dtype=np.int64) * L
    qual = np.full(len(bases), 30, np.uint8)
    nsites, gapped, multi = _compare(oracle, cb, co, bases, qual, off, k=11, chrombits=1, quit2=False)
    assert gapped >= 20 and multi >= 10, (nsites, gapped, multi)
