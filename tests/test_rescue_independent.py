"""CPU: mate rescue — AbstractMapThread.rescue with quickRescue, findTipDeletions and slowRescue of the C restatement (oracle/mapper_oracle.c, through
its test entry point) must equal a second restatement written from the Java text (tests/pyrescue.py; the scan stated over all starts at once, every fill by the
reference's own C): the rescued sites appended to the loose read's list (every field), the paired scores set on the anchor's sites, the number of scans and fills."""
import ctypes as C

import numpy as np
import pytest

from bbmap_b200 import rescue as rs
from bbmap_b200 import sitelist as sl
from bbmap_b200 import workloads as wl
from bbmap_b200.mapper import map_cfg

import pyrescue
from test_sitelist_independent import _same, _to_sites


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _mutate(rng, read, nsub):
    read = read.copy()
    for _ in range(nsub):
        k = int(rng.integers(0, len(read)))
        read[k] = wl.ACGT[(int(np.searchsorted(wl.ACGT, min(read[k], ord("T")))) + int(rng.integers(1, 4))) % 4]
    return read


@pytest.mark.parametrize("seed,kw", [(701, {}), (702, dict(average_pair_dist=300)), (703, dict(same_strand_pairs=1)), (704, dict(max_rescue_mismatches=8))])
def test_rescue(oracle, seed, kw):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    rng = np.random.default_rng(seed)
    genome = wl.random_genome(40000, seed=seed).copy()
    genome[:300] = ord("N"); genome[-300:] = ord("N")
    for _ in range(6):                                            # repeats: several placements inside one search range
        a = int(rng.integers(2000, 30000)); b = a + int(rng.integers(200, 700)); genome[b:b + 150] = genome[a:a + 150]
    g8 = genome.view(np.int8)
    co = np.array([0, len(genome)], np.int64)
    cfg = map_cfg(paired=1, **kw)
    tc = rs.tipdel_cfg()
    lib = oracle.lib
    lib.orc_test_rescue.restype = C.c_int
    R = pyrescue.Rescuer(oracle, g8, cfg[0], search_range=int(tc["search_range"][0]), slow_rescue_padding=int(tc["slow_rescue_padding"][0]))
    same = bool(cfg["same_strand_pairs"][0])
    added = paired = tipped = 0
    for it in range(500):
        L1, L2 = (int(x) for x in rng.choice([50, 100, 150], size=2))
        f = int(rng.integers(1500, len(genome) - 3000))
        ins = int(rng.integers(L1 + 10, 650)) if rng.random() < 0.85 else int(rng.integers(1300, 2500))
        astrand = int(rng.integers(0, 2))
        # anchor on the plus strand at f: mate at f+ins-L2 (to the right); anchor on the minus strand: mate to the left
        apos = f if astrand == 0 else f + ins - L1
        mpos = f + ins - L2 if astrand == 0 else f
        mate = genome[mpos:mpos + L2].copy()
        kind = rng.random()
        if kind < 0.3:
            mate = _mutate(rng, mate, int(rng.integers(1, 6)))
        elif kind < 0.45:
            mate = _mutate(rng, mate, int(rng.integers(8, 40)))
        elif kind < 0.65 and L2 > 40:                            # a deletion near a tip: findTipDeletions territory
            t = int(rng.integers(4, 9)); d = int(rng.integers(2, 60))
            mate = np.concatenate([genome[mpos:mpos + L2 - t], genome[mpos + L2 - t + d: mpos + L2 + d]]) if rng.random() < 0.5 else \
                   np.concatenate([genome[mpos - d: mpos - d + t], genome[mpos + t: mpos + L2]])
        elif kind < 0.72:
            mate = wl.ACGT[rng.integers(0, 4, size=L2, dtype=np.uint8)]
        if rng.random() < 0.1:
            mate[int(rng.integers(0, L2))] = ord("N")
        fwd = np.ascontiguousarray(mate, np.uint8).view(np.int8)
        rev = wl.revcomp(np.ascontiguousarray(mate, np.uint8)).copy().view(np.int8)
        # `bases` the search uses must read like the plus strand at the mate's locus
        mate_strand = astrand if same else astrand ^ 1
        basesP, basesM = (fwd, rev) if mate_strand == 0 else (rev, fwd)
        qual = rng.integers(2, 41, size=L2).astype(np.int8) if rng.random() < 0.7 else None
        nA = int(rng.integers(1, 4))
        A = np.zeros(4, sl.SS_DTYPE)
        maxa = 70 + 100 * (L1 - 1)
        for i in range(nA):
            A[i]["chrom"] = 1; A[i]["strand"] = astrand if i == 0 else int(rng.integers(0, 2))
            A[i]["start"] = apos if i == 0 else int(rng.integers(1500, len(genome) - 3000)); A[i]["stop"] = A[i]["start"] + L1 - 1 + (int(rng.integers(0, 4)) if rng.random() < 0.2 else 0)
            sc = maxa - int(rng.integers(0, 400)) * i - (0 if rng.random() < 0.4 else int(rng.integers(0, 900)))
            A[i]["slow_score"] = sc; A[i]["score"] = sc; A[i]["quick_score"] = sc // 2
            A[i]["paired_score"] = sc + 50 if rng.random() < 0.1 else 0; A[i]["rescued"] = 1 if rng.random() < 0.05 else 0
        order = np.argsort(-A["slow_score"][:nA], kind="stable"); A[:nA] = A[:nA][order]
        cap = 8
        Lst = np.zeros(cap, sl.SS_DTYPE); nL = np.array([int(rng.integers(0, 3))], np.int32)
        maxl = 70 + 100 * (L2 - 1)
        for i in range(int(nL[0])):
            Lst[i]["chrom"] = 1; Lst[i]["strand"] = int(rng.integers(0, 2)); Lst[i]["start"] = int(rng.integers(1500, 30000)); Lst[i]["stop"] = Lst[i]["start"] + L2 - 1
            sc = int(maxl * rng.uniform(0.3, 1.0)) - 300 * i
            Lst[i]["slow_score"] = sc; Lst[i]["score"] = sc
        search_dist = min(int(cfg["max_pair_dist"][0]), 2 * int(cfg["average_pair_dist"][0]) + 100)
        sa = _to_sites(A, nA); sl_ = _to_sites(Lst, int(nL[0]))
        counts = np.zeros(2, np.int64)
        a2 = A.copy(); l2 = Lst.copy(); n2 = nL.copy()
        st = lib.orc_test_rescue(_p(a2), C.c_int(nA), C.c_int(L1), _p(l2), _p(n2), C.c_int(cap), _p(np.ascontiguousarray(basesP)), _p(np.ascontiguousarray(basesM)),
                                 _p(qual), C.c_int(L2), C.c_int(search_dist), _p(g8), _p(co), _p(cfg), _p(tc), C.c_int(258), _p(counts))
        assert st == 0
        s0, f0 = R.scans, R.fills
        R.rescue(sa, L1, sl_, np.ascontiguousarray(basesP), np.ascontiguousarray(basesM), qual, search_dist)
        _same(sa, a2, nA, (it, "anchor"))
        _same(sl_, l2, int(n2[0]), (it, "loose"))
        assert (R.scans - s0, R.fills - f0) == (int(counts[0]), int(counts[1])), (it, R.scans - s0, R.fills - f0, counts)
        added += len(sl_) - int(nL[0]); paired += sum(1 for s in sa if s.pairedScore > 0)
        tipped += sum(1 for s in sl_[int(nL[0]):] if s.stop - s.start + 1 != L2)
    assert added > 100 and paired > 60 and tipped > 4, (added, paired, tipped)
