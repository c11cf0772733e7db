"""CPU: index construction and analysis — the C restatement the CUDA index build is checked against (oracle/index_oracle.c) must equal a
second restatement written from the Java text alone (tests/pyindex.py, numpy): every block's `starts` and `sites`, COUNTS, the 1001-entry
lengthHistogram, MAX_USABLE_LENGTH / MAX_USABLE_LENGTH2 / POINTS_PER_SITE."""
import os

import numpy as np
import pytest

from bbmap_b200 import workloads as wl
from bbmap_b200.index import pack_chromosomes

import pyindex


def _genome(seed, sizes):
    rng = np.random.Generator(np.random.PCG64(seed))
    scafs = []
    for n in sizes:
        s = wl.ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)].copy()
        if n >= 30000:
            s[1000:1400] = ord("A")                                              # period 1: banned keys
            s[1500:1900] = np.tile(np.frombuffer(b"AC", np.uint8), 200)          # period 2: banned keys
            s[2000:11000] = np.tile(np.frombuffer(b"ACG", np.uint8), 3000)       # period 3: > 2000 sites, clumpy -> COUNTS zeroed
            s[11500:14000] = np.tile(np.frombuffer(b"ACGTT", np.uint8), 500)     # period 5: clumps, but the lists stay below 2000
            s[14100:14130] = ord("N")
            s[14200:14230] = np.frombuffer(b"acgtuACGTURYKMacgtnnACGTacgtACG", np.uint8)[:30]   # lower case and U index like upper case; IUPAC does not
            unit = s[100:400].copy()
            for r in range(5):
                p = int(rng.integers(15000, n - 400)); s[p:p + 300] = unit
        scafs.append(s)
    return scafs


def _compare(oracle, bytes_, off, k, chrombits):
    ecfg, eblocks, ecounts, ehist = oracle.index_build(bytes_, off, k, chrombits)
    c = ecfg[0]
    cb = int(c["chrombits"]); cpb = int(c["chroms_per_block"])
    assert cpb == 1 << cb and int(c["shift_length"]) == 31 - cb
    nch = len(off) - 1
    blocks = []
    chrom = 1
    while chrom <= nch:
        base = chrom & ~(cpb - 1)
        a, bmax = max(1, base), min(nch, base + cpb - 1)
        chroms = [bytes_[int(off[i - 1]): int(off[i])] for i in range(a, bmax + 1)]
        blocks.append(pyindex.build_block(chroms, a, k, cb))
        chrom = bmax + 1
    assert len(blocks) == len(eblocks)
    for (gs, gt), (es, et) in zip(blocks, eblocks):
        assert np.array_equal(gs, es)
        assert np.array_equal(gt, et)
    counts, hist, mul, mul2, pps = pyindex.analyze(blocks, k, float(c["fraction_to_exclude"]), int(c["max_average_list_to_search"]))
    assert np.array_equal(counts, ecounts)
    assert hist == ehist.tolist()
    assert (mul, mul2, pps) == (int(c["max_usable_length"]), int(c["max_usable_length2"]), int(c["points_per_site"]))
    return eblocks, ecounts


@pytest.mark.parametrize("k,sizes,chrombits,maxlen", [(10, (60000, 5000, 80000), -1, 120000), (11, (30000,) * 5, 1, 120000), (9, (3000,), 0, None),
                                                      (8, (40000, 40000, 40000), 2, 60000)])
def test_index_equals_independent_restatement(oracle, k, sizes, chrombits, maxlen):
    scafs = _genome(400 + k, sizes)
    bytes_, off, _ = pack_chromosomes(scafs, max_length=maxlen if maxlen else (1 << 29) - 200000)
    eblocks, ecounts = _compare(oracle, bytes_, off, k, chrombits)
    if sizes[0] >= 30000 and k >= 10:
        # the period-3 keys were indexed (long lists) and then dropped as clumpy
        key = 0
        for ch in (b"ACG" * 5)[:k]:
            key = (key << 2) | b"ACGT".index(ch)
        lens = sum(int(s[key + 1]) - int(s[key]) for s, _ in eblocks)
        assert lens > 2000 and ecounts[key] == 0


def test_index_phix_k13(oracle):
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phix.npz"))
    bytes_, off, _ = pack_chromosomes([d["genome"]])
    _compare(oracle, bytes_, off, 13, -1)
