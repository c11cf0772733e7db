"""CPU: the ingest and seeding rows — the C restatement the CUDA kernels are checked against (oracle/host_oracle.c) must equal a second
restatement written from the Java text alone (tests/pyseed.py: plain Python lists, numpy float32 scalars for Java float, its own tables) on every
output: validated bases / qualities / junk flag / minus-strand bases; number of seeds, offsets, keys, key scores, base scores, minus-strand offsets and keys."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl
from bbmap_b200.keyring import default_cfg

import pyseed


def _s8(a):
    return [int(x) for x in np.asarray(a).view(np.int8)]


def _seed_compare(oracle, bases, qual, off, max_keys=96):
    cfg = default_cfg()
    exp = oracle.seed_batch(bases, qual, off, cfg, max_keys)
    c = cfg[0]
    seeded = 0
    for r in range(len(off) - 1):
        a, b = int(off[r]), int(off[r + 1])
        got = pyseed.quick_map_seed(_s8(bases[a:b]), None if qual is None else _s8(qual[a:b]), int(c["keylen"]), int(c["maxDesiredKeys"]),
                                    int(c["baseKeyHitScore"]), int(c["minApproxHitsToKeep"]), float(c["keyDensity"]), float(c["maxKeyDensity"]),
                                    float(c["minKeyDensity"]))
        n = int(exp["nkeys"][r])
        if got is None:
            assert n <= 0, (r, n)
            continue
        seeded += 1
        assert n == len(got["offsets"]), (r, n, got["offsets"])
        for name in ("offsets", "keys", "keyScores", "offsetsM", "keysM"):
            assert exp[name][r, :n].tolist() == got[name], (r, name, exp[name][r, :n].tolist(), got[name])
        assert _s8(exp["baseScores"][a:b]) == got["baseScores"], r
    return seeded, exp


def test_seeding_random_reads(oracle):
    bases, qual, off = wl.make_read_batch(1500, seed=191)
    seeded, exp = _seed_compare(oracle, bases, qual, off)
    assert seeded > 1200 and len(np.unique(exp["nkeys"])) > 3


def test_seeding_reference_key_counts(oracle):
    # 150 bp flat-Q30 reads get 18 seeds, 100 bp reads 15
    for length, want, seed in ((150, 18, 192), (100, 15, 193)):
        b, q, o = wl.make_read_batch(60, seed=seed, lengths=(length,), flat_q=30, n_rate=0)
        seeded, exp = _seed_compare(oracle, b, q, o)
        assert seeded == 60 and (exp["nkeys"] == want).all()


def test_seeding_edge_cases(oracle):
    bases, qual, off = wl.make_read_batch(700, seed=194, lengths=(12, 13, 14, 30, 64, 300, 600))
    _seed_compare(oracle, bases, qual, off)
    _seed_compare(oracle, bases, None, off)                       # FASTA path: no qualities, keys over N are -1
    b = bases.copy(); q = qual.copy()
    for r in range(0, 700, 7):
        b[off[r]: off[r + 1]] = ord("N"); q[off[r]: off[r + 1]] = 0
    seeded, exp = _seed_compare(oracle, b, q, off)
    assert (exp["nkeys"][::7] <= 0).all()
    q[:] = 2
    _seed_compare(oracle, b, q, off)                              # probAllErrors / avgQuality gates
    rng = np.random.default_rng(195)
    q2 = rng.integers(0, 42, len(q)).astype(np.int8)              # ragged qualities incl. zeros on defined bases
    _seed_compare(oracle, bases, q2, off)


def test_seeding_phix_shipped_reads(oracle):
    import os
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "phix.npz"))
    total = 0
    for m in ("r1", "r2"):
        seeded, _ = _seed_compare(oracle, d[m + "_bases"], d[m + "_qual"], d[m + "_off"])
        total += seeded
    assert total >= 190


@pytest.mark.parametrize("flags", range(16))
def test_validate_all_switches(oracle, flags):
    """Read.validate with the 16 combinations of FIX_JUNK / U_TO_T / TO_UPPER_CASE / LOWER_CASE_TO_N, every ASCII byte value, with and without qualities."""
    rng = np.random.default_rng(300 + flags)
    n = 120
    lens = rng.integers(0, 90, n)
    off = np.zeros(n + 1, np.int64); off[1:] = np.cumsum(lens)
    bases = rng.integers(0, 128, int(off[-1])).astype(np.int8)
    common = np.frombuffer(b"ACGTNacgtnUuXx.-*RYKMSWBDHV", np.int8)
    pick = rng.random(len(bases)) < 0.7
    bases[pick] = common[rng.integers(0, len(common), int(pick.sum()))]
    qual = rng.integers(-3, 60, len(bases)).astype(np.int8)
    for quality in (qual, None):
        eb, eq, ebm, efl = oracle.ingest_batch(bases, quality, off, flags)
        for r in range(n):
            a, b = int(off[r]), int(off[r + 1])
            gb, gq, junk = pyseed.validate(_s8(bases[a:b]), None if quality is None else _s8(quality[a:b]), fix_junk=bool(flags & 1),
                                           u_to_t=bool(flags & 2), to_upper=bool(flags & 4), lower_to_n=bool(flags & 8))
            assert _s8(eb[a:b]) == gb, (flags, r, bytes(bases[a:b].view(np.uint8)), _s8(eb[a:b]), gb)
            if quality is not None:
                assert _s8(eq[a:b]) == gq, (flags, r)
            assert bool(efl[r] & 1) == junk, (flags, r)
            if all(0 <= x < 128 and pyseed.BASE_TO_COMP_EXT[x] >= 0 for x in gb):
                assert _s8(ebm[a:b]) == pyseed.reverse_complement_bases(gb), (flags, r)
