"""GPU parity: Read.validate + reverse complement on the device (a0) vs the C restatement, all switch combinations, ragged
read lengths (incl. empty reads), every byte value."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _batch(seed, n, maxlen, weird=0.05):
    rng = np.random.Generator(np.random.PCG64(seed))
    lens = rng.integers(0, maxlen + 1, size=n); lens[rng.integers(0, n, size=3)] = 0
    off = np.zeros(n + 1, np.int64); np.cumsum(lens, out=off[1:])
    tot = int(off[-1])
    bases = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, size=tot)].copy()
    alpha = np.frombuffer(b"acgtuUNnXx-.*?RYKMSWBDHVrykmswbdhv @[`{", np.uint8)
    m = rng.random(tot) < weird
    bases[m] = alpha[rng.integers(0, len(alpha), size=int(m.sum()))]
    m = rng.random(tot) < weird / 5
    bases[m] = rng.integers(0, 256, size=int(m.sum())).astype(np.uint8)
    qual = rng.integers(-3, 60, size=tot).astype(np.int8)
    return bases, qual, off


@pytest.mark.parametrize("maxlen", [37, 151, 600, 5000])
def test_ingest_parity(oracle, maxlen):
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    from bbmap_b200 import reads
    msa = MultiStateAligner11tsCUDA(device=0)
    try:
        for flags in range(16):
            bases, qual, off = _batch(100 + flags, 3000 if maxlen < 1000 else 300, maxlen)
            for q in (qual, None):
                eb, eq, em, ef = oracle.ingest_batch(bases, q, off, flags)
                gb, gq, gm, gf = reads.validate_batch(msa.h, bases, q, off, flags)
                assert np.array_equal(gb, eb), (maxlen, flags)
                assert q is None or np.array_equal(gq, eq)
                assert np.array_equal(gm, em) and np.array_equal(gf, ef)
        assert ef.any() or flags & 1
    finally:
        msa.close()


def test_ingest_defaults_on_clean_reads(oracle):
    """Clean ACGT reads with Q in [2,41] are untouched and the minus strand is the textbook reverse complement."""
    from bbmap_b200.msa import MultiStateAligner11tsCUDA
    from bbmap_b200 import reads, workloads as wl
    rng = np.random.Generator(np.random.PCG64(3))
    n, L = 1000, 150
    bases = wl.ACGT[rng.integers(0, 4, size=n * L, dtype=np.uint8)]
    qual = rng.integers(2, 42, size=n * L).astype(np.int8)
    off = np.arange(n + 1, dtype=np.int64) * L
    msa = MultiStateAligner11tsCUDA(device=0)
    try:
        gb, gq, gm, gf = reads.validate_batch(msa.h, bases, qual, off)
    finally:
        msa.close()
    assert np.array_equal(gb.view(np.uint8), bases) and np.array_equal(gq, qual) and not gf.any()
    for r in (0, 17, n - 1):
        assert np.array_equal(gm.view(np.uint8)[r * L:(r + 1) * L], wl.revcomp(bases[r * L:(r + 1) * L]))
