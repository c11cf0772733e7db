"""GPU parity: KeyRing seeding kernel vs the C restatement of QualityTools/KeyRing (bit-exact ints, incl. float-driven
choices).  The restatement itself is 'parity unpinned' against Java (no JVM); the reference's own invariants are checked."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kr():
    from bbmap_b200.keyring import KeyRingCUDA
    k = KeyRingCUDA()
    yield k
    k.close()


def _check(oracle, kr, bases, qual, off, maxKeys=96):
    from bbmap_b200.keyring import default_cfg
    cfg = default_cfg()
    exp = oracle.seed_batch(bases, qual, off, cfg, maxKeys)
    got = kr.seed_batch(bases, qual, off, cfg, maxKeys)
    for k in ("nkeys", "offsets", "keys", "keyScores", "baseScores", "offsetsM", "keysM"):
        if not np.array_equal(got[k], exp[k]):
            bad = np.nonzero((got[k] != exp[k]).reshape(len(got[k]), -1).any(axis=1))[0] if got[k].ndim > 1 else np.nonzero(got[k] != exp[k])[0]
            raise AssertionError("%s differs at %s: got %s exp %s" % (k, bad[:5], got[k][bad[0]], exp[k][bad[0]]))
    return exp


def test_seed_random(oracle, kr):
    bases, qual, off = wl.make_read_batch(20000, seed=91)
    exp = _check(oracle, kr, bases, qual, off)
    n = exp["nkeys"]
    assert (n > 0).mean() > 0.9 and len(np.unique(n)) > 3
    # reference invariants: offsets strictly ascending, in range (BBIndex.checkOffsets, current/align2/BBIndex.java:200-205)
    lens = np.diff(off)
    for r in np.nonzero(n > 0)[0][:2000]:
        o = exp["offsets"][r, : n[r]]
        assert (np.diff(o) > 0).all() and o[0] >= 0 and o[-1] + 13 <= lens[r]
    # 150 bp flat Q30 reads get 18 seeds, 100 bp get 15
    b2, q2, o2 = wl.make_read_batch(200, seed=92, lengths=(150,), flat_q=30, n_rate=0)
    assert (_check(oracle, kr, b2, q2, o2)["nkeys"] == 18).all()
    b3, q3, o3 = wl.make_read_batch(200, seed=93, lengths=(100,), flat_q=30, n_rate=0)
    assert (_check(oracle, kr, b3, q3, o3)["nkeys"] == 15).all()


def test_seed_edge_cases(oracle, kr):
    # no qualities (FASTA path), short reads, long reads (not staged in shared memory), mostly-N reads
    bases, qual, off = wl.make_read_batch(3000, seed=94, lengths=(12, 13, 14, 30, 64, 300, 600))
    _check(oracle, kr, bases, None, off)
    _check(oracle, kr, bases, qual, off)
    b = bases.copy(); q = qual.copy()
    for r in range(0, 3000, 7):
        b[off[r]: off[r + 1]] = ord("N"); q[off[r]: off[r + 1]] = 0
    exp = _check(oracle, kr, b, q, off)
    assert (exp["nkeys"][::7] <= 0).all()
    # low-quality everything
    q[:] = 2
    _check(oracle, kr, b, q, off)
