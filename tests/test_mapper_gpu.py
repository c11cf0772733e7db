"""GPU: bbm_map_batch_host (the whole chain behind one C-ABI call) == the sequential CPU chain (oracle/chain.py) on the same reads: every field of the
read record, the primary match string, FLAG/POS/MAPQ/CIGAR.  Unpaired reads; parity of the Java-only stages is UNPINNED against Java (no JVM)."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl

pytestmark = pytest.mark.gpu

REC_FIELDS = ("chrom", "start", "stop", "strand", "map_score", "flags", "match_len", "cz3_sub", "tip_penalty", "status")


def _genome(kind, seed):
    g = wl.random_genome(300_000, seed=seed)
    if kind == "repeats":                       # planted repeat families: several sites per read, clearzone / ambiguity paths
        rng = np.random.Generator(np.random.PCG64(seed + 100))
        for _ in range(6):
            unit = wl.ACGT[rng.integers(0, 4, size=400, dtype=np.uint8)]
            for _c in range(5):
                u = unit.copy()
                m = rng.random(400) < 0.01
                u[m] = wl.ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
                q = int(rng.integers(0, len(g) - 400)); g[q:q + 400] = u
    return g


def _compare(dev, ref, n):
    ms = ref["match_stride"]
    for f in REC_FIELDS:
        a, b = dev["recs"][f], ref["recs"][f]
        bad = np.nonzero(a != b)[0]
        assert len(bad) == 0, "record field %s differs for reads %s: device %s, oracle %s" % (f, bad[:5], a[bad[:5]], b[bad[:5]])
    dm = dev["match"][: n * ms].reshape(n, ms); om = ref["match"][: n * ms].reshape(n, ms)
    live = np.arange(ms)[None, :] < ref["recs"]["match_len"][:, None]
    assert np.array_equal(dm[live], om[live]), "primary match strings differ"
    assert dev["sam"].tobytes() == ref["sam"].tobytes(), "SAM fields differ"


@pytest.mark.parametrize("kind,seed,L", [("plain", 11, 150), ("repeats", 12, 150), ("plain", 13, 100), ("repeats", 14, 250)])
def test_map_batch_single_equals_cpu_chain(kind, seed, L):
    from bbmap_b200.mapper import BBMapCUDA
    from oracle import chain, oracle as orc
    g = _genome(kind, seed)
    m = BBMapCUDA([g])
    try:
        R = wl.make_mapping_reads(m.cb, m.co, m.table, 1500, L=L, seed=seed + 1, sub_rate=0.015, indel_rate=0.02 / 3)
        o = orc.get()
        idx = o.index_build(m.cb, m.co, 13, -1)
        ref = chain.map_single(o, idx, m.cb, m.co, m.table, R["bases"], R["qual"], R["off"])
        dev = m.map_batch(R["bases"], R["qual"], R["off"], match_stride=ref["match_stride"])
        n = len(R["off"]) - 1
        assert ref["site_overflow"] == 0 and int(dev["stats"]["site_overflow_reads"]) == 0
        _compare(dev, ref, n)
        assert (ref["recs"]["flags"] & 1).mean() > 0.97
        assert int(dev["stats"]["realign_fills"]) == ref["realign_fills"] and int(dev["stats"]["slow_alignments"]) == ref["slow_alignments"]
    finally:
        m.close()
