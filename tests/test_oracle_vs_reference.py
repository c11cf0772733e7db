"""CPU: the oracle's port fills are bit-identical to the reference's own C (oracle/_ref) on randomised
tasks — result vector, iteration counter, the whole `packed` matrix, and everything derived from it."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl
from oracle import oracle as orc

MAXR, MAXC = 601, 3000


def _need_ref(oracle):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (reference mount absent)")


@pytest.mark.parametrize("bw,ratio", [(0, 0.0), (12, 0.0), (40, 0.0), (0, 0.18)])
@pytest.mark.parametrize("mode", ["java", "raw_limited", "raw_unlimited"])
def test_batch_port_equals_reference(oracle, bw, ratio, mode):
    _need_ref(oracle)
    genome = wl.random_genome(20000, seed=11)
    flags = wl.TF_SCORE | wl.TF_TRACEBACK | {"java": 0, "raw_limited": wl.TF_RAW_LIMITED, "raw_unlimited": wl.TF_RAW_UNLIMITED}[mode]
    n = 120 if mode == "raw_unlimited" else 400
    reads, tasks = wl.make_msa_tasks(genome, n, seed=5 + bw, flags=flags, ratio=0.56 if bw != 12 else 0.336)
    if mode == "raw_limited":
        tasks["min_score"] -= 120
    moff = wl.match_offsets(tasks)
    o1, m1, c1 = oracle.run_batch(reads, genome, tasks, match_off=moff, bandwidth=bw, ratio=ratio, kind="port", threads=2)
    o2, m2, c2 = oracle.run_batch(reads, genome, tasks, match_off=moff, bandwidth=bw, ratio=ratio, kind="reference", threads=3)
    assert c1 == c2 and c1 > 0
    assert o1.tobytes() == o2.tobytes()
    assert m1.tobytes() == m2.tobytes()
    # the workload must exercise success, failure and (in java mode) both fill paths
    if mode != "raw_unlimited":
        assert (o1["result"][:, 4] == 1).any() and (o1["result"][:, 4] == 0).any()
        assert (o1["match_len"] > 0).any()


def test_packed_matrix_bit_exact(oracle):
    _need_ref(oracle)
    genome = wl.random_genome(20000, seed=12)
    reads, tasks = wl.make_msa_tasks(genome, 40, seed=9, flags=wl.TF_RAW_LIMITED)
    pa = oracle.new_packed(MAXR, MAXC); pb = pa.copy()
    for i, t in enumerate(tasks):
        r = reads[t["read_off"]: t["read_off"] + t["read_len"]].view(np.int8)
        g = genome.view(np.int8)
        bw = (0, 12, 40)[i % 3]
        ra, ia = oracle.fill_limited(r, g, t["ref_start"], t["ref_end"], t["min_score"] - 120, pa, MAXR, MAXC, bandwidth=bw, kind="port")
        rb, ib = oracle.fill_limited(r, g, t["ref_start"], t["ref_end"], t["min_score"] - 120, pb, MAXR, MAXC, bandwidth=bw, kind="reference")
        assert ra.tolist() == rb.tolist() and ia == ib
        assert np.array_equal(pa, pb)
        if i % 5 == 0:
            ra, ia = oracle.fill_unlimited(r, g, t["ref_start"], t["ref_end"], pa, MAXR, MAXC, kind="port")
            rb, ib = oracle.fill_unlimited(r, g, t["ref_start"], t["ref_end"], pb, MAXR, MAXC, kind="reference")
            assert ra.tolist() == rb.tolist() and ia == ib
            assert np.array_equal(pa, pb)
