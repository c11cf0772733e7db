"""GPU parity: BandedAligner CUDA kernel vs the oracle (port, and the reference's own C when built) — edits and all five
return values, all four directions, exact/inexact, swap rules, even/odd/wide bands, N bases."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["thread", "warp"])
def band(request):
    """Both kernels: one thread per pair for bands of up to 15 cells with the wider pairs handed to the warp-per-pair kernel (default), and the
    warp-per-pair kernel for every pair."""
    from bbmap_b200.banded import BandedAlignerCUDA
    b = BandedAlignerCUDA()
    b.L.bbm_set_option(b.h, b"banded_thread", 1 if request.param == "thread" else 0)
    yield b
    b.close()


@pytest.mark.parametrize("widths,maxlen", [(None, 1200), ((3, 5, 11, 21, 53, 64, 101, 127), 500), ((1, 2, 4, 31, 32, 33, 63), 300), ((1, 2, 3, 4, 5, 6, 7, 8, 9), 400),
                                           ((8, 9, 10, 11, 12, 13, 14, 15, 16, 17), 400)])
def test_banded_random(oracle, band, widths, maxlen):
    q, r, tasks = wl.make_banded_tasks(3000, seed=18, min_len=10, max_len=maxlen, widths=widths,
                                       max_edits=(2, 5, 26) if widths is None else (0, 1, 2, 5, 16, 26, 40, 63))
    kind = "reference" if oracle.has_reference else "port"
    exp = oracle.banded_batch(q, r, tasks, kind=kind, threads=8)
    got = band.align_batch(q, r, tasks)
    bad = np.nonzero([got[i].tobytes() != exp[i].tobytes() for i in range(len(tasks))])[0]
    assert len(bad) == 0, "%d differ; first %d: task=%s got=%s exp=%s" % (len(bad), bad[0], tasks[bad[0]], got[bad[0]], exp[bad[0]])
    assert len(np.unique(exp["edits"])) > 5 and len(np.unique(exp["rv"][:, 4])) > 2


def test_banded_long_and_empty(oracle, band):
    q, r, tasks = wl.make_banded_tasks(200, seed=19, min_len=2000, max_len=5000)
    exp = oracle.banded_batch(q, r, tasks, kind="port", threads=8)
    got = band.align_batch(q, r, tasks)
    assert got.tobytes() == exp.tobytes()
    # len<1 (start beyond the end) returns 0 and leaves lastRow=-1 (jni/BandedAlignerJNI.c:153-169)
    t = tasks[:4].copy()
    t["dir"] = [0, 1, 2, 3]
    t["qstart"] = t["query_len"]; t["rstart"] = t["ref_len"]
    t["qstart"][1] = -1; t["qstart"][2] = -1; t["rstart"][2] = -1; t["rstart"][3] = -1
    exp = oracle.banded_batch(q, r, t, kind="port")
    got = band.align_batch(q, r, t)
    assert got.tobytes() == exp.tobytes()
    assert len(band.align_batch(q, r, tasks[:0])) == 0
