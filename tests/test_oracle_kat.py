"""CPU: the oracle (port AND the reference's own C) reproduces the Appendix-B known answers."""
import numpy as np
import pytest

from kat import KATS, REF, MINSCORE, B, rle
from oracle import oracle as orc

MAXR, MAXC = 601, 3000


@pytest.mark.parametrize("kind", ["port", "reference"])
@pytest.mark.parametrize("kat", KATS, ids=[k[0] for k in KATS])
def test_kat_fill(oracle, kat, kind):
    if kind == "reference" and not oracle.has_reference:
        pytest.skip("oracle/_ref not built (reference mount absent)")
    name, read, a, b, fn, bw, result, iters, score2, match = kat
    packed = oracle.new_packed(MAXR, MAXC)
    if fn == "limited":
        res, it = oracle.fill_limited(B(read), B(REF), a, b, MINSCORE, packed, MAXR, MAXC, bandwidth=bw, kind=kind)
    else:
        res, it = oracle.fill_unlimited(B(read), B(REF), a, b, packed, MAXR, MAXC, kind=kind)
    assert res.tolist() == result
    if iters is not None:
        assert it == iters


@pytest.mark.parametrize("kind", ["port", "reference"])
def test_kat_score_traceback(oracle, kind):
    if kind == "reference" and not oracle.has_reference:
        pytest.skip("oracle/_ref not built")
    ref = B(REF)
    for name, read, a, b, fn, bw, result, iters, score2, match in KATS:
        r = B(read)
        tasks = np.zeros(1, orc.TASK_DTYPE)
        tasks[0] = (0, 0, len(r), len(ref), a, b, MINSCORE, (orc.TF_RAW_LIMITED if fn == "limited" else orc.TF_RAW_UNLIMITED) | orc.TF_SCORE | orc.TF_TRACEBACK)
        outs, mbuf, cells = oracle.run_batch(r, ref, tasks, bandwidth=bw, kind=kind)
        o = outs[0]
        assert o["result"][: len(result)].tolist() == result, name
        if iters is not None:
            assert o["iterations"] == iters == cells
        if score2 is None:
            assert o["score_len"] == 0 and o["match_len"] == -1
        else:
            assert o["score_len"] == len(score2)
            assert o["score"][: len(score2)].tolist() == score2, name
            assert rle(mbuf[: o["match_len"]]) == match, name
