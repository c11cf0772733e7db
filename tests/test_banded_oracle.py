"""CPU: BandedAligner oracle — port vs the reference's own C, plus the Appendix-B known answers."""
import numpy as np
import pytest

from bbmap_b200 import workloads as wl
from kat import lcg_seq, B

COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def rc(s):
    return "".join(COMP[c] for c in reversed(s))


def test_banded_kats(oracle):
    ref = lcg_seq(120, 777); q = ref[:50] + ref[51:100]
    s = list(ref[:100])
    for p in (30, 60):
        s[p] = "A" if s[p] != "A" else "C"
    s = "".join(s)
    cases = [
        (0, ref[:100], ref[:100], 0, 0, 5, 11, (0, [99, 99, 99, 0, 0])),
        (0, q, ref[:100], 0, 0, 5, 11, (2, [98, 98, 98, 2, 0])),
        (2, q, ref[:100], 98, 99, 5, 11, (2, [0, 0, 98, 2, -1])),
        (1, rc(q), ref[:100], 98, 0, 5, 11, (2, [0, 98, 98, 2, 0])),
        (3, rc(q), ref[:100], 0, 99, 5, 11, (2, [98, 0, 98, 2, -1])),
        (0, s, ref[:100], 0, 0, 5, 11, (2, [99, 99, 99, 2, 0])),
        (0, s, ref[:100], 0, 0, 1, 11, (2, [59, 59, 60, 2, 0])),
    ]
    for d, qq, rr, qs, rs, me, mw, exp in cases:
        assert oracle.banded(d, B(qq), B(rr), qs, rs, me, True, mw) == exp


@pytest.mark.parametrize("widths", [None, (3, 5, 11, 21, 53, 64)])
def test_banded_port_equals_reference(oracle, widths):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref not built")
    q, r, tasks = wl.make_banded_tasks(1500, seed=8, min_len=20, max_len=700, widths=widths)
    a = oracle.banded_batch(q, r, tasks, kind="port", threads=2)
    b = oracle.banded_batch(q, r, tasks, kind="reference", threads=3)
    assert a.tobytes() == b.tobytes()
    assert len(np.unique(a["edits"])) > 5
