"""CPU: bbm_break_reads (host C++) against a plain-Python statement of ReformatReads.breakReads (current/jgi/ReformatReads.java:1179-1219): which reads
are dropped, where the pieces are cut, how they are named, and the paired-input error."""
import numpy as np
import pytest

from bbmap_b200 import lib as _lib
from bbmap_b200.reads import break_reads


def _py_break(reads, quals, names, mx, mn):
    mn = max(0, mn)
    out = []
    for i, (b, q, nm) in enumerate(zip(reads, quals, names)):
        if len(b) < mn:
            continue
        if mx < 1 or len(b) <= mx:
            out.append((b, q, nm, i, 0)); continue
        limit = len(b) - mn
        num, start = 1, 0
        while start < limit:
            stop = min(start + mx, len(b))
            out.append((b[start:stop], None if q is None else q[start:stop], nm + b"_%d" % num, i, start))
            num += 1; start += mx
    return out


def _pack(chunks):
    off = np.zeros(len(chunks) + 1, np.int64)
    off[1:] = np.cumsum([len(c) for c in chunks])
    return np.frombuffer(b"".join(chunks), np.int8) if off[-1] else np.zeros(0, np.int8), off


@pytest.mark.parametrize("mx,mn", [(500, 0), (500, 50), (100, 100), (0, 40), (37, 13), (600, 1)])
@pytest.mark.parametrize("with_q", [True, False])
def test_break_reads_matches_reference_rule(mx, mn, with_q):
    rng = np.random.default_rng(mx * 7 + mn)
    lens = [int(x) for x in rng.choice([0, 1, 12, 13, 36, 37, 38, 99, 100, 101, 150, 499, 500, 501, 550, 999, 1000, 1001, 1040, 2500], size=200)]
    reads = [bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n).tobytes()) for n in lens]
    quals = [bytes(rng.integers(0, 42, size=n, dtype=np.uint8).tobytes()) if with_q else None for n in lens]
    names = [b"read%d/x y" % i for i in range(len(lens))]
    b, ro = _pack(reads); nm, no = _pack(names)
    q = _pack(quals)[0] if with_q else None
    got = break_reads(b, q, ro, nm, no, mx, mn)
    exp = _py_break(reads, quals, names, mx, mn)
    assert len(got["src"]) == len(exp)
    for i, (eb, eq, en, src, st) in enumerate(exp):
        a, z = int(got["read_off"][i]), int(got["read_off"][i + 1])
        assert got["bases"][a:z].tobytes() == eb, i
        if with_q:
            assert got["quality"][a:z].tobytes() == eq, i
        assert got["names"][int(got["name_off"][i]): int(got["name_off"][i + 1])].tobytes() == en, i
        assert (int(got["src"][i]), int(got["piece_start"][i])) == (src, st)
    if mx > 0:
        assert (np.diff(got["read_off"]) <= mx).all()
    assert (np.diff(got["read_off"]) >= min(mn, 1) * 0).all()


def test_break_reads_errors():
    b, ro = _pack([b"A" * 700, b"C" * 100]); nm, no = _pack([b"a", b"b"])
    with pytest.raises(_lib.BbmError):
        break_reads(b, None, ro, nm, no, 500, 0, paired=True)          # the reference asserts: paired input cannot be broken
    with pytest.raises(_lib.BbmError):
        break_reads(b, None, ro, nm, no, 0, 0)
    with pytest.raises(_lib.BbmError):
        break_reads(b, None, ro, nm, no, 50, 60)
    ok = break_reads(b[:0], None, np.zeros(1, np.int64), nm[:0], np.zeros(1, np.int64), 500, 0)
    assert len(ok["src"]) == 0
    # paired input whose reads all fit passes through
    b2, ro2 = _pack([b"A" * 300, b"C" * 100])
    same = break_reads(b2, None, ro2, nm, no, 500, 0, paired=True)
    assert same["bases"].tobytes() == b2.tobytes() and same["names"].tobytes() == nm.tobytes()
