"""CPU: the C restatement of SamLine's record fields (oracle/sam_oracle.c) — invariants the SAM format and the reference's own
assertions impose (SamLine.java:744-746: the CIGAR consumes exactly the read; POS >= 1 for mapped reads; flag bits) and hand-checked
vectors."""
import numpy as np

from bbmap_b200 import sam
from sam_cases import cigar_stats, make_cases


def _cig(cbuf, coff, outs, i):
    return bytes(cbuf[coff[i]:coff[i] + outs["cigar_len"][i]].view(np.uint8)).decode()


def test_known_vectors(oracle):
    scaf = (np.array([0, 1], np.int32), np.array([8000], np.int32), np.array([5386], np.int32))
    cases = [  # match, start, stop, flags, expected cigar 1.4, cigar 1.3, pos
        (b"m" * 100, 8100, 8199, sam.RF_MAPPED | sam.RF_PERFECT, "100=", "100M", 101),
        (b"m" * 40 + b"S" + b"m" * 59, 8100, 8199, sam.RF_MAPPED, "40=1X59=", "100M", 101),
        (b"m" * 30 + b"DDD" + b"m" * 70, 8000, 8102, sam.RF_MAPPED | sam.RF_MINUS, "30=3D70=", "30M3D70M", 1),
        (b"m" * 30 + b"II" + b"m" * 68, 9000, 9097, sam.RF_MAPPED, "30=2I68=", "30M2I68M", 1001),
        (b"CC" + b"m" * 98, 7998, 8097, sam.RF_MAPPED, "2S98=", "2S98M", 1),                 # clipped bases before the scaffold start
        (b"m" * 95 + b"N" * 5, 8100, 8199, sam.RF_MAPPED, "95=5M", "100M", 101),
    ]
    for ver in (1.4, 1.3):
        tasks = np.zeros(len(cases), sam.SAM_TASK_DTYPE); bufs = []; off = 0
        for i, (m, a, b, fl, c14, c13, pos) in enumerate(cases):
            tasks[i] = (off, len(m), 1, a, b, 100, 9000, -1, fl, 0); bufs.append(np.frombuffer(m, np.uint8)); off += len(m)
        outs, cbuf, coff = oracle.sam_batch(tasks, np.concatenate(bufs), scaf, sam.default_cfg(ver))
        for i, (m, a, b, fl, c14, c13, pos) in enumerate(cases):
            assert _cig(cbuf, coff, outs, i) == (c14 if ver > 1.3 else c13), (i, ver, _cig(cbuf, coff, outs, i))
            assert outs["pos"][i] == pos and outs["scaffold"][i] == 0 and outs["rnext"][i] == -1 and outs["tlen"][i] == 0
            assert outs["flag"][i] == (0x10 if fl & sam.RF_MINUS else 0)
    # toMapq: score 9000, length 100 -> round((9000-4000)*1.6 * (1.5*log2(100)+36) / 10000)
    assert outs["mapq"][1] == int(np.floor(np.float32(5000 * 1.6) * (np.float32(1.5) * np.float32(np.log2(100)) + 36) / np.float32(10000) + 0.5))


def test_invariants(oracle):
    tasks, mbuf, scaf = make_cases()
    for ver in (1.4, 1.3):
        outs, cbuf, coff = oracle.sam_batch(tasks, mbuf, scaf, sam.default_cfg(ver))
        assert (outs["cigar_len"] != -2).all()
        ncig = 0
        for i in range(len(tasks)):
            t, o = tasks[i], outs[i]
            unmapped = bool(o["flag"] & 0x4)
            if not unmapped:
                assert o["pos"] >= 1 and o["mapq"] >= 1 and o["scaffold"] >= 0
            else:
                assert o["mapq"] == 0 and o["cigar_len"] == -1
            assert bool(o["flag"] & 0x10) == bool(t["flags"] & sam.RF_MINUS)
            assert bool(o["flag"] & 0x1) == (t["mate"] >= 0)
            if t["mate"] >= 0:
                assert bool(o["flag"] & 0x80) == bool(t["flags"] & sam.RF_PAIRNUM1) and bool(o["flag"] & 0x40) != bool(o["flag"] & 0x80)
                m = outs[t["mate"]]
                assert bool(o["flag"] & 0x8) == bool(m["flag"] & 0x4) and bool(o["flag"] & 0x20) == bool(m["flag"] & 0x10)
                assert int(o["tlen"]) * int(m["tlen"]) <= 0          # opposite signs (magnitudes can differ at scaffold ends: SamLine.java:205 clamps pos1 only)
                if o["flag"] & 0x2:
                    assert o["rnext"] == -2 and not unmapped
            if o["cigar_len"] > 0:
                ncig += 1
                q, r = cigar_stats(_cig(cbuf, coff, outs, i))
                assert q == t["read_len"], (i, _cig(cbuf, coff, outs, i))       # SamLine.java:744: cigarlen == bases.length
        assert ncig > 3000


def test_tasks_from_lists_by_hand(oracle):
    """Read.setFromTopSite / clearSite (stream/Read.java:1171-1190, 1213-1224, 1278-1286) as bbmap_b200.sam.tasks_from_lists states them, and the
    records SamLine makes of them: a plus-strand read, a minus-strand perfect read, an ambiguous read, an empty list, a cleared mapping."""
    from bbmap_b200 import sitelist as sl
    lists = np.zeros((5, 3), sl.SS_DTYPE); nss = np.array([2, 1, 1, 0, 1], np.int32); ro = np.arange(6, dtype=np.int64) * 100
    fl = np.zeros(5, sl.READ_OUT_DTYPE); fl["flags"] = [1, 1 | 2, 1 | 4, 0, 0]
    for r, (st, strand, perfect, slow) in enumerate([(8100, 0, 0, 9000), (8200, 1, 1, 9970), (8300, 0, 0, 7000), (0, 0, 0, 0), (8400, 1, 0, 6000)]):
        lists[r, 0]["chrom"] = 1; lists[r, 0]["start"] = st; lists[r, 0]["stop"] = st + 99; lists[r, 0]["strand"] = strand
        lists[r, 0]["perfect"] = perfect; lists[r, 0]["slow_score"] = slow; lists[r, 0]["score"] = slow + 5       # mapScore is the slow score
    lists[0, 1] = lists[0, 0]; lists[0, 1]["start"] = 20000                                                        # only the top site counts
    t = sam.tasks_from_lists(lists, nss, ro, fl)
    assert list(t["flags"]) == [sam.RF_MAPPED, sam.RF_MAPPED | sam.RF_MINUS | sam.RF_PERFECT, sam.RF_MAPPED | sam.RF_AMBIGUOUS, 0, 0]
    assert list(t["chrom"]) == [1, 1, 1, -1, -1] and list(t["start"]) == [8100, 8200, 8300, -1, -1] and list(t["stop"]) == [8199, 8299, 8399, -1, -1]
    assert list(t["score"]) == [9000, 9970, 7000, 0, 0] and (t["mate"] == -1).all() and (t["read_len"] == 100).all() and (t["match_len"] == 0).all()
    mo = np.arange(6, dtype=np.int64) * 104
    t2 = sam.tasks_from_lists(lists, nss, ro, fl, mo)
    assert list(t2["match_len"]) == [104, 104, 104, 0, 0] and list(t2["match_off"]) == [0, 104, 208, 312, 416]
    scaf = sam.scaffold_table([(1, 8000, 5000)], 1)
    out, _, _ = oracle.sam_batch(t, np.zeros(1, np.int8), scaf, sam.default_cfg())
    assert list(out["flag"]) == [0, 16, 0, 4, 4] and list(out["pos"][:3]) == [101, 201, 301] and (out["cigar_len"] == -1).all()
    assert out["mapq"][1] > out["mapq"][0] > out["mapq"][2] and list(out["scaffold"]) == [0, 0, 0, -1, -1]
