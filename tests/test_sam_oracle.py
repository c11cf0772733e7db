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
