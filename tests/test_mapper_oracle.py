"""CPU: the sequential restatement of the mapper tail (oracle/mapper_oracle.c, oracle/chain.py) against what can be pinned without a JVM:
hand-worked values of MSA.score(match), the reference's shipped phiX pairs (configs[0]; truth in the read name) through the whole paired
chain, structural invariants the reference asserts (CIGAR consumes the read, match length == mapped length, proper-pair FLAG / TLEN symmetry)."""
import ctypes as C
import re

import numpy as np
import pytest

from bbmap_b200 import workloads as wl
from bbmap_b200.index import pack_chromosomes


def _score(oracle, s):
    b = np.frombuffer(s.encode(), np.int8).copy()
    return oracle.lib.orc_score_match(b.ctypes.data_as(C.c_void_p), C.c_int(len(b)))


def test_score_match_hand_values(oracle):
    """MSA.score(byte[] match) (current/align2/MSA.java:488-560) with the 11ts constants (MultiStateAligner11tsJNI.java:1489-1563)."""
    assert _score(oracle, "m" * 100) == 70 + 99 * 100                               # calcMatchScore: first match 70, then 100 each
    # one substitution in the middle: 9970 - 100 - 127 - 30
    assert _score(oracle, "m" * 50 + "S" + "m" * 49) == 9713
    # KAT3: 3-base deletion: 9970 - 472 - 2*33 - 30
    assert _score(oracle, "m" * 49 + "DDD" + "m" * 51) == 9402
    # KAT4: 2-base insertion (two read bases not matched): 98 matches in two runs + INS + INS2
    assert _score(oracle, "m" * 50 + "II" + "m" * 48) == (70 + 49 * 100) + (-395 - 39) + (70 + 47 * 100)
    # a substitution right after a single match is POINTS_SUBR (-147), after an N it is POINTS_SUB2 (-51)
    assert _score(oracle, "mS" + "m" * 10) == 70 - 147 + 70 + 9 * 100
    assert _score(oracle, "mmNS" + "m" * 10) == (70 + 100) + 0 - 51 + (70 + 9 * 100)
    # clipped symbols score nothing; X/Y score like insertions
    assert _score(oracle, "CCC" + "m" * 10) == 70 + 9 * 100
    assert _score(oracle, "XX" + "m" * 10) == (-395 - 39) + 70 + 9 * 100
    # long deletion: 472 + 4*33 + 15*9 + 60*1 + ceil((100-80)/4)*1 for 100 bases (calcDelScore, approximateGaps needs len > 256)
    assert _score(oracle, "m" * 20 + "D" * 100 + "m" * 20) == 2 * (70 + 19 * 100) - (472 + 4 * 33 + 15 * 9 + 60 * 1 + ((100 - 80 + 3) // 4) * 1)


def _phix_pairs():
    from test_search_oracle import phix
    d, cb, co, table = phix()
    n = len(d["r1_off"]) - 1
    L = 100
    assert (np.diff(d["r1_off"]) == L).all() and (np.diff(d["r2_off"]) == L).all()
    bases = np.empty(2 * n * L, np.uint8); qual = np.empty(2 * n * L, np.uint8)
    bases.reshape(n, 2, L)[:, 0] = d["r1_bases"].reshape(n, L); bases.reshape(n, 2, L)[:, 1] = d["r2_bases"].reshape(n, L)
    qual.reshape(n, 2, L)[:, 0] = d["r1_qual"].reshape(n, L); qual.reshape(n, 2, L)[:, 1] = d["r2_qual"].reshape(n, L)
    truth = np.empty((2 * n, 5), np.int32); truth[0::2] = d["r1_truth"]; truth[1::2] = d["r2_truth"]
    return cb, co, table, bases, qual, np.arange(2 * n + 1, dtype=np.int64) * L, truth


def _cigar_consumes(cig, L):
    q = sum(int(n) for n, op in re.findall(r"(\d+)([=XMIS])", cig))
    return q == L


def check_mapping_invariants(res, off, truth=None, thresh=20, min_correct=0.9):
    recs, sam, ms = res["recs"], res["sam"], res["match_stride"]
    n = len(recs)
    mapped = (recs["flags"] & 1) != 0
    for r in np.nonzero(mapped)[0]:
        m = res["match"][r * ms:r * ms + int(recs["match_len"][r])].tobytes()
        L = int(off[r + 1] - off[r])
        assert sum(1 for c in m if c not in b"D") == L, (r, m)                                         # match string consumes the read
        assert sum(1 for c in m if c not in b"I") == int(recs["stop"][r] - recs["start"][r] + 1), (r, m)  # SiteScore.lengthsAgree
        cig = res["cigar"][int(res["cigar_off"][r]):int(res["cigar_off"][r]) + int(sam["cigar_len"][r])].tobytes().decode()
        assert _cigar_consumes(cig, L), (r, cig)                                                       # SamLine.java:744
        assert 0 <= sam["mapq"][r] <= 50 and sam["pos"][r] >= 1
    if truth is not None:
        ok = mapped & (recs["chrom"] == truth[:, 0]) & (recs["strand"] == truth[:, 1]) & \
            ((np.abs(recs["start"] - truth[:, 2]) <= thresh) | (np.abs(recs["stop"] - truth[:, 3]) <= thresh))
        assert ok.mean() >= min_correct, ok.mean()
    return mapped


def test_phix_pairs_whole_chain(oracle):
    """configs[0]: the reference's 100 shipped phiX pairs through processReadPair (restated): mapped at the origin the read names give, mated, proper-pair flags."""
    from oracle import chain
    cb, co, table, bases, qual, off, truth = _phix_pairs()
    idx = oracle.index_build(cb, co, 13, -1)
    res = chain.map_pairs(oracle, idx, cb, co, table, bases, qual, off)
    mapped = check_mapping_invariants(res, off, truth)
    assert res["site_overflow"] == 0 and (res["recs"]["status"] == 0).all()
    f = res["recs"]["flags"]; sam = res["sam"]
    assert mapped.mean() >= 0.95 and ((f & 8) != 0).mean() >= 0.9          # 8 of the 200 shipped reads carry too many errors / too low quality to map
    proper = (sam["flag"] & 2) != 0
    assert proper.mean() >= 0.9
    a, b = sam[0::2], sam[1::2]
    both = proper[0::2] & proper[1::2]
    assert (a["tlen"][both] == -b["tlen"][both]).all() and (a["pnext"][both] == b["pos"][both]).all() and (b["pnext"][both] == a["pos"][both]).all()
    assert ((a["flag"][both] & 0x40) != 0).all() and ((b["flag"][both] & 0x80) != 0).all()
    assert (((a["flag"][both] >> 4) & 1) != ((b["flag"][both] >> 4) & 1)).all()                          # opposite strands
    # SAM text: one line per read, 11 mandatory fields + NM/AM tags for mapped reads
    lines = chain.sam_lines(res, off, names=[b"r%d/%d" % (i // 2, i % 2 + 1) for i in range(len(f))], scaf_names=[b"phiX"], paired=True)
    for r, ln in enumerate(lines):
        fld = ln.rstrip(b"\n").split(b"\t")
        assert len(fld) >= 11 and fld[0] == b"r%d" % (r // 2) and int(fld[1]) == sam["flag"][r] and int(fld[3]) == sam["pos"][r]
        assert len(fld[9]) == 100 and len(fld[10]) == 100
        if mapped[r]:
            assert fld[2] == b"phiX" and any(x.startswith(b"NM:i:") for x in fld[11:]) and any(x.startswith(b"AM:i:") for x in fld[11:])


def test_phix_single_whole_chain(oracle):
    from oracle import chain
    cb, co, table, bases, qual, off, truth = _phix_pairs()
    idx = oracle.index_build(cb, co, 13, -1)
    res = chain.map_single(oracle, idx, cb, co, table, bases, qual, off)
    mapped = check_mapping_invariants(res, off, truth)
    assert mapped.mean() >= 0.9 and (res["recs"]["status"] == 0).all()      # without the mate, 15 of the 200 shipped reads stay unmapped


def test_pairs_with_damaged_and_missing_mates(oracle):
    """Rescue places mates the index search cannot (heavy damage); an unmappable mate leaves its partner mapped but unpaired (FLAG 0x8)."""
    from oracle import chain
    g = wl.random_genome(200_000, seed=41)
    cb, co, table = pack_chromosomes([g])
    R = wl.make_mapping_reads(cb, co, table, 400, seed=42, sub_rate=0.01, indel_rate=0.0)
    L = 150; n = 800
    rng = np.random.Generator(np.random.PCG64(43))
    bases = R["bases"].copy()
    damaged = [2 * i + 1 for i in range(0, 60)]
    for r in damaged:                                   # every 8th base substituted: no 13-mer seed survives, 19 mismatches <= MAX_RESCUE_MISMATCHES
        v = bases[r * L:(r + 1) * L]
        v[4::8] = wl.ACGT[(np.searchsorted(wl.ACGT, v[4::8]) + 1) % 4]
    junk = [2 * i + 1 for i in range(100, 120)]
    for r in junk:
        bases[r * L:(r + 1) * L] = wl.ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]
    idx = oracle.index_build(cb, co, 13, -1)
    res = chain.map_pairs(oracle, idx, cb, co, table, bases, R["qual"], R["off"])
    check_mapping_invariants(res, R["off"])
    f = res["recs"]["flags"]; sam = res["sam"]; tr = R["truth"]
    assert res["rescue_scans"] > 0 and ((f & 16) != 0).sum() >= 10                     # some mates only rescue could place
    resc = np.nonzero((f & 16) != 0)[0]
    assert (np.abs(res["recs"]["start"][resc] - tr[resc, 2]) <= 20).mean() > 0.9        # ... and at their true origin
    for r in junk:
        assert not (f[r] & 1) and (f[r - 1] & 1) and not (f[r - 1] & 8)
        assert sam["flag"][r - 1] & 0x8 and sam["flag"][r] & 0x4 and sam["pos"][r] == sam["pos"][r - 1]   # unmapped mate is placed at its partner (SamLine.java:236-246)
    assert res["mated"] == int(((f[0::2] & 8) != 0).sum())
