"""CPU: SAM record fields — the C restatement the CUDA kernel is checked against (oracle/sam_oracle.c) must equal a second restatement
written from the Java text alone (tests/pysam_fields.py: Read objects mutated by the SamLine constructor as the reference does, records built in
file order) on FLAG, POS, MAPQ, RNAME, RNEXT, PNEXT, TLEN and the CIGAR text, SAM 1.4 and 1.3, soft-clipping on and off, with an intron limit."""
import numpy as np
import pytest

from bbmap_b200 import sam
from sam_cases import make_cases

import pysam_fields as ps


def _table(scaf):
    off, loc, ln = scaf
    t = []
    for ch in range(1, len(off)):
        for g in range(int(off[ch - 1]), int(off[ch])):
            t.append((ch, int(loc[g]), int(ln[g])))
    return t


@pytest.mark.parametrize("version,soft_clip,intron_limit", [(1.4, 1, 2 ** 31 - 1), (1.3, 1, 2 ** 31 - 1), (1.4, 0, 2 ** 31 - 1), (1.4, 1, 1), (1.3, 0, 1)])
def test_sam_fields_equal_independent_restatement(oracle, version, soft_clip, intron_limit):
    tasks, mbuf, scaf = make_cases(n=3000, seed=505)
    off, loc, ln = scaf
    for i in range(5, len(tasks), 37):                       # alignments that span two scaffolds: the constructor unmaps them (and un-pairs the mate)
        ch = 1 if i % 2 else 3
        g = int(off[ch - 1])
        tasks["chrom"][i] = ch; tasks["start"][i] = int(loc[g]) + int(ln[g]) - 60; tasks["stop"][i] = int(loc[g]) + int(ln[g]) + 330
    cfg = sam.default_cfg(version)
    cfg["soft_clip"] = soft_clip; cfg["intron_limit"] = intron_limit
    outs, cbuf, coff = oracle.sam_batch(tasks, mbuf, scaf, cfg)
    S = ps.Scaffolds(_table(scaf), int(cfg["inter_scaffold_padding"][0]))
    reads = []
    for t in tasks:
        f = int(t["flags"])
        m = None if t["match_len"] == 0 else bytes(mbuf[int(t["match_off"]): int(t["match_off"]) + int(t["match_len"])])
        reads.append(ps.PyRead(int(t["chrom"]), int(t["start"]), int(t["stop"]), int(t["read_len"]), int(t["score"]), m, bool(f & sam.RF_MAPPED),
                               bool(f & sam.RF_MINUS), bool(f & sam.RF_PERFECT), bool(f & sam.RF_AMBIGUOUS), bool(f & sam.RF_SECONDARY),
                               bool(f & sam.RF_DISCARDED), bool(f & sam.RF_PAIRED), 1 if f & sam.RF_PAIRNUM1 else 0))
    for i, t in enumerate(tasks):
        if t["mate"] >= 0:
            reads[i].mate = reads[int(t["mate"])]
    seen = dict(cigar=0, tlen=0, multi=0, clipped=0)
    for i, t in enumerate(tasks):
        was_mapped = reads[i].mapped
        sl = ps.PySamLine(reads[i], reads[i].pairnum, S, v14=version > 1.3, soft_clip=bool(soft_clip), intron_limit=intron_limit,
                          penalize_ambig=bool(cfg["penalize_ambig"][0]))
        o = outs[i]
        enc = lambda x: -1 if x in (None, "*") else (-2 if x == "=" else x)
        got = (sl.flag, sl.pos, sl.mapq, enc(sl.rname), enc(sl.rnext), sl.pnext, sl.tlen)
        exp = (int(o["flag"]), int(o["pos"]), int(o["mapq"]), int(o["scaffold"]), int(o["rnext"]), int(o["pnext"]), int(o["tlen"]))
        assert got == exp, (i, got, exp, t)
        cig = None if o["cigar_len"] < 0 else bytes(cbuf[coff[i]: coff[i] + o["cigar_len"]].view(np.uint8)).decode()
        assert sl.cigar == cig, (i, sl.cigar, cig)
        seen["cigar"] += cig is not None; seen["tlen"] += sl.tlen != 0; seen["multi"] += bool(t["flags"] & sam.RF_MAPPED) and not reads[i].mapped
        seen["clipped"] += cig is not None and "S" in cig
    assert seen["cigar"] > 2000 and seen["tlen"] > 100 and seen["multi"] > 5 and (seen["clipped"] > 50 or not soft_clip), seen
