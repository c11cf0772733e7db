"""CPU: BBMapThread.processRead END TO END from independent restatements only — Read.validate and the seeds (tests/pyseed.py), BBIndex.find on a binary QuadHeap
(tests/pyfind.py), removeOutOfBounds / trimList / scoreNoIndels(Read) / findTipDeletions(Read) (tests/pysitelist.py, pyreadpolicies.py), scoreSlow
(tests/pyscoreslow.py), the final policy, the genMatchString loop with realign_new, clearzone 3 and the tip penalty (tests/test_chain_tail_independent._finish), the
SAM record (tests/pysam_fields.py); every MultiStateAligner fill by the reference's own C, every walk by tests/pywalk.py — against the sequential C chain the CUDA
mapper is tested against (oracle/chain.map_single): locus, strand, mapScore, flags, match string, FLAG / POS / MAPQ / CIGAR.  Reads whose candidate sites carry gap
arrays, reads with more than 16 candidate sites (the C chain's list cap) and reads whose match string keeps X / Y / C symbols are left out."""
import functools

import numpy as np
import pytest

from bbmap_b200 import rescue as rs
from bbmap_b200 import sitelist as sl
from bbmap_b200 import workloads as wl
from bbmap_b200.index import pack_chromosomes
from bbmap_b200.keyring import default_cfg
from bbmap_b200.mapper import map_cfg
from bbmap_b200.sam import default_cfg as sam_default_cfg
from oracle import chain

import pyfind
import pyreadpolicies as prp
import pyrealign
import pysam_fields as psf
import pyscoreslow
import pyseed
import pysitelist as ps
from test_chain_tail_independent import _finish


def _reads(g, rng, n, L):
    out = []
    for i in range(n):
        p = int(rng.integers(0, len(g) - L - 40))
        r = g[p:p + L + 30].copy()
        u = rng.random()
        if u < 0.35:
            q = int(rng.integers(20, L - 20)); d = int(rng.integers(1, 9))
            r = np.concatenate([r[:q], r[q + d:]]) if rng.random() < 0.5 else np.concatenate([r[:q], wl.ACGT[rng.integers(0, 4, size=d, dtype=np.uint8)], r[q:]])
        r = r[:L].copy()
        m = rng.random(L) < (0.0 if u > 0.8 else 0.02)
        r[m] = wl.ACGT[rng.integers(0, 4, size=int(m.sum()), dtype=np.uint8)]
        if rng.random() < 0.05:
            r[int(rng.integers(0, L))] = ord("N")
        if rng.random() < 0.04:
            r = wl.ACGT[rng.integers(0, 4, size=L, dtype=np.uint8)]          # unmappable
        out.append(r if i % 2 == 0 else wl.revcomp(r))
    bases = np.concatenate(out); off = np.arange(n + 1, dtype=np.int64) * L
    qual = rng.integers(12, 41, size=len(bases)).astype(np.uint8)
    qual[bases == ord("N")] = 0
    return bases, qual, off


@pytest.mark.parametrize("seed,L", [(1001, 150), (1002, 100)])
def test_process_read_end_to_end(oracle, seed, L):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    rng = np.random.Generator(np.random.PCG64(seed))
    g = wl.ACGT[rng.integers(0, 4, size=90000, dtype=np.uint8)]
    unit = g[5000:5300].copy()
    for c in range(4):
        p = 12000 + 15000 * c; g[p:p + 300] = unit
        if c % 2:
            g[p + 150] = wl.ACGT[(int(np.searchsorted(wl.ACGT, g[p + 150])) + 1) & 3]
    cb, co, table = pack_chromosomes([g])
    bases, qual, off = _reads(g, rng, 150, L)
    for i in range(0, 150, 5):                                        # reads from the repeat family: several sites, ambiguity, clearzone 3
        p = 12000 + 15000 * int(rng.integers(0, 4)) + int(rng.integers(0, 140))
        r = g[p:p + L]
        bases[off[i]:off[i + 1]] = r if i % 2 == 0 else wl.revcomp(r)
    idx = oracle.index_build(cb, co, 13, -1)
    ref = chain.map_single(oracle, idx, cb, co, table, bases, qual, off)
    icfg, blocks, counts, hist = idx
    py = pyfind.BBIndexPy(icfg, blocks, counts, hist, cb, co, quit_after_two_perfects=True)
    scfg = default_cfg()[0]
    pcfg = sl.policy_cfg(); mcfg = map_cfg(); wcfg = sl.slow_cfg(); tcfg = rs.tipdel_cfg()
    samcfg = sam_default_cfg()
    cb8 = np.ascontiguousarray(cb).view(np.int8)
    ref8 = cb8[int(co[0]): int(co[1])]
    refs = {1: ref8.tolist()}
    R = pyrealign.Realigner(oracle, ref8)
    packed = oracle.new_packed(601, 3000)
    S = psf.Scaffolds([(c, s, ln) for c, s, ln in table], 300)
    single = lambda c, a, b: S.is_single(c, a, b)
    maxidx = {1: len(ref8) - 1}
    ms = ref["match_stride"]
    done = mapped = skipped = with_indel = 0
    for r in range(len(off) - 1):
        a, b = int(off[r]), int(off[r + 1])
        rb, rq, junk = pyseed.validate([int(x) for x in bases[a:b].view(np.int8)], [int(x) for x in qual[a:b].view(np.int8)])
        rm = pyseed.reverse_complement_bases(rb)
        seed_ = pyseed.quick_map_seed(rb, rq, int(scfg["keylen"]), int(scfg["maxDesiredKeys"]), int(scfg["baseKeyHitScore"]), int(scfg["minApproxHitsToKeep"]),
                                      float(scfg["keyDensity"]), float(scfg["maxKeyDensity"]), float(scfg["minKeyDensity"]))
        e = ref["recs"][r]; ef = int(e["flags"])
        if seed_ is None:
            assert not ef & 1, r
            done += 1
            continue
        found = py.find(bytes(rb), seed_["baseScores"], seed_["offsets"], seed_["keyScores"])
        if found["gapfix"] or len(found["sites"]) > 16 or any(s.gaps is not None for s in found["sites"]) or e["status"]:
            skipped += 1
            continue
        sites = [ps.Site(s.chrom, s.strand, s.start, s.stop, s.hits, s.score, s.score, 0, 0, bool(s.perfect), bool(s.semiperfect), False, None) for s in found["sites"]]
        bp8 = np.array(rb, np.int8); bm8 = np.array(rm, np.int8)
        ps.remove_out_of_bounds(sites, L, maxidx, single, True, 2522)
        ps.trim_policy(sites, L, pcfg[0])
        near = prp.score_no_indels_read(sites, rb, rm, refs)
        sites.sort(key=functools.cmp_to_key(ps.compare_to))
        if near < 1:
            prp.find_tip_deletions_read(sites, rb, rm, rq, refs, {1: 0}, int(tcfg["search_range"][0]), int(tcfg["slow_rescue_padding"][0]))
            pyscoreslow.score_slow(oracle, packed, sites, bp8, bm8, ref8, wcfg[0])
        fin = ps.final_policy(sites, L, pcfg[0])
        got = _finish(R, sites, fin, bp8, bm8, pcfg, mcfg[0])
        mlen = int(e["match_len"])
        m_exp = ref["match"][r * ms: r * ms + mlen].tobytes() if mlen > 0 else None
        if (m_exp is not None and any(c in m_exp for c in b"XYC")) or (got["match"] is not None and any(c in got["match"] for c in b"XYC")):
            skipped += 1
            continue
        assert got["mapped"] == bool(ef & 1), (r, got, e)
        done += 1
        if not got["mapped"]:
            continue
        assert (got["chrom"], got["start"], got["stop"], got["strand"], got["map_score"], got["perfect"], got["ambiguous"], got["cz3"], got["pen"]) == \
               (int(e["chrom"]), int(e["start"]), int(e["stop"]), int(e["strand"]), int(e["map_score"]), bool(ef & 2), bool(ef & 4), int(e["cz3_sub"]), int(e["tip_penalty"])), (r, got, e)
        assert got["match"] == m_exp, (r, got["match"], m_exp)
        # the SAM record of the read
        rd = psf.PyRead(got["chrom"], got["start"], got["stop"], L, got["map_score"], got["match"], True, got["strand"] == 1, got["perfect"], got["ambiguous"], False, False,
                        False, 0)
        line = psf.PySamLine(rd, 0, S, v14=bool(samcfg["version14"][0]), soft_clip=bool(samcfg["soft_clip"][0]), intron_limit=int(samcfg["intron_limit"][0]),
                             penalize_ambig=bool(samcfg["penalize_ambig"][0]))
        so = ref["sam"][r]
        cig = bytes(ref["cigar"][int(ref["cigar_off"][r]): int(ref["cigar_off"][r]) + int(so["cigar_len"])].view(np.uint8)).decode() if so["cigar_len"] >= 0 else None
        assert (line.flag, line.pos, line.mapq, line.cigar) == (int(so["flag"]), int(so["pos"]), int(so["mapq"]), cig), (r, line.flag, line.pos, line.mapq, line.cigar, so, cig)
        mapped += 1; with_indel += (b"D" in got["match"]) or (b"I" in got["match"])
    assert done > 110 and mapped > 95 and with_indel > 20, (done, mapped, skipped, with_indel)
