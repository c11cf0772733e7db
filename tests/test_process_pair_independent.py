"""CPU: BBMapThread.processReadPair END TO END from independent restatements only — the seeds and BBIndex.find of both mates, pairSiteScoresInitial, the paired
trimList, scoreNoIndels / findTipDeletions / scoreSlow per mate, rescue in both directions (quickRescue, slowRescue), removeLowQualitySitesPaired,
pairSiteScoresFinal, the paired clearzone, canPair, genMatchString per mate, removeDuplicateBestSites (tests/pyseed.py, pyfind.py, pysitelist.py, pypairing.py,
pyreadpolicies.py, pyscoreslow.py, pyrescue.py, pygenmatch.py / pyrealign.py; every fill by the reference's own C) — against the sequential C chain the CUDA mapper is
tested against (oracle/chain.map_pairs): each mate's locus, strand, mapScore, mapped / perfect / ambiguous / paired / rescued flags and match string.  Pairs with a gap
array on any candidate site, more than 16 candidates, X / Y / C left in a match string, or a status bit are left out."""
import functools

import numpy as np
import pytest

from bbmap_b200 import rescue as rs
from bbmap_b200 import sitelist as sl
from bbmap_b200 import workloads as wl
from bbmap_b200.index import pack_chromosomes
from bbmap_b200.keyring import default_cfg
from bbmap_b200.mapper import map_cfg
from oracle import chain

import pyclip
import pyfind
import pygenmatch
import pypairing as pp
import pyreadpolicies as prp
import pyrealign
import pyrescue
import pysam_fields as psf
import pyscoreslow
import pyseed
import pysitelist as ps

MIN_TRIM_PAIRED, MAX_TRIM = 2, 800
_sort = lambda lst: lst.sort(key=functools.cmp_to_key(ps.compare_to))


class Mate:
    def __init__(self, bases, qual):
        self.b, self.q, _ = pyseed.validate(bases, qual)
        self.m = pyseed.reverse_complement_bases(self.b)
        self.b8, self.m8 = np.array(self.b, np.int8), np.array(self.m, np.int8)
        self.L = len(self.b)
        self.max_sw = 70 + 100 * (self.L - 1)
        self.sites = []
        self.perfect = self.ambiguous = self.paired = self.mapped = False
        self.match = None
        self.map_score = 0
        self.top = None


def _paired_clearzone(m, pcfg):
    top = m.sites[0]
    if m.perfect:
        return int(pcfg["clearzonep"])
    F = np.float32
    if top.score >= int(F(m.max_sw) * F(pcfg["cz1b_scale"]) - F(pcfg["cz1b_flat"])):
        return int(pcfg["clearzone1"])
    if top.score >= int(F(m.max_sw) * F(pcfg["cz1c_scale"]) - F(pcfg["cz1c_flat"])):
        return int(pcfg["clearzone1b"])
    return int(pcfg["clearzone1c"])


def process_pair(oracle, py, R, Rsc, packed, m1, m2, refs, ref8, maxidx, single, scfg, pcfg, mcfg, wcfg, tcfg):
    mates = (m1, m2)
    ok = []
    for m in mates:
        s = pyseed.quick_map_seed(m.b, m.q, int(scfg["keylen"]), int(scfg["maxDesiredKeys"]), int(scfg["baseKeyHitScore"]), int(scfg["minApproxHitsToKeep"]),
                                  float(scfg["keyDensity"]), float(scfg["maxKeyDensity"]), float(scfg["minKeyDensity"]))
        ok.append(s is not None)
        if s is not None:
            found = py.find(bytes(m.b), s["baseScores"], s["offsets"], s["keyScores"])
            if found["gapfix"] or len(found["sites"]) > 16 or any(x.gaps is not None for x in found["sites"]):
                return "skip"
            m.sites = [ps.Site(x.chrom, x.strand, x.start, x.stop, x.hits, x.score, x.score, 0, 0, bool(x.perfect), bool(x.semiperfect), False, None) for x in found["sites"]]
            ps.remove_out_of_bounds(m.sites, m.L, maxidx, single, True, 2522)
    if not ok[0] and not ok[1]:
        return "discarded"
    pp.pair_site_scores_initial(m1.sites, m1.L, m2.sites, m2.L, mcfg, MAX_TRIM, trim=bool(pcfg["trim_list"]))
    if pcfg["trim_list"]:
        for m in mates:
            if len(m.sites) > MIN_TRIM_PAIRED:
                _sort(m.sites)
            ps.trim_list(m.sites, True, m.max_sw, False, MIN_TRIM_PAIRED, MAX_TRIM)
    for m in mates:
        for s in m.sites:
            s.score = s.quickScore
    for m in mates:
        if m.sites:
            near = prp.score_no_indels_read(m.sites, m.b, m.m, refs)
            _sort(m.sites)
            if near < 1:
                prp.find_tip_deletions_read(m.sites, m.b, m.m, m.q, refs, {1: 0}, int(tcfg["search_range"]), int(tcfg["slow_rescue_padding"]))
            if any(s.gaps is not None for s in m.sites):
                return "skip"
            pyscoreslow.score_slow(oracle, packed, m.sites, m.b8, m.m8, ref8, wcfg)
            ps.merge_duplicate_sites(m.sites, True)
    if mcfg["do_rescue"]:
        dist = min(int(mcfg["max_pair_dist"]), 2 * int(mcfg["average_pair_dist"]) + 100)
        for anchor, loose in ((m1, m2), (m2, m1)):
            if anchor.sites and any(s.pairedScore == 0 for s in anchor.sites):
                _sort(anchor.sites)
                pp.remove_low_quality_sites_paired(anchor.sites, anchor.max_sw, mcfg["min_ratio_pre_rescue"], mcfg["min_ratio_pre_rescue"])
                Rsc.rescue(anchor.sites, anchor.L, loose.sites, loose.b8, loose.m8, loose.q, dist)
                ps.merge_duplicate_sites(loose.sites, True)
    for m in mates:
        if len(m.sites) > 1:
            _sort(m.sites)
    for m in mates:
        pp.remove_low_quality_sites_paired(m.sites, m.max_sw, mcfg["min_ratio"], mcfg["min_ratio_paired"])
    pp.pair_site_scores_final(m1.sites, m1.L, m2.sites, m2.L, mcfg, MAX_TRIM)
    for m in mates:
        if m.sites:
            _sort(m.sites)
        top = m.sites[0] if m.sites else None
        m.perfect = bool(top is not None and (top.slowScore == m.max_sw or top.perfect))          # Read.setPerfectFlag without a match string
    for m in mates:
        if len(m.sites) > 1 and ps.count_top_scores(m.sites, _paired_clearzone(m, pcfg)) > 1:
            m.ambiguous = True
    if m1.sites and m2.sites and pp.can_pair(m1.sites[0], m2.sites[0], m1.L, m2.L, mcfg):
        m1.paired = m2.paired = True
    for m in mates:                                                    # setFromTopSite
        m.mapped = bool(m.sites)
        m.map_score = m.sites[0].slowScore if m.sites else 0
    for m, other in ((m1, m2), (m2, m1)):
        if m.sites:
            cs = [pyclip.ClipSite(s, None) for s in m.sites]
            was_paired = m.paired
            m.paired = pygenmatch.gen_match_string(R, cs, m.b8, m.m8, m.max_sw, mcfg, False, m.paired)
            if was_paired and not m.paired:
                other.paired = False
            m.sites = [c.s for c in cs]
            top = cs[0]
            m.match = bytes(top.match) if top.match is not None else None
            m.map_score = top.s.slowScore
            m.perfect = bool(top.s.perfect)
            m.top = top.s
    for m, other in ((m1, m2), (m2, m1)):
        if m.mapped and m.map_score <= 0:                              # r.mapScore<=0 && r.sites!=null: clearMapping
            m.mapped = False; m.sites = []; m.paired = False; other.paired = False; m.match = None; m.map_score = 0
    for m in mates:
        if len(m.sites) >= 2:                                          # removeDuplicateBestSites
            t = m.sites[0]
            while len(m.sites) > 1 and (m.sites[-1].chrom, m.sites[-1].strand, m.sites[-1].start, m.sites[-1].stop) == (t.chrom, t.strand, t.start, t.stop):
                m.sites.pop()
    return "ok"


@pytest.mark.parametrize("seed", [1101, 1102])
def test_process_read_pair_end_to_end(oracle, seed):
    if not oracle.has_reference:
        pytest.skip("oracle/_ref/libbbref.so not built (needs /root/reference)")
    rng = np.random.Generator(np.random.PCG64(seed))
    g = wl.ACGT[rng.integers(0, 4, size=90000, dtype=np.uint8)]
    unit = g[5000:5300].copy()
    for c in range(4):
        p = 12000 + 15000 * c; g[p:p + 300] = unit
    cb, co, table = pack_chromosomes([g])
    npairs = 110
    RP = wl.make_mapping_reads(cb, co, table, npairs, L=150, seed=seed + 1, sub_rate=0.015, indel_rate=0.02 / 3)
    bases, qual, off = RP["bases"].copy(), RP["qual"].copy(), RP["off"]
    for p in range(0, npairs, 4):                                      # a damaged mate: rescue has to find it
        r = 2 * p + int(rng.integers(0, 2)); a, b = int(off[r]), int(off[r + 1])
        k = rng.choice(150, size=int(rng.integers(12, 30)), replace=False)
        bases[a + k] = wl.ACGT[rng.integers(0, 4, size=len(k), dtype=np.uint8)]
    for p in range(1, npairs, 9):                                      # an unmappable mate
        r = 2 * p + 1; a, b = int(off[r]), int(off[r + 1]); bases[a:b] = wl.ACGT[rng.integers(0, 4, size=150, dtype=np.uint8)]
    for p in range(3, npairs, 7):                                      # improper pairs: the mate comes from 40-60 kbp away, or from the same strand next door
        r = 2 * p + 1; a, b = int(off[r]), int(off[r + 1])
        t = RP["truth"][2 * p]
        if p % 2:
            q = (int(t[2]) - 8000 + int(rng.integers(40000, 60000))) % (len(g) - 200)
            rd = g[q:q + 150]
            bases[a:b] = rd if rng.random() < 0.5 else wl.revcomp(rd)
        else:
            q = max(0, min(len(g) - 200, int(t[2]) - 8000 + int(rng.integers(200, 400))))
            rd = g[q:q + 150]
            bases[a:b] = rd if int(t[1]) == 0 else wl.revcomp(rd)       # same strand as mate 1
    qual = rng.integers(14, 41, size=len(bases)).astype(np.uint8)
    idx = oracle.index_build(cb, co, 13, -1)
    ref = chain.map_pairs(oracle, idx, cb, co, table, bases, qual, off)
    icfg, blocks, counts, hist = idx
    py = pyfind.BBIndexPy(icfg, blocks, counts, hist, cb, co, quit_after_two_perfects=False)
    scfg = default_cfg()[0]
    pcfg = sl.policy_cfg()[0]; mcfg_arr = map_cfg(paired=1); mcfg = mcfg_arr[0]
    wcfg = sl.slow_cfg(paired=1, min_ratio=mcfg["min_ratio"], min_ratio_pre_rescue=mcfg["min_ratio_pre_rescue"])[0]
    tcfg = rs.tipdel_cfg()[0]
    ref8 = np.ascontiguousarray(cb).view(np.int8)[int(co[0]): int(co[1])]
    refs = {1: ref8.tolist()}
    R = pyrealign.Realigner(oracle, ref8)
    Rsc = pyrescue.Rescuer(oracle, ref8, mcfg, search_range=int(tcfg["search_range"]), slow_rescue_padding=int(tcfg["slow_rescue_padding"]))
    packed = oracle.new_packed(601, 3000)
    S = psf.Scaffolds([(c, s, ln) for c, s, ln in table], 300)
    single = lambda c, a, b: S.is_single(c, a, b)
    maxidx = {1: len(ref8) - 1}
    ms = ref["match_stride"]
    done = skipped = paired = rescued = unpaired_both = 0
    for p in range(npairs):
        ra, rb_ = 2 * p, 2 * p + 1
        if ref["recs"]["status"][ra] or ref["recs"]["status"][rb_]:
            skipped += 1
            continue
        mk = lambda r: Mate([int(x) for x in bases[int(off[r]): int(off[r + 1])].view(np.int8)], [int(x) for x in qual[int(off[r]): int(off[r + 1])].view(np.int8)])
        m1, m2 = mk(ra), mk(rb_)
        res = process_pair(oracle, py, R, Rsc, packed, m1, m2, refs, ref8, maxidx, single, scfg, pcfg, mcfg, wcfg, tcfg)
        if res == "skip":
            skipped += 1
            continue
        bad = False
        for r, m in ((ra, m1), (rb_, m2)):
            e = ref["recs"][r]; mlen = int(e["match_len"])
            m_exp = ref["match"][r * ms: r * ms + mlen].tobytes() if mlen > 0 else None
            if (m_exp is not None and any(c in m_exp for c in b"XYC")) or (m.match is not None and any(c in m.match for c in b"XYC")):
                bad = True
        if bad:
            skipped += 1
            continue
        for r, m in ((ra, m1), (rb_, m2)):
            e = ref["recs"][r]; ef = int(e["flags"]); mlen = int(e["match_len"])
            m_exp = ref["match"][r * ms: r * ms + mlen].tobytes() if mlen > 0 else None
            assert m.mapped == bool(ef & 1), (p, r, res, m.mapped, e)
            if not m.mapped:
                continue
            top = m.sites[0]
            assert (top.chrom, top.start, top.stop, top.strand, m.map_score, m.perfect, m.ambiguous, m.paired, bool(top.rescued)) == \
                   (int(e["chrom"]), int(e["start"]), int(e["stop"]), int(e["strand"]), int(e["map_score"]), bool(ef & 2), bool(ef & 4), bool(ef & 8), bool(ef & 16)), (p, r, e)
            assert m.match == m_exp, (p, r, m.match, m_exp)
            rescued += bool(top.rescued)
        done += 1; paired += m1.paired; unpaired_both += (m1.mapped and m2.mapped and not m1.paired)
    assert done > 80 and paired > 50 and rescued > 4 and unpaired_both > 5, (done, skipped, paired, rescued, unpaired_both)
