"""TEST INFRASTRUCTURE — an independent restatement (Python, from the Java text) of AbstractMapThread.genMatchString / genMatchStringForSite for the default flag set
(current/align2/AbstractMapThread.java:860-1068; GEN_MATCH_FAST, no secondary alignments, USE_SS_MATCH_FOR_PRIMARY) on sites without a gap array: which sites get
a match string, the two realign_new calls per site, the re-sort loop with mergeDuplicateSites, which sites are dropped, the `paired` flag.  Built on
tests/pyrealign.py (fills by the reference's own C), tests/pyclip.py and tests/pysitelist.py.  Shares no code with oracle/mapper_oracle.c."""
import numpy as np

import pyclip
import pysitelist as ps

F = np.float32
MAX_COLUMNS = 3000


def gen_match_string_for_site(R, cs, basesP8, basesM8, max_sw, cfg):
    s = cs.s
    bases8 = basesP8 if s.strand == 0 else basesM8
    bases = bases8.tolist()
    mult = F(cfg["min_ratio_paired"]) if cfg["paired"] else F(cfg["min_ratio"])
    mult = mult * F(1)
    min_msa_limit = -1 + int(mult * F(max_sw))
    pad0 = int(cfg["slow_align_padding"])
    max_indel = int(cfg["max_indel"])
    if s.perfect:
        cs.match = [ord("m")] * len(bases)
    else:
        old = s.slowScore
        padding = 0 if (s.perfect or s.semiperfect) else max(pad0, 6)
        R.realign(cs, bases8, padding, 1, min_msa_limit, max_indel < 1, False)
        lp, rp = cs.left_padding_needed(4, 5), cs.right_padding_needed(4, 5)
        if s.slowScore < old or lp > 0 or rp > 0:
            extra = (80 if max_indel > 0 else 20) + pad0
            remaining = MAX_COLUMNS - (s.stop - s.start + 1) - 2
            extra = max(0, min(_jdiv(remaining, 2), extra))
            R.realign(cs, bases8, extra, 2, min_msa_limit, False, True)
        if max_sw == s.slowScore:                       # SiteScore.setPerfectFlag(maxScore, bases)
            s.perfect = s.semiperfect = True
        else:
            pyclip.set_perfect(s, bases, R.ref)
    cs.clip_tip_indels(bases, R.ref, 4, 10)
    return s.slowScore


def gen_match_string(R, sites, basesP8, basesM8, max_sw, cfg, set_ss_score, paired):
    """sites: list of pyclip.ClipSite (match None), sorted as the reference has them; edited in place.  Returns the read's `paired` flag afterwards."""
    if not sites:
        return paired
    best = -(1 << 31)
    changed = 0
    i = 0
    while i < len(sites):
        cs = sites[i]
        if i > 0 and best >= cs.s.slowScore:
            break
        old_slow, old_score = cs.s.slowScore, cs.s.score
        if cs.match is None:
            gen_match_string_for_site(R, cs, basesP8, basesM8, max_sw, cfg)
            if set_ss_score:
                cs.s.score = cs.s.slowScore
        if i > 0 and cs.match is None and not paired:
            del sites[i]
        else:
            if old_score != cs.s.score or old_slow != cs.s.slowScore:
                changed += 1
            best = max(cs.s.slowScore, best)
        i += 1
    in_order = all(sites[k].s.score <= sites[k - 1].s.score for k in range(1, len(sites)))
    needs_sorting = changed > 0 and not in_order
    while needs_sorting:
        needs_sorting = False
        R.resorts = getattr(R, "resorts", 0) + 1
        top = sites[0]
        _merge_duplicates_exact(sites)
        import functools
        sites.sort(key=functools.cmp_to_key(lambda a, b: ps.compare_to(a.s, b.s)))
        i = 0
        while i < len(sites):
            cs = sites[i]
            if cs.match is None:
                gen_match_string_for_site(R, cs, basesP8, basesM8, max_sw, cfg)
                if set_ss_score:
                    cs.s.score = cs.s.slowScore
                if i > 0 and cs.match is None:
                    del sites[i]
                else:
                    needs_sorting = True
                i -= 1
            if i > 0 or True:                              # !PRINT_SECONDARY_ALIGNMENTS: only the first position is looked at
                break
            i += 1
        if paired and sites[0] is not top:
            paired = False
    return paired


def _merge_duplicates_exact(sites):
    """Tools.mergeDuplicateSites(list, false, false) on ClipSites: same position AND same gaps; the survivor keeps its own match string."""
    if len(sites) < 2:
        return
    import functools
    sites.sort(key=functools.cmp_to_key(lambda a, b: ps.pcomp(a.s, b.s)))
    a = sites[0]
    for i in range(1, len(sites)):
        b = sites[i]
        if a.s.positional_match(b.s, True):
            a.s.set_slow_score(max(a.s.slowScore, b.s.slowScore))
            a.s.pairedScore = 0 if (a.s.pairedScore <= a.s.slowScore and b.s.pairedScore <= a.s.slowScore) else max(0, a.s.pairedScore, b.s.pairedScore)
            a.s.score = max(a.s.score, b.s.score)
            a.s.perfect = a.s.perfect or b.s.perfect
            a.s.semiperfect = a.s.semiperfect or b.s.semiperfect
            sites[i] = None
        else:
            a = b
    sites[:] = [x for x in sites if x is not None]


def _jdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q
