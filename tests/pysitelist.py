"""TEST INFRASTRUCTURE — an independent restatement (Python objects in Python lists, written from the Java text alone) of the per-read site-list policies of
the unpaired mapping loop:
  the TRIM_LIST block of processRead          current/align2/BBMapThread.java:426-429  (Collections.sort + trimList :140-249)
  Tools.trimSiteList / trimSitesBelowCutoff   current/align2/Tools.java:654-672, 1106-1160
  the tail of processRead                     current/align2/BBMapThread.java:478-553  (mergeDuplicateSites, sort, setPerfectFlag, clearzone / ambiguity,
                                              score gate, removeLowQualitySitesUnpaired)
  Tools.mergeDuplicateSites / countTopScores / removeLowQualitySitesUnpaired / condenseStrict   Tools.java:542-568, 696-758, 911-927, 985-1003
  SiteScore.compareTo / PCOMP / positionalMatch / setSlowScore / setPairedScore   current/stream/SiteScore.java:55-76, 353-395, 962-983
  Read.setPerfectFlag                         current/stream/Read.java:2494-2513 (no match string exists yet at that point: testMatchPerfection(false) is false)
Shares no code with oracle/sitelist_oracle.c.  Java float arithmetic is numpy float32, one rounding per operation."""
import functools

import numpy as np

F = np.float32


class Site:
    def __init__(self, chrom, strand, start, stop, hits, score, quick, slow, paired, perfect, semiperfect, rescued, gaps, tag=None):
        self.chrom, self.strand, self.start, self.stop, self.hits = chrom, strand, start, stop, hits
        self.score, self.quickScore, self.slowScore, self.pairedScore = score, quick, slow, paired
        self.perfect, self.semiperfect, self.rescued, self.gaps, self.tag = perfect, semiperfect, rescued, gaps, tag

    def set_slow_score(self, x):
        if x <= 0:
            self.pairedScore = self.slowScore = x
        elif self.pairedScore <= 0:
            self.slowScore = x
        elif self.pairedScore > 0:
            self.pairedScore = x + (self.pairedScore - self.slowScore) if self.slowScore > 0 else x + 1
        self.slowScore = x

    def positional_match(self, b, test_gaps):
        if (self.chrom, self.strand, self.start, self.stop) != (b.chrom, b.strand, b.start, b.stop):
            return False
        if not test_gaps or (self.gaps is None and b.gaps is None):
            return True
        if (self.gaps is None) != (b.gaps is None):
            return False
        return list(self.gaps) == list(b.gaps)


def compare_to(a, b):
    for x in (b.score - a.score, b.slowScore - a.slowScore, b.pairedScore - a.pairedScore, b.quickScore - a.quickScore, a.chrom - b.chrom):
        if x != 0:
            return x
    return a.start - b.start


def pcomp(a, b):
    for x, y in ((a.chrom, b.chrom), (a.start, b.start), (a.stop, b.stop), (a.strand, b.strand)):
        if x != y:
            return x - y
    for x, y in ((a.score, b.score), (a.slowScore, b.slowScore), (a.quickScore, b.quickScore)):
        if x != y:
            return y - x
    if a.perfect != b.perfect:
        return -1 if a.perfect else 1
    if a.rescued != b.rescued:
        return 1 if a.rescued else -1
    return 0


def sort_sites(lst, cmp=compare_to):
    lst.sort(key=functools.cmp_to_key(cmp))          # list.sort is stable, like Collections.sort


def trim_sites_below_cutoff(ssl, cutoff, retain_paired, retain_semiperfect, min_retain, max_retain):
    if len(ssl) <= min_retain:
        return
    del ssl[max_retain:]
    removed = 0
    max_remove = len(ssl) - min_retain
    for i in range(len(ssl) - 1, -1, -1):
        ss = ssl[i]
        if not retain_semiperfect or not ss.semiperfect:
            if ss.score < cutoff and (not retain_paired or ss.pairedScore <= 0):
                ssl[i] = None
                removed += 1
                if removed >= max_remove:
                    break
    if removed:
        ssl[:] = [x for x in ssl if x is not None]


def trim_site_list(ssl, fraction, retain_paired, retain_semiperfect, min_retain, max_retain):
    if not ssl:
        return -999999
    if len(ssl) == 1:
        return ssl[0].score
    if 1 < min_retain < len(ssl):
        mx = ssl[0].score
    else:
        mx = max(max(s.score for s in ssl), -999999)
    cutoff = int(F(mx) * F(fraction))
    trim_sites_below_cutoff(ssl, cutoff, retain_paired, retain_semiperfect, min_retain, max_retain)
    return mx


def trim_list(lst, retain_paired, max_score, special_case_perfect, min_retain, max_retain):
    """BBMapThread.trimList with USE_AFFINE_SCORE."""
    if not lst:
        return -99999
    if len(lst) == 1:
        return lst[0].score
    t = lambda f, mr=min_retain: trim_site_list(lst, f, retain_paired, True, mr, max_retain)
    highest = t(.6)
    if highest == max_score and special_case_perfect:
        t(.94)
        if len(lst) > 8:
            t(.99)
        return highest
    mstr2 = 1 if min_retain <= 1 else min_retain + 1
    for size, frac in ((4, .65), (8, .7), (12, .75), (16, .8), (20, .85), (24, .9), (32, .95)):
        if len(lst) > size:
            t(frac)
    for size, frac in ((40, .97), (48, .99)):
        if len(lst) > size:
            t(frac, mstr2)
    return highest


def trim_policy(sites, read_len, cfg):
    """The TRIM_LIST block of processRead.  Returns the highest quick score (or None when the block is skipped)."""
    max_sw = 70 + (read_len - 1) * 100
    if cfg["trim_list"] and len(sites) > 1:
        if cfg["min_trim_sites_to_retain"] > 1:
            sort_sites(sites)
        return trim_list(sites, False, max_sw, True, int(cfg["min_trim_sites_to_retain"]), int(cfg["max_trim_sites_to_retain"]))
    return None


def merge_duplicate_sites(lst, merge_different_gaps=True):
    if len(lst) < 2:
        return 0
    sort_sites(lst, pcomp)
    removed = 0
    a = lst[0]
    for i in range(1, len(lst)):
        b = lst[i]
        exact = a.positional_match(b, True)
        if exact or (merge_different_gaps and a.positional_match(b, False)):
            better = a
            if not exact:
                if a.score != b.score:
                    better = a if a.score > b.score else b
                elif a.slowScore != b.slowScore:
                    better = a if a.slowScore > b.slowScore else b
                elif a.pairedScore != b.pairedScore:
                    better = a if a.pairedScore > b.pairedScore else b
            a.set_slow_score(max(a.slowScore, b.slowScore))
            a.pairedScore = 0 if (a.pairedScore <= a.slowScore and b.pairedScore <= a.slowScore) else max(0, a.pairedScore, b.pairedScore)
            a.score = max(a.score, b.score)
            a.perfect = a.perfect or b.perfect
            a.semiperfect = a.semiperfect or b.semiperfect
            if not exact:
                a.gaps = better.gaps
            removed += 1
            lst[i] = None
        else:
            a = b
    if removed:
        lst[:] = [x for x in lst if x is not None]
    return removed


def count_top_scores(lst, thresh):
    if not lst:
        return 0
    top = lst[0]
    limit = top.score - thresh
    count = 1
    for s in lst[1:]:
        if s.score < limit:
            break
        if top.start != s.start and top.stop != s.stop:
            count += 1
    return count


def final_policy(sites, read_len, cfg):
    """The tail of processRead after scoreSlow.  Returns dict(mapped, perfect, ambiguous, clearzone, best_sites); `sites` is edited in place."""
    max_sw = 70 + (read_len - 1) * 100
    out = dict(mapped=False, perfect=False, ambiguous=False, clearzone=None, best_sites=None)
    if sites:
        merge_duplicate_sites(sites)
        sort_sites(sites)
    top = sites[0] if sites else None
    out["perfect"] = bool(top is not None and (top.slowScore == max_sw or top.perfect))
    if len(sites) > 1:
        score = top.score
        if out["perfect"]:
            clearzone = int(cfg["clearzonep"])
        else:
            cz1b = F(max_sw) * F(cfg["cz1b_scale"]) - F(cfg["cz1b_flat"])
            cz1c = F(max_sw) * F(cfg["cz1c_scale"]) - F(cfg["cz1c_flat"])
            if F(score) > cz1b:
                clearzone = int((F((max_sw - score) * int(cfg["clearzone1b"])) + (F(score) - cz1b) * F(int(cfg["clearzone1"]))) / (F(max_sw) - cz1b))
            elif F(score) > cz1c:
                clearzone = int(((cz1b - F(score)) * F(int(cfg["clearzone1c"])) + (F(score) - cz1c) * F(int(cfg["clearzone1b"]))) / (cz1b - cz1c))
            else:
                clearzone = int(cfg["clearzone1c"])
        out["clearzone"] = clearzone
        nbest = count_top_scores(sites, clearzone)
        if nbest > 1:
            out["ambiguous"] = True
        else:
            e, lim1e = int(cfg["clearzone1e"]), int(cfg["clearzone_limit1e"])
            lim = (int(F(4) * F(lim1e)) if out["perfect"] else (2 * lim1e if score + e >= max_sw else lim1e)) + 1
            if len(sites) > lim and clearzone < e:
                nbest = count_top_scores(sites, e)
                if nbest > lim:
                    out["ambiguous"] = True
        out["best_sites"] = nbest
    if sites:
        lim = int(F(max_sw) * F(cfg["min_align_ratio"]))
        if sites[0].score < lim:
            del sites[:]
        else:
            thresh = min(lim, max(1, lim - int(cfg["clearzone3"])))
            if sites[0].score < thresh:
                del sites[:]
            else:
                for i in range(len(sites) - 1, 1, -1):
                    if sites[i].slowScore < thresh:
                        del sites[i]
    out["mapped"] = len(sites) > 0
    return out


def calc_tip_score_penalty(mapped, match, bases, map_score, tiplen):
    """AbstractMapThread.calcTipScorePenalty (current/align2/AbstractMapThread.java:2499-2567) on a long-format match string.  Returns (penalty, status):
    status 1 = the string ends before tiplen+1 read symbols were seen (the Java would run off the array), 2 = short format (digits)."""
    L = len(bases)
    max_score = 70 + (L - 1) * 100
    if not mapped or match is None or L < 2 * tiplen:
        return 0, 0

    def scan(order):
        points, cpos, prev = 0, 0, "m"
        it = iter(order)
        while cpos <= tiplen:
            try:
                b = chr(match[next(it)])
            except StopIteration:
                return None
            if b == "m":
                cpos += 1
            elif b == "D":
                if prev != "D":
                    points += 2 * (tiplen + 2 - cpos)
            elif b in "NC":
                points += tiplen + 2 - cpos
                cpos += 1
            elif b.isdigit():
                return "short"
            else:
                points += 2 * (tiplen + 2 - cpos)
                cpos += 1
            prev = b
        return points

    left = scan(range(len(match)))
    if left == "short":
        return 0, 2
    if left is None:
        return 0, 1
    right = scan(range(len(match) - 1, -1, -1))
    if right is None:
        return 0, 1
    points = left + right
    last = L - 1
    b = bases[0]
    if b != ord("N") and b == bases[1]:
        i = 2
        while i <= tiplen and bases[i] == b:
            points += 1; i += 1
    b = bases[last]
    if b != ord("N") and b == bases[last - 1]:
        i = last - 2
        while i >= last - tiplen and bases[i] == b:
            points += 1; i -= 1
    if points < 1:
        return 0, 0
    f = (F(80) * F(points)) / (F(points) + F(80))
    penalty = int(f * F(.0022) * F(max_score))
    max_penalty = map_score - (abs(max_score) // 10 if max_score >= 0 else -(abs(max_score) // 10))
    if max_penalty <= 0:
        return 0, 0
    return min(penalty, max_penalty), 0


def apply_score_penalty(sites, penalty):
    if penalty > 0:
        for s in sites:
            s.set_slow_score(s.slowScore - penalty)
            s.score = s.score - penalty


# ---------------- GapTools (current/align2/GapTools.java:10-14, 26-91, 126-175) and removeOutOfBounds (AbstractMapThread.java:2444-2479) ----------------
GAPBUFFER2, GAPLEN, MINGAP = 128, 128, 256


def _jdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def calc_gref_len(a, b, gaps):
    total = b - a + 1
    if gaps is None:
        return total
    for i in range(2, len(gaps), 2):
        total -= max(0, _jdiv(gaps[i] - gaps[i - 1] - GAPBUFFER2, GAPLEN)) * (GAPLEN - 1)
    return total


def fix_gaps(a, b, gaps, min_gap=MINGAP):
    """Returns the gap array (the same list object, edited in place, where the reference keeps its array) or None."""
    if gaps is None:
        return None
    if not (a <= gaps[-1] and b >= gaps[0]):               # Tools.overlap
        return None
    changed = 0
    if gaps[0] != a:
        gaps[0] = a; changed += 1
    if gaps[-1] != b:
        gaps[-1] = b; changed += 1
    for i in range(len(gaps)):
        if gaps[i] < a:
            gaps[i] = a; changed += 1
        elif gaps[i] > b:
            gaps[i] = b; changed += 1
    for i in range(1, len(gaps)):
        if gaps[i - 1] > gaps[i]:
            gaps[i] = gaps[i - 1]; changed += 1
    if changed == 0:
        return gaps
    gaps[0], gaps[-1] = a, b
    remove = 0
    for i in range(0, len(gaps), 2):
        gaps[i] = min(max(gaps[i], a), b); gaps[i + 1] = min(max(gaps[i + 1], a), b)
        if gaps[i] == gaps[i + 1]:
            remove += 1
    if remove == 0:
        return gaps
    ranges = [[gaps[i], gaps[i + 1]] for i in range(0, len(gaps), 2)]       # fixGaps2
    for i in range(1, len(ranges)):
        r1, r2 = ranges[i - 1], ranges[i]
        if r1 is not None and r2[0] - r1[1] <= min_gap:
            r2[0] = min(r1[0], r2[0]); r2[1] = max(r1[1], r2[1])
            ranges[i - 1] = None
    ranges = [r for r in ranges if r is not None]
    if len(ranges) < 2:
        return None
    return [x for r in ranges for x in r]


def set_stop(site, b):
    """SiteScore.setStop (stream/SiteScore.java:944-951): with a gap array the array is always re-fixed."""
    site.stop = b
    if site.gaps is not None:
        site.gaps[-1] = b
        site.gaps = fix_gaps(site.start, site.stop, site.gaps)


def remove_out_of_bounds(sites, read_len, chrom_max_index, is_single_scaffold, sam_out, expected_len_limit):
    """chrom_max_index: chrom -> ChromosomeArray.maxIndex; is_single_scaffold(chrom, start, stop) -> bool.  Returns the number of sites removed."""
    initial = len(sites)
    i = 0
    while i < len(sites):
        ss = sites[i]
        if ss.start < 0 or ss.stop > chrom_max_index[ss.chrom]:
            del sites[i]
            continue
        if sam_out and not is_single_scaffold(ss.chrom, ss.start, ss.stop):
            del sites[i]
            continue
        if calc_gref_len(ss.start, ss.stop, ss.gaps) >= expected_len_limit:
            set_stop(ss, ss.start + min(read_len + 40, expected_len_limit))
            if ss.gaps is not None:
                ss.gaps = fix_gaps(ss.start, ss.stop, ss.gaps)          # GapTools.fixGaps(ss) assigns its result to ss.gaps
        i += 1
    return initial - len(sites)


def check_gaps(site):
    """SiteScore.CHECKGAPS (stream/SiteScore.java:952-959)."""
    g = site.gaps
    if g is None:
        return True
    if len(g) == 0 or len(g) % 2 == 1:
        return False
    if any(g[i - 1] > g[i] for i in range(1, len(g))):
        return False
    return g[0] == site.start and g[-1] == site.stop


def set_limits(site, a, b):
    """SiteScore.setLimits (stream/SiteScore.java:905-914)."""
    site.start, site.stop = a, b
    if site.gaps is not None:
        site.gaps[0] = a
        site.gaps[-1] = b
        if not check_gaps(site):
            site.gaps = fix_gaps(site.start, site.stop, site.gaps)
