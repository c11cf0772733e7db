"""TEST INFRASTRUCTURE — an independent restatement (Python objects, from the Java text alone) of the pairing helpers of processReadPair:
  BBMapThread.pairSiteScoresInitial           current/align2/BBMapThread.java:736-940
  AbstractMapThread.pairSiteScoresFinal       current/align2/AbstractMapThread.java:1919-2095  (as called with trim = setScore = true, BBMapThread.java:1137)
  AbstractMapThread.canPair                   current/align2/AbstractMapThread.java:2097-2163
built on the SiteScore objects, comparators and Tools.trimSitesBelowCutoff of tests/pysitelist.py.  Shares no code with oracle/mapper_oracle.c."""
import numpy as np

import pysitelist as ps

F = np.float32
OUTER_DIST_MULT, OUTER_DIST_DIV = 14, 32
PLUS = 0


def _dists(s1, s2, require_correct):
    if require_correct and s1.strand != s2.strand:
        first_is_left = s1.strand == PLUS
    else:
        first_is_left = s1.start <= s2.start
    if first_is_left:
        return s2.start - s1.stop, s2.stop - s1.start
    return s1.start - s2.stop, s1.stop - s2.start


def _jdiv(a, b):
    """Java int division truncates toward zero."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q


def _walk_pairs(A, B, max_pair_dist):
    """The double loop both functions share: for every site of the first read, the sites of the mate on the same chromosome that start no further than
    max_pair_dist behind its end — with the reference's moving lower bound j (which is only advanced, never reset)."""
    ilimit, jlimit = len(A) - 1, len(B) - 1
    i = j = 0
    while i <= ilimit and j <= jlimit:
        s1, s2 = A[i], B[j]
        while j < jlimit and (s2.chrom < s1.chrom or (s2.chrom == s1.chrom and s1.start - s2.stop > max_pair_dist)):
            j += 1
            s2 = B[j]
        for k in range(j, jlimit + 1):
            s2 = B[k]
            if s2.chrom > s1.chrom or s2.start - s1.stop > max_pair_dist:
                break
            yield s1, s2
        i += 1


def pair_site_scores_initial(A, lenA, B, lenB, cfg, max_trim, trim=True):
    if not A or not B:
        return 0
    ps.sort_sites(A, ps.pcomp); ps.sort_sites(B, ps.pcomp)
    for s in A + B:
        s.pairedScore = 0
    max1 = max2 = -1
    max_read = max(lenA, lenB)
    outer_limit = (max_read * OUTER_DIST_MULT) // OUTER_DIST_DIV
    inner_limit = int(cfg["max_pair_dist"])
    apd = int(cfg["average_pair_dist"])
    expected = apd + lenA + lenB
    same_strand, require = bool(cfg["same_strand_pairs"]), bool(cfg["require_correct_strands"])
    perfect_pairs = 0
    for s1, s2 in _walk_pairs(A, B, inner_limit):
        inner, outer = _dists(s1, s2, require)
        if outer >= outer_limit and inner <= inner_limit:
            strand_ok = (s1.strand == s2.strand) == same_strand
            if strand_ok or not require:
                dev = abs(apd - inner)
                if strand_ok:
                    p1 = s1.score + 1 + max(1, _jdiv(s2.score, 2) - _jdiv(dev * s2.score, 32 * expected + 100))
                    p2 = s2.score + 1 + max(1, _jdiv(s1.score, 2) - _jdiv(dev * s1.score, 32 * expected + 100))
                else:
                    p1 = s1.score + max(0, _jdiv(s2.score, 16))
                    p2 = s2.score + max(0, _jdiv(s1.score, 16))
                got1 = p1 > s1.pairedScore
                if got1:
                    s1.pairedScore = p1
                    max1 = max(s1.score, max1)
                got2 = p2 > s2.pairedScore
                if got2:
                    s2.pairedScore = p2
                    max2 = max(s2.score, max2)
                if got1 and got2 and outer >= max_read and dev <= expected and s1.perfect and s2.perfect:
                    perfect_pairs += 1
    for s in A + B:
        if s.pairedScore > s.score:
            s.score = s.pairedScore
    if trim:
        if perfect_pairs > 0:
            ps.trim_sites_below_cutoff(A, int(F(max1) * F(.94)), False, True, 1, max_trim)
            ps.trim_sites_below_cutoff(B, int(F(max2) * F(.94)), False, True, 1, max_trim)
        else:
            if len(A) > 4:
                ps.trim_sites_below_cutoff(A, int(F(max1) * F(.9)), True, True, 1, max_trim)
            if len(B) > 4:
                ps.trim_sites_below_cutoff(B, int(F(max2) * F(.9)), True, True, 1, max_trim)
    return perfect_pairs


def pair_site_scores_final(A, lenA, B, lenB, cfg, max_trim):
    for s in A + B:
        s.pairedScore = 0
    if not A or not B:
        return
    ps.sort_sites(A, ps.pcomp); ps.sort_sites(B, ps.pcomp)
    max1 = max2 = -1
    mult1 = min(F(0.5), max(F(0.25), F(lenA) / (F(4) * F(lenB))))
    mult2 = min(F(0.5), max(F(0.25), F(lenB) / (F(4) * F(lenA))))
    outer_limit = (max(lenA, lenB) * OUTER_DIST_MULT) // OUTER_DIST_DIV
    mpd = int(cfg["max_pair_dist"]); apd = int(cfg["average_pair_dist"])
    expected = apd + lenA + lenB
    same_strand, require = bool(cfg["same_strand_pairs"]), bool(cfg["require_correct_strands"])
    for s1, s2 in _walk_pairs(A, B, mpd):
        inner, outer = _dists(s1, s2, require)
        if outer >= outer_limit and inner <= mpd:
            strand_ok = (s1.strand == s2.strand) == same_strand
            if strand_ok or not require:
                dev = abs(apd - inner)
                if strand_ok:
                    den = max(100, 10 * expected + 100)
                    p1 = s1.score + 1 + max(1, int(F(s2.score) * mult1) - _jdiv(dev * s2.score, den))
                    p2 = s2.score + 1 + max(1, int(F(s1.score) * mult2) - _jdiv(dev * s1.score, den))
                else:
                    p1 = s1.score + _jdiv(s2.score, 16)
                    p2 = s2.score + _jdiv(s1.score, 16)
                s1.pairedScore = max(s1.pairedScore, p1)
                s2.pairedScore = max(s2.pairedScore, p2)
                max1 = max(s1.score, max1)
                max2 = max(s2.score, max2)
    for s in A + B:
        if s.pairedScore > s.score:
            s.score = s.pairedScore
    f = min(F(cfg["secondary_site_score_ratio"]), F(0.95))
    ps.trim_sites_below_cutoff(A, int(F(max1) * f), False, True, 1, max_trim)
    ps.trim_sites_below_cutoff(B, int(F(max2) * f), False, True, 1, max_trim)


def can_pair(s1, s2, len1, len2, cfg):
    if s1.chrom != s2.chrom:
        return False
    require = bool(cfg["require_correct_strands"])
    if require and ((s1.strand == s2.strand) != bool(cfg["same_strand_pairs"])):
        return False
    inner, outer = _dists(s1, s2, require)
    return outer >= (max(len1, len2) * OUTER_DIST_MULT) // OUTER_DIST_DIV and inner <= int(cfg["max_pair_dist"])


def remove_low_quality_sites_paired(lst, max_sw, mult_single, mult_paired):
    """Tools.removeLowQualitySitesPaired (current/align2/Tools.java:934-958).  The list is edited in place; returns the number removed."""
    if not lst:
        return 0
    initial = len(lst)
    thresh = int(F(max_sw) * F(mult_single))
    thresh_paired = int(F(max_sw) * F(mult_paired))
    if lst[0].score < thresh_paired:
        del lst[:]
        return initial
    for i in range(len(lst) - 1, -1, -1):
        s = lst[i]
        if s.slowScore < (thresh_paired if s.pairedScore > 0 else thresh):
            del lst[i]
    return initial - len(lst)


def is_bad_pair(r, m, require_correct_strands, same_strand_pairs, maxdist):
    """Read.isBadPair (current/stream/Read.java:1305-1331); r, m: objects with mapped, paired, chrom, start, stop, strand (0 plus, 1 minus)."""
    if m is None or r.paired:
        return False
    if not r.mapped or not m.mapped:
        return False
    if r.chrom != m.chrom:
        return True
    inner = m.start - r.stop if r.start <= m.start else r.start - m.stop
    if inner > maxdist:
        return True
    if require_correct_strands and ((r.strand == m.strand) != same_strand_pairs):
        return True
    if not same_strand_pairs:
        if r.strand == 0 and m.strand == 1:
            if r.start >= m.stop:
                return True
        elif r.strand == 1 and m.strand == 0:
            if m.start >= r.stop:
                return True
    return False
