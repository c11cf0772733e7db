"""TEST INFRASTRUCTURE — an independent restatement (Python, from the Java text) of two per-read steps of processRead / processReadPair:
  AbstractMapThread.scoreNoIndels(Read, ...)      current/align2/AbstractMapThread.java:762-855   (QUICK_MATCH_STRINGS off)
  AbstractMapThread.findTipDeletions(Read, ...)   :1073-1105
on SiteScore objects (tests/pysitelist.Site), with MSA.scoreNoIndels of tests/pygapped.py, setPerfect of tests/pyclip.py and the per-site tip search of
tests/pyrescue.py.  Sites without a gap array keep None; scoreNoIndels drops the gap array of a near-perfect site as the reference does."""
import pyclip
import pygapped
import pyrescue
import pysitelist as ps


def score_no_indels_read(sites, basesP, basesM, refs_by_chrom):
    """sites of one read (edited in place); basesP/basesM: lists of ints; refs_by_chrom: chrom -> list of ints.  Returns the reference's return value
    (number of near-perfect scores, negated when slow alignment is forced)."""
    if not sites:
        return 0
    L = len(basesP)
    max_sw = 70 + (L - 1) * 100
    max_imperfect = max_sw + min(-472, -395 - 100)
    near = 0
    force_slow = False
    for ss in sites:
        old = ss.score
        bases = basesP if ss.strand == 0 else basesM
        ref = refs_by_chrom[ss.chrom]
        if ss.perfect:
            near += 1
            ss.set_slow_score(max_sw)
            ss.score = max_sw
            ss.gaps = None
        else:
            sc = pygapped.score_no_indels(bases, ref, ss.start)
            if sc < old and old >= max_imperfect and ss.stop - ss.start + 1 != L:
                sc2 = pygapped.score_no_indels(bases, ref, ss.stop - L + 1)
                if sc2 >= max_imperfect:
                    sc = sc2
                    _set_start(ss, ss.stop - L + 1)
                    pyclip.set_perfect(ss, bases, ref)
            ss.set_slow_score(sc)
            ss.score = sc
            if sc >= max_imperfect:
                near += 1
                ps.set_stop(ss, ss.start + L - 1)
                ss.gaps = None
                if sc >= max_sw:
                    ss.perfect = ss.semiperfect = True
                else:
                    pyclip.set_perfect(ss, bases, ref)
            elif old >= max_imperfect:
                force_slow = True
    return -near if force_slow else near


def _set_start(ss, a):
    """SiteScore.setStart (stream/SiteScore.java:935-943)."""
    ss.start = a
    if ss.gaps is not None:
        ss.gaps[0] = a
        if ss.gaps[0] > ss.gaps[1]:
            ss.gaps = ps.fix_gaps(ss.start, ss.stop, ss.gaps)


def find_tip_deletions_read(sites, basesP, basesM, quality, refs_by_chrom, min_index_by_chrom, search_range, slow_rescue_padding):
    """Returns the number of sites changed."""
    L = len(basesP)
    max_sw = 70 + (L - 1) * 100
    max_imperfect = max_sw + min(-472, -395 - 100)
    n = pyrescue.TIPLEN
    if quality is None:
        right = left = True
    elif n > len(quality):
        right = left = False                       # the Read helpers return 0 for n > length
    else:
        q = [int(x) for x in quality]
        tail, head = q[len(q) - n:], q[:n]
        right = min(tail) >= pyrescue.TIP_MIN_Q and sum(max(x, 0) for x in tail) // n >= pyrescue.TIP_AVG_Q
        left = min(head) >= pyrescue.TIP_MIN_Q and sum(max(x, 0) for x in head) // n >= pyrescue.TIP_AVG_Q
    if not right and not left:
        return 0
    changed_sites = 0
    for ss in sites:
        bases = basesP if ss.strand == 0 else basesM
        ref = refs_by_chrom[ss.chrom]
        if not ss.semiperfect and ss.slowScore < max_imperfect:
            if ss.gaps is not None:
                raise NotImplementedError("gapped site")
            if pyrescue.find_tip_deletions(ss, bases, ref, min_index_by_chrom[ss.chrom], max_imperfect, right, left, search_range, slow_rescue_padding):
                changed_sites += 1
                ss.set_slow_score(pygapped.score_no_indels(bases, ref, ss.start))
                if ss.slowScore == max_sw:
                    ss.stop = ss.start + L - 1
                    ss.perfect = ss.semiperfect = True
                else:
                    ss.perfect = False
                    pyclip.set_perfect(ss, bases, ref)
    return changed_sites
