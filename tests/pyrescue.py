"""TEST INFRASTRUCTURE — an independent restatement (Python, from the Java text) of mate rescue for the default flag set:
  AbstractMapThread.rescue          current/align2/AbstractMapThread.java:1144-1230
  AbstractMapThread.slowRescue      :1232-1296   (QUICK_MATCH_STRINGS off)
  AbstractMapThread.findTipDeletions(SiteScore, ...)   :1107-1141
  quickRescue                       :2303-2405   — stated over all candidate starts at once (numpy) + the reference's selection rule
  findTipDeletionsRight / Left      :2178-2294   — the order-free brute-force statement
  Read.min/avgQualityFirst/LastNBases   current/stream/Read.java:1759-1813
Alignments: MSA.fillAndScoreLimited of tests/pygapped.py (fills by the reference's own C); setPerfect: tests/pyclip.py.  Shares no code with oracle/."""
import numpy as np

import pyclip
import pygapped
import pysitelist as ps

F = np.float32
TIPLEN, TIP_MIN_Q, TIP_AVG_Q = 8, 6, 14
ALIGN_COLUMNS = 3000


def quick_rescue(bases8, ref8, min_index, chrom, strand, loc, search_dist, search_right, ideal_start, max_mm):
    L = len(bases8)
    if L < 10:
        return None
    lo, hi = (max(min_index, loc), min(len(ref8) - L, loc + search_dist)) if search_right else (max(min_index, loc - search_dist), min(len(ref8) - L, loc))
    if hi < lo:
        return None
    win = np.lib.stride_tricks.sliding_window_view(ref8[lo: hi + L], L)
    eq = (win == bases8[None, :]) & (bases8 != ord("N"))[None, :]
    mism = L - eq.sum(axis=1)
    cur = np.zeros(len(win), np.int64); contig = np.zeros(len(win), np.int64)
    for j in range(L):
        contig = np.where(eq[:, j], contig, np.maximum(contig, cur))
        cur = np.where(eq[:, j], cur + 1, 0)
    score = (L - mism) + contig
    limit = max_mm + 1
    first = lo
    best = (-1, 0, 0, 1 << 40)
    s = lo if search_right else hi
    while lo <= s <= hi:
        k = s - first
        m, sc, ad = int(mism[k]), int(score[k]), abs(s - ideal_start)
        if m <= limit and (sc > best[2] or (sc == best[2] and ad < best[3])):
            best = (s, m, sc, ad)
            limit = m
            if m == 0:
                if search_right:
                    hi = min(hi, ideal_start + ad)
                else:
                    lo = max(lo, ideal_start - ad)
        s += 1 if search_right else -1
    if best[0] < 0:
        return None
    start, mm = best[0], best[1]
    score_out = 70 + 100 * (L - 1 - mm)
    site = ps.Site(chrom, strand, start, start + L - 1, 0, score_out, score_out, 0, 0, False, False, True, None)
    pyclip.set_perfect(site, bases8.tolist(), ref8.tolist())
    site.set_slow_score(mm)
    return site


def _tip(bases, ref, mn, orig, dist, tiplen, right):
    L = len(bases)
    if right:
        if orig < mn + tiplen - 1:
            return 0
        differs = lambda s, j: bases[L - 1 - j] != ref[s - j]
    else:
        if orig + tiplen >= len(ref) or mn >= orig:
            return 0
        differs = lambda s, j: bases[j] != ref[s + j]
    om = last = contig = 0
    for i in range(tiplen):
        if contig >= 5:
            break
        if differs(orig, i):
            om += 1; last = i; contig = 0
        else:
            contig += 1
    if om < 3:
        return 0
    tl = last + 1
    if tl < 4:
        return 0
    if right:
        starts = range(orig + 1, min(len(ref) - 1, orig + min(dist, 30 * om)) + 1)
    else:
        starts = range(orig - 1, max(mn, orig - min(dist, 16 + 16 * om + 8 * tl)) - 1, -1)
    best_m, best_s = om, None
    for s in starts:
        m = sum(1 for j in range(tl) if differs(s, j))
        if m < best_m:
            best_m, best_s = m, s
    if best_s is None or best_m > 2 or om - best_m < 2:
        return 0
    return abs(best_s - orig)


def find_tip_deletions(ss, bases, ref, min_index, max_imperfect, look_right, look_left, search_range, slow_rescue_padding):
    if ss.slowScore >= max_imperfect or len(bases) <= 2 * TIPLEN:
        return False
    room = lambda: ALIGN_COLUMNS - (slow_rescue_padding + 8 + max(len(bases), ss.stop - ss.start))
    max_search = min(search_range, room())
    if max_search < 1:
        return False
    changed = False
    if look_right:
        x = _tip(bases, ref, min_index, ss.stop, max_search, TIPLEN, True)
        if x > 0:
            ss.stop += x
            changed = True
            max_search = min(max_search, room())
            if max_search < 1:
                return changed
    if look_left:
        y = _tip(bases, ref, min_index, ss.start, max_search, TIPLEN, False)
        if y > 0:
            ss.start -= y
            changed = True
    return changed


class Rescuer:
    def __init__(self, oracle, ref8, cfg, search_range=100, slow_rescue_padding=8, clearzone1e=258, min_index=0):
        self.oracle, self.ref8, self.ref, self.cfg = oracle, ref8, ref8.tolist(), cfg
        self.packed = oracle.new_packed(601, 3000)
        self.search_range, self.pad, self.cz1e, self.min_index = search_range, slow_rescue_padding, clearzone1e, min_index
        self.scans = self.fills = 0

    def in_bounds(self, ss):
        return ss.start >= 0 and ss.stop <= len(self.ref) - 1

    def slow_rescue(self, bases8, ss, max_score, max_imperfect, tip_right, tip_left):
        bases = bases8.tolist()
        L = len(bases)
        no_indel = pygapped.score_no_indels(bases, self.ref, ss.start)
        old_start = ss.start
        if no_indel < max_imperfect and int(self.cfg["max_indel"]) > 0:
            ss.set_slow_score(no_indel)
            if tip_right or tip_left:
                if find_tip_deletions(ss, bases, self.ref, self.min_index, max_imperfect, tip_right, tip_left, self.search_range, self.pad):
                    no_indel = pygapped.score_no_indels(bases, self.ref, ss.start)
            min_msa_limit = -self.cz1e + int(F(self.cfg["min_ratio_paired"]) * F(max_score))
            minscore = max(no_indel, min_msa_limit)
            self.fills += 1
            sv, _, _ = pygapped.fill_and_score_limited(self.oracle, self.packed, 601, 3000, bases8, self.ref8, ss.start - self.pad, ss.stop + self.pad, minscore, None)
            if sv is not None:
                ss.set_slow_score(sv[0])
                ss.score = ss.slowScore
                ss.start, ss.stop = sv[1], sv[2]
            else:
                ss.set_slow_score(no_indel)
                ss.score = ss.slowScore
                ss.start = old_start
                ss.stop = ss.start + L - 1
        else:
            ss.set_slow_score(no_indel)
            ss.score = ss.slowScore
            ss.stop = ss.start + L - 1
        ss.pairedScore = ss.score + 1
        ss.perfect = ss.slowScore == max_score
        if ss.perfect:
            ss.semiperfect = True
        else:
            pyclip.set_perfect(ss, bases, self.ref)

    def rescue(self, anchor_sites, anchor_len, loose_sites, basesP8, basesM8, qual_loose, search_dist):
        cfg = self.cfg
        if search_dist > int(cfg["max_rescue_dist"]) or not anchor_sites:
            return
        L = len(basesP8)
        max_loose = 70 + (L - 1) * 100
        max_anchor = 70 + (anchor_len - 1) * 100
        max_imperfect = max_loose + min(-472, -395 - 100)
        best_loose = loose_sites[0].slowScore if loose_sites else 0
        best_anchor = anchor_sites[0].slowScore
        if best_loose == max_loose and best_anchor == max_anchor and anchor_sites[0].pairedScore > 0:
            return
        rescue_limit = int(F(0.95) * F(best_anchor))
        retain = max(int(F(0.68) * F(best_loose)), int(F(0.4) * F(max_loose)))
        retain2 = max(int(F(0.95) * F(best_loose)), int(F(0.55) * F(max_loose)))
        max_mm = 5 if best_loose > max_imperfect else min(int(cfg["max_rescue_mismatches"]), int(F(0.60) * F(L) - F(1)))
        find_tip = self.search_range > 0 and best_loose < max_imperfect
        if qual_loose is None:
            right = left = find_tip
        else:
            q = [int(x) for x in qual_loose]
            n = TIPLEN
            if n > len(q):                                       # the four Read helpers return 0 when n exceeds the read
                min_last = avg_last = min_first = avg_first = 0
            else:
                tail, head = q[len(q) - n:], q[:n]
                min_last, avg_last = min(tail), sum(max(x, 0) for x in tail) // n
                min_first, avg_first = min(head), sum(max(x, 0) for x in head) // n
            right = find_tip and min_last >= TIP_MIN_Q and avg_last >= TIP_AVG_Q
            left = find_tip and min_first >= TIP_MIN_Q and avg_first >= TIP_AVG_Q
        same = bool(cfg["same_strand_pairs"])
        apd = int(cfg["average_pair_dist"])
        for ssa in anchor_sites:
            if ssa.slowScore < rescue_limit:
                break
            if ssa.pairedScore == 0 and not ssa.rescued:
                into = ssa.stop - ssa.start - 1 + (anchor_len * 11 // 16)
                strand = ssa.strand if same else ssa.strand ^ 1
                search_right = (strand == 0) if same else (strand == 1)
                anchor_left_of_mate = (ssa.strand == 0)                # plus-strand anchor: the mate lies to the right (either pairing mode)
                bases8 = (basesM8 if ssa.strand == 1 else basesP8) if same else (basesM8 if ssa.strand == 0 else basesP8)
                if anchor_left_of_mate:
                    loc, ideal = ssa.stop - into, ssa.stop + apd
                else:
                    loc, ideal = ssa.start + into, ssa.start - apd
                self.scans += 1
                ss = quick_rescue(bases8, self.ref8, self.min_index, ssa.chrom, strand, loc, search_dist + into, search_right, ideal, max_mm)
                if ss is not None and self.in_bounds(ss):
                    mm = ss.slowScore
                    ss.set_slow_score(0)
                    if mm <= max_mm:
                        self.slow_rescue(bases8, ss, max_loose, max_imperfect, right, left)
                        if ss.score > retain and self.in_bounds(ss):
                            if ss.score > retain2:
                                ss.pairedScore = max(ss.pairedScore, ss.slowScore + _jdiv(ssa.slowScore, 4))
                                ssa.pairedScore = max(ssa.pairedScore, ssa.slowScore + _jdiv(ss.slowScore, 4))
                            loose_sites.append(ss)


def _jdiv(a, b):
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b > 0) else -q
