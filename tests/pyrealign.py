"""TEST INFRASTRUCTURE — an independent restatement (Python, from the Java text) of TranslateColorspaceRead.realign_new
(current/align2/TranslateColorspaceRead.java:229-660), sites with a gap array included (fills on makeGref's reference, the gapped-length form of the window rule, setLimits /
fixGaps): the fixXY / clipTipIndels preamble, the padding rules, scoreNoIndels with its
match string, up to three limited fills with growing windows (plus the unlimited one the plus-strand block alone has), traceback, setLimits / fixLimitsXY, the
left/right-padding test with one level of recursion, setPerfect.  The two strand blocks of the reference differ in two places and both are kept: the minus-strand
block's first window adjustment has no `else` branch, and it has no fillUnlimited fallback.  Alignments: tests/pygapped.py (fills by the reference's own C),
walks: tests/pywalk.py; helpers: tests/pyclip.py.  Shares no code with oracle/mapper_oracle.c."""
import numpy as np

import pyclip
import pygapped
import pysitelist as ps

MAX_COLUMNS = 3000


def _adjust(newlen, left, right, with_else):
    lim = MAX_COLUMNS - 80
    if newlen >= lim:
        while newlen >= lim and left > right:
            newlen -= 1; left -= 1
        while newlen >= lim and left < right:
            newlen -= 1; right -= 1
        while newlen >= lim:
            newlen -= 2; left -= 1; right -= 1
    elif with_else:
        x = max(0, min(20, (MAX_COLUMNS - newlen) // 2 - 40))
        left, right = max(x, left), max(x, right)
    return left, right


class Realigner:
    def __init__(self, oracle, ref8, maxR=601):
        self.oracle, self.ref8, self.ref = oracle, ref8, ref8.tolist()
        self.packed = oracle.new_packed(maxR, MAX_COLUMNS)
        self.maxR = maxR
        self.fills = 0

    def _fill(self, bases8, lo, hi, minscore, gaps=None):
        self.fills += 1
        return pygapped.fill_and_score_limited(self.oracle, self.packed, self.maxR, MAX_COLUMNS, bases8, self.ref8, lo, hi, minscore, gaps)

    def realign(self, cs, bases8, padding, recur, min_valid, forbid_indels=False, fix_xy=False):
        """cs: pyclip.ClipSite (edited in place); bases8: numpy int8 of the site's strand."""
        s = cs.s
        bases = bases8.tolist()
        L = len(bases)
        max_index = len(self.ref) - 1
        if cs.contains_xy():
            cs.fix_xy(bases, self.ref)
        cs.clip_tip_indels(bases, self.ref, 4, 10)
        padding = max(min(padding, (MAX_COLUMNS - L) // 2 - 20), 0)
        if ps.calc_gref_len(s.start, s.stop, s.gaps) > MAX_COLUMNS - 20:
            ps.set_stop(s, s.start + min(L + 40, MAX_COLUMNS - 20))
            if s.gaps is not None:
                s.gaps = ps.fix_gaps(s.start, s.stop, s.gaps)
        if s.start < 0:
            pyclip._set_start(s, 0)
        if s.stop > max_index:
            ps.set_stop(s, max_index)
        span = s.stop - s.start + 1
        if span < L:
            padding = max(padding, min(L, L - span + 10) // 2 + 1)
        padding = max(0, min(padding, (MAX_COLUMNS - max(L, ps.calc_gref_len(s.start, s.stop, s.gaps))) // 2 - 100))
        if forbid_indels:
            padding = 0
        max_q = 70 + (L - 1) * 100
        max_i = max_q + min(-472, -395 - 100)
        plus = s.strand == 0
        # scoreNoIndelsAndMakeMatchString writes into the old array when it has the read's length, else into a fresh (zeroed) one
        old = cs.match if (cs.match is not None and len(cs.match) == L) else None
        no_indel, m = pygapped.score_no_indels(bases, self.ref, s.start, True)
        if m is None:                                     # -99999: the site hangs over the array; the array passed in is left as it was
            cs.match = list(old) if old is not None else [0] * L
        else:
            lo, hi = (0 if s.start >= 0 else -s.start), L - max(0, s.start + L - len(self.ref))
            cs.match = [m[i] if lo <= i < hi else (old[i] if old is not None else 0) for i in range(L)]
        if no_indel >= max_i or forbid_indels:
            ps.set_stop(s, s.start + L - 1)
            s.set_slow_score(no_indel)
        else:
            lo, hi = max(s.start - padding, 0), min(s.stop + padding, max_index)
            lim = max(no_indel, min_valid)

            def newlen(a, b, l, r):                                        # the window the next fill would need: columns, or gapped-reference length
                if s.gaps is None:
                    return b - a + 1 + l + r
                return max(L, ps.calc_gref_len(a, b, s.gaps)) + 1 + l + r

            sv, ms, mx = self._fill(bases8, lo, hi, lim, s.gaps)
            if sv is not None and len(sv) > 6:
                old0 = sv[0]
                epl, epr = _adjust(newlen(lo, hi, sv[6], sv[7]), sv[6], sv[7], with_else=plus or s.gaps is not None)
                lo, hi = max(0, lo - epl), min(max_index, hi + epr)
                sv, ms, mx = self._fill(bases8, lo, hi, lim, s.gaps)
                if sv is None or sv[0] < old0:
                    epl, epr = _adjust(newlen(lo, hi, epl, epr), epl, epr, with_else=True)
                    lo, hi = max(0, lo - epl), min(max_index, hi + epr)
                    sv, ms, mx = self._fill(bases8, lo, hi, lim, s.gaps)
                    if plus and lo > 0 and hi < max_index and (sv is None or sv[0] < old0):
                        lo, hi = max(s.start - 8, 0), min(s.stop + 8, max_index)
                        sv, ms, mx = self._fill(bases8, lo, hi, 0, s.gaps)  # fillUnlimited
            if sv is not None:
                cs.match = list(ms)
                ps.set_limits(s, sv[1], sv[2])
                y = 0
                for c in reversed(cs.match):                               # fixLimitsXY: trailing Y symbols extend the stop
                    if c != ord("Y"):
                        break
                    y += 1
                if y:
                    ps.set_limits(s, s.start, s.stop + y)
                s.set_slow_score(sv[0])
            else:
                ps.set_stop(s, s.start + L - 1)
                s.set_slow_score(no_indel)
        lp, rp = cs.left_padding_needed(4, 5), cs.right_padding_needed(4, 5)
        if s.stop < max_index and s.start > 0 and (lp > 0 or rp > 0):
            if recur > 0:
                s.gaps = ps.fix_gaps(s.start, s.stop, s.gaps)
                self.realign(cs, bases8, min(10 + max(lp, rp), (MAX_COLUMNS - L) // 2 - 20), recur - 1, min_valid, forbid_indels, fix_xy)
            elif fix_xy and cs.contains_xy():
                cs.fix_xy(bases, self.ref)
        pyclip.set_perfect(s, bases, self.ref)
