"""TEST INFRASTRUCTURE — independent restatements (plain Python, from the Java text alone) of
  * MSA.scoreNoIndels / scoreNoIndelsAndMakeMatchString      current/align2/MultiStateAligner11tsJNI.java:1033-1089,1243-1318   
  * makeGref, translateTo/FromGappedCoordinate               :668-801                                                            
  * the wrapper around the native fill: MSA.fillAndScoreLimited (MSA.java:103-134), fillLimited / fillLimitedX dispatch and the -120
    (:115-164), score(…, gapped) and traceback(…, gapped) (:362-375, 499-531)
The fills themselves are done by the reference's own C (oracle/_ref); the walks by tests/pywalk.py.  Shares no code with oracle/msa_oracle.c."""
import numpy as np

import pywalk

GAPBUFFER = 64
GAPBUFFER2 = 128
GAPLEN = 128
GAPC = ord("-")
GREFLIMIT2_CUSHION = 128
N = ord("N")
SUB_ARRAY = None


def _sub(t):
    """POINTS_SUB_ARRAY[t] (:1610-1625): -127 for the first substitution of a run, -51 up to LIMIT_FOR_COST_3 = 5, then -25."""
    return -127 if t <= 1 else (-51 if t <= 5 else -25)


def score_no_indels(read, ref, ref_start, want_match=False):
    """read, ref: sequences of signed byte values.  -> score, or (score, match bytearray) with want_match (match positions outside the
    scored range stay 0, as in the freshly allocated Java array)."""
    n = len(read)
    ref_stop = ref_start + n
    if want_match and (ref_start < 0 or ref_stop > len(ref)):
        return -99999, None
    lo, hi = 0, n
    if ref_start < 0:
        lo = -ref_start
    if ref_stop > len(ref):
        hi -= ref_stop - len(ref)
    score = 0
    mode = None
    streak = 0
    match = bytearray(n)
    for i in range(lo, hi):
        c, r = read[i], ref[ref_start + i]
        if c == r and c != N:
            if mode == "MS":
                streak += 1; score += 100
            else:
                streak = 0; score += 70
            mode = "MS"; match[i] = ord("m")
        elif c < 0 or c == N:
            match[i] = N
        elif r < 0 or r == N:
            match[i] = N
        else:
            streak = streak + 1 if mode == "SUB" else 0
            score += _sub(streak + 1)
            mode = "SUB"; match[i] = ord("S")
    return (score, match) if want_match else score


class Gref:
    def __init__(self, ref, gaps, start, stop, buflen=3002):
        """makeGref(ref, gaps, refStartLoc, refEndLoc): exons copied, every intron shortened to GAPBUFFER + (gap % GAPLEN) bases, `div` gap
        symbols, GAPBUFFER bases; then up to GREFLIMIT2_CUSHION bases of cushion."""
        g = list(gaps)
        g[0] = min(g[0], start)
        g[-1] = max(g[-1], stop)
        self.origin = g[0]
        out = []
        for i in range(0, len(g), 2):
            x, y = g[i], g[i + 1]
            out.extend(ref[x:y + 1])
            if i + 2 < len(g):
                z = g[i + 2]
                gap = z - y - 1
                assert gap >= GAPBUFFER2 + GAPLEN
                rem = gap % GAPLEN
                div = (gap - GAPBUFFER2) // GAPLEN
                out.extend(ref[y + 1: y + GAPBUFFER + rem + 1])
                out.extend([GAPC] * div)
                out.extend(ref[z - GAPBUFFER: z])
        self.limit = len(out)
        lim = min(buflen, self.limit + GREFLIMIT2_CUSHION)
        r = stop + 1
        self.limit2 = 0
        for i in range(self.limit, lim):
            out.append(ref[r] if r < len(ref) else N)
            self.limit2 = i
            r += 1
        self.bytes = out

    def to_gapped(self, point):
        if point <= self.origin:
            return point - self.origin
        j = self.origin
        for i in range(self.limit2):
            if j == point:
                return i
            j += GAPLEN if self.bytes[i] == GAPC else 1
        raise RuntimeError("Out of bounds.")

    def from_gapped(self, point):
        if point <= 0:
            return self.origin + point
        j = self.origin
        for i in range(self.limit2):
            if i == point:
                return j
            j += GAPLEN if self.bytes[i] == GAPC else 1
        raise RuntimeError("Out of bounds.")


def takes_unlimited(rows, columns, min_score, bandwidth=0, ratio=0.0):
    """The dispatch at the top of fillLimitedX (:132-141)."""
    if bandwidth < 1 and ratio <= 0:
        halfband = 0
    else:
        halfband = max(min(9999999 if bandwidth < 1 else bandwidth, 9999999 if ratio <= 0 else 8 + int(np.float32(rows) * np.float32(ratio))), columns - rows + 8) // 2
    return min_score < 1 or columns + rows < 90 or ((halfband < 1 or halfband * 3 > columns) and columns > rows + min(170, rows + 20))


def fill_and_score_limited(oracle, packed, maxRows, maxColumns, read, ref, ref_start, ref_end, min_score, gaps):
    """MSA.fillAndScoreLimited + traceback as BBMapThread.scoreSlow drives them.  read/ref: numpy int8 arrays.
    -> (score list or None, match bytes or None, [rows, maxCol, maxState, maxScore])."""
    a = max(0, ref_start)
    b = min(len(ref) - 1, ref_end)
    rd = [int(x) for x in read]
    if gaps is None or len(gaps) == 0:
        gref = None
        fill_ref, fa, fb = ref, a, b
    else:
        gref = Gref(ref.tolist(), [int(x) for x in gaps], a, b)
        fill_ref = np.array(gref.bytes, np.int8)
        fa, fb = 0, gref.limit
    rows, columns = len(rd), fb - fa + 1
    if takes_unlimited(rows, columns, min_score):
        res, _ = oracle.fill_unlimited(read, fill_ref, fa, fb, packed, maxRows, maxColumns, kind="reference")
        max4 = [int(x) for x in res[:4]]
    else:
        res, _ = oracle.fill_limited(read, fill_ref, fa, fb, min_score - 120, packed, maxRows, maxColumns, kind="reference")
        if res[4] == 1:
            return None, None, None
        max4 = [int(x) for x in res[:4]]
    M = pywalk.Matrix(memoryview(packed).cast("B").cast("i"), maxRows, maxColumns)
    if gref is None:
        sv = pywalk.score2(M, rows, columns, a, b, max4[0], max4[1], max4[2])
        ms = pywalk.traceback2(M, bytes(read.view(np.uint8)), bytes(ref.view(np.uint8)), columns, a, max4[0], max4[1], max4[2])
    else:
        gstart, gstop = gref.to_gapped(a), gref.to_gapped(b)
        sv = pywalk.score2(M, rows, columns, gstart, gstop, max4[0], max4[1], max4[2])
        sv[1] = gref.from_gapped(sv[1])
        sv[2] = gref.from_gapped(sv[2])
        ms = pywalk.traceback2(M, bytes(read.view(np.uint8)), bytes(fill_ref.view(np.uint8)), columns, gstart, max4[0], max4[1], max4[2])
    return sv, ms, max4


# ---------------- MSA.score(match) (current/align2/MSA.java:488-557, calc*Score :726-747, MultiStateAligner11tsJNI.java:1347-1425) ----------------
def _del_score(n):
    score = -472
    if n > 256:                                   # approximateGaps: a run longer than MINGAP is costed like the gapped reference would be
        rem, div = n % 128, (n - 128) // 128
        score += div * -2                          # POINTS_GAP = -GAPCOST = -max(1, GAPLEN/64)
        n = rem + 128
    if n > 80:
        score += ((n - 80 + 3) // 4) * -1
        n = 80
    if n > 20:
        score += (n - 20) * -1
        n = 20
    if n > 5:
        score += (n - 5) * -9
        n = 5
    if n > 1:
        score += (n - 1) * -33
    return score


def _ins_score(n):
    total = 0
    for i in range(1, n + 1):                      # POINTS_INS_ARRAY_C: cumulative, clamped at MIN_SCORE
        total = max(-1046575, total + (-8 if i > 20 else (-23 if i > 5 else (-39 if i > 1 else -395))))
    return total


def _sub_score(n):
    score = -127
    if n > 5:
        score += (n - 5) * -25
        n = 5
    if n > 1:
        score += (n - 1) * -51
    return score


def score_match(match):
    """match: bytes in long format.  Runs come from itertools.groupby instead of the reference's mode / current bookkeeping."""
    import itertools
    score = 0
    prev_mode, prev_len = "0", 0
    for sym, grp in itertools.groupby(match):
        c, n = chr(sym), len(list(grp))
        if c == "m":
            score += 70 + (n - 1) * 100
        elif c == "S":
            score += _sub_score(n)
            if prev_mode in "NR":
                score += -51 - -127
            elif prev_mode == "m" and prev_len < 2:
                score += -147 - -127
        elif c == "D":
            score += _del_score(n)
        elif c in "IXY":
            score += _ins_score(n)
        elif c in "CNR":
            pass                                   # POINTS_NOCALL = POINTS_NOREF = 0
        else:
            raise AssertionError("Unhandled symbol " + c)
        prev_mode, prev_len = c, n
    return score
