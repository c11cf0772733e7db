"""TEST INFRASTRUCTURE — an independent restatement (Python, from the Java text alone) of the SiteScore helpers realign_new / genMatchString lean on:
  leftPaddingNeeded / rightPaddingNeeded     current/stream/SiteScore.java:447-491 (the right-hand loop's `mloc>=tiplen` test included, as written)
  clipTipIndels / clipLeftTipIndel / clipRightTipIndel / unclip     :493-672
  fixXY                                      :674-826
  setPerfect                                 :239-292
  ChromosomeArray.get                        current/dna/ChromosomeArray.java:232-234 (N at and beyond maxIndex)
  Read.calcMatchLength                       current/stream/Read.java:1419-1470 (long format: every symbol but I advances the reference)
(incrementStart / incrementStop re-fix a gap array the way setStart / setStop do).  The match string is a Python list of characters; MSA.score(match) is tests/pygapped.score_match, setSlowScore the one of
tests/pysitelist.Site.  Shares no code with oracle/mapper_oracle.c."""
import pygapped
import pysitelist as _ps

N = ord("N")
DEFINED = {ord(c) for c in "ACGT"}


class ClipSite:
    def __init__(self, site, match):
        """site: a pysitelist.Site; match: bytes or None."""
        self.s = site
        self.match = None if match is None else list(match)

    # -- small accessors --
    def mapped_length(self):
        return self.s.stop - self.s.start + 1

    def match_length(self):
        return sum(1 for c in self.match if c != ord("I"))

    def lengths_agree(self):
        return True if self.match is None else self.match_length() == self.mapped_length()

    def contains_xy(self):
        if not self.match:
            return False
        return chr(self.match[0]) in "XY" or chr(self.match[-1]) in "XY"

    # -- padding --
    def left_padding_needed(self, tiplen, max_indel):
        if not self.match:
            return 0
        ins = xy = 0
        for mloc, c in enumerate(self.match):
            ch = chr(c)
            if ch == "I":
                ins += 1
            elif ch in "XY":
                xy += 1
            elif ch == "D":
                return ins + xy
            elif mloc >= tiplen:
                break
        return ins + xy if (ins > max_indel or xy > 0 or self.match[0] == ord("I")) else 0

    def right_padding_needed(self, tiplen, max_indel):
        if not self.match:
            return 0
        ins = xy = 0
        for mloc in range(len(self.match) - 1, -1, -1):
            ch = chr(self.match[mloc])
            if ch == "I":
                ins += 1
            elif ch in "XY":
                xy += 1
            elif ch == "D":
                return ins + xy
            elif mloc >= tiplen:          # (sic)
                break
        return ins + xy if (ins > max_indel or xy > 0 or self.match[-1] == ord("I")) else 0

    # -- clipping --
    def _clip_left(self, tiplen, max_indel):
        m = self.match
        if m is None or len(m) < max_indel or chr(m[0]) in "CYX":
            return False
        neutral = ins = dele = 0
        mloc = 0
        while mloc < len(m):
            ch = chr(m[mloc])
            if ch == "I":
                ins += 1
            elif ch == "D":
                dele += 1
            else:
                neutral += 1
                if mloc >= tiplen:
                    break
            mloc += 1
        while mloc >= 0 and mloc < len(m) and m[mloc] == ord("m"):      # the Java reads match[mloc] with mloc == length when the loop ran off the end: it would throw
            mloc -= 1; neutral -= 1
        if ins <= max_indel and dele <= 4 * max_indel:
            return False
        total = neutral + ins + dele
        if dele > 0:
            m = [c for c in m[:total] if c != ord("D")] + m[total:]
        for i in range(neutral + ins):
            m[i] = ord("C")
        self.match = m
        _set_start(self.s, self.s.start - (ins - dele))
        return True

    def _clip_right(self, tiplen, max_indel):
        m = self.match
        if m is None or len(m) < max_indel or chr(m[-1]) in "CYX":
            return False
        last = len(m) - 1
        neutral = ins = dele = 0
        lowest = last - tiplen
        mloc = last
        while mloc >= 0:
            ch = chr(m[mloc])
            if ch == "I":
                ins += 1
            elif ch == "D":
                dele += 1
            else:
                neutral += 1
                if mloc <= lowest:
                    break
            mloc -= 1
        while 0 <= mloc < len(m) and m[mloc] == ord("m"):
            mloc += 1; neutral -= 1
        if ins <= max_indel and dele <= 4 * max_indel:
            return False
        total = neutral + ins + dele
        limit = len(m) - total
        if dele > 0:
            m = m[:limit] + [c for c in m[limit:] if c != ord("D")]
        for i in range(limit, len(m)):
            m[i] = ord("C")
        self.match = m
        _ps.set_stop(self.s, self.s.stop + (ins - dele))
        return True

    def unclip(self, bases, ca):
        m = self.match
        if not m or (m[0] != ord("C") and m[-1] != ord("C")):
            return False
        rloc, cloc = self.s.start, 0
        for i, sym in enumerate(m):
            ch = chr(sym)
            if ch == "C":
                c, r = bases[cloc], ca(rloc)
                m[i] = N if (c not in DEFINED or r not in DEFINED) else (ord("m") if c == r else ord("S"))
                rloc += 1; cloc += 1
            elif ch in "mNSXY":
                rloc += 1; cloc += 1
            elif ch == "I":
                cloc += 1
            elif ch == "D":
                rloc += 1
            else:
                raise RuntimeError("Unsupported symbol")
        return True

    def clip_tip_indels(self, bases, ref, tiplen, max_indel):
        if self.match is None or len(self.match) < max_indel:
            return False
        left = self._clip_left(tiplen, max_indel)
        right = self._clip_right(tiplen, max_indel)
        if left or right:
            self.unclip(bases, _getter(ref))
            self._rescore()
            set_perfect(self.s, bases, ref)
        return left or right

    def _rescore(self):
        old = self.s.slowScore
        self.s.set_slow_score(pygapped.score_match(bytes(self.match)))
        self.s.score = self.s.score + (self.s.slowScore - old)

    # -- fixXY(bases, nullifyOnFailure=false, msa) --
    def fix_xy(self, bases, ref):
        if not self.contains_xy():
            return True
        ca = _getter(ref)
        m = self.match
        success = True
        MAX_SUBS, MAX_RATE = 5, 0.4

        def rewrite(mloc, c, r):
            if r == N or c == N:
                m[mloc] = N
            elif c == r:
                m[mloc] = ord("m")
            else:
                m[mloc] = ord("S")
                return 1
            return 0

        lead = 0
        while lead < len(m) and chr(m[lead]) in "XY":
            lead += 1
        if lead >= len(m) or lead >= len(bases):
            success = False
        elif lead > 0:
            subs, first_sub = 0, -1
            for mloc in range(lead - 1, -1, -1):                 # from the last X/Y on the left back to position 0
                if rewrite(mloc, bases[mloc], ca(self.s.start + mloc)):
                    subs += 1
                    if subs == 1:
                        first_sub = mloc
            if self.mapped_length() != self.match_length():
                _set_start(self.s, self.s.start - lead)
            if subs > MAX_SUBS and subs > lead * MAX_RATE:
                for i in range(first_sub + 1):
                    m[i] = ord("C")
        if success:
            mloc = len(m) - 1
            while mloc >= 0 and chr(m[mloc]) in "XY":
                mloc -= 1
            dif = len(m) - 1 - mloc
            if mloc < 0:
                success = False
            elif dif > 0:
                first = mloc + 1
                num_x = len(m) - first
                rloc, cloc = self.s.stop - dif + 1, len(bases) - dif
                subs, first_sub = 0, -1
                if cloc < 0:
                    success = False
                else:
                    for k in range(first, len(m)):
                        if rewrite(k, bases[cloc], ca(rloc)):
                            subs += 1
                            if subs == 1:
                                first_sub = k
                        rloc += 1; cloc += 1
                if success:
                    if self.mapped_length() != self.match_length():
                        _ps.set_stop(self.s, self.s.stop + num_x)
                    if subs > MAX_SUBS and subs > num_x * MAX_RATE:
                        for i in range(first_sub, len(m)):
                            m[i] = ord("C")
        success = success and not self.contains_xy()
        if self.match is not None:
            self._rescore()
        set_perfect(self.s, bases, ref)
        return success


def _set_start(site, a):
    """SiteScore.setStart (stream/SiteScore.java:935-943)."""
    site.start = a
    if site.gaps is not None:
        site.gaps[0] = a
        if site.gaps[0] > site.gaps[1]:
            site.gaps = _ps.fix_gaps(site.start, site.stop, site.gaps)


def _getter(ref):
    """ChromosomeArray.get: N below minIndex (0) and at or beyond maxIndex (= array length - 1)."""
    mx = len(ref) - 1
    return lambda loc: N if (loc < 0 or loc >= mx) else ref[loc]


def set_perfect(s, bases, ref):
    """SiteScore.setPerfect(bases)."""
    if len(bases) != s.stop - s.start + 1:
        s.perfect = s.semiperfect = False
        return False
    perfect = semi = True
    refloc, readloc, n = s.start, 0, 0
    mx = min(s.stop, len(ref) - 1)
    nlimit = len(bases) // 2
    if s.start < 0:
        n -= s.start; readloc -= s.start; refloc -= s.start
        perfect = False
    if s.stop >= len(ref):
        n += s.stop - len(ref) + 1
        perfect = False
    if n > nlimit:
        s.perfect = s.semiperfect = False
        return False
    while refloc <= mx:
        c, r = bases[readloc], ref[refloc]
        if c != r or c == N:
            perfect = False
            if c == N:
                semi = False
            if r != N:
                s.perfect = s.semiperfect = False
                return False
            n += 1
            if n > nlimit:
                s.perfect = s.semiperfect = False
                return False
        refloc += 1; readloc += 1
    semi = semi and n <= nlimit
    perfect = perfect and semi and n == 0
    s.perfect, s.semiperfect = perfect, semi
    return perfect
