"""TEST INFRASTRUCTURE — an independent restatement (plain Python objects, from the Java text alone) of the SAM record fields the reference derives
from a mapped Read: SamLine(Read, fragNum) (current/stream/SamLine.java:82-355) with makeFlag (:2134-2151), toMapq (:1703-1723), toCigar13/14
(:600-750), countLeadingClip / countLeadingIndels / countTrailingClip / countTrailingIndels (:924-1020), and the scaffold lookups
Data.scaffoldIndex / isSingleScaffold / scaffoldRelativeLoc (current/dna/Data.java:1089-1140).  Shares no code with oracle/sam_oracle.c.

Reads are objects that the constructor MUTATES the way the reference does (a multi-scaffold alignment unmaps the read, drops its match string and
clears `paired` on both mates), and records are built in file order (mate 1, then mate 2), so the second record of a pair sees those effects."""
import bisect
import math

import numpy as np

F = np.float32
INV_LOG2 = 1 / math.log(2)


class PyRead:
    def __init__(self, chrom, start, stop, length, score, match, mapped, minus, perfect, ambiguous, secondary, discarded, paired, pairnum):
        self.chrom, self.start, self.stop, self.length, self.score, self.match = chrom, start, stop, length, score, match
        self.mapped, self.minus, self.perfect, self.ambiguous = mapped, minus, perfect, ambiguous
        self.secondary, self.discarded, self.paired, self.pairnum = secondary, discarded, paired, pairnum
        self.mate = None


class Scaffolds:
    def __init__(self, table, padding=300):
        """table: [(chrom, loc, length)], chrom 1-based."""
        self.locs, self.lens, self.gidx = {}, {}, {}
        for g, (ch, loc, ln) in enumerate(sorted(table)):
            self.locs.setdefault(ch, []).append(loc); self.lens.setdefault(ch, []).append(ln); self.gidx.setdefault(ch, []).append(g)
        self.padding = padding

    def _search(self, array, key):
        """Arrays.binarySearch: index if present, else -(insertion point)-1."""
        i = bisect.bisect_left(array, key)
        return i if i < len(array) and array[i] == key else -1 - i

    def index(self, chrom, loc):
        array = self.locs[chrom]
        if len(array) < 2:
            return 0
        loc += self.padding // 2
        idx = self._search(array, loc)
        return idx if idx >= 0 else max(0, (-1 - idx) - 1)

    def is_single(self, chrom, loc1, loc2):
        array = self.locs[chrom]
        if len(array) < 2:
            return True
        idx = self._search(array, loc1 + self.padding)
        scaf = idx if idx >= 0 else max(0, (-1 - idx) - 1)
        if scaf == len(array) - 1:
            return True
        lower, upper = array[scaf] - self.padding, array[scaf + 1]
        if loc2 < lower or loc1 > upper:
            return False
        return loc2 < upper


def java_round(x):
    return int(math.floor(float(x) + 0.5))


def to_mapq(score, length, mapped, ambig, penalize_ambig=True):
    if not mapped or length < 1:
        return 0
    if ambig and penalize_ambig:
        adjusted = (F(score) * F(3)) / (F(100) * F(length))
        return max(1, java_round(adjusted))
    score2 = F(score - length * 40) * F(1.6)
    mx = F(1.5) * F(math.log(length) * INV_LOG2) + F(36)
    adjusted = (score2 * mx) / (F(100) * F(length))
    return max(4, java_round(adjusted))


def leading_clip(match):
    if not match or match[0] != ord("C"):
        return 0
    n = 0
    for b in match:
        if b != ord("C"):
            break
        n += 1
    return n


def trailing_clip(match):
    if match is None:
        return 0
    n = 0
    for b in reversed(match):
        if b != ord("C"):
            break
        n += 1
    return n


def leading_indels(rloc, match):
    if match is None or rloc >= 0:
        return 0
    dels = inss = 0
    for b in match:
        if rloc >= 0:
            break
        if b == ord("D"):
            dels += 1; rloc += 1
        elif b == ord("I"):
            inss += 1
        else:
            rloc += 1
    return dels - inss


def trailing_indels(rloc, rlen, match):
    """As written in the reference this returns 0 for every reachable input: it bails out for rloc >= 0, and for rloc < 0 the loop condition
    rloc >= rlen is false at once."""
    if match is None or rloc >= 0:
        return 0
    assert rloc < rlen
    return 0


def to_cigar(match, read_start, read_stop, reflen, v14, soft_clip, intron_limit):
    if match is None or read_start == read_stop:
        return None
    parts = []
    count = 0
    mode = last = "="
    refloc = read_start

    def flush(md, cnt):
        parts.append(str(cnt) + ("N" if md == "D" and cnt > intron_limit else md))

    for m in match:
        c = chr(m)
        sfd = False
        if soft_clip and (refloc < 0 or refloc >= reflen):
            mode = "S"
            if c != "I":
                refloc += 1
            if c == "D":
                sfd = True
        elif v14 and c in "ms":
            mode = "="; refloc += 1
        elif v14 and c == "S":
            mode = "X"; refloc += 1
        elif not v14 and c in "msSNB":
            mode = "M"; refloc += 1
        elif c in "IXY":
            mode = "I"
        elif c == "D":
            mode = "D"; refloc += 1
        elif c == "C":
            mode = "S"; refloc += 1
        elif v14 and c in "NB":
            mode = "M"; refloc += 1
        else:
            raise RuntimeError("Invalid match string character")
        if mode != last:
            if count > 0:
                flush(last, count)
            count = 0
            last = mode
        count += 1
        if sfd:
            count -= 1
    flush(mode, count)
    return "".join(parts)


class PySamLine:
    def __init__(self, r1, frag_num, S, v14=True, soft_clip=True, intron_limit=2 ** 31 - 1, penalize_ambig=True):
        r2 = r1.mate
        perfect = r1.perfect
        idx1 = idx2 = -1
        a1 = a2 = b1 = b2 = 0
        scaflen = scaflen2 = 0
        name1 = name2 = "*"
        if r1.mapped:
            if S.is_single(r1.chrom, r1.start, r1.stop):
                idx1 = S.index(r1.chrom, (r1.start + r1.stop) // 2)
                name1 = S.gidx[r1.chrom][idx1]
                scaflen = S.lens[r1.chrom][idx1]
                a1 = r1.start - S.locs[r1.chrom][idx1]
                b1 = a1 - r1.start + r1.stop
            else:
                r1.mapped = False; r1.paired = False; r1.match = None
                if r2 is not None:
                    r2.paired = False
        if r2 is not None and r2.mapped:
            if S.is_single(r2.chrom, r2.start, r2.stop):
                idx2 = S.index(r2.chrom, (r2.start + r2.stop) // 2)
                name2 = S.gidx[r2.chrom][idx2]
                scaflen2 = S.lens[r2.chrom][idx2]
                a2 = r2.start - S.locs[r2.chrom][idx2]
                b2 = a2 - r2.start + r2.stop
            else:
                r2.mapped = False; r2.paired = False; r2.match = None
                r1.paired = False
        same = r2 is not None and idx1 > -1 and idx1 == idx2 and r1.chrom == r2.chrom
        # makeFlag
        flag = 0
        if r2 is not None:
            flag |= 0x1
            if r1.mapped and r1.match is not None and same and r1.paired and r2.mapped and r2.match is not None:
                flag |= 0x2
            flag |= 0x40 if frag_num == 0 else 0x80
        if not r1.mapped:
            flag |= 0x4
        if r2 is not None and not r2.mapped:
            flag |= 0x8
        if r1.minus:
            flag |= 0x10
        if r2 is not None and r2.minus:
            flag |= 0x20
        if r1.secondary:
            flag |= 0x100
        if r1.discarded:
            flag |= 0x200
        self.flag = flag
        self.rname = name1 if r1.mapped else (name2 if (r2 is not None and r2.mapped) else None)
        if r1.mapped:
            pos0 = (a1 + 1) + leading_clip(r1.match) + leading_indels(a1, r1.match)
            pos1 = (b1 + 1) - trailing_clip(r1.match) - trailing_indels(b1, scaflen, r1.match)
            pos1 = min(pos1, scaflen)
            pos0 = max(pos0, 1)
        else:
            pos0 = pos1 = 0
        if r2 is not None and r2.mapped:
            pos0m = (a2 + 1) + leading_clip(r2.match) + leading_indels(a2, r2.match)
            pos1m = (b2 + 1) - trailing_clip(r2.match) - trailing_indels(b2, scaflen, r2.match)
            if pos1m > scaflen:
                pos1 = scaflen                         # (sic) the reference clamps pos1, against the FIRST read's scaffold length
            pos0m = max(pos0m, 1)
        else:
            pos0m = pos1m = 0
        tlen = 0
        if r2 is None:
            pos, pnext = pos0, pos0m
        elif r1.mapped and r2.mapped:
            pos, pnext = pos0, pos0m
            if same:
                tlen = 1 + (max(pos1, pos1m) - min(pos0, pos0m))
        elif r1.mapped:
            pos, pnext = pos0, pos0
        elif r2.mapped:
            pos, pnext = pos0m, pos0m
        else:
            pos, pnext = pos0, pos0m
        self.pos, self.pnext = pos, pnext
        self.mapq = to_mapq(r1.score, r1.length, r1.mapped, r1.ambiguous, penalize_ambig)
        inbounds = r1.mapped and a1 >= 0 and b1 < scaflen
        self.cigar = None
        if r1.mapped and r1.match is not None:
            non_m = any(b > ord("9") and b != ord("m") for b in r1.match)
            non_nms = any(b > ord("9") and chr(b) not in "msNS" for b in r1.match)
            if v14:
                self.cigar = "%d=" % r1.length if (inbounds and perfect and not non_m) else to_cigar(r1.match, a1, b1, scaflen, True, soft_clip, intron_limit)
            else:
                self.cigar = "%dM" % r1.length if (inbounds and (perfect or not non_nms)) else to_cigar(r1.match, a1, b1, scaflen, False, soft_clip, intron_limit)
        if r2 is None or (not r1.mapped and not r2.mapped):
            self.rnext = "*"
        elif r1.mapped and r2.mapped:
            self.rnext = "=" if same else name2
        else:
            self.rnext = "="
        if not (r2 is None or r1.start < r2.start or (r1.start == r2.start and r1.pairnum == 0)):
            tlen = -tlen
        self.tlen = tlen
