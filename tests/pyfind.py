"""TEST INFRASTRUCTURE ONLY — a second, structurally independent statement of BBIndex.find in plain Python.

oracle/search_oracle.c and bbmap_b200/csrc/search.cu were written by the same hand in the same shape (arg-min over the live columns instead of
a heap, skip-ahead shortcuts in the kernel, flat arrays), so a slip in reading `slowWalk3` / `extendScore` made once would pass on both sides.
This file is written from the Java text alone, object by object, in the reference's own order of operations:

  * a real binary `QuadHeap` of `Quad` objects with the reference's percolation code          (align2/QuadHeap.java:15-75, align2/Quad.java)
  * `find` → key filtering with the five relaxations, `shrink2`, greedy trimming             (align2/BBIndex.java:403-535, 266-350; align2/Solver.java:48-152)
  * `prescanAllBlocks` → `findMaxQscore2`                                                    (BBIndex.java:642-741, 2294-2450)
  * `find(block, strand)` → `slowWalk3` with one heap pop at a time, no shortcuts             (BBIndex.java:742-782, 1219-1706)
  * `quickScore`, `scoreLeft/Right`, `scoreY`, `scoreZ2`, `maxQuickScore`, `extendScore`,
    `makeGapArray`, `calcApproxHitsCutoff`                                                   (BBIndex.java:2482-2511, 2558-2914, 2967-3035, 3267-3294; AbstractIndex.java:52-80)
  * `MSA.calcAffineScore`, `SiteScore.setPerfect`                                            (MultiStateAligner11tsJNI.java:871-941; stream/SiteScore.java:239-292)

Java `float` arithmetic is numpy float32 with one rounding per operation; `(int)` is truncation.  Inputs are what the seeding stage hands over
(offsets, keyScores, baseScores) and the index arrays (starts/sites per block, COUNTS, lengthHistogram, the BBIndex statics in the cfg record);
tests/test_find_independent.py asserts that every emitted SiteScore and `bestScores[]` equal the C restatement's on phiX, multi-block and
planted-repeat genomes."""
import numpy as np

F = np.float32

BASE_HIT_SCORE = 100
MAX_INDEL = 16000
MAX_INDEL2 = 2 * MAX_INDEL
Z_SCORE_MULT = 20
Y_SCORE_MULT = 10
HIT_FRACTION_TO_RETAIN = F(0.85)
MIN_HIT_LISTS_TO_RETAIN = 6
SMALL_GENOME_LIST = 20
MIN_APPROX_HITS_TO_KEEP = 1
MAX_HITS_REDUCTION_PERFECT = 0
MIN_SCORE_MULT = F(0.15)                  # USE_AFFINE_SCORE
MIN_QSCORE_MULT = F(0.025)
MIN_QSCORE_MULT2 = F(0.1)
DYNAMIC_SCORE_THRESH = F(0.84)
DYNAMIC_QSCORE_THRESH = F(0.6)
DYNAMIC_QSCORE_THRESH_PERFECT = F(0.8)
PRESCAN_QSCORE_THRESH = F(DYNAMIC_QSCORE_THRESH * F(0.95))
LIMIT_SUBSUMPTION_LENGTH_TO_2X = True
GAPBUFFER2, GAPLEN = 128, 128
MINGAP = GAPBUFFER2 + GAPLEN
POINTS_GAP = -max(1, GAPLEN // 64)

# MultiStateAligner11tsJNI.java:1489-1516, 1565-1610
POINTS_NOCALL, POINTS_MATCH, POINTS_MATCH2, POINTS_SUB, POINTS_SUB2, POINTS_SUB3 = 0, 70, 100, -127, -51, -25
POINTS_INS, POINTS_INS2, POINTS_INS3, POINTS_INS4 = -395, -39, -23, -8
POINTS_DEL, POINTS_DEL2, POINTS_DEL3, POINTS_DEL4, POINTS_DEL5 = -472, -33, -9, -1, -1
TIMESLIP = 4
MASK5 = TIMESLIP - 1
LIMIT_FOR_COST_3, LIMIT_FOR_COST_4, LIMIT_FOR_COST_5 = 5, 20, 80
MIN_SCORE = -(((1 << 20) - 1) - 2000)


def _tables():
    ins_c = [0] * 604; sub = [0] * 604
    for i in range(1, 604):
        pts = POINTS_INS4 if i > LIMIT_FOR_COST_4 else POINTS_INS3 if i > LIMIT_FOR_COST_3 else POINTS_INS2 if i > 1 else POINTS_INS
        ins_c[i] = max(MIN_SCORE, pts + ins_c[i - 1])
        sub[i] = POINTS_SUB3 if i > LIMIT_FOR_COST_3 else POINTS_SUB2 if i > 1 else POINTS_SUB
    return ins_c, sub


POINTS_INS_ARRAY_C, POINTS_SUB_ARRAY = _tables()

# Solver.java:224-235
POINTS_PER_LIST, POINTS_PER_BASE1, BONUS_POINTS_FOR_END_LIST, POINTS_FOR_TOTAL_LIST_WIDTH, MULT_FOR_SPACING_PENALTY = 30000, 6000, 40000, 5500, -30
EARLY_TERMINATION_SCORE = -50 * 2000       # static initialiser: evaluated while POINTS_PER_SITE still holds its default -50

BASE_TO_NUMBER = {ord(c): i for i, c in enumerate("ACGT")}
BASE_TO_NUMBER.update({ord(c): i for i, c in enumerate("acgt")})
BASE_TO_NUMBER[ord("U")] = 3; BASE_TO_NUMBER[ord("u")] = 3
COMPLEMENT = {ord(a): ord(b) for a, b in zip("ACGTUNacgtun", "TGCAANtgcaan")}


def jint(x):
    """(int) of a Java float: truncation toward zero."""
    return int(x)


def absdif(a, b):
    return a - b if a > b else b - a


def overlap(a1, b1, a2, b2):
    return a2 <= b1 and b2 >= a1


class Quad:
    __slots__ = ("column", "row", "site", "list")

    def __init__(self, col):
        self.column, self.row, self.site, self.list = col, 0, 0, None

    def compare_to(self, o):
        x = self.site - o.site
        return self.column - o.column if x == 0 else x


class QuadHeap:
    """align2/QuadHeap.java — 1-based array heap; `percDown` sifts a new leaf toward the root, `percUp` sifts the root toward the leaves."""

    def __init__(self, max_size):
        n = max_size + 1
        if n & 1:
            n += 1
        self.array = [None] * n
        self.size = 0

    def clear(self):
        self.size = 0

    def is_empty(self):
        return self.size == 0

    def peek(self):
        return None if self.size == 0 else self.array[1]

    def add(self, t):
        self.size += 1
        self.array[self.size] = t
        self._perc_down(self.size)

    def poll(self):
        if self.size == 0:
            return None
        t = self.array[1]
        self.array[1] = self.array[self.size]
        self.array[self.size] = None
        self.size -= 1
        if self.size > 0:
            self._perc_up(1)
        return t

    def _perc_down(self, loc):
        if loc == 1:
            return
        nxt = loc // 2
        a = self.array[loc]; b = self.array[nxt]
        while loc > 1 and a.compare_to(b) < 0:
            self.array[loc] = b
            loc = nxt
            nxt //= 2
            b = self.array[nxt]
        self.array[loc] = a

    def _perc_up(self, loc):
        n1 = loc * 2; n2 = n1 + 1
        if n1 > self.size:
            return
        a = self.array[loc]; b = self.array[n1]; c = self.array[n2] if n2 < len(self.array) else None
        if n2 > self.size:
            c = None                       # slots past `size` are nulled by poll()
        if c is None or b.compare_to(c) < 1:
            if a.compare_to(b) > 0:
                self.array[n1] = a; self.array[loc] = b
                self._perc_up(n1)
        else:
            if a.compare_to(c) > 0:
                self.array[n2] = a; self.array[loc] = c
                self._perc_up(n2)


class SiteScore:
    def __init__(self, chrom, strand, start, stop, hits, score, perfect):
        self.chrom, self.strand, self.start, self.stop, self.hits, self.score = chrom, strand, start, stop, hits, score
        self.perfect = perfect
        self.semiperfect = perfect
        self.gaps = None

    def overlaps(self, o):
        return self.chrom == o.chrom and self.strand == o.strand and overlap(self.start, self.stop, o.start, o.stop)


class GapFixNeeded(Exception):
    """a subsumption moved the limits of a site that carries a gap array (GapTools.fixGaps): the C restatement flags the read instead"""


class Block:
    def __init__(self, starts, sites):
        self.starts, self.sites = starts, sites

    def length_key(self, key):                 # Block.length(int key)  (Block.java:62-66)
        x = int(self.starts[key + 1]) - int(self.starts[key])
        if x == 0:
            return 0
        return x if int(self.sites[int(self.starts[key])]) != -1 else 0

    def length(self, start, stop):             # Block.length(int start, int stop)
        if start == stop or int(self.sites[start]) == -1:
            return 0
        return stop - start


class BBIndexPy:
    def __init__(self, cfg, blocks, counts, hist, chrom_bytes, chrom_off, quit_after_two_perfects=True):
        c = cfg[0] if hasattr(cfg, "dtype") and cfg.shape else cfg
        self.KEYLEN = int(c["keylen"])
        self.NUM_CHROM_BITS = int(c["chrombits"])
        self.CHROMS_PER_BLOCK = 1 << self.NUM_CHROM_BITS
        self.SHIFT_LENGTH = 32 - 1 - self.NUM_CHROM_BITS
        self.SITE_MASK = 0xFFFFFFFF >> (self.NUM_CHROM_BITS + 1)
        self.CHROM_MASK_LOW = self.CHROMS_PER_BLOCK - 1
        self.CHROM_MASK_HIGH = ~self.CHROM_MASK_LOW
        self.MAX_HITS_REDUCTION2 = int(c["max_hits_reduction2"]); self.MAXIMUM_MAX_HITS_REDUCTION = int(c["maximum_max_hits_reduction"])
        self.HIT_REDUCTION_DIV = int(c["hit_reduction_div"]); self.POINTS_PER_SITE = int(c["points_per_site"])
        self.MAX_AVERAGE_LIST_TO_SEARCH = int(c["max_average_list_to_search"]); self.MAX_AVERAGE_LIST_TO_SEARCH2 = int(c["max_average_list_to_search2"])
        self.MAX_SHORTEST_LIST_TO_SEARCH = int(c["max_shortest_list_to_search"]); self.MAX_USABLE_LENGTH = int(c["max_usable_length"])
        self.BASE_KEY_HIT_SCORE = BASE_HIT_SCORE * self.KEYLEN
        self.INV_BASE_KEY_HIT_SCORE = F(1) / F(self.BASE_KEY_HIT_SCORE)
        self.INDEL_PENALTY = self.BASE_KEY_HIT_SCORE // 2 - 1
        self.INDEL_PENALTY_MULT = 20
        self.MAX_PENALTY_FOR_MISALIGNED_HIT = self.BASE_KEY_HIT_SCORE - (1 + self.BASE_KEY_HIT_SCORE // 8)
        self.SCOREZ_1KEY = Z_SCORE_MULT * self.KEYLEN
        self.blocks = [Block(s, t) for s, t in blocks]
        self.COUNTS = counts
        self.lengthHistogram = hist
        self.chroms = [None] + [np.asarray(chrom_bytes[chrom_off[i]:chrom_off[i + 1]]).view(np.uint8).tobytes() for i in range(len(chrom_off) - 1)]
        self.minChrom, self.maxChrom = 1, len(chrom_off) - 1
        self.QUIT_AFTER_TWO_PERFECTS = quit_after_two_perfects
        self.heap = QuadHeap(255)
        self.status_gapfix = False

    # ---- codecs (BBIndex.java:3038-3060) ----
    def to_number(self, site, chrom):
        return ((chrom & self.CHROM_MASK_LOW) << self.SHIFT_LENGTH) | site

    def number_to_chrom(self, number, base_chrom):
        return (number >> self.SHIFT_LENGTH) + (base_chrom & self.CHROM_MASK_HIGH)

    def number_to_site(self, number):
        return number & self.SITE_MASK

    def base_chrom(self, chrom):
        return max(0, chrom & self.CHROM_MASK_HIGH)

    def block_of(self, chrom):
        """index[chrom]: every chromosome of a block points at the same Block object; blocks[0] holds chromosome 1."""
        return self.blocks[((chrom & self.CHROM_MASK_HIGH) - (1 & self.CHROM_MASK_HIGH)) // self.CHROMS_PER_BLOCK]

    def count(self, key):
        return int(self.COUNTS[key])

    # ---- key filtering ----
    def count_hits(self, keys, max_len):
        n = 0
        for i, key in enumerate(keys):
            if key >= 0:
                ln = self.count(key)
                if 0 < ln < max_len:
                    n += 1
                else:
                    keys[i] = -1
        return n

    @staticmethod
    def shrink2(offsets, keys, key_scores):
        keep = [i for i, k in enumerate(keys) if k >= 0]
        return [offsets[i] for i in keep], [keys[i] for i in keep], [key_scores[i] for i in keep]

    def value_of_element(self, offsets, lengths, key_weight, chunk, lists, index):
        numlists = len(lists)
        if numlists < 1:
            return 0
        prospect = lists[index]
        if lengths[prospect] == 0:
            return -999999
        valuep = POINTS_PER_LIST + (POINTS_PER_LIST * 2 // numlists) + ((POINTS_PER_LIST * 10) // lengths[prospect])
        valuem = self.POINTS_PER_SITE * lengths[prospect]
        if prospect == 0 or prospect == len(offsets) - 1:
            valuep += BONUS_POINTS_FOR_END_LIST
        if numlists == 1:
            valuep += (POINTS_FOR_TOTAL_LIST_WIDTH + POINTS_PER_BASE1) * chunk
            return int(F(valuep) * key_weight) + valuem
        first, last = lists[0], lists[-1]
        offL = -1 if prospect == first else offsets[lists[index - 1]]
        offP = offsets[prospect]
        offR = offsets[-1] + 1 if prospect == last else offsets[lists[index + 1]]
        old_left, old_right, new_space = offP - offL, offR - offP, offR - offL
        valuep += ((old_left * old_left + old_right * old_right) - (new_space * new_space)) * MULT_FOR_SPACING_PENALTY
        if prospect == first:
            uniquely = offR - offP
        elif prospect == last:
            uniquely = offP - offL
        else:
            b = offR - (offL + chunk)
            uniquely = b if b > 0 else 0
        if prospect == first or prospect == last:
            valuep += (POINTS_PER_BASE1 + POINTS_FOR_TOTAL_LIST_WIDTH) * uniquely
        else:
            valuep += POINTS_PER_BASE1 * uniquely
        return int(F(valuep) * key_weight) + valuem

    def find_worst_greedy(self, offsets, lengths, weights, chunk, lists):
        mn = (1 << 63) - 1
        worst = -1
        clamp = lambda v: max(-(1 << 31), min((1 << 31) - 1, v))
        for i in range(len(lists)):
            value = self.value_of_element(offsets, lengths, weights[i], chunk, lists, i)
            if value < mn:
                if mn < EARLY_TERMINATION_SCORE and i != 0:
                    return i, clamp(value)
                mn = value; worst = i
        return worst, clamp(mn)

    def trim_by_greedy(self, offsets, key_scores, max_hit_lists, keys):
        weights = [F(ks) * self.INV_BASE_KEY_HIT_SCORE for ks in key_scores]
        H = self.lengthHistogram
        limit = max(SMALL_GENOME_LIST, int(H[self.MAX_AVERAGE_LIST_TO_SEARCH])) * len(keys)
        limit2 = max(SMALL_GENOME_LIST, int(H[self.MAX_AVERAGE_LIST_TO_SEARCH2]))
        limit3 = max(SMALL_GENOME_LIST, int(H[self.MAX_SHORTEST_LIST_TO_SEARCH]))
        total = 0; initial = 0
        shortest = (1 << 31) - 2; shortest2 = (1 << 31) - 1
        lengths = []
        for key in keys:
            x = self.count(key)
            lengths.append(x); total += x; initial += 0 if x == 0 else 1
            if x > 0 and x < shortest2:
                shortest2 = x
                if shortest2 < shortest:
                    shortest2 = shortest; shortest = x
        if initial < MIN_APPROX_HITS_TO_KEEP:
            return initial
        if shortest > limit3:
            for i in range(len(keys)):
                keys[i] = -1
            return 0
        hits_count = initial
        while hits_count >= MIN_APPROX_HITS_TO_KEEP and (total > limit or total // initial > limit2 or hits_count > max_hit_lists):
            lists = [i for i in range(len(lengths)) if lengths[i] > 0][:hits_count]
            worst_index, worst_value = self.find_worst_greedy(offsets, lengths, weights, self.KEYLEN, lists)
            worst = lists[worst_index]
            total -= lengths[worst]
            if worst_value > 0 or lengths[worst] < SMALL_GENOME_LIST:
                return hits_count
            hits_count -= 1
            lengths[worst] = 0
            keys[worst] = -1
        return hits_count

    # ---- hit lists of one block ----
    def get_hits(self, keys, chrom):
        b = self.block_of(chrom)
        starts, stops, n = [], [], 0
        for key in keys:
            st = sp = -1
            if key >= 0:
                ln = self.count(key)
                if 0 < ln < (1 << 31) - 1:
                    len2 = b.length_key(key)
                    if len2 > 0:
                        st = int(b.starts[key]); sp = st + len2; n += 1
            starts.append(st); stops.append(sp)
        return n, starts, stops

    @staticmethod
    def shrink(starts, stops, offsets, key_scores):
        keep = [i for i in range(len(offsets)) if starts[i] >= 0]
        if len(keep) == len(offsets):
            return starts, stops, offsets, key_scores
        return [starts[i] for i in keep], [stops[i] for i in keep], [offsets[i] for i in keep], [key_scores[i] for i in keep]

    # ---- quick scores ----
    def max_score_z(self, offsets):
        score = 0; a0 = b0 = -1
        for a in offsets:
            if b0 < a:
                score += b0 - a0; a0 = a
            b0 = a + self.KEYLEN
        score += b0 - a0
        return score * Z_SCORE_MULT

    def max_quick_score(self, offsets, key_scores):
        return sum(key_scores) + self.max_score_z(offsets) + Y_SCORE_MULT * (offsets[-1] - offsets[0])

    def _indel_penalty(self, offset):
        return min(self.INDEL_PENALTY + self.INDEL_PENALTY_MULT * offset, self.MAX_PENALTY_FOR_MISALIGNED_HIT)

    def score_right(self, locs, key_scores, center, num_hits):
        score = 0; loc = locs[center]
        for i in range(center + 1, num_hits):
            if locs[i] >= 0:
                prev = loc; loc = locs[i]
                offset = absdif(loc, prev)
                if offset <= MAX_INDEL:
                    score += key_scores[i]
                    if offset != 0:
                        score -= self._indel_penalty(offset)
                else:
                    loc = prev
        return score

    def score_left(self, locs, key_scores, center):
        score = 0; loc = locs[center]
        for i in range(center - 1, -1, -1):
            if locs[i] >= 0:
                prev = loc; loc = locs[i]
                offset = absdif(loc, prev)
                if offset <= MAX_INDEL:
                    score += key_scores[i]
                    if offset != 0:
                        score -= self._indel_penalty(offset)
                else:
                    loc = prev
        return score

    @staticmethod
    def score_y(locs, center, offsets):
        c = locs[center]; right = -1; i = len(offsets) - 1
        while right < center:
            if locs[i] == c:
                right = i
            i -= 1
        return offsets[right] - offsets[center]

    def quick_score(self, locs, key_scores, center, offsets, num_approx, num_hits):
        if num_approx == 1:
            return key_scores[center]
        x = key_scores[center] + self.score_left(locs, key_scores, center) + self.score_right(locs, key_scores, center, num_hits) - center
        return x + Y_SCORE_MULT * self.score_y(locs, center, offsets)

    def score_z2(self, locs, center, offsets, num_approx, num_hits):
        if num_approx == 1:
            return self.SCOREZ_1KEY
        c = locs[center]
        max_loc = c + MAX_INDEL2; min_loc = max(0, c - MAX_INDEL)
        score = 0; a0 = b0 = -1
        for i in range(num_hits):
            if min_loc <= locs[i] <= max_loc:
                a = offsets[i]
                if b0 < a:
                    score += b0 - a0; a0 = a
                b0 = a + self.KEYLEN
        score += b0 - a0
        return score * Z_SCORE_MULT

    def calc_approx_hits_cutoff(self, keys, hits, current, perfect):
        reduction = min(max(hits // self.HIT_REDUCTION_DIV, self.MAX_HITS_REDUCTION2), max(self.MAXIMUM_MAX_HITS_REDUCTION, keys // 8))
        r = max(MIN_APPROX_HITS_TO_KEEP, current, hits - reduction)
        if perfect:
            r = max(r, keys - MAX_HITS_REDUCTION_PERFECT)
        return r

    # ---- heap set-up shared by the two walks ----
    def _first_site(self, a, offset, base_chrom):
        if (a & self.SITE_MASK) >= offset:
            return a - offset
        ch = self.number_to_chrom(a, base_chrom)
        st2 = max(self.number_to_site(a) - offset, 0)
        return self.to_number(st2, ch)

    def _load_heap(self, b, starts, stops, offsets, base_chrom):
        self.heap.clear()
        triples, values, sizes = [], [], []
        for i in range(len(offsets)):
            t = Quad(i)
            t.row = starts[i]; t.list = b.sites
            t.site = self._first_site(int(b.sites[starts[i]]), offsets[i], base_chrom)
            sizes.append(b.length(starts[i], stops[i]))
            triples.append(t); values.append(t.site)
            self.heap.add(t)
        return triples, values, sizes

    def _count_nearby(self, values, site, minsite, num_hits, cutoff):
        approx = 0; max_nearby = site
        chances = num_hits - cutoff; column = 0
        maxsite = site + MAX_INDEL2
        while column < num_hits and chances >= 0:
            x = values[column]
            if minsite <= x <= maxsite:
                if x > max_nearby:
                    max_nearby = x
                approx += 1
            else:
                chances -= 1
            column += 1
        return approx, max_nearby

    # ---- prescan ----
    def find_max_qscore2(self, starts, stops, offsets, key_scores, base_chrom_, prev_max_hits, early_exit, perfect_only):
        num_hits = len(offsets)
        base_chrom = self.base_chrom(base_chrom_)
        b = self.block_of(base_chrom_)
        heap = self.heap
        triples, values, sizes = self._load_heap(b, starts, stops, offsets, base_chrom)
        max_quick = self.max_quick_score(offsets, key_scores)
        top_q = -999999999; max_hits = 0
        if perfect_only:
            cutoff = num_hits; indel_cutoff = 0
        else:
            cutoff = max(prev_max_hits, min(MIN_APPROX_HITS_TO_KEEP, num_hits - 1)); indel_cutoff = MAX_INDEL2
        while not heap.is_empty():
            t = heap.peek()
            site, center = t.site, t.column
            approx, _ = self._count_nearby(values, site, site - min(MAX_INDEL, indel_cutoff), num_hits, cutoff)
            if approx >= cutoff:
                q = self.quick_score(values, key_scores, center, offsets, approx, num_hits) + self.score_z2(values, center, offsets, approx, num_hits)
                if q > top_q:
                    max_hits = max(approx, max_hits)
                    cutoff = max(cutoff, approx - 1)
                    top_q = q
                    if q >= max_quick and early_exit:
                        return top_q, max_hits
            while heap.peek().site == site:
                t2 = heap.poll()
                row, col = t2.row + 1, t2.column
                if row < stops[col]:
                    t2.row = row
                    t2.site = self._first_site(int(t2.list[row]), offsets[col], base_chrom)
                    values[col] = t2.site
                    heap.add(t2)
                elif early_exit and (perfect_only or heap.size < cutoff):
                    return top_q, max_hits
                if heap.is_empty():
                    break
        return top_q, max_hits

    def prescan_all_blocks(self, best, pm, all_covered):
        keysP, key_scoresP, offsetsP = pm[0]
        best_q = 0; max_hits = 0; min_hits_to_score = MIN_APPROX_HITS_TO_KEEP
        max_quick = self.max_quick_score(offsetsP, key_scoresP)
        ncyc = 2 * len(self.blocks)
        counts = [len(keysP)] * (ncyc + 2); scores = [max_quick] * (ncyc + 2)
        cycle = 0
        chrom = self.minChrom
        while chrom <= self.maxChrom:
            for pmi in range(2):
                keys, key_scores, offsets = pm[pmi]
                n, starts, stops = self.get_hits(keys, chrom)
                if n < min_hits_to_score:
                    scores[cycle] = -9999; counts[cycle] = 0
                else:
                    if n < len(keys):
                        starts, stops, offsets, key_scores = self.shrink(starts, stops, offsets, key_scores)
                    q, h = self.find_max_qscore2(starts, stops, offsets, key_scores, chrom, min_hits_to_score, True, best_q >= max_quick and all_covered)
                    scores[cycle] = q; counts[cycle] = h
                    best_q = max(q, best_q); max_hits = max(max_hits, h)
                    if best_q >= max_quick and all_covered:
                        min_hits_to_score = max(min_hits_to_score, max_hits)
                        best[1] = max(best[1], max_hits); best[3] = max(best[3], best_q)
                        return counts, scores
                cycle += 1
            chrom = (chrom & self.CHROM_MASK_HIGH) + self.CHROMS_PER_BLOCK
        best[1] = max(best[1], max_hits); best[3] = max(best[3], best_q)
        return counts, scores

    # ---- extension ----
    def extend_score(self, bases, base_scores, offsets, values, chrom, center, loc_array, num_hits):
        K = self.KEYLEN
        center_val = values[center]
        center_loc = self.number_to_site(center_val)
        min_val = center_val - MAX_INDEL; max_val = center_val + MAX_INDEL2
        ref = self.chroms[chrom]
        L = len(bases)
        for i in range(L):
            loc_array[i] = -1
        keynum = 0
        for i in range(num_hits):
            value = values[i]
            if min_val <= value <= max_val:
                refbase = self.number_to_site(value)
                keynum += 1
                callbase = offsets[i]
                misses = 0
                cloc = callbase + K - 1; rloc = refbase + cloc
                while cloc >= 0 and rloc >= 0 and rloc < len(ref):
                    old = loc_array[cloc]
                    if old == refbase:
                        break
                    if misses > 0 and old >= 0:
                        break
                    if bases[cloc] == ref[rloc]:
                        if old < 0 or refbase == center_loc:
                            loc_array[cloc] = refbase
                    else:
                        misses += 1
                        if old >= 0 or keynum > 1:
                            break
                    cloc -= 1; rloc -= 1
        for i in range(num_hits):
            value = values[i]
            if min_val <= value <= max_val:
                refbase = self.number_to_site(value)
                callbase = offsets[i]
                misses = 0
                cloc = callbase + K; rloc = refbase + cloc
                while cloc < L and rloc < len(ref):
                    old = loc_array[cloc]
                    if old == refbase:
                        break
                    if misses > 0 and old >= 0:
                        break
                    if bases[cloc] == ref[rloc]:
                        if old < 0 or refbase == center_loc:
                            loc_array[cloc] = refbase
                    else:
                        misses += 1
                        if old >= 0:
                            break
                    cloc += 1; rloc += 1
        for i in range(L):
            if bases[i] == ord("N"):
                loc_array[i] = -2
        return calc_affine_score(loc_array, base_scores)

    def set_perfect(self, ss, bases):
        """SiteScore.setPerfect(bases) (stream/SiteScore.java:239-292)."""
        L = len(bases)
        if L != ss.stop - ss.start + 1:
            ss.perfect = ss.semiperfect = False
            return
        ref = self.chroms[ss.chrom]
        ss.perfect = ss.semiperfect = True
        refloc, readloc, N = ss.start, 0, 0
        mx = min(ss.stop, len(ref) - 1); nlimit = L // 2
        if ss.start < 0:
            N -= ss.start; readloc -= ss.start; refloc -= ss.start
            ss.perfect = False
        if ss.stop >= len(ref):
            N += ss.stop - len(ref) + 1
            ss.perfect = False
        if N > nlimit:
            ss.perfect = ss.semiperfect = False
            return
        bn = ord("N")
        while refloc <= mx:
            c = bases[readloc]; r = ref[refloc]
            if c != r or c == bn:
                ss.perfect = False
                if c == bn:
                    ss.semiperfect = False
                bail = r != bn
                if not bail:
                    N += 1
                    bail = N > nlimit
                if bail:
                    ss.semiperfect = False
                    return
            refloc += 1; readloc += 1
        ss.semiperfect = ss.semiperfect and N <= nlimit
        ss.perfect = ss.perfect and ss.semiperfect and N == 0

    @staticmethod
    def _move_limits(ss):
        if ss.gaps is not None:
            raise GapFixNeeded()

    # ---- slowWalk3 ----
    def slow_walk3(self, starts, stops, bases, base_scores, key_scores, offsets, base_chrom_, strand, ssl, best, all_covered, max_score, fully_defined):
        num_keys = len(offsets)
        max_quick = self.max_quick_score(offsets, key_scores)
        starts, stops, offsets, key_scores = self.shrink(starts, stops, offsets, key_scores)
        num_hits = len(offsets)
        filter_by_qscore = num_keys >= 5
        min_score = jint(MIN_SCORE_MULT * F(max_score))
        min_quick = jint(MIN_QSCORE_MULT * F(max_quick))
        base_chrom = self.base_chrom(base_chrom_)
        b = self.block_of(base_chrom_)
        heap = self.heap
        L = len(bases)
        loc_array = [0] * L
        top = best[0]
        cutoff = max(min_score, jint(F(top) * DYNAMIC_SCORE_THRESH))
        qcutoff = max(best[2], min_quick)
        bestq = best[3]; max_hits = best[1]; perfects = best[5]
        hits_cutoff = self.calc_approx_hits_cutoff(num_keys, max_hits, MIN_APPROX_HITS_TO_KEEP, top >= max_score)
        if hits_cutoff > num_hits:
            return
        short_circuit = all_covered and num_keys == num_hits and filter_by_qscore
        if top >= max_score:
            qcutoff = max(qcutoff, jint(F(max_quick) * DYNAMIC_QSCORE_THRESH_PERFECT))
        triples, values, sizes = self._load_heap(b, starts, stops, offsets, base_chrom)

        def finish():
            best[0] = max(best[0], top); best[1] = max(best[1], max_hits); best[2] = max(best[2], qcutoff); best[3] = max(best[3], bestq)
            best[4] = max_quick; best[5] = perfects

        prev = None
        while not heap.is_empty():
            t = heap.peek()
            site, center = t.site, t.column
            approx, max_nearby = self._count_nearby(values, site, site - MAX_INDEL, num_hits, hits_cutoff)
            if approx >= hits_cutoff:
                q = self.quick_score(values, key_scores, center, offsets, approx, num_hits) if filter_by_qscore else qcutoff
                q += self.score_z2(values, center, offsets, approx, num_hits)
                map_start, map_stop = site, max_nearby
                loc_valid = False
                if q < qcutoff:
                    score = -1
                else:
                    chrom = self.number_to_chrom(site, base_chrom)
                    if short_circuit and q == max_quick:
                        score = max_score
                    else:
                        score = self.extend_score(bases, base_scores, offsets, values, chrom, center, loc_array, num_hits)
                        loc_valid = True
                        located = [x for x in loc_array if x > -1]
                        if not located:
                            raise AssertionError("anomaly: extendScore located no base")
                        map_start = self.to_number(min(located), chrom)
                        map_stop = self.to_number(max(located), chrom)
                    if score == max_score:
                        qcutoff = max(qcutoff, jint(F(max_quick) * DYNAMIC_QSCORE_THRESH_PERFECT))
                        hits_cutoff = self.calc_approx_hits_cutoff(num_keys, max_hits, MIN_APPROX_HITS_TO_KEEP, True)
                    if score >= cutoff:
                        qcutoff = max(qcutoff, jint(F(q) * DYNAMIC_QSCORE_THRESH))
                        bestq = max(q, bestq)
                if score >= cutoff:
                    if score > top:
                        max_hits = max(approx, max_hits)
                        hits_cutoff = self.calc_approx_hits_cutoff(num_keys, max_hits, hits_cutoff, top >= max_score)
                        cutoff = max(cutoff, jint(F(score) * DYNAMIC_SCORE_THRESH))
                        if score >= max_score:
                            cutoff = max(cutoff, jint(F(score) * F(0.95)))
                        top = score
                    chrom = self.number_to_chrom(map_start, base_chrom)
                    site2 = self.number_to_site(map_start)
                    site3 = self.number_to_site(map_stop) + L - 1
                    gap_array = None
                    if site3 - site2 >= MINGAP + L:
                        assert loc_valid
                        gap_array = make_gap_array(loc_array, site2, MINGAP)
                        if gap_array is not None:
                            gap_array[0] = min(gap_array[0], site2)
                            gap_array[-1] = max(gap_array[-1], site3)
                    ss = None
                    perfect1 = score == max_score and fully_defined
                    inbounds = site2 >= 0 and site3 < len(self.chroms[chrom])
                    if inbounds and gap_array is None and prev is not None and prev.chrom == chrom and prev.strand == strand and overlap(prev.start, prev.stop, site2, site3):
                        better = max(score, prev.score)
                        min_start = min(prev.start, site2); max_stop = max(prev.stop, site3)
                        perfect2 = prev.score == max_score and fully_defined
                        short_enough = (not LIMIT_SUBSUMPTION_LENGTH_TO_2X) or (max_stop - min_start < 2 * L)
                        if prev.start == site2 and prev.stop == site3:
                            prev.score = better
                            prev.perfect = prev.perfect or perfect1 or perfect2
                            if prev.perfect:
                                prev.semiperfect = True
                        elif short_enough and prev.start == site2 and not prev.semiperfect:
                            if perfect2:
                                pass
                            elif perfect1:
                                prev.stop = site3; self._move_limits(prev)
                                if not prev.perfect:
                                    perfects += 1
                                prev.perfect = prev.semiperfect = True
                            else:
                                prev.stop = max_stop; self._move_limits(prev)
                                self.set_perfect(prev, bases)
                            prev.score = better
                        elif short_enough and prev.stop == site3 and not prev.semiperfect:
                            if perfect2:
                                pass
                            elif perfect1:
                                prev.start = site2; self._move_limits(prev)
                                if not prev.perfect:
                                    perfects += 1
                                prev.perfect = prev.semiperfect = True
                            else:
                                prev.start = min_start; self._move_limits(prev)
                                self.set_perfect(prev, bases)
                            prev.score = better
                        else:                                           # SUBSUME_OVERLAPPING_SITES is false: class 5, a new site
                            ss = SiteScore(chrom, strand, site2, site3, approx, score, perfect1)
                            if not perfect1:
                                self.set_perfect(ss, bases)
                    elif inbounds:
                        ss = SiteScore(chrom, strand, site2, site3, approx, score, perfect1)
                        if not perfect1:
                            self.set_perfect(ss, bases)
                        ss.gaps = gap_array
                    if ss is not None:
                        ssl.append(ss)
                        stop_now = False
                        if ss.perfect:
                            if prev is None or not prev.perfect or not ss.overlaps(prev):
                                perfects += 1
                                if self.QUIT_AFTER_TWO_PERFECTS and perfects >= 2:
                                    stop_now = True
                        if stop_now:
                            break
                        prev = ss
            returned = False
            while heap.peek().site == site:
                t2 = heap.poll()
                row, col = t2.row + 1, t2.column
                if row < stops[col]:
                    t2.row = row
                    t2.site = self._first_site(int(t2.list[row]), offsets[col], base_chrom)
                    values[col] = t2.site
                    heap.add(t2)
                elif heap.size < hits_cutoff:
                    returned = True
                    break
                if heap.is_empty():
                    break
            if returned:
                break
        finish()

    # ---- BBIndex.find(basesP, ...) ----
    def find(self, basesP, base_scoresP, offsetsP, key_scoresP):
        """-> dict(sites=[SiteScore...], num_hits, max_score, max_quick_score, best_scores[6] or None when find() returned before the walks)."""
        K = self.KEYLEN
        basesP = bytes(basesP); L = len(basesP)
        base_scoresP = [int(x) for x in base_scoresP]
        offsetsP = [int(x) for x in offsetsP]; key_scoresP = [int(x) for x in key_scoresP]
        out = {"sites": [], "num_hits": 0, "max_score": None, "max_quick_score": None, "best_scores": None, "gapfix": False}

        def to_number(a, bb):
            o = 0
            for i in range(a, bb + 1):
                x = BASE_TO_NUMBER.get(basesP[i], -1)
                if x < 0:
                    return -1
                o = (o << 2) | x
            return o
        keys_original = [to_number(o, o + K - 1) for o in offsetsP]
        keysP = list(keys_original)
        max_len = self.MAX_USABLE_LENGTH
        num_hits = self.count_hits(keysP, max_len)
        if num_hits > 0:
            trigger = (3 * len(keysP)) // 4
            for lim, ml in ((4, (max_len * 3) // 2), (3, max_len * 2), (3, max_len * 3), (2, max_len * 5)):
                if num_hits < lim and num_hits < trigger:
                    keysP = list(keys_original)
                    num_hits = self.count_hits(keysP, ml)
        if num_hits < len(keysP):
            offsetsP, keysP, key_scoresP = self.shrink2(offsetsP, keysP, key_scoresP)
        max_lists = max(jint(HIT_FRACTION_TO_RETAIN * F(len(keysP))), MIN_HIT_LISTS_TO_RETAIN)
        num_hits = self.trim_by_greedy(offsetsP, key_scoresP, max_lists, keysP)
        out["num_hits"] = num_hits
        if num_hits < MIN_APPROX_HITS_TO_KEEP:
            return out
        if num_hits < len(keysP):
            offsetsP, keysP, key_scoresP = self.shrink2(offsetsP, keysP, key_scoresP)
        n = len(offsetsP)
        offsetsM = [L - (offsetsP[n - 1 - i] + K) for i in range(n)]
        keysM = [rc_key(keysP[n - 1 - i], K) for i in range(n)]
        basesM = bytes(COMPLEMENT.get(c, c) for c in reversed(basesP))
        base_scoresM = base_scoresP[::-1]; key_scoresM = key_scoresP[::-1]
        max_quick = self.max_quick_score(offsetsP, key_scoresP)
        best = [0] * 6
        prescan = num_hits >= 5
        precounts = prescores = None
        hits_cutoff = 0
        qscore_cutoff = jint(MIN_QSCORE_MULT * F(max_quick))
        all_covered = True
        if offsetsP[0] != 0 or offsetsP[-1] != L - K:
            all_covered = False
        else:
            for i in range(1, n):
                if offsetsP[i] > offsetsP[i - 1] + K:
                    all_covered = False
                    break
        pretend = all_covered or n >= len(keys_original) - 4 or (n >= 9 and (offsetsP[-1] - offsetsP[0] + K) > max(40, jint(F(L) * F(0.75))))
        if prescan:
            precounts, prescores = self.prescan_all_blocks(best, [(keysP, key_scoresP, offsetsP), (keysM, key_scoresM, offsetsM)], pretend)
            if best[1] < MIN_APPROX_HITS_TO_KEEP:
                return out
            if F(best[3]) < F(max_quick) * MIN_QSCORE_MULT2:
                return out
            if best[3] >= max_quick and pretend:
                hits_cutoff = self.calc_approx_hits_cutoff(n, best[1], MIN_APPROX_HITS_TO_KEEP, True)
                qscore_cutoff = max(qscore_cutoff, jint(F(best[3]) * DYNAMIC_QSCORE_THRESH_PERFECT))
            else:
                hits_cutoff = self.calc_approx_hits_cutoff(n, best[1], MIN_APPROX_HITS_TO_KEEP, False)
                qscore_cutoff = max(qscore_cutoff, jint(F(best[3]) * PRESCAN_QSCORE_THRESH))
        max_score = POINTS_MATCH + (L - 1) * POINTS_MATCH2 + sum(base_scoresP)
        fully_defined = all(c in BASE_TO_NUMBER for c in basesP)
        out["max_score"] = max_score; out["max_quick_score"] = max_quick
        ssl = out["sites"]
        cycle = 0
        chrom = self.minChrom
        try:
            while chrom <= self.maxChrom:
                for strand, (keys, bases, bscores, kscores, offsets) in enumerate(((keysP, basesP, base_scoresP, key_scoresP, offsetsP),
                                                                                 (keysM, basesM, base_scoresM, key_scoresM, offsetsM))):
                    if precounts is None or precounts[cycle] >= hits_cutoff or prescores[cycle] >= qscore_cutoff:
                        nh, starts, stops = self.get_hits(keys, chrom)
                        if nh >= MIN_APPROX_HITS_TO_KEEP:
                            self.slow_walk3(starts, stops, bases, bscores, kscores, offsets, chrom, strand, ssl, best, all_covered, max_score, fully_defined)
                    if self.QUIT_AFTER_TWO_PERFECTS and best[5] >= 2:
                        raise StopIteration
                    cycle += 1
                chrom = (chrom & self.CHROM_MASK_HIGH) + self.CHROMS_PER_BLOCK
        except StopIteration:
            pass
        except GapFixNeeded:
            out["gapfix"] = True
        out["best_scores"] = best
        return out


def rc_key(kmer, k):
    """AminoAcid.reverseComplementBinaryFast (dna/AminoAcid.java:258-271), without the byte table."""
    out = 0
    for _ in range(k):
        out = (out << 2) | ((~kmer) & 3)
        kmer >>= 2
    return out


def calc_affine_score(loc_array, base_scores):
    """MSA.calcAffineScore(locArray, baseScores, bases) (MultiStateAligner11tsJNI.java:871-941)."""
    score = 0; last_loc = -3; last_value = -1; time_in_mode = 0
    for i, loc in enumerate(loc_array):
        if loc > 0:
            if loc == last_value:
                score += POINTS_MATCH2 + base_scores[i]
            elif loc == last_loc or last_loc < 0:
                score += POINTS_MATCH + base_scores[i]
            elif loc < last_loc:
                score += POINTS_MATCH + base_scores[i] + POINTS_DEL
                dif = last_loc - loc + 1
                if dif > MINGAP:
                    rem = dif % GAPLEN
                    div = (dif - GAPBUFFER2) // GAPLEN
                    score += div * POINTS_GAP
                    dif = rem + GAPBUFFER2
                if dif > LIMIT_FOR_COST_5:
                    score += ((dif - LIMIT_FOR_COST_5 + MASK5) // TIMESLIP) * POINTS_DEL5
                    dif = LIMIT_FOR_COST_5
                if dif > LIMIT_FOR_COST_4:
                    score += (dif - LIMIT_FOR_COST_4) * POINTS_DEL4
                    dif = LIMIT_FOR_COST_4
                if dif > LIMIT_FOR_COST_3:
                    score += (dif - LIMIT_FOR_COST_3) * POINTS_DEL3
                    dif = LIMIT_FOR_COST_3
                if dif > 1:
                    score += (dif - 1) * POINTS_DEL2
                time_in_mode = 1
            else:
                score += POINTS_MATCH + base_scores[i] + POINTS_INS_ARRAY_C[min(loc - last_loc, 5)]
                time_in_mode = 1
            last_loc = loc
        elif loc == -1:
            if last_value < 0 and time_in_mode > 0:
                time_in_mode += 1
                score += POINTS_SUB_ARRAY[time_in_mode]
            else:
                score += POINTS_SUB
                time_in_mode = 1
        else:
            time_in_mode = 0
            score += POINTS_NOCALL
        last_value = loc
    return score


def make_gap_array(loc_array, min_loc, min_gap):
    """BBIndex.makeGapArray (BBIndex.java:2837-2878); destroys loc_array like the original."""
    do_sort = False
    if loc_array[0] < 0:
        loc_array[0] = min_loc
    for i in range(1, len(loc_array)):
        if loc_array[i] < 0:
            loc_array[i] = loc_array[i - 1] + 1
        else:
            loc_array[i] += i
        if loc_array[i] < loc_array[i - 1]:
            do_sort = True
    if do_sort:
        loc_array.sort()
    gaps = sum(1 for i in range(1, len(loc_array)) if loc_array[i] - loc_array[i - 1] > min_gap)
    if gaps < 1:
        return None
    out = [loc_array[0]]
    for i in range(1, len(loc_array)):
        if loc_array[i] - loc_array[i - 1] > min_gap:
            out += [loc_array[i - 1], loc_array[i]]
    out.append(loc_array[-1])
    return out
