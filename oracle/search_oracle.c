/*
 * search_oracle.c — TEST INFRASTRUCTURE ONLY.
 * C restatement of BBMap's index search for one read, Java-only in the reference:
 *   BBIndex.find                      current/align2/BBIndex.java:403-639   (key filtering, strand/block loop, bestScores[6])
 *   countHits / getHits / shrink(2)   :353-391, 783-851
 *   trimExcessHitListsByGreedy        :266-350   + Solver.findWorstGreedy / valueOfElement  current/align2/Solver.java:48-152
 *   prescanAllBlocks / findMaxQscore2 :642-741, 2294-2450
 *   slowWalk3                         :1219-1706 (heap walk = repeated arg-min over (site, column), Quad.java:18-22)
 *   quickScore / scoreLeft / scoreRight / scoreZ2 / maxScoreZ / maxQuickScore   :2472-2511, 2882-2914, 2948-3035
 *   AbstractIndex.scoreY              current/align2/AbstractIndex.java:52-80
 *   extendScore / makeGapArray        :2558-2878
 *   MSA.calcAffineScore(locArray, baseScores, bases)   current/align2/MultiStateAligner11tsJNI.java:871-941
 *   calcApproxHitsCutoff              :3267-3294
 *   SiteScore.setPerfect              current/stream/SiteScore.java:239-292
 *   AminoAcid.reverseComplementBases  current/dna/AminoAcid.java:203-211
 * PARITY UNPINNED against Java (no JVM).  Unsupported corner (flagged, never silently wrong): subsuming into a previous
 * site that carries a gap array (needs GapTools.fixGaps).
 */
#pragma GCC optimize ("fp-contract=off")
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "search_oracle.h"

extern int orc_rcomp_key_fast(int kmer, int k);

enum { MAX_INDEL = 16000, MAX_INDEL2 = 32000, MINGAP = 256, GAPLEN = 128, GAPBUFFER2 = 128,
       Y_SCORE_MULT = 10, Z_SCORE_MULT = 20, BASE_HIT_SCORE = 100, INDEL_PENALTY_MULT = 20,
       MIN_HIT_LISTS_TO_RETAIN = 6, SMALL_GENOME_LIST = 20, MAXK = ORC_MAX_KEYS };
static const float MIN_SCORE_MULT = 0.15f, MIN_QSCORE_MULT = 0.025f, MIN_QSCORE_MULT2 = 0.1f, DYNAMIC_SCORE_THRESH = 0.84f,
                   DYNAMIC_QSCORE_THRESH = 0.6f, DYNAMIC_QSCORE_THRESH_PERFECT = 0.8f, HIT_FRACTION_TO_RETAIN = 0.85f;
#define PRESCAN_QSCORE_THRESH (DYNAMIC_QSCORE_THRESH * .95f)

static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int absdif(int a, int b) { return a > b ? a - b : b - a; }

typedef struct {
    const orc_search_index* X;
    int K, baseKeyHitScore, indelPenalty, maxPenaltyMisaligned, scoreZ1Key;
    int siteMask, shift, lowMask, highMask, cpb;
    /* per-walk scratch */
    int values[MAXK], sizes[MAXK], rows[MAXK], stopsA[MAXK], active[MAXK];
    int locArray[ORC_MAX_READ];
    int status;
} ctx_t;

static inline int count_key(const ctx_t* c, int key) { return c->X->counts[key]; }
static inline int to_number(const ctx_t* c, int site, int chrom) { return ((chrom & c->lowMask) << c->shift) | site; }
static inline int number_to_chrom(const ctx_t* c, int number, int baseChrom) { return (int)((uint32_t)number >> c->shift) + (baseChrom & c->highMask); }
static inline int number_to_site(const ctx_t* c, int number) { return number & c->siteMask; }
static inline int base_chrom(const ctx_t* c, int chrom) { return imax(0, chrom & c->highMask); }
static inline const orc_search_block* block_of(const ctx_t* c, int chrom) { return &c->X->blocks[((chrom & c->highMask) - (1 & c->highMask)) / c->cpb]; }   /* blocks[0] holds chromosome 1 (index[baseChrom(1)] in the reference) */

static int block_length(const orc_search_block* b, int key) {        /* Block.length(key), Block.java:62-66 */
    const int x = b->starts[key + 1] - b->starts[key];
    if (x == 0) return 0;
    return b->sites[b->starts[key]] != -1 ? x : 0;
}

/* ---------------- calcApproxHitsCutoff (BBIndex.java:3267-3294; not perfect/semiperfect mode) ---------------- */
static int approx_hits_cutoff(const ctx_t* c, int keys, int hits, int currentCutoff, int perfect) {
    const orc_index_cfg* g = c->X->cfg;
    const int mahtk = 1;
    const int reduction = imin(imax(hits / g->hit_reduction_div, g->max_hits_reduction2), imax(g->maximum_max_hits_reduction, keys / 8));
    int r = hits - reduction;
    r = imax(mahtk, imax(currentCutoff, r));
    if (perfect) r = imax(r, keys - 0);
    return r;
}

static int max_score_z(const ctx_t* c, const int* offsets, int n) {
    int score = 0, a0 = -1, b0 = -1;
    for (int i = 0; i < n; i++) { const int a = offsets[i]; if (b0 < a) { score += b0 - a0; a0 = a; } b0 = a + c->K; }
    score += b0 - a0;
    return score * Z_SCORE_MULT;
}
static int max_quick_score(const ctx_t* c, const int* offsets, const int* keyScores, int n) {
    int x = 0;
    for (int i = 0; i < n; i++) x += keyScores[i];
    const int y = Y_SCORE_MULT * (offsets[n - 1] - offsets[0]);
    x += max_score_z(c, offsets, n);
    return x + y;
}

static int score_right(const ctx_t* c, const int* locs, const int* keyScores, int centerIndex, int numHits) {
    int score = 0, prev, loc = locs[centerIndex];
    for (int i = centerIndex + 1; i < numHits; i++) {
        if (locs[i] >= 0) {
            prev = loc; loc = locs[i];
            const int offset = absdif(loc, prev);
            if (offset <= MAX_INDEL) {
                score += keyScores[i];
                if (offset != 0) score -= imin(c->indelPenalty + INDEL_PENALTY_MULT * offset, c->maxPenaltyMisaligned);
            } else loc = prev;
        }
    }
    return score;
}
static int score_left(const ctx_t* c, const int* locs, const int* keyScores, int centerIndex) {
    int score = 0, prev, loc = locs[centerIndex];
    for (int i = centerIndex - 1; i >= 0; i--) {
        if (locs[i] >= 0) {
            prev = loc; loc = locs[i];
            const int offset = absdif(loc, prev);
            if (offset <= MAX_INDEL) {
                score += keyScores[i];
                if (offset != 0) score -= imin(c->indelPenalty + INDEL_PENALTY_MULT * offset, c->maxPenaltyMisaligned);
            } else loc = prev;
        }
    }
    return score;
}
static int score_y(const int* locs, int centerIndex, const int* offsets, int n) {
    const int center = locs[centerIndex];
    int rightIndex = -1;
    for (int i = n - 1; rightIndex < centerIndex; i--) if (locs[i] == center) rightIndex = i;
    return offsets[rightIndex] - offsets[centerIndex];
}
static int quick_score(const ctx_t* c, const int* locs, const int* keyScores, int centerIndex, const int* offsets, int numApproxHits, int numHits) {
    if (numApproxHits == 1) return keyScores[centerIndex];
    const int x = keyScores[centerIndex] + score_left(c, locs, keyScores, centerIndex) + score_right(c, locs, keyScores, centerIndex, numHits) - centerIndex;
    const int y = Y_SCORE_MULT * score_y(locs, centerIndex, offsets, numHits);
    return x + y;
}
static int score_z2(const ctx_t* c, const int* locs, int centerIndex, const int* offsets, int numApproxHits, int numHits) {
    if (numApproxHits == 1) return c->scoreZ1Key;
    const int center = locs[centerIndex];
    const int maxLoc = center + MAX_INDEL2, minLoc = imax(0, center - MAX_INDEL);
    int score = 0, a0 = -1, b0 = -1;
    for (int i = 0; i < numHits; i++) {
        const int loc = locs[i];
        if (loc >= minLoc && loc <= maxLoc) { const int a = offsets[i]; if (b0 < a) { score += b0 - a0; a0 = a; } b0 = a + c->K; }
    }
    score += b0 - a0;
    return score * Z_SCORE_MULT;
}

/* ---------------- MSA.calcAffineScore(locArray, baseScores, bases) ---------------- */
static int calc_affine_score(const int* locArray, const int8_t* baseScores, int len) {
    static const int INSC[6] = {0, -395, -434, -473, -512, -551};     /* POINTS_INS_ARRAY_C[0..5] */
    int score = 0, lastLoc = -3, lastValue = -1, timeInMode = 0;
    for (int i = 0; i < len; i++) {
        const int loc = locArray[i];
        if (loc > 0) {
            if (loc == lastValue) score += 100 + baseScores[i];
            else if (loc == lastLoc || lastLoc < 0) score += 70 + baseScores[i];
            else if (loc < lastLoc) {
                score += 70 + baseScores[i];
                score += -472;
                int dif = lastLoc - loc + 1;
                if (dif > MINGAP) { const int rem = dif % GAPLEN, div = (dif - GAPBUFFER2) / GAPLEN; score += div * -2; dif = rem + GAPBUFFER2; }
                if (dif > 80) { score += ((dif - 80 + 3) / 4) * -1; dif = 80; }
                if (dif > 20) { score += (dif - 20) * -1; dif = 20; }
                if (dif > 5) { score += (dif - 5) * -9; dif = 5; }
                if (dif > 1) score += (dif - 1) * -33;
                timeInMode = 1;
            } else {
                score += 70 + baseScores[i] + INSC[imin(loc - lastLoc, 5)];
                timeInMode = 1;
            }
            lastLoc = loc;
        } else if (loc == -1) {
            if (lastValue < 0 && timeInMode > 0) { timeInMode++; score += timeInMode > 5 ? -25 : (timeInMode > 1 ? -51 : -127); }
            else { score += -127; timeInMode = 1; }
        } else { timeInMode = 0; }
        lastValue = loc;
    }
    return score;
}

/* ---------------- extendScore (BBIndex.java:2558-2757, USE_AFFINE_SCORE, KFILTER<2) ---------------- */
static int extend_score(ctx_t* c, const int8_t* bases, const int8_t* baseScores, int len, const int* offsets, const int* values,
                        int chrom, int centerIndex, int* locArray, int numHits) {
    const int centerVal = values[centerIndex], centerLoc = number_to_site(c, centerVal);
    const int minVal = centerVal - MAX_INDEL, maxVal = centerVal + MAX_INDEL2;
    const int8_t* ref = c->X->chroms + c->X->chrom_off[chrom - 1];
    const int refLen = (int)(c->X->chrom_off[chrom] - c->X->chrom_off[chrom - 1]);
    const int K = c->K;
    for (int i = 0; i < len; i++) locArray[i] = -1;
    for (int i = 0, keynum = 0; i < numHits; i++) {
        const int value = values[i];
        if (value >= minVal && value <= maxVal) {
            const int refbase = number_to_site(c, value);
            keynum++;
            const int callbase = offsets[i];
            int misses = 0;
            for (int cloc = callbase + K - 1, rloc = refbase + cloc; cloc >= 0 && rloc >= 0 && rloc < refLen; cloc--, rloc--) {
                const int old = locArray[cloc];
                if (old == refbase) break;
                if (misses > 0 && old >= 0) break;
                if (bases[cloc] == ref[rloc]) { if (old < 0 || refbase == centerLoc) locArray[cloc] = refbase; }
                else { misses++; if (old >= 0 || keynum > 1) break; }
            }
        }
    }
    for (int i = 0; i < numHits; i++) {
        const int value = values[i];
        if (value >= minVal && value <= maxVal) {
            const int refbase = number_to_site(c, value);
            const int callbase = offsets[i];
            int misses = 0;
            for (int cloc = callbase + K, rloc = refbase + cloc; cloc < len && rloc < refLen; cloc++, rloc++) {
                const int old = locArray[cloc];
                if (old == refbase) break;
                if (misses > 0 && old >= 0) break;
                if (bases[cloc] == ref[rloc]) { if (old < 0 || refbase == centerLoc) locArray[cloc] = refbase; }
                else { misses++; if (old >= 0) break; }
            }
        }
    }
    for (int i = 0; i < len; i++) if (bases[i] == 'N') locArray[i] = -2;
    return calc_affine_score(locArray, baseScores, len);
}

static int cmp_int(const void* a, const void* b) { const int x = *(const int*)a, y = *(const int*)b; return x < y ? -1 : (x > y ? 1 : 0); }

/* makeGapArray (BBIndex.java:2837-2878); destroys locArray.  Returns #ints written (0 = null). */
static int make_gap_array(int* locArray, int len, int minLoc, int minGap, int* out, int cap, int* overflow) {
    int gaps = 0, doSort = 0;
    if (locArray[0] < 0) locArray[0] = minLoc;
    for (int i = 1; i < len; i++) {
        if (locArray[i] < 0) locArray[i] = locArray[i - 1] + 1; else locArray[i] += i;
        if (locArray[i] < locArray[i - 1]) doSort = 1;
    }
    if (doSort) qsort(locArray, (size_t)len, sizeof(int), cmp_int);
    for (int i = 1; i < len; i++) if (locArray[i] - locArray[i - 1] > minGap) gaps++;
    if (gaps < 1) return 0;
    const int n = 2 + gaps * 2;
    if (n > cap) { *overflow = 1; return 0; }
    out[0] = locArray[0]; out[n - 1] = locArray[len - 1];
    for (int i = 1, j = 1; i < len; i++) if (locArray[i] - locArray[i - 1] > minGap) { out[j] = locArray[i - 1]; out[j + 1] = locArray[i]; j += 2; }
    return n;
}

/* SiteScore.setPerfect(bases) */
static void set_perfect(const ctx_t* c, orc_site* s, const int8_t* bases, int len) {
    if (len != s->stop - s->start + 1) { s->perfect = 0; s->semiperfect = 0; return; }
    const int8_t* ref = c->X->chroms + c->X->chrom_off[s->chrom - 1];
    const int refLen = (int)(c->X->chrom_off[s->chrom] - c->X->chrom_off[s->chrom - 1]);
    int perfect = 1, semiperfect = 1;
    int refloc = s->start, readloc = 0, N = 0;
    const int max = imin(s->stop, refLen - 1), nlimit = len / 2;
    if (s->start < 0) { N -= s->start; readloc -= s->start; refloc -= s->start; perfect = 0; }
    if (s->stop >= refLen) { N += (s->stop - refLen + 1); perfect = 0; }
    if (N > nlimit) { s->perfect = 0; s->semiperfect = 0; return; }
    for (; refloc <= max; refloc++, readloc++) {
        const int8_t cb = bases[readloc], r = ref[refloc];
        if (cb != r || cb == 'N') {
            perfect = 0;
            if (cb == 'N') semiperfect = 0;
            if (r != 'N' || (N = N + 1) > nlimit) { s->perfect = (int8_t)perfect; s->semiperfect = 0; return; }
        }
    }
    semiperfect = (semiperfect && (N <= nlimit));
    perfect = (perfect && semiperfect && (N == 0));
    s->perfect = (int8_t)perfect; s->semiperfect = (int8_t)semiperfect;
}

/* ---------------- getHits / shrink ---------------- */
static int get_hits(const ctx_t* c, const int* keys, int n, int chrom, int* starts, int* stops) {
    int numHits = 0;
    const orc_search_block* b = block_of(c, chrom);
    for (int i = 0; i < n; i++) {
        const int key = keys[i];
        starts[i] = -1; stops[i] = -1;
        if (key >= 0) {
            const int len = count_key(c, key);
            if (len > 0) {                                 /* maxLen = Integer.MAX_VALUE */
                const int len2 = block_length(b, key);
                if (len2 > 0) { starts[i] = b->starts[key]; stops[i] = starts[i] + len2; numHits++; }
            }
        }
    }
    return numHits;
}
static int shrink_hits(int* starts, int* stops, int* offsets, int* keyScores, int n) {
    int j = 0;
    for (int i = 0; i < n; i++) if (starts[i] >= 0) { starts[j] = starts[i]; stops[j] = stops[i]; offsets[j] = offsets[i]; keyScores[j] = keyScores[i]; j++; }
    return j;
}

/* translate a raw site to (site - offset), clamped at the chromosome start (BBIndex.java:1305-1313 etc.) */
static inline int site_minus_offset(const ctx_t* c, int a, int offset, int baseChrom) {
    if ((a & c->siteMask) >= offset) return a - offset;
    const int ch = number_to_chrom(c, a, baseChrom), st = number_to_site(c, a);
    return to_number(c, imax(st - offset, 0), ch);
}

/* The heap (QuadHeap ordered by (site, column), Quad.java:18-22) is replaced by an arg-min over the active columns: the
 * order is total, so the sequence of (site, column) visited is identical. */
static int heap_peek(const ctx_t* c, int numHits) {
    int best = -1;
    for (int i = 0; i < numHits; i++) if (c->active[i] && (best < 0 || c->values[i] < c->values[best])) best = i;
    return best;
}
static int heap_size(const ctx_t* c, int numHits) { int n = 0; for (int i = 0; i < numHits; i++) n += c->active[i]; return n; }

/* ---------------- findMaxQscore2 ---------------- */
static void find_max_qscore2(ctx_t* c, const int* starts, const int* stops, const int* offsets, const int* keyScores, int numHits,
                             int baseChrom_, int prevMaxHits, int earlyExit, int perfectOnly, int* outScore, int* outHits) {
    const int baseChrom = base_chrom(c, baseChrom_);
    const orc_search_block* b = block_of(c, baseChrom_);
    for (int i = 0; i < numHits; i++) {
        c->sizes[i] = stops[i] - starts[i];
        c->rows[i] = starts[i]; c->stopsA[i] = stops[i]; c->active[i] = 1;
        c->values[i] = site_minus_offset(c, b->sites[starts[i]], offsets[i], baseChrom);
    }
    const int maxQuickScore = max_quick_score(c, offsets, keyScores, numHits);
    int topQscore = -999999999, maxHits = 0, approxHitsCutoff, indelCutoff;
    if (perfectOnly) { approxHitsCutoff = numHits; indelCutoff = 0; }
    else { approxHitsCutoff = imax(prevMaxHits, imin(1, numHits - 1)); indelCutoff = MAX_INDEL2; }
    int t;
    while ((t = heap_peek(c, numHits)) >= 0) {
        const int site = c->values[t], centerIndex = t;
        int approxHits = 0;
        {
            const int minsite = site - imin(MAX_INDEL, indelCutoff), maxsite = site + MAX_INDEL2;
            for (int column = 0, chances = numHits - approxHitsCutoff; column < numHits && chances >= 0; column++) {
                const int x = c->values[column];
                if (x >= minsite && x <= maxsite) approxHits++; else chances--;
            }
        }
        if (approxHits >= approxHitsCutoff) {
            int qscore = quick_score(c, c->values, keyScores, centerIndex, offsets, approxHits, numHits);
            qscore += score_z2(c, c->values, centerIndex, offsets, approxHits, numHits);
            if (qscore > topQscore) {
                maxHits = imax(approxHits, maxHits);
                approxHitsCutoff = imax(approxHitsCutoff, approxHits - 1);
                topQscore = qscore;
                if (qscore >= maxQuickScore && earlyExit) { *outScore = topQscore; *outHits = maxHits; return; }
            }
        }
        int t2;
        while ((t2 = heap_peek(c, numHits)) >= 0 && c->values[t2] == site) {
            const int row = c->rows[t2] + 1, col = t2;
            if (row < c->stopsA[col]) {
                c->rows[col] = row;
                c->values[col] = site_minus_offset(c, b->sites[row], offsets[col], baseChrom);
            } else {
                c->active[col] = 0;
                /* NOTE: values[col] keeps its last site, exactly like the reference's valueArray */
                if (earlyExit && (perfectOnly || heap_size(c, numHits) < approxHitsCutoff)) { *outScore = topQscore; *outHits = maxHits; return; }
            }
        }
    }
    *outScore = topQscore; *outHits = maxHits;
}

/* ---------------- slowWalk3 ---------------- */
static void slow_walk3(ctx_t* c, int* starts, int* stops, const int8_t* bases, const int8_t* baseScores, int len, int* keyScores, int* offsets,
                       int numKeys, int baseChrom_, int strand, int obeyLimits, orc_read_result* R, int* bestScores, int allBasesCovered,
                       int maxScore, int fullyDefined, int quitAfterTwoPerfects, int* prevIdx) {
    const int maxQuickScore = max_quick_score(c, offsets, keyScores, numKeys);
    const int numHits = shrink_hits(starts, stops, offsets, keyScores, numKeys);
    const int filter_by_qscore = (numKeys >= 5);
    const int minScore = obeyLimits ? (int)(MIN_SCORE_MULT * maxScore) : (int)(MIN_SCORE_MULT * 1.25f * maxScore);
    const int minQuickScore = (int)(MIN_QSCORE_MULT * maxQuickScore);
    const int baseChrom = base_chrom(c, baseChrom_);
    const orc_search_block* b = block_of(c, baseChrom_);
    int currentTopScore = bestScores[0];
    int cutoff = imax(minScore, (int)(currentTopScore * DYNAMIC_SCORE_THRESH));
    int qcutoff = imax(bestScores[2], minQuickScore);
    int bestqscore = bestScores[3], maxHits = bestScores[1], perfectsFound = bestScores[5];
    int approxHitsCutoff = approx_hits_cutoff(c, numKeys, maxHits, 1, currentTopScore >= maxScore);
    if (approxHitsCutoff > numHits) return;
    const int shortCircuit = (allBasesCovered && numKeys == numHits && filter_by_qscore);
    if (currentTopScore >= maxScore) qcutoff = imax(qcutoff, (int)(maxQuickScore * DYNAMIC_QSCORE_THRESH_PERFECT));
    for (int i = 0; i < numHits; i++) {
        c->sizes[i] = stops[i] - starts[i];
        c->rows[i] = starts[i]; c->stopsA[i] = stops[i]; c->active[i] = 1;
        c->values[i] = site_minus_offset(c, b->sites[starts[i]], offsets[i], baseChrom);
    }
    int* locArray = c->locArray;
    int prev = -1;       /* index of prevSS in R->sites (a site made during THIS walk), -1 = null */
    int t, quit = 0;
    while (!quit && (t = heap_peek(c, numHits)) >= 0) {
        const int site = c->values[t], centerIndex = t;
        int maxNearbySite = site, approxHits = 0;
        {
            const int minsite = site - MAX_INDEL, maxsite = site + MAX_INDEL2;
            for (int column = 0, chances = numHits - approxHitsCutoff; column < numHits && chances >= 0; column++) {
                const int x = c->values[column];
                if (x >= minsite && x <= maxsite) { maxNearbySite = (x > maxNearbySite ? x : maxNearbySite); approxHits++; } else chances--;
            }
        }
        if (approxHits >= approxHitsCutoff) {
            int score;
            int qscore = filter_by_qscore ? quick_score(c, c->values, keyScores, centerIndex, offsets, approxHits, numHits) : qcutoff;
            qscore += score_z2(c, c->values, centerIndex, offsets, approxHits, numHits);
            int mapStart = site, mapStop = maxNearbySite;
            int locArrayValid = 0;
            if (qscore < qcutoff) score = -1;
            else {
                const int chrom = number_to_chrom(c, site, baseChrom);
                if (shortCircuit && qscore == maxQuickScore) score = maxScore;
                else {
                    score = extend_score(c, bases, baseScores, len, offsets, c->values, chrom, centerIndex, locArray, numHits);
                    locArrayValid = 1;
                    int mn = INT32_MAX, mx = INT32_MIN;
                    for (int i = 0; i < len; i++) { const int x = locArray[i]; if (x > -1) { if (x < mn) mn = x; if (x > mx) mx = x; } }
                    if (mn < 0 || mx < 0) { score = -99999; c->status |= ORC_ST_ANOMALY; }
                    mapStart = to_number(c, mn, chrom); mapStop = to_number(c, mx, chrom);
                }
                if (score == maxScore) {
                    qcutoff = imax(qcutoff, (int)(maxQuickScore * DYNAMIC_QSCORE_THRESH_PERFECT));
                    approxHitsCutoff = approx_hits_cutoff(c, numKeys, maxHits, 1, 1);
                }
                if (score >= cutoff) { qcutoff = imax(qcutoff, (int)(qscore * DYNAMIC_QSCORE_THRESH)); bestqscore = imax(qscore, bestqscore); }
            }
            if (score >= cutoff) {
                if (score > currentTopScore) {
                    maxHits = imax(approxHits, maxHits);
                    approxHitsCutoff = approx_hits_cutoff(c, numKeys, maxHits, approxHitsCutoff, currentTopScore >= maxScore);
                    cutoff = imax(cutoff, (int)(score * DYNAMIC_SCORE_THRESH));
                    if (score >= maxScore) cutoff = imax(cutoff, (int)(score * 0.95f));
                    currentTopScore = score;
                }
                const int chrom = number_to_chrom(c, mapStart, baseChrom);
                const int site2 = number_to_site(c, mapStart), site3 = number_to_site(c, mapStop) + len - 1;
                int gapArr[ORC_MAX_GAPS]; int ngap = 0;   /* ORC_MAX_GAPS-1 ints fit in orc_site */
                if (site3 - site2 >= MINGAP + len) {
                    int ov = 0;
                    (void)locArrayValid;
                    ngap = make_gap_array(locArray, len, site2, MINGAP, gapArr, ORC_MAX_GAPS - 1, &ov);
                    if (ov) c->status |= ORC_ST_GAP_OVERFLOW;
                    if (ngap > 0) { gapArr[0] = imin(gapArr[0], site2); gapArr[ngap - 1] = imax(gapArr[ngap - 1], site3); }
                }
                const int perfect1 = (score == maxScore && fullyDefined);
                const int chromLen = (int)(c->X->chrom_off[chrom] - c->X->chrom_off[chrom - 1]);
                const int inbounds = (site2 >= 0 && site3 < chromLen);
                int made = -1;
                orc_site* P = prev >= 0 ? &R->sites[prev] : 0;
                if (inbounds && ngap == 0 && P && P->chrom == chrom && P->strand == strand && (site2 <= P->stop && site3 >= P->start)) {
                    const int betterScore = imax(score, P->score);
                    const int minStart = imin(P->start, site2), maxStop = imax(P->stop, site3);
                    const int perfect2 = (P->score == maxScore && fullyDefined);
                    const int shortEnough = (maxStop - minStart < 2 * len);
                    if (P->start == site2 && P->stop == site3) {
                        P->score = betterScore;
                        P->perfect = (int8_t)(P->perfect || perfect1 || perfect2);
                        if (P->perfect) P->semiperfect = 1;
                    } else if (shortEnough && P->start == site2 && !P->semiperfect) {
                        if (P->ngaps) c->status |= ORC_ST_GAPFIX;
                        if (perfect2) {}
                        else if (perfect1) { P->stop = site3; if (!P->perfect) perfectsFound++; P->perfect = P->semiperfect = 1; }
                        else { P->stop = maxStop; set_perfect(c, P, bases, len); }
                        P->score = betterScore;
                    } else if (shortEnough && P->stop == site3 && !P->semiperfect) {
                        if (P->ngaps) c->status |= ORC_ST_GAPFIX;
                        if (perfect2) {}
                        else if (perfect1) { P->start = site2; if (!P->perfect) perfectsFound++; P->perfect = P->semiperfect = 1; }
                        else { P->start = minStart; set_perfect(c, P, bases, len); }
                        P->score = betterScore;
                    } else {
                        made = 1;
                    }
                } else if (inbounds) made = 1;
                if (made > 0) {
                    if (R->nsites >= ORC_MAX_SITES) { c->status |= ORC_ST_SITE_OVERFLOW; }
                    else {
                        orc_site* S = &R->sites[R->nsites];
                        memset(S, 0, sizeof(*S));
                        S->chrom = chrom; S->strand = (int8_t)strand; S->start = site2; S->stop = site3; S->hits = approxHits; S->score = score;
                        S->perfect = (int8_t)perfect1; S->semiperfect = (int8_t)perfect1;
                        if (!perfect1) set_perfect(c, S, bases, len);
                        /* gaps are attached only on the "new site" path that is not an overlap of prevSS (BBIndex.java:1640) */
                        if (!(P && inbounds && ngap == 0 && P->chrom == chrom && P->strand == strand && (site2 <= P->stop && site3 >= P->start))) {
                            S->ngaps = ngap; for (int g = 0; g < ngap; g++) S->gaps[g] = gapArr[g];
                        }
                        const int idx = R->nsites++;
                        if (S->perfect) {
                            const int overlapsPrev = P && P->chrom == S->chrom && P->strand == S->strand && (S->start <= P->stop && S->stop >= P->start);
                            if (!P || !P->perfect || !overlapsPrev) {
                                perfectsFound++;
                                if (quitAfterTwoPerfects && perfectsFound >= 2) { prev = idx; quit = 1; }
                            }
                        }
                        prev = idx;
                    }
                }
            }
        }
        if (quit) break;
        int t2, ret = 0;
        while ((t2 = heap_peek(c, numHits)) >= 0 && c->values[t2] == site) {
            const int row = c->rows[t2] + 1, col = t2;
            if (row < c->stopsA[col]) {
                c->rows[col] = row;
                c->values[col] = site_minus_offset(c, b->sites[row], offsets[col], baseChrom);
            } else {
                c->active[col] = 0;
                if (heap_size(c, numHits) < approxHitsCutoff) { ret = 1; break; }
            }
        }
        if (ret) break;
    }
    bestScores[0] = imax(bestScores[0], currentTopScore);
    bestScores[1] = imax(bestScores[1], maxHits);
    bestScores[2] = imax(bestScores[2], qcutoff);
    bestScores[3] = imax(bestScores[3], bestqscore);
    bestScores[4] = maxQuickScore;
    bestScores[5] = perfectsFound;
    (void)prevIdx;
}

/* ---------------- Solver (greedy removal of the least useful hit list) ---------------- */
static int64_t value_of_element(const ctx_t* c, const int* offsets, int noffsets, const int* lengths, float keyWeight, int chunk,
                                const int* lists, int numlists, int index) {
    const int64_t POINTS_PER_LIST = 30000, POINTS_PER_BASE1 = 6000, BONUS_END = 40000, POINTS_WIDTH = 5500, MULT_SPACING = -30;
    const int64_t POINTS_PER_SITE = c->X->cfg->points_per_site;
    if (numlists < 1) return 0;
    const int prospect = lists[index];
    if (lengths[prospect] == 0) return -999999;
    int64_t valuep = POINTS_PER_LIST + (POINTS_PER_LIST * 2 / numlists) + ((POINTS_PER_LIST * 10) / lengths[prospect]);
    const int64_t valuem = POINTS_PER_SITE * lengths[prospect];
    if (prospect == 0 || prospect == noffsets - 1) valuep += BONUS_END;
    if (numlists == 1) { valuep += (POINTS_WIDTH + POINTS_PER_BASE1) * chunk; return ((int64_t)((float)valuep * keyWeight)) + valuem; }
    const int first = lists[0], last = lists[numlists - 1];
    const int offL = (prospect == first ? -1 : offsets[lists[index - 1]]);
    const int offP = offsets[prospect];
    const int offR = (prospect == last ? offsets[noffsets - 1] + 1 : offsets[lists[index + 1]]);
    const int oldLeftSpace = offP - offL, oldRightSpace = offR - offP, newSpace = offR - offL;
    const int64_t spaceScore = (int64_t)((oldLeftSpace * oldLeftSpace + oldRightSpace * oldRightSpace) - (newSpace * newSpace)) * MULT_SPACING;
    valuep += spaceScore;
    int uniquelyCovered;
    if (prospect == first) uniquelyCovered = offR - offP;
    else if (prospect == last) uniquelyCovered = offP - offL;
    else { const int a = offL + chunk, bb = offR - a; uniquelyCovered = (bb > 0 ? bb : 0); }
    if (prospect == first || prospect == last) valuep += (POINTS_PER_BASE1 + POINTS_WIDTH) * uniquelyCovered;
    else valuep += POINTS_PER_BASE1 * uniquelyCovered;
    return ((int64_t)((float)valuep * keyWeight)) + valuem;
}
static void find_worst_greedy(const ctx_t* c, const int* offsets, int noffsets, const int* lengths, const float* weights, int chunk,
                              const int* lists, int numlists, int* r) {
    const int64_t EARLY = -50LL * 2000;       /* Solver.EARLY_TERMINATION_SCORE, fixed at class-load time */
    int64_t min = INT64_MAX; int worstIndex = -1;
    for (int i = 0; i < numlists; i++) {
        const int64_t value = value_of_element(c, offsets, noffsets, lengths, weights[i], chunk, lists, numlists, i);
        if (value < min) {
            if (min < EARLY && i != 0) { r[0] = i; r[1] = (int)(value < INT32_MIN ? INT32_MIN : value > INT32_MAX ? INT32_MAX : value); return; }
            min = value; worstIndex = i;
        }
    }
    r[0] = worstIndex; r[1] = (int)(min < INT32_MIN ? INT32_MIN : min > INT32_MAX ? INT32_MAX : min);
}
static int trim_by_greedy(ctx_t* c, const int* offsets, const int* keyScores, int n, int maxHitLists, int* keys) {
    const int* hist = c->X->hist; const orc_index_cfg* g = c->X->cfg;
    float keyWeights[MAXK]; int lengths[MAXK], lists[MAXK];
    const float inv = 1.f / c->baseKeyHitScore;
    for (int i = 0; i < n; i++) keyWeights[i] = keyScores[i] * inv;
    const int limit = imax(SMALL_GENOME_LIST, hist[g->max_average_list_to_search]) * n;
    const int limit2 = imax(SMALL_GENOME_LIST, hist[g->max_average_list_to_search2]);
    const int limit3 = imax(SMALL_GENOME_LIST, hist[g->max_shortest_list_to_search]);
    int sum = 0, initialHitCount = 0, shortest = INT32_MAX - 1, shortest2 = INT32_MAX;
    for (int i = 0; i < n; i++) {
        const int x = count_key(c, keys[i]);
        lengths[i] = x; sum += x; initialHitCount += (x == 0 ? 0 : 1);
        if (x > 0 && x < shortest2) { shortest2 = x; if (shortest2 < shortest) { shortest2 = shortest; shortest = x; } }
    }
    if (initialHitCount < 1) return initialHitCount;
    if (shortest > limit3) { for (int i = 0; i < n; i++) keys[i] = -1; return 0; }
    int hitsCount = initialHitCount;
    while (hitsCount >= 1 && (sum > limit || sum / initialHitCount > limit2 || hitsCount > maxHitLists)) {
        for (int i = 0, j = 0; j < hitsCount; i++) if (lengths[i] > 0) lists[j++] = i;
        int r[2];
        find_worst_greedy(c, offsets, n, lengths, keyWeights, c->K, lists, hitsCount, r);
        const int worst = lists[r[0]], worstValue = r[1];
        sum -= lengths[worst];
        if (worstValue > 0 || lengths[worst] < SMALL_GENOME_LIST) return hitsCount;
        hitsCount--; lengths[worst] = 0; keys[worst] = -1;
    }
    return hitsCount;
}

static int count_hits(const ctx_t* c, int* keys, int n, int maxLen) {
    int numHits = 0;
    for (int i = 0; i < n; i++) {
        const int key = keys[i];
        if (key >= 0) { const int len = count_key(c, key); if (len > 0 && len < maxLen) numHits++; else keys[i] = -1; }
    }
    return numHits;
}
static int shrink2(int* offsets, int* keys, int* keyScores, int n) {
    int j = 0;
    for (int i = 0; i < n; i++) if (keys[i] >= 0) { offsets[j] = offsets[i]; keys[j] = keys[i]; keyScores[j] = keyScores[i]; j++; }
    return j;
}

static int8_t comp_base(int8_t b) {     /* baseToComplementExtended for the bytes a validated read can hold */
    switch (b) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'U': return 'A'; case 'N': return 'N';
                 case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'u': return 'a'; case 'n': return 'n'; default: return b; }
}

/* ---------------- BBIndex.find for one read ---------------- */
void orc_search_read(const orc_search_index* X, const int8_t* basesP, int len, const int8_t* baseScoresP, const int32_t* offsetsIn,
                     const int32_t* keyScoresIn, int nkeys, int quitAfterTwoPerfects, orc_read_result* R) {
    ctx_t* c = (ctx_t*)calloc(1, sizeof(ctx_t));
    const orc_index_cfg* g = X->cfg;
    c->X = X; c->K = g->keylen; c->baseKeyHitScore = BASE_HIT_SCORE * c->K;
    c->indelPenalty = (c->baseKeyHitScore / 2) - 1;
    c->maxPenaltyMisaligned = c->baseKeyHitScore - (1 + c->baseKeyHitScore / 8);
    c->scoreZ1Key = Z_SCORE_MULT * c->K;
    c->shift = g->shift_length; c->cpb = g->chroms_per_block; c->lowMask = c->cpb - 1; c->highMask = ~c->lowMask;
    c->siteMask = (int)(0xFFFFFFFFu >> (g->chrombits + 1));
    memset(R, 0, sizeof(*R));
    const int K = c->K, obeyLimits = 1;
    int offsetsP[MAXK], keysP[MAXK], keyScoresP[MAXK], keysOriginal[MAXK];
    if (nkeys < 1 || nkeys > MAXK || len > ORC_MAX_READ) { R->status = ORC_ST_BADARG; free(c); return; }
    int n = nkeys;
    for (int i = 0; i < n; i++) {
        offsetsP[i] = offsetsIn[i]; keyScoresP[i] = keyScoresIn[i];
        int key = 0, bad = 0;
        for (int p = offsetsIn[i]; p < offsetsIn[i] + K; p++) {
            const int8_t ch = basesP[p]; int x;
            switch (ch) { case 'A': case 'a': x = 0; break; case 'C': case 'c': x = 1; break; case 'G': case 'g': x = 2; break; case 'T': case 't': case 'U': case 'u': x = 3; break; default: x = -1; }
            if (x < 0) { bad = 1; break; }
            key = (key << 2) | x;
        }
        keysOriginal[i] = bad ? -1 : key; keysP[i] = keysOriginal[i];
    }
    const int numKeysOriginal = n;
    const int maxLen = g->max_usable_length;
    int numHits = count_hits(c, keysP, n, maxLen);
    if (numHits > 0) {
        const int trigger = (3 * n) / 4;
        if (numHits < 4 && numHits < trigger) { memcpy(keysP, keysOriginal, sizeof(int) * n); numHits = count_hits(c, keysP, n, (maxLen * 3) / 2); }
        if (numHits < 3 && numHits < trigger) { memcpy(keysP, keysOriginal, sizeof(int) * n); numHits = count_hits(c, keysP, n, maxLen * 2); }
        if (numHits < 3 && numHits < trigger) { memcpy(keysP, keysOriginal, sizeof(int) * n); numHits = count_hits(c, keysP, n, maxLen * 3); }
        if (numHits < 2 && numHits < trigger) { memcpy(keysP, keysOriginal, sizeof(int) * n); numHits = count_hits(c, keysP, n, maxLen * 5); }
    }
    if (numHits < n) n = shrink2(offsetsP, keysP, keyScoresP, n);
    if (n > 0) {     /* TRIM_BY_GREEDY && obeyLimits */
        const int maxLists = imax((int)(HIT_FRACTION_TO_RETAIN * n), MIN_HIT_LISTS_TO_RETAIN);
        numHits = trim_by_greedy(c, offsetsP, keyScoresP, n, maxLists, keysP);
    }
    R->num_hits = numHits;
    if (numHits < 1) { R->status = c->status; free(c); return; }
    if (numHits < n) n = shrink2(offsetsP, keysP, keyScoresP, n);
    /* minus strand */
    int offsetsM[MAXK], keysM[MAXK], keyScoresM[MAXK];
    int8_t basesM[ORC_MAX_READ], baseScoresM[ORC_MAX_READ];
    for (int i = 0; i < n; i++) { offsetsM[i] = len - (offsetsP[n - 1 - i] + K); keysM[i] = orc_rcomp_key_fast(keysP[n - 1 - i], K); keyScoresM[i] = keyScoresP[n - 1 - i]; }
    for (int i = 0; i < len; i++) { basesM[i] = comp_base(basesP[len - 1 - i]); baseScoresM[i] = baseScoresP[len - 1 - i]; }
    const int maxQuickScore = max_quick_score(c, offsetsP, keyScoresP, n);
    int bestScores[6] = {0, 0, 0, 0, 0, 0};
    const int prescan_qscore = (numHits >= 5);
    int precounts[2 * ORC_MAX_BLOCKS], prescores[2 * ORC_MAX_BLOCKS]; int havePre = 0;
    int hitsCutoff = 0, qscoreCutoff = (int)(MIN_QSCORE_MULT * maxQuickScore);
    int allBasesCovered = 1;
    if (offsetsP[0] != 0) allBasesCovered = 0;
    else if (offsetsP[n - 1] != (len - K)) allBasesCovered = 0;
    else for (int i = 1; i < n; i++) if (offsetsP[i] > offsetsP[i - 1] + K) { allBasesCovered = 0; break; }
    const int pretend = (allBasesCovered || n >= numKeysOriginal - 4 || (n >= 9 && (offsetsP[n - 1] - offsetsP[0] + K) > imax(40, (int)(len * .75f))));
    const int nchroms = X->nchroms, minChrom = 1, maxChrom = nchroms;
    int st[MAXK], sp[MAXK], of[MAXK], ks[MAXK];
    if (prescan_qscore) {
        /* prescanAllBlocks */
        int bestqscore = 0, maxHits = 0, minHitsToScore = 1, cycle = 0, early = 0;
        const int ncyc = 2 * X->nblocks;
        for (int i = 0; i < ncyc; i++) { precounts[i] = n; prescores[i] = maxQuickScore; }
        havePre = 1;
        for (int chrom = minChrom; chrom <= maxChrom && !early; chrom = ((chrom & c->highMask) + c->cpb)) {
            for (int pmi = 0; pmi < 2 && !early; pmi++, cycle++) {
                const int* keys = pmi == 0 ? keysP : keysM;
                memcpy(of, pmi == 0 ? offsetsP : offsetsM, sizeof(int) * n); memcpy(ks, pmi == 0 ? keyScoresP : keyScoresM, sizeof(int) * n);
                int nh = get_hits(c, keys, n, chrom, st, sp);
                if (nh < minHitsToScore) { prescores[cycle] = -9999; precounts[cycle] = 0; }
                else {
                    if (nh < n) nh = shrink_hits(st, sp, of, ks, n);
                    int ts, th;
                    find_max_qscore2(c, st, sp, of, ks, nh, chrom, minHitsToScore, 1, bestqscore >= maxQuickScore && pretend, &ts, &th);
                    prescores[cycle] = ts; precounts[cycle] = th;
                    bestqscore = imax(ts, bestqscore); maxHits = imax(maxHits, th);
                    if (bestqscore >= maxQuickScore && pretend) { minHitsToScore = imax(minHitsToScore, maxHits); early = 1; }
                }
            }
        }
        bestScores[1] = imax(bestScores[1], maxHits); bestScores[3] = imax(bestScores[3], bestqscore);
        if (bestScores[1] < 1) { R->status = c->status; free(c); return; }
        if ((float)bestScores[3] < maxQuickScore * MIN_QSCORE_MULT2) { R->status = c->status; free(c); return; }
        if (bestScores[3] >= maxQuickScore && pretend) {
            hitsCutoff = approx_hits_cutoff(c, n, bestScores[1], 1, 1);
            qscoreCutoff = imax(qscoreCutoff, (int)(bestScores[3] * DYNAMIC_QSCORE_THRESH_PERFECT));
        } else {
            hitsCutoff = approx_hits_cutoff(c, n, bestScores[1], 1, 0);
            qscoreCutoff = imax(qscoreCutoff, (int)(bestScores[3] * PRESCAN_QSCORE_THRESH));
        }
    }
    int maxScore = 70 + (len - 1) * 100;
    for (int i = 0; i < len; i++) maxScore += baseScoresP[i];
    int fullyDefined = 1;
    for (int i = 0; i < len; i++) { const int8_t ch = basesP[i]; if (!(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'U' || ch == 'a' || ch == 'c' || ch == 'g' || ch == 't' || ch == 'u')) { fullyDefined = 0; break; } }
    R->max_score = maxScore; R->max_quick_score = maxQuickScore;
    int cycle = 0;
    for (int chrom = minChrom; chrom <= maxChrom; chrom = ((chrom & c->highMask) + c->cpb)) {
        for (int strand = 0; strand < 2; strand++) {
            if (!havePre || precounts[cycle] >= hitsCutoff || prescores[cycle] >= qscoreCutoff) {
                const int* keys = strand == 0 ? keysP : keysM;
                memcpy(of, strand == 0 ? offsetsP : offsetsM, sizeof(int) * n); memcpy(ks, strand == 0 ? keyScoresP : keyScoresM, sizeof(int) * n);
                const int nh = get_hits(c, keys, n, chrom, st, sp);
                if (nh >= 1)
                    slow_walk3(c, st, sp, strand == 0 ? basesP : basesM, strand == 0 ? baseScoresP : baseScoresM, len, ks, of, n, chrom, strand,
                               obeyLimits, R, bestScores, allBasesCovered, maxScore, fullyDefined, quitAfterTwoPerfects, 0);
            }
            cycle++;
            if (quitAfterTwoPerfects && bestScores[5] >= 2) goto done;
        }
    }
done:
    for (int i = 0; i < 6; i++) R->best_scores[i] = bestScores[i];
    R->status = c->status;
    free(c);
}

void orc_search_batch(const orc_search_index* X, const int8_t* bases, const int8_t* baseScores, const int64_t* read_off, int64_t nreads,
                      const int32_t* nkeys, const int32_t* offsets, const int32_t* keyScores, int32_t maxKeys, int quitAfterTwoPerfects,
                      orc_read_result* results) {
    for (int64_t r = 0; r < nreads; r++) {
        const int64_t o = read_off[r]; const int len = (int)(read_off[r + 1] - o);
        if (nkeys[r] <= 0) { memset(&results[r], 0, sizeof(orc_read_result)); continue; }
        orc_search_read(X, bases + o, len, baseScores + o, offsets + r * maxKeys, keyScores + r * maxKeys, nkeys[r], quitAfterTwoPerfects, &results[r]);
    }
}
