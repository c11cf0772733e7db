// search.cu — BBIndex.find on the GPU (north-star kernel 2): seeds -> candidate sites.
//
// Per read, the reference's own phase order (current/align2/BBIndex.java:403-639):
//   key filtering      countHits with up to four relaxations (:376-440), shrink2, trimExcessHitListsByGreedy (:266-350) with
//                      Solver.findWorstGreedy/valueOfElement (current/align2/Solver.java:48-152)
//   prescanAllBlocks   (:642-741) -> findMaxQscore2 (:2294-2450): merge of the hit lists per block and strand, quickScore
//                      (:2490-2511) + scoreZ2 (:2882-2914) only -> per-cycle hit / qscore cut-offs
//   slowWalk3          (:1219-1706) per block, plus strand first: the same merge with dynamic cut-offs, extendScore
//                      (:2558-2757) + MSA.calcAffineScore (MultiStateAligner11tsJNI.java:871-941), subsumption of overlapping
//                      sites, SiteScore.setPerfect (current/stream/SiteScore.java:239-292), makeGapArray (:2837-2878)
// bestScores[6] carries over between blocks and strands exactly as in the reference (:540-550).
//
// Round-1 layout: one thread per read, its working arrays in local memory; the reference's QuadHeap (ordered by
// (site, column), Quad.java:18-22 — a total order) is an arg-min over the <=96 active columns, which visits the same
// (site, column) sequence.  Index (starts/sites/COUNTS) and chromosome bytes are gathered straight from HBM/L2.
// Unsupported corner (flagged in `status`, never silently wrong): subsumption into a previous site that carries a gap
// array (needs GapTools.fixGaps).
#include <cstring>
#include <cstdio>
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

constexpr int MAX_INDEL = 16000, MAX_INDEL2 = 32000, MINGAP = 256, GAPLEN = 128, GAPBUFFER2 = 128;
constexpr int Y_SCORE_MULT = 10, Z_SCORE_MULT = 20, BASE_HIT_SCORE = 100, INDEL_PENALTY_MULT = 20;
constexpr int MIN_HIT_LISTS_TO_RETAIN = 6, SMALL_GENOME_LIST = 20, MAXK = 96, SEARCH_MAX_READ = 608, SEARCH_MAX_BLOCKS = 64;
constexpr float MIN_SCORE_MULT = 0.15f, MIN_QSCORE_MULT = 0.025f, MIN_QSCORE_MULT2 = 0.1f, DYNAMIC_SCORE_THRESH = 0.84f,
                DYNAMIC_QSCORE_THRESH = 0.6f, DYNAMIC_QSCORE_THRESH_PERFECT = 0.8f, HIT_FRACTION_TO_RETAIN = 0.85f;
constexpr float PRESCAN_QSCORE_THRESH = DYNAMIC_QSCORE_THRESH * .95f;

struct SearchBlock { const int* starts; const int* sites; };
struct SearchIndex {
    const bbm_index_cfg* cfg; const SearchBlock* blocks; int nblocks; int nchroms;
    const int* counts; const int* hist; const int8_t* chroms; const long long* chrom_off;
};
struct ReadState {           // where this read's sites go
    bbm_site* sites; int maxSites; int nsites;
};
// Per-read working arrays of the heap walks, one slot per hit list: current (site - offset) value, current / end index into the
// block's sites array (a list is exhausted when row == stop), key offset and key score.  For batches with at most 32 keys per read
// they live in shared memory, laid out [slot][thread] (bank = thread: conflict-free whatever slot each lane touches); otherwise in a
// per-thread global pool.  `S` is the slot stride in elements.
// (a ring in shared memory was measured slower: this kernel lives on L1 hits of its thread-local arrays, and every KB of shared memory
// is taken from the L1 carve-out)
constexpr int RING_KEYS = 32, RING_CAP = 2 * RING_KEYS, RING_STRIDE = 1;
struct ctx_t {
    const SearchIndex* X;
    int K, baseKeyHitScore, indelPenalty, maxPenaltyMisaligned, scoreZ1Key;
    int siteMask, shift, lowMask, highMask, cpb;
    int S;
    int* values; int* rows; int* stops; short* of; short* ks;
    int* locArray;
    int* ringV; signed char* ringC;           // [RING_CAP] in the thread's pool block: sorted (value, column) view of the live heads (singleton fast path)
    int status;
    long long tExtend, tPrescan, tWalk, tFilter;      // clock64 per phase (only summed when SearchParams.prof is set)
};
#define VAL(i)  c->values[(i) * c->S]
#define ROW(i)  c->rows[(i) * c->S]
#define STOP(i) c->stops[(i) * c->S]
#define OFS(i)  ((int)c->of[(i) * c->S])
#define KSC(i)  ((int)c->ks[(i) * c->S])
#define RV(p)   c->ringV[(p) * RING_STRIDE]
#define RC(p)   c->ringC[(p) * RING_STRIDE]

__device__ __forceinline__ int rcomp_fast_dev(int kmer, int k) {     // AminoAcid.reverseComplementBinaryFast (dna/AminoAcid.java:258-271)
    int out = 0;
    const int extra = k & 3;
    for (int i = 0; i < extra; ++i) { out = (out << 2) | ((~kmer) & 3); kmer >>= 2; }
    k -= extra;
    for (int i = 0; i < k; i += 4) {
        int b = kmer & 0xFF, r = 0;
        for (int j = 0; j < 4; ++j) { r = (r << 2) | ((~b) & 3); b >>= 2; }
        out = (out << 8) | (int)(short)r;
        kmer >>= 8;
    }
    return out;
}

__device__ __forceinline__ int absdif(int a, int b) { return a > b ? a - b : b - a; }



__device__ __forceinline__ int count_key(const ctx_t* c, int key) { return c->X->counts[key]; }
__device__ __forceinline__ int to_number(const ctx_t* c, int site, int chrom) { return ((chrom & c->lowMask) << c->shift) | site; }
__device__ __forceinline__ int number_to_chrom(const ctx_t* c, int number, int baseChrom) { return (int)((uint32_t)number >> c->shift) + (baseChrom & c->highMask); }
__device__ __forceinline__ int number_to_site(const ctx_t* c, int number) { return number & c->siteMask; }
__device__ __forceinline__ int base_chrom(const ctx_t* c, int chrom) { return imax(0, chrom & c->highMask); }
__device__ __forceinline__ const SearchBlock* block_of(const ctx_t* c, int chrom) { return &c->X->blocks[((chrom & c->highMask) - (1 & c->highMask)) / c->cpb]; }   // blocks[0] holds chromosome 1 (index[baseChrom(1)] in the reference)

__device__ int block_length(const SearchBlock* b, int key) {        /* Block.length(key), Block.java:62-66 */
    const int x = b->starts[key + 1] - b->starts[key];
    if (x == 0) return 0;
    return b->sites[b->starts[key]] != -1 ? x : 0;
}

/* ---------------- calcApproxHitsCutoff (BBIndex.java:3267-3294; not perfect/semiperfect mode) ---------------- */
__device__ int approx_hits_cutoff(const ctx_t* c, int keys, int hits, int currentCutoff, int perfect) {
    const bbm_index_cfg* g = c->X->cfg;
    const int mahtk = 1;
    const int reduction = imin(imax(hits / g->hit_reduction_div, g->max_hits_reduction2), imax(g->maximum_max_hits_reduction, keys / 8));
    int r = hits - reduction;
    r = imax(mahtk, imax(currentCutoff, r));
    if (perfect) r = imax(r, keys - 0);
    return r;
}

__device__ int max_score_z(const ctx_t* c, const int* offsets, int n) {
    int score = 0, a0 = -1, b0 = -1;
    for (int i = 0; i < n; i++) { const int a = offsets[i]; if (b0 < a) { score += b0 - a0; a0 = a; } b0 = a + c->K; }
    score += b0 - a0;
    return score * Z_SCORE_MULT;
}
__device__ int max_quick_score(const ctx_t* c, const int* offsets, const int* keyScores, int n) {
    int x = 0;
    for (int i = 0; i < n; i++) x += keyScores[i];
    const int y = Y_SCORE_MULT * (offsets[n - 1] - offsets[0]);
    x += max_score_z(c, offsets, n);
    return x + y;
}

__device__ int score_right(const ctx_t* c, int centerIndex, int numHits) {
    int score = 0, prev, loc = VAL(centerIndex);
    for (int i = centerIndex + 1; i < numHits; i++) {
        const int v = VAL(i);
        if (v >= 0) {
            prev = loc; loc = v;
            const int offset = absdif(loc, prev);
            if (offset <= MAX_INDEL) {
                score += KSC(i);
                if (offset != 0) score -= imin(c->indelPenalty + INDEL_PENALTY_MULT * offset, c->maxPenaltyMisaligned);
            } else loc = prev;
        }
    }
    return score;
}
__device__ int score_left(const ctx_t* c, int centerIndex) {
    int score = 0, prev, loc = VAL(centerIndex);
    for (int i = centerIndex - 1; i >= 0; i--) {
        const int v = VAL(i);
        if (v >= 0) {
            prev = loc; loc = v;
            const int offset = absdif(loc, prev);
            if (offset <= MAX_INDEL) {
                score += KSC(i);
                if (offset != 0) score -= imin(c->indelPenalty + INDEL_PENALTY_MULT * offset, c->maxPenaltyMisaligned);
            } else loc = prev;
        }
    }
    return score;
}
__device__ int score_y(const ctx_t* c, int centerIndex, int n) {
    const int center = VAL(centerIndex);
    int rightIndex = -1;
    for (int i = n - 1; rightIndex < centerIndex; i--) if (VAL(i) == center) rightIndex = i;
    return OFS(rightIndex) - OFS(centerIndex);
}
__device__ int quick_score(const ctx_t* c, int centerIndex, int numApproxHits, int numHits) {
    if (numApproxHits == 1) return KSC(centerIndex);
    const int x = KSC(centerIndex) + score_left(c, centerIndex) + score_right(c, centerIndex, numHits) - centerIndex;
    const int y = Y_SCORE_MULT * score_y(c, centerIndex, numHits);
    return x + y;
}
__device__ int score_z2(const ctx_t* c, int centerIndex, int numApproxHits, int numHits) {
    if (numApproxHits == 1) return c->scoreZ1Key;
    const int center = VAL(centerIndex);
    const int maxLoc = center + MAX_INDEL2, minLoc = imax(0, center - MAX_INDEL);
    int score = 0, a0 = -1, b0 = -1;
    for (int i = 0; i < numHits; i++) {
        const int loc = VAL(i);
        if (loc >= minLoc && loc <= maxLoc) { const int a = OFS(i); if (b0 < a) { score += b0 - a0; a0 = a; } b0 = a + c->K; }
    }
    score += b0 - a0;
    return score * Z_SCORE_MULT;
}
// maxQuickScore over the ctx arrays (offsets / key scores of the current strand, before shrinking)
__device__ int max_quick_score_ctx(const ctx_t* c, int n) {
    int x = 0, score = 0, a0 = -1, b0 = -1;
    for (int i = 0; i < n; i++) { x += KSC(i); const int a = OFS(i); if (b0 < a) { score += b0 - a0; a0 = a; } b0 = a + c->K; }
    score += b0 - a0;
    return x + score * Z_SCORE_MULT + Y_SCORE_MULT * (OFS(n - 1) - OFS(0));
}

/* ---------------- MSA.calcAffineScore(locArray, baseScores, bases) ---------------- */
__device__ int calc_affine_score(const int* locArray, const int8_t* baseScores, int len) {
    const int INSC[6] = {0, -395, -434, -473, -512, -551};     /* POINTS_INS_ARRAY_C[0..5] */
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
__device__ int extend_score(ctx_t* c, const int8_t* bases, const int8_t* baseScores, int len,
                        int chrom, int centerIndex, int* locArray, int numHits) {
    const int centerVal = VAL(centerIndex), centerLoc = number_to_site(c, centerVal);
    const int minVal = centerVal - MAX_INDEL, maxVal = centerVal + MAX_INDEL2;
    const int8_t* ref = c->X->chroms + c->X->chrom_off[chrom - 1];
    const int refLen = (int)(c->X->chrom_off[chrom] - c->X->chrom_off[chrom - 1]);
    const int K = c->K;
    for (int i = 0; i < len; i++) locArray[i] = -1;
    for (int i = 0, keynum = 0; i < numHits; i++) {
        const int value = VAL(i);
        if (value >= minVal && value <= maxVal) {
            const int refbase = number_to_site(c, value);
            keynum++;
            const int callbase = OFS(i);
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
        const int value = VAL(i);
        if (value >= minVal && value <= maxVal) {
            const int refbase = number_to_site(c, value);
            const int callbase = OFS(i);
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


/* makeGapArray (BBIndex.java:2837-2878); destroys locArray.  Returns #ints written (0 = null). */
__device__ int make_gap_array(int* locArray, int len, int minLoc, int minGap, int* out, int cap, int* overflow) {
    int gaps = 0, doSort = 0;
    if (locArray[0] < 0) locArray[0] = minLoc;
    for (int i = 1; i < len; i++) {
        if (locArray[i] < 0) locArray[i] = locArray[i - 1] + 1; else locArray[i] += i;
        if (locArray[i] < locArray[i - 1]) doSort = 1;
    }
    if (doSort) { for (int i = 1; i < len; i++) { const int v = locArray[i]; int j = i - 1; while (j >= 0 && locArray[j] > v) { locArray[j + 1] = locArray[j]; j--; } locArray[j + 1] = v; } }
    for (int i = 1; i < len; i++) if (locArray[i] - locArray[i - 1] > minGap) gaps++;
    if (gaps < 1) return 0;
    const int n = 2 + gaps * 2;
    if (n > cap) { *overflow = 1; return 0; }
    out[0] = locArray[0]; out[n - 1] = locArray[len - 1];
    for (int i = 1, j = 1; i < len; i++) if (locArray[i] - locArray[i - 1] > minGap) { out[j] = locArray[i - 1]; out[j + 1] = locArray[i]; j += 2; }
    return n;
}

/* SiteScore.setPerfect(bases) */
__device__ void set_perfect(const ctx_t* c, bbm_site* s, const int8_t* bases, int len) {
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
// getHits (BBIndex.java:353-373) straight into the ctx arrays: ROW/STOP = start/stop of the key's list in this block (-1 = none),
// OFS/KSC = the strand's offsets and key scores.
__device__ int get_hits(ctx_t* c, const int* keys, const int* offsets, const int* keyScores, int n, int chrom) {
    int numHits = 0;
    const SearchBlock* b = block_of(c, chrom);
    for (int i = 0; i < n; i++) {
        const int key = keys[i];
        int st = -1, sp = -1;
        if (key >= 0) {
            const int len = count_key(c, key);
            if (len > 0) {                                 /* maxLen = Integer.MAX_VALUE */
                const int len2 = block_length(b, key);
                if (len2 > 0) { st = b->starts[key]; sp = st + len2; numHits++; }
            }
        }
        ROW(i) = st; STOP(i) = sp;
        c->of[i * c->S] = (short)offsets[i]; c->ks[i * c->S] = (short)keyScores[i];
    }
    return numHits;
}
__device__ int shrink_hits(ctx_t* c, int n) {
    int j = 0;
    for (int i = 0; i < n; i++) if (ROW(i) >= 0) { if (j != i) { ROW(j) = ROW(i); STOP(j) = STOP(i); c->of[j * c->S] = c->of[i * c->S]; c->ks[j * c->S] = c->ks[i * c->S]; } j++; }
    return j;
}

/* translate a raw site to (site - offset), clamped at the chromosome start (BBIndex.java:1305-1313 etc.) */
__device__ __forceinline__ int site_minus_offset(const ctx_t* c, int a, int offset, int baseChrom) {
    if ((a & c->siteMask) >= offset) return a - offset;
    const int ch = number_to_chrom(c, a, baseChrom), st = number_to_site(c, a);
    return to_number(c, imax(st - offset, 0), ch);
}

/* The heap (QuadHeap ordered by (site, column), Quad.java:18-22) is replaced by an arg-min over the live columns (row < stop): the
 * order is total, so the sequence of (site, column) visited is identical. */
__device__ int heap_peek(const ctx_t* c, int numHits) {
    int best = -1, bestVal = 0;
    for (int i = 0; i < numHits; i++) {
        if (ROW(i) < STOP(i)) { const int v = VAL(i); if (best < 0 || v < bestVal) { best = i; bestVal = v; } }
    }
    return best;
}

/* ---------------- exact skip-ahead for long hit lists ----------------
 * Both walks pop the globally smallest (site, column) one at a time and first count how many list heads lie in
 * [site - MAX_INDEL, site + MAX_INDEL2]; a site with fewer than approxHitsCutoff such heads changes nothing but the heap
 * (BBIndex.java:1340-1365, 2352-2376).  All live heads are >= site, heads only grow, and an exhausted list keeps its last value, which
 * only falls out of range as the walk moves on.  So if no exhausted list is in range now and h_k is the k-th smallest live head, every
 * site below T = h_k - MAX_INDEL2 sees fewer than k heads in range: with k = approxHitsCutoff those sites are no-ops and the lists can be
 * advanced to T by binary search instead of one pop at a time.  On a human-sized index (hit lists ~46 long, 18 keys, 2 blocks) this
 * removes ~99 % of the heap steps of slowWalk3.  In the prescan the cutoff starts at 1; there k = 2 is used: the skipped sites are then
 * singletons, whose only effect is topQscore = max(topQscore, keyScore + scoreZ1Key) (quickScore/scoreZ2 with one hit, :2490-2492,
 * 2883), which the caller applies for the one list that moved.  Exhaustion during the skip triggers the same exits as a pop would.
 * Returns a bit mask: bit 0 = the caller's exhaustion rule fired, bit 1 = at least one list moved; *movedCol = a list that moved. */
__device__ int skip_ahead(ctx_t* c, const SearchBlock* b, int baseChrom, int numHits, int k, int exitBelow, bool exitOnAnyExhaust,
                          int* nActive, int* movedCol) {
    if (k < 2 || *nActive < k) return 0;
    // smallest live head, and the exhausted lists' stale values
    int s = 0x7fffffff, staleMax = -0x7fffffff - 1;
    for (int i = 0; i < numHits; i++) { const int v = VAL(i); if (ROW(i) < STOP(i)) s = imin(s, v); else staleMax = imax(staleMax, v); }
    if (staleMax >= s - MAX_INDEL) return 0;
    // k-th smallest live head in (value, column) order: k selection passes from below, or nActive-k+1 passes from above if that is fewer
    int hk, hkCol;
    if (k <= *nActive - k + 1) {
        hk = -0x7fffffff - 1; hkCol = -1;
        for (int pass = 0; pass < k; pass++) {
            int best = 0x7fffffff, bestCol = -1;
            for (int i = 0; i < numHits; i++) {
                if (ROW(i) >= STOP(i)) continue;
                const int v = VAL(i);
                if ((v > hk || (v == hk && i > hkCol)) && (v < best || bestCol < 0)) { best = v; bestCol = i; }
            }
            hk = best; hkCol = bestCol;
        }
    } else {
        hk = 0x7fffffff; hkCol = 0x7fffffff;
        for (int pass = 0; pass < *nActive - k + 1; pass++) {
            int best = -0x7fffffff - 1, bestCol = -1;
            for (int i = 0; i < numHits; i++) {
                if (ROW(i) >= STOP(i)) continue;
                const int v = VAL(i);
                if ((v < hk || (v == hk && i < hkCol)) && (v > best || (v == best && i > bestCol) || bestCol < 0)) { best = v; bestCol = i; }
            }
            hk = best; hkCol = bestCol;
        }
    }
    if (hk < -0x40000000 + MAX_INDEL2) return 0;
    const int T = hk - MAX_INDEL2;
    if (T <= s) return 0;
    int flags = 0;
    for (int col = 0; col < numHits; col++) {
        if (ROW(col) >= STOP(col) || VAL(col) >= T) continue;
        // first element of the list whose translated value is >= T (the translation is monotone in the raw site)
        const int ofs = OFS(col);
        int lo = ROW(col) + 1, hi = STOP(col);
        while (lo < hi) { const int mid = lo + ((hi - lo) >> 1); if (site_minus_offset(c, b->sites[mid], ofs, baseChrom) >= T) hi = mid; else lo = mid + 1; }
        flags |= 2; *movedCol = col;
        if (lo < STOP(col)) { ROW(col) = lo; VAL(col) = site_minus_offset(c, b->sites[lo], ofs, baseChrom); }
        else {
            VAL(col) = site_minus_offset(c, b->sites[STOP(col) - 1], ofs, baseChrom);      // the list keeps its last site
            ROW(col) = STOP(col);
            (*nActive)--;
            if (exitOnAnyExhaust || *nActive < exitBelow) return flags | 1;
        }
    }
    return flags;
}

/* ---------------- findMaxQscore2 ---------------- */
__device__ void find_max_qscore2(ctx_t* c, int numHits, int baseChrom_, int prevMaxHits, int earlyExit, int perfectOnly, int* outScore, int* outHits) {
    const int baseChrom = base_chrom(c, baseChrom_);
    const SearchBlock* b = block_of(c, baseChrom_);
    for (int i = 0; i < numHits; i++) VAL(i) = site_minus_offset(c, b->sites[ROW(i)], OFS(i), baseChrom);
    const int maxQuickScore = max_quick_score_ctx(c, numHits);
    int topQscore = -999999999, maxHits = 0, approxHitsCutoff, indelCutoff;
    if (perfectOnly) { approxHitsCutoff = numHits; indelCutoff = 0; }
    else { approxHitsCutoff = imax(prevMaxHits, imin(1, numHits - 1)); indelCutoff = MAX_INDEL2; }
    int t, nActive = numHits;
    long long total = 0;
    for (int i = 0; i < numHits; i++) total += STOP(i) - ROW(i);
    const bool longLists = total >= 4LL * numHits;        // skip-ahead only pays when the lists are long
    int staleMax = -0x7fffffff - 1;          // largest last-site of an exhausted list (it still counts as a hit within MAX_INDEL above it)
    while (true) {
        if (longLists && numHits >= 2 && approxHitsCutoff >= 2) {
            int moved = -1;
            const int f = skip_ahead(c, b, baseChrom, numHits, approxHitsCutoff, approxHitsCutoff, earlyExit && perfectOnly, &nActive, &moved);
            if ((f & 1) && earlyExit) { *outScore = topQscore; *outHits = maxHits; return; }
        } else if (longLists && c->ringV && numHits >= 2 && numHits <= RING_KEYS && nActive >= 2) {
            // Singleton fast path (cutoff 1: every site is scored).  While the two smallest live heads are more than MAX_INDEL2 apart and
            // no exhausted list is in range, the smallest head is a site with exactly one hit: quickScore = its key score, scoreZ2 =
            // scoreZ1Key (BBIndex.java:2490-2492, 2883), approxHits-1 = 0 leaves the cutoff alone.  The live heads are kept sorted in a
            // small ring (insertion from the back).
            int front = 0, back = 0;
            for (int i = 0; i < numHits; i++) {
                if (ROW(i) >= STOP(i)) { staleMax = imax(staleMax, VAL(i)); continue; }
                const int v = VAL(i);
                int p = back++;
                while (p > front && (RV(p - 1) > v || (RV(p - 1) == v && RC(p - 1) > i))) { RV(p) = RV(p - 1); RC(p) = RC(p - 1); p--; }
                RV(p) = v; RC(p) = (signed char)i;
            }
            while (back - front >= 2 && RV(front) <= 0x7fffffff - MAX_INDEL2 && RV(front + 1) > RV(front) + MAX_INDEL2 && (long long)RV(front) - MAX_INDEL > staleMax) {
                const int col = RC(front), s0 = RV(front);
                front++;
                const int q = KSC(col) + c->scoreZ1Key;
                if (q > topQscore) { maxHits = imax(1, maxHits); topQscore = q; }
                const int row = ROW(col) + 1;
                ROW(col) = row;
                if (row < STOP(col)) {
                    const int v = site_minus_offset(c, b->sites[row], OFS(col), baseChrom);
                    VAL(col) = v;
                    if (back == RING_CAP) { const int n = back - front; for (int i = 0; i < n; i++) { RV(i) = RV(front + i); RC(i) = RC(front + i); } front = 0; back = n; }
                    int p = back++;
                    while (p > front && (RV(p - 1) > v || (RV(p - 1) == v && RC(p - 1) > col))) { RV(p) = RV(p - 1); RC(p) = RC(p - 1); p--; }
                    RV(p) = v; RC(p) = (signed char)col;
                } else {
                    nActive--; staleMax = imax(staleMax, s0);          /* VAL(col) keeps its last site */
                    if (earlyExit && nActive < approxHitsCutoff) { *outScore = topQscore; *outHits = maxHits; return; }
                }
            }
        }
        if ((t = heap_peek(c, numHits)) < 0) break;
        const int site = VAL(t), centerIndex = t;
        int approxHits = 0;
        {
            const int minsite = site - imin(MAX_INDEL, indelCutoff), maxsite = site + MAX_INDEL2;
            for (int column = 0, chances = numHits - approxHitsCutoff; column < numHits && chances >= 0; column++) {
                const int x = VAL(column);
                if (x >= minsite && x <= maxsite) approxHits++; else chances--;
            }
        }
        if (approxHits >= approxHitsCutoff) {
            int qscore = quick_score(c, centerIndex, approxHits, numHits);
            qscore += score_z2(c, centerIndex, approxHits, numHits);
            if (qscore > topQscore) {
                maxHits = imax(approxHits, maxHits);
                approxHitsCutoff = imax(approxHitsCutoff, approxHits - 1);
                topQscore = qscore;
                if (qscore >= maxQuickScore && earlyExit) { *outScore = topQscore; *outHits = maxHits; return; }
            }
        }
        // pop every heap entry that sits on `site`.  The heap order is (site, column), so they come out in ascending column order, and a
        // column whose next site is again `site` (clamped at the chromosome start) is popped again at once: one pass over the columns.
        for (int col = 0; col < numHits; col++) {
            while (ROW(col) < STOP(col) && VAL(col) == site) {
                const int row = ROW(col) + 1;
                ROW(col) = row;
                if (row < STOP(col)) VAL(col) = site_minus_offset(c, b->sites[row], OFS(col), baseChrom);
                else {
                    nActive--;          /* VAL(col) keeps its last site, exactly like the reference's valueArray */
                    if (earlyExit && (perfectOnly || nActive < approxHitsCutoff)) { *outScore = topQscore; *outHits = maxHits; return; }
                }
            }
        }
    }
    *outScore = topQscore; *outHits = maxHits;
}

/* ---------------- slowWalk3 ---------------- */
__device__ void slow_walk3(ctx_t* c, const int8_t* bases, const int8_t* baseScores, int len,
                       int numKeys, int baseChrom_, int strand, int obeyLimits, ReadState* R, int* bestScores, int allBasesCovered,
                       int maxScore, int fullyDefined, int quitAfterTwoPerfects) {
    const int maxQuickScore = max_quick_score_ctx(c, numKeys);
    const int numHits = shrink_hits(c, numKeys);
    const int filter_by_qscore = (numKeys >= 5);
    const int minScore = obeyLimits ? (int)(MIN_SCORE_MULT * maxScore) : (int)(MIN_SCORE_MULT * 1.25f * maxScore);
    const int minQuickScore = (int)(MIN_QSCORE_MULT * maxQuickScore);
    const int baseChrom = base_chrom(c, baseChrom_);
    const SearchBlock* b = block_of(c, baseChrom_);
    int currentTopScore = bestScores[0];
    int cutoff = imax(minScore, (int)(currentTopScore * DYNAMIC_SCORE_THRESH));
    int qcutoff = imax(bestScores[2], minQuickScore);
    int bestqscore = bestScores[3], maxHits = bestScores[1], perfectsFound = bestScores[5];
    int approxHitsCutoff = approx_hits_cutoff(c, numKeys, maxHits, 1, currentTopScore >= maxScore);
    if (approxHitsCutoff > numHits) return;
    const int shortCircuit = (allBasesCovered && numKeys == numHits && filter_by_qscore);
    if (currentTopScore >= maxScore) qcutoff = imax(qcutoff, (int)(maxQuickScore * DYNAMIC_QSCORE_THRESH_PERFECT));
    for (int i = 0; i < numHits; i++) VAL(i) = site_minus_offset(c, b->sites[ROW(i)], OFS(i), baseChrom);
    int* locArray = c->locArray;
    int prev = -1;       /* index of prevSS in R->sites (a site made during THIS walk), -1 = null */
    int t, quit = 0, nActive = numHits;
    long long total = 0;
    for (int i = 0; i < numHits; i++) total += STOP(i) - ROW(i);
    const bool longLists = total >= 4LL * numHits;
    while (!quit) {
        if (longLists) {
            int moved = -1;
            if (skip_ahead(c, b, baseChrom, numHits, approxHitsCutoff, approxHitsCutoff, false, &nActive, &moved) & 1) break;
        }
        if ((t = heap_peek(c, numHits)) < 0) break;
        const int site = VAL(t), centerIndex = t;
        int maxNearbySite = site, approxHits = 0;
        {
            const int minsite = site - MAX_INDEL, maxsite = site + MAX_INDEL2;
            for (int column = 0, chances = numHits - approxHitsCutoff; column < numHits && chances >= 0; column++) {
                const int x = VAL(column);
                if (x >= minsite && x <= maxsite) { maxNearbySite = (x > maxNearbySite ? x : maxNearbySite); approxHits++; } else chances--;
            }
        }
        if (approxHits >= approxHitsCutoff) {
            int score;
            int qscore = filter_by_qscore ? quick_score(c, centerIndex, approxHits, numHits) : qcutoff;
            qscore += score_z2(c, centerIndex, approxHits, numHits);
            int mapStart = site, mapStop = maxNearbySite;
            int locArrayValid = 0;
            if (qscore < qcutoff) score = -1;
            else {
                const int chrom = number_to_chrom(c, site, baseChrom);
                if (shortCircuit && qscore == maxQuickScore) score = maxScore;
                else {
                    const long long tx0 = clock64();
                    score = extend_score(c, bases, baseScores, len, chrom, centerIndex, locArray, numHits);
                    c->tExtend += clock64() - tx0;
                    locArrayValid = 1;
                    int mn = 0x7fffffff, mx = (-0x7fffffff-1);
                    for (int i = 0; i < len; i++) { const int x = locArray[i]; if (x > -1) { if (x < mn) mn = x; if (x > mx) mx = x; } }
                    if (mn < 0 || mx < 0) { score = -99999; c->status |= BBM_ST_ANOMALY; }
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
                int gapArr[BBM_MAX_GAPS]; int ngap = 0;   /* BBM_MAX_GAPS-1 ints fit in bbm_site */
                if (site3 - site2 >= MINGAP + len) {
                    int ov = 0;
                    (void)locArrayValid;
                    ngap = make_gap_array(locArray, len, site2, MINGAP, gapArr, BBM_MAX_GAPS - 1, &ov);
                    if (ov) c->status |= BBM_ST_GAP_OVERFLOW;
                    if (ngap > 0) { gapArr[0] = imin(gapArr[0], site2); gapArr[ngap - 1] = imax(gapArr[ngap - 1], site3); }
                }
                const int perfect1 = (score == maxScore && fullyDefined);
                const int chromLen = (int)(c->X->chrom_off[chrom] - c->X->chrom_off[chrom - 1]);
                const int inbounds = (site2 >= 0 && site3 < chromLen);
                int made = -1;
                bbm_site* P = prev >= 0 ? &R->sites[prev] : 0;
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
                        if (P->ngaps) c->status |= BBM_ST_GAPFIX;
                        if (perfect2) {}
                        else if (perfect1) { P->stop = site3; if (!P->perfect) perfectsFound++; P->perfect = P->semiperfect = 1; }
                        else { P->stop = maxStop; set_perfect(c, P, bases, len); }
                        P->score = betterScore;
                    } else if (shortEnough && P->stop == site3 && !P->semiperfect) {
                        if (P->ngaps) c->status |= BBM_ST_GAPFIX;
                        if (perfect2) {}
                        else if (perfect1) { P->start = site2; if (!P->perfect) perfectsFound++; P->perfect = P->semiperfect = 1; }
                        else { P->start = minStart; set_perfect(c, P, bases, len); }
                        P->score = betterScore;
                    } else {
                        made = 1;
                    }
                } else if (inbounds) made = 1;
                if (made > 0) {
                    if (R->nsites >= R->maxSites) { c->status |= BBM_ST_SITE_OVERFLOW; }
                    else {
                        bbm_site* S = &R->sites[R->nsites];
                        { bbm_site z = {}; *S = z; }
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
        int ret = 0;
        for (int col = 0; col < numHits && !ret; col++) {          // same (site, column) pop order as the heap, see findMaxQscore2
            while (ROW(col) < STOP(col) && VAL(col) == site) {
                const int row = ROW(col) + 1;
                ROW(col) = row;
                if (row < STOP(col)) VAL(col) = site_minus_offset(c, b->sites[row], OFS(col), baseChrom);
                else {
                    nActive--;
                    if (nActive < approxHitsCutoff) { ret = 1; break; }
                }
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
}

/* ---------------- Solver (greedy removal of the least useful hit list) ---------------- */
__device__ long long value_of_element(const ctx_t* c, const int* offsets, int noffsets, const int* lengths, float keyWeight, int chunk,
                                const int* lists, int numlists, int index) {
    const long long POINTS_PER_LIST = 30000, POINTS_PER_BASE1 = 6000, BONUS_END = 40000, POINTS_WIDTH = 5500, MULT_SPACING = -30;
    const long long POINTS_PER_SITE = c->X->cfg->points_per_site;
    if (numlists < 1) return 0;
    const int prospect = lists[index];
    if (lengths[prospect] == 0) return -999999;
    long long valuep = POINTS_PER_LIST + (POINTS_PER_LIST * 2 / numlists) + ((POINTS_PER_LIST * 10) / lengths[prospect]);
    const long long valuem = POINTS_PER_SITE * lengths[prospect];
    if (prospect == 0 || prospect == noffsets - 1) valuep += BONUS_END;
    if (numlists == 1) { valuep += (POINTS_WIDTH + POINTS_PER_BASE1) * chunk; return ((long long)((float)valuep * keyWeight)) + valuem; }
    const int first = lists[0], last = lists[numlists - 1];
    const int offL = (prospect == first ? -1 : offsets[lists[index - 1]]);
    const int offP = offsets[prospect];
    const int offR = (prospect == last ? offsets[noffsets - 1] + 1 : offsets[lists[index + 1]]);
    const int oldLeftSpace = offP - offL, oldRightSpace = offR - offP, newSpace = offR - offL;
    const long long spaceScore = (long long)((oldLeftSpace * oldLeftSpace + oldRightSpace * oldRightSpace) - (newSpace * newSpace)) * MULT_SPACING;
    valuep += spaceScore;
    int uniquelyCovered;
    if (prospect == first) uniquelyCovered = offR - offP;
    else if (prospect == last) uniquelyCovered = offP - offL;
    else { const int a = offL + chunk, bb = offR - a; uniquelyCovered = (bb > 0 ? bb : 0); }
    if (prospect == first || prospect == last) valuep += (POINTS_PER_BASE1 + POINTS_WIDTH) * uniquelyCovered;
    else valuep += POINTS_PER_BASE1 * uniquelyCovered;
    return ((long long)((float)valuep * keyWeight)) + valuem;
}
__device__ void find_worst_greedy(const ctx_t* c, const int* offsets, int noffsets, const int* lengths, const float* weights, int chunk,
                              const int* lists, int numlists, int* r) {
    const long long EARLY = -50LL * 2000;       /* Solver.EARLY_TERMINATION_SCORE, fixed at class-load time */
    long long min = 0x7fffffffffffffffLL; int worstIndex = -1;
    for (int i = 0; i < numlists; i++) {
        const long long value = value_of_element(c, offsets, noffsets, lengths, weights[i], chunk, lists, numlists, i);
        if (value < min) {
            if (min < EARLY && i != 0) { r[0] = i; r[1] = (int)(value < (-0x7fffffff-1) ? (-0x7fffffff-1) : value > 0x7fffffff ? 0x7fffffff : value); return; }
            min = value; worstIndex = i;
        }
    }
    r[0] = worstIndex; r[1] = (int)(min < (-0x7fffffff-1) ? (-0x7fffffff-1) : min > 0x7fffffff ? 0x7fffffff : min);
}
__device__ int trim_by_greedy(ctx_t* c, const int* offsets, const int* keyScores, int n, int maxHitLists, int* keys) {
    const int* hist = c->X->hist; const bbm_index_cfg* g = c->X->cfg;
    float keyWeights[MAXK]; int lengths[MAXK], lists[MAXK];
    const float inv = 1.f / c->baseKeyHitScore;
    for (int i = 0; i < n; i++) keyWeights[i] = keyScores[i] * inv;
    const int limit = imax(SMALL_GENOME_LIST, hist[g->max_average_list_to_search]) * n;
    const int limit2 = imax(SMALL_GENOME_LIST, hist[g->max_average_list_to_search2]);
    const int limit3 = imax(SMALL_GENOME_LIST, hist[g->max_shortest_list_to_search]);
    int sum = 0, initialHitCount = 0, shortest = 0x7fffffff - 1, shortest2 = 0x7fffffff;
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

__device__ int count_hits(const ctx_t* c, int* keys, int n, int maxLen) {
    int numHits = 0;
    for (int i = 0; i < n; i++) {
        const int key = keys[i];
        if (key >= 0) { const int len = count_key(c, key); if (len > 0 && len < maxLen) numHits++; else keys[i] = -1; }
    }
    return numHits;
}
__device__ int shrink2(int* offsets, int* keys, int* keyScores, int n) {
    int j = 0;
    for (int i = 0; i < n; i++) if (keys[i] >= 0) { offsets[j] = offsets[i]; keys[j] = keys[i]; keyScores[j] = keyScores[i]; j++; }
    return j;
}

__device__ int8_t comp_base(int8_t b) {     /* baseToComplementExtended for the bytes a validated read can hold */
    switch (b) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; case 'U': return 'A'; case 'N': return 'N';
                 case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; case 'u': return 'a'; case 'n': return 'n'; default: return b; }
}


// ---------------- BBIndex.find for one read (current/align2/BBIndex.java:403-639) ----------------
// `phases`: 1 = key filtering, 2 = prescan, 4 = walk.  7 runs BBIndex.find in one go; the split launches run one phase per kernel and carry
// the state between them in `mid` (int[midStride] per read: header, the filtered key arrays of both strands, the per-cycle prescan
// results), so that all lanes of a warp execute the same phase.
constexpr int MID_HDR = 24;
__device__ void search_read(const SearchIndex* X, const int8_t* basesP, int len, const int8_t* baseScoresP, const int* offsetsIn,
                            const int* keyScoresIn, int nkeys, int quitAfterTwoPerfects, bbm_search_head* H, ReadState* R, ctx_t* c,
                            int8_t* basesM, int8_t* baseScoresM, int phases, int* mid, int MK) {
    const bbm_index_cfg* g = X->cfg;
    c->X = X; c->K = g->keylen; c->baseKeyHitScore = BASE_HIT_SCORE * c->K;
    c->indelPenalty = (c->baseKeyHitScore / 2) - 1;
    c->maxPenaltyMisaligned = c->baseKeyHitScore - (1 + c->baseKeyHitScore / 8);
    c->scoreZ1Key = Z_SCORE_MULT * c->K;
    c->shift = g->shift_length; c->cpb = g->chroms_per_block; c->lowMask = c->cpb - 1; c->highMask = ~c->lowMask;
    c->siteMask = (int)(0xFFFFFFFFu >> (g->chrombits + 1));
    c->status = 0;
    const int K = c->K, obeyLimits = 1;
    int offsetsP[MAXK], keysP[MAXK], keyScoresP[MAXK];
    int offsetsM[MAXK], keysM[MAXK], keyScoresM[MAXK];
    if (nkeys < 1 || nkeys > MAXK || len > SEARCH_MAX_READ) { H->status = BBM_ST_BADARG; if (mid) mid[0] = 0; return; }
    int n = nkeys, numHits = 0, numKeysOriginal = nkeys;
    if (phases & 1) {
        int keysOriginal[MAXK];
        for (int i = 0; i < n; i++) {
            offsetsP[i] = offsetsIn[i]; keyScoresP[i] = keyScoresIn[i];
            int key = 0; bool bad = false;
            for (int p = offsetsIn[i]; p < offsetsIn[i] + K; p++) {        // KeyRing.makeKeys / ChromosomeArray.toNumber
                const int ch = basesP[p];
                const int u = ch & 0xDF;
                const int x = (ch & 0x80) ? -1 : (u == 'A' ? 0 : (u == 'C' ? 1 : (u == 'G' ? 2 : ((u == 'T' || u == 'U') ? 3 : -1))));
                if (x < 0) { bad = true; break; }
                key = (key << 2) | x;
            }
            keysOriginal[i] = bad ? -1 : key; keysP[i] = keysOriginal[i];
        }
        const int maxLen = g->max_usable_length;
        const long long tf0 = clock64();
        numHits = count_hits(c, keysP, n, maxLen);
        if (numHits > 0) {
            const int trigger = (3 * n) / 4;
            if (numHits < 4 && numHits < trigger) { for (int i = 0; i < n; i++) keysP[i] = keysOriginal[i]; numHits = count_hits(c, keysP, n, (maxLen * 3) / 2); }
            if (numHits < 3 && numHits < trigger) { for (int i = 0; i < n; i++) keysP[i] = keysOriginal[i]; numHits = count_hits(c, keysP, n, maxLen * 2); }
            if (numHits < 3 && numHits < trigger) { for (int i = 0; i < n; i++) keysP[i] = keysOriginal[i]; numHits = count_hits(c, keysP, n, maxLen * 3); }
            if (numHits < 2 && numHits < trigger) { for (int i = 0; i < n; i++) keysP[i] = keysOriginal[i]; numHits = count_hits(c, keysP, n, maxLen * 5); }
        }
        if (numHits < n) n = shrink2(offsetsP, keysP, keyScoresP, n);
        if (n > 0) {     // TRIM_BY_GREEDY && obeyLimits
            const int maxLists = imax((int)(HIT_FRACTION_TO_RETAIN * n), MIN_HIT_LISTS_TO_RETAIN);
            numHits = trim_by_greedy(c, offsetsP, keyScoresP, n, maxLists, keysP);
        }
        c->tFilter += clock64() - tf0;
        H->num_hits = numHits;
        if (numHits < 1) { H->status = c->status; if (mid) mid[0] = 0; return; }
        if (numHits < n) n = shrink2(offsetsP, keysP, keyScoresP, n);
        // minus strand: KeyRing.reverseOffsets / reverseComplementKeys
        for (int i = 0; i < n; i++) { offsetsM[i] = len - (offsetsP[n - 1 - i] + K); keysM[i] = rcomp_fast_dev(keysP[n - 1 - i], K); keyScoresM[i] = keyScoresP[n - 1 - i]; }
        if (mid) {
            mid[0] = n; mid[1] = numHits; mid[2] = c->status; mid[10] = 0; mid[11] = 0; mid[12] = 0;
            int* a = mid + MID_HDR;
            for (int i = 0; i < n; i++) { a[i] = keysP[i]; a[MK + i] = keysM[i]; a[2 * MK + i] = offsetsP[i]; a[3 * MK + i] = offsetsM[i]; a[4 * MK + i] = keyScoresP[i]; a[5 * MK + i] = keyScoresM[i]; }
        }
        if (!(phases & 6)) return;
    } else {
        n = mid[0]; numHits = mid[1]; c->status = mid[2];
        if (n < 1) return;
        const int* a = mid + MID_HDR;
        for (int i = 0; i < n; i++) { keysP[i] = a[i]; keysM[i] = a[MK + i]; offsetsP[i] = a[2 * MK + i]; offsetsM[i] = a[3 * MK + i]; keyScoresP[i] = a[4 * MK + i]; keyScoresM[i] = a[5 * MK + i]; }
    }
    // Tools.reverseAndCopy / AminoAcid.reverseComplementBases: the walk scores minus-strand sites against the reverse complement
    if (phases & 4) for (int i = 0; i < len; i++) { basesM[i] = comp_base(basesP[len - 1 - i]); baseScoresM[i] = baseScoresP[len - 1 - i]; }
    const int maxQuickScore = max_quick_score(c, offsetsP, keyScoresP, n);
    int bestScores[6] = {0, 0, 0, 0, 0, 0};
    const bool prescan_qscore = (numHits >= 5);
    int precounts[2 * SEARCH_MAX_BLOCKS], prescores[2 * SEARCH_MAX_BLOCKS]; bool havePre = false;
    int hitsCutoff = 0, qscoreCutoff = (int)(MIN_QSCORE_MULT * maxQuickScore);
    bool allBasesCovered = true;
    if (offsetsP[0] != 0) allBasesCovered = false;
    else if (offsetsP[n - 1] != (len - K)) allBasesCovered = false;
    else for (int i = 1; i < n; i++) if (offsetsP[i] > offsetsP[i - 1] + K) { allBasesCovered = false; break; }
    const bool pretend = (allBasesCovered || n >= numKeysOriginal - 4 || (n >= 9 && (offsetsP[n - 1] - offsetsP[0] + K) > imax(40, (int)(len * .75f))));
    const int minChrom = 1, maxChrom = X->nchroms;
    int* midPre = mid ? mid + MID_HDR + 6 * MK : nullptr;
    if (!(phases & 2)) {
        // prescan results of the earlier launch
        if (mid[9]) { H->status = c->status; return; }
        bestScores[1] = mid[4]; bestScores[3] = mid[5]; hitsCutoff = mid[6]; qscoreCutoff = mid[7]; havePre = mid[8] != 0;
        if (havePre) { const int ncyc = 2 * X->nblocks; for (int i = 0; i < ncyc; i++) { precounts[i] = midPre[i]; prescores[i] = midPre[ncyc + i]; } }
    } else if (mid && phases == 2 && mid[10] == 0 && mid[11] == 1) {
        return;                                   // the warp-per-read prescan already filled mid for this read
    } else if (prescan_qscore) {
        int bestqscore = 0, maxHits = 0, minHitsToScore = 1, cycle = 0; bool early = false;
        const int ncyc = 2 * X->nblocks;
        for (int i = 0; i < ncyc; i++) { precounts[i] = n; prescores[i] = maxQuickScore; }
        havePre = true;
        for (int chrom = minChrom; chrom <= maxChrom && !early; chrom = ((chrom & c->highMask) + c->cpb)) {
            for (int pmi = 0; pmi < 2 && !early; pmi++, cycle++) {
                int nh = get_hits(c, pmi == 0 ? keysP : keysM, pmi == 0 ? offsetsP : offsetsM, pmi == 0 ? keyScoresP : keyScoresM, n, chrom);
                if (nh < minHitsToScore) { prescores[cycle] = -9999; precounts[cycle] = 0; }
                else {
                    if (nh < n) nh = shrink_hits(c, n);
                    int ts, th;
                    const long long tp0 = clock64();
                    find_max_qscore2(c, nh, chrom, minHitsToScore, 1, bestqscore >= maxQuickScore && pretend, &ts, &th);
                    c->tPrescan += clock64() - tp0;
                    prescores[cycle] = ts; precounts[cycle] = th;
                    bestqscore = imax(ts, bestqscore); maxHits = imax(maxHits, th);
                    if (bestqscore >= maxQuickScore && pretend) { minHitsToScore = imax(minHitsToScore, maxHits); early = true; }
                }
            }
        }
        bestScores[1] = imax(bestScores[1], maxHits); bestScores[3] = imax(bestScores[3], bestqscore);
        bool dead = false;
        if (bestScores[1] < 1) dead = true;
        else if ((float)bestScores[3] < __fmul_rn((float)maxQuickScore, MIN_QSCORE_MULT2)) dead = true;
        if (dead) { H->status = c->status; if (mid) mid[9] = 1; return; }
        if (bestScores[3] >= maxQuickScore && pretend) {
            hitsCutoff = approx_hits_cutoff(c, n, bestScores[1], 1, 1);
            qscoreCutoff = imax(qscoreCutoff, (int)(bestScores[3] * DYNAMIC_QSCORE_THRESH_PERFECT));
        } else {
            hitsCutoff = approx_hits_cutoff(c, n, bestScores[1], 1, 0);
            qscoreCutoff = imax(qscoreCutoff, (int)(bestScores[3] * PRESCAN_QSCORE_THRESH));
        }
    }
    if ((phases & 2) && mid) {
        mid[2] = c->status; mid[4] = bestScores[1]; mid[5] = bestScores[3]; mid[6] = hitsCutoff; mid[7] = qscoreCutoff; mid[8] = havePre ? 1 : 0; mid[9] = 0;
        if (havePre) { const int ncyc = 2 * X->nblocks; for (int i = 0; i < ncyc; i++) { midPre[i] = precounts[i]; midPre[ncyc + i] = prescores[i]; } }
    }
    if (!(phases & 4)) return;
    int maxScore = 70 + (len - 1) * 100;                        // msa.maxQuality(baseScores)
    bool fullyDefined = true;
    for (int i = 0; i < len; i++) { maxScore += baseScoresP[i]; fullyDefined = fullyDefined && base_defined(basesP[i]); }
    H->max_score = maxScore; H->max_quick_score = maxQuickScore;
    int cycle = 0; bool done = false;
    for (int chrom = minChrom; chrom <= maxChrom && !done; chrom = ((chrom & c->highMask) + c->cpb)) {
        for (int strand = 0; strand < 2 && !done; strand++) {
            if (!havePre || precounts[cycle] >= hitsCutoff || prescores[cycle] >= qscoreCutoff) {
                const int nh = get_hits(c, strand == 0 ? keysP : keysM, strand == 0 ? offsetsP : offsetsM, strand == 0 ? keyScoresP : keyScoresM, n, chrom);
                const long long tw0 = clock64();
                if (nh >= 1)
                    slow_walk3(c, strand == 0 ? basesP : basesM, strand == 0 ? baseScoresP : baseScoresM, len, n, chrom, strand,
                               obeyLimits, R, bestScores, allBasesCovered, maxScore, fullyDefined, quitAfterTwoPerfects);
                c->tWalk += clock64() - tw0;
            }
            cycle++;
            if (quitAfterTwoPerfects && bestScores[5] >= 2) done = true;
        }
    }
    for (int i = 0; i < 6; i++) H->best_scores[i] = bestScores[i];
    H->status = c->status;
}

struct SearchParams {
    SearchIndex X;
    const int8_t* bases; const int8_t* baseScores; const long long* read_off; long long nreads;
    const int* nkeys; const int* offsets; const int* keyScores; int maxKeys; int quitAfterTwoPerfects;
    bbm_search_head* heads; bbm_site* sites; int maxSites;
    char* pool; unsigned int* counter;
    int phases; int* mid; int midStride;      // split launches: one phase of BBIndex.find per kernel, state carried in mid[read][midStride]
    unsigned long long* prof;     // optional: 5 cycle counters {total, filter, prescan, walk (incl. extend), extend}
};

constexpr int SEARCH_THREADS = 64;
constexpr int SEARCH_FAST_KEYS = 32;           // batches with at most this many keys per read keep the walk arrays in shared memory
// per-thread block in the global pool: locArray[608] int, basesM[608], baseScoresM[608], then (pool variant only) the walk arrays for 96 keys
constexpr size_t SEARCH_POOL_FIXED = (size_t)SEARCH_MAX_READ * 4 + 2 * SEARCH_MAX_READ + (size_t)RING_CAP * 4 + RING_CAP;
constexpr size_t SEARCH_POOL_BYTES = SEARCH_POOL_FIXED + (size_t)MAXK * (3 * 4 + 2 * 2);

template <bool SHARED>
__global__ void __launch_bounds__(SEARCH_THREADS, SHARED ? 7 : 8) search_kernel(SearchParams P) {
    __shared__ int sVal[SHARED ? SEARCH_FAST_KEYS * SEARCH_THREADS : 1], sRow[SHARED ? SEARCH_FAST_KEYS * SEARCH_THREADS : 1],
                   sStop[SHARED ? SEARCH_FAST_KEYS * SEARCH_THREADS : 1];
    __shared__ short sOf[SHARED ? SEARCH_FAST_KEYS * SEARCH_THREADS : 1], sKs[SHARED ? SEARCH_FAST_KEYS * SEARCH_THREADS : 1];
    const long long slot = (long long)blockIdx.x * SEARCH_THREADS + threadIdx.x;
    char* mine = P.pool + slot * SEARCH_POOL_BYTES;
    ctx_t ctx; ctx_t* c = &ctx;
    c->locArray = (int*)mine;
    int8_t* basesM = (int8_t*)(mine + (size_t)SEARCH_MAX_READ * 4);
    int8_t* baseScoresM = basesM + SEARCH_MAX_READ;
    c->ringV = (int*)(baseScoresM + SEARCH_MAX_READ); c->ringC = (signed char*)(c->ringV + RING_CAP);
    if (SHARED) {
        c->S = SEARCH_THREADS;
        c->values = sVal + threadIdx.x; c->rows = sRow + threadIdx.x; c->stops = sStop + threadIdx.x; c->of = sOf + threadIdx.x; c->ks = sKs + threadIdx.x;
    } else {
        c->S = 1;
        int* w = (int*)(mine + SEARCH_POOL_FIXED);
        c->values = w; c->rows = w + MAXK; c->stops = w + 2 * MAXK; c->of = (short*)(w + 3 * MAXK); c->ks = c->of + MAXK;
    }
    c->tExtend = 0; c->tPrescan = 0; c->tWalk = 0; c->tFilter = 0;
    const long long tk0 = clock64();
    for (;;) {
        const unsigned r = atomicAdd(P.counter, 1u);
        if ((long long)r >= P.nreads) break;
        bbm_search_head* H = P.heads + r;
        if (P.phases & 1) { bbm_search_head z = {}; *H = z; }
        if (P.phases == 4 && P.mid) { const int* m = P.mid + (long long)r * P.midStride; if (m[0] >= 1 && m[0] <= 32 && m[12] == 1) continue; }   // walked by walk_warp_kernel
        ReadState R; R.sites = P.sites + (long long)r * P.maxSites; R.maxSites = P.maxSites; R.nsites = 0;
        const long long o = P.read_off[r]; const int len = (int)(P.read_off[r + 1] - o);
        const int nk = P.nkeys[r];
        if (SHARED && nk > SEARCH_FAST_KEYS) H->status = BBM_ST_BADARG;       // cannot happen: the host picks this kernel from maxKeys
        else if (nk > 0)
            search_read(&P.X, P.bases + o, len, P.baseScores + o, P.offsets + (long long)r * P.maxKeys, P.keyScores + (long long)r * P.maxKeys,
                        nk, P.quitAfterTwoPerfects, H, &R, c, basesM, baseScoresM, P.phases, P.mid ? P.mid + (long long)r * P.midStride : nullptr, P.maxKeys);
        else if (P.mid && (P.phases & 1)) P.mid[(long long)r * P.midStride] = 0;
        if (P.phases & 4) H->nsites = R.nsites;
    }
    if (P.prof) {
        atomicAdd(P.prof + 0, (unsigned long long)(clock64() - tk0)); atomicAdd(P.prof + 1, (unsigned long long)c->tFilter);
        atomicAdd(P.prof + 2, (unsigned long long)c->tPrescan); atomicAdd(P.prof + 3, (unsigned long long)c->tWalk); atomicAdd(P.prof + 4, (unsigned long long)c->tExtend);
    }
}


// =====================  warp-per-read prescan (phase 2 of the split launches)  =====================
// prescanAllBlocks / findMaxQscore2 (BBIndex.java:642-741, 2294-2450) with one WARP per read: lane c owns hit list c of the current
// (block, strand) — its cursor, head value and prefetched next value live in registers, the index gathers of a read are issued by all
// lanes at once.  The reference pops the merged lists one site at a time; here every iteration first retires, in one go, the longest
// prefix of the merged order that consists of *isolated* sites: sites whose window [site-MAX_INDEL, site+MAX_INDEL2] contains no other
// list head at the time they are popped (decided from the heads and next values of all lists, ranked through shared memory) and no
// last value of an exhausted list.  Such a site has exactly one hit: it is scored keyScore + scoreZ1Key if the cutoff is 1 and ignored
// otherwise (:2352-2376, 2490-2492, 2883), so a whole run of them only raises topQscore to the largest of their scores.  Any other site
// goes through the reference's own step (count, quickScore, scoreZ2, pops), evaluated redundantly by all lanes from shuffled values.
// Reads with more than 32 keys are left to the thread-per-read kernel (mid[10] = 1).
constexpr int PW_WARPS = 4;
constexpr unsigned FULL = 0xffffffffu;
struct PwSlot { int val, nx, flags; };     // flags: bit0 hasNx, bit1 retire

__device__ __forceinline__ int jadd(int a, int b) { return (int)((unsigned)a + (unsigned)b); }   // Java int arithmetic (wraps)

__global__ void __launch_bounds__(PW_WARPS * 32) prescan_warp_kernel(SearchParams P) {
    __shared__ PwSlot slots[PW_WARPS][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    const SearchIndex* X = &P.X;
    const bbm_index_cfg* g = X->cfg;
    ctx_t cc; ctx_t* c = &cc;
    c->X = X; c->K = g->keylen; c->baseKeyHitScore = BASE_HIT_SCORE * c->K;
    c->indelPenalty = (c->baseKeyHitScore / 2) - 1;
    c->maxPenaltyMisaligned = c->baseKeyHitScore - (1 + c->baseKeyHitScore / 8);
    c->scoreZ1Key = Z_SCORE_MULT * c->K;
    c->shift = g->shift_length; c->cpb = g->chroms_per_block; c->lowMask = c->cpb - 1; c->highMask = ~c->lowMask;
    c->siteMask = (int)(0xFFFFFFFFu >> (g->chrombits + 1));
    const int K = c->K, MK = P.maxKeys;
    for (;;) {
        unsigned r = 0;
        if (lane == 0) r = atomicAdd(P.counter, 1u);
        r = __shfl_sync(FULL, r, 0);
        if ((long long)r >= P.nreads) break;
        int* mid = P.mid + (long long)r * P.midStride;
        const int n = mid[0];
        if (n < 1) continue;
        if (n > 32) { if (lane == 0) { mid[10] = 1; mid[11] = 1; } continue; }
        const int numHitsRead = mid[1], status0 = mid[2];
        const int len = (int)(P.read_off[r + 1] - P.read_off[r]), numKeysOriginal = P.nkeys[r];
        const int* a = mid + MID_HDR;
        int keyP = -1, keyM = -1, ofsP = 0, ofsM = 0, kscP = 0, kscM = 0;
        if (lane < n) { keyP = a[lane]; keyM = a[MK + lane]; ofsP = a[2 * MK + lane]; ofsM = a[3 * MK + lane]; kscP = a[4 * MK + lane]; kscM = a[5 * MK + lane]; }
        // maxQuickScore of the read (plus-strand arrays), allBasesCovered, pretend (BBIndex.java:470-486)
        int maxQuickScore;
        {
            int x = lane < n ? kscP : 0;
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
            int score = 0, a0 = -1, b0 = -1;
            for (int i = 0; i < n; i++) { const int av = __shfl_sync(FULL, ofsP, i); if (b0 < av) { score += b0 - a0; a0 = av; } b0 = av + K; }
            score += b0 - a0;
            maxQuickScore = x + score * Z_SCORE_MULT + Y_SCORE_MULT * (__shfl_sync(FULL, ofsP, n - 1) - __shfl_sync(FULL, ofsP, 0));
        }
        const int ofsFirst = __shfl_sync(FULL, ofsP, 0), ofsLast = __shfl_sync(FULL, ofsP, n - 1);
        const int ofsPrev = __shfl_up_sync(FULL, ofsP, 1);
        const bool gapBad = __any_sync(FULL, lane >= 1 && lane < n && ofsP > ofsPrev + K);
        const bool allBasesCovered = (ofsFirst == 0) && (ofsLast == (len - K)) && !gapBad;
        const bool pretend = (allBasesCovered || n >= numKeysOriginal - 4 || (n >= 9 && (ofsLast - ofsFirst + K) > imax(40, (int)(len * .75f))));
        int hitsCutoff = 0, qscoreCutoff = (int)(MIN_QSCORE_MULT * maxQuickScore);
        int best1 = 0, best3 = 0; bool havePre = false, dead = false;
        int* midPre = mid + MID_HDR + 6 * MK;
        if (numHitsRead >= 5) {
            int bestqscore = 0, maxHitsAll = 0, minHitsToScore = 1, cycle = 0; bool early = false;
            const int ncyc = 2 * X->nblocks;
            if (lane < ncyc) { midPre[lane] = n; midPre[ncyc + lane] = maxQuickScore; }
            for (int i = 32 + lane; i < ncyc; i += 32) { midPre[i] = n; midPre[ncyc + i] = maxQuickScore; }
            havePre = true;
            for (int chrom = 1; chrom <= X->nchroms && !early; chrom = ((chrom & c->highMask) + c->cpb)) {
                const int baseChrom = base_chrom(c, chrom);
                const SearchBlock* b = block_of(c, chrom);
                for (int pmi = 0; pmi < 2 && !early; pmi++, cycle++) {
                    // ---- getHits (:353-373): all lists of the read at once ----
                    const int key = pmi == 0 ? keyP : keyM;
                    int st = -1, sp = -1;
                    if (lane < n && key >= 0 && X->counts[key] > 0) {
                        const int s0 = b->starts[key], x = b->starts[key + 1] - s0;
                        if (x > 0 && b->sites[s0] != -1) { st = s0; sp = s0 + x; }
                    }
                    const unsigned vmask = __ballot_sync(FULL, st >= 0);
                    const int nh = __popc(vmask);
                    int ts = -9999, th = 0;
                    if (nh >= minHitsToScore) {
                        // ---- shrink (:783-813): lane col takes the col-th list that has hits ----
                        const int src = lane < nh ? __fns(vmask, 0, lane + 1) : 0;
                        int row = __shfl_sync(FULL, st, src), stop = __shfl_sync(FULL, sp, src);
                        const int ofs = __shfl_sync(FULL, pmi == 0 ? ofsP : ofsM, src), ksc = __shfl_sync(FULL, pmi == 0 ? kscP : kscM, src);
                        const bool isCol = lane < nh;
                        bool live = isCol;
                        int val = 0, nx = 0; bool hasNx = false;
                        if (isCol) {
                            val = site_minus_offset(c, b->sites[row], ofs, baseChrom);
                            hasNx = row + 1 < stop;
                            if (hasNx) nx = site_minus_offset(c, b->sites[row + 1], ofs, baseChrom);
                        }
                        // maxQuickScore of these columns (:2316)
                        int mqs;
                        {
                            int x = isCol ? ksc : 0;
#pragma unroll
                            for (int o = 16; o >= 1; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
                            int score = 0, a0 = -1, b0 = -1;
                            for (int i = 0; i < nh; i++) { const int av = __shfl_sync(FULL, ofs, i); if (b0 < av) { score += b0 - a0; a0 = av; } b0 = av + K; }
                            score += b0 - a0;
                            mqs = x + score * Z_SCORE_MULT + Y_SCORE_MULT * (__shfl_sync(FULL, ofs, nh - 1) - __shfl_sync(FULL, ofs, 0));
                        }
                        int topQscore = -999999999, maxHits = 0;
                        int cutoff = imax(minHitsToScore, imin(1, nh - 1));
                        int nActive = nh, staleMax = -0x7fffffff - 1;
                        bool done = false;
                        while (!done) {
                            const unsigned liveMask = __ballot_sync(FULL, live);
                            if (!liveMask) break;
                            const int nLive = __popc(liveMask);
                            // ---- rank the live heads by (value, column), move them to rank order through shared memory ----
                            int rank = 0;
                            for (int j = 0; j < nh; j++) {
                                const int vj = __shfl_sync(FULL, val, j);
                                if (((liveMask >> j) & 1u) && (vj < val || (vj == val && j < lane))) rank++;
                            }
                            if (live) { PwSlot q; q.val = val; q.nx = nx; q.flags = hasNx ? 1 : 0; slots[wib][rank] = q; }
                            __syncwarp();
                            bool iso = false;
                            {
                                const int big = 0x7fffffff, small = -0x7fffffff - 1;
                                const bool pos = lane < nLive;                                     // this lane now stands for rank `lane`
                                PwSlot me; me.val = 0; me.nx = 0; me.flags = 0;
                                if (pos) me = slots[wib][lane];
                                const int pv = me.val;
                                int pmin = (pos && (me.flags & 1)) ? me.nx : big;                  // next values of the lists ranked at or below me
                                int pmax = (pos && !(me.flags & 1)) ? me.val : small;              // lists without one leave their value behind when popped
#pragma unroll
                                for (int o = 1; o < 32; o <<= 1) {                                 // inclusive scans in rank order
                                    const int tmin = __shfl_up_sync(FULL, pmin, o), tmax = __shfl_up_sync(FULL, pmax, o);
                                    if (lane >= o) { pmin = imin(pmin, tmin); pmax = imax(pmax, tmax); }
                                }
                                int exMin = __shfl_up_sync(FULL, pmin, 1), exMax = __shfl_up_sync(FULL, pmax, 1);
                                if (lane == 0) { exMin = big; exMax = small; }
                                const int nextVal = (pos && lane + 1 < nLive) ? slots[wib][lane + 1].val : big;
                                const bool room = pv <= 0x7fffffff - MAX_INDEL2;
                                const long long hi = (long long)pv + MAX_INDEL2, lo = (long long)pv - MAX_INDEL;
                                iso = pos && room && (lane + 1 >= nLive || (long long)nextVal > hi) && (exMin == big || (long long)exMin > hi) &&
                                      ((long long)exMax < lo) && ((long long)staleMax < lo);
                            }
                            const unsigned nonIso = __ballot_sync(FULL, lane < nLive && !iso);
                            const int firstNonIso = nonIso ? __ffs(nonIso) - 1 : nLive;
                            if (lane < nLive) slots[wib][lane].flags |= (lane < firstNonIso) ? 2 : 0;
                            __syncwarp();
                            if (firstNonIso > 0) {
                                // ---- retire the run of isolated sites ----
                                const bool retire = live && (slots[wib][rank].flags & 2);
                                int q = retire ? ksc + c->scoreZ1Key : -0x7fffffff - 1;
#pragma unroll
                                for (int o = 16; o >= 1; o >>= 1) q = imax(q, __shfl_xor_sync(FULL, q, o));
                                if (cutoff <= 1 && q > topQscore) { maxHits = imax(1, maxHits); topQscore = q; }
                                bool exhausted = false;
                                if (retire) {
                                    row++;
                                    if (hasNx) { val = nx; hasNx = row + 1 < stop; if (hasNx) nx = site_minus_offset(c, b->sites[row + 1], ofs, baseChrom); }
                                    else { live = false; exhausted = true; }
                                }
                                const unsigned exMask = __ballot_sync(FULL, exhausted);
                                if (exMask) {
                                    int sv = exhausted ? val : -0x7fffffff - 1;
#pragma unroll
                                    for (int o = 16; o >= 1; o >>= 1) sv = imax(sv, __shfl_xor_sync(FULL, sv, o));
                                    staleMax = imax(staleMax, sv);
                                    nActive -= __popc(exMask);
                                    if (nActive < cutoff) done = true;
                                }
                                __syncwarp();
                                continue;
                            }
                            __syncwarp();
                            // ---- the reference's own step for the smallest head (:2340-2440) ----
                            const int centerIndex = __ffs(__ballot_sync(FULL, live && rank == 0)) - 1;
                            const int site = __shfl_sync(FULL, val, centerIndex);
                            const int minsite = jadd(site, -MAX_INDEL), maxsite = jadd(site, MAX_INDEL2);
                            const int approxHits = __popc(__ballot_sync(FULL, isCol && val >= minsite && val <= maxsite));
                            if (approxHits >= cutoff) {
                                int qscore;
                                if (approxHits == 1) qscore = __shfl_sync(FULL, ksc, centerIndex) + c->scoreZ1Key;
                                else {
                                    // quickScore (:2490-2511) = key score + scoreLeft + scoreRight - centerIndex + Y * scoreY; scoreZ2 (:2882-2914)
                                    int sc = __shfl_sync(FULL, ksc, centerIndex);
                                    int loc = site, prev;
                                    for (int i = centerIndex - 1; i >= 0; i--) {
                                        const int v = __shfl_sync(FULL, val, i), ks = __shfl_sync(FULL, ksc, i);
                                        if (v >= 0) {
                                            prev = loc; loc = v;
                                            const int offset = absdif(loc, prev);
                                            if (offset <= MAX_INDEL) { sc += ks; if (offset != 0) sc -= imin(c->indelPenalty + INDEL_PENALTY_MULT * offset, c->maxPenaltyMisaligned); }
                                            else loc = prev;
                                        }
                                    }
                                    loc = site;
                                    for (int i = centerIndex + 1; i < nh; i++) {
                                        const int v = __shfl_sync(FULL, val, i), ks = __shfl_sync(FULL, ksc, i);
                                        if (v >= 0) {
                                            prev = loc; loc = v;
                                            const int offset = absdif(loc, prev);
                                            if (offset <= MAX_INDEL) { sc += ks; if (offset != 0) sc -= imin(c->indelPenalty + INDEL_PENALTY_MULT * offset, c->maxPenaltyMisaligned); }
                                            else loc = prev;
                                        }
                                    }
                                    sc -= centerIndex;
                                    const unsigned eq = __ballot_sync(FULL, isCol && val == site);
                                    const int rightIndex = 31 - __clz(eq);
                                    sc += Y_SCORE_MULT * (__shfl_sync(FULL, ofs, rightIndex) - __shfl_sync(FULL, ofs, centerIndex));
                                    const int maxLoc = jadd(site, MAX_INDEL2), minLoc = imax(0, jadd(site, -MAX_INDEL));
                                    int z = 0, a0 = -1, b0 = -1;
                                    for (int i = 0; i < nh; i++) {
                                        const int v = __shfl_sync(FULL, val, i), av = __shfl_sync(FULL, ofs, i);
                                        if (v >= minLoc && v <= maxLoc) { if (b0 < av) { z += b0 - a0; a0 = av; } b0 = av + K; }
                                    }
                                    z += b0 - a0;
                                    qscore = sc + z * Z_SCORE_MULT;
                                }
                                if (qscore > topQscore) {
                                    maxHits = imax(approxHits, maxHits);
                                    cutoff = imax(cutoff, approxHits - 1);
                                    topQscore = qscore;
                                    if (qscore >= mqs) { done = true; continue; }
                                }
                            }
                            // pops: every live list sitting on `site` advances (again if its next site is `site` too)
                            for (;;) {
                                const bool hit = live && val == site;
                                if (!__any_sync(FULL, hit)) break;
                                bool exhausted = false;
                                if (hit) {
                                    row++;
                                    if (hasNx) { val = nx; hasNx = row + 1 < stop; if (hasNx) nx = site_minus_offset(c, b->sites[row + 1], ofs, baseChrom); }
                                    else { live = false; exhausted = true; }
                                }
                                const unsigned exMask = __ballot_sync(FULL, exhausted);
                                if (exMask) {
                                    int sv = exhausted ? val : -0x7fffffff - 1;
#pragma unroll
                                    for (int o = 16; o >= 1; o >>= 1) sv = imax(sv, __shfl_xor_sync(FULL, sv, o));
                                    staleMax = imax(staleMax, sv);
                                    nActive -= __popc(exMask);
                                    if (nActive < cutoff) { done = true; break; }
                                }
                            }
                        }
                        ts = topQscore; th = maxHits;
                        bestqscore = imax(ts, bestqscore); maxHitsAll = imax(maxHitsAll, th);
                        if (bestqscore >= maxQuickScore && pretend) { minHitsToScore = imax(minHitsToScore, maxHitsAll); early = true; }
                    } else { ts = -9999; th = 0; }
                    if (lane == 0) { midPre[cycle] = th; midPre[ncyc + cycle] = ts; }
                }
            }
            best1 = maxHitsAll; best3 = bestqscore;
            if (best1 < 1) dead = true;
            else if ((float)best3 < __fmul_rn((float)maxQuickScore, MIN_QSCORE_MULT2)) dead = true;
            if (!dead) {
                if (best3 >= maxQuickScore && pretend) {
                    hitsCutoff = approx_hits_cutoff(c, n, best1, 1, 1);
                    qscoreCutoff = imax(qscoreCutoff, (int)(best3 * DYNAMIC_QSCORE_THRESH_PERFECT));
                } else {
                    hitsCutoff = approx_hits_cutoff(c, n, best1, 1, 0);
                    qscoreCutoff = imax(qscoreCutoff, (int)(best3 * PRESCAN_QSCORE_THRESH));
                }
            }
        }
        if (lane == 0) {
            mid[2] = status0; mid[4] = best1; mid[5] = best3; mid[6] = hitsCutoff; mid[7] = qscoreCutoff; mid[8] = havePre ? 1 : 0; mid[9] = dead ? 1 : 0; mid[10] = 0; mid[11] = 1;
            if (dead) P.heads[r].status = status0;
        }
        __syncwarp();
    }
}

#include "search_walk_warp.cuh"

}  // namespace bbm

using namespace bbm;

extern "C" size_t bbm_search_pool_bytes() { return SEARCH_POOL_BYTES; }
extern "C" int bbm_search_threads() { return SEARCH_THREADS; }
extern "C" int bbm_launch_search_prescan_warp(const bbm_index_cfg* d_cfg, const void* d_blocks, int nblocks, int nchroms, const int* d_counts,
                                              const long long* read_off, long long nreads, const int* nkeys, int maxKeys, bbm_search_head* heads,
                                              unsigned int* counter, int blocks, int* mid, int midStride, cudaStream_t st) {
    SearchParams P; memset(&P, 0, sizeof(P));
    P.X.cfg = d_cfg; P.X.blocks = (const SearchBlock*)d_blocks; P.X.nblocks = nblocks; P.X.nchroms = nchroms; P.X.counts = d_counts;
    P.read_off = read_off; P.nreads = nreads; P.nkeys = nkeys; P.maxKeys = maxKeys; P.heads = heads; P.counter = counter; P.mid = mid; P.midStride = midStride;
    prescan_warp_kernel<<<blocks, PW_WARPS * 32, 0, st>>>(P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_search_walk_warp(const bbm_index_cfg* d_cfg, const void* d_blocks, int nblocks, int nchroms, const int* d_counts, const int8_t* d_chroms,
                                           const long long* d_chrom_off, const int8_t* bases, const int8_t* baseScores, const long long* read_off, long long nreads,
                                           const int* nkeys, int maxKeys, int quit2, bbm_search_head* heads, bbm_site* sites, int maxSites, unsigned int* counter, int blocks,
                                           int* mid, int midStride, cudaStream_t st) {
    SearchParams P;
    memset(&P, 0, sizeof P);
    P.X.cfg = d_cfg; P.X.blocks = (const SearchBlock*)d_blocks; P.X.nblocks = nblocks; P.X.nchroms = nchroms; P.X.counts = d_counts;
    P.X.chroms = d_chroms; P.X.chrom_off = d_chrom_off;
    P.bases = bases; P.baseScores = baseScores; P.read_off = read_off; P.nreads = nreads; P.nkeys = nkeys; P.maxKeys = maxKeys; P.quitAfterTwoPerfects = quit2;
    P.heads = heads; P.sites = sites; P.maxSites = maxSites; P.counter = counter; P.phases = 4; P.mid = mid; P.midStride = midStride;
    walk_warp_kernel<<<blocks, WW_WARPS * 32, 0, st>>>(P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_search_mid_stride(int maxKeys, int nblocks) { return MID_HDR + 6 * maxKeys + 4 * nblocks; }
extern "C" int bbm_launch_search(const bbm_index_cfg* d_cfg, const void* d_blocks, int nblocks, int nchroms, const int* d_counts, const int* d_hist,
                                 const int8_t* d_chroms, const long long* d_chrom_off, const int8_t* bases, const int8_t* baseScores,
                                 const long long* read_off, long long nreads, const int* nkeys, const int* offsets, const int* keyScores, int maxKeys,
                                 int quitAfterTwoPerfects, bbm_search_head* heads, bbm_site* sites, int maxSites, void* pool,
                                 unsigned int* counter, unsigned long long* prof, int blocks, int forcePool, int phases, int* mid, int midStride,
                                 cudaStream_t st) {
    SearchParams P; P.prof = prof; P.phases = phases; P.mid = mid; P.midStride = midStride;
    P.X.cfg = d_cfg; P.X.blocks = (const SearchBlock*)d_blocks; P.X.nblocks = nblocks; P.X.nchroms = nchroms; P.X.counts = d_counts; P.X.hist = d_hist;
    P.X.chroms = d_chroms; P.X.chrom_off = d_chrom_off;
    P.bases = bases; P.baseScores = baseScores; P.read_off = read_off; P.nreads = nreads; P.nkeys = nkeys; P.offsets = offsets; P.keyScores = keyScores;
    P.maxKeys = maxKeys; P.quitAfterTwoPerfects = quitAfterTwoPerfects; P.heads = heads; P.sites = sites; P.maxSites = maxSites;
    P.pool = (char*)pool; P.counter = counter;
    if (maxKeys <= SEARCH_FAST_KEYS && !forcePool) search_kernel<true><<<blocks, SEARCH_THREADS, 0, st>>>(P);
    else search_kernel<false><<<blocks, SEARCH_THREADS, 0, st>>>(P);
    return (int)cudaGetLastError();
}
