// sitelist_dev.cuh — device helpers shared by the list-policy kernels (sitelist.cu) and the mapper kernels (genmatch.cu, pairing.cu):
// SiteScore comparators and setters, stable sorts, Tools.trimSiteList / mergeDuplicateSites / countTopScores, MSA.scoreNoIndels,
// SiteScore.setPerfect, GapTools.fixGaps and the clearzone-3 fraction.  Reference lines are cited at each function.
#pragma once
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

constexpr int SL_MAX_CAP = 64;

__device__ __forceinline__ int ss_compare(const bbm_ss& a, const bbm_ss& o) {      // SiteScore.compareTo
    int x = o.score - a.score; if (x) return x;
    x = o.slow_score - a.slow_score; if (x) return x;
    x = o.paired_score - a.paired_score; if (x) return x;
    x = o.quick_score - a.quick_score; if (x) return x;
    x = a.chrom - o.chrom; if (x) return x;
    return a.start - o.start;
}
__device__ __forceinline__ int ss_pcomp(const bbm_ss& a, const bbm_ss& b) {        // SiteScore.PCOMP
    if (a.chrom != b.chrom) return a.chrom - b.chrom;
    if (a.start != b.start) return a.start - b.start;
    if (a.stop != b.stop) return a.stop - b.stop;
    if (a.strand != b.strand) return a.strand - b.strand;
    if (a.score != b.score) return b.score - a.score;
    if (a.slow_score != b.slow_score) return b.slow_score - a.slow_score;
    if (a.quick_score != b.quick_score) return b.quick_score - a.quick_score;
    if (a.perfect != b.perfect) return a.perfect ? -1 : 1;
    if (a.rescued != b.rescued) return a.rescued ? 1 : -1;
    return 0;
}
template <bool POSITIONAL>
static __device__ void stable_sort(bbm_ss* v, int n) {        // Collections.sort is stable; so is insertion sort
    for (int i = 1; i < n; i++) {
        const bbm_ss x = v[i]; int j = i - 1;
        while (j >= 0 && (POSITIONAL ? ss_pcomp(v[j], x) : ss_compare(v[j], x)) > 0) { v[j + 1] = v[j]; j--; }
        if (j + 1 != i) v[j + 1] = x;
    }
}
static __device__ int compact(bbm_ss* v, int n, unsigned long long dead) {
    if (!dead) return n;
    int k = 0;
    for (int i = 0; i < n; i++) if (!((dead >> i) & 1ull)) { if (k != i) v[k] = v[i]; k++; }
    return k;
}

// Tools.trimSiteList + trimSitesBelowCutoff (retainSemiperfect = true)
static __device__ int trim_site_list(bbm_ss* v, int& n, float frac, bool retainPaired, int minS, int maxS) {
    if (n == 0) return -999999;
    if (n == 1) return v[0].score;
    int maxScore = -999999;
    if (minS > 1 && minS < n) maxScore = v[0].score;
    else for (int i = 0; i < n; i++) maxScore = imax(maxScore, v[i].score);
    const int cutoff = (int)__fmul_rn((float)maxScore, frac);
    if (n <= minS) return maxScore;
    while (n > maxS) n--;
    int removed = 0; const int maxToRemove = n - minS;
    unsigned long long dead = 0;
    for (int i = n - 1; i >= 0; i--) {
        if (!v[i].semiperfect && v[i].score < cutoff && (!retainPaired || v[i].paired_score <= 0)) {
            dead |= 1ull << i; removed++;
            if (removed >= maxToRemove) break;
        }
    }
    n = compact(v, n, dead);
    return maxScore;
}
static __device__ int trim_list(bbm_ss* v, int& n, bool retainPaired, int maxScore, bool specialCasePerfect, int minS, int maxS) {
    if (n == 0) return -99999;
    if (n == 1) return v[0].score;
    const int highest = trim_site_list(v, n, .6f, retainPaired, minS, maxS);
    if (highest == maxScore && specialCasePerfect) {
        trim_site_list(v, n, .94f, retainPaired, minS, maxS);
        if (n > 8) trim_site_list(v, n, .99f, retainPaired, minS, maxS);
        return highest;
    }
    const int mstr2 = (minS <= 1 ? 1 : minS + 1);
    if (n > 4) trim_site_list(v, n, .65f, retainPaired, minS, maxS);
    if (n > 8) trim_site_list(v, n, .7f, retainPaired, minS, maxS);
    if (n > 12) trim_site_list(v, n, .75f, retainPaired, minS, maxS);
    if (n > 16) trim_site_list(v, n, .8f, retainPaired, minS, maxS);
    if (n > 20) trim_site_list(v, n, .85f, retainPaired, minS, maxS);
    if (n > 24) trim_site_list(v, n, .9f, retainPaired, minS, maxS);
    if (n > 32) trim_site_list(v, n, .95f, retainPaired, minS, maxS);
    if (n > 40) trim_site_list(v, n, .97f, retainPaired, mstr2, maxS);
    if (n > 48) trim_site_list(v, n, .99f, retainPaired, mstr2, maxS);
    return highest;
}

// SiteScore.setSlowScore (stream/SiteScore.java:962-983): also moves pairedScore
__device__ __forceinline__ void set_slow_score(bbm_ss& s, int x) {
    if (x <= 0) { s.paired_score = x; }
    else if (s.paired_score > 0) s.paired_score = (s.slow_score > 0) ? x + (s.paired_score - s.slow_score) : x + 1;
    s.slow_score = x;
}

__device__ __forceinline__ int max_quality(int len) { return 70 + (len - 1) * 100; }                    // …JNI.java:1321-1323
__device__ __forceinline__ int max_imperfect(int len) { return max_quality(len) + imin(-472, -395 - 100); }   // :1331-1336

// MSA.scoreNoIndels (…JNI.java:1033-1089), score only (the kernel in noindel.cu also writes match strings)
static __device__ int score_no_indels(const int8_t* __restrict__ read, int len, const int8_t* __restrict__ ref, int refLen, int refStart) {
    int readStart = 0, readStop = len;
    const long long refStop = (long long)refStart + len;
    if (refStart < 0) readStart = -refStart;
    if (refStop > refLen) readStop -= (int)(refStop - refLen);
    int score = 0, mode = -1, timeInMode = 0;
    for (int k = readStart; k < readStop; ++k) {
        const int c = read[k], r = ref[refStart + k];
        if (c == r && c != 'N') { if (mode == 0) { timeInMode++; score += 100; } else { timeInMode = 0; score += 70; } mode = 0; }
        else if (c < 0 || c == 'N') {}
        else if (r < 0 || r == 'N') {}
        else { if (mode == 3) timeInMode++; else timeInMode = 0; score += timeInMode == 0 ? -127 : (timeInMode < 5 ? -51 : -25); mode = 3; }
    }
    return score;
}

// SiteScore.setPerfect(bases) (stream/SiteScore.java:239-291)
static __device__ void ss_set_perfect(bbm_ss& s, const int8_t* __restrict__ bases, int len, const int8_t* __restrict__ ref, int refLen) {
    if (len != s.stop - s.start + 1) { s.perfect = 0; s.semiperfect = 0; return; }
    bool perfect = true, semiperfect = true;
    int refloc = s.start, readloc = 0, N = 0;
    const int mx = imin(s.stop, refLen - 1), nlimit = len / 2;
    if (s.start < 0) { N -= s.start; readloc -= s.start; refloc -= s.start; perfect = false; }
    if (s.stop >= refLen) { N += (s.stop - refLen + 1); perfect = false; }
    if (N > nlimit) { s.perfect = 0; s.semiperfect = 0; return; }
    for (; refloc <= mx; refloc++, readloc++) {
        const int8_t c = bases[readloc], r = ref[refloc];
        if (c != r || c == 'N') {
            perfect = false;
            if (c == 'N') semiperfect = false;
            if (r != 'N' || (N = N + 1) > nlimit) { s.perfect = 0; s.semiperfect = 0; return; }
        }
    }
    semiperfect = (semiperfect && (N <= nlimit));
    perfect = (perfect && semiperfect && (N == 0));
    s.perfect = perfect ? 1 : 0; s.semiperfect = semiperfect ? 1 : 0;
}

static __device__ bool positional_match(const bbm_ss& a, const bbm_ss& b, bool testGaps) {
    if (a.chrom != b.chrom || a.strand != b.strand || a.start != b.start || a.stop != b.stop) return false;
    if (!testGaps || (a.ngaps == 0 && b.ngaps == 0)) return true;
    if (a.ngaps != b.ngaps) return false;            // covers "one is null" as well
    for (int i = 0; i < a.ngaps; i++) if (a.gaps[i] != b.gaps[i]) return false;
    return true;
}

// Tools.mergeDuplicateSites(list, true, true)
static __device__ int merge_duplicate_sites(bbm_ss* v, int n) {
    if (n < 2) return n;
    stable_sort<true>(v, n);
    unsigned long long dead = 0;
    int ai = 0;
    for (int i = 1; i < n; i++) {
        bbm_ss& a = v[ai]; const bbm_ss& b = v[i];
        const bool same = positional_match(a, b, true);
        if (same || positional_match(a, b, false)) {
            if (!same) {    // same outermost boundaries, different gaps: keep the gaps of the better one (decided before the scores are merged)
                bool takeB;
                if (a.score != b.score) takeB = b.score > a.score;
                else if (a.slow_score != b.slow_score) takeB = b.slow_score > a.slow_score;
                else if (a.paired_score != b.paired_score) takeB = b.paired_score > a.paired_score;
                else takeB = false;
                if (takeB) { a.ngaps = b.ngaps; for (int g = 0; g < BBM_MAX_GAPS - 1; g++) a.gaps[g] = b.gaps[g]; }
            }
            set_slow_score(a, imax(a.slow_score, b.slow_score));          // a.setSlowScore(max(...)) moves a positive pairedScore first (Tools.java:733)
            a.paired_score = (a.paired_score <= a.slow_score && b.paired_score <= a.slow_score) ? 0 : imax(0, imax(a.paired_score, b.paired_score));
            a.score = imax(a.score, b.score);
            a.perfect = (a.perfect || b.perfect) ? 1 : 0; a.semiperfect = (a.semiperfect || b.semiperfect) ? 1 : 0;
            dead |= 1ull << i;
        } else ai = i;
    }
    return compact(v, n, dead);
}

static __device__ int count_top_scores(const bbm_ss* v, int n, int thresh) {
    if (n == 0) return 0;
    int count = 1; const int limit = v[0].score - thresh;
    for (int i = 1; i < n; i++) {
        if (v[i].score < limit) break;
        if (v[0].start != v[i].start && v[0].stop != v[i].stop) count++;
    }
    return count;
}

// ---- gap arrays: GapTools.fixGaps and the SiteScore setters that call it ----
constexpr int SL_MINGAP = 256;                 // Shared.MINGAP = GAPBUFFER2 + GAPLEN (align2/Shared.java:20-24)

// GapTools.fixGaps(a, b, gaps, minGap) + fixGaps2, in place; returns the new number of ints (0 = null)
static __device__ int fix_gaps(int a, int b, int* gaps, int n, int minGap) {
    if (n == 0) return 0;
    if (!(gaps[0] <= b && gaps[n - 1] >= a)) return 0;
    int changed = 0;
    if (gaps[0] != a) { gaps[0] = a; changed++; }
    if (gaps[n - 1] != b) { gaps[n - 1] = b; changed++; }
    for (int i = 0; i < n; i++) { if (gaps[i] < a) { gaps[i] = a; changed++; } else if (gaps[i] > b) { gaps[i] = b; changed++; } }
    for (int i = 1; i < n; i++) if (gaps[i - 1] > gaps[i]) { gaps[i] = gaps[i - 1]; changed++; }
    if (changed == 0) return n;
    gaps[0] = a; gaps[n - 1] = b;
    int remove = 0;
    for (int i = 0; i < n; i += 2) {
        gaps[i] = imin(imax(gaps[i], a), b); gaps[i + 1] = imin(imax(gaps[i + 1], a), b);
        if (gaps[i] == gaps[i + 1]) remove++;
    }
    if (remove == 0) return n;
    const int m = n / 2; unsigned dead = 0;
    for (int i = 1; i < m; i++) {
        if (gaps[2 * i] - gaps[2 * i - 1] <= minGap) {
            gaps[2 * i] = imin(gaps[2 * i - 2], gaps[2 * i]); gaps[2 * i + 1] = imax(gaps[2 * i - 1], gaps[2 * i + 1]);
            dead |= 1u << (i - 1);
        }
    }
    int k = 0;
    for (int i = 0; i < m; i++) if (!((dead >> i) & 1u)) { gaps[2 * k] = gaps[2 * i]; gaps[2 * k + 1] = gaps[2 * i + 1]; k++; }
    return k < 2 ? 0 : 2 * k;
}
static __device__ bool check_gaps(const bbm_ss& s) {                                   // SiteScore.CHECKGAPS
    if (s.ngaps == 0) return true;
    if (s.ngaps & 1) return false;
    for (int i = 1; i < s.ngaps; i++) if (s.gaps[i - 1] > s.gaps[i]) return false;
    return s.gaps[0] == s.start && s.gaps[s.ngaps - 1] == s.stop;
}
static __device__ void ss_set_limits(bbm_ss& s, int a, int b) {                        // SiteScore.setLimits
    s.start = a; s.stop = b;
    if (s.ngaps > 0) { s.gaps[0] = a; s.gaps[s.ngaps - 1] = b; if (!check_gaps(s)) s.ngaps = fix_gaps(a, b, s.gaps, s.ngaps, SL_MINGAP); }
}
static __device__ void ss_set_stop(bbm_ss& s, int b) {                                 // SiteScore.setStop
    s.stop = b;
    if (s.ngaps > 0) { s.gaps[s.ngaps - 1] = b; s.ngaps = fix_gaps(s.start, b, s.gaps, s.ngaps, SL_MINGAP); }
}
static __device__ int calc_gref_len(const bbm_ss& s) {                                 // GapTools.calcGrefLen (GAPBUFFER2 = GAPLEN = 128)
    int total = s.stop - s.start + 1;
    for (int i = 2; i < s.ngaps; i += 2) total -= imax(0, (s.gaps[i] - s.gaps[i - 1] - 128) / 128) * 127;
    return total;
}

static __device__ void ss_set_start(bbm_ss& s, int a) {                                // SiteScore.setStart (stream/SiteScore.java:933-942)
    s.start = a;
    if (s.ngaps > 0) { s.gaps[0] = a; if (s.gaps[0] > s.gaps[1]) s.ngaps = fix_gaps(a, s.stop, s.gaps, s.ngaps, SL_MINGAP); }
}

__device__ __forceinline__ float cz3_mult(int i) { return i == 1 ? 1.f : i == 2 ? .75f : i == 3 ? .5f : i == 4 ? .25f : i == 5 ? .125f : .0625f; }   // CZ3_MULTS :2809
__device__ __forceinline__ float calc_cz3_fraction(int score1, int score2, int cz3, float inv) {
    const int dif = score1 - score2;
    if (dif >= cz3) return 0.f;
    const float f = __fmul_rn((float)(cz3 - dif), inv);
    const float a = __fmul_rn(2.f, __fmul_rn(f, f));
    return __fadd_rn(__fadd_rn(f, a), __fmul_rn(a, f));
}

// findTipDeletionsRight / findTipDeletionsLeft, scalar forms (AbstractMapThread.java:2178-2294)
static __device__ int tip_right(const int8_t* __restrict__ bases, int len, const int8_t* __restrict__ ref, int refLen, int minIndex, int originalStop, int searchDist, int tiplen) {
    if (originalStop < minIndex + tiplen - 1 || originalStop >= refLen) return 0;
    const int tipCoord = len - 1;
    int lastMismatch = 0, originalMismatches = 0, contig = 0;
    for (int i = 0; i < tiplen && contig < 5; i++) {
        if (bases[tipCoord - i] != ref[originalStop - i]) { originalMismatches++; lastMismatch = i; contig = 0; } else contig++;
    }
    if (originalMismatches < 3) return 0;
    int minMismatches = originalMismatches, bestStart = originalStop;
    tiplen = lastMismatch + 1;
    if (tiplen < 4) return 0;
    searchDist = imin(searchDist, 30 * originalMismatches);
    const int last = imin(refLen - 1, originalStop + searchDist);
    for (int start = originalStop + 1; start <= last && minMismatches > 0; start++) {
        int m = 0;
        for (int j = 0; j < tiplen && m < minMismatches; j++) m += (bases[tipCoord - j] != ref[start - j]) ? 1 : 0;
        if (m < minMismatches) { bestStart = start; minMismatches = m; }
    }
    if (minMismatches > 2 || originalMismatches - minMismatches < 2) return 0;
    return bestStart - originalStop;
}
static __device__ int tip_left(const int8_t* __restrict__ bases, const int8_t* __restrict__ ref, int refLen, int minIndex, int originalStart, int searchDist, int tiplen) {
    if (originalStart + tiplen >= refLen || minIndex >= originalStart) return 0;
    int lastMismatch = 0, originalMismatches = 0, contig = 0;
    for (int i = 0; i < tiplen && contig < 5; i++) {
        if (bases[i] != ref[originalStart + i]) { originalMismatches++; lastMismatch = i; contig = 0; } else contig++;
    }
    if (originalMismatches < 3) return 0;
    int minMismatches = originalMismatches, bestStart = originalStart;
    tiplen = lastMismatch + 1;
    if (tiplen < 4) return 0;
    searchDist = imin(searchDist, 16 + 16 * originalMismatches + 8 * tiplen);
    const int last = imax(minIndex, originalStart - searchDist);
    for (int start = originalStart - 1; start >= last && minMismatches > 0; start--) {
        int m = 0;
        for (int j = 0; j < tiplen && m < minMismatches; j++) m += (bases[j] != ref[start + j]) ? 1 : 0;
        if (m < minMismatches) { bestStart = start; minMismatches = m; }
    }
    if (minMismatches > 2 || originalMismatches - minMismatches < 2) return 0;
    return originalStart - bestStart;
}

}  // namespace bbm
