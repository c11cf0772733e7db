// rescue.cu — the brute-force ungapped scans around candidate sites.
//   AbstractMapThread.findTipDeletions(SiteScore,...)   current/align2/AbstractMapThread.java:1107-1141
//   AbstractMapThread.findTipDeletionsRight / Left       :2178-2235, :2238-2294
//   AbstractMapThread.quickRescue                        :2303-2405  (+ SiteScore.setPerfect, stream/SiteScore.java:239-291)
//
// One warp per task; lane l evaluates candidate start base±l, so the 32 lanes read 32 consecutive reference bytes per read
// position (one sector) and the read byte is a broadcast.  The reference's loops are sequential with a running best that also
// cuts later candidates short; both scans are restated so that the running state never has to be carried lane to lane:
//   * tip deletions accept only strictly fewer mismatches, so the result is the first start (in scan order) that reaches the
//     minimum full mismatch count — a lexicographic (count, order) arg-min, reduced with shuffles; a candidate the reference
//     cuts short has count >= the running minimum and can never be accepted;
//   * quickRescue's acceptance depends on (mismatches, score, |start-idealStart|) of the running best and shrinks the scan
//     bound after a perfect hit, so each chunk of 32 starts is evaluated in parallel (a lane stops counting once it exceeds
//     the bound that held when the chunk started — bounds only tighten) and the survivors are then resolved in scan order with
//     ballots; every lane keeps the (uniform) running state.
#include <climits>
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

constexpr unsigned RFULL = 0xffffffffu;

// lexicographic min of (count, order) over the warp; returns the pair to every lane
__device__ __forceinline__ void warp_argmin(int& cnt, int& ord) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
        const int c2 = __shfl_xor_sync(RFULL, cnt, d), o2 = __shfl_xor_sync(RFULL, ord, d);
        if (c2 < cnt || (c2 == cnt && o2 < ord)) { cnt = c2; ord = o2; }
    }
}

// findTipDeletionsRight (:2178-2235).  All lanes return the same value.
__device__ int tipdel_right(const int8_t* __restrict__ bases, int len, const int8_t* __restrict__ ref, int refLen, int minIndex,
                            int originalStop, int searchDist, int tiplen, int lane) {
    if (originalStop < minIndex + tiplen - 1) return 0;
    if (originalStop >= refLen) return 0;           // Java would throw; callers pass in-bounds sites
    const int tipCoord = len - 1;
    int lastMismatch = 0, originalMismatches = 0, contig = 0;
    for (int i = 0; i < tiplen && contig < 5; i++) {
        if (bases[tipCoord - i] != ref[originalStop - i]) { originalMismatches++; lastMismatch = i; contig = 0; }
        else contig++;
    }
    if (originalMismatches < 3) return 0;
    tiplen = lastMismatch + 1;
    if (tiplen < 4) return 0;
    searchDist = imin(searchDist, 30 * originalMismatches);
    const int lastIndexToStart = imin(refLen - 1, originalStop + searchDist);
    int bestCnt = INT_MAX, bestOrd = INT_MAX;
    for (int start = originalStop + 1 + lane; start <= lastIndexToStart; start += 32) {
        int m = 0;
        for (int j = 0; j < tiplen; j++) m += (bases[tipCoord - j] != ref[start - j]) ? 1 : 0;
        const int ord = start - originalStop;
        if (m < bestCnt) { bestCnt = m; bestOrd = ord; }         // per lane the order is ascending, so strict < keeps the first
    }
    warp_argmin(bestCnt, bestOrd);
    const int minMismatches = imin(originalMismatches, bestCnt);
    if (minMismatches > 2 || originalMismatches - minMismatches < 2) return 0;
    return bestOrd;                                              // bestCnt < originalMismatches here, so a start was accepted
}

// findTipDeletionsLeft (:2238-2294)
__device__ int tipdel_left(const int8_t* __restrict__ bases, const int8_t* __restrict__ ref, int refLen, int minIndex,
                           int originalStart, int searchDist, int tiplen, int lane) {
    if (originalStart + tiplen >= refLen) return 0;
    if (minIndex >= originalStart) return 0;
    int lastMismatch = 0, originalMismatches = 0, contig = 0;
    for (int i = 0; i < tiplen && contig < 5; i++) {
        if (bases[i] != ref[originalStart + i]) { originalMismatches++; lastMismatch = i; contig = 0; }
        else contig++;
    }
    if (originalMismatches < 3) return 0;
    tiplen = lastMismatch + 1;
    if (tiplen < 4) return 0;
    searchDist = imin(searchDist, 16 + 16 * originalMismatches + 8 * tiplen);
    const int lastIndexToStart = imax(minIndex, originalStart - searchDist);
    int bestCnt = INT_MAX, bestOrd = INT_MAX;
    for (int start = originalStart - 1 - lane; start >= lastIndexToStart; start -= 32) {
        int m = 0;
        for (int j = 0; j < tiplen; j++) m += (bases[j] != ref[start + j]) ? 1 : 0;
        const int ord = originalStart - start;
        if (m < bestCnt) { bestCnt = m; bestOrd = ord; }
    }
    warp_argmin(bestCnt, bestOrd);
    const int minMismatches = imin(originalMismatches, bestCnt);
    if (minMismatches > 2 || originalMismatches - minMismatches < 2) return 0;
    return bestOrd;
}

__global__ void __launch_bounds__(128) tipdel_kernel(const int8_t* __restrict__ reads, const int8_t* __restrict__ refs,
                                                     const bbm_tipdel_task* __restrict__ tasks, long long n, bbm_tipdel_cfg cfg,
                                                     bbm_tipdel_out* __restrict__ outs) {
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n) return;
    const bbm_tipdel_task T = tasks[t];
    const int8_t* bases = reads + T.read_off;
    const int8_t* ref = refs + T.ref_off;
    const int len = T.read_len;
    int start = T.start, stop = T.stop, right = 0, left = 0;
    // findTipDeletions(SiteScore, ...) :1107-1141
    bool go = !(T.slow_score >= T.max_imperfect) && !(len <= 2 * cfg.max_tiplen);
    int maxSearch = cfg.search_range;
    if (go) {
        maxSearch = imin(maxSearch, cfg.align_columns - (cfg.slow_rescue_padding + 8 + imax(len, stop - start)));
        if (maxSearch < 1) go = false;
    }
    if (go && (T.flags & 1)) {
        const int x = tipdel_right(bases, len, ref, T.ref_len, T.min_index, stop, maxSearch, cfg.max_tiplen, lane);
        if (x > 0) {
            stop += x; right = x;
            maxSearch = imin(maxSearch, cfg.align_columns - (cfg.slow_rescue_padding + 8 + imax(len, stop - start)));
            if (maxSearch < 1) go = false;
        }
    }
    if (go && (T.flags & 2)) {
        const int y = tipdel_left(bases, ref, T.ref_len, T.min_index, start, maxSearch, cfg.max_tiplen, lane);
        if (y > 0) { start -= y; left = y; }
    }
    if (lane == 0) { bbm_tipdel_out o; o.start = start; o.stop = stop; o.right = right; o.left = left; outs[t] = o; }
}

// ---------------- quickRescue (:2303-2405) ----------------
__global__ void __launch_bounds__(128) rescue_kernel(const int8_t* __restrict__ reads, const int8_t* __restrict__ refs,
                                                     const bbm_rescue_task* __restrict__ tasks, long long n, bbm_rescue_cfg cfg,
                                                     bbm_rescue_out* __restrict__ outs) {
    const long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= n) return;
    const bbm_rescue_task T = tasks[t];
    const int8_t* __restrict__ bases = reads + T.read_off;
    const int8_t* __restrict__ ref = refs + T.ref_off;
    const int len = T.read_len, refLen = T.ref_len, idealStart = T.ideal_start;
    bbm_rescue_out o; o.start = -1; o.stop = -1; o.mismatches = 0; o.max_contig = 0; o.score = 0; o.perfect = 0; o.in_bounds = 0; o.pad_ = 0;
    if (len < 10) { if (lane == 0) outs[t] = o; return; }
    const bool searchRight = (T.flags & 1) != 0;
    int lowerBound, upperBound;
    if (searchRight) { lowerBound = imax(T.min_index, T.loc); upperBound = imin(refLen - len, T.loc + T.search_dist); }
    else { lowerBound = imax(T.min_index, T.loc - T.search_dist); upperBound = imin(refLen - len, T.loc); }
    int minMismatches = T.max_mismatches + 1;
    int maxContigMatches = 0, bestScore = 0, bestStart = -1, bestAbsdif = INT_MAX;
    // scan order: ascending from lowerBound (searchRight) or descending from upperBound
    const int dir = searchRight ? 1 : -1;
    int base = searchRight ? lowerBound : upperBound;
    for (;;) {
        // chunk of 32 starts in scan order (bounds are uniform across the warp)
        if (searchRight ? (base > upperBound) : (base < lowerBound)) break;
        const int start = base + dir * lane;
        const bool inRange = searchRight ? (start <= upperBound) : (start >= lowerBound);
        int mismatches = 0, contig = 0;
        if (inRange) {
            int currentContig = 0;
            const int bound = minMismatches;
            for (int j = 0; j < len && mismatches <= bound; j++) {
                const int8_t c = bases[j], r = ref[start + j];
                if (c != r || c == 'N') { mismatches++; contig = imax(contig, currentContig); currentContig = 0; }
                else currentContig++;
            }
        }
        unsigned cand = __ballot_sync(RFULL, inRange && mismatches <= minMismatches);
        while (cand) {
            const int L = __ffs(cand) - 1;
            const int mL = __shfl_sync(RFULL, mismatches, L), cL = __shfl_sync(RFULL, contig, L), sL = __shfl_sync(RFULL, start, L);
            cand &= cand - 1;
            if (searchRight ? (sL > upperBound) : (sL < lowerBound)) { cand = 0; break; }     // the bound shrank below this start
            const int score = (len - mL) + cL;
            const int ad = sL > idealStart ? sL - idealStart : idealStart - sL;
            if (mL <= minMismatches && (score > bestScore || (score == bestScore && ad < bestAbsdif))) {
                bestStart = sL; minMismatches = mL; maxContigMatches = cL; bestScore = score; bestAbsdif = ad;
                if (mL == 0) { if (searchRight) upperBound = imin(upperBound, idealStart + ad); else lowerBound = imax(lowerBound, idealStart - ad); }
            }
        }
        base += dir * 32;
    }
    if (bestStart >= 0) {
        o.start = bestStart; o.stop = bestStart + len - 1; o.mismatches = minMismatches; o.max_contig = maxContigMatches;
        o.score = cfg.use_affine ? cfg.points_match + cfg.points_match2 * (len - 1 - minMismatches)
                                 : maxContigMatches + cfg.base_hit_score * (len - minMismatches);
        // SiteScore.setPerfect(bases) for a site exactly as long as the read that starts at >= 0: the flags follow from
        //   nN  = positions with read base 'N'
        //   bad = positions with c != r and (r != 'N' or the N-th reference 'N' beyond nlimit)   -> semiperfect = false at once
        // evaluated by all lanes over strided positions (the early return only changes which of two false values is reported)
        const int nlimit = len / 2;
        int nReadN = 0, nMis = 0, nRefNMis = 0, nHard = 0;
        for (int j = lane; j < len; j += 32) {
            const int8_t c = bases[j], r = ref[bestStart + j];
            if (c != r || c == 'N') {
                nMis++;
                if (c == 'N') nReadN++;
                if (r != 'N') nHard++; else nRefNMis++;
            }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            nReadN += __shfl_xor_sync(RFULL, nReadN, d); nMis += __shfl_xor_sync(RFULL, nMis, d);
            nRefNMis += __shfl_xor_sync(RFULL, nRefNMis, d); nHard += __shfl_xor_sync(RFULL, nHard, d);
        }
        // N counts mismatching positions whose reference base is 'N'; the loop bails out (semiperfect=false) on a hard mismatch or N>nlimit
        const bool semiperfect = (nHard == 0) && (nRefNMis <= nlimit) && (nReadN == 0);
        const bool perfect = semiperfect && (nMis == 0);
        o.perfect = (perfect ? 1 : 0) | (semiperfect ? 2 : 0);
        o.in_bounds = (o.start >= 0 && o.stop <= T.max_index) ? 1 : 0;
    }
    if (lane == 0) outs[t] = o;
}

}  // namespace bbm

extern "C" int bbm_launch_tipdel(const int8_t* reads, const int8_t* refs, const bbm_tipdel_task* tasks, long long n, const bbm_tipdel_cfg* cfg,
                                 bbm_tipdel_out* outs, cudaStream_t st) {
    const long long threads = n * 32;
    bbm::tipdel_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(reads, refs, tasks, n, *cfg, outs);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_rescue(const int8_t* reads, const int8_t* refs, const bbm_rescue_task* tasks, long long n, const bbm_rescue_cfg* cfg,
                                 bbm_rescue_out* outs, cudaStream_t st) {
    const long long threads = n * 32;
    bbm::rescue_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(reads, refs, tasks, n, *cfg, outs);
    return (int)cudaGetLastError();
}
