// search_walk_warp.cuh — slowWalk3 + extendScore + calcAffineScore with one WARP per read (phase 4 of the split launches).
// Included by search.cu inside namespace bbm, after the thread-per-read code whose scalar helpers it reuses.
//
// Reference: BBIndex.find's block/strand loop (current/align2/BBIndex.java:612-636), slowWalk3 (:1219-1706), quickScore / scoreLeft / scoreRight /
// scoreY / scoreZ2 (:2490-2511, 2882-2914, 2967-3035; AbstractIndex.java:52-80), extendScore (:2558-2757), MSA.calcAffineScore
// (MultiStateAligner11tsJNI.java:871-941), makeGapArray (:2837-2878), SiteScore.setPerfect (stream/SiteScore.java:239-292).
//
// Why: the thread-per-read walk runs at 3 active lanes in the list bookkeeping and at 1.1 in extendScore / calcAffineScore (every lane reaches its
// rare extension at a different time), `profiles/r02h`.  Here lane c owns hit list c of the current (block, strand): cursor, end and head value in
// registers.  Per heap step the warp
//   * skips ahead exactly as the thread kernel does (the k-th smallest live head by all-pairs ranking with shuffles; every list below
//     T = h_k - MAX_INDEL2 advances by its own binary search, all lists at once),
//   * finds the smallest (site, column) by a 64-bit min-reduction, counts approxHits with the reference's `chances` cut-off from two ballots,
//   * evaluates quickScore / scoreZ2 from shuffled values,
//   * and — the part that was serial — extends all keys into locArray (shared memory) 32 read positions at a time: per key and direction the
//     reference's loop stops at the first position that is already filled with this location, or filled at all once a mismatch was seen, or
//     mismatching over a filled cell (leftwards: or for any key but the first); with `old`, the match bit and the running mismatch count of a
//     chunk in hand these are two ballots and a prefix population count, and the writes of the positions before the stop go out together.
//   * calcAffineScore is a scan: each position needs locArray[i-1] (shuffle), the last located position before it (highest set bit of a ballot
//     below the lane, carried across chunks) and, for a substitution, the length of the run of unlocated positions it stands in (lowest clear
//     bit of the substitution ballot below the lane) — timeInMode of the reference is exactly that run length, because only a preceding
//     unlocated position lets it grow and every other kind of position restarts it.
// Site emission and subsumption are warp-uniform scalar code (every lane holds the same values; lane 0 writes).  Results are bit-identical to the
// thread-per-read walk and to both CPU restatements (tests/test_search_gpu.py runs all launch variants).
#pragma once

constexpr int WW_WARPS = 4;
struct WwShared {
    int loc[SEARCH_MAX_READ];
    int8_t bases[SEARCH_MAX_READ], bs[SEARCH_MAX_READ];
    int gap[BBM_MAX_GAPS + 2];
};

__device__ __forceinline__ int ww_max(int v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = imax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int ww_min(int v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v = imin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int ww_sum(int v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ int ww_first(unsigned m) { return m ? __ffs(m) - 1 : 32; }

// deletion cost of calcAffineScore for dif = lastLoc - loc + 1 (…JNI.java:889-910)
__device__ __forceinline__ int ww_del_cost(int dif) {
    int s = 0;
    if (dif > MINGAP) { const int rem = dif % GAPLEN, div = (dif - GAPBUFFER2) / GAPLEN; s += div * -2; dif = rem + GAPBUFFER2; }
    if (dif > 80) { s += ((dif - 80 + 3) / 4) * -1; dif = 80; }
    if (dif > 20) { s += (dif - 20) * -1; dif = 20; }
    if (dif > 5) { s += (dif - 5) * -9; dif = 5; }
    if (dif > 1) s += (dif - 1) * -33;
    return s;
}

// MSA.calcAffineScore over sh.loc / sh.bs, 32 positions per step
__device__ int ww_calc_affine(const WwShared& sh, int len, int lane, unsigned lt) {
    int score = 0, carryLastLoc = -3, carryLastValue = -1, carryRun = 0;
    for (int base = 0; base < len; base += 32) {
        const int i = base + lane;
        const bool in = i < len;
        const int loc = in ? sh.loc[i] : 0;               // 0 = "neither located nor a substitution": contributes nothing
        const bool pos = in && loc > 0, sub = in && loc == -1;
        int lastValue = __shfl_up_sync(FULL, loc, 1);
        if (lane == 0) lastValue = carryLastValue;
        const unsigned posMask = __ballot_sync(FULL, pos), subMask = __ballot_sync(FULL, sub);
        const unsigned below = posMask & lt;
        const int src = below ? 31 - __clz(below) : 0;
        int lastLoc = __shfl_sync(FULL, loc, src);
        if (!below) lastLoc = carryLastLoc;
        int term = 0;
        if (pos) {
            const int bsv = sh.bs[i];
            if (loc == lastValue) term = 100 + bsv;
            else if (loc == lastLoc || lastLoc < 0) term = 70 + bsv;
            else if (loc < lastLoc) term = 70 + bsv - 472 + ww_del_cost(lastLoc - loc + 1);
            else { const int d = imin(loc - lastLoc, 5); term = 70 + bsv + (d <= 0 ? 0 : (-395 - 39 * (d - 1))); }
        } else if (sub) {
            const unsigned nz = ~subMask & lt;             // positions below me in this chunk that are not substitutions
            const int run = nz ? lane - (31 - __clz(nz)) : lane + 1 + carryRun;
            term = run > 5 ? -25 : (run > 1 ? -51 : -127);
        }
        score += ww_sum(term);
        // carries for the next chunk (the last chunk's values are never used)
        const int hiPos = posMask ? 31 - __clz(posMask) : 0;
        const int lp = __shfl_sync(FULL, loc, hiPos);
        if (posMask) carryLastLoc = lp;
        carryLastValue = __shfl_sync(FULL, loc, 31);
        const unsigned nzAll = ~subMask;                   // run of substitutions ending at lane 31
        carryRun = (subMask >> 31) ? (nzAll ? 31 - (31 - __clz(nzAll)) : 32 + carryRun) : 0;
    }
    return score;
}

// SiteScore.setPerfect(bases) for the site (chrom, start, stop); bases = sh.bases (the strand being walked)
__device__ void ww_set_perfect(const ctx_t* c, const WwShared& sh, int len, int chrom, int start, int stop, int lane, int* perfectOut, int* semiOut) {
    if (len != stop - start + 1) { *perfectOut = 0; *semiOut = 0; return; }
    const int8_t* ref = c->X->chroms + c->X->chrom_off[chrom - 1];
    const int refLen = (int)(c->X->chrom_off[chrom] - c->X->chrom_off[chrom - 1]);
    int perfect = 1, semiperfect = 1, refloc = start, readloc = 0, N = 0;
    const int mx = imin(stop, refLen - 1), nlimit = len / 2;
    if (start < 0) { N -= start; readloc -= start; refloc -= start; perfect = 0; }
    if (stop >= refLen) { N += (stop - refLen + 1); perfect = 0; }
    if (N > nlimit) { *perfectOut = 0; *semiOut = 0; return; }
    // the loop (:270-284) leaves early only with semiperfect=false, so its outcome is a function of three counts
    int anyMis = 0, anyReadN = 0, anyRefNotN = 0, refNCount = 0;
    for (int o = lane; refloc + o <= mx; o += 32) {
        const int8_t cb = sh.bases[readloc + o], r = ref[refloc + o];
        if (cb != r || cb == 'N') { anyMis = 1; if (cb == 'N') anyReadN = 1; if (r != 'N') anyRefNotN = 1; else refNCount++; }
    }
    anyMis = __any_sync(FULL, anyMis); anyReadN = __any_sync(FULL, anyReadN); anyRefNotN = __any_sync(FULL, anyRefNotN); refNCount = ww_sum(refNCount);
    if (anyMis) perfect = 0;
    if (anyReadN || anyRefNotN) semiperfect = 0;
    N += refNCount;
    semiperfect = (semiperfect && (N <= nlimit));
    perfect = (perfect && semiperfect && (N == 0));
    *perfectOut = perfect; *semiOut = semiperfect;
}

__global__ void __launch_bounds__(WW_WARPS * 32) walk_warp_kernel(SearchParams P) {
    __shared__ WwShared shAll[WW_WARPS];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const unsigned lt = (1u << lane) - 1u;
    WwShared& sh = shAll[wib];
    const SearchIndex* X = &P.X;
    const bbm_index_cfg* g = X->cfg;
    ctx_t cc; ctx_t* c = &cc;
    c->X = X; c->K = g->keylen; c->baseKeyHitScore = BASE_HIT_SCORE * c->K;
    c->indelPenalty = (c->baseKeyHitScore / 2) - 1;
    c->maxPenaltyMisaligned = c->baseKeyHitScore - (1 + c->baseKeyHitScore / 8);
    c->scoreZ1Key = Z_SCORE_MULT * c->K;
    c->shift = g->shift_length; c->cpb = g->chroms_per_block; c->lowMask = c->cpb - 1; c->highMask = ~c->lowMask;
    c->siteMask = (int)(0xFFFFFFFFu >> (g->chrombits + 1));
    const int K = c->K, MK = P.maxKeys, quit2 = P.quitAfterTwoPerfects;
    for (;;) {
        unsigned r = 0;
        if (lane == 0) r = atomicAdd(P.counter, 1u);
        r = __shfl_sync(FULL, r, 0);
        if ((long long)r >= P.nreads) break;
        int* mid = P.mid + (long long)r * P.midStride;
        const int n = mid[0];
        if (n < 1 || n > 32) continue;                       // nothing to search / left to the thread-per-read launch (mid[12] stays 0)
        bbm_search_head* H = P.heads + r;
        int status = mid[2];
        __syncwarp();
        if (lane == 0) mid[12] = 1;
        if (mid[9]) { if (lane == 0) { H->status = status; H->nsites = 0; } continue; }         // the prescan ruled the read out (:592-593)
        const long long ro = P.read_off[r];
        const int len = (int)(P.read_off[r + 1] - ro);
        const int8_t* basesP = P.bases + ro; const int8_t* baseScoresP = P.baseScores + ro;
        const int* a = mid + MID_HDR;
        int keyP = -1, keyM = -1, ofsP = 0, ofsM = 0, kscP = 0, kscM = 0;
        if (lane < n) { keyP = a[lane]; keyM = a[MK + lane]; ofsP = a[2 * MK + lane]; ofsM = a[3 * MK + lane]; kscP = a[4 * MK + lane]; kscM = a[5 * MK + lane]; }
        int maxQuickScore;
        {
            const int x = ww_sum(lane < n ? kscP : 0);
            int score = 0, a0 = -1, b0 = -1;
            for (int i = 0; i < n; i++) { const int av = __shfl_sync(FULL, ofsP, i); if (b0 < av) { score += b0 - a0; a0 = av; } b0 = av + K; }
            score += b0 - a0;
            maxQuickScore = x + score * Z_SCORE_MULT + Y_SCORE_MULT * (__shfl_sync(FULL, ofsP, n - 1) - __shfl_sync(FULL, ofsP, 0));
        }
        const int ofsFirst = __shfl_sync(FULL, ofsP, 0), ofsLast = __shfl_sync(FULL, ofsP, n - 1);
        const int ofsPrev = __shfl_up_sync(FULL, ofsP, 1);
        const bool gapBad = __any_sync(FULL, lane >= 1 && lane < n && ofsP > ofsPrev + K);
        const bool allBasesCovered = (ofsFirst == 0) && (ofsLast == (len - K)) && !gapBad;
        int bestScores[6] = {0, mid[4], 0, mid[5], 0, 0};
        const int hitsCutoff = mid[6], qscoreCutoff = mid[7];
        const bool havePre = mid[8] != 0;
        const int* midPre = mid + MID_HDR + 6 * MK;
        const int ncyc = 2 * X->nblocks;
        // maxScore = msa.maxQuality(baseScores), fullyDefined (:609-610)
        int maxScore, fullyDefined;
        {
            int s = 0, undef = 0;
            for (int i = lane; i < len; i += 32) { s += baseScoresP[i]; undef |= base_defined(basesP[i]) ? 0 : 1; }
            maxScore = 70 + (len - 1) * 100 + ww_sum(s);
            fullyDefined = __any_sync(FULL, undef) ? 0 : 1;
        }
        if (lane == 0) { H->max_score = maxScore; H->max_quick_score = maxQuickScore; }
        bbm_site* sites = P.sites + (long long)r * P.maxSites;
        int nsites = 0;
        int cycle = 0; bool doneRead = false;
        int stagedStrand = -1;
        for (int chromB = 1; chromB <= X->nchroms && !doneRead; chromB = ((chromB & c->highMask) + c->cpb)) {
            const int baseChrom = base_chrom(c, chromB);
            const SearchBlock* b = block_of(c, chromB);
            for (int strand = 0; strand < 2 && !doneRead; strand++) {
                const bool searchIt = !havePre || midPre[cycle] >= hitsCutoff || midPre[ncyc + cycle] >= qscoreCutoff;
                if (searchIt) {
                    // ---- getHits (:353-373) ----
                    const int key = strand == 0 ? keyP : keyM;
                    int st = -1, sp = -1;
                    if (lane < n && key >= 0 && X->counts[key] > 0) {
                        const int s0 = b->starts[key], x = b->starts[key + 1] - s0;
                        if (x > 0 && b->sites[s0] != -1) { st = s0; sp = s0 + x; }
                    }
                    const unsigned vmask = __ballot_sync(FULL, st >= 0);
                    const int numHits = __popc(vmask), numKeys = n;
                    if (numHits >= 1) {
                        // ======================= slowWalk3 =======================
                        if (stagedStrand != strand) {          // bases / baseScores of the strand being walked (Tools.reverseAndCopy, reverseComplementBases)
                            __syncwarp();
                            for (int i = lane; i < len; i += 32) {
                                sh.bases[i] = strand == 0 ? basesP[i] : comp_base(basesP[len - 1 - i]);
                                sh.bs[i] = strand == 0 ? baseScoresP[i] : baseScoresP[len - 1 - i];
                            }
                            stagedStrand = strand;
                            __syncwarp();
                        }
                        const int src = lane < numHits ? __fns(vmask, 0, lane + 1) : 0;       // shrink (:783-813)
                        int row = __shfl_sync(FULL, st, src); const int stop = __shfl_sync(FULL, sp, src);
                        const int ofs = __shfl_sync(FULL, strand == 0 ? ofsP : ofsM, src), ksc = __shfl_sync(FULL, strand == 0 ? kscP : kscM, src);
                        const bool isCol = lane < numHits;
                        const unsigned colMask = numHits >= 32 ? FULL : ((1u << numHits) - 1u);
                        const int filter_by_qscore = (numKeys >= 5);
                        const int minScore = (int)(MIN_SCORE_MULT * maxScore);
                        const int minQuickScore = (int)(MIN_QSCORE_MULT * maxQuickScore);
                        int currentTopScore = bestScores[0];
                        int cutoff = imax(minScore, (int)(currentTopScore * DYNAMIC_SCORE_THRESH));
                        int qcutoff = imax(bestScores[2], minQuickScore);
                        int bestqscore = bestScores[3], maxHits = bestScores[1], perfectsFound = bestScores[5];
                        int approxHitsCutoff = approx_hits_cutoff(c, numKeys, maxHits, 1, currentTopScore >= maxScore);
                        if (approxHitsCutoff <= numHits) {
                            const int shortCircuit = (allBasesCovered && numKeys == numHits && filter_by_qscore);
                            if (currentTopScore >= maxScore) qcutoff = imax(qcutoff, (int)(maxQuickScore * DYNAMIC_QSCORE_THRESH_PERFECT));
                            int val = 0; bool live = isCol;
                            if (isCol) val = site_minus_offset(c, b->sites[row], ofs, baseChrom);
                            const int total = ww_sum(isCol ? stop - row : 0);
                            const bool longLists = total >= 4 * numHits;
                            int nActive = numHits;
                            // prevSS: a copy in every lane; lane 0 writes changes through to sites[prevIdx]
                            int prevIdx = -1, pChrom = 0, pStart = 0, pStop = 0, pScore = 0, pPerfect = 0, pSemi = 0, pNgaps = 0;
                            bool quit = false;
                            while (!quit) {
                                // ---------------- exact skip-ahead (see skip_ahead) ----------------
                                if (longLists && approxHitsCutoff >= 2 && nActive >= approxHitsCutoff) {
                                    const unsigned liveMask = __ballot_sync(FULL, live);
                                    const int s = ww_min(live ? val : 0x7fffffff);
                                    const int staleMax = ww_max((isCol && !live) ? val : (-0x7fffffff - 1));
                                    if (!(staleMax >= s - MAX_INDEL)) {
                                        int rank = 0;
                                        for (int j = 0; j < numHits; j++) {
                                            const int vj = __shfl_sync(FULL, val, j);
                                            if (((liveMask >> j) & 1u) && (vj < val || (vj == val && j < lane))) rank++;
                                        }
                                        const unsigned kth = __ballot_sync(FULL, live && rank == approxHitsCutoff - 1);
                                        const int hk = __shfl_sync(FULL, val, kth ? __ffs(kth) - 1 : 0);
                                        if (kth && !(hk < -0x40000000 + MAX_INDEL2)) {
                                            const int T = hk - MAX_INDEL2;
                                            if (T > s) {
                                                bool exhausted = false;
                                                if (live && val < T) {
                                                    int lo = row + 1, hi = stop;
                                                    while (lo < hi) { const int m2 = lo + ((hi - lo) >> 1); if (site_minus_offset(c, b->sites[m2], ofs, baseChrom) >= T) hi = m2; else lo = m2 + 1; }
                                                    if (lo < stop) { row = lo; val = site_minus_offset(c, b->sites[lo], ofs, baseChrom); }
                                                    else { val = site_minus_offset(c, b->sites[stop - 1], ofs, baseChrom); row = stop; live = false; exhausted = true; }
                                                }
                                                nActive -= __popc(__ballot_sync(FULL, exhausted));
                                                if (nActive < approxHitsCutoff) break;
                                            }
                                        }
                                    }
                                }
                                // ---------------- heap.peek(): smallest (site, column) ----------------
                                long long hk64 = live ? (((long long)val << 6) | lane) : 0x7fffffffffffffffLL;
#pragma unroll
                                for (int o = 16; o >= 1; o >>= 1) { const long long t = __shfl_xor_sync(FULL, hk64, o); hk64 = t < hk64 ? t : hk64; }
                                if (hk64 == 0x7fffffffffffffffLL) break;
                                const int centerIndex = (int)(hk64 & 63), site = (int)(hk64 >> 6);
                                // ---------------- approxHits with the `chances` cut-off (:1340-1352) ----------------
                                int maxNearbySite = site, approxHits = 0;
                                {
                                    const int minsite = site - MAX_INDEL, maxsite = site + MAX_INDEL2;
                                    const unsigned inr = __ballot_sync(FULL, isCol && val >= minsite && val <= maxsite);
                                    const int chances0 = numHits - approxHitsCutoff;
                                    const bool examined = isCol && __popc(~inr & colMask & lt) <= chances0;
                                    const unsigned cnt = __ballot_sync(FULL, examined && ((inr >> lane) & 1u));
                                    approxHits = __popc(cnt);
                                    maxNearbySite = imax(site, ww_max(((cnt >> lane) & 1u) ? val : (-0x7fffffff - 1)));
                                }
                                if (approxHits >= approxHitsCutoff) {
                                    int score;
                                    // ---- quickScore (:2490-2511) + scoreZ2 (:2882-2914) ----
                                    int qscore;
                                    if (approxHits == 1) qscore = (filter_by_qscore ? __shfl_sync(FULL, ksc, centerIndex) : qcutoff) + c->scoreZ1Key;
                                    else {
                                        int z = 0, a0 = -1, b0 = -1;
                                        const int maxLoc = site + MAX_INDEL2, minLoc = imax(0, site - MAX_INDEL);
                                        for (int i = 0; i < numHits; i++) {
                                            const int v = __shfl_sync(FULL, val, i), av = __shfl_sync(FULL, ofs, i);
                                            if (v >= minLoc && v <= maxLoc) { if (b0 < av) { z += b0 - a0; a0 = av; } b0 = av + K; }
                                        }
                                        z += b0 - a0;
                                        if (filter_by_qscore) {
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
                                            for (int i = centerIndex + 1; i < numHits; i++) {
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
                                            qscore = sc + z * Z_SCORE_MULT;
                                        } else qscore = qcutoff + z * Z_SCORE_MULT;
                                    }
                                    int mapStart = site, mapStop = maxNearbySite;
                                    if (qscore < qcutoff) score = -1;
                                    else {
                                        const int chrom = number_to_chrom(c, site, baseChrom);
                                        if (shortCircuit && qscore == maxQuickScore) score = maxScore;
                                        else {
                                            // ================= extendScore (:2558-2757), 32 read positions at a time =================
                                            const int centerVal = site, centerLoc = number_to_site(c, centerVal);
                                            const int minVal = centerVal - MAX_INDEL, maxVal = centerVal + MAX_INDEL2;
                                            const int8_t* ref = X->chroms + X->chrom_off[chrom - 1];
                                            const int refLen = (int)(X->chrom_off[chrom] - X->chrom_off[chrom - 1]);
                                            __syncwarp();
                                            for (int i = lane; i < len; i += 32) sh.loc[i] = -1;
                                            __syncwarp();
                                            for (int i = 0, keynum = 0; i < numHits; i++) {                  // leftwards from the end of each key
                                                const int value = __shfl_sync(FULL, val, i);
                                                if (!(value >= minVal && value <= maxVal)) continue;
                                                const int refbase = number_to_site(c, value), callbase = __shfl_sync(FULL, ofs, i);
                                                keynum++;
                                                int misses = 0;
                                                for (int top = callbase + K - 1; top >= 0; top -= 32) {
                                                    const int cloc = top - lane, rloc = refbase + cloc;
                                                    const bool valid = cloc >= 0 && rloc >= 0 && rloc < refLen;
                                                    const int old = valid ? sh.loc[cloc] : 0;
                                                    const bool match = valid && sh.bases[cloc] == ref[rloc];
                                                    const unsigned mm = __ballot_sync(FULL, valid && !match);
                                                    const int before = misses + __popc(mm & lt);
                                                    const unsigned pre = __ballot_sync(FULL, !valid || old == refbase || (before > 0 && old >= 0));
                                                    const unsigned post = __ballot_sync(FULL, valid && !match && (old >= 0 || keynum > 1));
                                                    const int nproc = imin(ww_first(pre), ww_first(post) + 1);       // positions 0..nproc-1 of this chunk are processed
                                                    if (lane < nproc && match && (old < 0 || refbase == centerLoc)) sh.loc[cloc] = refbase;
                                                    misses += __popc(mm & (nproc >= 32 ? FULL : ((1u << nproc) - 1u)));
                                                    __syncwarp();
                                                    if (pre | post) break;                                           // a stop anywhere in the chunk (also after its last position) ends this key's loop
                                                }
                                            }
                                            for (int i = 0; i < numHits; i++) {                                  // rightwards from behind each key
                                                const int value = __shfl_sync(FULL, val, i);
                                                if (!(value >= minVal && value <= maxVal)) continue;
                                                const int refbase = number_to_site(c, value), callbase = __shfl_sync(FULL, ofs, i);
                                                int misses = 0;
                                                for (int bot = callbase + K; bot < len; bot += 32) {
                                                    const int cloc = bot + lane, rloc = refbase + cloc;
                                                    const bool valid = cloc < len && rloc < refLen;
                                                    const int old = valid ? sh.loc[cloc] : 0;
                                                    const bool match = valid && sh.bases[cloc] == ref[rloc];
                                                    const unsigned mm = __ballot_sync(FULL, valid && !match);
                                                    const int before = misses + __popc(mm & lt);
                                                    const unsigned pre = __ballot_sync(FULL, !valid || old == refbase || (before > 0 && old >= 0));
                                                    const unsigned post = __ballot_sync(FULL, valid && !match && old >= 0);
                                                    const int nproc = imin(ww_first(pre), ww_first(post) + 1);
                                                    if (lane < nproc && match && (old < 0 || refbase == centerLoc)) sh.loc[cloc] = refbase;
                                                    misses += __popc(mm & (nproc >= 32 ? FULL : ((1u << nproc) - 1u)));
                                                    __syncwarp();
                                                    if (pre | post) break;                                           // a stop anywhere in the chunk (also after its last position) ends this key's loop
                                                }
                                            }
                                            int mn = 0x7fffffff, mx = (-0x7fffffff - 1);
                                            for (int i = lane; i < len; i += 32) {
                                                if (sh.bases[i] == 'N') sh.loc[i] = -2;
                                                const int x = sh.loc[i];
                                                if (x > -1) { mn = imin(mn, x); mx = imax(mx, x); }
                                            }
                                            __syncwarp();
                                            mn = ww_min(mn); mx = ww_max(mx);
                                            score = ww_calc_affine(sh, len, lane, lt);
                                            if (mn < 0 || mx < 0) { score = -99999; status |= BBM_ST_ANOMALY; }
                                            mapStart = to_number(c, mn, chrom); mapStop = to_number(c, mx, chrom);
                                        }
                                        if (score == maxScore) {
                                            qcutoff = imax(qcutoff, (int)(maxQuickScore * DYNAMIC_QSCORE_THRESH_PERFECT));
                                            approxHitsCutoff = approx_hits_cutoff(c, numKeys, maxHits, 1, 1);
                                        }
                                        if (score >= cutoff) { qcutoff = imax(qcutoff, (int)(qscore * DYNAMIC_QSCORE_THRESH)); bestqscore = imax(qscore, bestqscore); }
                                    }
#ifdef WW_DEBUG_READ
                                    if (r == WW_DEBUG_READ && lane == 0) printf("WW read %u strand %d site %d center %d approx %d cutoffHits %d qscore %d qcutoff %d score %d cutoff %d top %d\n", r, strand, site, centerIndex, approxHits, approxHitsCutoff, qscore, qcutoff, score, cutoff, currentTopScore);
#endif
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
                                        int ngap = 0;
                                        if (site3 - site2 >= MINGAP + len) {                                 // makeGapArray: rare (spliced reads); one lane, result through shared memory
                                            __syncwarp();
                                            if (lane == 0) {
                                                int ov = 0;
                                                int ng = make_gap_array(sh.loc, len, site2, MINGAP, sh.gap, BBM_MAX_GAPS - 1, &ov);
                                                if (ng > 0) { sh.gap[0] = imin(sh.gap[0], site2); sh.gap[ng - 1] = imax(sh.gap[ng - 1], site3); }
                                                sh.gap[BBM_MAX_GAPS] = ng; sh.gap[BBM_MAX_GAPS + 1] = ov;
                                            }
                                            __syncwarp();
                                            ngap = sh.gap[BBM_MAX_GAPS];
                                            if (sh.gap[BBM_MAX_GAPS + 1]) status |= BBM_ST_GAP_OVERFLOW;
                                        }
                                        const int perfect1 = (score == maxScore && fullyDefined);
                                        const int chromLen = (int)(X->chrom_off[chrom] - X->chrom_off[chrom - 1]);
                                        const int inbounds = (site2 >= 0 && site3 < chromLen);
                                        const bool haveP = prevIdx >= 0;
                                        const bool overlapP = haveP && pChrom == chrom && (site2 <= pStop && site3 >= pStart);      // prevSS was made in this walk: same strand
                                        int made = -1;
                                        if (inbounds && ngap == 0 && overlapP) {
                                            const int betterScore = imax(score, pScore);
                                            const int minStart = imin(pStart, site2), maxStop = imax(pStop, site3);
                                            const int perfect2 = (pScore == maxScore && fullyDefined);
                                            const int shortEnough = (maxStop - minStart < 2 * len);
                                            bool changed = true;
                                            if (pStart == site2 && pStop == site3) {
                                                pScore = betterScore;
                                                pPerfect = (pPerfect || perfect1 || perfect2) ? 1 : 0;
                                                if (pPerfect) pSemi = 1;
                                            } else if (shortEnough && pStart == site2 && !pSemi) {
                                                if (pNgaps) status |= BBM_ST_GAPFIX;
                                                if (perfect2) {}
                                                else if (perfect1) { pStop = site3; if (!pPerfect) perfectsFound++; pPerfect = pSemi = 1; }
                                                else { pStop = maxStop; ww_set_perfect(c, sh, len, pChrom, pStart, pStop, lane, &pPerfect, &pSemi); }
                                                pScore = betterScore;
                                            } else if (shortEnough && pStop == site3 && !pSemi) {
                                                if (pNgaps) status |= BBM_ST_GAPFIX;
                                                if (perfect2) {}
                                                else if (perfect1) { pStart = site2; if (!pPerfect) perfectsFound++; pPerfect = pSemi = 1; }
                                                else { pStart = minStart; ww_set_perfect(c, sh, len, pChrom, pStart, pStop, lane, &pPerfect, &pSemi); }
                                                pScore = betterScore;
                                            } else { made = 1; changed = false; }
                                            if (changed && lane == 0) {
                                                bbm_site* S = &sites[prevIdx];
                                                S->start = pStart; S->stop = pStop; S->score = pScore; S->perfect = (int8_t)pPerfect; S->semiperfect = (int8_t)pSemi;
                                            }
                                        } else if (inbounds) made = 1;
                                        if (made > 0) {
                                            if (nsites >= P.maxSites) status |= BBM_ST_SITE_OVERFLOW;
                                            else {
                                                int sPerfect = perfect1, sSemi = perfect1;
                                                if (!perfect1) ww_set_perfect(c, sh, len, chrom, site2, site3, lane, &sPerfect, &sSemi);
                                                const bool attachGaps = !(haveP && inbounds && ngap == 0 && overlapP);
                                                const int sNgaps = attachGaps ? ngap : 0;
                                                if (lane == 0) {
                                                    bbm_site* S = &sites[nsites];
                                                    { bbm_site z = {}; *S = z; }
                                                    S->chrom = chrom; S->strand = (int8_t)strand; S->start = site2; S->stop = site3; S->hits = approxHits; S->score = score;
                                                    S->perfect = (int8_t)sPerfect; S->semiperfect = (int8_t)sSemi;
                                                    S->ngaps = sNgaps; for (int q = 0; q < sNgaps; q++) S->gaps[q] = sh.gap[q];
                                                }
                                                const int idx = nsites++;
                                                if (sPerfect) {
                                                    const bool overlapsPrev = haveP && pChrom == chrom && (site2 <= pStop && site3 >= pStart);
                                                    if (!haveP || !pPerfect || !overlapsPrev) {
                                                        perfectsFound++;
                                                        if (quit2 && perfectsFound >= 2) quit = true;
                                                    }
                                                }
                                                prevIdx = idx; pChrom = chrom; pStart = site2; pStop = site3; pScore = score; pPerfect = sPerfect; pSemi = sSemi; pNgaps = sNgaps;
                                            }
                                        }
                                    }
                                }
                                if (quit) break;
                                // ---------------- pops: every live list sitting on `site` advances (:1640-1690) ----------------
                                bool ret = false;
                                for (;;) {
                                    const bool hit = live && val == site;
                                    if (!__any_sync(FULL, hit)) break;
                                    bool exhausted = false;
                                    if (hit) {
                                        row++;
                                        if (row < stop) val = site_minus_offset(c, b->sites[row], ofs, baseChrom);
                                        else { live = false; exhausted = true; }
                                    }
                                    nActive -= __popc(__ballot_sync(FULL, exhausted));
                                    if (nActive < approxHitsCutoff) { ret = true; break; }
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
                    }
                }
                cycle++;
                if (quit2 && bestScores[5] >= 2) doneRead = true;
            }
        }
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < 6; i++) H->best_scores[i] = bestScores[i];
            H->status = status; H->nsites = nsites;
        }
    }
}
