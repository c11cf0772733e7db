// pairing.cu — the pair logic of BBMapThread.processReadPair (current/align2/BBMapThread.java:943-1362) on the device, one thread per pair:
//   pairSiteScoresInitial :736-940 + the paired trimList / score reset :988-1017
//   the rescue block :1061-1100 with AbstractMapThread.rescue / slowRescue (current/align2/AbstractMapThread.java:1144-1306): per anchor site one
//     quickRescue scan (rescue.cu, warp per task) and, for what it finds, one fillAndScoreLimited (the batched aligner); the sites of a pair are
//     appended to the mate's list in anchor order, then Tools.mergeDuplicateSites
//   Tools.removeLowQualitySitesPaired (Tools.java:934-958), pairSiteScoresFinal / canPair (AbstractMapThread.java:1919-2170), the paired clearzone
//     rule :1147-1176, Read.setFromTopSite, Read.isBadPair (stream/Read.java:1305-1331)
//   after genMatchString: the anomaly blocks, removeDuplicateBestSites, AMBIGUOUS_TOSS, toLocalAlignment for X/Y/C tips, statistics :1228-1352
// Reads 2p and 2p+1 are the mates of pair p.  The two rescue directions are sequential (the second one sees the list the first one extended).
#include <cuda_runtime.h>
#include "sitelist_dev.cuh"
#include "mapper_kernels.cuh"
#include "genmatch_dev.cuh"

namespace bbm {

// Tools.trimSitesBelowCutoff(list, cutoff, retainPaired, retainSemiperfect = true, minS, maxS)
static __device__ int trim_below_cutoff(bbm_ss* v, int n, int cutoff, bool retainPaired, int minS, int maxS) {
    if (n <= minS) return n;
    while (n > maxS) n--;
    int removed = 0; const int maxToRemove = n - minS;
    unsigned long long dead = 0;
    for (int i = n - 1; i >= 0; i--) {
        if (!v[i].semiperfect && v[i].score < cutoff && (!retainPaired || v[i].paired_score <= 0)) {
            dead |= 1ull << i; removed++;
            if (removed >= maxToRemove) break;
        }
    }
    return compact(v, n, dead);
}
static __device__ __forceinline__ void pair_dists(const bbm_ss& a, const bbm_ss& b, bool requireCorrect, int& inner, int& outer) {
    const bool first = (requireCorrect && a.strand != b.strand) ? (a.strand == 0) : (a.start <= b.start);
    if (first) { inner = b.start - a.stop; outer = b.stop - a.start; }
    else { inner = a.start - b.stop; outer = a.stop - b.start; }
}
static __device__ int remove_low_quality_paired(bbm_ss* v, int n, int maxSw, float multSingle, float multPaired) {
    if (n == 0) return 0;
    const int th = (int)__fmul_rn((float)maxSw, multSingle), thp = (int)__fmul_rn((float)maxSw, multPaired);
    if (v[0].score < thp) return 0;
    unsigned long long dead = 0;
    for (int i = 0; i < n; i++) if ((v[i].paired_score > 0) ? (v[i].slow_score < thp) : (v[i].slow_score < th)) dead |= 1ull << i;
    return compact(v, n, dead);
}

struct PairCtx {
    const PairParams& P; long long p; bbm_ss* v[2]; int n[2]; int len[2]; int maxSw[2];
};
static __device__ PairCtx pair_ctx(const PairParams& P, long long p) {
    PairCtx X = { P, p, { P.lists + (2 * p) * P.cap, P.lists + (2 * p + 1) * P.cap }, { P.nss[2 * p], P.nss[2 * p + 1] }, {0, 0}, {0, 0} };
    for (int e = 0; e < 2; e++) { X.len[e] = (int)(P.read_off[2 * p + e + 1] - P.read_off[2 * p + e]); X.maxSw[e] = max_quality(X.len[e]); }
    return X;
}

// ---------------- PAIR_INIT ----------------
static __device__ void pair_initial(PairCtx& X) {
    const PairParams& P = X.P; const bbm_map_cfg& cfg = P.cfg;
    bbm_ss* a = X.v[0]; bbm_ss* b = X.v[1];
    int& na = X.n[0]; int& nb = X.n[1];
    const int maxTrim = P.pc.max_trim_sites_to_retain;
    if (na >= 1 && nb >= 1) {
        stable_sort<true>(a, na); stable_sort<true>(b, nb);
        for (int i = 0; i < na; i++) a[i].paired_score = 0;
        for (int i = 0; i < nb; i++) b[i].paired_score = 0;
        int maxPaired1 = -1, maxPaired2 = -1, numPerfectPairs = 0;
        const int ilimit = na - 1, jlimit = nb - 1, maxReadLen = imax(X.len[0], X.len[1]);
        const int outerDistLimit = (maxReadLen * 14) / 32, innerDistLimit = cfg.max_pair_dist;
        const int apd = cfg.average_pair_dist, expectedFragLength = apd + X.len[0] + X.len[1];
        const bool sameStrand = cfg.same_strand_pairs != 0, requireCorrect = cfg.require_correct_strands != 0;
        for (int i = 0, j = 0; i <= ilimit && j <= jlimit; i++) {
            bbm_ss& s1 = a[i];
            while (j < jlimit && (b[j].chrom < s1.chrom || (b[j].chrom == s1.chrom && s1.start - b[j].stop > innerDistLimit))) j++;
            for (int k = j; k <= jlimit; k++) {
                bbm_ss& s2 = b[k];
                if (s2.chrom > s1.chrom) break;
                if (s2.start - s1.stop > innerDistLimit) break;
                int innerdist, outerdist;
                pair_dists(s1, s2, requireCorrect, innerdist, outerdist);
                if (outerdist >= outerDistLimit && innerdist <= innerDistLimit) {
                    const bool strandOK = ((s1.strand == s2.strand) == sameStrand);
                    if (strandOK || !requireCorrect) {
                        bool paired1 = false, paired2 = false;
                        const int deviation = apd > innerdist ? apd - innerdist : innerdist - apd;
                        int ps1, ps2;
                        if (strandOK) {
                            ps1 = s1.score + 1 + imax(1, s2.score / 2 - ((deviation * s2.score) / (32 * expectedFragLength + 100)));
                            ps2 = s2.score + 1 + imax(1, s1.score / 2 - ((deviation * s1.score) / (32 * expectedFragLength + 100)));
                        } else { ps1 = s1.score + imax(0, s2.score / 16); ps2 = s2.score + imax(0, s1.score / 16); }
                        if (ps1 > s1.paired_score) { paired1 = true; s1.paired_score = ps1; maxPaired1 = imax(s1.score, maxPaired1); }
                        if (ps2 > s2.paired_score) { paired2 = true; s2.paired_score = ps2; maxPaired2 = imax(s2.score, maxPaired2); }
                        if (paired1 && paired2 && outerdist >= maxReadLen && deviation <= expectedFragLength && s1.perfect && s2.perfect) numPerfectPairs++;
                    }
                }
            }
        }
        for (int i = 0; i < na; i++) if (a[i].paired_score > a[i].score) a[i].score = a[i].paired_score;
        for (int i = 0; i < nb; i++) if (b[i].paired_score > b[i].score) b[i].score = b[i].paired_score;
        if (numPerfectPairs > 0) {
            na = trim_below_cutoff(a, na, (int)__fmul_rn((float)maxPaired1, .94f), false, 1, maxTrim);
            nb = trim_below_cutoff(b, nb, (int)__fmul_rn((float)maxPaired2, .94f), false, 1, maxTrim);
        } else {
            if (na > 4) na = trim_below_cutoff(a, na, (int)__fmul_rn((float)maxPaired1, .9f), true, 1, maxTrim);
            if (nb > 4) nb = trim_below_cutoff(b, nb, (int)__fmul_rn((float)maxPaired2, .9f), true, 1, maxTrim);
        }
    }
    for (int e = 0; e < 2; e++) {
        if (X.n[e] > 2) stable_sort<false>(X.v[e], X.n[e]);                      // MIN_TRIM_SITES_TO_RETAIN_PAIRED = 2 (BBMapThread.java:63)
        trim_list(X.v[e], X.n[e], true, X.maxSw[e], false, 2, maxTrim);
        for (int i = 0; i < X.n[e]; i++) X.v[e][i].score = X.v[e][i].quick_score;
    }
}

// ---------------- rescue ----------------
// dir 0: read 2p anchors, 2p+1 is loose; dir 1 the other way round
static __device__ void rescue_prep(PairCtx& X, int dir) {
    const PairParams& P = X.P; const bbm_map_cfg& cfg = P.cfg;
    const long long p = X.p;
    int* ps = P.pstate + p * PAIR_STATE;
    if (dir == 0) {
        int u0 = 0, u1 = 0;
        for (int i = 0; i < X.n[0]; i++) u0 += (X.v[0][i].paired_score == 0);
        for (int i = 0; i < X.n[1]; i++) u1 += (X.v[1][i].paired_score == 0);
        ps[PS_UNPAIRED0] = u0; ps[PS_UNPAIRED1] = u1;
    }
    const int e = dir, o = dir ^ 1;
    const long long ra = 2 * p + e, rl = 2 * p + o;
    for (int i = 0; i < P.cap; i++) P.rtask_of[ra * P.cap + i] = -1;
    ps[PS_ACTIVE0 + dir] = 0;
    if (!cfg.do_rescue || !(ps[PS_UNPAIRED0 + e] > 0 && X.n[e] > 0)) return;
    ps[PS_ACTIVE0 + dir] = 1;
    bbm_ss* A = X.v[e]; int& nA = X.n[e];
    stable_sort<false>(A, nA);
    nA = remove_low_quality_paired(A, nA, X.maxSw[e], cfg.min_ratio_pre_rescue, cfg.min_ratio_pre_rescue);
    const int searchDist = imin(cfg.max_pair_dist, 2 * cfg.average_pair_dist + 100);
    if (searchDist > cfg.max_rescue_dist || nA == 0) return;
    const bbm_ss* L = X.v[o]; const int nL = X.n[o];
    const int lenA = X.len[e], lenL = X.len[o];
    const int maxLooseSw = X.maxSw[o], maxAnchorSw = X.maxSw[e], maxImperfect = max_imperfect(lenL);
    const int bestLoose = nL == 0 ? 0 : L[0].slow_score, bestAnchor = A[0].slow_score;
    if (bestLoose == maxLooseSw && bestAnchor == maxAnchorSw && A[0].paired_score > 0) return;
    const int rescueScoreLimit = (int)__fmul_rn(0.95f, (float)bestAnchor);
    const int retain1 = imax((int)__fmul_rn(0.68f, (float)bestLoose), (int)__fmul_rn(0.4f, (float)maxLooseSw));
    const int retain2 = imax((int)__fmul_rn(0.95f, (float)bestLoose), (int)__fmul_rn(0.55f, (float)maxLooseSw));
    const int maxMismatches = (bestLoose > maxImperfect) ? 5 : imin(cfg.max_rescue_mismatches, (int)__fsub_rn(__fmul_rn(0.60f, (float)lenL), 1.f));
    bool findRight = bestLoose < maxImperfect, findLeft = findRight;
    if (findRight && P.quality) {
        const int T = P.tc.max_tiplen;
        const int8_t* q = P.quality + P.read_off[rl];
        int minL = 0, avgL = 0, minF = 0, avgF = 0;
        if (T <= lenL) {
            int x = 0; minL = q[lenL - T];
            for (int i = lenL - T; i < lenL; i++) { const int b = q[i]; x += (b < 0 ? 0 : b); minL = imin(minL, b); }
            avgL = x / T;
            x = 0; minF = q[0];
            for (int i = 0; i < T; i++) { const int b = q[i]; x += (b < 0 ? 0 : b); minF = imin(minF, b); }
            avgF = x / T;
        }
        findRight = (minL >= 6 && avgL >= 14); findLeft = (minF >= 6 && avgF >= 14);
    }
    ps[PS_RETAIN1] = retain1; ps[PS_RETAIN2] = retain2; ps[PS_MAXMM] = maxMismatches; ps[PS_FIND] = (findRight ? 1 : 0) | (findLeft ? 2 : 0);
    for (int ia = 0; ia < nA; ia++) {
        const bbm_ss& ssa = A[ia];
        if (ssa.slow_score < rescueScoreLimit) break;
        if (!(ssa.paired_score == 0 && !ssa.rescued)) continue;
        const int searchIntoAnchor = ssa.stop - ssa.start - 1 + (lenA * 11 / 16);
        int loc, idealStart; bool minusBases;
        const int strand = cfg.same_strand_pairs ? ssa.strand : (ssa.strand ^ 1);
        const bool searchRight = cfg.same_strand_pairs ? (strand == 0) : (strand == 1);
        if (cfg.same_strand_pairs) {
            if (ssa.strand == 1) { minusBases = true; loc = ssa.start + searchIntoAnchor; idealStart = ssa.start - cfg.average_pair_dist; }
            else { minusBases = false; loc = ssa.stop - searchIntoAnchor; idealStart = ssa.stop + cfg.average_pair_dist; }
        } else {
            if (ssa.strand == 0) { minusBases = true; loc = ssa.stop - searchIntoAnchor; idealStart = ssa.stop + cfg.average_pair_dist; }
            else { minusBases = false; loc = ssa.start + searchIntoAnchor; idealStart = ssa.start - cfg.average_pair_dist; }
        }
        const int k = atomicAdd(P.counters + 1, 1);
        if (k >= P.maxTasks) { ps[PS_STATUS0 + o] |= BBM_MAP_ST_LIST_OVERFLOW; continue; }
        const int refLen = (int)(P.chrom_off[ssa.chrom] - P.chrom_off[ssa.chrom - 1]);
        bbm_rescue_task T = {};
        T.read_off = ((minusBases ? P.basesM : P.basesP) - P.basesP) + P.read_off[rl];
        T.ref_off = P.chrom_off[ssa.chrom - 1]; T.read_len = lenL; T.ref_len = refLen; T.min_index = 0; T.max_index = refLen - 1;
        T.loc = loc; T.search_dist = searchDist + searchIntoAnchor; T.ideal_start = idealStart; T.max_mismatches = maxMismatches; T.flags = searchRight ? 1 : 0;
        P.rtasks[k] = T;
        RescueAux a = {};
        a.pair = (int)p; a.dir = dir; a.anchor = ia; a.chrom = ssa.chrom; a.strand = strand; a.minus = minusBases ? 1 : 0; a.msa_req = -1; a.valid = 0;
        P.raux[k] = a;
        P.rtask_of[ra * P.cap + ia] = k;
    }
}

// thread per rescue task, after the quickRescue scan: slowRescue up to its alignment request (:1262-1281)
__global__ void __launch_bounds__(128) rescue_mid_kernel(PairParams P, int ntasks) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= ntasks) return;
    RescueAux a = P.raux[k];
    const bbm_rescue_out O = P.routs[k];
    const bbm_map_cfg& cfg = P.cfg;
    const long long p = a.pair; const int o = a.dir ^ 1;
    const long long rl = 2 * p + o;
    const int* ps = P.pstate + p * PAIR_STATE;
    a.valid = 0; a.msa_req = -1;
    if (O.start >= 0 && O.in_bounds && O.mismatches <= ps[PS_MAXMM]) {
        const int lenL = (int)(P.read_off[rl + 1] - P.read_off[rl]);
        const int8_t* bases = (a.minus ? P.basesM : P.basesP) + P.read_off[rl];
        const int refLen = (int)(P.chrom_off[a.chrom] - P.chrom_off[a.chrom - 1]);
        const int8_t* ref = P.refs + P.chrom_off[a.chrom - 1];
        const int maxLooseSw = max_quality(lenL), maxImperfect = max_imperfect(lenL);
        bbm_ss ss = {};
        ss.chrom = a.chrom; ss.strand = (int8_t)a.strand; ss.start = O.start; ss.stop = O.stop; ss.hits = 0; ss.quick_score = O.score; ss.score = O.score;
        ss.perfect = (O.perfect & 1) ? 1 : 0; ss.semiperfect = (O.perfect & 2) ? 1 : 0; ss.rescued = 1; ss.slow_score = 0; ss.paired_score = 0;
        int sw = score_no_indels(bases, lenL, ref, refLen, ss.start);
        a.old_start = ss.start; a.valid = 1;
        if (sw < maxImperfect && cfg.max_indel > 0) {
            set_slow_score(ss, sw);
            const int find = ps[PS_FIND];
            if (find && ss.slow_score < maxImperfect && lenL > 2 * P.tc.max_tiplen) {                 // findTipDeletions(ss, bases, maxImperfectScore, right, left) :1107-1141
                int maxSearch = imin(P.tc.search_range, P.tc.align_columns - (P.tc.slow_rescue_padding + 8 + imax(lenL, ss.stop - ss.start)));
                bool changed = false;
                if (maxSearch >= 1) {
                    bool go = true;
                    if (find & 1) {
                        const int x = tip_right(bases, lenL, ref, refLen, 0, ss.stop, maxSearch, P.tc.max_tiplen);
                        if (x > 0) {
                            ss_set_stop(ss, ss.stop + x); changed = true;
                            maxSearch = imin(maxSearch, P.tc.align_columns - (P.tc.slow_rescue_padding + 8 + imax(lenL, ss.stop - ss.start)));
                            if (maxSearch < 1) go = false;
                        }
                    }
                    if (go && (find & 2)) { const int y = tip_left(bases, ref, refLen, 0, ss.start, maxSearch, P.tc.max_tiplen); if (y > 0) { ss_set_start(ss, ss.start - y); changed = true; } }
                }
                if (changed) sw = score_no_indels(bases, lenL, ref, refLen, ss.start);
            }
            const int minMsaLimit = -P.clearzone1e + (int)__fmul_rn(cfg.min_ratio_paired, (float)maxLooseSw);
            const int minscore = imax(sw, minMsaLimit);
            const int q = atomicAdd(P.counters + 2, 1);
            bbm_msa_task t = {};
            t.read_off = ((a.minus ? P.basesM : P.basesP) - P.basesP) + P.read_off[rl];
            t.ref_off = P.chrom_off[a.chrom - 1]; t.read_len = lenL; t.ref_len = refLen;
            t.ref_start = ss.start - P.tc.slow_rescue_padding; t.ref_end = ss.stop + P.tc.slow_rescue_padding; t.min_score = minscore;
            t.flags = BBM_TF_CLAMP | BBM_TF_SCORE;
            P.mtasks[q] = t;
            a.msa_req = q;
        }
        a.sw = sw;
        P.rsites[k] = ss;
    }
    P.raux[k] = a;
}

// thread per pair: finish slowRescue for the pair's tasks in anchor order, append, merge (:1188-1204, 1282-1305; processReadPair :1086, :1093)
static __device__ void rescue_apply(PairCtx& X, int dir) {
    const PairParams& P = X.P;
    const long long p = X.p;
    int* ps = P.pstate + p * PAIR_STATE;
    if (!ps[PS_ACTIVE0 + dir]) return;
    const int e = dir, o = dir ^ 1;
    const long long ra = 2 * p + e, rl = 2 * p + o;
    bbm_ss* A = X.v[e]; bbm_ss* L = X.v[o]; int& nL = X.n[o];
    const int lenL = X.len[o], maxLooseSw = X.maxSw[o];
    for (int ia = 0; ia < X.n[e]; ia++) {
        const int k = P.rtask_of[ra * P.cap + ia];
        if (k < 0) continue;
        const RescueAux a = P.raux[k];
        if (!a.valid) continue;
        bbm_ss ss = P.rsites[k];
        const int8_t* bases = (a.minus ? P.basesM : P.basesP) + P.read_off[rl];
        const int refLen = (int)(P.chrom_off[a.chrom] - P.chrom_off[a.chrom - 1]);
        const int8_t* ref = P.refs + P.chrom_off[a.chrom - 1];
        if (a.msa_req >= 0) {
            const bbm_msa_out m = P.mouts[a.msa_req];
            if (m.status != 0) ps[PS_STATUS0 + o] |= BBM_MAP_ST_ALIGNER;
            if (m.status == 0 && m.score_len > 0) { set_slow_score(ss, m.score[0]); ss.score = ss.slow_score; ss_set_start(ss, m.score[1]); ss_set_stop(ss, m.score[2]); }
            else { set_slow_score(ss, a.sw); ss.score = ss.slow_score; ss_set_start(ss, a.old_start); ss_set_stop(ss, ss.start + lenL - 1); }
        } else { set_slow_score(ss, a.sw); ss.score = ss.slow_score; ss_set_stop(ss, ss.start + lenL - 1); }
        ss.paired_score = ss.score + 1;
        ss.perfect = (ss.slow_score == maxLooseSw) ? 1 : 0;
        if (ss.perfect) ss.semiperfect = 1; else ss_set_perfect(ss, bases, lenL, ref, refLen);
        const bool inb = ss.start >= 0 && ss.stop <= refLen - 1;
        if (ss.score > ps[PS_RETAIN1] && inb) {
            bbm_ss& ssa = A[ia];
            if (ss.score > ps[PS_RETAIN2]) {
                ss.paired_score = imax(ss.paired_score, ss.slow_score + ssa.slow_score / 4);
                ssa.paired_score = imax(ssa.paired_score, ssa.slow_score + ss.slow_score / 4);
            }
            if (nL < P.cap) { L[nL++] = ss; } else ps[PS_STATUS0 + o] |= BBM_MAP_ST_LIST_OVERFLOW;
        }
    }
    nL = merge_duplicate_sites(L, nL);
}

// ---------------- PAIR_FINAL: after rescue up to setFromTopSite (:1113-1195) ----------------
static __device__ void pair_final(PairCtx& X) {
    const PairParams& P = X.P; const bbm_map_cfg& cfg = P.cfg; const bbm_policy_cfg& pc = P.pc;
    const int maxTrim = pc.max_trim_sites_to_retain;
    for (int e = 0; e < 2; e++) if (X.n[e] > 1) stable_sort<false>(X.v[e], X.n[e]);
    for (int e = 0; e < 2; e++) X.n[e] = remove_low_quality_paired(X.v[e], X.n[e], X.maxSw[e], cfg.min_ratio, cfg.min_ratio_paired);
    bbm_ss* a = X.v[0]; bbm_ss* b = X.v[1]; int& na = X.n[0]; int& nb = X.n[1];
    for (int i = 0; i < na; i++) a[i].paired_score = 0;
    for (int i = 0; i < nb; i++) b[i].paired_score = 0;
    if (na >= 1 && nb >= 1) {                                               // pairSiteScoresFinal(r, r2, true, true, ...)
        stable_sort<true>(a, na); stable_sort<true>(b, nb);
        int maxPaired1 = -1, maxPaired2 = -1;
        const float q1 = __fdiv_rn((float)X.len[0], __fmul_rn(4.f, (float)X.len[1])), q2 = __fdiv_rn((float)X.len[1], __fmul_rn(4.f, (float)X.len[0]));
        const float mult1 = fminf(0.5f, fmaxf(0.25f, q1)), mult2 = fminf(0.5f, fmaxf(0.25f, q2));
        const int ilimit = na - 1, jlimit = nb - 1;
        const int outerDistLimit = (imax(X.len[0], X.len[1]) * 14) / 32, MPD = cfg.max_pair_dist, apd = cfg.average_pair_dist;
        const int expectedFragLength = apd + X.len[0] + X.len[1];
        const bool sameStrand = cfg.same_strand_pairs != 0, requireCorrect = cfg.require_correct_strands != 0;
        for (int i = 0, j = 0; i <= ilimit && j <= jlimit; i++) {
            bbm_ss& s1 = a[i];
            while (j < jlimit && (b[j].chrom < s1.chrom || (b[j].chrom == s1.chrom && s1.start - b[j].stop > MPD))) j++;
            for (int k = j; k <= jlimit; k++) {
                bbm_ss& s2 = b[k];
                if (s2.chrom > s1.chrom) break;
                if (s2.start - s1.stop > MPD) break;
                int innerdist, outerdist;
                pair_dists(s1, s2, requireCorrect, innerdist, outerdist);
                if (outerdist >= outerDistLimit && innerdist <= MPD) {
                    const bool strandOK = ((s1.strand == s2.strand) == sameStrand);
                    if (strandOK || !requireCorrect) {
                        const int deviation = apd > innerdist ? apd - innerdist : innerdist - apd;
                        int ps1, ps2;
                        if (strandOK) {
                            const int den = imax(100, 10 * expectedFragLength + 100);
                            ps1 = s1.score + 1 + imax(1, (int)__fmul_rn((float)s2.score, mult1) - ((deviation * s2.score) / den));
                            ps2 = s2.score + 1 + imax(1, (int)__fmul_rn((float)s1.score, mult2) - ((deviation * s1.score) / den));
                        } else { ps1 = s1.score + s2.score / 16; ps2 = s2.score + s1.score / 16; }
                        s1.paired_score = imax(s1.paired_score, ps1);
                        s2.paired_score = imax(s2.paired_score, ps2);
                        maxPaired1 = imax(s1.score, maxPaired1);
                        maxPaired2 = imax(s2.score, maxPaired2);
                    }
                }
            }
        }
        for (int i = 0; i < na; i++) if (a[i].paired_score > a[i].score) a[i].score = a[i].paired_score;
        for (int i = 0; i < nb; i++) if (b[i].paired_score > b[i].score) b[i].score = b[i].paired_score;
        const float f = fminf(cfg.secondary_site_score_ratio, 0.95f);
        na = trim_below_cutoff(a, na, (int)__fmul_rn((float)maxPaired1, f), false, 1, maxTrim);
        nb = trim_below_cutoff(b, nb, (int)__fmul_rn((float)maxPaired2, f), false, 1, maxTrim);
    }
    for (int e = 0; e < 2; e++) if (X.n[e] > 0) stable_sort<false>(X.v[e], X.n[e]);
    int fl[2] = {0, 0};
    for (int e = 0; e < 2; e++) {
        const bbm_ss* v = X.v[e]; const int n = X.n[e], maxSw = X.maxSw[e];
        const bool perfect = n > 0 && (v[0].slow_score == maxSw || v[0].perfect);
        bool ambiguous = false;
        if (n > 1) {
            int cz;
            if (perfect) cz = pc.clearzonep;
            else if (v[0].score >= (int)__fsub_rn(__fmul_rn((float)maxSw, pc.cz1b_scale), pc.cz1b_flat)) cz = pc.clearzone1;
            else if (v[0].score >= (int)__fsub_rn(__fmul_rn((float)maxSw, pc.cz1c_scale), pc.cz1c_flat)) cz = pc.clearzone1b;
            else cz = pc.clearzone1c;
            if (count_top_scores(v, n, cz) > 1) ambiguous = true;
        }
        fl[e] = (perfect ? 2 : 0) | (ambiguous ? 4 : 0);
    }
    bool paired = false;
    if (na > 0 && nb > 0) {                                                 // canPair(top1, top2, ...)
        const bbm_ss& s1 = a[0]; const bbm_ss& s2 = b[0];
        bool ok = s1.chrom == s2.chrom;
        if (ok && cfg.require_correct_strands && ((s1.strand == s2.strand) != (cfg.same_strand_pairs != 0))) ok = false;
        if (ok) {
            int inner, outer;
            pair_dists(s1, s2, cfg.require_correct_strands != 0, inner, outer);
            ok = outer >= (imax(X.len[0], X.len[1]) * 14) / 32 && inner <= cfg.max_pair_dist;
        }
        paired = ok;
    }
    for (int e = 0; e < 2; e++) {
        fl[e] |= (paired ? 8 : 0) | (X.n[e] > 0 ? 1 : 0);
        if (X.n[e] > 0) fl[e] = (fl[e] & ~2) | (X.v[e][0].perfect ? 2 : 0);     // setFromSite: setPerfect(ss.perfect)
    }
    if (cfg.kill_bad_pairs && !paired && X.n[0] > 0 && X.n[1] > 0) {         // Read.isBadPair on the top sites
        const bbm_ss& r = a[0]; const bbm_ss& m = b[0];
        bool bad = false;
        if (r.chrom != m.chrom) bad = true;
        else {
            const int inner = (r.start <= m.start) ? (m.start - r.stop) : (r.start - m.stop);
            if (inner > cfg.max_pair_dist) bad = true;
            else if (cfg.require_correct_strands && ((r.strand == m.strand) != (cfg.same_strand_pairs != 0))) bad = true;
            else if (!cfg.same_strand_pairs) {
                if (r.strand == 0 && m.strand == 1) { if (r.start >= m.stop) bad = true; }
                else if (r.strand == 1 && m.strand == 0) { if (m.start >= r.stop) bad = true; }
            }
        }
        if (bad) {
            const int x = a[0].slow_score / X.len[0], y = b[0].slow_score / X.len[1];
            const int k = (x >= y) ? 1 : 0;                                   // clearAnswers(false) on the weaker mate
            X.n[k] = 0; fl[k] = 0;
        }
    }
    P.rflags[2 * X.p] = fl[0]; P.rflags[2 * X.p + 1] = fl[1];
}

__global__ void __launch_bounds__(128) pair_kernel(PairParams P, int op) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.npairs) return;
    PairCtx X = pair_ctx(P, p);
    if (op == PAIR_OP_INIT) {
        int* ps = P.pstate + p * PAIR_STATE;
        for (int i = 0; i < PAIR_STATE; i++) ps[i] = 0;
        if (P.nkeys[2 * p] < 0 && P.nkeys[2 * p + 1] < 0) { X.n[0] = 0; X.n[1] = 0; ps[PS_DISCARDED] = 1; }    // both quickMaps < 0 (:964-975)
        else pair_initial(X);
    } else if (op == PAIR_OP_RESCUE_PREP0) rescue_prep(X, 0);
    else if (op == PAIR_OP_RESCUE_PREP1) rescue_prep(X, 1);
    else if (op == PAIR_OP_RESCUE_APPLY0) rescue_apply(X, 0);
    else if (op == PAIR_OP_RESCUE_APPLY1) rescue_apply(X, 1);
    else if (op == PAIR_OP_FINAL) pair_final(X);
    P.nss[2 * p] = X.n[0]; P.nss[2 * p + 1] = X.n[1];
}

// ---------------- PAIR_FINISH: after genMatchString (:1228-1352) ----------------
__global__ void __launch_bounds__(128) pair_finish_kernel(PairParams P, FinParams F) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.npairs) return;
    PairCtx X = pair_ctx(P, p);
    const bbm_map_cfg& cfg = P.cfg;
    const int* ps = P.pstate + p * PAIR_STATE;
    bbm_map_rec q[2]; int slot[2]; int st[2];
    bool paired = (P.rflags[2 * p] & 8) != 0;
    for (int e = 0; e < 2; e++) {
        const long long r = 2 * p + e;
        const int* gs = F.state + r * GM_STATE;
        st[e] = gs[S_STATUS] | ps[PS_STATUS0 + e];
        if (gs[S_TOPCHANGED]) paired = false;                             // the top site changed identity while paired (:936-939)
    }
    for (int e = 0; e < 2; e++) {
        const long long r = 2 * p + e;
        bbm_ss* v = X.v[e]; int& n = X.n[e];
        bbm_map_rec& Q = q[e];
        Q = bbm_map_rec{}; Q.chrom = -1; Q.start = -1; Q.stop = -1; Q.match_slot = -1;
        int fl = P.rflags[r] & 7;
        slot[e] = -1;
        if (n > 0) {                                            // r.* = top site after genMatchString (:944-953)
            Q.chrom = v[0].chrom; Q.strand = v[0].strand; Q.start = v[0].start; Q.stop = v[0].stop; Q.map_score = v[0].slow_score;
            fl = (fl & ~2) | (v[0].perfect ? 2 : 0) | (v[0].rescued ? 16 : 0) | 1;
            slot[e] = v[0].has_match ? v[0].has_match - 1 : -1;
            Q.match_len = slot[e] >= 0 ? F.mlen[r * GM_SLOTS + slot[e]] : 0;
        } else fl &= ~1;
        Q.flags = fl;
    }
    for (int e = 0; e < 2; e++) {                               // anomaly blocks: mapScore <= 0 with a list -> clearMapping
        if (X.n[e] > 0 && q[e].map_score <= 0) { X.n[e] = 0; q[e].flags &= ~1; paired = false; }
    }
    for (int e = 0; e < 2; e++) if (X.n[e] > 1) {               // removeDuplicateBestSites
        bbm_ss* v = X.v[e]; int& n = X.n[e]; const bbm_ss t = v[0];
        while (n > 1 && t.chrom == v[n - 1].chrom && t.strand == v[n - 1].strand && t.start == v[n - 1].start && t.stop == v[n - 1].stop) n--;
    }
    for (int e = 0; e < 2; e++) if ((q[e].flags & 4) && cfg.ambiguous_toss) { X.n[e] = 0; q[e].flags &= ~1; paired = false; }
    for (int e = 0; e < 2; e++) {                               // toLocalAlignment for X/Y/C tips
        const long long r = 2 * p + e;
        if (!(q[e].flags & 1) || X.n[e] == 0 || slot[e] < 0 || q[e].match_len < 1) continue;
        int8_t* m = F.mslots + (r * GM_SLOTS + slot[e]) * F.ms; int& ml = F.mlen[r * GM_SLOTS + slot[e]];
        const int8_t a = m[0], b = m[ml - 1];
        if (!(a == 'X' || b == 'Y' || a == 'C' || b == 'C')) continue;
        bbm_ss top = X.v[e][0];
        const int8_t* bases = (top.strand == 0 ? P.basesP : P.basesM) + P.read_off[r];
        const int refLen = (int)(P.chrom_off[top.chrom] - P.chrom_off[top.chrom - 1]);
        int f2 = q[e].flags & 7, rs = q[e].start, rp = q[e].stop, msc = q[e].map_score;
        const bool ok = gm_to_local(top, m, ml, F.ms, bases, X.len[e], P.refs + P.chrom_off[top.chrom - 1], refLen, 1, rs, rp, msc, f2, st[e]);
        if (!ok) { X.n[e] = 0; q[e].flags &= ~1; paired = false; }
        else { X.v[e][0] = top; q[e].start = rs; q[e].stop = rp; q[e].map_score = msc; q[e].flags = (q[e].flags & ~7) | (f2 & 7); q[e].match_len = ml; }
    }
    for (int e = 0; e < 2; e++) {
        const long long r = 2 * p + e;
        bbm_map_rec& Q = q[e];
        if (!(Q.flags & 1) || X.n[e] == 0) { Q.chrom = -1; Q.strand = 0; Q.start = -1; Q.stop = -1; Q.match_len = 0; Q.map_score = 0; Q.match_slot = -1; Q.flags &= ~1; X.n[e] = 0; }
        else Q.match_slot = slot[e];
        Q.flags = (Q.flags & ~8) | (paired ? 8 : 0);
        if (P.nkeys[r] < 0) Q.flags |= 32;
        Q.status = st[e];
        bbm_ss* v = X.v[e];
        for (int i = 0; i < X.n[e]; i++) { v[i].hits &= 0xffff; v[i].has_match = v[i].has_match ? 1 : 0; }
        P.nss[r] = X.n[e];
        F.recs[r] = Q;
    }
    if (paired && (q[0].flags & 1)) {                           // calcStatistics1: numMated, innerLengthSum
        int inner = (q[0].start <= q[1].start) ? (q[1].start - q[0].stop) : (q[0].start - q[1].stop);
        inner = imax(-160, imin(cfg.max_pair_dist, inner));
        atomicAdd((unsigned long long*)P.stats, 1ull);
        atomicAdd((unsigned long long*)P.stats + 1, (unsigned long long)(long long)inner);
    }
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_pair_state_ints() { return PAIR_STATE; }
extern "C" int bbm_launch_pair(const PairParams* P, int op, cudaStream_t st) {
    pair_kernel<<<(unsigned)((P->npairs + 127) / 128), 128, 0, st>>>(*P, op);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_rescue_mid(const PairParams* P, int ntasks, cudaStream_t st) {
    rescue_mid_kernel<<<(unsigned)((ntasks + 127) / 128), 128, 0, st>>>(*P, ntasks);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_pair_finish(const PairParams* P, const FinParams* F, cudaStream_t st) {
    pair_finish_kernel<<<(unsigned)((P->npairs + 127) / 128), 128, 0, st>>>(*P, *F);
    return (int)cudaGetLastError();
}
