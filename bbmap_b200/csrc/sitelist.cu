// sitelist.cu — the per-read site-list policies of the unpaired mapping loop, on the device (part of SURVEY.md §8 f1).
//   BBMapThread.processRead list handling     current/align2/BBMapThread.java:420-431, 440-443, 478-553
//   BBMapThread.trimList (affine branch)      :140-249
//   Tools.trimSiteList / trimSitesBelowCutoff / mergeDuplicateSites / countTopScores / removeLowQualitySitesUnpaired
//                                             current/align2/Tools.java:654-673, 1106-1160, 697-760, 913-931, 986-1003
//   SiteScore.compareTo / PCOMP / positionalMatch / setPerfect   current/stream/SiteScore.java:55-76, 379-395, 353-365, 239-291
//   AbstractMapThread.scoreNoIndels(Read,...) current/align2/AbstractMapThread.java:762-855
//
// One thread per read: the lists are a handful of 80-byte records (a read keeps 1-3 sites after trimming), every policy is a short
// sequential pass whose outcome depends on the order of the elements, and the reads of a batch are independent.  Removals are kept as a
// bit mask and applied by one compaction, which is what the reference's "set to null, then condenseStrict" does.  Java's float
// arithmetic (int -> float conversions, one rounding per operation, truncation by the (int) cast) is spelled out with __fmul_rn /
// __fadd_rn / __fdiv_rn so that no multiply-add is contracted.
#include <cuda_runtime.h>
#include "sitelist_dev.cuh"

namespace bbm {


struct SitelistParams {
    int op; bbm_ss* lists; int* nss; long long nreads; int cap; const long long* read_off;
    const int8_t* basesP; const int8_t* basesM; const int8_t* refs; const long long* chrom_off;
    bbm_policy_cfg cfg; bbm_read_out* out;
};

__global__ void __launch_bounds__(128) sitelist_kernel(SitelistParams P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    bbm_ss* v = P.lists + r * P.cap;
    int n = P.nss[r];
    const int len = (int)(P.read_off[r + 1] - P.read_off[r]);
    const bbm_policy_cfg& cfg = P.cfg;
    bbm_read_out o; o.near_perfect = 0; o.flags = 0; o.clearzone = 0; o.best_sites = 0;
    if (P.op == BBM_SL_TRIM) {
        if (cfg.trim_list && n > 1) {
            if (cfg.min_trim_sites_to_retain > 1) stable_sort<false>(v, n);
            o.best_sites = trim_list(v, n, false, max_quality(len), true, cfg.min_trim_sites_to_retain, cfg.max_trim_sites_to_retain);
        }
    } else if (P.op == BBM_SL_MERGE) {
        n = merge_duplicate_sites(v, n);                 // Tools.mergeDuplicateSites(r.sites, true, true) on its own (processReadPair :1043, :1059)
    } else if (P.op == BBM_SL_NOINDEL) {
        const int maxSw = max_quality(len), maxImp = max_imperfect(len);
        int numNear = 0, best = -0x7fffffff - 1; bool forceSlow = false;
        for (int j = 0; j < n; j++) {
            bbm_ss ss = v[j];
            const int oldScore = ss.score, sslen = ss.stop - ss.start + 1;
            const int8_t* bases = (ss.strand == 0 ? P.basesP : P.basesM) + P.read_off[r];
            const int8_t* ref = P.refs + P.chrom_off[ss.chrom - 1];
            const int refLen = (int)(P.chrom_off[ss.chrom] - P.chrom_off[ss.chrom - 1]);
            if (ss.perfect) {
                numNear++;
                set_slow_score(ss, maxSw); ss.score = maxSw; ss.ngaps = 0;
            } else {
                int sni = score_no_indels(bases, len, ref, refLen, ss.start);
                if (sni < oldScore && oldScore >= maxImp && sslen != len) {
                    const int s2 = score_no_indels(bases, len, ref, refLen, ss.stop - len + 1);
                    if (s2 >= maxImp) { sni = s2; ss.start = ss.stop - len + 1; ss_set_perfect(ss, bases, len, ref, refLen); }
                }
                set_slow_score(ss, sni); ss.score = sni;
                if (sni >= maxImp) {
                    numNear++;
                    ss.stop = ss.start + len - 1; ss.ngaps = 0;
                    if (sni >= maxSw) { ss.perfect = 1; ss.semiperfect = 1; }
                    else ss_set_perfect(ss, bases, len, ref, refLen);
                    if (cfg.quick_match_strings && !ss.perfect && (cfg.print_secondary || sni >= best)) ss.has_match = 1;
                } else if (oldScore >= maxImp) forceSlow = true;
                else if (cfg.print_secondary) forceSlow = true;
            }
            best = imax(ss.slow_score, best);
            v[j] = ss;
        }
        o.near_perfect = n == 0 ? 0 : (forceSlow ? -numNear : numNear);
        stable_sort<false>(v, n);
    } else {
        const int maxSw = max_quality(len);
        int flags = 0, clearzone = 0, numBest = 0;
        if (n > 0) { n = merge_duplicate_sites(v, n); stable_sort<false>(v, n); }
        const bool perfect = n > 0 && (v[0].slow_score == maxSw || v[0].perfect);
        if (n > 1) {
            const int score = v[0].score;
            if (perfect) clearzone = cfg.clearzonep;
            else {
                const float fmax = (float)maxSw, fs = (float)score;
                const float cz1blimit = __fsub_rn(__fmul_rn(fmax, cfg.cz1b_scale), cfg.cz1b_flat);
                const float cz1climit = __fsub_rn(__fmul_rn(fmax, cfg.cz1c_scale), cfg.cz1c_flat);
                if (fs > cz1blimit) {
                    const float t1 = (float)((maxSw - score) * cfg.clearzone1b);                       // int product first, as in Java
                    const float t2 = __fmul_rn(__fsub_rn(fs, cz1blimit), (float)cfg.clearzone1);
                    clearzone = (int)__fdiv_rn(__fadd_rn(t1, t2), __fsub_rn(fmax, cz1blimit));
                } else if (fs > cz1climit) {
                    const float t1 = __fmul_rn(__fsub_rn(cz1blimit, fs), (float)cfg.clearzone1c);
                    const float t2 = __fmul_rn(__fsub_rn(fs, cz1climit), (float)cfg.clearzone1b);
                    clearzone = (int)__fdiv_rn(__fadd_rn(t1, t2), __fsub_rn(cz1blimit, cz1climit));
                } else clearzone = cfg.clearzone1c;
            }
            numBest = count_top_scores(v, n, clearzone);
            if (numBest > 1) flags |= 4;
            else {
                const int lim = (perfect ? (int)__fmul_rn(4.f, (float)cfg.clearzone_limit1e)
                                         : (score + cfg.clearzone1e >= maxSw ? 2 * cfg.clearzone_limit1e : cfg.clearzone_limit1e)) + 1;
                if (n > lim && clearzone < cfg.clearzone1e) {
                    numBest = count_top_scores(v, n, cfg.clearzone1e);
                    if (numBest > lim) flags |= 4;
                }
            }
        }
        if (n > 0) {
            const int lim = (int)__fmul_rn((float)maxSw, cfg.min_align_ratio);
            if (v[0].score < lim) n = 0;
            else {
                const int thresh = imin(lim, imax(1, lim - cfg.clearzone3));      // Tools.removeLowQualitySitesUnpaired(list, thresh)
                unsigned long long dead = 0;
                for (int i = n - 1; i > 1; i--) if (v[i].slow_score < thresh) dead |= 1ull << i;
                n = compact(v, n, dead);
            }
        }
        if (n > 0) flags |= 1;
        if (perfect && n > 0) flags |= 2;
        o.flags = flags; o.clearzone = clearzone; o.best_sites = numBest;
    }
    P.nss[r] = n;
    P.out[r] = o;
}

__global__ void __launch_bounds__(128) sitelist_from_search_kernel(const bbm_search_head* __restrict__ heads, const bbm_site* __restrict__ sites, long long nreads,
                                                                   int maxSites, bbm_ss* __restrict__ lists, int* __restrict__ nss, int cap) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    const int n = imin(imin(heads[r].nsites, maxSites), cap);
    for (int i = 0; i < n; i++) {
        const bbm_site s = sites[r * maxSites + i];
        bbm_ss q = {};
        q.chrom = s.chrom; q.start = s.start; q.stop = s.stop; q.hits = s.hits; q.score = s.score; q.quick_score = s.score;
        q.strand = s.strand; q.perfect = s.perfect; q.semiperfect = s.semiperfect; q.ngaps = s.ngaps;
        for (int g = 0; g < BBM_MAX_GAPS - 1; g++) q.gaps[g] = s.gaps[g];
        lists[r * cap + i] = q;
    }
    nss[r] = n;
}



// =====================  removeOutOfBounds (AbstractMapThread.java:2444-2479), quickMap's step right after the index search (:739)  =====================
__device__ int sl_bsearch_java(const int* a, int n, int key) {       // Arrays.binarySearch
    int lo = 0, hi = n - 1;
    while (lo <= hi) { const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1); const int v = a[mid]; if (v < key) lo = mid + 1; else if (v > key) hi = mid - 1; else return mid; }
    return -(lo + 1);
}
__device__ bool sl_is_single_scaffold(const int* loc, int n, int pad, int loc1, int loc2) {      // Data.isSingleScaffold (dna/Data.java:1112-1140)
    if (n < 2) return true;
    const int idx = sl_bsearch_java(loc, n, loc1 + pad);
    const int scaf = idx >= 0 ? idx : imax(0, (-1 - idx) - 1);
    if (scaf == n - 1) return true;
    const int lowerBound = loc[scaf] - pad, upperBound = loc[scaf + 1];
    if (loc2 < lowerBound || loc1 > upperBound) return false;
    return loc2 < upperBound;
}
struct BoundsParams {
    bbm_ss* lists; int* nss; long long nreads; int cap; const long long* read_off; const int* chrom_max_index;
    const int* scaf_off; const int* scaf_loc; int pad, sam_out, expected_len_limit; bbm_read_out* out;
};
__global__ void __launch_bounds__(128) sitelist_bounds_kernel(BoundsParams P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    bbm_ss* v = P.lists + r * P.cap;
    const int n = P.nss[r], len = (int)(P.read_off[r + 1] - P.read_off[r]);
    unsigned long long dead = 0; int flags = 0;
    for (int i = 0; i < n; i++) {
        bbm_ss& ss = v[i];
        bool removed = false;
        if (ss.start < 0 || ss.stop > P.chrom_max_index[ss.chrom - 1]) removed = true;
        else if (P.sam_out && P.scaf_off) {
            const int base = P.scaf_off[ss.chrom - 1], cnt = P.scaf_off[ss.chrom] - base;
            if (!sl_is_single_scaffold(P.scaf_loc + base, cnt, P.pad, ss.start, ss.stop)) removed = true;
        }
        if (removed) { dead |= 1ull << i; continue; }
        if (calc_gref_len(ss) >= P.expected_len_limit) {
            ss_set_stop(ss, ss.start + imin(len + 40, P.expected_len_limit));
            if (ss.ngaps > 0) ss.ngaps = fix_gaps(ss.start, ss.stop, ss.gaps, ss.ngaps, SL_MINGAP);      // :2470 calls fixGaps once more
        }
    }
    const int n2 = compact(v, n, dead);
    P.nss[r] = n2;
    bbm_read_out o; o.near_perfect = 0; o.flags = flags; o.clearzone = 0; o.best_sites = n - n2;
    P.out[r] = o;
}

// =====================  findTipDeletions(Read, ...) (AbstractMapThread.java:1073-1104)  =====================
// Scalar forms of findTipDeletionsRight/Left (:2178-2294; the warp-per-task forms live in rescue.cu): here the unit is the read, the
// scans are short (<= 100 starts x 8 bases) and most sites stop after the 8-base tip check.
struct TipParams {
    bbm_ss* lists; const int* nss; long long nreads; int cap; const long long* read_off; const int8_t* basesP; const int8_t* basesM;
    const int8_t* quality; const int8_t* refs; const long long* chrom_off; const int* chrom_min_index; bbm_tipdel_cfg tc; bbm_read_out* out;
};

__global__ void __launch_bounds__(128) sitelist_tipdel_kernel(TipParams P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    bbm_ss* v = P.lists + r * P.cap;
    const int n = P.nss[r], len = (int)(P.read_off[r + 1] - P.read_off[r]), TIPLEN = P.tc.max_tiplen;
    bbm_read_out o; o.near_perfect = 0; o.flags = 0; o.clearzone = 0; o.best_sites = 0;
    bool findRight = true, findLeft = true;
    if (len == 0) { P.out[r] = o; return; }
    if (P.quality) {                                   // Read.min/avgQuality{First,Last}NBases (stream/Read.java:1760-1815)
        const int8_t* q = P.quality + P.read_off[r];
        int minL = 0, avgL = 0, minF = 0, avgF = 0;
        if (TIPLEN <= len) {
            int x = 0; minL = q[len - TIPLEN];
            for (int i = len - TIPLEN; i < len; i++) { const int b = q[i]; x += (b < 0 ? 0 : b); minL = imin(minL, b); }
            avgL = x / TIPLEN;
            x = 0; minF = q[0];
            for (int i = 0; i < TIPLEN; i++) { const int b = q[i]; x += (b < 0 ? 0 : b); minF = imin(minF, b); }
            avgF = x / TIPLEN;
        }
        findRight = (minL >= 6 && avgL >= 14); findLeft = (minF >= 6 && avgF >= 14);
    }
    if (findRight || findLeft) {
        const int maxSw = max_quality(len), maxImp = max_imperfect(len);
        for (int j = 0; j < n; j++) {
            bbm_ss ss = v[j];
            if (ss.semiperfect || ss.slow_score >= maxImp) continue;
            const int8_t* bases = (ss.strand == 0 ? P.basesP : P.basesM) + P.read_off[r];
            const int8_t* ref = P.refs + P.chrom_off[ss.chrom - 1];
            const int refLen = (int)(P.chrom_off[ss.chrom] - P.chrom_off[ss.chrom - 1]);
            const int minIndex = P.chrom_min_index ? P.chrom_min_index[ss.chrom - 1] : 0;
            bool changed = false;
            if (len > 2 * TIPLEN) {
                int maxSearch = imin(P.tc.search_range, P.tc.align_columns - (P.tc.slow_rescue_padding + 8 + imax(len, ss.stop - ss.start)));
                if (maxSearch >= 1) {
                    bool go = true;
                    if (findRight) {
                        const int x = tip_right(bases, len, ref, refLen, minIndex, ss.stop, maxSearch, TIPLEN);
                        if (x > 0) {
                            ss_set_stop(ss, ss.stop + x); changed = true;
                            maxSearch = imin(maxSearch, P.tc.align_columns - (P.tc.slow_rescue_padding + 8 + imax(len, ss.stop - ss.start)));
                            if (maxSearch < 1) go = false;
                        }
                    }
                    if (go && findLeft) {
                        const int y = tip_left(bases, ref, refLen, minIndex, ss.start, maxSearch, TIPLEN);
                        if (y > 0) { ss_set_start(ss, ss.start - y); changed = true; }
                    }
                }
            }
            if (changed) {
                o.best_sites++;
                ss.has_match = 0;
                set_slow_score(ss, score_no_indels(bases, len, ref, refLen, ss.start));
                if (ss.slow_score == maxSw) { ss_set_stop(ss, ss.start + len - 1); ss.perfect = 1; ss.semiperfect = 1; }
                else { ss.perfect = 0; ss_set_perfect(ss, bases, len, ref, refLen); }
                v[j] = ss;
            }
        }
    }
    P.out[r] = o;
}

// =====================  scoreSlow in rounds (BBMapThread.scoreSlow, current/align2/BBMapThread.java:252-386)  =====================
// scoreSlow walks the sites of one read in order because its limit ratchets: minMsaLimit = max(minMsaLimit, slowScore - CLEARZONE3)
// after every site (:375-376).  Reads are independent, so the batched form is rounds: round k handles the k-th site of every read —
//   SLOW_PREP   the per-site preamble (:270-300) and the fillAndScoreLimited request (MSA.java:136-143: window = site +- pad, clamped)
//   [one bbm_msa batch over all reads of the round]
//   SLOW_RETRY  "more padding needed" (:303-326): score array of length 8 -> widen the site by the suggested pads, ask again with
//               SLOW_ALIGN_PADDING+EXTRA_PADDING;  [second batch, only if some read asked]
//   SLOW_APPLY  keep the better array, setSlowScore/setLimits, ratchet, perfect/semiperfect flags (:327-385)
// Default flags only: QUICK_MATCH_STRINGS=false (no traceback / fixXY / clipTipIndels inside scoreSlow).  Sites that carry a gap array go to
// the gapped aligner (bbm_msa_gapped: makeGref on the device) as a second packed request list; SiteScore.setLimits / setStop keep the gap
// array consistent through GapTools.fixGaps (GapTools.java:26-72,126-175; stream/SiteScore.java:905-914,943-958).
//
// Look-ahead.  The ONLY state scoreSlow carries from one site of a read to the next is minMsaLimit (minMatch matters to QUICK_MATCH_STRINGS alone), and a
// site sees it only through minscore = max(swscoreNoIndel, minMsaLimit).  So from a read's second site on, a round takes a WINDOW of its next sites at
// once, each prepared with the limit as it stands, and SLOW_APPLY walks the window in list order: a site whose minscore would be different under the
// limit as ratcheted by the sites before it is put back exactly as it was (a copy is kept) together with everything behind it, and waits for the next
// round.  Every result that is applied was therefore computed with the very limit the reference would have used; sites are sorted best-first, so the
// limit almost never moves after the first site and a read from a 64-copy repeat takes 5 rounds instead of 64 (human-scale scoreSlow 185 -> see DESIGN).
struct SlowParams {
    int phase, round, window; bbm_ss* lists; const int* nss; long long nreads; int cap; const long long* read_off;
    const int8_t* basesP; const int8_t* basesM; const int8_t* refs; const long long* chrom_off; const int* run;
    bbm_slow_cfg cfg; int* state;       // [nreads][20]: 0 minMsaLimit, 1 minMatch, 14 status, 17 cursor (next site), 18 first slot of this round's window, 19 window length
    int* slots;                         // [pool][20] per site in flight: 2 aligned, 3 expectedLen, 4 minscore, 5 old_len, 6..13 old[8], 15 request slot, 16 gapped request, 17 swscoreNoIndel
    bbm_ss* backup;                     // [pool] the site as it was before SLOW_PREP touched it
    bbm_msa_task* tasks; const bbm_msa_out* outs;                  // plain requests, packed
    bbm_gapped_task* gtasks; int* gaps; const bbm_msa_out* gouts;  // gapped requests, packed; gap arrays at gaps[slot * BBM_MAX_GAPS]
    int* counters;                                                 // [0] reads active in this round, [1] plain requests, [2] gapped requests, [3] slot pool cursor, [4] alignments applied (whole call)
};
constexpr int SLOW_PREP = 0, SLOW_RETRY = 1, SLOW_APPLY = 2;
constexpr int SLOW_STATE = 20;
// one fillAndScoreLimited(bases, ss, pad, minscore) request: plain or gapped list, slot remembered in the read's state
__device__ void slow_request(const SlowParams& P, long long r, int len, const bbm_ss& ss, int pad, int minscore, int* st) {
    bbm_msa_task task = {};
    task.read_off = ((ss.strand == 0 ? P.basesP : P.basesM) - P.basesP) + P.read_off[r];
    task.ref_off = P.chrom_off[ss.chrom - 1]; task.read_len = len;
    task.ref_len = (int)(P.chrom_off[ss.chrom] - P.chrom_off[ss.chrom - 1]);
    task.ref_start = ss.start - pad; task.ref_end = ss.stop + pad;
    task.min_score = minscore;
    if (ss.ngaps > 0) {
        const int slot = atomicAdd(P.counters + 2, 1);
        task.flags = BBM_TF_SCORE;                                   // the gapped path clamps the window itself (MSA.java:104-105)
        bbm_gapped_task g; g.t = task; g.gaps_off = slot * BBM_MAX_GAPS; g.ngaps = ss.ngaps;
        for (int i = 0; i < ss.ngaps; i++) P.gaps[slot * BBM_MAX_GAPS + i] = ss.gaps[i];
        P.gtasks[slot] = g; st[15] = slot; st[16] = 1;
    } else {
        const int slot = atomicAdd(P.counters + 1, 1);
        task.flags = BBM_TF_CLAMP | BBM_TF_SCORE;
        P.tasks[slot] = task; st[15] = slot; st[16] = 0;
    }
}

__global__ void __launch_bounds__(128) scoreslow_kernel(SlowParams P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    int* st = P.state + r * SLOW_STATE;
    const int len = (int)(P.read_off[r + 1] - P.read_off[r]);
    const bbm_slow_cfg& cfg = P.cfg;
    if (P.phase == SLOW_PREP) {
        if (P.round == 0) {
            const int maxSw = max_quality(len);
            const int lim = -cfg.clearzone1e + (int)__fmul_rn(cfg.paired ? cfg.min_ratio_pre_rescue : cfg.min_ratio, (float)maxSw);
            st[0] = lim; st[1] = imax(-300, lim - cfg.clearzone3); st[14] = 0; st[17] = 0;
        }
        const int cur = st[17], n = P.nss[r];
        st[19] = 0;
        if (P.run[r] != 0 && cur < n) {
            atomicAdd(P.counters, 1);
            const int w = (cur < 1) ? 1 : imin(P.window, n - cur);       // the first site sets the limit every later one is judged by
            const int base = atomicAdd(P.counters + 3, w);
            st[18] = base; st[19] = w;
            for (int g = 0; g < w; g++) {
                const int k = cur + g;
                int* ps = P.slots + (long long)(base + g) * SLOW_STATE;
                bbm_ss ss = P.lists[r * P.cap + k];
                P.backup[base + g] = ss;
                ps[2] = 0;
                if (ss.stop - ss.start != len - 1) { set_slow_score(ss, 0); ss.semiperfect = 0; ss.perfect = 0; }
                const int sw = ss.slow_score;
                if (sw < max_imperfect(len) && !ss.semiperfect) {
                    const int expectedLen = calc_gref_len(ss);
                    if (expectedLen >= cfg.expected_len_limit) ss_set_stop(ss, ss.start + imin(len + 40, cfg.expected_len_limit));
                    const int minscore = imax(sw, st[0]);
                    ps[2] = 1; ps[3] = expectedLen; ps[4] = minscore; ps[17] = sw;
                    slow_request(P, r, len, ss, cfg.slow_align_padding, minscore, ps);
                }
                P.lists[r * P.cap + k] = ss;
            }
        }
    } else if (P.phase == SLOW_RETRY) {
        const int cur = st[17], w = st[19], base = st[18];
        for (int g = 0; g < w; g++) {
            int* ps = P.slots + (long long)(base + g) * SLOW_STATE;
            if (ps[2] != 1) continue;
            const int k = cur + g;
            const bbm_msa_out o = ps[16] ? P.gouts[ps[15]] : P.outs[ps[15]];
            if (o.status != 0) ps[18] = BBM_SLOW_ALIGNER_ERROR; else ps[18] = 0;
            const int n = (o.status == 0) ? o.score_len : 0;
            ps[5] = n;
#pragma unroll
            for (int q = 0; q < 8; q++) ps[6 + q] = o.score[q];
            if (n > 6 && (o.score[3] + o.score[4] + ps[3] < cfg.expected_len_limit)) {
                bbm_ss ss = P.lists[r * P.cap + k];
                ss_set_limits(ss, ss.start - o.score[6], ss.stop + o.score[7]);
                ps[2] = 2;
                slow_request(P, r, len, ss, cfg.slow_align_padding + cfg.extra_padding, ps[4], ps);
                P.lists[r * P.cap + k] = ss;
            }
        }
    } else {
        const int cur = st[17], w = st[19], base = st[18];
        for (int g = 0; g < w; g++) {
            const int k = cur + g;
            const int* ps = P.slots + (long long)(base + g) * SLOW_STATE;
            if (g > 0 && ps[2] >= 1 && imax(ps[17], st[0]) != ps[4]) {
                // the limit moved under this request: the site and everything behind it go back to what they were and are asked again next round
                for (int q = g; q < w; q++) P.lists[r * P.cap + cur + q] = P.backup[base + q];
                break;
            }
            bbm_ss ss = P.lists[r * P.cap + k];
            int n = 0, a0 = 0, a1 = 0, a2 = 0;
            if (ps[2] >= 1) { n = ps[5]; a0 = ps[6]; a1 = ps[7]; a2 = ps[8]; st[14] |= ps[18]; atomicAdd(P.counters + 4, ps[2]); }
            if (ps[2] == 2) {
                const bbm_msa_out o = ps[16] ? P.gouts[ps[15]] : P.outs[ps[15]];
                if (o.status != 0) st[14] |= BBM_SLOW_ALIGNER_ERROR;
                const int n2 = (o.status == 0) ? o.score_len : 0;
                if (!(n2 == 0 || o.score[0] < a0)) { n = n2; a0 = o.score[0]; a1 = o.score[1]; a2 = o.score[2]; }
            }
            if (n > 0) { set_slow_score(ss, a0); ss_set_limits(ss, a1, a2); }
            ss.score = ss.slow_score;
            st[1] = imax(st[1], ss.slow_score);
            st[0] = imax(st[0], ss.slow_score - cfg.clearzone3);
            const int maxSw = max_quality(len);
            ss.perfect = (ss.slow_score == maxSw) ? 1 : 0;
            if (ss.perfect) ss.semiperfect = 1;
            else if (!ss.semiperfect) {
                const int8_t* bases = (ss.strand == 0 ? P.basesP : P.basesM) + P.read_off[r];
                ss_set_perfect(ss, bases, len, P.refs + P.chrom_off[ss.chrom - 1], (int)(P.chrom_off[ss.chrom] - P.chrom_off[ss.chrom - 1]));
            }
            P.lists[r * P.cap + k] = ss;
            st[17] = k + 1;
        }
        st[19] = 0;
    }
}

// =====================  after the primary site's match string: clearzone 3, final score gate, tip penalty  =====================
// BBMapThread.processRead :667-684, 698-700, 706-709; AbstractMapThread.applyClearzone3 :1820-1870, calcCZ3_fraction :1893-1911,
// calcTipScorePenalty :2499-2567, applyScorePenalty :2601-2609.  Java float arithmetic: one rounding per operation, no contraction.
struct Cz3Params { bbm_ss* lists; int* nss; long long nreads; int cap; const long long* read_off; bbm_policy_cfg cfg; int ambiguous_toss; bbm_read_out* io; };
__global__ void __launch_bounds__(128) sitelist_cz3_kernel(Cz3Params P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    bbm_ss* v = P.lists + r * P.cap;
    int n = P.nss[r];
    const int len = (int)(P.read_off[r + 1] - P.read_off[r]);
    const int maxSw = max_quality(len);
    const bbm_policy_cfg& cfg = P.cfg;
    int flags = P.io[r].flags, subi = 0;
    if (n > 1) {                                                         // removeDuplicateBestSites (:1328-1349): copies of the top site at the end of the list
        const bbm_ss t = v[0];
        while (n > 1 && t.chrom == v[n - 1].chrom && t.strand == v[n - 1].strand && t.start == v[n - 1].start && t.stop == v[n - 1].stop) n--;
    }
    if (n == 0) flags &= ~1;
    int mapScore = n > 0 ? v[0].slow_score : 0;
    if ((cfg.clearzone3 > cfg.clearzone1 || cfg.clearzone3 > cfg.clearzonep) && n > 0 && !(flags & 4) && mapScore > 0) {
        const float q = __fdiv_rn((float)maxSw, (float)mapScore);
        const float cz3v2 = __fmul_rn((float)cfg.clearzone3, 1.25f < q ? 1.25f : q);
        const int cz3 = (int)cz3v2; const float inv = __fdiv_rn(1.f, cz3v2);
        if ((flags & 1) && n >= 2) {                                     // applyClearzone3
            const int score1 = mapScore;
            float sub = 0.f;
            const int mx = imin(7, n);
            int prevScore = v[0].slow_score;
            for (int i = 1; i < mx; i++) {
                const int s2 = v[i].slow_score;
                if (i > 2 && s2 < prevScore) break;
                const float f = calc_cz3_fraction(score1, s2, cz3, inv);
                if (f <= 0.f) break;
                sub = __fadd_rn(sub, __fmul_rn(f, cz3_mult(i)));
                prevScore = s2;
            }
            if (sub > 0.f) {
                const float asym = __fadd_rn(4.f, __fmul_rn(0.03f, (float)len));
                sub = __fmul_rn(sub, 1.8f);
                const float sub2 = __fmul_rn((float)cz3, __fdiv_rn(__fmul_rn(asym, sub), __fadd_rn(sub, asym)));
                subi = (int)__fadd_rn(sub2, 0.5f);
                if (subi >= mapScore - 300) subi = mapScore - 300;
                if (subi <= 0) subi = 0;
                else for (int i = 0; i < n; i++) { bbm_ss ss = v[i]; set_slow_score(ss, ss.slow_score - subi); ss.score -= subi; v[i] = ss; }
            }
        }
        if (subi > 0) {
            mapScore -= subi;
            if (mapScore < (int)__fmul_rn((float)maxSw, cfg.min_align_ratio)) flags |= 4;
        }
    }
    if ((flags & 4) && P.ambiguous_toss) { n = 0; flags &= ~1; mapScore = 0; }
    if (n == 0 || (!(flags & 4) && (float)mapScore < __fmul_rn((float)maxSw, cfg.min_align_ratio))) { n = 0; flags &= ~1; mapScore = 0; }   // r.clearMapping()
    P.nss[r] = n;
    bbm_read_out o = P.io[r];
    o.flags = flags; o.near_perfect = mapScore; o.best_sites = subi;
    P.io[r] = o;
}

struct TipPenParams {
    bbm_ss* lists; const int* nss; long long nreads; int cap; const long long* read_off; const int8_t* bases; const int8_t* match; const long long* match_off;
    const bbm_read_out* flags; int tiplen; int* penalty; int* status;
};
__device__ __forceinline__ int tip_points(int8_t b, int8_t prev, int tiplen, int& cpos) {      // one match symbol of either tip loop (:2507-2525 / :2529-2542)
    if (b == 'm') { cpos++; return 0; }
    if (b == 'D') return prev != 'D' ? 2 * (tiplen + 2 - cpos) : 0;
    const int w = (b == 'N' || b == 'C') ? 1 : 2;
    const int p = w * (tiplen + 2 - cpos);
    cpos++;
    return p;
}
__global__ void __launch_bounds__(128) sitelist_tip_penalty_kernel(TipPenParams P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    bbm_ss* v = P.lists + r * P.cap;
    const int n = P.nss[r], tiplen = P.tiplen;
    const int len = (int)(P.read_off[r + 1] - P.read_off[r]);
    const int mlen = (int)(P.match_off[r + 1] - P.match_off[r]);
    const int8_t* match = P.match + P.match_off[r];
    const int8_t* bases = P.bases + P.read_off[r];
    int st = 0, pen = 0;
    if ((P.flags[r].flags & 1) && n > 0 && mlen > 0 && len >= 2 * tiplen) {
        const int maxScore = max_quality(len), mapScore = v[0].slow_score;
        int points = 0;
        int8_t prev = 'm';
        for (int i = 0, cpos = 0; cpos <= tiplen; i++) {
            if (i >= mlen) { st |= 1; break; }
            const int8_t b = match[i];
            if (b >= '0' && b <= '9') { st |= 2; break; }
            points += tip_points(b, prev, tiplen, cpos);
            prev = b;
        }
        prev = 'm';
        if (!st) for (int i = mlen - 1, cpos = 0; cpos <= tiplen; i--) {
            if (i < 0) { st |= 1; break; }
            const int8_t b = match[i];
            points += tip_points(b, prev, tiplen, cpos);
            prev = b;
        }
        if (!st) {
            const int last = len - 1;
            int8_t b = bases[0];
            if (b != 'N' && b == bases[1]) for (int i = 2; i <= tiplen && bases[i] == b; i++) points++;
            b = bases[last];
            if (b != 'N' && b == bases[last - 1]) for (int i = last - 2; i >= (last - tiplen) && bases[i] == b; i--) points++;
            if (points >= 1) {
                const float f = __fdiv_rn(__fmul_rn(80.f, (float)points), __fadd_rn((float)points, 80.f));
                const int penalty = (int)__fmul_rn(__fmul_rn(f, .0022f), (float)maxScore);
                const int maxPenalty = mapScore - maxScore / 10;
                if (maxPenalty > 0) pen = imin(penalty, maxPenalty);
            }
        }
        if (pen > 0) for (int i = 0; i < n; i++) { bbm_ss ss = v[i]; set_slow_score(ss, ss.slow_score - pen); ss.score -= pen; v[i] = ss; }   // applyScorePenalty
    }
    P.penalty[r] = pen;
    if (P.status) P.status[r] = st;
}

// Read.setFromTopSite / setFromSite (stream/Read.java:1171-1190, 1213-1224) for an unpaired read, as the record SamLine(Read,int) reads: chrom, strand, start, stop,
// mapScore = slowScore, perfect = ss.perfect; an empty list is Read.clearSite (:1278-1286).  Gapped top sites keep start/stop (fixGaps only touches the gap array).
__global__ void __launch_bounds__(128) sam_tasks_from_lists_kernel(const bbm_ss* __restrict__ lists, const int* __restrict__ nss, long long nreads, int cap,
                                                                   const long long* __restrict__ read_off, const bbm_read_out* __restrict__ flags,
                                                                   const long long* __restrict__ match_off, bbm_sam_task* __restrict__ tasks) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    bbm_sam_task t;
    t.match_off = match_off ? match_off[r] : 0;
    t.match_len = match_off ? (int)(match_off[r + 1] - match_off[r]) : 0;
    t.read_len = (int)(read_off[r + 1] - read_off[r]);
    t.mate = -1; t.pad_ = 0;
    const int f = flags[r].flags;
    if (nss[r] > 0 && (f & 1)) {
        const bbm_ss ss = lists[r * cap];
        t.chrom = ss.chrom; t.start = ss.start; t.stop = ss.stop; t.score = ss.slow_score;
        t.flags = BBM_RF_MAPPED | (ss.strand ? BBM_RF_MINUS : 0) | (ss.perfect ? BBM_RF_PERFECT : 0) | ((f & 4) ? BBM_RF_AMBIGUOUS : 0);
    } else {
        t.chrom = -1; t.start = -1; t.stop = -1; t.score = 0; t.match_len = 0;
        t.flags = (f & 4) ? BBM_RF_AMBIGUOUS : 0;
    }
    tasks[r] = t;
}

}  // namespace bbm

extern "C" int bbm_launch_sitelist_cz3(bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const bbm_policy_cfg* cfg,
                                       int ambiguous_toss, bbm_read_out* io, cudaStream_t st) {
    bbm::Cz3Params P;
    P.lists = lists; P.nss = nss; P.nreads = nreads; P.cap = cap; P.read_off = read_off; P.cfg = *cfg; P.ambiguous_toss = ambiguous_toss; P.io = io;
    bbm::sitelist_cz3_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_sitelist_tip_penalty(bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off, const int8_t* bases,
                                               const int8_t* match, const long long* match_off, const bbm_read_out* flags, int tiplen, int* penalty,
                                               int* status, cudaStream_t st) {
    bbm::TipPenParams P;
    P.lists = lists; P.nss = nss; P.nreads = nreads; P.cap = cap; P.read_off = read_off; P.bases = bases; P.match = match; P.match_off = match_off;
    P.flags = flags; P.tiplen = tiplen; P.penalty = penalty; P.status = status;
    bbm::sitelist_tip_penalty_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}

extern "C" int bbm_sitelist_max_cap() { return bbm::SL_MAX_CAP; }
extern "C" int bbm_launch_sitelist(int op, bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const int8_t* basesP,
                                   const int8_t* basesM, const int8_t* refs, const long long* chrom_off, const bbm_policy_cfg* cfg, bbm_read_out* out,
                                   cudaStream_t st) {
    bbm::SitelistParams P;
    P.op = op; P.lists = lists; P.nss = nss; P.nreads = nreads; P.cap = cap; P.read_off = read_off; P.basesP = basesP; P.basesM = basesM;
    P.refs = refs; P.chrom_off = chrom_off; P.cfg = *cfg; P.out = out;
    bbm::sitelist_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_sitelist_from_search(const bbm_search_head* heads, const bbm_site* sites, long long nreads, int maxSites, bbm_ss* lists,
                                               int* nss, int cap, cudaStream_t st) {
    bbm::sitelist_from_search_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(heads, sites, nreads, maxSites, lists, nss, cap);
    return (int)cudaGetLastError();
}

extern "C" int bbm_launch_scoreslow(int phase, int round, bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off,
                                    const int8_t* basesP, const int8_t* basesM, const int8_t* refs, const long long* chrom_off, const int* run,
                                    const bbm_slow_cfg* cfg, int* state, bbm_msa_task* tasks, const bbm_msa_out* outs, bbm_gapped_task* gtasks, int* gaps,
                                    const bbm_msa_out* gouts, int* counters, int window, int* slots, bbm_ss* backup, cudaStream_t st) {
    bbm::SlowParams P;
    P.window = window; P.slots = slots; P.backup = backup;
    P.phase = phase; P.round = round; P.lists = lists; P.nss = nss; P.nreads = nreads; P.cap = cap; P.read_off = read_off; P.basesP = basesP;
    P.basesM = basesM; P.refs = refs; P.chrom_off = chrom_off; P.run = run; P.cfg = *cfg; P.state = state; P.tasks = tasks; P.outs = outs; P.gtasks = gtasks; P.gaps = gaps; P.gouts = gouts;
    P.counters = counters;
    bbm::scoreslow_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_scoreslow_state_ints() { return bbm::SLOW_STATE; }

extern "C" int bbm_launch_sitelist_tipdel(bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off, const int8_t* basesP,
                                          const int8_t* basesM, const int8_t* quality, const int8_t* refs, const long long* chrom_off,
                                          const int* chrom_min_index, const bbm_tipdel_cfg* tc, bbm_read_out* out, cudaStream_t st) {
    bbm::TipParams P;
    P.lists = lists; P.nss = nss; P.nreads = nreads; P.cap = cap; P.read_off = read_off; P.basesP = basesP; P.basesM = basesM; P.quality = quality;
    P.refs = refs; P.chrom_off = chrom_off; P.chrom_min_index = chrom_min_index; P.tc = *tc; P.out = out;
    bbm::sitelist_tipdel_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}

extern "C" int bbm_launch_sitelist_bounds(bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const int* chrom_max_index,
                                          const int* scaf_off, const int* scaf_loc, int pad, int sam_out, int expected_len_limit, bbm_read_out* out, cudaStream_t st) {
    bbm::BoundsParams P;
    P.lists = lists; P.nss = nss; P.nreads = nreads; P.cap = cap; P.read_off = read_off; P.chrom_max_index = chrom_max_index; P.scaf_off = scaf_off;
    P.scaf_loc = scaf_loc; P.pad = pad; P.sam_out = sam_out; P.expected_len_limit = expected_len_limit; P.out = out;
    bbm::sitelist_bounds_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_sam_tasks_from_lists(const bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off, const bbm_read_out* flags,
                                               const long long* match_off, bbm_sam_task* tasks, cudaStream_t st) {
    bbm::sam_tasks_from_lists_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(lists, nss, nreads, cap, read_off, flags, match_off, tasks);
    return (int)cudaGetLastError();
}
