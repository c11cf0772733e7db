// genmatch.cu — the primary site's match string and the tail of BBMapThread.processRead, on the device.
//
//   genMatchString / genMatchStringForSite        current/align2/AbstractMapThread.java:860-1068
//   TranslateColorspaceRead.realign_new           current/align2/TranslateColorspaceRead.java:229-660
//   SiteScore.fixXY / clipTipIndels / unclip / leftPaddingNeeded / rightPaddingNeeded / fixLimitsXY / setPerfectFlag / isPerfect / isSemiPerfect
//                                                 current/stream/SiteScore.java:175-236, 431-840, 916-931
//   MSA.score(match) / toLocalAlignment           current/align2/MSA.java:216-470, 488-560;  calcDelScore / calcInsScore  MultiStateAligner11tsJNI.java:1347-1421
//   processRead after the list is final           current/align2/BBMapThread.java:557-709
//
// The reference walks one read at a time and calls the aligner synchronously, up to four fills per realign_new level, two levels of
// recursion and two realign_new calls per site.  Here every read is a small coroutine: one thread runs the read's control flow until it
// needs an alignment, appends the request (fillLimited + score + traceback in one task) to a packed list and parks with a resume label;
// the host runs the batched aligner on the list (the same kernels scoreSlow uses) and relaunches; reads finish independently.  A site's
// match string lives in one of GM_SLOTS byte slots owned by the read (bbm_ss.has_match = slot + 1), so sorts and merges move it with
// the record.
#include <cuda_runtime.h>
#include "sitelist_dev.cuh"
#include "mapper_kernels.cuh"
#include "genmatch_dev.cuh"

namespace bbm {




// ---------------- the coroutine ----------------
struct GmRead {
    const GmParams& P; long long r; int* st; bbm_ss* v; int n; int len, maxSw;
    __device__ int8_t* slot(int s) const { return P.mslots + (r * GM_SLOTS + s) * P.ms; }
    __device__ int& slen(int s) const { return P.mlen[r * GM_SLOTS + s]; }
    __device__ const int8_t* bases(const bbm_ss& ss) const { return (ss.strand == 0 ? P.basesP : P.basesM) + P.read_off[r]; }
    __device__ const int8_t* ref(const bbm_ss& ss, int& refLen) const { refLen = (int)(P.chrom_off[ss.chrom] - P.chrom_off[ss.chrom - 1]); return P.refs + P.chrom_off[ss.chrom - 1]; }
    __device__ int alloc_slot() { for (int s = 0; s < GM_SLOTS; s++) if (!(st[S_SLOTMASK] & (1 << s))) { st[S_SLOTMASK] |= (1 << s); slen(s) = 0; return s; } return -1; }
    __device__ void free_slot(int s) { st[S_SLOTMASK] &= ~(1 << s); }
};

__device__ void gm_request(GmRead& G, const bbm_ss& ss, int minLoc, int maxLoc, int minScore, bool unlimited) {
    const GmParams& P = G.P;
    bbm_msa_task task = {};
    task.read_off = ((ss.strand == 0 ? P.basesP : P.basesM) - P.basesP) + P.read_off[G.r];
    task.ref_off = P.chrom_off[ss.chrom - 1]; task.read_len = G.len;
    task.ref_len = (int)(P.chrom_off[ss.chrom] - P.chrom_off[ss.chrom - 1]);
    task.ref_start = minLoc; task.ref_end = maxLoc; task.min_score = minScore;
    task.flags = BBM_TF_SCORE | BBM_TF_TRACEBACK | (unlimited ? BBM_TF_RAW_UNLIMITED : 0);
    const int need = G.len + (maxLoc - minLoc + 1) + 264;
    G.st[S_FILLS]++;
    if (ss.ngaps > 0) {
        const int k = atomicAdd(P.counters + 2, 1);
        bbm_gapped_task g; g.t = task; g.gaps_off = k * BBM_MAX_GAPS; g.ngaps = ss.ngaps;
        for (int i = 0; i < ss.ngaps; i++) P.gaps[k * BBM_MAX_GAPS + i] = ss.gaps[i];
        P.gtasks[k] = g; G.st[S_REQ] = k; G.st[S_REQ_GAPPED] = 1;
        atomicMax(P.counters + 4, need);
    } else {
        const int k = atomicAdd(P.counters + 1, 1);
        P.tasks[k] = task; G.st[S_REQ] = k; G.st[S_REQ_GAPPED] = 0;
        atomicMax(P.counters + 3, need);
    }
    atomicAdd(P.counters, 1);
}

__device__ void gm_run(GmRead& G) {
    const GmParams& P = G.P;
    int* st = G.st; bbm_ss* v = G.v;
    const int len = G.len, maxSw = G.maxSw;
    const int maxI = maxSw + imin(-472, -395 - 100);
    int pc = st[S_PC];
    int n8 = 0; int sc0 = 0, sc1 = 0, sc2 = 0;          // result of the fill this launch resumes from
    const int8_t* rm = nullptr; int rmlen = -1;
    if (pc >= PC_RA_FILL1 && pc <= PC_RA_FILL4) {
        const bool g = st[S_REQ_GAPPED] != 0;
        const bbm_msa_out o = g ? P.gouts[st[S_REQ]] : P.outs[st[S_REQ]];
        if (o.status != 0) st[S_STATUS] |= BBM_MAP_ST_ALIGNER;
        n8 = (o.status == 0) ? o.score_len : 0;
        sc0 = o.score[0]; sc1 = o.score[1]; sc2 = o.score[2];
        if (pc == PC_RA_FILL1) { st[S_EPL] = (n8 > 6) ? o.score[6] : 0; st[S_EPR] = (n8 > 6) ? o.score[7] : 0; }     // extraPadLeft / extraPadRight live on across the retries
        rm = (g ? P.gmatch + (long long)st[S_REQ] * P.gstride : P.rmatch + (long long)st[S_REQ] * P.rstride);
        rmlen = o.match_len;
    }
    for (;;) {
        switch (pc) {
        case PC_BEGIN: {                               // do { genMatchString } while (top.score < second.score)   (BBMapThread.java:594-613)
            if (G.n == 0) { pc = PC_DONE; break; }
            if (!st[S_FIRST]) { if (P.setSSScore) stable_sort<false>(v, G.n); }
            st[S_BEST] = -0x7fffffff - 1; st[S_CHANGED] = 0; st[S_SITE] = 0;
            pc = PC_GMS_LOOP; break;
        }
        case PC_GMS_LOOP: {
            const int i = st[S_SITE];
            if (i >= G.n || (i > 0 && st[S_BEST] >= v[i].slow_score)) { pc = PC_GMS_SORT; st[S_RET_SITE] = -1; break; }
            st[S_OLDSLOW] = v[i].slow_score; st[S_OLDSCORE] = v[i].score;
            if (v[i].has_match == 0) { st[S_RET_SITE] = PC_GMS_AFTER_SITE; pc = PC_SITE_BEGIN; }
            else { st[S_RET_SITE] = 0; pc = PC_GMS_AFTER_SITE; }
            break;
        }
        case PC_GMS_AFTER_SITE: {
            const int i = st[S_SITE];
            if (st[S_RET_SITE] == PC_GMS_AFTER_SITE && P.setSSScore) v[i].score = v[i].slow_score;
            if (st[S_OLDSCORE] != v[i].score || st[S_OLDSLOW] != v[i].slow_score) st[S_CHANGED]++;
            st[S_BEST] = imax(v[i].slow_score, st[S_BEST]);
            st[S_SITE] = i + 1;
            pc = PC_GMS_LOOP; break;
        }
        case PC_GMS_SORT: {                             // needsSorting loop (:914-940); entered once with ret = -1 to take the decision
            bool needs;
            if (st[S_RET_SITE] == -1) {
                bool ordered = true;
                for (int i = 1; i < G.n; i++) if (v[i].score > v[i - 1].score) { ordered = false; break; }     // Read.CHECKORDER (stream/Read.java:3141-3150)
                needs = st[S_CHANGED] > 0 && !ordered;
            } else needs = true;
            if (!needs) { pc = PC_GMS_FINISH; break; }
            st[S_TOPSERIAL] = v[0].hits >> 16;
            {   // Tools.mergeDuplicateSites(list, false, false)
                stable_sort<true>(v, G.n);
                unsigned long long dead = 0; int ai = 0;
                for (int i = 1; i < G.n; i++) {
                    bbm_ss& a = v[ai]; const bbm_ss& b = v[i];
                    if (positional_match(a, b, true)) {
                        set_slow_score(a, imax(a.slow_score, b.slow_score));
                        a.paired_score = (a.paired_score <= a.slow_score && b.paired_score <= a.slow_score) ? 0 : imax(0, imax(a.paired_score, b.paired_score));
                        a.score = imax(a.score, b.score);
                        a.perfect = (a.perfect || b.perfect) ? 1 : 0; a.semiperfect = (a.semiperfect || b.semiperfect) ? 1 : 0;
                        if (b.has_match) G.free_slot(b.has_match - 1);
                        dead |= 1ull << i;
                    } else ai = i;
                }
                G.n = compact(v, G.n, dead);
            }
            stable_sort<false>(v, G.n);
            if (v[0].has_match == 0) { st[S_SITE] = 0; st[S_RET_SITE] = PC_GMS_AFTER_TOP; pc = PC_SITE_BEGIN; break; }
            if ((v[0].hits >> 16) != st[S_TOPSERIAL]) st[S_TOPCHANGED] = 1;
            pc = PC_GMS_FINISH; break;
        }
        case PC_GMS_AFTER_TOP: {
            if (P.setSSScore) v[0].score = v[0].slow_score;
            if ((v[0].hits >> 16) != st[S_TOPSERIAL]) st[S_TOPCHANGED] = 1;
            st[S_RET_SITE] = 1;                         // needsSorting = true
            pc = PC_GMS_SORT; break;
        }
        case PC_GMS_FINISH: {
            if (!P.setSSScore) { pc = PC_DONE; break; }                      // paired reads call genMatchString once (BBMapThread.java:1193-1215)
            v[0].score = v[0].slow_score;                                    // r.topSite().setScore(r.topSite().slowScore)
            st[S_FIRST] = 0;
            if (G.n > 1 && v[0].score < v[1].score) pc = PC_BEGIN; else pc = PC_DONE;
            break;
        }
        // ---------------- genMatchStringForSite ----------------
        case PC_SITE_BEGIN: {
            bbm_ss& ss = v[st[S_SITE]];
            const float mult = P.cfg.paired ? P.cfg.min_ratio_paired : P.cfg.min_ratio;     // secondary = false: x 1f
            st[S_MINMSA] = -1 + (int)__fmul_rn(mult, (float)maxSw);
            const int s = G.alloc_slot();
            if (s < 0) { st[S_STATUS] |= BBM_MAP_ST_SLOTS; pc = PC_DONE; break; }            // flagged; the read keeps what it has
            ss.has_match = s + 1;
            if (ss.perfect) {
                if (len > P.ms) { st[S_STATUS] |= BBM_MAP_ST_MATCH_OVERFLOW; G.slen(s) = 0; }
                else { int8_t* m = G.slot(s); for (int i = 0; i < len; i++) m[i] = 'm'; G.slen(s) = len; }
                pc = PC_SITE_END; break;
            }
            st[S_SITE_OLDSCORE] = ss.slow_score;
            st[S_PADDING] = ss.semiperfect ? 0 : imax(P.cfg.slow_align_padding, 6);
            st[S_RECUR] = 1; st[S_FORBID] = P.cfg.max_indel < 1; st[S_FIXXY] = 0;
            st[S_RET_RA] = PC_SITE_AFTER_R1; pc = PC_RA_BEGIN; break;
        }
        case PC_SITE_AFTER_R1: {
            bbm_ss& ss = v[st[S_SITE]];
            const int s = ss.has_match - 1;
            if (ss.ngaps > 0) ss.ngaps = fix_gaps(ss.start, ss.stop, ss.gaps, ss.ngaps, SL_MINGAP);
            const int lp = gm_left_padding(G.slot(s), G.slen(s), 4, 5), rp = gm_right_padding(G.slot(s), G.slen(s), 4, 5);
            if (ss.slow_score < st[S_SITE_OLDSCORE] || lp > 0 || rp > 0) {
                int extra = (P.cfg.max_indel > 0 ? 80 : 20) + P.cfg.slow_align_padding;
                const int remaining = GM_MAXCOLS - calc_gref_len(ss) - 2;
                extra = imax(0, imin(remaining / 2, extra));
                st[S_PADDING] = extra; st[S_RECUR] = 2; st[S_FORBID] = 0; st[S_FIXXY] = 1;
                st[S_RET_RA] = PC_SITE_AFTER_R2; pc = PC_RA_BEGIN; break;
            }
            pc = PC_SITE_AFTER_R2; st[S_RET_RA] = 0; break;
        }
        case PC_SITE_AFTER_R2: {
            bbm_ss& ss = v[st[S_SITE]];
            if (st[S_RET_RA] == PC_SITE_AFTER_R2 && ss.ngaps > 0) ss.ngaps = fix_gaps(ss.start, ss.stop, ss.gaps, ss.ngaps, SL_MINGAP);
            if (maxSw == ss.slow_score) { ss.perfect = 1; ss.semiperfect = 1; }                // setPerfectFlag(maxSwScore, bases)
            else { int refLen; const int8_t* ref = G.ref(ss, refLen); ss_set_perfect(ss, G.bases(ss), len, ref, refLen); }
            pc = PC_SITE_END; break;
        }
        case PC_SITE_END: {
            bbm_ss& ss = v[st[S_SITE]];
            const int s = ss.has_match - 1;
            int refLen; const int8_t* ref = G.ref(ss, refLen);
            gm_clip_tip_indels(ss, G.slot(s), G.slen(s), G.bases(ss), len, ref, refLen, 4, 10);
            pc = st[S_RET_SITE]; break;
        }
        // ---------------- realign_new ----------------
        case PC_RA_BEGIN: {
            bbm_ss& ss = v[st[S_SITE]];
            const int s = ss.has_match - 1;
            int8_t* m = G.slot(s); int& ml = G.slen(s);
            int refLen; const int8_t* ref = G.ref(ss, refLen);
            const int8_t* bases = G.bases(ss);
            const int maxIndex = refLen - 1;
            if (gm_contains_xy(m, ml)) gm_fix_xy(ss, m, ml, bases, len, ref, refLen);
            gm_clip_tip_indels(ss, m, ml, bases, len, ref, refLen, 4, 10);
            int padding = imax(imin(st[S_PADDING], (GM_MAXCOLS - len) / 2 - 20), 0);
            if (calc_gref_len(ss) > GM_MAXCOLS - 20) {
                ss_set_stop(ss, ss.start + imin(len + 40, GM_MAXCOLS - 20));
                if (ss.ngaps > 0) ss.ngaps = fix_gaps(ss.start, ss.stop, ss.gaps, ss.ngaps, SL_MINGAP);
            }
            if (ss.start < 0) ss_set_start(ss, 0);
            if (ss.stop > maxIndex) ss_set_stop(ss, maxIndex);
            { const int b = ss.stop - ss.start + 1; if (b < len) { const int c = imin(len, len - b + 10) / 2; padding = imax(padding, c + 1); } }
            padding = imax(0, imin(padding, (GM_MAXCOLS - imax(len, calc_gref_len(ss))) / 2 - 100));
            if (st[S_FORBID]) padding = 0;
            if (len > P.ms) { st[S_STATUS] |= BBM_MAP_ST_MATCH_OVERFLOW; ml = 0; pc = st[S_RET_RA]; break; }
            if (ml != len) { for (int i = 0; i < len; i++) m[i] = 0; ml = len; }
            const int scoreNoIndel = gm_noindel_match(bases, len, ref, refLen, ss.start, m);
            st[S_NOINDEL] = scoreNoIndel;
            if (scoreNoIndel >= maxI || st[S_FORBID]) {
                ss_set_stop(ss, ss.start + len - 1);
                set_slow_score(ss, scoreNoIndel);
                pc = PC_RA_AFTER; break;
            }
            st[S_MINLOC] = imax(ss.start - padding, 0); st[S_MAXLOC] = imin(ss.stop + padding, maxIndex);
            st[S_LIM] = imax(scoreNoIndel, st[S_MINMSA]);
            gm_request(G, ss, st[S_MINLOC], st[S_MAXLOC], st[S_LIM], false);
            st[S_PC] = PC_RA_FILL1; P.nss[G.r] = G.n; return;
        }
        case PC_RA_FILL1: case PC_RA_FILL2: case PC_RA_FILL3: case PC_RA_FILL4: {
            bbm_ss& ss = v[st[S_SITE]];
            const int s = ss.has_match - 1;
            int refLen; G.ref(ss, refLen);
            const int maxIndex = refLen - 1;
            const bool gapped = ss.ngaps > 0;
            bool again = false;
            if (pc == PC_RA_FILL1 && n8 > 6) {
                st[S_OLD0] = sc0;
                int epl = st[S_EPL], epr = st[S_EPR];
                gm_adjust_pads(gapped, gapped ? imax(len, gm_gref_len(st[S_MINLOC], st[S_MAXLOC], ss)) : 0, st[S_MAXLOC] - st[S_MINLOC] + 1, epl, epr, gapped || ss.strand == 0);
                st[S_EPL] = epl; st[S_EPR] = epr;
                st[S_MINLOC] = imax(0, st[S_MINLOC] - epl); st[S_MAXLOC] = imin(maxIndex, st[S_MAXLOC] + epr);
                gm_request(G, ss, st[S_MINLOC], st[S_MAXLOC], st[S_LIM], false);
                st[S_PC] = PC_RA_FILL2; again = true;
            } else if (pc == PC_RA_FILL2 && (n8 == 0 || sc0 < st[S_OLD0])) {
                int epl = st[S_EPL], epr = st[S_EPR];          // the pads of the second fill carry over, not what the second fill suggested (:416-445)
                gm_adjust_pads(gapped, gapped ? imax(len, gm_gref_len(st[S_MINLOC], st[S_MAXLOC], ss)) : 0, st[S_MAXLOC] - st[S_MINLOC] + 1, epl, epr, true);
                st[S_EPL] = epl; st[S_EPR] = epr;
                st[S_MINLOC] = imax(0, st[S_MINLOC] - epl); st[S_MAXLOC] = imin(maxIndex, st[S_MAXLOC] + epr);
                gm_request(G, ss, st[S_MINLOC], st[S_MAXLOC], st[S_LIM], false);
                st[S_PC] = PC_RA_FILL3; again = true;
            } else if (pc == PC_RA_FILL3 && ss.strand == 0 && st[S_MINLOC] > 0 && st[S_MAXLOC] < maxIndex && (n8 == 0 || sc0 < st[S_OLD0])) {
                st[S_MINLOC] = imax(ss.start - 8, 0); st[S_MAXLOC] = imin(ss.stop + 8, maxIndex);
                gm_request(G, ss, st[S_MINLOC], st[S_MAXLOC], 0, true);
                st[S_PC] = PC_RA_FILL4; again = true;
            }
            if (again) { P.nss[G.r] = G.n; return; }
            // this fill is the last one: max / score are its result
            if (n8 > 0) {
                int8_t* m = G.slot(s);
                if (rmlen < 0 || rmlen > P.ms) { st[S_STATUS] |= BBM_MAP_ST_MATCH_OVERFLOW; G.slen(s) = 0; }
                else { for (int i = 0; i < rmlen; i++) m[i] = rm[i]; G.slen(s) = rmlen; }
                ss_set_limits(ss, sc1, sc2);
                gm_fix_limits_xy(ss, m, G.slen(s));
                set_slow_score(ss, sc0);
            } else {
                ss_set_stop(ss, ss.start + len - 1);
                set_slow_score(ss, st[S_NOINDEL]);
            }
            pc = PC_RA_AFTER; break;
        }
        case PC_RA_AFTER: {
            bbm_ss& ss = v[st[S_SITE]];
            const int s = ss.has_match - 1;
            int refLen; const int8_t* ref = G.ref(ss, refLen);
            const int maxIndex = refLen - 1;
            const int lp = gm_left_padding(G.slot(s), G.slen(s), 4, 5), rp = gm_right_padding(G.slot(s), G.slen(s), 4, 5);
            if (ss.stop < maxIndex && ss.start > 0 && (lp > 0 || rp > 0)) {
                if (st[S_RECUR] > 0) {
                    if (ss.ngaps > 0) ss.ngaps = fix_gaps(ss.start, ss.stop, ss.gaps, ss.ngaps, SL_MINGAP);
                    st[S_PADDING] = imin(10 + imax(lp, rp), (GM_MAXCOLS - len) / 2 - 20);
                    st[S_RECUR]--;
                    pc = PC_RA_BEGIN; break;                 // the recursive call; the caller's trailing setPerfect repeats what the callee's did
                } else if (st[S_FIXXY] && gm_contains_xy(G.slot(s), G.slen(s))) gm_fix_xy(ss, G.slot(s), G.slen(s), G.bases(ss), len, ref, refLen);
            }
            ss_set_perfect(ss, G.bases(ss), len, ref, refLen);
            pc = st[S_RET_RA]; break;
        }
        default: pc = PC_DONE; break;
        }
        if (pc == PC_DONE) { st[S_PC] = PC_DONE; P.nss[G.r] = G.n; return; }
    }
}

__global__ void __launch_bounds__(128) genmatch_kernel(GmParams P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    int* st = P.state + r * GM_STATE;
    if (P.first) {
        for (int i = 0; i < GM_STATE; i++) st[i] = 0;
        st[S_PC] = PC_BEGIN; st[S_FIRST] = 1;
        bbm_ss* v = P.lists + r * P.cap;
        const int n = P.nss[r];
        for (int i = 0; i < n; i++) { v[i].has_match = 0; v[i].hits = (v[i].hits & 0xffff) | (i << 16); }      // serial number = object identity of the SiteScore
        for (int s = 0; s < GM_SLOTS; s++) P.mlen[r * GM_SLOTS + s] = 0;
    }
    if (st[S_PC] == PC_DONE) return;
    const int len = (int)(P.read_off[r + 1] - P.read_off[r]);
    GmRead G = { P, r, st, P.lists + r * P.cap, P.nss[r], len, max_quality(len) };
    gm_run(G);
}

// ---------------- after genMatchString: the rest of processRead (BBMapThread.java:624-709) ----------------

__global__ void __launch_bounds__(128) map_finish_kernel(FinParams P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    bbm_ss* v = P.lists + r * P.cap;
    int n = P.nss[r];
    const int len = (int)(P.read_off[r + 1] - P.read_off[r]);
    const int maxSw = max_quality(len);
    const bbm_policy_cfg& cfg = P.pc;
    int flags = P.flags[r].flags & 7;
    int status = P.state[r * GM_STATE + S_STATUS];
    int mapScore = n > 0 ? v[0].slow_score : 0, rstart = -1, rstop = -1;
    if (n > 0) { flags = (flags & ~2) | (v[0].perfect ? 2 : 0); rstart = v[0].start; rstop = v[0].stop; }
    if (n > 1) {                                                         // removeDuplicateBestSites
        const bbm_ss t = v[0];
        while (n > 1 && t.chrom == v[n - 1].chrom && t.strand == v[n - 1].strand && t.start == v[n - 1].start && t.stop == v[n - 1].stop) n--;
    }
    if (n > 0 && mapScore <= 0) { mapScore = 0; n = 0; }                 // "failed cigar string generation" (:630-638)
    if (n == 0) flags &= ~1;
    int subi = 0;
    if ((cfg.clearzone3 > cfg.clearzone1 || cfg.clearzone3 > cfg.clearzonep) && n > 0 && !(flags & 4)) {
        const float q = __fdiv_rn((float)maxSw, (float)mapScore);
        const float cz3v2 = __fmul_rn((float)cfg.clearzone3, 1.25f < q ? 1.25f : q);
        const int cz3 = (int)cz3v2; const float inv = __fdiv_rn(1.f, cz3v2);
        if ((flags & 1) && n >= 2) {                                     // applyClearzone3
            float sub = 0.f;
            const int mx = imin(7, n);
            for (int i = 1; i < mx; i++) {
                if (i > 2 && v[i].slow_score < v[i - 1].slow_score) break;
                const float f = calc_cz3_fraction(mapScore, v[i].slow_score, cz3, inv);
                if (f <= 0.f) break;
                sub = __fadd_rn(sub, __fmul_rn(f, cz3_mult(i)));
            }
            if (sub > 0.f) {
                const float asym = __fadd_rn(4.f, __fmul_rn(0.03f, (float)len));
                sub = __fmul_rn(sub, 1.8f);
                const float sub2 = __fmul_rn((float)cz3, __fdiv_rn(__fmul_rn(asym, sub), __fadd_rn(sub, asym)));
                subi = (int)__fadd_rn(sub2, 0.5f);
                if (subi >= mapScore - 300) subi = mapScore - 300;
                if (subi <= 0) subi = 0;
                else {
                    for (int i = 0; i < n; i++) { bbm_ss ss = v[i]; set_slow_score(ss, ss.slow_score - subi); ss.score -= subi; v[i] = ss; }
                    mapScore -= subi;
                    if (mapScore < (int)__fmul_rn((float)maxSw, cfg.min_align_ratio)) flags |= 4;
                }
            }
        }
    }
    if ((flags & 4) && P.cfg.ambiguous_toss) { n = 0; flags &= ~1; mapScore = 0; }
    int slot = (n > 0 && v[0].has_match) ? v[0].has_match - 1 : -1;
    if ((flags & 1) && n > 0 && slot >= 0 && P.mlen[r * GM_SLOTS + slot] > 0) {
        int8_t* m = P.mslots + (r * GM_SLOTS + slot) * P.ms; int& ml = P.mlen[r * GM_SLOTS + slot];
        const int8_t a = m[0], b = m[ml - 1];
        if (a == 'X' || b == 'Y' || a == 'C' || b == 'C') {              // r.containsXYC(); LOCAL_ALIGN is off by default
            bbm_ss top = v[0];
            const int8_t* bases = (top.strand == 0 ? P.basesP : P.basesM) + P.read_off[r];
            const int refLen = (int)(P.chrom_off[top.chrom] - P.chrom_off[top.chrom - 1]);
            const bool ok = gm_to_local(top, m, ml, P.ms, bases, len, P.refs + P.chrom_off[top.chrom - 1], refLen, 1, rstart, rstop, mapScore, flags, status);
            v[0] = top;
            if (!ok) { n = 0; flags &= ~1; mapScore = 0; }
        }
    }
    if (n == 0 || (!(flags & 4) && (float)mapScore < __fmul_rn((float)maxSw, cfg.min_align_ratio))) { n = 0; flags &= ~1; mapScore = 0; }   // r.clearMapping()
    int pen = 0;
    if (P.cfg.penalize_ambig && n > 0 && (flags & 1) && slot >= 0 && len >= 14) {        // calcTipScorePenalty(r, maxSwScore, 7) + applyScorePenalty
        const int tiplen = 7;
        const int8_t* match = P.mslots + (r * GM_SLOTS + slot) * P.ms; const int mlen = P.mlen[r * GM_SLOTS + slot];
        const int8_t* bases = P.basesP + P.read_off[r];
        int points = 0; bool bad = mlen < 1;
        int8_t prev = 'm';
        for (int i = 0, cpos = 0; cpos <= tiplen && !bad; i++) {
            if (i >= mlen) { bad = true; break; }
            const int8_t b = match[i];
            if (b == 'm') cpos++;
            else if (b == 'D') { if (prev != 'D') points += 2 * (tiplen + 2 - cpos); }
            else if (b == 'N' || b == 'C') { points += (tiplen + 2 - cpos); cpos++; }
            else { points += 2 * (tiplen + 2 - cpos); cpos++; }
            prev = b;
        }
        prev = 'm';
        for (int i = mlen - 1, cpos = 0; cpos <= tiplen && !bad; i--) {
            if (i < 0) { bad = true; break; }
            const int8_t b = match[i];
            if (b == 'm') cpos++;
            else if (b == 'D') { if (prev != 'D') points += 2 * (tiplen + 2 - cpos); }
            else if (b == 'N' || b == 'C') { points += (tiplen + 2 - cpos); cpos++; }
            else { points += 2 * (tiplen + 2 - cpos); cpos++; }
            prev = b;
        }
        if (bad) { if (mlen >= 1) status |= BBM_MAP_ST_TIP; }
        else {
            const int last = len - 1;
            int8_t b = bases[0];
            if (b != 'N' && b == bases[1]) for (int i = 2; i <= tiplen && bases[i] == b; i++) points++;
            b = bases[last];
            if (b != 'N' && b == bases[last - 1]) for (int i = last - 2; i >= (last - tiplen) && bases[i] == b; i--) points++;
            if (points >= 1) {
                const float f = __fdiv_rn(__fmul_rn(80.f, (float)points), __fadd_rn((float)points, 80.f));
                const int penalty = (int)__fmul_rn(__fmul_rn(f, .0022f), (float)maxSw);
                const int maxPenalty = mapScore - maxSw / 10;
                if (maxPenalty > 0) pen = imin(penalty, maxPenalty);
            }
        }
        if (pen > 0) { mapScore -= pen; for (int i = 0; i < n; i++) { bbm_ss ss = v[i]; set_slow_score(ss, ss.slow_score - pen); ss.score -= pen; v[i] = ss; } }
    }
    bbm_map_rec rec = {};
    rec.flags = flags; rec.map_score = mapScore; rec.cz3_sub = subi; rec.tip_penalty = pen; rec.status = status;
    if (n > 0 && (flags & 1)) {
        rec.chrom = v[0].chrom; rec.strand = v[0].strand; rec.start = rstart; rec.stop = rstop;
        rec.match_len = slot >= 0 ? P.mlen[r * GM_SLOTS + slot] : 0;
        rec.match_slot = slot;
    } else { rec.chrom = -1; rec.start = -1; rec.stop = -1; rec.strand = 0; rec.match_len = 0; rec.match_slot = -1; }
    for (int i = 0; i < n; i++) { v[i].hits &= 0xffff; v[i].has_match = v[i].has_match ? 1 : 0; }
    P.nss[r] = n;
    P.recs[r] = rec;
}

// Read fields -> the record SamLine(Read,int) reads
__global__ void __launch_bounds__(128) map_sam_tasks_kernel(const bbm_map_rec* __restrict__ recs, long long nreads, const long long* __restrict__ read_off, long long ms,
                                                            int paired, bbm_sam_task* __restrict__ tasks) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    const bbm_map_rec q = recs[r];
    bbm_sam_task t;
    t.match_off = (r * GM_SLOTS + (q.match_slot < 0 ? 0 : q.match_slot)) * ms;
    t.match_len = q.match_len; t.chrom = q.chrom; t.start = q.start; t.stop = q.stop;
    t.read_len = (int)(read_off[r + 1] - read_off[r]);
    t.score = q.map_score;
    t.mate = paired ? (int)(r ^ 1) : -1;
    const int f = q.flags;
    t.flags = ((f & 1) ? BBM_RF_MAPPED : 0) | (((f & 1) && q.strand == 1) ? BBM_RF_MINUS : 0) | ((f & 2) ? BBM_RF_PERFECT : 0) | ((f & 4) ? BBM_RF_AMBIGUOUS : 0) |
              ((f & 32) ? BBM_RF_DISCARDED : 0) | ((f & 8) ? BBM_RF_PAIRED : 0) | ((paired && (r & 1)) ? BBM_RF_PAIRNUM1 : 0);
    t.pad_ = 0;
    tasks[r] = t;
}

// ---------------- small glue kernels of the chain ----------------
// processRead :455-465 / processReadPair :1023-1052: scoreSlow runs when scoreNoIndels found no near-perfect site (always for pairs), findTipDeletions
// only without a near-perfect site
__global__ void __launch_bounds__(256) map_runmask_kernel(const bbm_read_out* __restrict__ out, const int* __restrict__ nss, long long n, int paired,
                                                          int* __restrict__ run, int* __restrict__ masked) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int none = out[r].near_perfect < 1;
    run[r] = (paired || none) ? 1 : 0;
    masked[r] = none ? nss[r] : 0;
}
__global__ void __launch_bounds__(256) map_arange_kernel(long long* __restrict__ off, long long n, long long stride) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k <= n) off[k] = k * stride;
}
__global__ void __launch_bounds__(256) map_overflow_kernel(const bbm_search_head* __restrict__ heads, long long n, int maxSites, int* __restrict__ counter) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (heads[r].nsites > maxSites || (heads[r].status & BBM_ST_SITE_OVERFLOW)) atomicAdd(counter, 1);
}
__global__ void __launch_bounds__(256) map_status_kernel(const bbm_search_head* __restrict__ heads, int maxSites, const int* __restrict__ slowStatus, const int* __restrict__ nkeys,
                                                         bbm_map_rec* __restrict__ recs, long long n, unsigned long long* __restrict__ counters) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    bbm_map_rec q = recs[r];
    if (heads[r].nsites > maxSites || (heads[r].status & BBM_ST_SITE_OVERFLOW)) { q.status |= BBM_MAP_ST_SITE_OVERFLOW; atomicAdd(counters + 2, 1ull); }
    if (slowStatus[r]) q.status |= BBM_MAP_ST_SLOW;
    if (nkeys[r] < 0) q.flags |= 32;                                     // quickMap returned < 0: r.setDiscarded(true) (:409-415)
    if (q.flags & 1) atomicAdd(counters, 1ull);
    if (q.status) atomicAdd(counters + 1, 1ull);
    recs[r] = q;
}
// primary match strings -> the caller's fixed-stride buffer (one warp per read)
__global__ void __launch_bounds__(256) map_copy_match_kernel(const bbm_map_rec* __restrict__ recs, const int8_t* __restrict__ mslots, long long ms, int8_t* __restrict__ out,
                                                             long long stride, long long n) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= n) return;
    const bbm_map_rec q = recs[r];
    if (q.match_slot < 0) return;
    const int8_t* m = mslots + (r * GM_SLOTS + q.match_slot) * ms;
    const long long len = q.match_len < stride ? q.match_len : stride;
    for (long long i = lane; i < len; i += 32) out[r * stride + i] = m[i];
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_genmatch_state_ints() { return GM_STATE; }
extern "C" int bbm_genmatch_slots() { return GM_SLOTS; }
extern "C" int bbm_launch_genmatch(const GmParams* P, cudaStream_t st) {
    genmatch_kernel<<<(unsigned)((P->nreads + 127) / 128), 128, 0, st>>>(*P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_map_finish(const FinParams* P, cudaStream_t st) {
    map_finish_kernel<<<(unsigned)((P->nreads + 127) / 128), 128, 0, st>>>(*P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_map_sam_tasks(const bbm_map_rec* recs, long long nreads, const long long* read_off, long long ms, int paired, bbm_sam_task* tasks, cudaStream_t st) {
    map_sam_tasks_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(recs, nreads, read_off, ms, paired, tasks);
    return (int)cudaGetLastError();
}

extern "C" int bbm_launch_map_runmask(const bbm_read_out* out, const int* nss, long long n, int paired, int* run, int* masked, cudaStream_t st) {
    map_runmask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(out, nss, n, paired, run, masked);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_map_arange(long long* off, long long n, long long stride, cudaStream_t st) {
    map_arange_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(off, n, stride);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_map_overflow(const bbm_search_head* heads, long long n, int maxSites, int* counter, cudaStream_t st) {
    map_overflow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(heads, n, maxSites, counter);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_map_status(const bbm_search_head* heads, int maxSites, const int* slowStatus, const int* nkeys, bbm_map_rec* recs, long long n,
                                     unsigned long long* counters, cudaStream_t st) {
    map_status_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(heads, maxSites, slowStatus, nkeys, recs, n, counters);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_map_copy_match(const bbm_map_rec* recs, const int8_t* mslots, long long ms, int8_t* out, long long stride, long long n, cudaStream_t st) {
    map_copy_match_kernel<<<(unsigned)((n * 32 + 255) / 256), 256, 0, st>>>(recs, mslots, ms, out, stride, n);
    return (int)cudaGetLastError();
}
