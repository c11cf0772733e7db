/* TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of BBMapThread.scoreSlow (current/align2/BBMapThread.java:252-386) for the default
 * flag set (QUICK_MATCH_STRINGS=false), one read after the other and one site after the other exactly as the reference walks them, on top of
 * the oracle's MultiStateAligner11ts (msa_oracle.c: MSA.fillAndScoreLimited, MSA.java:103-143, gapped references included), with SiteScore.setLimits /
 * setStop and GapTools.fixGaps for sites that carry a gap array.  Parity UNPINNED against Java (no JVM). */
#include <string.h>
#include "msa_oracle.h"
#include "sitelist_oracle.h"

typedef struct { int32_t paired; float min_ratio, min_ratio_pre_rescue; int32_t clearzone1e, clearzone3, slow_align_padding, extra_padding, expected_len_limit; } orc_slow_cfg;

static int imax2(int a, int b) { return a > b ? a : b; }
static int imin2(int a, int b) { return a < b ? a : b; }

/* SiteScore.setPerfect(bases), as in sitelist_oracle.c */
static void set_perfect(orc_ss* s, const int8_t* bases, int len, const int8_t* ref, int refLen)
{
    if (len != s->stop - s->start + 1) { s->perfect = 0; s->semiperfect = 0; return; }
    int perfect = 1, semiperfect = 1, refloc = s->start, readloc = 0, N = 0;
    const int max = imin2(s->stop, refLen - 1), nlimit = len / 2;
    if (s->start < 0) { N -= s->start; readloc -= s->start; refloc -= s->start; perfect = 0; }
    if (s->stop >= refLen) { N += (s->stop - refLen + 1); perfect = 0; }
    if (N > nlimit) { s->perfect = 0; s->semiperfect = 0; return; }
    for (; refloc <= max; refloc++, readloc++) {
        const int8_t c = bases[readloc], r = ref[refloc];
        if (c != r || c == 'N') {
            perfect = 0;
            if (c == 'N') semiperfect = 0;
            if (r != 'N' || (N = N + 1) > nlimit) { s->perfect = 0; s->semiperfect = 0; return; }
        }
    }
    semiperfect = (semiperfect && (N <= nlimit));
    perfect = (perfect && semiperfect && (N == 0));
    s->perfect = (int8_t)perfect; s->semiperfect = (int8_t)semiperfect;
}

int64_t orc_score_slow_with(orc_msa* msa, orc_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int64_t* read_off,
                            const int8_t* refs, const int64_t* chrom_off, const int32_t* run, const orc_slow_cfg* cfg, int32_t* status)
{
    int64_t alignments = 0;
    for (int64_t r = 0; r < nreads; r++) {
        if (status) status[r] = 0;
        if (!run[r]) continue;
        const int len = (int)(read_off[r + 1] - read_off[r]);
        const int maxSwScore = 70 + (len - 1) * 100, maxImperfectSwScore = maxSwScore + imin2(-472, -395 - 100);
        int minMsaLimit = -cfg->clearzone1e + (int)((cfg->paired ? cfg->min_ratio_pre_rescue : cfg->min_ratio) * (float)maxSwScore);
        int minMatch = imax2(-300, minMsaLimit - cfg->clearzone3);
        for (int i = 0; i < nss[r]; i++) {
            orc_ss* ss = &lists[r * cap + i];
            const int8_t* bases = (ss->strand == 0 ? basesP : basesM) + read_off[r];
            const int8_t* ref = refs + chrom_off[ss->chrom - 1]; const int refLen = (int)(chrom_off[ss->chrom] - chrom_off[ss->chrom - 1]);
            if (ss->stop - ss->start != len - 1) { orc_ss_set_slow_score(ss, 0); ss->semiperfect = 0; ss->perfect = 0; }
            const int swscoreNoIndel = ss->slow_score;
            int32_t arr[8], old[8], max4[4]; int n = 0;
            if (swscoreNoIndel < maxImperfectSwScore && !ss->semiperfect) {
                const int expectedLen = orc_calc_gref_len(ss);
                if (expectedLen >= cfg->expected_len_limit) orc_ss_set_stop(ss, ss->start + imin2(len + 40, cfg->expected_len_limit));
                int pad = cfg->slow_align_padding;
                const int minscore = imax2(swscoreNoIndel, minMsaLimit);
                int32_t g[ORC_MAX_GAPS];
                memcpy(g, ss->gaps, sizeof(ss->gaps));
                n = orc_msa_fillAndScoreLimited(msa, bases, len, ref, refLen, ss->start - pad, ss->stop + pad, minscore, ss->ngaps ? g : 0, ss->ngaps, max4, arr);
                alignments++;
                if (n > 6 && (arr[3] + arr[4] + expectedLen < cfg->expected_len_limit)) {
                    const int oldn = n; memcpy(old, arr, sizeof(arr));
                    orc_ss_set_limits(ss, ss->start - arr[6], ss->stop + arr[7]);
                    pad = cfg->slow_align_padding + cfg->extra_padding;
                    memcpy(g, ss->gaps, sizeof(ss->gaps));
                    n = orc_msa_fillAndScoreLimited(msa, bases, len, ref, refLen, ss->start - pad, ss->stop + pad, minscore, ss->ngaps ? g : 0, ss->ngaps, max4, arr);
                    alignments++;
                    if (n == 0 || arr[0] < old[0]) { n = oldn; memcpy(arr, old, sizeof(arr)); }
                }
            }
            if (n > 0) { orc_ss_set_slow_score(ss, arr[0]); orc_ss_set_limits(ss, arr[1], arr[2]); }
            ss->score = ss->slow_score;
            minMatch = imax2(minMatch, ss->slow_score);
            minMsaLimit = imax2(minMsaLimit, ss->slow_score - cfg->clearzone3);
            ss->perfect = (ss->slow_score == maxSwScore);
            if (ss->perfect) ss->semiperfect = 1;
            else if (!ss->semiperfect) set_perfect(ss, bases, len, ref, refLen);
        }
        (void)minMatch;
    }
    return alignments;
}

int64_t orc_score_slow(orc_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int64_t* read_off,
                       const int8_t* refs, const int64_t* chrom_off, const int32_t* run, const orc_slow_cfg* cfg, int32_t* status)
{
    orc_msa* msa = orc_msa_new(601, 3000);
    const int64_t a = orc_score_slow_with(msa, lists, nss, nreads, cap, basesP, basesM, read_off, refs, chrom_off, run, cfg, status);
    orc_msa_free(msa);
    return a;
}
