/* TEST INFRASTRUCTURE ONLY — see sitelist_oracle.h.  Restates, statement for statement, on an array that plays the ArrayList:
 *   BBMapThread.processRead, the list handling around the alignment stages      current/align2/BBMapThread.java:420-431, 440-443, 478-553
 *   BBMapThread.trimList (affine branch)                                        :140-249
 *   Tools.trimSiteList / trimSitesBelowCutoff / condenseStrict                  current/align2/Tools.java:654-673, 1106-1160, 542-566
 *   Tools.mergeDuplicateSites / countTopScores / removeLowQualitySitesUnpaired  :697-760, 913-931, 986-1003
 *   SiteScore.compareTo, PositionComparator, positionalMatch, setPerfect        current/stream/SiteScore.java:55-76, 379-395, 353-365, 239-291
 *   AbstractMapThread.scoreNoIndels(Read, ...)                                  current/align2/AbstractMapThread.java:762-855
 *   Read.setPerfectFlag (match == null at this point)                           current/stream/Read.java:2494-2512
 *   MSA.maxQuality / maxImperfectScore                                          current/align2/MultiStateAligner11tsJNI.java:1321-1336
 */
#include "sitelist_oracle.h"
#include "host_oracle.h"
#include "rescue_oracle.h"
#include <string.h>

static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* SiteScore.compareTo: higher scores first */
static int ss_compare(const orc_ss* a, const orc_ss* o) {
    int x = o->score - a->score; if (x) return x;
    x = o->slow_score - a->slow_score; if (x) return x;
    x = o->paired_score - a->paired_score; if (x) return x;
    x = o->quick_score - a->quick_score; if (x) return x;
    x = a->chrom - o->chrom; if (x) return x;
    return a->start - o->start;
}
/* SiteScore.PCOMP */
static int ss_pcomp(const orc_ss* a, const orc_ss* b) {
    if (a->chrom != b->chrom) return a->chrom - b->chrom;
    if (a->start != b->start) return a->start - b->start;
    if (a->stop != b->stop) return a->stop - b->stop;
    if (a->strand != b->strand) return a->strand - b->strand;
    if (a->score != b->score) return b->score - a->score;
    if (a->slow_score != b->slow_score) return b->slow_score - a->slow_score;
    if (a->quick_score != b->quick_score) return b->quick_score - a->quick_score;
    if (a->perfect != b->perfect) return a->perfect ? -1 : 1;
    if (a->rescued != b->rescued) return a->rescued ? 1 : -1;
    return 0;
}
/* Collections.sort is stable: insertion sort keeps equal elements in order */
static void stable_sort(orc_ss* v, int n, int (*cmp)(const orc_ss*, const orc_ss*)) {
    for (int i = 1; i < n; i++) {
        orc_ss x = v[i]; int j = i - 1;
        while (j >= 0 && cmp(&v[j], &x) > 0) { v[j + 1] = v[j]; j--; }
        v[j + 1] = x;
    }
}

/* Tools.trimSitesBelowCutoff with retainSemiperfect = true */
static int trim_below_cutoff(orc_ss* v, int n, int cutoff, int retainPaired, int minS, int maxS) {
    if (n <= minS) return n;
    while (n > maxS) n--;
    int removed = 0; const int maxToRemove = n - minS;
    char dead[4096]; memset(dead, 0, (size_t)n);
    for (int i = n - 1; i >= 0; i--) {
        const orc_ss* ss = &v[i];
        if (!ss->semiperfect) {
            if (ss->score < cutoff && (!retainPaired || ss->paired_score <= 0)) {
                dead[i] = 1; removed++;
                if (removed >= maxToRemove) break;
            }
        }
    }
    if (removed > 0) { int k = 0; for (int i = 0; i < n; i++) if (!dead[i]) v[k++] = v[i]; n = k; }
    return n;
}
/* Tools.trimSiteList; *n is updated, returns maxScore */
static int trim_site_list(orc_ss* v, int* n, float frac, int retainPaired, int minS, int maxS) {
    if (*n == 0) return -999999;
    if (*n == 1) return v[0].score;
    int maxScore = -999999;
    if (minS > 1 && minS < *n) maxScore = v[0].score;
    else for (int i = 0; i < *n; i++) maxScore = imax(maxScore, v[i].score);
    const int cutoff = (int)((float)maxScore * frac);
    *n = trim_below_cutoff(v, *n, cutoff, retainPaired, minS, maxS);
    return maxScore;
}
/* BBMapThread.trimList, USE_AFFINE_SCORE branch */
static int trim_list(orc_ss* v, int* n, int retainPaired, int maxScore, int specialCasePerfect, int minS, int maxS) {
    if (*n == 0) return -99999;
    if (*n == 1) return v[0].score;
    const int highest = trim_site_list(v, n, .6f, retainPaired, minS, maxS);
    if (highest == maxScore && specialCasePerfect) {
        trim_site_list(v, n, .94f, retainPaired, minS, maxS);
        if (*n > 8) trim_site_list(v, n, .99f, retainPaired, minS, maxS);
        return highest;
    }
    const int mstr2 = (minS <= 1 ? 1 : minS + 1);
    if (*n > 4) trim_site_list(v, n, .65f, retainPaired, minS, maxS);
    if (*n > 8) trim_site_list(v, n, .7f, retainPaired, minS, maxS);
    if (*n > 12) trim_site_list(v, n, .75f, retainPaired, minS, maxS);
    if (*n > 16) trim_site_list(v, n, .8f, retainPaired, minS, maxS);
    if (*n > 20) trim_site_list(v, n, .85f, retainPaired, minS, maxS);
    if (*n > 24) trim_site_list(v, n, .9f, retainPaired, minS, maxS);
    if (*n > 32) trim_site_list(v, n, .95f, retainPaired, minS, maxS);
    if (*n > 40) trim_site_list(v, n, .97f, retainPaired, mstr2, maxS);
    if (*n > 48) trim_site_list(v, n, .99f, retainPaired, mstr2, maxS);
    return highest;
}

/* GapTools.fixGaps(a, b, gaps, minGap) + fixGaps2 (current/align2/GapTools.java:26-72, 126-175).  Works in place on gaps[0..n); returns the new number of
 * ints, 0 = null. */
int orc_fix_gaps(int a, int b, int32_t* gaps, int n, int minGap)
{
    if (n == 0) return 0;
    if (!(gaps[0] <= b && gaps[n - 1] >= a)) return 0;                 /* Tools.overlap(a, b, g0, gN) */
    int changed = 0;
    if (gaps[0] != a) { gaps[0] = a; changed++; }
    if (gaps[n - 1] != b) { gaps[n - 1] = b; changed++; }
    for (int i = 0; i < n; i++) { if (gaps[i] < a) { gaps[i] = a; changed++; } else if (gaps[i] > b) { gaps[i] = b; changed++; } }
    for (int i = 1; i < n; i++) if (gaps[i - 1] > gaps[i]) { gaps[i] = gaps[i - 1]; changed++; }
    if (changed == 0) return n;
    gaps[0] = a; gaps[n - 1] = b;
    int remove = 0;
    for (int i = 0; i < n; i += 2) {
        gaps[i] = gaps[i] < a ? a : (gaps[i] > b ? b : gaps[i]);
        gaps[i + 1] = gaps[i + 1] < a ? a : (gaps[i + 1] > b ? b : gaps[i + 1]);
        if (gaps[i] == gaps[i + 1]) remove++;
    }
    if (remove == 0) return n;
    /* fixGaps2: merge ranges closer than minGap, left to right */
    int dead[ORC_MAX_GAPS]; const int m = n / 2;
    for (int i = 0; i < m; i++) dead[i] = 0;
    for (int i = 1; i < m; i++) {
        if (!dead[i - 1]) {
            if (gaps[2 * i] - gaps[2 * i - 1] <= minGap) {
                gaps[2 * i] = imin(gaps[2 * i - 2], gaps[2 * i]);
                gaps[2 * i + 1] = imax(gaps[2 * i - 1], gaps[2 * i + 1]);
                dead[i - 1] = 1;
            }
        }
    }
    int k = 0;
    for (int i = 0; i < m; i++) if (!dead[i]) { gaps[2 * k] = gaps[2 * i]; gaps[2 * k + 1] = gaps[2 * i + 1]; k++; }
    if (k < 2) return 0;
    return 2 * k;
}
/* SiteScore.CHECKGAPS (stream/SiteScore.java:951-958) */
static int check_gaps(const orc_ss* s)
{
    if (s->ngaps == 0) return 1;
    if (s->ngaps & 1) return 0;
    for (int i = 1; i < s->ngaps; i++) if (s->gaps[i - 1] > s->gaps[i]) return 0;
    return s->gaps[0] == s->start && s->gaps[s->ngaps - 1] == s->stop;
}
/* SiteScore.setLimits / setStop (stream/SiteScore.java:905-914, 943-950); MINGAP = 256 (Shared.java:24) */
void orc_ss_set_limits(orc_ss* s, int a, int b)
{
    s->start = a; s->stop = b;
    if (s->ngaps > 0) { s->gaps[0] = a; s->gaps[s->ngaps - 1] = b; if (!check_gaps(s)) s->ngaps = orc_fix_gaps(s->start, s->stop, s->gaps, s->ngaps, 256); }
}
void orc_ss_set_stop(orc_ss* s, int b)
{
    s->stop = b;
    if (s->ngaps > 0) { s->gaps[s->ngaps - 1] = b; s->ngaps = orc_fix_gaps(s->start, s->stop, s->gaps, s->ngaps, 256); }
}
/* GapTools.calcGrefLen (:75-91) */
int orc_calc_gref_len(const orc_ss* s)
{
    int total = s->stop - s->start + 1;
    for (int i = 2; i < s->ngaps; i += 2) total -= imax(0, (s->gaps[i] - s->gaps[i - 1] - 128) / 128) * 127;
    return total;
}

/* SiteScore.setStart (stream/SiteScore.java:933-942) */
void orc_ss_set_start(orc_ss* s, int a)
{
    s->start = a;
    if (s->ngaps > 0) { s->gaps[0] = a; if (s->gaps[0] > s->gaps[1]) s->ngaps = orc_fix_gaps(a, s->stop, s->gaps, s->ngaps, 256); }
}

/* SiteScore.setSlowScore (stream/SiteScore.java:962-983): also moves pairedScore */
void orc_ss_set_slow_score(orc_ss* s, int x) {
    if (x <= 0) { s->paired_score = s->slow_score = x; }
    else if (s->paired_score <= 0) { s->slow_score = x; }
    else { if (s->slow_score > 0) s->paired_score = x + (s->paired_score - s->slow_score); else s->paired_score = x + 1; }
    s->slow_score = x;
}

static int max_quality(int len) { return 70 + (len - 1) * 100; }
static int max_imperfect(int len) { return max_quality(len) + imin(-472, -395 - 100); }

void orc_sitelist_trim(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int32_t* read_len, const orc_policy_cfg* cfg, orc_read_out* out)
{
    for (int64_t r = 0; r < nreads; r++) {
        orc_ss* v = lists + r * cap; int n = nss[r];
        out[r].near_perfect = 0; out[r].flags = 0; out[r].clearzone = 0; out[r].best_sites = 0;
        if (cfg->trim_list && n > 1) {
            if (cfg->min_trim_sites_to_retain > 1) stable_sort(v, n, ss_compare);
            out[r].best_sites = trim_list(v, &n, 0, max_quality(read_len[r]), 1, cfg->min_trim_sites_to_retain, cfg->max_trim_sites_to_retain);
        }
        nss[r] = n;
    }
}

/* SiteScore.setPerfect(bases) in full (sites of any length, possibly hanging over the array) */
static void ss_set_perfect(orc_ss* s, const int8_t* bases, int len, const int8_t* ref, int refLen)
{
    if (len != s->stop - s->start + 1) { s->perfect = 0; s->semiperfect = 0; return; }
    int perfect = 1, semiperfect = 1;
    int refloc = s->start, readloc = 0, N = 0;
    const int max = imin(s->stop, refLen - 1), nlimit = len / 2;
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

void orc_sitelist_noindel(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int64_t* read_off,
                          const int8_t* refs, const int64_t* chrom_off, const orc_policy_cfg* cfg, orc_read_out* out)
{
    for (int64_t r = 0; r < nreads; r++) {
        orc_ss* v = lists + r * cap; const int n = nss[r];
        const int len = (int)(read_off[r + 1] - read_off[r]);
        const int maxSw = max_quality(len), maxImp = max_imperfect(len);
        int numPerfect = 0, numNear = 0, best = (-2147483647 - 1), forceSlow = 0;
        for (int j = 0; j < n; j++) {
            orc_ss* ss = &v[j];
            const int oldScore = ss->score, sslen = ss->stop - ss->start + 1;
            const int8_t* bases = (ss->strand == 0 ? basesP : basesM) + read_off[r];
            const int8_t* ref = refs + chrom_off[ss->chrom - 1]; const int refLen = (int)(chrom_off[ss->chrom] - chrom_off[ss->chrom - 1]);
            int sni;
            if (ss->perfect) {
                numNear++;
                sni = maxSw; orc_ss_set_slow_score(ss, sni); ss->score = sni; ss->ngaps = 0;
            } else {
                sni = orc_score_no_indels(bases, len, ref, refLen, ss->start, 0);
                if (sni < oldScore && oldScore >= maxImp && sslen != len) {
                    const int s2 = orc_score_no_indels(bases, len, ref, refLen, ss->stop - len + 1, 0);
                    if (s2 >= maxImp) { sni = s2; ss->start = ss->stop - len + 1; ss_set_perfect(ss, bases, len, ref, refLen); }
                }
                orc_ss_set_slow_score(ss, sni); ss->score = sni;
                if (sni >= maxImp) {
                    numNear++;
                    ss->stop = ss->start + len - 1; ss->ngaps = 0;
                    if (sni >= maxSw) { numPerfect++; ss->perfect = ss->semiperfect = 1; }
                    else ss_set_perfect(ss, bases, len, ref, refLen);
                    if (cfg->quick_match_strings && !ss->perfect && (cfg->print_secondary || sni >= best)) ss->has_match = 1;
                } else if (oldScore >= maxImp) forceSlow = 1;
                else if (cfg->print_secondary) forceSlow = 1;
            }
            best = imax(ss->slow_score, best);
        }
        out[r].near_perfect = n == 0 ? 0 : (forceSlow ? -numNear : numNear);
        (void)numPerfect;
        stable_sort(v, n, ss_compare);           /* BBMapThread.java:442 */
    }
}

static int positional_match(const orc_ss* a, const orc_ss* b, int testGaps) {
    if (a->chrom != b->chrom || a->strand != b->strand || a->start != b->start || a->stop != b->stop) return 0;
    if (!testGaps || (a->ngaps == 0 && b->ngaps == 0)) return 1;
    if ((a->ngaps == 0) != (b->ngaps == 0)) return 0;
    if (a->ngaps != b->ngaps) return 0;
    for (int i = 0; i < a->ngaps; i++) if (a->gaps[i] != b->gaps[i]) return 0;
    return 1;
}
static int max3i(int a, int b, int c) { return imax(a, imax(b, c)); }

/* Tools.mergeDuplicateSites(list, true, true); returns the new size (the assertion of doAssertions is not restated) */
static int merge_duplicate_sites(orc_ss* v, int n) {
    if (n < 2) return n;
    stable_sort(v, n, ss_pcomp);
    char dead[4096]; memset(dead, 0, (size_t)n);
    int removed = 0, ai = 0;
    for (int i = 1; i < n; i++) {
        orc_ss* a = &v[ai]; orc_ss* b = &v[i];
        if (positional_match(a, b, 1)) {
            orc_ss_set_slow_score(a, imax(a->slow_score, b->slow_score));
            a->paired_score = (a->paired_score <= a->slow_score && b->paired_score <= a->slow_score) ? 0 : max3i(0, a->paired_score, b->paired_score);
            a->score = imax(a->score, b->score);
            a->perfect = (a->perfect || b->perfect); a->semiperfect = (a->semiperfect || b->semiperfect);
            removed++; dead[i] = 1;
        } else if (positional_match(a, b, 0)) {
            const orc_ss* better;
            if (a->score != b->score) better = (a->score > b->score ? a : b);
            else if (a->slow_score != b->slow_score) better = (a->slow_score > b->slow_score ? a : b);
            else if (a->paired_score != b->paired_score) better = (a->paired_score > b->paired_score ? a : b);
            else better = a;
            const int bg = better->ngaps; int g[ORC_MAX_GAPS]; memcpy(g, better->gaps, sizeof(better->gaps));
            orc_ss_set_slow_score(a, imax(a->slow_score, b->slow_score));
            a->paired_score = (a->paired_score <= a->slow_score && b->paired_score <= a->slow_score) ? 0 : max3i(0, a->paired_score, b->paired_score);
            a->score = imax(a->score, b->score);
            a->perfect = (a->perfect || b->perfect); a->semiperfect = (a->semiperfect || b->semiperfect);
            a->ngaps = bg; memcpy(a->gaps, g, sizeof(a->gaps));
            removed++; dead[i] = 1;
        } else ai = i;
    }
    if (removed > 0) { int k = 0; for (int i = 0; i < n; i++) if (!dead[i]) v[k++] = v[i]; n = k; }
    return n;
}
static int count_top_scores(const orc_ss* v, int n, int thresh) {
    if (n == 0) return 0;
    int count = 1; const int limit = v[0].score - thresh;
    for (int i = 1; i < n; i++) {
        if (v[i].score < limit) break;
        if (v[0].start != v[i].start && v[0].stop != v[i].stop) count++;
    }
    return count;
}
static int remove_low_quality_unpaired(orc_ss* v, int n, int thresh) {
    if (n == 0) return 0;
    if (v[0].score < thresh) return 0;
    for (int i = n - 1; i > 1; i--)
        if (v[i].slow_score < thresh) { for (int k = i; k + 1 < n; k++) v[k] = v[k + 1]; n--; }
    return n;
}

/* BBMapThread.processRead :478-553 */
void orc_sitelist_final(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int32_t* read_len, const orc_policy_cfg* cfg, orc_read_out* out)
{
    for (int64_t r = 0; r < nreads; r++) {
        orc_ss* v = lists + r * cap; int n = nss[r];
        const int maxSw = max_quality(read_len[r]);
        int flags = 0, clearzone = 0, numBest = 0;
        if (n > 0) { n = merge_duplicate_sites(v, n); stable_sort(v, n, ss_compare); }
        const int perfect = n > 0 && (v[0].slow_score == maxSw || v[0].perfect);           /* Read.setPerfectFlag, match == null */
        if (n > 1) {
            const int score = v[0].score;
            if (perfect) clearzone = cfg->clearzonep;
            else {
                const float cz1blimit = ((float)maxSw * cfg->cz1b_scale - cfg->cz1b_flat);
                const float cz1climit = ((float)maxSw * cfg->cz1c_scale - cfg->cz1c_flat);
                if ((float)score > cz1blimit)
                    clearzone = (int)(((float)((maxSw - score) * cfg->clearzone1b) + ((float)score - cz1blimit) * (float)cfg->clearzone1) / ((float)maxSw - cz1blimit));   /* int product first, as in Java */
                else if ((float)score > cz1climit)
                    clearzone = (int)(((cz1blimit - (float)score) * (float)cfg->clearzone1c + ((float)score - cz1climit) * (float)cfg->clearzone1b) / (cz1blimit - cz1climit));
                else clearzone = cfg->clearzone1c;
            }
            numBest = count_top_scores(v, n, clearzone);
            if (numBest > 1) flags |= 4;
            else {
                const int lim = (perfect ? (int)(4.f * (float)cfg->clearzone_limit1e) : score + cfg->clearzone1e >= maxSw ? 2 * cfg->clearzone_limit1e : cfg->clearzone_limit1e) + 1;
                if (n > lim && clearzone < cfg->clearzone1e) {
                    numBest = count_top_scores(v, n, cfg->clearzone1e);
                    if (numBest > lim) flags |= 4;
                }
            }
        }
        if (n > 0) {
            const int lim = (int)((float)maxSw * cfg->min_align_ratio);
            if (v[0].score < lim) n = 0;
            else n = remove_low_quality_unpaired(v, n, imin(lim, imax(1, lim - cfg->clearzone3)));
        }
        if (n > 0) flags |= 1;
        if (perfect && n > 0) flags |= 2;
        nss[r] = n;
        out[r].near_perfect = 0; out[r].flags = flags; out[r].clearzone = clearzone; out[r].best_sites = numBest;
    }
}

/* AbstractMapThread.findTipDeletions(Read r, basesP, basesM, maxSwScore, maxImperfectScore) (current/align2/AbstractMapThread.java:1073-1104) with the
 * quality gate of Read.min/avgQuality{First,Last}NBases (current/stream/Read.java:1760-1815; quality == NULL: the FASTA path).  out[r].best_sites
 * counts the sites that changed.  Sites with gaps would go through setStart/setStop -> GapTools.fixGaps: skipped and flagged (flags bit 3). */
void orc_sitelist_tipdel(orc_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int8_t* quality,
                         const int64_t* read_off, const int8_t* refs, const int64_t* chrom_off, const int32_t* chrom_min_index, const orc_tipdel_cfg* tc, orc_read_out* out)
{
    const int TIPLEN = tc->max_tiplen;
    for (int64_t r = 0; r < nreads; r++) {
        orc_ss* v = lists + r * cap; const int n = nss[r];
        const int len = (int)(read_off[r + 1] - read_off[r]);
        out[r].near_perfect = 0; out[r].flags = 0; out[r].clearzone = 0; out[r].best_sites = 0;
        if (len == 0) continue;
        int findRight = 1, findLeft = 1;
        if (quality) {
            const int8_t* q = quality + read_off[r];
            int minL = 0, avgL = 0, minF = 0, avgF = 0;
            if (TIPLEN <= len) {
                int x = 0; minL = q[len - TIPLEN];
                for (int i = len - TIPLEN; i < len; i++) { const int b = q[i]; x += (b < 0 ? 0 : b); if (b < minL) minL = b; }
                avgL = x / TIPLEN;
                x = 0; minF = q[0];
                for (int i = 0; i < TIPLEN; i++) { const int b = q[i]; x += (b < 0 ? 0 : b); if (i >= 1 && b < minF) minF = b; }
                avgF = x / TIPLEN;
            }
            findRight = (minL >= 6 && avgL >= 14); findLeft = (minF >= 6 && avgF >= 14);
        }
        if (!findRight && !findLeft) continue;
        const int maxSw = max_quality(len), maxImp = max_imperfect(len);
        for (int j = 0; j < n; j++) {
            orc_ss* ss = &v[j];
            if (ss->semiperfect || ss->slow_score >= maxImp) continue;
            const int8_t* bases = (ss->strand == 0 ? basesP : basesM) + read_off[r];
            const int8_t* ref = refs + chrom_off[ss->chrom - 1]; const int refLen = (int)(chrom_off[ss->chrom] - chrom_off[ss->chrom - 1]);
            const int minIndex = chrom_min_index ? chrom_min_index[ss->chrom - 1] : 0;
            /* findTipDeletions(ss, bases, maxImperfectScore, lookRight, lookLeft) :1107-1141 */
            int changed = 0;
            if (len > 2 * TIPLEN) {
                int maxSearch = tc->search_range;
                maxSearch = imin(maxSearch, tc->align_columns - (tc->slow_rescue_padding + 8 + imax(len, ss->stop - ss->start)));
                if (maxSearch >= 1) {
                    int go = 1;
                    if (findRight) {
                        const int x = orc_find_tip_deletions_right(bases, len, ref, refLen, minIndex, ss->stop, maxSearch, TIPLEN);
                        if (x > 0) {
                            orc_ss_set_stop(ss, ss->stop + x); changed = 1;
                            maxSearch = imin(maxSearch, tc->align_columns - (tc->slow_rescue_padding + 8 + imax(len, ss->stop - ss->start)));
                            if (maxSearch < 1) go = 0;
                        }
                    }
                    if (go && findLeft) {
                        const int y = orc_find_tip_deletions_left(bases, len, ref, refLen, minIndex, ss->start, maxSearch, TIPLEN);
                        if (y > 0) { orc_ss_set_start(ss, ss->start - y); changed = 1; }
                    }
                }
            }
            if (changed) {
                out[r].best_sites++;
                ss->has_match = 0;
                orc_ss_set_slow_score(ss, orc_score_no_indels(bases, len, ref, refLen, ss->start, 0));
                if (ss->slow_score == maxSw) { orc_ss_set_stop(ss, ss->start + len - 1); ss->perfect = ss->semiperfect = 1; }
                else { ss->perfect = 0; ss_set_perfect(ss, bases, len, ref, refLen); }
            }
        }
    }
}

/* AbstractMapThread.removeOutOfBounds(r, ..., SAM_OUT, EXPECTED_LEN_LIMIT) (current/align2/AbstractMapThread.java:2444-2479), called by quickMap right
 * after the index search (:739): sites hanging over the chromosome array go, with SAM output so do sites that span two scaffolds
 * (Data.isSingleScaffold, current/dna/Data.java:1112-1140), and over-long sites are cut to read length + 40.  scaf_off == NULL: no scaffold table
 * (Data.scaffoldLocs == null).  A gapped site that needs cutting would go through fixGaps: left alone and flagged (flags bit 3). */
static int bsearch_java(const int32_t* a, int n, int key) {
    int lo = 0, hi = n - 1;
    while (lo <= hi) { const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1); const int v = a[mid]; if (v < key) lo = mid + 1; else if (v > key) hi = mid - 1; else return mid; }
    return -(lo + 1);
}
static int is_single_scaffold(const int32_t* loc, int n, int pad, int loc1, int loc2) {
    if (n < 2) return 1;
    const int idx = bsearch_java(loc, n, loc1 + pad);
    const int scaf = idx >= 0 ? idx : imax(0, (-1 - idx) - 1);
    if (scaf == n - 1) return 1;
    const int lowerBound = loc[scaf] - pad, upperBound = loc[scaf + 1];
    if (loc2 < lowerBound || loc1 > upperBound) return 0;
    return loc2 < upperBound;
}
void orc_sitelist_bounds(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int32_t* read_len, const int32_t* chrom_max_index,
                         const int32_t* scaf_off, const int32_t* scaf_loc, int32_t inter_scaffold_padding, int32_t sam_out, int32_t expected_len_limit,
                         orc_read_out* out)
{
    for (int64_t r = 0; r < nreads; r++) {
        orc_ss* v = lists + r * cap; int n = nss[r];
        const int initial = n; int flags = 0;
        for (int i = 0; i < n; i++) {
            orc_ss* ss = &v[i];
            int removed = 0;
            if (ss->start < 0 || ss->stop > chrom_max_index[ss->chrom - 1]) removed = 1;
            else if (sam_out && scaf_off) {
                const int base = scaf_off[ss->chrom - 1], cnt = scaf_off[ss->chrom] - base;
                if (!is_single_scaffold(scaf_loc + base, cnt, inter_scaffold_padding, ss->start, ss->stop)) removed = 1;
            }
            if (removed) { for (int k = i; k + 1 < n; k++) v[k] = v[k + 1]; n--; i--; continue; }
            if (orc_calc_gref_len(ss) >= expected_len_limit) {
                orc_ss_set_stop(ss, ss->start + imin(read_len[r] + 40, expected_len_limit));
                if (ss->ngaps > 0) ss->ngaps = orc_fix_gaps(ss->start, ss->stop, ss->gaps, ss->ngaps, 256);
            }
        }
        nss[r] = n;
        out[r].near_perfect = 0; out[r].flags = flags; out[r].clearzone = 0; out[r].best_sites = initial - n;
    }
}

/* ---- what processRead does to the list after the match string of the primary site exists (unpaired reads) ----
 *   the clearzone-3 block and the final score gate          current/align2/BBMapThread.java:667-684, 698-700
 *   AbstractMapThread.applyClearzone3 / calcCZ3_fraction    current/align2/AbstractMapThread.java:1820-1870, 1893-1911 (CZ3_MULTS :2809)
 *   AbstractMapThread.calcTipScorePenalty / applyScorePenalty   :2499-2567, 2601-2609 (called with tiplen 7, BBMapThread.java:706-709)
 * r.mapScore is the top site's slowScore (Read.setFromSite, current/stream/Read.java:1178; the reference asserts the equality, :1826).
 * All float arithmetic is Java's: single precision, one rounding per operation. */
static const float CZ3_MULTS[7] = { 0.f, 1.f, .75f, 0.5f, 0.25f, 0.125f, 0.0625f };

static float calc_cz3_fraction(int score1, int score2, int cz3, float inv)
{
    const int dif = score1 - score2;
    if (dif >= cz3) return 0.f;
    const int dif2 = cz3 - dif;
    const float f = (float)dif2 * inv;
    const float f2 = f * f;
    const float a = 2.f * f2;
    return (f + a) + a * f;                     /* f+2f*f2+2f*f2*f, left to right */
}

/* returns the amount subtracted (0 = applyClearzone3 returned false) */
static int apply_clearzone3(orc_ss* v, int n, int len, int flags, int cz3, float inv)
{
    if (!(flags & 1) || (flags & 4) || n < 2) return 0;
    const int score1 = v[0].slow_score, mapScore = v[0].slow_score;
    float sub = 0.f;
    const int max = imin(7, n);
    for (int i = 1; i < max; i++) {
        if (i > 2 && v[i].slow_score < v[i - 1].slow_score) break;
        const float f = calc_cz3_fraction(score1, v[i].slow_score, cz3, inv);
        if (f <= 0.f) break;
        sub += f * CZ3_MULTS[i];
    }
    if (sub <= 0.f) return 0;
    const float asymptote = 4.f + 0.03f * (float)len;
    sub = sub * 1.8f;
    const float sub2 = (float)cz3 * ((asymptote * sub) / (sub + asymptote));
    int subi = (int)(sub2 + 0.5f);
    if (subi >= mapScore - 300) subi = mapScore - 300;
    if (subi <= 0) return 0;
    for (int i = 0; i < n; i++) { orc_ss_set_slow_score(&v[i], v[i].slow_score - subi); v[i].score -= subi; }
    return subi;
}

/* out[r] is in/out: flags as the final policy left them; afterwards flags updated, near_perfect = r.mapScore, best_sites = the amount subtracted */
void orc_sitelist_clearzone3(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int32_t* read_len, const orc_policy_cfg* cfg,
                             int32_t ambiguous_toss, orc_read_out* out)
{
    for (int64_t r = 0; r < nreads; r++) {
        orc_ss* v = lists + r * cap; int n = nss[r];
        const int maxSw = max_quality(read_len[r]);
        int flags = out[r].flags, subi = 0;
        /* removeDuplicateBestSites (AbstractMapThread.java:1328-1349; processRead :624-628): copies of the top site at the end of the list */
        while (n > 1 && v[0].chrom == v[n - 1].chrom && v[0].strand == v[n - 1].strand && v[0].start == v[n - 1].start && v[0].stop == v[n - 1].stop) n--;
        if (n == 0) flags &= ~1;
        int mapScore = n > 0 ? v[0].slow_score : 0;
        if ((cfg->clearzone3 > cfg->clearzone1 || cfg->clearzone3 > cfg->clearzonep) && n > 0 && !(flags & 4) && mapScore > 0) {
            const float q = (float)maxSw / (float)mapScore;
            const float cz3v2 = (float)cfg->clearzone3 * (1.25f < q ? 1.25f : q);
            subi = apply_clearzone3(v, n, read_len[r], flags, (int)cz3v2, 1.f / cz3v2);
            if (subi > 0) {
                mapScore -= subi;
                const int minScore = (int)((float)maxSw * cfg->min_align_ratio);
                if (mapScore < minScore) flags |= 4;
            }
        }
        if ((flags & 4) && ambiguous_toss) { n = 0; flags &= ~1; mapScore = 0; }
        if (n == 0 || (!(flags & 4) && (float)mapScore < (float)maxSw * cfg->min_align_ratio)) { n = 0; flags &= ~1; mapScore = 0; }   /* r.clearMapping() */
        nss[r] = n;
        out[r].flags = flags; out[r].near_perfect = mapScore; out[r].best_sites = subi;
    }
}

/* calcTipScorePenalty(r, maxScore, tiplen) on a long-format match string; status bit0: the string ended before tiplen+1 read positions
 * were seen (the reference would throw), bit1: short-format digits (Read.toLongMatchString is not restated). */
static int calc_tip_score_penalty(const int8_t* match, int mlen, const int8_t* bases, int len, int mapped, int mapScore, int maxScore, int tiplen, int* status)
{
    if (!mapped || mlen <= 0 || len < 2 * tiplen) return 0;
    int points = 0;
    int8_t prev = 'm';
    for (int i = 0, cpos = 0; cpos <= tiplen; i++) {
        if (i >= mlen) { *status |= 1; return 0; }
        const int8_t b = match[i];
        if (b == 'm') cpos++;
        else if (b == 'D') { if (prev != 'D') points += 2 * (tiplen + 2 - cpos); }
        else if (b == 'N' || b == 'C') { points += (tiplen + 2 - cpos); cpos++; }
        else {
            if (b >= '0' && b <= '9') { *status |= 2; return 0; }
            points += 2 * (tiplen + 2 - cpos); cpos++;
        }
        prev = b;
    }
    prev = 'm';
    for (int i = mlen - 1, cpos = 0; cpos <= tiplen; i--) {
        if (i < 0) { *status |= 1; return 0; }
        const int8_t b = match[i];
        if (b == 'm') cpos++;
        else if (b == 'D') { if (prev != 'D') points += 2 * (tiplen + 2 - cpos); }
        else if (b == 'N' || b == 'C') { points += (tiplen + 2 - cpos); cpos++; }
        else { points += 2 * (tiplen + 2 - cpos); cpos++; }
        prev = b;
    }
    const int last = len - 1;
    int8_t b = bases[0];
    if (b != 'N' && b == bases[1]) for (int i = 2; i <= tiplen && bases[i] == b; i++) points++;
    b = bases[last];
    if (b != 'N' && b == bases[last - 1]) for (int i = last - 2; i >= (last - tiplen) && bases[i] == b; i--) points++;
    if (points < 1) return 0;
    const float asymptote = 80.f;
    const float f = (asymptote * (float)points) / ((float)points + asymptote);
    const int penalty = (int)((f * .0022f) * (float)maxScore);
    const int maxPenalty = mapScore - maxScore / 10;
    if (maxPenalty <= 0) return 0;
    return imin(penalty, maxPenalty);
}

/* PENALIZE_AMBIG block of processRead: penalty[r] = calcTipScorePenalty(r, maxSwScore, tiplen), then applyScorePenalty.  flags[r] bit0 = r.mapped();
 * match of read r = match[match_off[r] .. match_off[r+1]) (empty = r.match == null); bases = r.bases (the read as sequenced). */
void orc_sitelist_tip_penalty(orc_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int64_t* read_off, const int8_t* bases,
                              const int8_t* match, const int64_t* match_off, const orc_read_out* flags, int32_t tiplen, int32_t* penalty, int32_t* status)
{
    for (int64_t r = 0; r < nreads; r++) {
        orc_ss* v = lists + r * cap; const int n = nss[r];
        const int len = (int)(read_off[r + 1] - read_off[r]);
        int st = 0;
        const int p = calc_tip_score_penalty(match + match_off[r], (int)(match_off[r + 1] - match_off[r]), bases + read_off[r], len,
                                             (flags[r].flags & 1) && n > 0, n > 0 ? v[0].slow_score : 0, max_quality(len), tiplen, &st);
        if (p > 0) for (int i = 0; i < n; i++) { orc_ss_set_slow_score(&v[i], v[i].slow_score - p); v[i].score -= p; }
        penalty[r] = p; status[r] = st;
    }
}

/* ---- list primitives for the paired chain (mapper_oracle.c) ---- */
void orc_sl_sort(orc_ss* v, int n, int positional) { stable_sort(v, n, positional ? ss_pcomp : ss_compare); }
int orc_sl_trim_below_cutoff(orc_ss* v, int n, int cutoff, int retainPaired, int minS, int maxS) { return trim_below_cutoff(v, n, cutoff, retainPaired, minS, maxS); }
int orc_sl_trim_list(orc_ss* v, int* n, int retainPaired, int maxScore, int specialCasePerfect, int minS, int maxS) { return trim_list(v, n, retainPaired, maxScore, specialCasePerfect, minS, maxS); }
int orc_sl_merge_duplicates(orc_ss* v, int n) { return merge_duplicate_sites(v, n); }
int orc_sl_count_top_scores(const orc_ss* v, int n, int thresh) { return count_top_scores(v, n, thresh); }
void orc_sl_set_perfect(orc_ss* s, const int8_t* bases, int len, const int8_t* ref, int refLen) { ss_set_perfect(s, bases, len, ref, refLen); }
