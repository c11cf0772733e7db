/* TEST INFRASTRUCTURE ONLY — see rescue_oracle.h.
 * Restates, loop for loop (sequential early exits included):
 *   AbstractMapThread.findTipDeletions(SiteScore,...)   current/align2/AbstractMapThread.java:1107-1141
 *   AbstractMapThread.findTipDeletionsRight / Left       :2178-2235, :2238-2294
 *   AbstractMapThread.quickRescue                        :2303-2405
 *   SiteScore.setPerfect(bases) for a site as long as the read   current/stream/SiteScore.java:239-291
 *   SiteScore.isInBounds                                 :425-428
 */
#include "rescue_oracle.h"
#include <limits.h>

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

int orc_find_tip_deletions_right(const int8_t* bases, int len, const int8_t* ref, int refLen, int minIndex, int originalStop, int searchDist, int tiplen)
{
    if (originalStop < minIndex + tiplen - 1) return 0;
    if (originalStop >= refLen) return 0;                 /* Java would throw ArrayIndexOutOfBounds; sites passed here are in bounds */
    int minMismatches = tiplen, bestStart = originalStop;
    const int tipCoord = len - 1;
    int lastMismatch = 0, originalMismatches = 0, contig = 0;
    for (int i = 0; i < tiplen && contig < 5; i++) {
        if (bases[tipCoord - i] != ref[originalStop - i]) { originalMismatches++; lastMismatch = i; contig = 0; }
        else contig++;
    }
    if (originalMismatches < 3) return 0;
    minMismatches = originalMismatches;
    tiplen = lastMismatch + 1;
    if (tiplen < 4) return 0;
    searchDist = imin(searchDist, 30 * originalMismatches);
    const int lastIndexToStart = imin(refLen - 1, originalStop + searchDist);
    for (int start = originalStop + 1; start <= lastIndexToStart && minMismatches > 0; start++) {
        int mismatches = 0;
        for (int j = 0; j < tiplen && mismatches < minMismatches; j++)
            if (bases[tipCoord - j] != ref[start - j]) mismatches++;
        if (mismatches < minMismatches) { bestStart = start; minMismatches = mismatches; }
    }
    if (minMismatches > 2 || originalMismatches - minMismatches < 2) return 0;
    return bestStart - originalStop;
}

int orc_find_tip_deletions_left(const int8_t* bases, int len, const int8_t* ref, int refLen, int minIndex, int originalStart, int searchDist, int tiplen)
{
    (void)len;
    if (originalStart + tiplen >= refLen) return 0;
    if (minIndex >= originalStart) return 0;
    int minMismatches = tiplen, bestStart = originalStart;
    int lastMismatch = 0, originalMismatches = 0, contig = 0;
    for (int i = 0; i < tiplen && contig < 5; i++) {
        if (bases[i] != ref[originalStart + i]) { originalMismatches++; lastMismatch = i; contig = 0; }
        else contig++;
    }
    if (originalMismatches < 3) return 0;
    minMismatches = originalMismatches;
    tiplen = lastMismatch + 1;
    if (tiplen < 4) return 0;
    searchDist = imin(searchDist, 16 + 16 * originalMismatches + 8 * tiplen);
    const int lastIndexToStart = imax(minIndex, originalStart - searchDist);
    for (int start = originalStart - 1; start >= lastIndexToStart && minMismatches > 0; start--) {
        int mismatches = 0;
        for (int j = 0; j < tiplen && mismatches < minMismatches; j++)
            if (bases[j] != ref[start + j]) mismatches++;
        if (mismatches < minMismatches) { bestStart = start; minMismatches = mismatches; }
    }
    if (minMismatches > 2 || originalMismatches - minMismatches < 2) return 0;
    return originalStart - bestStart;
}

/* findTipDeletions(SiteScore ss, bases, maxImperfectScore, lookRight, lookLeft) */
void orc_tipdel_batch(const int8_t* reads, const int8_t* refs, const orc_tipdel_task* tasks, int64_t n, const orc_tipdel_cfg* cfg, orc_tipdel_out* outs)
{
    for (int64_t t = 0; t < n; ++t) {
        const orc_tipdel_task* T = &tasks[t];
        orc_tipdel_out* o = &outs[t];
        const int8_t* bases = reads + T->read_off; const int8_t* ref = refs + T->ref_off;
        const int len = T->read_len;
        int start = T->start, stop = T->stop;
        o->start = start; o->stop = stop; o->right = 0; o->left = 0;
        if (T->slow_score >= T->max_imperfect) continue;
        if (len <= 2 * cfg->max_tiplen) continue;
        int maxSearch = cfg->search_range;
        maxSearch = imin(maxSearch, cfg->align_columns - (cfg->slow_rescue_padding + 8 + imax(len, stop - start)));
        if (maxSearch < 1) continue;
        if (T->flags & 1) {
            const int x = orc_find_tip_deletions_right(bases, len, ref, T->ref_len, T->min_index, stop, maxSearch, cfg->max_tiplen);
            if (x > 0) {
                stop += x; o->stop = stop; o->right = x;
                maxSearch = imin(maxSearch, cfg->align_columns - (cfg->slow_rescue_padding + 8 + imax(len, stop - start)));
                if (maxSearch < 1) continue;
            }
        }
        if (T->flags & 2) {
            const int y = orc_find_tip_deletions_left(bases, len, ref, T->ref_len, T->min_index, start, maxSearch, cfg->max_tiplen);
            if (y > 0) { start -= y; o->start = start; o->left = y; }
        }
    }
}

static int absdif(int a, int b) { return a > b ? a - b : b - a; }

void orc_rescue_batch(const int8_t* reads, const int8_t* refs, const orc_rescue_task* tasks, int64_t n, const orc_rescue_cfg* cfg, orc_rescue_out* outs)
{
    for (int64_t t = 0; t < n; ++t) {
        const orc_rescue_task* T = &tasks[t];
        orc_rescue_out* o = &outs[t];
        const int8_t* bases = reads + T->read_off; const int8_t* ref = refs + T->ref_off;
        const int len = T->read_len, refLen = T->ref_len;
        o->start = -1; o->stop = -1; o->mismatches = 0; o->max_contig = 0; o->score = 0; o->perfect = 0; o->in_bounds = 0; o->pad_ = 0;
        if (len < 10) continue;
        const int searchRight = T->flags & 1, idealStart = T->ideal_start;
        int lowerBound, upperBound;
        if (searchRight) { lowerBound = imax(T->min_index, T->loc); upperBound = imin(refLen - len, T->loc + T->search_dist); }
        else { lowerBound = imax(T->min_index, T->loc - T->search_dist); upperBound = imin(refLen - len, T->loc); }
        int minMismatches = T->max_mismatches + 1;
        int maxContigMatches = 0, bestScore = 0, bestStart = -1, bestAbsdif = INT_MAX;
        if (searchRight) {
            for (int start = lowerBound; start <= upperBound; start++) {
                int mismatches = 0, contig = 0, currentContig = 0;
                for (int j = 0; j < len && mismatches <= minMismatches; j++) {
                    const int8_t c = bases[j], r = ref[start + j];
                    if (c != r || c == 'N') { mismatches++; contig = imax(contig, currentContig); currentContig = 0; }
                    else currentContig++;
                }
                const int score = (len - mismatches) + contig, ad = absdif(start, idealStart);
                if (mismatches <= minMismatches && (score > bestScore || (score == bestScore && ad < bestAbsdif))) {
                    bestStart = start; minMismatches = mismatches; maxContigMatches = contig; bestScore = score; bestAbsdif = ad;
                    if (mismatches == 0) upperBound = imin(upperBound, idealStart + ad);
                }
            }
        } else {
            for (int start = upperBound; start >= lowerBound; start--) {
                int mismatches = 0, contig = 0, currentContig = 0;
                for (int j = 0; j < len && mismatches <= minMismatches; j++) {
                    const int8_t c = bases[j], r = ref[start + j];
                    if (c != r || c == 'N') { mismatches++; contig = imax(contig, currentContig); currentContig = 0; }
                    else currentContig++;
                }
                const int score = (len - mismatches) + contig, ad = absdif(start, idealStart);
                if (mismatches <= minMismatches && (score > bestScore || (score == bestScore && ad < bestAbsdif))) {
                    bestStart = start; minMismatches = mismatches; maxContigMatches = contig; bestScore = score; bestAbsdif = ad;
                    if (mismatches == 0) lowerBound = imax(lowerBound, idealStart - ad);
                }
            }
        }
        if (bestStart < 0) continue;
        o->start = bestStart; o->stop = bestStart + len - 1; o->mismatches = minMismatches; o->max_contig = maxContigMatches;
        o->score = cfg->use_affine ? cfg->points_match + cfg->points_match2 * (len - 1 - minMismatches)
                                   : maxContigMatches + cfg->base_hit_score * (len - minMismatches);
        /* SiteScore.setPerfect(bases): the site is exactly as long as the read and starts at >= minIndex >= 0 */
        {
            int perfect = 1, semiperfect = 1, N = 0;
            const int stop = o->stop, nlimit = len / 2, max = imin(stop, refLen - 1);
            if (stop >= refLen) { N += stop - refLen + 1; perfect = 0; }
            if (N > nlimit) { perfect = 0; semiperfect = 0; }
            else {
                int refloc = bestStart, readloc = 0, bail = 0;
                for (; refloc <= max; refloc++, readloc++) {
                    const int8_t c = bases[readloc], r = ref[refloc];
                    if (c != r || c == 'N') {
                        perfect = 0;
                        if (c == 'N') semiperfect = 0;
                        if (r != 'N' || (N = N + 1) > nlimit) { semiperfect = 0; bail = 1; break; }
                    }
                }
                if (bail) { perfect = 0; }          /* returns semiperfect(false) with perfect already false */
                else { semiperfect = (semiperfect && N <= nlimit); perfect = (perfect && semiperfect && N == 0); }
            }
            o->perfect = (perfect ? 1 : 0) | (semiperfect ? 2 : 0);
        }
        o->in_bounds = (o->start >= 0 && o->stop <= T->max_index) ? 1 : 0;
    }
}
