/* TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of what BBMapThread.processRead does once the site list is final
 * (current/align2/BBMapThread.java:557-709): the primary site's match string
 *   genMatchString                  current/align2/AbstractMapThread.java:860-966
 *   genMatchStringForSite           :968-1068
 *   TranslateColorspaceRead.realign_new   current/align2/TranslateColorspaceRead.java:229-660
 *   SiteScore.fixXY / clipTipIndels / clipLeftTipIndel / clipRightTipIndel / unclip / leftPaddingNeeded / rightPaddingNeeded /
 *   fixLimitsXY / setPerfectFlag / isPerfect / isSemiPerfect        current/stream/SiteScore.java:175-236, 431-840, 916-931
 *   MSA.score(match) / toLocalAlignment      current/align2/MSA.java:216-470, 488-560;  calcDelScore / calcInsScore  …JNI.java:1347-1421
 * and the rest of the tail: removeDuplicateBestSites, the mapScore<=0 gate, applyClearzone3, AMBIGUOUS_TOSS, toLocalAlignment for
 * X/Y/C tips, the final ratio gate and the tip-score penalty (BBMapThread.java:624-709).
 * One read after the other, one site after the other, calling the oracle's sequential MultiStateAligner11ts (msa_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs may call this.  Parity UNPINNED against Java (no JVM). */
#include <stdlib.h>
#include <string.h>
#include "msa_oracle.h"
#include "sitelist_oracle.h"
#include "host_oracle.h"
#include "mapper_oracle.h"

static int imax2(int a, int b) { return a > b ? a : b; }
static int imin2(int a, int b) { return a < b ? a : b; }

#define MINGAP 256
#define MAXCOLS 3000

typedef struct {
    orc_ss s;
    int8_t* match; int mlen;          /* ss.match (NULL = null) */
    int serial;                       /* object identity of the SiteScore (topSite()!=top tests) */
} msite;

typedef struct {
    orc_msa* msa; const orc_map_cfg* cfg;
    const int8_t* refs; const int64_t* chrom_off;
    int64_t fills; int status;
} mctx;

static const int8_t* chrom_ptr(const mctx* C, int chrom, int* refLen) { *refLen = (int)(C->chrom_off[chrom] - C->chrom_off[chrom - 1]); return C->refs + C->chrom_off[chrom - 1]; }
static int8_t ca_get(const int8_t* ref, int refLen, int loc) { return (loc < 0 || loc >= refLen - 1) ? (int8_t)'N' : ref[loc]; }   /* ChromosomeArray.get: loc>=maxIndex -> 'N' */
static int fully_defined(int c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }     /* AminoAcid.isFullyDefined on upper-case input */

static void set_match(msite* m, const int8_t* src, int n) {
    if (m->match) free(m->match);
    m->match = (int8_t*)malloc((size_t)(n > 0 ? n : 1)); m->mlen = n;
    if (n > 0 && src) memcpy(m->match, src, (size_t)n);
}

/* Read.calcMatchLength on a long-format string (no digits): reference length */
static int match_ref_length(const int8_t* m, int n) { int len = 0; for (int i = 0; i < n; i++) if (m[i] != 'I') len++; return len; }
static int lengths_agree(const msite* m) { return m->match == NULL ? 1 : match_ref_length(m->match, m->mlen) == m->s.stop - m->s.start + 1; }
static int match_contains_xy(const msite* m) {
    if (!m->match || m->mlen < 1) return 0;
    const int8_t a = m->match[0], b = m->match[m->mlen - 1];
    return a == 'X' || a == 'Y' || b == 'X' || b == 'Y';
}

/* MSA.calcDelScore(len, true) / calcInsScore / calcSubScore / calcMatchScore for the 11ts constants */
static int calc_del_score(int len) {
    if (len <= 0) return 0;
    int score = -472;
    if (len > MINGAP) { const int rem = len % 128, div = (len - 128) / 128; score += div * (-2); len = rem + 128; }
    if (len > 80) { score += ((len - 80 + 3) / 4) * (-1); len = 80; }
    if (len > 20) { score += (len - 20) * (-1); len = 20; }
    if (len > 5) { score += (len - 5) * (-9); len = 5; }
    if (len > 1) score += (len - 1) * (-33);
    return score;
}
static int calc_ins_score(int len) {        /* POINTS_INS_ARRAY_C[len]: -395, then -39 x4, -23 x15, -8 ... clamped at MIN_SCORE */
    if (len <= 0) return 0;
    int s = 0;
    for (int i = 1; i <= len; i++) { const int p = i == 1 ? -395 : (i < 6 ? -39 : (i < 21 ? -23 : -8)); s = imax2(-1046575, p + s); }
    return s;
}
static int calc_sub_score(int len) { int score = -127; if (len > 5) { score += (len - 5) * (-25); len = 5; } if (len > 1) score += (len - 1) * (-51); return score; }
static int calc_match_score(int len) { return 70 + (len - 1) * 100; }

static int run_points(int8_t mode, int current, int8_t prevMode, int prevStreak) {
    if (mode == 'm') return calc_match_score(current);
    if (mode == 'S') { int s = calc_sub_score(current); if (prevMode == 'N' || prevMode == 'R') s += -51 - (-127); else if (prevMode == 'm' && prevStreak < 2) s += -147 - (-127); return s; }
    if (mode == 'D') return calc_del_score(current);
    if (mode == 'I' || mode == 'X' || mode == 'Y') return calc_ins_score(current);
    return 0;      /* C, N, R */
}
/* MSA.score(byte[] match) (MSA.java:488-560) */
int orc_score_match(const int8_t* match, int n) {
    if (!match || n < 1) return 0;
    int8_t mode = match[0], prevMode = '0'; int current = 0, prevStreak = 0, score = 0;
    for (int i = 0; i < n; i++) {
        const int8_t c = match[i];
        if (mode == c) current++;
        else { score += run_points(mode, current, prevMode, prevStreak); prevMode = mode; prevStreak = current; mode = c; current = 1; }
    }
    if (current > 0) score += run_points(mode, current, prevMode, prevStreak);
    return score;
}

/* SiteScore.setPerfect(bases) — same statement as scoreslow_oracle.c */
static void set_perfect(orc_ss* s, const int8_t* bases, int len, const int8_t* ref, int refLen) {
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
static int is_perfect(const orc_ss* s, const int8_t* bases, int len, const int8_t* ref, int refLen) {
    if (len != s->stop - s->start + 1 || s->start < 0) return 0;
    if (s->stop >= refLen) return 0;
    for (int i = 0; i < len; i++) { const int8_t c = bases[i], r = ref[s->start + i]; if (c != r || c == 'N') return 0; }
    return 1;
}
static int is_semiperfect(const orc_ss* s, const int8_t* bases, int len, const int8_t* ref, int refLen) {
    if (len != s->stop - s->start + 1) return 0;
    int readStart = 0, readStop = len, maxNoref = len / 2;
    const int refStop = s->start + len;
    if (s->start < 0) readStart = -s->start;
    if (refStop > refLen) readStop -= (refStop - refLen);
    for (int i = readStart; i < readStop; i++) {
        const int8_t c = bases[i], r = ref[s->start + i];
        if (c == 'N') return 0;
        if (c != r) { maxNoref--; if (maxNoref < 0 || r != 'N') return 0; }
    }
    return 1;
}

/* leftPaddingNeeded / rightPaddingNeeded (SiteScore.java:447-491) — literal, including the right-hand loop's `mloc>=tiplen` test */
static int left_padding_needed(const msite* m, int tiplen, int maxIndel) {
    if (!m->match || m->mlen < 1) return 0;
    int insertion = 0, xy = 0;
    for (int mloc = 0; mloc < m->mlen; mloc++) {
        const int8_t c = m->match[mloc];
        if (c == 'I') insertion++;
        else if (c == 'X' || c == 'Y') xy++;
        else if (c == 'D') return insertion + xy;
        else { if (mloc >= tiplen) break; }
    }
    if (insertion > maxIndel || xy > 0 || m->match[0] == 'I') return insertion + xy;
    return 0;
}
static int right_padding_needed(const msite* m, int tiplen, int maxIndel) {
    if (!m->match || m->mlen < 1) return 0;
    const int lastIndex = m->mlen - 1;
    int insertion = 0, xy = 0;
    for (int mloc = lastIndex; mloc >= 0; mloc--) {
        const int8_t c = m->match[mloc];
        if (c == 'I') insertion++;
        else if (c == 'X' || c == 'Y') xy++;
        else if (c == 'D') return insertion + xy;
        else { if (mloc >= tiplen) break; }
    }
    if (insertion > maxIndel || xy > 0 || m->match[lastIndex] == 'I') return insertion + xy;
    return 0;
}

static void fix_gaps_ss(orc_ss* s) { if (s->ngaps > 0) s->ngaps = orc_fix_gaps(s->start, s->stop, s->gaps, s->ngaps, MINGAP); }

static int clip_left_tip_indel(msite* m, int tiplen, int maxIndel) {
    if (!m->match || m->mlen < maxIndel) return 0;
    int8_t* match = m->match;
    if (match[0] == 'C' || match[0] == 'Y' || match[0] == 'X') return 0;
    int neutral = 0, insertion = 0, deletion = 0;
    {
        int mloc = 0;
        for (; mloc < m->mlen; mloc++) {
            const int8_t c = match[mloc];
            if (c == 'I') insertion++;
            else if (c == 'D') deletion++;
            else { neutral++; if (mloc >= tiplen) break; }
        }
        if (mloc >= m->mlen) mloc = m->mlen - 1;        /* the Java would index past the end here (all-indel strings do not occur) */
        while (mloc >= 0 && match[mloc] == 'm') { mloc--; neutral--; }
    }
    if (insertion <= maxIndel && deletion <= 4 * maxIndel) return 0;
    int sum = neutral + insertion + deletion;
    if (deletion > 0) {
        int i = 0, j = 0;
        for (; i < sum; i++) { if (match[i] != 'D') { match[j] = match[i]; j++; } }
        for (; i < m->mlen; i++, j++) match[j] = match[i];
        m->mlen = j;
    }
    sum = neutral + insertion;
    for (int i = 0; i < sum; i++) match[i] = 'C';
    orc_ss_set_start(&m->s, m->s.start - (insertion - deletion));
    return 1;
}
static int clip_right_tip_indel(msite* m, int tiplen, int maxIndel) {
    if (!m->match || m->mlen < maxIndel) return 0;
    int8_t* match = m->match;
    const int lastIndex = m->mlen - 1;
    if (match[lastIndex] == 'C' || match[lastIndex] == 'Y' || match[lastIndex] == 'X') return 0;
    int neutral = 0, insertion = 0, deletion = 0;
    {
        int mloc = lastIndex;
        for (const int min = lastIndex - tiplen; mloc >= 0; mloc--) {
            const int8_t c = match[mloc];
            if (c == 'I') insertion++;
            else if (c == 'D') deletion++;
            else { neutral++; if (mloc <= min) break; }
        }
        if (mloc < 0) mloc = 0;
        while (mloc < m->mlen && match[mloc] == 'm') { mloc++; neutral--; }
    }
    if (insertion <= maxIndel && deletion <= 4 * maxIndel) return 0;
    int sum = neutral + insertion + deletion;
    const int limit = m->mlen - sum;
    int newlen = m->mlen;
    if (deletion > 0) {
        int i = limit, j = limit;
        for (; i < m->mlen; i++) { if (match[i] != 'D') { match[j] = match[i]; j++; } }
        newlen = j;
    }
    m->mlen = newlen;
    for (int i = limit; i < m->mlen; i++) match[i] = 'C';
    orc_ss_set_stop(&m->s, m->s.stop + (insertion - deletion));
    return 1;
}
static int unclip(msite* m, const int8_t* bases, const int8_t* ref, int refLen) {
    if (!m->match || m->mlen < 1) return 0;
    if (m->match[0] != 'C' && m->match[m->mlen - 1] != 'C') return 0;
    for (int rloc = m->s.start, cloc = 0, mloc = 0; mloc < m->mlen; mloc++) {
        const int8_t x = m->match[mloc];
        if (x == 'C') {
            const int8_t c = bases[cloc], r = ca_get(ref, refLen, rloc);
            if (!fully_defined(c) || !fully_defined(r)) m->match[mloc] = 'N'; else m->match[mloc] = (int8_t)(c == r ? 'm' : 'S');
            rloc++; cloc++;
        } else if (x == 'I') cloc++;
        else if (x == 'D') rloc++;
        else { rloc++; cloc++; }
    }
    return 1;
}
/* SiteScore.clipTipIndels(bases, tiplen, maxIndel, msa) */
static int clip_tip_indels(const mctx* C, msite* m, const int8_t* bases, int len, int tiplen, int maxIndel) {
    if (!m->match || m->mlen < maxIndel) return 0;
    int refLen; const int8_t* ref = chrom_ptr(C, m->s.chrom, &refLen);
    const int left = clip_left_tip_indel(m, tiplen, maxIndel);
    const int right = clip_right_tip_indel(m, tiplen, maxIndel);
    if (left || right) {
        unclip(m, bases, ref, refLen);
        const int oldScore = m->s.slow_score;
        orc_ss_set_slow_score(&m->s, orc_score_match(m->match, m->mlen));
        m->s.score = m->s.score + (m->s.slow_score - oldScore);
        set_perfect(&m->s, bases, len, ref, refLen);
    }
    return left | right;
}
/* SiteScore.fixXY(bases, nullifyOnFailure=false, msa) */
static int fix_xy(const mctx* C, msite* m, const int8_t* bases, int len) {
    if (!match_contains_xy(m)) return 1;
    int refLen; const int8_t* ref = chrom_ptr(C, m->s.chrom, &refLen);
    int8_t* match = m->match; const int mlen = m->mlen;
    int success = 1;
    const float maxSubRate = 0.4f; const int maxSubs = 5;
    {
        int mloc = 0;
        while (mloc < mlen && (match[mloc] == 'X' || match[mloc] == 'Y')) mloc++;
        if (mloc >= mlen || mloc >= len) success = 0;
        else if (mloc > 0) {
            mloc--;
            const int numX = mloc + 1;
            int rloc = m->s.start + mloc, cloc = mloc, subs = 0, firstSub = -1;
            while (mloc >= 0) {
                const int8_t c = bases[cloc], r = ca_get(ref, refLen, rloc);
                if (r == 'N' || c == 'N') match[mloc] = 'N';
                else if (c == r) match[mloc] = 'm';
                else { match[mloc] = 'S'; subs++; if (subs == 1) firstSub = mloc; }
                mloc--; rloc--; cloc--;
            }
            if (success && (m->s.stop - m->s.start + 1) != match_ref_length(match, mlen)) orc_ss_set_start(&m->s, m->s.start - numX);
            if (subs > maxSubs && (float)subs > (float)numX * maxSubRate) for (int i = 0; i <= firstSub; i++) match[i] = 'C';
        }
    }
    if (success) {
        int mloc = mlen - 1;
        while (mloc >= 0 && (match[mloc] == 'X' || match[mloc] == 'Y')) mloc--;
        const int dif = mlen - 1 - mloc;
        if (mloc < 0) success = 0;
        else if (dif > 0) {
            mloc++;
            const int numX = mlen - mloc;
            int rloc = m->s.stop - dif + 1, cloc = len - dif, subs = 0, firstSub = -1;
            if (cloc < 0) success = 0;
            else while (mloc < mlen) {
                const int8_t c = bases[cloc], r = ca_get(ref, refLen, rloc);
                if (r == 'N' || c == 'N') match[mloc] = 'N';
                else if (c == r) match[mloc] = 'm';
                else { match[mloc] = 'S'; subs++; if (subs == 1) firstSub = mloc; }
                mloc++; rloc++; cloc++;
            }
            if (success) {
                if ((m->s.stop - m->s.start + 1) != match_ref_length(match, mlen)) orc_ss_set_stop(&m->s, m->s.stop + numX);
                if (subs > maxSubs && (float)subs > (float)numX * maxSubRate) for (int i = firstSub; i < mlen; i++) match[i] = 'C';
            }
        }
    }
    success = success && !match_contains_xy(m);
    {
        const int oldScore = m->s.slow_score;
        orc_ss_set_slow_score(&m->s, orc_score_match(m->match, m->mlen));
        m->s.score = m->s.score + (m->s.slow_score - oldScore);
    }
    set_perfect(&m->s, bases, len, ref, refLen);
    return success;
}
static void fix_limits_xy(msite* m) {
    if (!m->match || m->mlen < 1) return;
    int y = 0;
    for (int i = m->mlen - 1; i >= 0; i--) { if (m->match[i] == 'Y') y++; else break; }
    if (y != 0) orc_ss_set_limits(&m->s, m->s.start, m->s.stop + y);
}

/* the "how much more padding" block that realign_new repeats (TranslateColorspaceRead.java:372-395 etc.); withElse = the branch that
 * raises both pads to x exists (it is missing for ungapped sites on the minus strand the first time, :571-579) */
static void adjust_pads(int gapped, int greflen, int span, int* epl, int* epr, int withElse) {
    int newlen = gapped ? (greflen + 1 + *epl + *epr) : (span + *epl + *epr);
    if (newlen >= MAXCOLS - 80) {
        while (newlen >= MAXCOLS - 80 && *epl > *epr) { newlen--; (*epl)--; }
        while (newlen >= MAXCOLS - 80 && *epl < *epr) { newlen--; (*epr)--; }
        while (newlen >= MAXCOLS - 80) { newlen -= 2; (*epl)--; (*epr)--; }
    } else if (withElse) {
        const int x = imax2(0, imin2(20, ((MAXCOLS - newlen) / 2) - 40));
        *epl = imax2(x, *epl); *epr = imax2(x, *epr);
    }
}
static int gref_len_of(int a, int b, const int32_t* gaps, int ngaps) { orc_ss t; memset(&t, 0, sizeof t); t.start = a; t.stop = b; t.ngaps = ngaps; memcpy(t.gaps, gaps, sizeof(t.gaps)); return orc_calc_gref_len(&t); }

/* one msa.fillLimited + msa.score; returns score_len (0 = null) */
static int fill_and_score(mctx* C, msite* m, const int8_t* bases, int len, const int8_t* ref, int refLen, int minLoc, int maxLoc, int minScore, int unlimited,
                          int32_t* max4, int32_t* score8) {
    int32_t g[ORC_MAX_GAPS]; memcpy(g, m->s.gaps, sizeof(m->s.gaps));
    C->fills++;
    const int ok = unlimited ? orc_msa_fillUnlimited(C->msa, bases, len, ref, refLen, minLoc, maxLoc, m->s.ngaps ? g : 0, m->s.ngaps, max4)
                             : orc_msa_fillLimited(C->msa, bases, len, ref, refLen, minLoc, maxLoc, minScore, m->s.ngaps ? g : 0, m->s.ngaps, max4);
    if (ok <= 0) return 0;
    return orc_msa_score(C->msa, bases, ref, minLoc, maxLoc, max4[0], max4[1], max4[2], m->s.ngaps > 0, score8);
}

static void realign_new(mctx* C, msite* m, const int8_t* bases, int len, int padding, int recur, int minValidScore, int forbidIndels, int fixXY) {
    orc_ss* ss = &m->s;
    if (match_contains_xy(m)) fix_xy(C, m, bases, len);
    clip_tip_indels(C, m, bases, len, 4, 10);
    padding = imin2(padding, (MAXCOLS - len) / 2 - 20);
    padding = imax2(padding, 0);
    int refLen; const int8_t* ref = chrom_ptr(C, ss->chrom, &refLen);
    const int maxIndex = refLen - 1;
    {
        const int expectedLen = orc_calc_gref_len(ss);
        if (expectedLen > MAXCOLS - 20) { orc_ss_set_stop(ss, ss->start + imin2(len + 40, MAXCOLS - 20)); fix_gaps_ss(ss); }
    }
    if (ss->start < 0) orc_ss_set_start(ss, 0);
    if (ss->stop > maxIndex) orc_ss_set_stop(ss, maxIndex);
    {
        const int a = len, b = ss->stop - ss->start + 1;
        if (b < a) { const int c = imin2(len, a - b + 10) / 2; padding = imax2(padding, c + 1); }
    }
    padding = imax2(0, imin2(padding, (MAXCOLS - imax2(len, orc_calc_gref_len(ss))) / 2 - 100));
    if (forbidIndels) padding = 0;
    const int maxQ = 70 + (len - 1) * 100, maxI = maxQ + imin2(-472, -395 - 100);
    (void)maxQ;
    const int minusNoElse = (ss->strand != 0);          /* the minus-strand copy of the block lacks the `else` of the first adjustment for ungapped sites */
    int8_t* nm = (int8_t*)calloc((size_t)len + 1, 1);
    if (m->match && m->mlen == len) memcpy(nm, m->match, (size_t)len);
    const int scoreNoIndel = orc_score_no_indels(bases, len, ref, refLen, ss->start, nm);
    set_match(m, nm, len);
    free(nm);
    if (scoreNoIndel >= maxI || forbidIndels) {
        orc_ss_set_stop(ss, ss->start + len - 1);
        orc_ss_set_slow_score(ss, scoreNoIndel);
    } else {
        int minLoc = imax2(ss->start - padding, 0), maxLoc = imin2(ss->stop + padding, maxIndex);
        int32_t max4[4], score[8]; int n;
        const int lim = imax2(scoreNoIndel, minValidScore);
        n = fill_and_score(C, m, bases, len, ref, refLen, minLoc, maxLoc, lim, 0, max4, score);
        if (n > 6) {
            const int old0 = score[0];
            int epl = score[6], epr = score[7];
            const int gapped = ss->ngaps > 0;
            adjust_pads(gapped, gapped ? imax2(len, gref_len_of(minLoc, maxLoc, ss->gaps, ss->ngaps)) : 0, maxLoc - minLoc + 1, &epl, &epr, gapped || !minusNoElse);
            minLoc = imax2(0, minLoc - epl); maxLoc = imin2(maxIndex, maxLoc + epr);
            n = fill_and_score(C, m, bases, len, ref, refLen, minLoc, maxLoc, lim, 0, max4, score);
            if (n == 0 || score[0] < old0) {
                adjust_pads(gapped, gapped ? imax2(len, gref_len_of(minLoc, maxLoc, ss->gaps, ss->ngaps)) : 0, maxLoc - minLoc + 1, &epl, &epr, 1);
                minLoc = imax2(0, minLoc - epl); maxLoc = imin2(maxIndex, maxLoc + epr);
                n = fill_and_score(C, m, bases, len, ref, refLen, minLoc, maxLoc, lim, 0, max4, score);
                if (ss->strand == 0 && minLoc > 0 && maxLoc < maxIndex && (n == 0 || score[0] < old0)) {
                    minLoc = imax2(ss->start - 8, 0); maxLoc = imin2(ss->stop + 8, maxIndex);
                    n = fill_and_score(C, m, bases, len, ref, refLen, minLoc, maxLoc, 0, 1, max4, score);
                }
            }
        }
        if (n > 0) {
            const int cap = len + (maxLoc - minLoc + 1) + 70000;
            int8_t* tb = (int8_t*)malloc((size_t)cap);
            const int tl = orc_msa_traceback(C->msa, bases, ref, minLoc, maxLoc, max4[0], max4[1], max4[2], ss->ngaps > 0, tb, cap);
            if (tl < 0) { C->status |= ORC_MAP_ST_MATCH_OVERFLOW; set_match(m, tb, 0); }
            else set_match(m, tb, tl);
            free(tb);
            orc_ss_set_limits(ss, score[1], score[2]);
            fix_limits_xy(m);
            orc_ss_set_slow_score(ss, score[0]);
        } else {
            orc_ss_set_stop(ss, ss->start + len - 1);
            orc_ss_set_slow_score(ss, scoreNoIndel);
        }
    }
    const int lp = left_padding_needed(m, 4, 5), rp = right_padding_needed(m, 4, 5);
    if (ss->stop < maxIndex && ss->start > 0 && (lp > 0 || rp > 0)) {
        if (recur > 0) {
            fix_gaps_ss(ss);
            const int p_temp = imin2(10 + imax2(lp, rp), (MAXCOLS - len) / 2 - 20);
            realign_new(C, m, bases, len, p_temp, recur - 1, minValidScore, forbidIndels, fixXY);
        } else if (fixXY && match_contains_xy(m)) fix_xy(C, m, bases, len);
    }
    set_perfect(ss, bases, len, ref, refLen);
}

/* AbstractMapThread.genMatchStringForSite (GEN_MATCH_FAST) */
static void gen_match_string_for_site(mctx* C, msite* m, const int8_t* basesP, const int8_t* basesM, int len, int maxSwScore, int secondary) {
    const int8_t* bases = m->s.strand == 0 ? basesP : basesM;
    const orc_map_cfg* cfg = C->cfg;
    const float mult = (cfg->paired ? cfg->min_ratio_paired : cfg->min_ratio) * (secondary ? cfg->secondary_site_score_ratio : 1.f);
    const int minMsaLimit = -1 + (int)(mult * (float)maxSwScore);
    int refLen; const int8_t* ref = chrom_ptr(C, m->s.chrom, &refLen);
    if (m->s.perfect) {
        int8_t* p = (int8_t*)malloc((size_t)len); memset(p, 'm', (size_t)len); set_match(m, p, len); free(p);
    } else {
        const int oldScore = m->s.slow_score;
        const int padding = (m->s.perfect || m->s.semiperfect) ? 0 : imax2(cfg->slow_align_padding, 6);
        realign_new(C, m, bases, len, padding, 1, minMsaLimit, cfg->max_indel < 1, 0);
        fix_gaps_ss(&m->s);
        const int lp = left_padding_needed(m, 4, 5), rp = right_padding_needed(m, 4, 5);
        if (m->s.slow_score < oldScore || lp > 0 || rp > 0) {
            int extra = (cfg->max_indel > 0 ? 80 : 20) + cfg->slow_align_padding;
            const int expectedLen = orc_calc_gref_len(&m->s);
            const int remaining = MAXCOLS - expectedLen - 2;
            extra = imax2(0, imin2(remaining / 2, extra));
            realign_new(C, m, bases, len, extra, 2, minMsaLimit, 0, 1);
            fix_gaps_ss(&m->s);
        }
        if (maxSwScore == m->s.slow_score) { m->s.perfect = 1; m->s.semiperfect = 1; }       /* setPerfectFlag(maxSwScore, bases) */
        else set_perfect(&m->s, bases, len, ref, refLen);
    }
    clip_tip_indels(C, m, bases, len, 4, 10);
}

/* SiteScore.compareTo / PCOMP / positionalMatch (as in sitelist_oracle.c) */
static int ss_cmp(const orc_ss* a, const orc_ss* o) {
    int x = o->score - a->score; if (x) return x;
    x = o->slow_score - a->slow_score; if (x) return x;
    x = o->paired_score - a->paired_score; if (x) return x;
    x = o->quick_score - a->quick_score; if (x) return x;
    x = a->chrom - o->chrom; if (x) return x;
    return a->start - o->start;
}
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
static void msort(msite* v, int n, int positional) {          /* stable (Collections.sort) */
    for (int i = 1; i < n; i++) {
        const msite x = v[i]; int j = i - 1;
        while (j >= 0 && (positional ? ss_pcomp(&v[j].s, &x.s) : ss_cmp(&v[j].s, &x.s)) > 0) { v[j + 1] = v[j]; j--; }
        v[j + 1] = x;
    }
}
static int check_order(const msite* v, int n) { for (int i = 1; i < n; i++) if (v[i].s.score > v[i - 1].s.score) return 0; return 1; }   /* Read.CHECKORDER (stream/Read.java:3141-3150): scores only */
static int positional_match(const orc_ss* a, const orc_ss* b, int testGaps) {
    if (a->chrom != b->chrom || a->strand != b->strand || a->start != b->start || a->stop != b->stop) return 0;
    if (!testGaps || (a->ngaps == 0 && b->ngaps == 0)) return 1;
    if (a->ngaps != b->ngaps) return 0;
    for (int i = 0; i < a->ngaps; i++) if (a->gaps[i] != b->gaps[i]) return 0;
    return 1;
}
/* Tools.mergeDuplicateSites(list, false, false) */
static int merge_duplicates_nogaps(msite* v, int n) {
    if (n < 2) return n;
    msort(v, n, 1);
    int ai = 0; int k = 0;
    char* dead = (char*)calloc((size_t)n, 1);
    for (int i = 1; i < n; i++) {
        orc_ss* a = &v[ai].s; const orc_ss* b = &v[i].s;
        if (positional_match(a, b, 1)) {
            orc_ss_set_slow_score(a, imax2(a->slow_score, b->slow_score));
            a->paired_score = (a->paired_score <= a->slow_score && b->paired_score <= a->slow_score) ? 0 : imax2(0, imax2(a->paired_score, b->paired_score));
            a->score = imax2(a->score, b->score);
            a->perfect = (a->perfect || b->perfect); a->semiperfect = (a->semiperfect || b->semiperfect);
            dead[i] = 1;
        } else ai = i;
    }
    for (int i = 0; i < n; i++) { if (!dead[i]) v[k++] = v[i]; else if (v[i].match) free(v[i].match); }
    free(dead);
    return k;
}

/* AbstractMapThread.genMatchString(r, basesP, basesM, maxImperfect, maxSw, setSSScore, recur) for PRINT_SECONDARY_ALIGNMENTS=false.
 * *pairedFlag (may be NULL): r.paired(), cleared when the top site changes identity (:936-939). */
static void gen_match_string(mctx* C, msite* v, int* np, const int8_t* basesP, const int8_t* basesM, int len, int maxSwScore, int setSSScore, int* pairedFlag) {
    int n = *np;
    if (n == 0) return;
    int best = -0x7fffffff - 1, scoreChanged = 0;
    for (int i = 0; i < n; i++) {
        msite* ss = &v[i];
        if (i > 0 && best >= ss->s.slow_score) break;
        const int oldSlow = ss->s.slow_score, oldScore = ss->s.score;
        if (ss->match == NULL) {
            gen_match_string_for_site(C, ss, basesP, basesM, len, maxSwScore, 0);
            if (setSSScore) ss->s.score = ss->s.slow_score;
        }
        if (i > 0 && ss->match == NULL && !(pairedFlag && *pairedFlag)) { for (int k = i; k + 1 < n; k++) v[k] = v[k + 1]; n--; }   /* r.sites.remove(i); the loop's i++ then skips one (as in the Java) */
        else { if (oldScore != ss->s.score || oldSlow != ss->s.slow_score) scoreChanged++; best = imax2(ss->s.slow_score, best); }
    }
    int needsSorting = (scoreChanged > 0 && !check_order(v, n));
    while (needsSorting) {
        needsSorting = 0;
        const int top = v[0].serial;
        n = merge_duplicates_nogaps(v, n);
        msort(v, n, 0);
        if (n > 0 && v[0].match == NULL) {
            gen_match_string_for_site(C, &v[0], basesP, basesM, len, maxSwScore, 0);
            if (setSSScore) v[0].s.score = v[0].s.slow_score;
            needsSorting = 1;
        }
        if (pairedFlag && *pairedFlag && v[0].serial != top) *pairedFlag = 0;
    }
    *np = n;
}

/* MSA.toLocalAlignment(r, ss, basesM, minToClip, matchPointsMult = LOCAL_ALIGN_MATCH_POINT_RATIO = 1f, BBMapThread.java:35) on the top site;
 * rstart/rstop/rmapScore/flags are r.start, r.stop, r.mapScore and the read's flag bits */
static int to_local_alignment(mctx* C, msite* top, const int8_t* bases, int len, int minToClip, int* rstart, int* rstop, int* rmapScore, int* flags, int depth) {
    if (!top->match || top->mlen < 1) return 0;
    int refLen; const int8_t* ref = chrom_ptr(C, top->s.chrom, &refLen);
    if (top->match[0] == 'X' || top->match[top->mlen - 1] == 'Y') { fix_xy(C, top, bases, len); *rstart = top->s.start; *rstop = top->s.stop; }
    const int8_t* match = top->match; const int mlen = top->mlen;
    int maxScore = -1, startLocC = -1, stopLocC = -1, lastZeroC = 0, startLocM = -1, stopLocM = -1, lastZeroM = 0, startLocR = -1, stopLocR = -1, lastZeroR = 0;
    int8_t mode = match[0], prevMode = '0'; int current = 0, prevStreak = 0, cpos = 0, rpos = *rstart, score = 0;
    for (int mpos = 0; mpos <= mlen; mpos++) {
        const int atEnd = (mpos == mlen);
        const int8_t c = atEnd ? 0 : match[mpos];
        if (!atEnd && mode == c) { current++; continue; }
        if (atEnd && current <= 0) break;
        if (mode == 'm') {
            if (score <= 0) { score = 0; lastZeroC = cpos; lastZeroM = mpos - current; lastZeroR = rpos; }
            score += calc_match_score(current);
            cpos += current; rpos += current;
            if (score > maxScore) { maxScore = score; startLocC = lastZeroC; startLocM = lastZeroM; startLocR = lastZeroR; stopLocC = cpos - 1; stopLocM = mpos - 1; stopLocR = rpos - 1; }
        } else if (mode == 'S') { score += run_points('S', current, prevMode, prevStreak); cpos += current; rpos += current; }
        else if (mode == 'D') { score += calc_del_score(current); rpos += current; }
        else if (mode == 'I') { score += calc_ins_score(current); cpos += current; }
        else if (mode == 'X' || mode == 'Y') { score += calc_ins_score(current); cpos += current; rpos += current; }
        else { cpos += current; rpos += current; }          /* C, N, R */
        prevMode = mode; prevStreak = current; mode = c; current = 1;
    }
    if (startLocC < 0 || stopLocC < 0) { *flags |= 256; return 0; }        /* r.clearMapping() */
    int headTrimR = startLocC, headTrimM = startLocM, tailTrimR = len - stopLocC - 1, tailTrimM = mlen - stopLocM - 1;
    if (headTrimR <= minToClip && headTrimM <= minToClip) headTrimR = headTrimM = 0;
    if (tailTrimR <= minToClip && tailTrimM <= minToClip) tailTrimR = tailTrimM = 0;
    if (headTrimR == 0 && headTrimM == 0 && tailTrimR == 0 && tailTrimM == 0) return 0;
    const int headDelta = headTrimR - headTrimM, tailDelta = tailTrimR - tailTrimM;
    if (headDelta == 0 && tailDelta == 0) {
        for (int i = 0; i < headTrimM; i++) top->match[i] = 'C';
        for (int i = mlen - tailTrimM; i < mlen; i++) top->match[i] = 'C';
    } else {
        const int newlen = mlen - headTrimM - tailTrimM + headTrimR + tailTrimR;
        int8_t* m2 = (int8_t*)malloc((size_t)(newlen > 0 ? newlen : 1));
        for (int i = 0; i < headTrimR; i++) m2[i] = 'C';
        for (int i = newlen - tailTrimR; i < newlen; i++) m2[i] = 'C';
        for (int i = headTrimM, i2 = headTrimR, lim = newlen - tailTrimR; i2 < lim; i++, i2++) m2[i2] = top->match[i];
        set_match(top, m2, newlen); free(m2);
    }
    if (headTrimR != 0) *rstart = startLocR - headTrimR;
    if (tailTrimR != 0) *rstop = stopLocR + tailTrimR;
    maxScore = imax2(maxScore, top->s.slow_score);
    *rmapScore = maxScore;
    orc_ss_set_limits(&top->s, *rstart, *rstop);
    if (!top->s.perfect && is_perfect(&top->s, bases, len, ref, refLen)) {
        top->s.perfect = top->s.semiperfect = 1; *flags |= 2;
        memset(top->match, 'm', (size_t)top->mlen);
        orc_ss_set_slow_score(&top->s, maxScore);
    } else if (!top->s.semiperfect && is_semiperfect(&top->s, bases, len, ref, refLen)) {
        top->s.semiperfect = 1;
        int8_t* nm = (int8_t*)calloc((size_t)len + 1, 1);
        orc_score_no_indels(bases, len, ref, refLen, top->s.start, nm);          /* genMatchNoIndels */
        set_match(top, nm, len); free(nm);
        if (depth < 4) return to_local_alignment(C, top, bases, len, minToClip, rstart, rstop, rmapScore, flags, depth + 1);
    }
    return 1;
}

/* AbstractMapThread.calcTipScorePenalty (:2499-2567) on r.match / r.bases / r.mapScore */
static int tip_penalty(const int8_t* match, int mlen, const int8_t* basesAsSequenced, int len, int mapped, int mapScore, int maxScore, int tiplen, int* status) {
    if (!mapped || !match || mlen < 1 || len < 2 * tiplen) return 0;
    int points = 0; int8_t prev = 'm';
    for (int i = 0, cpos = 0; cpos <= tiplen; i++) {
        if (i >= mlen) { *status |= ORC_MAP_ST_TIP; return 0; }
        const int8_t b = match[i];
        if (b == 'm') cpos++;
        else if (b == 'D') { if (prev != 'D') points += 2 * (tiplen + 2 - cpos); }
        else if (b == 'N' || b == 'C') { points += (tiplen + 2 - cpos); cpos++; }
        else { points += 2 * (tiplen + 2 - cpos); cpos++; }
        prev = b;
    }
    prev = 'm';
    for (int i = mlen - 1, cpos = 0; cpos <= tiplen; i--) {
        if (i < 0) { *status |= ORC_MAP_ST_TIP; return 0; }
        const int8_t b = match[i];
        if (b == 'm') cpos++;
        else if (b == 'D') { if (prev != 'D') points += 2 * (tiplen + 2 - cpos); }
        else if (b == 'N' || b == 'C') { points += (tiplen + 2 - cpos); cpos++; }
        else { points += 2 * (tiplen + 2 - cpos); cpos++; }
        prev = b;
    }
    const int last = len - 1;
    int8_t b = basesAsSequenced[0];
    if (b != 'N' && b == basesAsSequenced[1]) for (int i = 2; i <= tiplen && basesAsSequenced[i] == b; i++) points++;
    b = basesAsSequenced[last];
    if (b != 'N' && b == basesAsSequenced[last - 1]) for (int i = last - 2; i >= (last - tiplen) && basesAsSequenced[i] == b; i--) points++;
    if (points < 1) return 0;
    const float asymptote = 80.f;
    const float f = ((asymptote * (float)points) / ((float)points + asymptote));
    const int penalty = (int)(f * .0022f * (float)maxScore);
    const int maxPenalty = mapScore - maxScore / 10;
    if (maxPenalty <= 0) return 0;
    return imin2(penalty, maxPenalty);
}

static float cz3_mult(int i) { static const float t[7] = {0.f, 1.f, .75f, .5f, .25f, .125f, .0625f}; return t[i]; }
static float cz3_fraction(int score1, int score2, int cz3, float inv) {
    const int dif = score1 - score2;
    if (dif >= cz3) return 0.f;
    const float f = (float)(cz3 - dif) * inv;
    const float f2 = f * f;
    return f + 2.f * f2 + 2.f * f2 * f;
}

/* The tail of BBMapThread.processRead for one unpaired read whose list went through the final policy (orc_sitelist_final): :557-709. */
static void finish_single(mctx* C, msite* v, int* np, const int8_t* basesP, const int8_t* basesM, int len, int inFlags, const orc_policy_cfg* pc, orc_map_rec* rec,
                          int8_t* match_out, int match_cap)
{
    int n = *np;
    const int maxSw = 70 + (len - 1) * 100;
    int flags = inFlags & 7;                 /* bit0 mapped, bit1 perfect, bit2 ambiguous */
    int mapScore = n > 0 ? v[0].s.slow_score : 0;
    int rstart = -1, rstop = -1;
    C->status = 0;
    if (n > 0) {                             /* MAKE_MATCH_STRING: do { genMatchString } while (top.score < second.score) */
        int first = 1;
        do {
            if (!first) msort(v, n, 0);
            gen_match_string(C, v, &n, basesP, basesM, len, maxSw, 1, NULL);
            v[0].s.score = v[0].s.slow_score;
            first = 0;
        } while (n > 1 && v[0].s.score < v[1].s.score);
        mapScore = v[0].s.slow_score;
        flags = (flags & ~2) | (v[0].s.perfect ? 2 : 0);
        rstart = v[0].s.start; rstop = v[0].s.stop;
    }
    if (n > 1) {                             /* removeDuplicateBestSites */
        const orc_ss* t = &v[0].s;
        while (n > 1 && t->chrom == v[n - 1].s.chrom && t->strand == v[n - 1].s.strand && t->start == v[n - 1].s.start && t->stop == v[n - 1].s.stop) { if (v[n - 1].match) free(v[n - 1].match); n--; }
    }
    if (n > 0 && mapScore <= 0) { mapScore = 0; flags &= ~1; for (int i = 0; i < n; i++) if (v[i].match) free(v[i].match); n = 0; }
    if (n == 0) flags &= ~1;
    int sub_applied = 0;
    if ((pc->clearzone3 > pc->clearzone1 || pc->clearzone3 > pc->clearzonep) && n > 0 && !(flags & 4)) {
        const float cz3v2 = (float)pc->clearzone3 * ((1.25f < ((float)maxSw / (float)mapScore)) ? 1.25f : ((float)maxSw / (float)mapScore));
        const int cz3 = (int)cz3v2; const float inv = 1.f / cz3v2;
        if ((flags & 1) && n >= 2) {
            const int score1 = v[0].s.slow_score;
            float sub = 0.f;
            const int mx = imin2(7, n);
            for (int i = 1; i < mx; i++) {
                if (i > 2 && v[i].s.slow_score < v[i - 1].s.slow_score) break;
                const float f = cz3_fraction(score1, v[i].s.slow_score, cz3, inv);
                if (f <= 0.f) break;
                sub += f * cz3_mult(i);
            }
            if (sub > 0.f) {
                const float asym = 4.f + 0.03f * (float)len;
                sub = sub * 1.8f;
                const float sub2 = (float)cz3 * ((asym * sub) / (sub + asym));
                int subi = (int)(sub2 + 0.5f);
                if (subi >= mapScore - 300) subi = mapScore - 300;
                if (subi > 0) {
                    for (int i = 0; i < n; i++) { orc_ss_set_slow_score(&v[i].s, v[i].s.slow_score - subi); v[i].s.score -= subi; }
                    mapScore -= subi; sub_applied = subi;
                    if (mapScore < (int)((float)maxSw * pc->min_align_ratio)) flags |= 4;
                }
            }
        }
    }
    if ((flags & 4) && C->cfg->ambiguous_toss) { for (int i = 0; i < n; i++) if (v[i].match) free(v[i].match); n = 0; flags &= ~1; mapScore = 0; }
    if ((flags & 1) && n > 0 && v[0].match && v[0].mlen > 0) {
        const int8_t a = v[0].match[0], b = v[0].match[v[0].mlen - 1];
        if (a == 'X' || b == 'Y' || a == 'C' || b == 'C') {              /* r.containsXYC() (LOCAL_ALIGN is off by default) */
            const int8_t* bases = v[0].s.strand == 0 ? basesP : basesM;
            int f2 = flags;
            to_local_alignment(C, &v[0], bases, len, 1, &rstart, &rstop, &mapScore, &f2, 0);
            if (f2 & 256) { for (int i = 0; i < n; i++) if (v[i].match) free(v[i].match); n = 0; flags &= ~1; mapScore = 0; }
            else flags = f2 & 7;
        }
    }
    if (n == 0 || (!(flags & 4) && (float)mapScore < (float)maxSw * pc->min_align_ratio)) { for (int i = 0; i < n; i++) if (v[i].match) free(v[i].match); n = 0; flags &= ~1; mapScore = 0; }
    int penalty = 0;
    if (C->cfg->penalize_ambig && n > 0) {
        penalty = tip_penalty(v[0].match, v[0].mlen, basesP, len, flags & 1, mapScore, maxSw, 7, &C->status);
        if (penalty > 0) { mapScore -= penalty; for (int i = 0; i < n; i++) { orc_ss_set_slow_score(&v[i].s, v[i].s.slow_score - penalty); v[i].s.score -= penalty; } }
    }
    memset(rec, 0, sizeof *rec);
    rec->flags = flags; rec->map_score = mapScore; rec->cz3_sub = sub_applied; rec->tip_penalty = penalty; rec->status = C->status;
    if (n > 0 && (flags & 1)) {
        rec->chrom = v[0].s.chrom; rec->strand = v[0].s.strand; rec->start = rstart; rec->stop = rstop;
        rec->match_len = v[0].match ? v[0].mlen : 0;
        if (rec->match_len > match_cap) { rec->status |= ORC_MAP_ST_MATCH_OVERFLOW; rec->match_len = 0; }
        if (rec->match_len > 0) memcpy(match_out, v[0].match, (size_t)rec->match_len);
    } else { rec->chrom = -1; rec->start = -1; rec->stop = -1; rec->strand = 0; rec->match_len = 0; }
    *np = n;
}

int64_t orc_map_finish_single(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int64_t* read_off,
                              const int8_t* refs, const int64_t* chrom_off, const orc_policy_cfg* pc, const orc_map_cfg* cfg, const orc_read_out* flags_in,
                              orc_map_rec* recs, int8_t* match_buf, int64_t match_stride)
{
    mctx C; memset(&C, 0, sizeof C);
    C.msa = orc_msa_new(601, 3000); C.cfg = cfg; C.refs = refs; C.chrom_off = chrom_off;
    msite* v = (msite*)calloc((size_t)cap, sizeof(msite));
    for (int64_t r = 0; r < nreads; r++) {
        int n = nss[r];
        const int len = (int)(read_off[r + 1] - read_off[r]);
        for (int i = 0; i < n; i++) { v[i].s = lists[r * cap + i]; v[i].match = NULL; v[i].mlen = 0; v[i].serial = i; v[i].s.has_match = 0; }
        finish_single(&C, v, &n, basesP + read_off[r], basesM + read_off[r], len, flags_in[r].flags, pc, &recs[r], match_buf + r * match_stride, (int)match_stride);
        for (int i = 0; i < n; i++) { lists[r * cap + i] = v[i].s; lists[r * cap + i].has_match = v[i].match ? 1 : 0; if (v[i].match) free(v[i].match); v[i].match = NULL; }
        nss[r] = n;
    }
    free(v);
    const int64_t fills = C.fills;
    orc_msa_free(C.msa);
    return fills;
}

/* =====================  paired reads: BBMapThread.processReadPair (current/align2/BBMapThread.java:943-1362)  =====================
 * pairSiteScoresInitial :736-940, the rescue block :1061-1100 with AbstractMapThread.rescue / slowRescue (:1144-1306), removeLowQualitySitesPaired
 * (Tools.java:934-958), pairSiteScoresFinal / canPair (AbstractMapThread.java:1919-2170), the paired clearzone rule :1147-1176, Read.isBadPair
 * (stream/Read.java:1305-1331), genMatchString per mate and the bookkeeping behind it (:1186-1352).  One pair after the other. */
#include "rescue_oracle.h"
typedef struct { int32_t paired; float min_ratio, min_ratio_pre_rescue; int32_t clearzone1e, clearzone3, slow_align_padding, extra_padding, expected_len_limit; } orc_slow_cfg2;
int64_t orc_score_slow_with(orc_msa* msa, orc_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int64_t* read_off,
                            const int8_t* refs, const int64_t* chrom_off, const int32_t* run, const orc_slow_cfg2* cfg, int32_t* status);

static int absdif_i(int a, int b) { return a > b ? a - b : b - a; }
static void pair_dists(const orc_ss* ss1, const orc_ss* ss2, int requireCorrect, int* inner, int* outer) {
    int plusFirst;
    if (requireCorrect && ss1->strand != ss2->strand) plusFirst = (ss1->strand == 0);
    else plusFirst = (ss1->start <= ss2->start);
    if (plusFirst) { *inner = ss2->start - ss1->stop; *outer = ss2->stop - ss1->start; }
    else { *inner = ss1->start - ss2->stop; *outer = ss1->stop - ss2->start; }
}

static int pair_site_scores_initial(orc_ss* a, int* na, int len1, orc_ss* b, int* nb, int len2, const orc_map_cfg* cfg, int maxTrim) {
    if (*na < 1 || *nb < 1) return 0;
    orc_sl_sort(a, *na, 1); orc_sl_sort(b, *nb, 1);
    for (int i = 0; i < *na; i++) a[i].paired_score = 0;
    for (int i = 0; i < *nb; i++) b[i].paired_score = 0;
    int maxPaired1 = -1, maxPaired2 = -1, numPerfectPairs = 0;
    const int ilimit = *na - 1, jlimit = *nb - 1, maxReadLen = imax2(len1, len2);
    const int outerDistLimit = (maxReadLen * 14) / 32, innerDistLimit = cfg->max_pair_dist;
    const int apd = cfg->average_pair_dist, expectedFragLength = apd + len1 + len2;
    const int sameStrand = cfg->same_strand_pairs, requireCorrect = cfg->require_correct_strands;
    for (int i = 0, j = 0; i <= ilimit && j <= jlimit; i++) {
        orc_ss* ss1 = &a[i]; orc_ss* ss2 = &b[j];
        while (j < jlimit && (ss2->chrom < ss1->chrom || (ss2->chrom == ss1->chrom && ss1->start - ss2->stop > innerDistLimit))) { j++; ss2 = &b[j]; }
        for (int k = j; k <= jlimit; k++) {
            ss2 = &b[k];
            if (ss2->chrom > ss1->chrom) break;
            if (ss2->start - ss1->stop > innerDistLimit) break;
            int innerdist, outerdist;
            pair_dists(ss1, ss2, requireCorrect, &innerdist, &outerdist);
            if (outerdist >= outerDistLimit && innerdist <= innerDistLimit) {
                const int strandOK = ((ss1->strand == ss2->strand) == (sameStrand != 0));
                if (strandOK || !requireCorrect) {
                    int paired1 = 0, paired2 = 0;
                    const int deviation = absdif_i(apd, innerdist);
                    int ps1, ps2;
                    if (strandOK) {
                        ps1 = ss1->score + 1 + imax2(1, ss2->score / 2 - ((deviation * ss2->score) / (32 * expectedFragLength + 100)));
                        ps2 = ss2->score + 1 + imax2(1, ss1->score / 2 - ((deviation * ss1->score) / (32 * expectedFragLength + 100)));
                    } else { ps1 = ss1->score + imax2(0, ss2->score / 16); ps2 = ss2->score + imax2(0, ss1->score / 16); }
                    if (ps1 > ss1->paired_score) { paired1 = 1; ss1->paired_score = imax2(ss1->paired_score, ps1); maxPaired1 = imax2(ss1->score, maxPaired1); }
                    if (ps2 > ss2->paired_score) { paired2 = 1; ss2->paired_score = imax2(ss2->paired_score, ps2); maxPaired2 = imax2(ss2->score, maxPaired2); }
                    if (paired1 && paired2 && outerdist >= maxReadLen && deviation <= expectedFragLength && ss1->perfect && ss2->perfect) numPerfectPairs++;
                }
            }
        }
    }
    for (int i = 0; i < *na; i++) if (a[i].paired_score > a[i].score) a[i].score = a[i].paired_score;
    for (int i = 0; i < *nb; i++) if (b[i].paired_score > b[i].score) b[i].score = b[i].paired_score;
    if (numPerfectPairs > 0) {
        *na = orc_sl_trim_below_cutoff(a, *na, (int)((float)maxPaired1 * .94f), 0, 1, maxTrim);
        *nb = orc_sl_trim_below_cutoff(b, *nb, (int)((float)maxPaired2 * .94f), 0, 1, maxTrim);
    } else {
        if (*na > 4) *na = orc_sl_trim_below_cutoff(a, *na, (int)((float)maxPaired1 * .9f), 1, 1, maxTrim);
        if (*nb > 4) *nb = orc_sl_trim_below_cutoff(b, *nb, (int)((float)maxPaired2 * .9f), 1, 1, maxTrim);
    }
    return numPerfectPairs;
}

static void pair_site_scores_final(orc_ss* a, int* na, int len1, orc_ss* b, int* nb, int len2, const orc_map_cfg* cfg, int maxTrim) {
    for (int i = 0; i < *na; i++) a[i].paired_score = 0;
    for (int i = 0; i < *nb; i++) b[i].paired_score = 0;
    if (*na < 1 || *nb < 1) return;
    orc_sl_sort(a, *na, 1); orc_sl_sort(b, *nb, 1);
    int maxPaired1 = -1, maxPaired2 = -1;
    const float q1 = (float)len1 / (4.f * (float)len2), q2 = (float)len2 / (4.f * (float)len1);
    const float mult1 = (0.5f < ((0.25f > q1) ? 0.25f : q1)) ? 0.5f : ((0.25f > q1) ? 0.25f : q1);
    const float mult2 = (0.5f < ((0.25f > q2) ? 0.25f : q2)) ? 0.5f : ((0.25f > q2) ? 0.25f : q2);
    const int ilimit = *na - 1, jlimit = *nb - 1;
    const int outerDistLimit = (imax2(len1, len2) * 14) / 32, MPD = cfg->max_pair_dist, apd = cfg->average_pair_dist;
    const int expectedFragLength = apd + len1 + len2;
    const int sameStrand = cfg->same_strand_pairs, requireCorrect = cfg->require_correct_strands;
    for (int i = 0, j = 0; i <= ilimit && j <= jlimit; i++) {
        orc_ss* ss1 = &a[i]; orc_ss* ss2 = &b[j];
        while (j < jlimit && (ss2->chrom < ss1->chrom || (ss2->chrom == ss1->chrom && ss1->start - ss2->stop > MPD))) { j++; ss2 = &b[j]; }
        for (int k = j; k <= jlimit; k++) {
            ss2 = &b[k];
            if (ss2->chrom > ss1->chrom) break;
            if (ss2->start - ss1->stop > MPD) break;
            int innerdist, outerdist;
            pair_dists(ss1, ss2, requireCorrect, &innerdist, &outerdist);
            if (outerdist >= outerDistLimit && innerdist <= MPD) {
                const int strandOK = ((ss1->strand == ss2->strand) == (sameStrand != 0));
                if (strandOK || !requireCorrect) {
                    const int deviation = absdif_i(apd, innerdist);
                    int ps1, ps2;
                    if (strandOK) {
                        const int den = imax2(100, 10 * expectedFragLength + 100);
                        ps1 = ss1->score + 1 + imax2(1, (int)((float)ss2->score * mult1) - ((deviation * ss2->score) / den));
                        ps2 = ss2->score + 1 + imax2(1, (int)((float)ss1->score * mult2) - ((deviation * ss1->score) / den));
                    } else { ps1 = ss1->score + ss2->score / 16; ps2 = ss2->score + ss1->score / 16; }
                    ss1->paired_score = imax2(ss1->paired_score, ps1);
                    ss2->paired_score = imax2(ss2->paired_score, ps2);
                    maxPaired1 = imax2(ss1->score, maxPaired1);
                    maxPaired2 = imax2(ss2->score, maxPaired2);
                }
            }
        }
    }
    for (int i = 0; i < *na; i++) if (a[i].paired_score > a[i].score) a[i].score = a[i].paired_score;
    for (int i = 0; i < *nb; i++) if (b[i].paired_score > b[i].score) b[i].score = b[i].paired_score;
    const float f = cfg->secondary_site_score_ratio < 0.95f ? cfg->secondary_site_score_ratio : 0.95f;
    *na = orc_sl_trim_below_cutoff(a, *na, (int)((float)maxPaired1 * f), 0, 1, maxTrim);
    *nb = orc_sl_trim_below_cutoff(b, *nb, (int)((float)maxPaired2 * f), 0, 1, maxTrim);
}

static int can_pair(const orc_ss* ss1, const orc_ss* ss2, int len1, int len2, const orc_map_cfg* cfg) {
    if (ss1->chrom != ss2->chrom) return 0;
    if (cfg->require_correct_strands) { const int strandOK = ((ss1->strand == ss2->strand) == (cfg->same_strand_pairs != 0)); if (!strandOK) return 0; }
    const int outerDistLimit = (imax2(len1, len2) * 14) / 32;
    int inner, outer;
    pair_dists(ss1, ss2, cfg->require_correct_strands, &inner, &outer);
    return outer >= outerDistLimit && inner <= cfg->max_pair_dist;
}

/* test entry points: the pairing helpers on their own, so that a second restatement can be compared with them (tests/test_pairing_independent.py) */
int orc_test_pair_initial(orc_ss* a, int32_t* na, int len1, orc_ss* b, int32_t* nb, int len2, const orc_map_cfg* cfg, int maxTrim) {
    int x = *na, y = *nb; const int r = pair_site_scores_initial(a, &x, len1, b, &y, len2, cfg, maxTrim); *na = x; *nb = y; return r;
}
void orc_test_pair_final(orc_ss* a, int32_t* na, int len1, orc_ss* b, int32_t* nb, int len2, const orc_map_cfg* cfg, int maxTrim) {
    int x = *na, y = *nb; pair_site_scores_final(a, &x, len1, b, &y, len2, cfg, maxTrim); *na = x; *nb = y;
}
int orc_test_can_pair(const orc_ss* ss1, const orc_ss* ss2, int len1, int len2, const orc_map_cfg* cfg) { return can_pair(ss1, ss2, len1, len2, cfg); }

/* Tools.removeLowQualitySitesPaired */
static int remove_low_quality_paired(orc_ss* v, int n, int maxSw, float multSingle, float multPaired) {
    if (n == 0) return 0;
    const int th = (int)((float)maxSw * multSingle), thp = (int)((float)maxSw * multPaired);
    if (v[0].score < thp) return 0;
    int k = 0;
    for (int i = 0; i < n; i++) {
        const int dead = (v[i].paired_score > 0) ? (v[i].slow_score < thp) : (v[i].slow_score < th);
        if (!dead) v[k++] = v[i];
    }
    return k;
}

typedef struct { int64_t rescue_scans, rescue_fills; } pstats;

/* AbstractMapThread.rescue(anchor, loose, basesP, basesM, searchDist) */
static void rescue_dir(mctx* C, orc_ss* A, int nA, int lenA, orc_ss* L, int* nL, int cap, const int8_t* basesP, const int8_t* basesM, const int8_t* qualL, int lenL,
                       int searchDist, const orc_tipdel_cfg* tc, int clearzone1e, pstats* ps, int* status)
{
    const orc_map_cfg* cfg = C->cfg;
    if (searchDist > cfg->max_rescue_dist) return;
    if (nA == 0) return;
    const int maxLooseSw = 70 + (lenL - 1) * 100, maxAnchorSw = 70 + (lenA - 1) * 100, maxImperfect = maxLooseSw + imin2(-472, -395 - 100);
    const int bestLoose = (*nL == 0) ? 0 : L[0].slow_score, bestAnchor = A[0].slow_score;
    if (bestLoose == maxLooseSw && bestAnchor == maxAnchorSw && A[0].paired_score > 0) return;
    const int rescueScoreLimit = (int)(0.95f * (float)bestAnchor);
    const int retainScoreLimit = imax2((int)(0.68f * (float)bestLoose), (int)(0.4f * (float)maxLooseSw));
    const int retainScoreLimit2 = imax2((int)(0.95f * (float)bestLoose), (int)(0.55f * (float)maxLooseSw));
    const int maxMismatches = (bestLoose > maxImperfect) ? 5 : imin2(cfg->max_rescue_mismatches, (int)(0.60f * (float)lenL - 1.f));
    const int findTip = bestLoose < maxImperfect;
    int findRight = findTip, findLeft = findTip;
    if (findTip && qualL) {
        const int T = tc->max_tiplen;
        int minL = 0, avgL = 0, minF = 0, avgF = 0;
        if (T <= lenL) {
            int x = 0; minL = qualL[lenL - T];
            for (int i = lenL - T; i < lenL; i++) { const int b = qualL[i]; x += (b < 0 ? 0 : b); if (b < minL) minL = b; }
            avgL = x / T;
            x = 0; minF = qualL[0];
            for (int i = 0; i < T; i++) { const int b = qualL[i]; x += (b < 0 ? 0 : b); if (i >= 1 && b < minF) minF = b; }
            avgF = x / T;
        }
        findRight = (minL >= 6 && avgL >= 14); findLeft = (minF >= 6 && avgF >= 14);
    }
    const orc_rescue_cfg rc = {70, 100, 1, 100};
    const int n0 = nA;
    for (int ia = 0; ia < n0; ia++) {
        orc_ss* ssa = &A[ia];
        if (ssa->slow_score < rescueScoreLimit) break;
        if (!(ssa->paired_score == 0 && !ssa->rescued)) continue;
        const int searchIntoAnchor = ssa->stop - ssa->start - 1 + (lenA * 11 / 16);
        int loc, idealStart; const int8_t* bases;
        const int8_t strand = (int8_t)(cfg->same_strand_pairs ? ssa->strand : (ssa->strand ^ 1));
        const int searchRight = cfg->same_strand_pairs ? (strand == 0) : (strand == 1);
        if (cfg->same_strand_pairs) {
            if (ssa->strand == 1) { bases = basesM; loc = ssa->start + searchIntoAnchor; idealStart = ssa->start - cfg->average_pair_dist; }
            else { bases = basesP; loc = ssa->stop - searchIntoAnchor; idealStart = ssa->stop + cfg->average_pair_dist; }
        } else {
            if (ssa->strand == 0) { bases = basesM; loc = ssa->stop - searchIntoAnchor; idealStart = ssa->stop + cfg->average_pair_dist; }
            else { bases = basesP; loc = ssa->start + searchIntoAnchor; idealStart = ssa->start - cfg->average_pair_dist; }
        }
        int refLen; const int8_t* ref = chrom_ptr(C, ssa->chrom, &refLen);
        orc_rescue_task T; memset(&T, 0, sizeof T);
        T.read_off = 0; T.ref_off = 0; T.read_len = lenL; T.ref_len = refLen; T.min_index = 0; T.max_index = refLen - 1; T.loc = loc;
        T.search_dist = searchDist + searchIntoAnchor; T.ideal_start = idealStart; T.max_mismatches = maxMismatches; T.flags = searchRight ? 1 : 0;
        orc_rescue_out O;
        orc_rescue_batch(bases, ref, &T, 1, &rc, &O);
        ps->rescue_scans++;
        if (O.start < 0 || !O.in_bounds) continue;
        const int mismatches = O.mismatches;
        if (mismatches > maxMismatches) continue;
        orc_ss ss; memset(&ss, 0, sizeof ss);
        ss.chrom = ssa->chrom; ss.strand = strand; ss.start = O.start; ss.stop = O.stop; ss.hits = 0; ss.quick_score = O.score; ss.score = O.score;
        ss.perfect = (O.perfect & 1) ? 1 : 0; ss.semiperfect = (O.perfect & 2) ? 1 : 0; ss.rescued = 1;
        ss.slow_score = 0; ss.paired_score = 0;           /* setSlowScore(minMismatches) then setSlowScore(0): pairedScore ends at 0 */
        /* slowRescue(bases, ss, maxLooseSwScore, maxImperfectScore, findRight, findLeft) */
        {
            int sw = orc_score_no_indels(bases, lenL, ref, refLen, ss.start, 0);
            const int oldStart = ss.start;
            if (sw < maxImperfect && cfg->max_indel > 0) {
                orc_ss_set_slow_score(&ss, sw);
                if (findRight || findLeft) {
                    orc_tipdel_task tt; memset(&tt, 0, sizeof tt);
                    tt.read_len = lenL; tt.ref_len = refLen; tt.min_index = 0; tt.start = ss.start; tt.stop = ss.stop; tt.slow_score = ss.slow_score; tt.max_imperfect = maxImperfect;
                    tt.flags = (findRight ? 1 : 0) | (findLeft ? 2 : 0);
                    orc_tipdel_out to;
                    orc_tipdel_batch(bases, ref, &tt, 1, tc, &to);
                    if (to.right > 0 || to.left > 0) {
                        if (to.right > 0) orc_ss_set_stop(&ss, to.stop);
                        if (to.left > 0) orc_ss_set_start(&ss, to.start);
                        sw = orc_score_no_indels(bases, lenL, ref, refLen, ss.start, 0);
                    }
                }
                const int minMsaLimit = -clearzone1e + (int)(cfg->min_ratio_paired * (float)maxLooseSw);
                const int minscore = imax2(sw, minMsaLimit);
                int32_t max4[4], arr[8];
                const int n8 = orc_msa_fillAndScoreLimited(C->msa, bases, lenL, ref, refLen, ss.start - tc->slow_rescue_padding, ss.stop + tc->slow_rescue_padding, minscore, 0, 0, max4, arr);
                ps->rescue_fills++;
                if (n8 > 0) { orc_ss_set_slow_score(&ss, arr[0]); ss.score = ss.slow_score; orc_ss_set_start(&ss, arr[1]); orc_ss_set_stop(&ss, arr[2]); }
                else { orc_ss_set_slow_score(&ss, sw); ss.score = ss.slow_score; orc_ss_set_start(&ss, oldStart); orc_ss_set_stop(&ss, ss.start + lenL - 1); }
            } else { orc_ss_set_slow_score(&ss, sw); ss.score = ss.slow_score; orc_ss_set_stop(&ss, ss.start + lenL - 1); }
            ss.paired_score = ss.score + 1;
            ss.perfect = (ss.slow_score == maxLooseSw);
            if (ss.perfect) ss.semiperfect = 1; else orc_sl_set_perfect(&ss, bases, lenL, ref, refLen);
        }
        const int inb = (ss.start >= 0 && ss.stop <= refLen - 1);
        if (ss.score > retainScoreLimit && inb) {
            if (ss.score > retainScoreLimit2) {
                ss.paired_score = imax2(ss.paired_score, ss.slow_score + ssa->slow_score / 4);
                ssa->paired_score = imax2(ssa->paired_score, ssa->slow_score + ss.slow_score / 4);
            }
            if (*nL < cap) { L[*nL] = ss; (*nL)++; } else *status |= ORC_MAP_ST_LIST_OVERFLOW;
        }
    }
}

static int paired_clearzone(const orc_ss* top, int perfect, int maxSw, const orc_policy_cfg* pc) {
    if (perfect) return pc->clearzonep;
    if (top->score >= (int)((float)maxSw * pc->cz1b_scale - pc->cz1b_flat)) return pc->clearzone1;
    if (top->score >= (int)((float)maxSw * pc->cz1c_scale - pc->cz1c_flat)) return pc->clearzone1b;
    return pc->clearzone1c;
}

/* Read.isBadPair(requireCorrectStrands, sameStrandPairs, maxdist) */
static int is_bad_pair(const orc_map_rec* r, const orc_map_rec* m, const orc_map_cfg* cfg) {
    if (r->flags & 8) return 0;
    if (!(r->flags & 1) || !(m->flags & 1)) return 0;
    if (r->chrom != m->chrom) return 1;
    { const int inner = (r->start <= m->start) ? (m->start - r->stop) : (r->start - m->stop); if (inner > cfg->max_pair_dist) return 1; }
    if (cfg->require_correct_strands && ((r->strand == m->strand) != (cfg->same_strand_pairs != 0))) return 1;
    if (!cfg->same_strand_pairs) {
        if (r->strand == 0 && m->strand == 1) { if (r->start >= m->stop) return 1; }
        else if (r->strand == 1 && m->strand == 0) { if (m->start >= r->stop) return 1; }
    }
    return 0;
}

static void set_from_top(orc_map_rec* q, const orc_ss* v, int n) {       /* Read.setFromTopSite (ambig=best): clearSite / setFromSite */
    if (n == 0) { q->chrom = -1; q->strand = 0; q->start = -1; q->stop = -1; q->map_score = 0; q->flags &= ~(1 | 16); return; }
    q->flags |= 1;
    q->chrom = v[0].chrom; q->strand = v[0].strand; q->start = v[0].start; q->stop = v[0].stop; q->map_score = v[0].slow_score;
    q->flags = (q->flags & ~(2 | 16)) | (v[0].perfect ? 2 : 0) | (v[0].rescued ? 16 : 0);
}
static void clear_mapping(orc_map_rec* q, orc_map_rec* mate, int* n) {   /* Read.clearMapping */
    q->chrom = -1; q->strand = 0; q->start = -1; q->stop = -1; q->map_score = 0; q->match_len = 0;
    *n = 0; q->flags &= ~(1 | 8); mate->flags &= ~8;
}

/* test entry point: one SiteScore with its match string through a helper of realign_new / genMatchString (tests/test_clip_independent.py).
 * op 0 leftPaddingNeeded, 1 rightPaddingNeeded, 2 clipTipIndels, 3 fixXY, 4 unclip, 5 setPerfect.  match is edited in place (capacity mcap). */
int orc_test_site_op(int op, orc_ss* s, int8_t* match, int32_t* mlen, int32_t mcap, const int8_t* bases, int len, const int8_t* refs, const int64_t* chrom_off,
                     int tiplen, int maxIndel) {
    mctx C; memset(&C, 0, sizeof C); C.refs = refs; C.chrom_off = chrom_off;
    msite m; memset(&m, 0, sizeof m); m.s = *s; m.match = NULL; m.mlen = 0;
    if (*mlen >= 0) set_match(&m, match, *mlen);
    int refLen; const int8_t* ref = chrom_ptr(&C, m.s.chrom, &refLen);
    int r = 0;
    if (op == 0) r = left_padding_needed(&m, tiplen, maxIndel);
    else if (op == 1) r = right_padding_needed(&m, tiplen, maxIndel);
    else if (op == 2) r = clip_tip_indels(&C, &m, bases, len, tiplen, maxIndel);
    else if (op == 3) r = fix_xy(&C, &m, bases, len);
    else if (op == 4) r = unclip(&m, bases, ref, refLen);
    else if (op == 5) { set_perfect(&m.s, bases, len, ref, refLen); r = m.s.perfect; }
    *s = m.s;
    if (m.match) { if (m.mlen <= mcap) memcpy(match, m.match, (size_t)m.mlen); *mlen = m.mlen; free(m.match); } else *mlen = -1;
    return r;
}
/* test entry point: realign_new on one site (tests/test_realign_independent.py); returns the number of fills it asked for */
int orc_test_realign_new(orc_ss* s, int8_t* match, int32_t* mlen, int32_t mcap, const int8_t* bases, int len, const int8_t* refs, const int64_t* chrom_off,
                         int padding, int recur, int minValidScore, int forbidIndels, int fixXY) {
    static orc_msa* msa = NULL;
    if (!msa) msa = orc_msa_new(601, MAXCOLS);
    mctx C; memset(&C, 0, sizeof C); C.refs = refs; C.chrom_off = chrom_off; C.msa = msa;
    msite m; memset(&m, 0, sizeof m); m.s = *s; m.match = NULL; m.mlen = 0;
    if (*mlen >= 0) set_match(&m, match, *mlen);
    realign_new(&C, &m, bases, len, padding, recur, minValidScore, forbidIndels, fixXY);
    *s = m.s;
    if (m.match) { if (m.mlen <= mcap) memcpy(match, m.match, (size_t)m.mlen); *mlen = m.mlen; free(m.match); } else *mlen = -1;
    return (int)C.fills;
}
/* test entry point: AbstractMapThread.rescue(anchor, loose, ...) on one pair of lists (tests/test_rescue_independent.py); returns the status bits */
int orc_test_rescue(orc_ss* A, int nA, int lenA, orc_ss* L, int32_t* nL, int cap, const int8_t* basesP, const int8_t* basesM, const int8_t* qualL, int lenL,
                    int searchDist, const int8_t* refs, const int64_t* chrom_off, const orc_map_cfg* cfg, const orc_tipdel_cfg* tc, int clearzone1e, int64_t* counts) {
    static orc_msa* msa = NULL;
    if (!msa) msa = orc_msa_new(601, MAXCOLS);
    mctx C; memset(&C, 0, sizeof C); C.refs = refs; C.chrom_off = chrom_off; C.msa = msa; C.cfg = cfg;
    pstats ps; memset(&ps, 0, sizeof ps);
    int status = 0, n = *nL;
    rescue_dir(&C, A, nA, lenA, L, &n, cap, basesP, basesM, qualL, lenL, searchDist, tc, clearzone1e, &ps, &status);
    *nL = n;
    if (counts) { counts[0] = ps.rescue_scans; counts[1] = ps.rescue_fills; }
    return status;
}
/* test entry point: AbstractMapThread.genMatchString on one read's list (tests/test_genmatch_independent.py); sites come in without match strings.
 * Returns the number of fills; writes the surviving sites, the `paired` flag and the top site's match string. */
int orc_test_gen_match_string(orc_ss* lists, int32_t* n, int cap, const int8_t* basesP, const int8_t* basesM, int len, const int8_t* refs, const int64_t* chrom_off,
                              const orc_map_cfg* cfg, int maxSwScore, int setSSScore, int32_t* paired, int8_t* top_match, int32_t* top_mlen, int32_t mcap) {
    static orc_msa* msa = NULL;
    if (!msa) msa = orc_msa_new(601, MAXCOLS);
    mctx C; memset(&C, 0, sizeof C); C.refs = refs; C.chrom_off = chrom_off; C.msa = msa; C.cfg = cfg;
    msite* v = (msite*)calloc((size_t)(cap > 0 ? cap : 1), sizeof(msite));
    int k = *n;
    for (int i = 0; i < k; i++) { v[i].s = lists[i]; v[i].match = NULL; v[i].mlen = 0; v[i].serial = i; }
    int pf = *paired;
    gen_match_string(&C, v, &k, basesP, basesM, len, maxSwScore, setSSScore, &pf);
    *paired = pf;
    for (int i = 0; i < k; i++) lists[i] = v[i].s;
    *top_mlen = -1;
    if (k > 0 && v[0].match) { if (v[0].mlen <= mcap) memcpy(top_match, v[0].match, (size_t)v[0].mlen); *top_mlen = v[0].mlen; }
    for (int i = 0; i < k; i++) if (v[i].match) free(v[i].match);          /* (removed sites were released by gen_match_string itself or leak here: test code) */
    free(v);
    *n = k;
    return (int)C.fills;
}
int orc_test_remove_low_quality_paired(orc_ss* v, int n, int maxSw, float multSingle, float multPaired) { return remove_low_quality_paired(v, n, maxSw, multSingle, multPaired); }
int orc_test_is_bad_pair(const orc_map_rec* r, const orc_map_rec* m, const orc_map_cfg* cfg) { return is_bad_pair(r, m, cfg); }

int64_t orc_map_pairs(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int8_t* quality,
                      const int64_t* read_off, const int8_t* refs, const int64_t* chrom_off, const int32_t* nkeys, const orc_policy_cfg* pc,
                      const orc_map_cfg* cfg, const void* slow_cfg, const orc_tipdel_cfg* tc, orc_map_rec* recs, int8_t* match_buf, int64_t match_stride,
                      int64_t* stats /* [8]: slow alignments, realign fills, rescue scans, rescue fills, mated pairs, inner length sum */)
{
    mctx C; memset(&C, 0, sizeof C);
    C.msa = orc_msa_new(601, 3000); C.cfg = cfg; C.refs = refs; C.chrom_off = chrom_off;
    const orc_slow_cfg2* sc = (const orc_slow_cfg2*)slow_cfg;
    msite* mv = (msite*)calloc((size_t)cap, sizeof(msite));
    pstats ps = {0, 0};
    int64_t slowAl = 0, mated = 0, innerSum = 0;
    const int maxTrim = pc->max_trim_sites_to_retain;
    for (int64_t p = 0; p + 1 < nreads; p += 2) {
        orc_ss* v[2] = { lists + p * cap, lists + (p + 1) * cap };
        int n[2] = { nss[p], nss[p + 1] };
        int len[2], maxSw[2], st[2] = {0, 0};
        const int8_t* bP[2]; const int8_t* bM[2]; const int8_t* ql[2];
        orc_map_rec* q[2] = { &recs[p], &recs[p + 1] };
        for (int e = 0; e < 2; e++) {
            len[e] = (int)(read_off[p + e + 1] - read_off[p + e]); maxSw[e] = 70 + (len[e] - 1) * 100;
            bP[e] = basesP + read_off[p + e]; bM[e] = basesM + read_off[p + e]; ql[e] = quality ? quality + read_off[p + e] : 0;
            memset(q[e], 0, sizeof(orc_map_rec)); q[e]->chrom = -1; q[e]->start = -1; q[e]->stop = -1;
        }
        if (nkeys[p] < 0 && nkeys[p + 1] < 0) { q[0]->flags |= 32; q[1]->flags |= 32; nss[p] = 0; nss[p + 1] = 0; continue; }
        pair_site_scores_initial(v[0], &n[0], len[0], v[1], &n[1], len[1], cfg, maxTrim);
        for (int e = 0; e < 2; e++) {
            if (n[e] > 2) orc_sl_sort(v[e], n[e], 0);
            orc_sl_trim_list(v[e], &n[e], 1, maxSw[e], 0, 2, maxTrim);
        }
        for (int e = 0; e < 2; e++) for (int i = 0; i < n[e]; i++) v[e][i].score = v[e][i].quick_score;
        for (int e = 0; e < 2; e++) {
            if (n[e] <= 0) continue;
            orc_read_out ro; int32_t nn = n[e]; int32_t one = 1;
            orc_sitelist_noindel(v[e], &nn, 1, cap, basesP, basesM, read_off + p + e, refs, chrom_off, pc, &ro);      /* scoreNoIndels + Collections.sort */
            if (ro.near_perfect < 1) { orc_read_out r2; orc_sitelist_tipdel(v[e], &nn, 1, cap, basesP, basesM, quality, read_off + p + e, refs, chrom_off, 0, tc, &r2); }
            int32_t sst = 0;
            slowAl += orc_score_slow_with(C.msa, v[e], &nn, 1, cap, basesP, basesM, read_off + p + e, refs, chrom_off, &one, sc, &sst);
            if (sst) st[e] |= ORC_MAP_ST_SLOW;
            n[e] = orc_sl_merge_duplicates(v[e], nn);
        }
        if (cfg->do_rescue) {
            int unpaired[2] = {0, 0};
            for (int e = 0; e < 2; e++) for (int i = 0; i < n[e]; i++) if (v[e][i].paired_score == 0) unpaired[e]++;
            const int searchDist = imin2(cfg->max_pair_dist, 2 * cfg->average_pair_dist + 100);
            for (int e = 0; e < 2; e++) {
                const int o = e ^ 1;
                if (unpaired[e] > 0 && n[e] > 0) {
                    orc_sl_sort(v[e], n[e], 0);
                    n[e] = remove_low_quality_paired(v[e], n[e], maxSw[e], cfg->min_ratio_pre_rescue, cfg->min_ratio_pre_rescue);
                    rescue_dir(&C, v[e], n[e], len[e], v[o], &n[o], cap, bP[o], bM[o], ql[o], len[o], searchDist, tc, sc->clearzone1e, &ps, &st[o]);
                    n[o] = orc_sl_merge_duplicates(v[o], n[o]);
                }
            }
        }
        for (int e = 0; e < 2; e++) if (n[e] > 1) orc_sl_sort(v[e], n[e], 0);
        for (int e = 0; e < 2; e++) n[e] = remove_low_quality_paired(v[e], n[e], maxSw[e], cfg->min_ratio, cfg->min_ratio_paired);
        pair_site_scores_final(v[0], &n[0], len[0], v[1], &n[1], len[1], cfg, maxTrim);
        for (int e = 0; e < 2; e++) if (n[e] > 0) orc_sl_sort(v[e], n[e], 0);
        int perfect[2], ambiguous[2] = {0, 0};
        for (int e = 0; e < 2; e++) {
            perfect[e] = n[e] > 0 && (v[e][0].slow_score == maxSw[e] || v[e][0].perfect);       /* Read.setPerfectFlag, match == null */
            if (n[e] > 1) {
                const int cz = paired_clearzone(&v[e][0], perfect[e], maxSw[e], pc);
                if (orc_sl_count_top_scores(v[e], n[e], cz) > 1) ambiguous[e] = 1;                 /* processAmbiguous(..., AMBIGUOUS_TOSS=false) keeps the list */
            }
        }
        int paired = 0;
        if (n[0] > 0 && n[1] > 0 && can_pair(&v[0][0], &v[1][0], len[0], len[1], cfg)) paired = 1;
        for (int e = 0; e < 2; e++) {
            q[e]->flags = (perfect[e] ? 2 : 0) | (ambiguous[e] ? 4 : 0) | (paired ? 8 : 0);
            set_from_top(q[e], v[e], n[e]);
        }
        if (cfg->kill_bad_pairs && is_bad_pair(q[0], q[1], cfg)) {
            const int x = q[0]->map_score / len[0], y = q[1]->map_score / len[1];
            const int k = (x >= y) ? 1 : 0;                   /* clearAnswers(false) on the weaker mate */
            q[k]->chrom = -1; q[k]->strand = 0; q[k]->start = -1; q[k]->stop = -1; q[k]->map_score = 0; q[k]->flags = 0; n[k] = 0;
        }
        /* genMatchString per mate (setSSScore = false) */
        int slotlen[2] = {0, 0};
        for (int e = 0; e < 2; e++) {
            if (n[e] <= 0) continue;
            for (int i = 0; i < n[e]; i++) { mv[i].s = v[e][i]; mv[i].match = NULL; mv[i].mlen = 0; mv[i].serial = i; mv[i].s.has_match = 0; }
            int pflag = (q[e]->flags & 8) ? 1 : 0; const int pflag0 = pflag;
            C.status = 0;
            gen_match_string(&C, mv, &n[e], bP[e], bM[e], len[e], maxSw[e], 0, &pflag);
            st[e] |= C.status;
            if (pflag0 && !pflag) { q[0]->flags &= ~8; q[1]->flags &= ~8; }
            /* r.start ... r.mapScore = top site (:944-953) */
            q[e]->start = mv[0].s.start; q[e]->stop = mv[0].s.stop; q[e]->chrom = mv[0].s.chrom; q[e]->strand = mv[0].s.strand; q[e]->map_score = mv[0].s.slow_score;
            q[e]->flags = (q[e]->flags & ~(2 | 16)) | (mv[0].s.perfect ? 2 : 0) | (mv[0].s.rescued ? 16 : 0);
            int8_t* mo = match_buf + (p + e) * match_stride;
            if (mv[0].match && mv[0].mlen <= match_stride) { memcpy(mo, mv[0].match, (size_t)mv[0].mlen); slotlen[e] = mv[0].mlen; }
            else if (mv[0].match) st[e] |= ORC_MAP_ST_MATCH_OVERFLOW;
            for (int i = 0; i < n[e]; i++) { v[e][i] = mv[i].s; v[e][i].has_match = mv[i].match ? 1 : 0; if (mv[i].match && i > 0) { free(mv[i].match); mv[i].match = NULL; } }
            /* keep the top site's string in mv[0] for toLocalAlignment below: stash it per mate */
            if (e == 0) { q[0]->pad_[0] = 0; }
            q[e]->match_len = slotlen[e];
            if (mv[0].match) { free(mv[0].match); mv[0].match = NULL; }
        }
        /* anomaly blocks (:1228-1253) */
        for (int e = 0; e < 2; e++) {
            const int o = e ^ 1;
            if (q[e]->map_score > 0 && n[e] == 0) clear_mapping(q[e], q[o], &n[e]);
            else if (q[e]->map_score <= 0 && n[e] > 0) clear_mapping(q[e], q[o], &n[e]);
        }
        for (int e = 0; e < 2; e++) if (n[e] > 1) {         /* removeDuplicateBestSites */
            const orc_ss* t = &v[e][0];
            while (n[e] > 1 && t->chrom == v[e][n[e] - 1].chrom && t->strand == v[e][n[e] - 1].strand && t->start == v[e][n[e] - 1].start && t->stop == v[e][n[e] - 1].stop) n[e]--;
        }
        for (int e = 0; e < 2; e++) if ((q[e]->flags & 4) && cfg->ambiguous_toss) {
            n[e] = 0; q[e]->chrom = -1; q[e]->strand = 0; q[e]->start = -1; q[e]->stop = -1; q[e]->map_score = 0; q[e]->flags &= ~(1 | 8); q[e ^ 1]->flags &= ~8; q[e]->match_len = 0;
        }
        for (int e = 0; e < 2; e++) {                        /* toLocalAlignment for X/Y/C tips */
            if (!(q[e]->flags & 1) || n[e] == 0 || q[e]->match_len < 1) continue;
            int8_t* mo = match_buf + (p + e) * match_stride;
            const int8_t a = mo[0], b = mo[q[e]->match_len - 1];
            if (!(a == 'X' || b == 'Y' || a == 'C' || b == 'C')) continue;
            msite top; top.s = v[e][0]; top.match = NULL; top.mlen = 0; top.serial = 0;
            set_match(&top, mo, q[e]->match_len);
            int f2 = q[e]->flags & 7, rs = q[e]->start, rp = q[e]->stop, msc = q[e]->map_score;
            to_local_alignment(&C, &top, v[e][0].strand == 0 ? bP[e] : bM[e], len[e], 1, &rs, &rp, &msc, &f2, 0);
            if (f2 & 256) clear_mapping(q[e], q[e ^ 1], &n[e]);
            else {
                v[e][0] = top.s; v[e][0].has_match = 1;
                q[e]->start = rs; q[e]->stop = rp; q[e]->map_score = msc; q[e]->flags = (q[e]->flags & ~7) | (f2 & 7);
                if (top.mlen <= match_stride) { memcpy(mo, top.match, (size_t)top.mlen); q[e]->match_len = top.mlen; } else { st[e] |= ORC_MAP_ST_MATCH_OVERFLOW; q[e]->match_len = 0; }
            }
            free(top.match);
        }
        for (int e = 0; e < 2; e++) {
            if (!(q[e]->flags & 1) || n[e] == 0) { q[e]->chrom = -1; q[e]->strand = 0; q[e]->start = -1; q[e]->stop = -1; q[e]->match_len = 0; if (!(q[e]->flags & 1)) q[e]->map_score = 0; }
            q[e]->status = st[e];
            if (nkeys[p + e] < 0) q[e]->flags |= 32;
            nss[p + e] = n[e];
        }
        if ((q[0]->flags & 8) && (q[0]->flags & 1)) {        /* calcStatistics1: numMated, innerLengthSum */
            int inner = (q[0]->start <= q[1]->start) ? (q[1]->start - q[0]->stop) : (q[0]->start - q[1]->stop);
            inner = imin2(cfg->max_pair_dist, inner); inner = imax2(-160, inner);
            mated++; innerSum += inner;
        }
    }
    free(mv);
    if (stats) { stats[0] = slowAl; stats[1] = C.fills; stats[2] = ps.rescue_scans; stats[3] = ps.rescue_fills; stats[4] = mated; stats[5] = innerSum; }
    orc_msa_free(C.msa);
    return mated;
}
