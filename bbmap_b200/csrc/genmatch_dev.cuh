// genmatch_dev.cuh — device helpers of the match-string stage shared by genmatch.cu and pairing.cu: MSA.score(match), SiteScore.fixXY /
// clipTipIndels / unclip / padding tests, scoreNoIndelsAndMakeMatchString, isPerfect / isSemiPerfect, MSA.toLocalAlignment.  Reference lines at each function.
#pragma once
#include "sitelist_dev.cuh"
#include "mapper_kernels.cuh"

namespace bbm {

constexpr int GM_MAXCOLS = 3000;       // msa.maxColumns (ALIGN_COLUMNS)

// ---------------- match-string arithmetic ----------------
__device__ __forceinline__ int gm_calc_del(int len) {                      // MSA11tsJNI.calcDelScore(len, true)
    if (len <= 0) return 0;
    int score = -472;
    if (len > 256) { const int rem = len % 128, div = (len - 128) / 128; score += div * (-2); len = rem + 128; }
    if (len > 80) { score += ((len - 80 + 3) / 4) * (-1); len = 80; }
    if (len > 20) { score += (len - 20) * (-1); len = 20; }
    if (len > 5) { score += (len - 5) * (-9); len = 5; }
    if (len > 1) score += (len - 1) * (-33);
    return score;
}
__device__ __forceinline__ int gm_calc_ins(int len) {                      // POINTS_INS_ARRAY_C[len]
    if (len <= 0) return 0;
    if (len == 1) return -395;
    if (len < 6) return -395 - 39 * (len - 1);
    if (len < 21) return -395 - 39 * 4 - 23 * (len - 5);
    const long long s = -395 - 39 * 4 - 23 * 15 - 8ll * (len - 20);
    return s < -1046575 ? -1046575 : (int)s;
}
__device__ __forceinline__ int gm_calc_sub(int len) { int score = -127; if (len > 5) { score += (len - 5) * (-25); len = 5; } if (len > 1) score += (len - 1) * (-51); return score; }
__device__ __forceinline__ int gm_run_points(int8_t mode, int current, int8_t prevMode, int prevStreak) {
    if (mode == 'm') return 70 + (current - 1) * 100;
    if (mode == 'S') { int s = gm_calc_sub(current); if (prevMode == 'N' || prevMode == 'R') s += 76; else if (prevMode == 'm' && prevStreak < 2) s += -20; return s; }
    if (mode == 'D') return gm_calc_del(current);
    if (mode == 'I' || mode == 'X' || mode == 'Y') return gm_calc_ins(current);
    return 0;
}
static __device__ int gm_score_match(const int8_t* match, int n) {                // MSA.score(byte[] match)
    if (n < 1) return 0;
    int8_t mode = match[0], prevMode = '0'; int current = 0, prevStreak = 0, score = 0;
    for (int i = 0; i < n; i++) {
        const int8_t c = match[i];
        if (mode == c) current++;
        else { score += gm_run_points(mode, current, prevMode, prevStreak); prevMode = mode; prevStreak = current; mode = c; current = 1; }
    }
    if (current > 0) score += gm_run_points(mode, current, prevMode, prevStreak);
    return score;
}
__device__ __forceinline__ int gm_ref_length(const int8_t* m, int n) { int len = 0; for (int i = 0; i < n; i++) len += (m[i] != 'I'); return len; }
__device__ __forceinline__ bool gm_contains_xy(const int8_t* m, int n) {
    if (n < 1) return false;
    const int8_t a = m[0], b = m[n - 1];
    return a == 'X' || a == 'Y' || b == 'X' || b == 'Y';
}
__device__ __forceinline__ int8_t gm_ca_get(const int8_t* ref, int refLen, int loc) { return (loc < 0 || loc >= refLen - 1) ? (int8_t)'N' : ref[loc]; }   // ChromosomeArray.get
__device__ __forceinline__ bool gm_defined(int c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

static __device__ int gm_left_padding(const int8_t* m, int n, int tiplen, int maxIndel) {
    if (n < 1) return 0;
    int insertion = 0, xy = 0;
    for (int i = 0; i < n; i++) {
        const int8_t c = m[i];
        if (c == 'I') insertion++;
        else if (c == 'X' || c == 'Y') xy++;
        else if (c == 'D') return insertion + xy;
        else if (i >= tiplen) break;
    }
    return (insertion > maxIndel || xy > 0 || m[0] == 'I') ? insertion + xy : 0;
}
static __device__ int gm_right_padding(const int8_t* m, int n, int tiplen, int maxIndel) {
    if (n < 1) return 0;
    int insertion = 0, xy = 0;
    for (int i = n - 1; i >= 0; i--) {
        const int8_t c = m[i];
        if (c == 'I') insertion++;
        else if (c == 'X' || c == 'Y') xy++;
        else if (c == 'D') return insertion + xy;
        else if (i >= tiplen) break;                                       // the reference tests mloc>=tiplen here as well (SiteScore.java:482)
    }
    return (insertion > maxIndel || xy > 0 || m[n - 1] == 'I') ? insertion + xy : 0;
}

// MSA.scoreNoIndelsAndMakeMatchString(read, ref, refStart, matchReturn): -99999 and an untouched string when the read hangs over the array
static __device__ int gm_noindel_match(const int8_t* __restrict__ read, int len, const int8_t* __restrict__ ref, int refLen, int refStart, int8_t* match) {
    if (refStart < 0 || (long long)refStart + len > refLen) return -99999;
    int score = 0, mode = -1, timeInMode = 0;
    for (int k = 0; k < len; ++k) {
        const int c = read[k], r = ref[refStart + k];
        if (c == r && c != 'N') { if (mode == 0) { timeInMode++; score += 100; } else { timeInMode = 0; score += 70; } mode = 0; match[k] = 'm'; }
        else if (c < 0 || c == 'N') match[k] = 'N';
        else if (r < 0 || r == 'N') match[k] = 'N';
        else { if (mode == 3) timeInMode++; else timeInMode = 0; score += timeInMode == 0 ? -127 : (timeInMode < 5 ? -51 : -25); mode = 3; match[k] = 'S'; }
    }
    return score;
}

static __device__ bool gm_is_perfect(const bbm_ss& s, const int8_t* bases, int len, const int8_t* ref, int refLen) {
    if (len != s.stop - s.start + 1 || s.start < 0 || s.stop >= refLen) return false;
    for (int i = 0; i < len; i++) { const int8_t c = bases[i]; if (c != ref[s.start + i] || c == 'N') return false; }
    return true;
}
static __device__ bool gm_is_semiperfect(const bbm_ss& s, const int8_t* bases, int len, const int8_t* ref, int refLen) {
    if (len != s.stop - s.start + 1) return false;
    int readStart = 0, readStop = len, maxNoref = len / 2;
    const int refStop = s.start + len;
    if (s.start < 0) readStart = -s.start;
    if (refStop > refLen) readStop -= (refStop - refLen);
    for (int i = readStart; i < readStop; i++) {
        const int8_t c = bases[i], r = ref[s.start + i];
        if (c == 'N') return false;
        if (c != r) { maxNoref--; if (maxNoref < 0 || r != 'N') return false; }
    }
    return true;
}

// ---------------- SiteScore edits that read or rewrite the match string ----------------
static __device__ bool gm_clip_left(bbm_ss& ss, int8_t* match, int& mlen, int tiplen, int maxIndel) {
    if (mlen < maxIndel) return false;
    if (match[0] == 'C' || match[0] == 'Y' || match[0] == 'X') return false;
    int neutral = 0, insertion = 0, deletion = 0, mloc = 0;
    for (; mloc < mlen; mloc++) {
        const int8_t c = match[mloc];
        if (c == 'I') insertion++;
        else if (c == 'D') deletion++;
        else { neutral++; if (mloc >= tiplen) break; }
    }
    if (mloc >= mlen) mloc = mlen - 1;
    while (mloc >= 0 && match[mloc] == 'm') { mloc--; neutral--; }
    if (insertion <= maxIndel && deletion <= 4 * maxIndel) return false;
    int sum = neutral + insertion + deletion;
    if (deletion > 0) {
        int i = 0, j = 0;
        for (; i < sum; i++) if (match[i] != 'D') match[j++] = match[i];
        for (; i < mlen; i++, j++) match[j] = match[i];
        mlen = j;
    }
    sum = neutral + insertion;
    for (int i = 0; i < sum; i++) match[i] = 'C';
    ss_set_start(ss, ss.start - (insertion - deletion));
    return true;
}
static __device__ bool gm_clip_right(bbm_ss& ss, int8_t* match, int& mlen, int tiplen, int maxIndel) {
    if (mlen < maxIndel) return false;
    const int lastIndex = mlen - 1;
    if (match[lastIndex] == 'C' || match[lastIndex] == 'Y' || match[lastIndex] == 'X') return false;
    int neutral = 0, insertion = 0, deletion = 0, mloc = lastIndex;
    for (const int mn = lastIndex - tiplen; mloc >= 0; mloc--) {
        const int8_t c = match[mloc];
        if (c == 'I') insertion++;
        else if (c == 'D') deletion++;
        else { neutral++; if (mloc <= mn) break; }
    }
    if (mloc < 0) mloc = 0;
    while (mloc < mlen && match[mloc] == 'm') { mloc++; neutral--; }
    if (insertion <= maxIndel && deletion <= 4 * maxIndel) return false;
    const int sum = neutral + insertion + deletion, limit = mlen - sum;
    if (deletion > 0) {
        int j = limit;
        for (int i = limit; i < mlen; i++) if (match[i] != 'D') match[j++] = match[i];
        mlen = j;
    }
    for (int i = limit; i < mlen; i++) match[i] = 'C';
    ss_set_stop(ss, ss.stop + (insertion - deletion));
    return true;
}
static __device__ void gm_unclip(const bbm_ss& ss, int8_t* match, int mlen, const int8_t* bases, const int8_t* ref, int refLen) {
    if (mlen < 1 || (match[0] != 'C' && match[mlen - 1] != 'C')) return;
    for (int rloc = ss.start, cloc = 0, mloc = 0; mloc < mlen; mloc++) {
        const int8_t x = match[mloc];
        if (x == 'C') {
            const int8_t c = bases[cloc], r = gm_ca_get(ref, refLen, rloc);
            match[mloc] = (!gm_defined(c) || !gm_defined(r)) ? 'N' : (c == r ? 'm' : 'S');
            rloc++; cloc++;
        } else if (x == 'I') cloc++;
        else if (x == 'D') rloc++;
        else { rloc++; cloc++; }
    }
}
static __device__ bool gm_clip_tip_indels(bbm_ss& ss, int8_t* match, int& mlen, const int8_t* bases, int len, const int8_t* ref, int refLen, int tiplen, int maxIndel) {
    if (mlen < maxIndel) return false;
    const bool left = gm_clip_left(ss, match, mlen, tiplen, maxIndel);
    const bool right = gm_clip_right(ss, match, mlen, tiplen, maxIndel);
    if (left || right) {
        gm_unclip(ss, match, mlen, bases, ref, refLen);
        const int oldScore = ss.slow_score;
        set_slow_score(ss, gm_score_match(match, mlen));
        ss.score = ss.score + (ss.slow_score - oldScore);
        ss_set_perfect(ss, bases, len, ref, refLen);
    }
    return left || right;
}
static __device__ bool gm_fix_xy(bbm_ss& ss, int8_t* match, int mlen, const int8_t* bases, int len, const int8_t* ref, int refLen) {
    if (!gm_contains_xy(match, mlen)) return true;
    bool success = true;
    const int maxSubs = 5;
    {
        int mloc = 0;
        while (mloc < mlen && (match[mloc] == 'X' || match[mloc] == 'Y')) mloc++;
        if (mloc >= mlen || mloc >= len) success = false;
        else if (mloc > 0) {
            mloc--;
            const int numX = mloc + 1;
            int rloc = ss.start + mloc, cloc = mloc, subs = 0, firstSub = -1;
            while (mloc >= 0) {
                const int8_t c = bases[cloc], r = gm_ca_get(ref, refLen, rloc);
                if (r == 'N' || c == 'N') match[mloc] = 'N';
                else if (c == r) match[mloc] = 'm';
                else { match[mloc] = 'S'; subs++; if (subs == 1) firstSub = mloc; }
                mloc--; rloc--; cloc--;
            }
            if ((ss.stop - ss.start + 1) != gm_ref_length(match, mlen)) ss_set_start(ss, ss.start - numX);
            if (subs > maxSubs && (float)subs > __fmul_rn((float)numX, 0.4f)) for (int i = 0; i <= firstSub; i++) match[i] = 'C';
        }
    }
    if (success) {
        int mloc = mlen - 1;
        while (mloc >= 0 && (match[mloc] == 'X' || match[mloc] == 'Y')) mloc--;
        const int dif = mlen - 1 - mloc;
        if (mloc < 0) success = false;
        else if (dif > 0) {
            mloc++;
            const int numX = mlen - mloc;
            int rloc = ss.stop - dif + 1, cloc = len - dif, subs = 0, firstSub = -1;
            if (cloc < 0) success = false;
            else while (mloc < mlen) {
                const int8_t c = bases[cloc], r = gm_ca_get(ref, refLen, rloc);
                if (r == 'N' || c == 'N') match[mloc] = 'N';
                else if (c == r) match[mloc] = 'm';
                else { match[mloc] = 'S'; subs++; if (subs == 1) firstSub = mloc; }
                mloc++; rloc++; cloc++;
            }
            if (success) {
                if ((ss.stop - ss.start + 1) != gm_ref_length(match, mlen)) ss_set_stop(ss, ss.stop + numX);
                if (subs > maxSubs && (float)subs > __fmul_rn((float)numX, 0.4f)) for (int i = firstSub; i < mlen; i++) match[i] = 'C';
            }
        }
    }
    success = success && !gm_contains_xy(match, mlen);
    const int oldScore = ss.slow_score;
    set_slow_score(ss, gm_score_match(match, mlen));
    ss.score = ss.score + (ss.slow_score - oldScore);
    ss_set_perfect(ss, bases, len, ref, refLen);
    return success;
}
static __device__ void gm_fix_limits_xy(bbm_ss& ss, const int8_t* match, int mlen) {
    int y = 0;
    for (int i = mlen - 1; i >= 0; i--) { if (match[i] == 'Y') y++; else break; }
    if (y != 0) ss_set_limits(ss, ss.start, ss.stop + y);
}
static __device__ void gm_adjust_pads(bool gapped, int greflen, int span, int& epl, int& epr, bool withElse) {
    int newlen = gapped ? (greflen + 1 + epl + epr) : (span + epl + epr);
    if (newlen >= GM_MAXCOLS - 80) {
        while (newlen >= GM_MAXCOLS - 80 && epl > epr) { newlen--; epl--; }
        while (newlen >= GM_MAXCOLS - 80 && epl < epr) { newlen--; epr--; }
        while (newlen >= GM_MAXCOLS - 80) { newlen -= 2; epl--; epr--; }
    } else if (withElse) {
        const int x = imax(0, imin(20, ((GM_MAXCOLS - newlen) / 2) - 40));
        epl = imax(x, epl); epr = imax(x, epr);
    }
}
static __device__ int gm_gref_len(int a, int b, const bbm_ss& ss) { bbm_ss t = ss; t.start = a; t.stop = b; return calc_gref_len(t); }


// MSA.toLocalAlignment(r, ss, basesM, minToClip, 1f) on the top site (MSA.java:216-470); returns false when the read must be unmapped
static __device__ bool gm_to_local(bbm_ss& top, int8_t* match, int& mlen, long long ms, const int8_t* bases, int len, const int8_t* ref, int refLen, int minToClip,
                            int& rstart, int& rstop, int& mapScore, int& flags, int& status) {
    for (int depth = 0; depth < 5; depth++) {
        if (mlen < 1) return true;
        if (match[0] == 'X' || match[mlen - 1] == 'Y') { gm_fix_xy(top, match, mlen, bases, len, ref, refLen); rstart = top.start; rstop = top.stop; }
        int maxScore = -1, startLocC = -1, stopLocC = -1, lastZeroC = 0, startLocM = -1, stopLocM = -1, lastZeroM = 0, startLocR = -1, stopLocR = -1, lastZeroR = 0;
        int8_t mode = match[0], prevMode = '0'; int current = 0, prevStreak = 0, cpos = 0, rpos = rstart, score = 0;
        for (int mpos = 0; mpos <= mlen; mpos++) {
            const bool atEnd = (mpos == mlen);
            const int8_t c = atEnd ? 0 : match[mpos];
            if (!atEnd && mode == c) { current++; continue; }
            if (atEnd && current <= 0) break;
            if (mode == 'm') {
                if (score <= 0) { score = 0; lastZeroC = cpos; lastZeroM = mpos - current; lastZeroR = rpos; }
                score += 70 + (current - 1) * 100;
                cpos += current; rpos += current;
                if (score > maxScore) { maxScore = score; startLocC = lastZeroC; startLocM = lastZeroM; startLocR = lastZeroR; stopLocC = cpos - 1; stopLocM = mpos - 1; stopLocR = rpos - 1; }
            } else if (mode == 'S') { score += gm_run_points('S', current, prevMode, prevStreak); cpos += current; rpos += current; }
            else if (mode == 'D') { score += gm_calc_del(current); rpos += current; }
            else if (mode == 'I') { score += gm_calc_ins(current); cpos += current; }
            else if (mode == 'X' || mode == 'Y') { score += gm_calc_ins(current); cpos += current; rpos += current; }
            else { cpos += current; rpos += current; }
            prevMode = mode; prevStreak = current; mode = c; current = 1;
        }
        if (startLocC < 0 || stopLocC < 0) return false;                    // r.clearMapping()
        int headTrimR = startLocC, headTrimM = startLocM, tailTrimR = len - stopLocC - 1, tailTrimM = mlen - stopLocM - 1;
        if (headTrimR <= minToClip && headTrimM <= minToClip) headTrimR = headTrimM = 0;
        if (tailTrimR <= minToClip && tailTrimM <= minToClip) tailTrimR = tailTrimM = 0;
        if (headTrimR == 0 && headTrimM == 0 && tailTrimR == 0 && tailTrimM == 0) return true;
        if (headTrimR == headTrimM && tailTrimR == tailTrimM) {
            for (int i = 0; i < headTrimM; i++) match[i] = 'C';
            for (int i = mlen - tailTrimM; i < mlen; i++) match[i] = 'C';
        } else {
            const int newlen = mlen - headTrimM - tailTrimM + headTrimR + tailTrimR;
            if (newlen > ms) { status |= BBM_MAP_ST_MATCH_OVERFLOW; return true; }
            const int lim = newlen - tailTrimR, delta = headTrimR - headTrimM;     // match2[i2] = match[i2 - delta] for i2 in [headTrimR, lim)
            if (delta > 0) for (int i2 = lim - 1; i2 >= headTrimR; i2--) match[i2] = match[i2 - delta];
            else if (delta < 0) for (int i2 = headTrimR; i2 < lim; i2++) match[i2] = match[i2 - delta];
            for (int i = 0; i < headTrimR; i++) match[i] = 'C';
            for (int i = lim; i < newlen; i++) match[i] = 'C';
            mlen = newlen;
        }
        if (headTrimR != 0) rstart = startLocR - headTrimR;
        if (tailTrimR != 0) rstop = stopLocR + tailTrimR;
        maxScore = imax(maxScore, top.slow_score);
        mapScore = maxScore;
        ss_set_limits(top, rstart, rstop);
        if (!top.perfect && gm_is_perfect(top, bases, len, ref, refLen)) {
            top.perfect = 1; top.semiperfect = 1; flags |= 2;
            for (int i = 0; i < mlen; i++) match[i] = 'm';
            set_slow_score(top, maxScore);
            return true;
        } else if (!top.semiperfect && gm_is_semiperfect(top, bases, len, ref, refLen)) {
            top.semiperfect = 1;
            if (len > ms) { status |= BBM_MAP_ST_MATCH_OVERFLOW; return true; }
            for (int i = 0; i < len; i++) match[i] = 0;
            mlen = len;
            gm_noindel_match(bases, len, ref, refLen, top.start, match);       // genMatchNoIndels
            continue;                                                            // return toLocalAlignment(...)
        }
        return true;
    }
    return true;
}


}  // namespace bbm
