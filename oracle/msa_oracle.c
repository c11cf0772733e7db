/*
 * msa_oracle.c — TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * A plain-C restatement of BBMap's MultiStateAligner11ts as seen through its JNI
 * plug-in boundary.  Two halves:
 *
 *   (1) "port" fills  — orc_fill_unlimited / orc_fill_limitedX restate
 *       reference jni/MultiStateAligner11tsJNI.c:100-314 and :361-704 (which are
 *       themselves line-for-line twins of current/align2/MultiStateAligner11ts.java
 *       :624-878 and :131-607).  They are validated against the reference's own C
 *       compiled from /root/reference into oracle/_ref (see oracle/Makefile and
 *       tests/test_oracle_vs_reference.py) — bit-exact packed matrix, result and
 *       iteration counter.
 *
 *   (2) the Java-only half of the plug-in, which cannot run here (no JVM):
 *       constructor init   MultiStateAligner11tsJNI.java:71-113
 *       fillLimited rule    :116-164   (dispatch limited/unlimited, minScore-=120)
 *       traceback2          :376-495
 *       score2              :537-658
 *       makeGref + coordinate translation :668-801
 *       tables              :1576-1625
 *       MSA.fillAndScoreLimited  current/align2/MSA.java:103-134
 *       scoreNoIndels       :1034-1089 (see msa_oracle_noindel.c)
 *     These run on a `packed` matrix filled by EITHER the reference's C (kind
 *     "reference") or the port fills (kind "port"); the fill backend is a pair of
 *     function pointers.
 *
 * PARITY PINNING: half (1) is pinned against the reference's own native code and
 * the known-answer vectors of SURVEY.md Appendix B.  Half (2) is pinned only by
 * those KAT columns (which were themselves produced by a restatement): it is
 * "parity unpinned" against Java itself.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include "msa_oracle.h"

/* ---- constants (MultiStateAligner11tsJNI.java:1489-1563 == jni/...JNI.c:18-98) ---- */
enum {
    ST_MS = 0, ST_DEL = 1, ST_INS = 2,
    TBITS = 11,
    TMASK = (1 << TBITS) - 1,
    MAXTIME = TMASK,
    MAXSCORE = ((1 << 20) - 1) - 2000,
    MINSCORE = -MAXSCORE,
    BADSCORE = MINSCORE - 1,
    P_MATCH = 70, P_MATCH2 = 100,
    P_SUB = -127, P_SUBR = -147, P_SUB2 = -51, P_SUB3 = -25,
    P_INS = -395, P_INS2 = -39, P_INS3 = -23, P_INS4 = -8,
    P_DEL = -472, P_DEL2 = -33, P_DEL3 = -9, P_DEL4 = -1, P_DEL5 = -1,
    P_DEL_REF_N = -10, P_GAP = -2, P_NOCALL = 0, P_NOREF = 0,
    LIM3 = 5, LIM4 = 20, LIM5 = 80, SLIP = 4, M5 = 3,
    BAR_I1 = 2, BAR_D1 = 3,
    GAPLEN_ = 128, GAPBUFFER_ = 64, GAPBUFFER2_ = 128, MINGAP_ = 256, CUSHION_ = 128
};
#define SMASK ((int32_t)0xFFFFF800)
#define OFF(x) ((int32_t)((x) * 2048))
#define BADOFF OFF(BADSCORE)

static inline int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }

void orc_msa_tables(int32_t* sub_off, int32_t* ins_off, int32_t* insC_off, int32_t* sub_pts, int32_t* ins_pts, int32_t* insC_pts) {
    /* MultiStateAligner11tsJNI.java:1576-1625 */
    int32_t ic = 0, icp = 0;
    if (sub_off) sub_off[0] = 0;
    if (ins_off) ins_off[0] = 0;
    if (insC_off) insC_off[0] = 0;
    if (sub_pts) sub_pts[0] = 0;
    if (ins_pts) ins_pts[0] = 0;
    if (insC_pts) insC_pts[0] = 0;
    for (int i = 1; i < ORC_TABLE_LEN; i++) {
        int p = i > LIM4 ? P_INS4 : i > LIM3 ? P_INS3 : i > 1 ? P_INS2 : P_INS;
        int s = i > LIM3 ? P_SUB3 : i > 1 ? P_SUB2 : P_SUB;
        ic = imax(OFF(MINSCORE), OFF(p) + ic);
        icp = imax(MINSCORE, p + icp);
        if (ins_off) ins_off[i] = OFF(p);
        if (sub_off) sub_off[i] = OFF(s);
        if (insC_off) insC_off[i] = ic;
        if (ins_pts) ins_pts[i] = p;
        if (sub_pts) sub_pts[i] = s;
        if (insC_pts) insC_pts[i] = icp;
    }
}

void orc_base_to_number(int8_t* t /*128*/) {
    /* dna/AminoAcid.java:615-624 */
    memset(t, -1, 128);
    const char* b = "ACGT";
    for (int i = 0; i < 4; i++) { t[(int)b[i]] = (int8_t)i; t[(int)b[i] + 32] = (int8_t)i; }
    t['U'] = 3; t['u'] = 3;
}

static int32_t del_offset(int32_t len) {
    /* calcDelScoreOffset, jni/...JNI.c:316-336 */
    if (len <= 0) return 0;
    int32_t s = OFF(P_DEL);
    if (len > LIM5) { s += ((len - LIM5 + M5) / SLIP) * OFF(P_DEL5); len = LIM5; }
    if (len > LIM4) { s += (len - LIM4) * OFF(P_DEL4); len = LIM4; }
    if (len > LIM3) { s += (len - LIM3) * OFF(P_DEL3); len = LIM3; }
    if (len > 1) s += (len - 1) * OFF(P_DEL2);
    return s;
}
static inline int32_t ins_offset(int32_t len, const int32_t* insC) { return len <= 0 ? 0 : insC[len]; }

static inline int32_t clampt(int32_t t) { return t > MAXTIME ? MAXTIME - M5 : t; }

/* ---- port of fillUnlimited (jni/...JNI.c:100-314) ---- */
void orc_fill_unlimited(const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
                        int32_t refStartLoc, int32_t refEndLoc, int32_t* result, int64_t* iterations,
                        int32_t* packed, const int32_t* SUBA, const int32_t* INSA,
                        int32_t maxRows, int32_t maxColumns) {
    (void)ref_length;
    const int32_t rows = read_length, columns = refEndLoc - refStartLoc + 1;
    const int32_t maxGain = (read_length - 1) * OFF(P_MATCH2) + OFF(P_MATCH);
    const int32_t subfloor = 0 - 2 * maxGain;
    const int32_t barI2 = rows - BAR_I1, barI2b = columns - 1, barD2 = rows - BAR_D1;
    const int64_t stride = maxColumns + 1, plane = (int64_t)(maxRows + 1) * stride;
    int32_t* M = packed; int32_t* D = packed + plane; int32_t* I = packed + 2 * plane;
    if (rows > maxRows || columns > maxColumns) { result[0] = -1; return; } /* reference exit(0)s here, :130-132 */

    for (int32_t row = 1; row <= rows; row++) {
        const int64_t up = (int64_t)(row - 1) * stride, cur = (int64_t)row * stride;
        const int8_t c1 = read[row - 1], c0 = row < 2 ? (int8_t)'?' : read[row - 2];
        for (int32_t col = 1; col <= columns; col++) {
            (*iterations)++;
            const int8_t r1 = ref[refStartLoc + col - 1];
            const int8_t r0 = col < 2 ? (int8_t)'!' : ref[refStartLoc + col - 2];
            const int match = (c1 == r1 && r1 != 'N'), prevMatch = (c0 == r0 && r0 != 'N'), gap = (r1 == '-');
            if (gap) {
                M[cur + col] = subfloor;
            } else {
                const int32_t dm = M[up + col - 1], sM = dm & SMASK, streak = dm & TMASK;
                const int32_t sD = D[up + col - 1] & SMASK, sI = I[up + col - 1] & SMASK;
                int32_t a, o;
                if (match) { a = sM + (prevMatch ? OFF(P_MATCH2) : OFF(P_MATCH)); o = OFF(P_MATCH); }
                else {
                    a = sM + ((r1 != 'N' && c1 != 'N') ? (prevMatch ? (streak <= 1 ? OFF(P_SUBR) : OFF(P_SUB)) : SUBA[streak + 1]) : OFF(P_NOCALL));
                    o = OFF(P_SUB);
                }
                const int32_t b = sD + o, c = sI + o;
                int32_t score, time;
                if (a >= b && a >= c) { score = a; time = match ? (prevMatch ? streak + 1 : 1) : (prevMatch ? 1 : streak + 1); }
                else if (b >= c) { score = b; time = 1; }
                else { score = c; time = 1; }
                M[cur + col] = score | clampt(time);
            }
            if (row < BAR_D1 || row > barD2) {
                D[cur + col] = subfloor;
            } else {
                const int32_t ld = D[cur + col - 1], streak = ld & TMASK;
                int32_t a = (M[cur + col - 1] & SMASK) + OFF(P_DEL);
                int32_t b = (ld & SMASK) + (streak == 0 ? OFF(P_DEL) : streak < LIM3 ? OFF(P_DEL2) : streak < LIM4 ? OFF(P_DEL3) :
                                             streak < LIM5 ? OFF(P_DEL4) : ((streak & M5) == 0 ? OFF(P_DEL5) : 0));
                if (r1 == 'N') { a += OFF(P_DEL_REF_N); b += OFF(P_DEL_REF_N); }
                else if (gap) { a += OFF(P_GAP); b += OFF(P_GAP); }
                int32_t score, time;
                if (a >= b) { score = a; time = 1; } else { score = b; time = streak + 1; }
                D[cur + col] = score | clampt(time);
            }
            if (gap || (row < BAR_I1 && col > 1) || (row > barI2 && col < barI2b)) {
                I[cur + col] = subfloor;
            } else {
                const int32_t ui = I[up + col], streak = ui & TMASK;
                const int32_t a = (M[up + col] & SMASK) + OFF(P_INS);
                const int32_t b = (ui & SMASK) + INSA[streak + 1];
                int32_t score, time;
                if (a >= b) { score = a; time = 1; } else { score = b; time = streak + 1; }
                I[cur + col] = score | clampt(time);
            }
        }
    }
    int32_t maxCol = -1, maxState = -1, maxScore = INT_MIN;
    const int64_t last = (int64_t)rows * stride;
    for (int st = 0; st < 3; st++)
        for (int32_t col = 1; col <= columns; col++) {
            const int32_t x = packed[st * plane + last + col] & SMASK;
            if (x > maxScore) { maxScore = x; maxCol = col; maxState = st; }
        }
    result[0] = rows; result[1] = maxCol; result[2] = maxState; result[3] = maxScore >> TBITS;
}

/* optional per-row trace of the visited interval (analysis aid for kernel design; not thread-safe) */
static int32_t* g_row_trace = 0;
void orc_set_row_trace(int32_t* buf) { g_row_trace = buf; }

/* ---- port of fillLimitedX (jni/...JNI.c:361-704) ---- */
void orc_fill_limitedX(const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
                       int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* result, int64_t* iterations,
                       int32_t* packed, const int32_t* SUBA, const int32_t* INSA, int32_t maxRows, int32_t maxColumns,
                       int32_t bandwidth, float bandwidthRatio, int32_t* vertLimit, int32_t* horizLimit,
                       const int8_t* baseToNumber, const int32_t* INSC) {
    (void)ref_length;
    const int32_t rows = read_length, columns = refEndLoc - refStartLoc + 1;
    const int64_t stride = maxColumns + 1, plane = (int64_t)(maxRows + 1) * stride;
    int32_t* M = packed; int32_t* D = packed + plane; int32_t* I = packed + 2 * plane;

    const int32_t halfband = (bandwidth < 1 && bandwidthRatio <= 0) ? 0 :
        imax(imin(bandwidth < 1 ? 9999999 : bandwidth, bandwidthRatio <= 0 ? 9999999 : 8 + (int32_t)(rows * bandwidthRatio)), (columns - rows + 8)) / 2;
    const int32_t barI2 = rows - BAR_I1, barI2b = columns - 1, barD2 = rows - BAR_D1;

    const int64_t last = (int64_t)rows * stride;
    for (int st = 0; st < 3; st++) for (int32_t i = 1; i < columns + 1; i++) packed[st * plane + last + i] = BADOFF;

    int32_t minGoodCol = 1, maxGoodCol = columns;
    const int32_t minScore_off = (int32_t)((uint32_t)minScore << TBITS);
    const int32_t maxGain = (read_length - 1) * OFF(P_MATCH2) + OFF(P_MATCH);
    const int32_t floor_ = minScore_off - maxGain;
    const int32_t subfloor = floor_ - 5 * OFF(P_MATCH2);

    vertLimit[rows] = minScore_off;
    int prevDefined = 0;
    for (int32_t i = rows - 1; i >= 0; i--) {
        const int8_t c = read[i];
        if (baseToNumber[(int)c] >= 0) { vertLimit[i] = imax(vertLimit[i + 1] - (prevDefined ? OFF(P_MATCH2) : OFF(P_MATCH)), floor_); prevDefined = 1; }
        else { vertLimit[i] = imax(vertLimit[i + 1] - OFF(P_NOCALL), floor_); prevDefined = 0; }
    }
    horizLimit[columns] = minScore_off;
    prevDefined = 0;
    for (int32_t i = columns - 1; i >= 0; i--) {
        const int8_t c = ref[refStartLoc + i];
        if (baseToNumber[(int)c] >= 0) { horizLimit[i] = imax(horizLimit[i + 1] - (prevDefined ? OFF(P_MATCH2) : OFF(P_MATCH)), floor_); prevDefined = 1; }
        else { horizLimit[i] = imax(horizLimit[i + 1] - (prevDefined && c == '-' ? OFF(P_DEL) : OFF(P_NOREF)), floor_); prevDefined = 0; }
    }

    for (int32_t row = 1; row <= rows; row++) {
        const int32_t colStart = halfband < 1 ? minGoodCol : imax(minGoodCol, row - halfband);
        const int32_t colStop = halfband < 1 ? maxGoodCol : imin(maxGoodCol, row + halfband * 2 - 1);
        minGoodCol = -1; maxGoodCol = -2;
        const int32_t vlimit = vertLimit[row];
        if (colStart < 0 || colStop < colStart) break;
        const int64_t up = (int64_t)(row - 1) * stride, cur = (int64_t)row * stride;
        if (g_row_trace) { g_row_trace[2 * row] = colStart; g_row_trace[2 * row + 1] = colStart - 1; }
        if (colStart > 1) { M[cur + colStart - 1] = subfloor; I[cur + colStart - 1] = subfloor; D[cur + colStart - 1] = subfloor; }
        const int8_t c1 = read[row - 1], c0 = row < 2 ? (int8_t)'?' : read[row - 2];

        for (int32_t col = colStart; col <= columns; col++) {
            const int8_t r1 = ref[refStartLoc + col - 1];
            const int8_t r0 = col < 2 ? (int8_t)'!' : ref[refStartLoc + col - 2];
            const int gap = (r1 == '-'), match = (c1 == r1 && r1 != 'N'), prevMatch = (c0 == r0 && r0 != 'N');
            (*iterations)++;
            const int32_t limit = imax(vlimit, horizLimit[col]);
            const int32_t limit3 = imax(floor_, match ? limit - OFF(P_MATCH2) : limit - OFF(P_SUB3));
            const int32_t delNeeded = imax(0, row - col - 1);
            const int32_t insNeeded = imax(0, (rows - row) - (columns - col) - 1);
            const int32_t delPenalty = del_offset(delNeeded);
            const int32_t insPenalty = ins_offset(insNeeded, INSC);

            const int32_t dM = M[up + col - 1], dD = D[up + col - 1], dI = I[up + col - 1];
            const int32_t lM = M[cur + col - 1], lD = D[cur + col - 1];
            const int32_t uM = M[up + col], uI = I[up + col];

            /* MS */
            if (gap || ((dM & SMASK) <= limit3 && (dD & SMASK) <= limit3 && (dI & SMASK) <= limit3)) {
                M[cur + col] = subfloor;
            } else {
                const int32_t sM = dM & SMASK, streak = dM & TMASK;
                int32_t a, o;
                if (match) { a = sM + (prevMatch ? OFF(P_MATCH2) : OFF(P_MATCH)); o = OFF(P_MATCH); }
                else {
                    a = sM + ((r1 != 'N' && c1 != 'N') ? (prevMatch ? (streak <= 1 ? OFF(P_SUBR) : OFF(P_SUB)) : SUBA[streak + 1]) : OFF(P_NOCALL));
                    o = OFF(P_SUB);
                }
                const int32_t b = (dD & SMASK) + o, c = (dI & SMASK) + o;
                int32_t score, time;
                if (a >= b && a >= c) { score = a; time = match ? (prevMatch ? streak + 1 : 1) : (prevMatch ? 1 : streak + 1); }
                else if (b >= c) { score = b; time = 1; }
                else { score = c; time = 1; }
                const int32_t limit2 = delNeeded > 0 ? limit - delPenalty : insNeeded > 0 ? limit - insPenalty : limit;
                if (score >= limit2) { maxGoodCol = col; if (minGoodCol < 0) minGoodCol = col; } else score = subfloor;
                M[cur + col] = score | clampt(time);
            }
            /* DEL */
            if (((lM & SMASK) <= limit && (lD & SMASK) <= limit) || row < BAR_D1 || row > barD2) {
                D[cur + col] = subfloor;
            } else {
                const int32_t streak = lD & TMASK;
                int32_t a = (lM & SMASK) + OFF(P_DEL);
                int32_t b = (lD & SMASK) + (streak == 0 ? OFF(P_DEL) : streak < LIM3 ? OFF(P_DEL2) : streak < LIM4 ? OFF(P_DEL3) :
                                             streak < LIM5 ? OFF(P_DEL4) : ((streak & M5) == 0 ? OFF(P_DEL5) : 0));
                if (r1 == 'N') { a += OFF(P_DEL_REF_N); b += OFF(P_DEL_REF_N); }
                else if (gap) { a += OFF(P_GAP); b += OFF(P_GAP); }
                int32_t score, time;
                if (a >= b) { score = a; time = 1; } else { score = b; time = streak + 1; }
                const int32_t limit2 = insNeeded > 0 ? limit - insPenalty :
                                       delNeeded > 0 ? limit - del_offset(time + delNeeded) + del_offset(time) : limit;
                if (score >= limit2) { maxGoodCol = col; if (minGoodCol < 0) minGoodCol = col; } else score = subfloor;
                D[cur + col] = score | clampt(time);
            }
            /* INS */
            if (gap || ((uM & SMASK) <= limit && (uI & SMASK) <= limit) || (row < BAR_I1 && col > 1) || (row > barI2 && col < barI2b)) {
                I[cur + col] = subfloor;
            } else {
                const int32_t streak = uI & TMASK;
                const int32_t a = (uM & SMASK) + OFF(P_INS);
                const int32_t b = (uI & SMASK) + INSA[streak + 1];
                int32_t score, time;
                if (a >= b) { score = a; time = 1; } else { score = b; time = streak + 1; }
                const int32_t limit2 = delNeeded > 0 ? limit - delPenalty :
                                       insNeeded > 0 ? limit - ins_offset(time + insNeeded, INSC) + ins_offset(time, INSC) : limit;
                if (score >= limit2) { maxGoodCol = col; if (minGoodCol < 0) minGoodCol = col; } else score = subfloor;
                I[cur + col] = score | clampt(time);
            }
            if (g_row_trace) g_row_trace[2 * row + 1] = col;
            if (col >= colStop) {
                if (col > colStop && (maxGoodCol < col || halfband > 0)) break;
                if (row > 1) { M[up + col + 1] = subfloor; I[up + col + 1] = subfloor; D[up + col + 1] = subfloor; }
            }
        }
    }

    int32_t maxCol = -1, maxState = -1, maxScore = INT_MIN;
    for (int st = 0; st < 3; st++)
        for (int32_t col = 1; col <= columns; col++) {
            const int32_t x = packed[st * plane + last + col] & SMASK;
            if (x > maxScore) { maxScore = x; maxCol = col; maxState = st; }
        }
    result[0] = rows; result[1] = maxCol; result[2] = maxState;
    if (maxScore < minScore_off) { result[3] = maxScore; result[4] = 1; }
    else { result[3] = maxScore >> TBITS; result[4] = 0; }
}

/* =====================  Java half of the plug-in  ===================== */

struct orc_msa {
    int32_t maxRows, maxColumns;
    int32_t* packed; int32_t* vertLimit; int32_t* horizLimit; int8_t* grefbuffer;
    int32_t sub[ORC_TABLE_LEN], ins[ORC_TABLE_LEN], insC[ORC_TABLE_LEN];
    int8_t b2n[128];
    int32_t rows, columns;
    int32_t greflimit, greflimit2, grefRefOrigin;
    int32_t bandwidth; float bandwidthRatio;
    int64_t iterationsLimited, iterationsUnlimited;
    int32_t lastPath; /* 0 limited, 1 unlimited */
    orc_fill_limited_fn fillL; orc_fill_unlimited_fn fillU;
};

orc_msa* orc_msa_new(int32_t maxRows, int32_t maxColumns) {
    /* MultiStateAligner11tsJNI.java:71-113 */
    orc_msa* m = (orc_msa*)calloc(1, sizeof(orc_msa));
    m->maxRows = maxRows; m->maxColumns = maxColumns;
    const int64_t stride = maxColumns + 1, plane = (int64_t)(maxRows + 1) * stride;
    m->packed = (int32_t*)calloc((size_t)(3 * plane), sizeof(int32_t));
    m->grefbuffer = (int8_t*)calloc((size_t)maxColumns + 2, 1);
    m->vertLimit = (int32_t*)malloc(sizeof(int32_t) * ((size_t)maxRows + 1));
    m->horizLimit = (int32_t*)malloc(sizeof(int32_t) * ((size_t)maxColumns + 1));
    for (int i = 0; i <= maxRows; i++) m->vertLimit[i] = BADOFF;
    for (int i = 0; i <= maxColumns; i++) m->horizLimit[i] = BADOFF;
    orc_msa_tables(m->sub, m->ins, m->insC, NULL, NULL, NULL);
    orc_base_to_number(m->b2n);
    for (int st = 0; st < 3; st++) {
        int32_t* P = m->packed + st * plane;
        for (int i = 1; i <= maxRows; i++) for (int j = 0; j <= maxColumns; j++) P[i * stride + j] |= BADOFF;
        for (int i = 0; i <= maxRows; i++) {
            const int32_t prev = i < 2 ? 0 : P[(int64_t)(i - 1) * stride];
            P[(int64_t)i * stride] = prev + m->ins[i < ORC_TABLE_LEN ? i : ORC_TABLE_LEN - 1];
        }
    }
    m->fillL = orc_fill_limitedX; m->fillU = orc_fill_unlimited;
    return m;
}
void orc_msa_free(orc_msa* m) { if (!m) return; free(m->packed); free(m->grefbuffer); free(m->vertLimit); free(m->horizLimit); free(m); }
void orc_msa_set_backend(orc_msa* m, orc_fill_limited_fn l, orc_fill_unlimited_fn u) { if (l) m->fillL = l; if (u) m->fillU = u; }
void orc_msa_set_band(orc_msa* m, int32_t bandwidth, float ratio) { m->bandwidth = bandwidth; m->bandwidthRatio = ratio; }
void orc_msa_set_shape(orc_msa* m, int32_t rows, int32_t columns) { m->rows = rows; m->columns = columns; }
int32_t* orc_msa_packed(orc_msa* m) { return m->packed; }
int64_t orc_msa_iterations(const orc_msa* m, int which) { return which ? m->iterationsUnlimited : m->iterationsLimited; }
int32_t orc_msa_last_path(const orc_msa* m) { return m->lastPath; }
int32_t orc_msa_greflimit(const orc_msa* m) { return m->greflimit; }
const int8_t* orc_msa_gref(const orc_msa* m) { return m->grefbuffer; }

static int fill_unlimited_x(orc_msa* m, const int8_t* read, int32_t rlen, const int8_t* ref, int32_t reflen, int32_t a, int32_t b, int32_t* max4) {
    int32_t res[4] = {0, 0, 0, 0};
    m->rows = rlen; m->columns = b - a + 1;       /* the C side recomputes these; score2 uses the Java fields */
    m->lastPath = 1;
    m->fillU(read, ref, rlen, reflen, a, b, res, &m->iterationsUnlimited, m->packed, m->sub, m->ins, m->maxRows, m->maxColumns);
    memcpy(max4, res, sizeof(res));
    return 1;
}

static int fill_limited_x(orc_msa* m, const int8_t* read, int32_t rlen, const int8_t* ref, int32_t reflen, int32_t a, int32_t b, int32_t minScore, int32_t* max4) {
    /* MultiStateAligner11tsJNI.java:132-164 */
    m->rows = rlen; m->columns = b - a + 1;
    const int32_t rows = m->rows, columns = m->columns, bandwidth = m->bandwidth; const float bwr = m->bandwidthRatio;
    const int32_t halfband = (bandwidth < 1 && bwr <= 0) ? 0 :
        imax(imin(bandwidth < 1 ? 9999999 : bandwidth, bwr <= 0 ? 9999999 : 8 + (int32_t)(rows * bwr)), (columns - rows + 8)) / 2;
    if (minScore < 1 || (columns + rows < 90) || ((halfband < 1 || halfband * 3 > columns) && (columns > rlen + imin(170, rlen + 20))))
        return fill_unlimited_x(m, read, rlen, ref, reflen, a, b, max4);
    minScore -= 120;
    int32_t res[5] = {0, 0, 0, 0, 0};
    m->lastPath = 0;
    m->fillL(read, ref, rlen, reflen, a, b, minScore, res, &m->iterationsLimited, m->packed, m->sub, m->ins, m->maxRows, m->maxColumns,
             bandwidth, bwr, m->vertLimit, m->horizLimit, m->b2n, m->insC);
    if (res[4] == 1) return 0;
    memcpy(max4, res, 4 * sizeof(int32_t));
    return 1;
}

static int8_t* make_gref(orc_msa* m, const int8_t* ref, int32_t reflen, int32_t* gaps, int32_t ngaps, int32_t refStartLoc, int32_t refEndLoc) {
    /* MultiStateAligner11tsJNI.java:668-757 */
    const int32_t g0 = gaps[0], gN = gaps[ngaps - 1];
    gaps[0] = imin(gaps[0], refStartLoc);
    gaps[ngaps - 1] = imax(gN, refEndLoc);
    m->grefRefOrigin = gaps[0];
    int8_t* gref = m->grefbuffer;
    const int32_t greflen = m->maxColumns + 2;
    int32_t gpos = 0;
    for (int32_t i = 0; i < ngaps; i += 2) {
        const int32_t x = gaps[i], y = gaps[i + 1];
        for (int32_t r = x; r <= y; r++, gpos++) { if (gpos >= greflen) goto overflow; gref[gpos] = ref[r]; }
        if (i + 2 < ngaps) {
            const int32_t z = gaps[i + 2], gap = z - y - 1;
            const int32_t rem = gap % GAPLEN_, lim = y + GAPBUFFER_ + rem, div = (gap - GAPBUFFER2_) / GAPLEN_;
            for (int32_t r = y + 1; r <= lim; r++, gpos++) { if (gpos >= greflen) goto overflow; gref[gpos] = ref[r]; }
            for (int32_t g = 0; g < div; g++, gpos++) { if (gpos >= greflen) goto overflow; gref[gpos] = '-'; }
            for (int32_t r = z - GAPBUFFER_; r < z; r++, gpos++) { if (gpos >= greflen) goto overflow; gref[gpos] = ref[r]; }
        }
    }
    m->greflimit = gpos;
    {
        const int32_t lim = imin(greflen, m->greflimit + CUSHION_);
        for (int32_t i = m->greflimit, r = refEndLoc + 1; i < lim; i++, r++) { gref[i] = (r < reflen ? ref[r] : (int8_t)'N'); m->greflimit2 = i; }
    }
    gaps[0] = g0; gaps[ngaps - 1] = gN;
    return gref;
overflow:
    gaps[0] = g0; gaps[ngaps - 1] = gN;
    return NULL;
}

static int32_t from_gapped(const orc_msa* m, int32_t point) {
    /* :759-779 */
    if (point <= 0) return m->grefRefOrigin + point;
    for (int32_t i = 0, j = m->grefRefOrigin; i < m->greflimit2; i++) {
        if (i == point) return j;
        j += (m->grefbuffer[i] == '-' ? GAPLEN_ : 1);
    }
    return INT_MIN;
}
static int32_t to_gapped(const orc_msa* m, int32_t point) {
    /* :781-801 */
    if (point <= m->grefRefOrigin) return point - m->grefRefOrigin;
    for (int32_t i = 0, j = m->grefRefOrigin; i < m->greflimit2; i++) {
        if (j == point) return i;
        j += (m->grefbuffer[i] == '-' ? GAPLEN_ : 1);
    }
    return INT_MIN;
}
int32_t orc_msa_to_gapped(const orc_msa* m, int32_t p) { return to_gapped(m, p); }
int32_t orc_msa_from_gapped(const orc_msa* m, int32_t p) { return from_gapped(m, p); }

int orc_msa_fillLimited(orc_msa* m, const int8_t* read, int32_t rlen, const int8_t* ref, int32_t reflen,
                        int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* gaps, int32_t ngaps, int32_t* max4) {
    /* :116-130 */
    if (!gaps || ngaps <= 0) return fill_limited_x(m, read, rlen, ref, reflen, refStartLoc, refEndLoc, minScore, max4);
    int8_t* gref = make_gref(m, ref, reflen, gaps, ngaps, refStartLoc, refEndLoc);
    if (!gref) return -1;
    return fill_limited_x(m, read, rlen, gref, m->maxColumns + 2, 0, m->greflimit, minScore, max4);
}

int orc_msa_fillUnlimited(orc_msa* m, const int8_t* read, int32_t rlen, const int8_t* ref, int32_t reflen,
                          int32_t refStartLoc, int32_t refEndLoc, int32_t* gaps, int32_t ngaps, int32_t* max4) {
    /* :166-175 */
    if (!gaps || ngaps <= 0) return fill_unlimited_x(m, read, rlen, ref, reflen, refStartLoc, refEndLoc, max4);
    int8_t* gref = make_gref(m, ref, reflen, gaps, ngaps, refStartLoc, refEndLoc);
    if (!gref) return -1;
    return fill_unlimited_x(m, read, rlen, gref, m->maxColumns + 2, 0, m->greflimit, max4);
}

/* predecessor rule shared by score2 / traceback2 (:391-447, :573-611) */
static inline int prev_state(const orc_msa* m, int state, int32_t row, int32_t col, int32_t* time_out) {
    const int64_t stride = m->maxColumns + 1, plane = (int64_t)(m->maxRows + 1) * stride;
    const int32_t* M = m->packed; const int32_t* D = M + plane; const int32_t* I = M + 2 * plane;
    const int32_t time = m->packed[state * plane + (int64_t)row * stride + col] & TMASK;
    *time_out = time;
    if (time > 1) return state;
    if (state == ST_MS) {
        const int64_t d = (int64_t)(row - 1) * stride + col - 1;
        const int32_t a = M[d] & SMASK, b = D[d] & SMASK, c = I[d] & SMASK;
        return (a >= b && a >= c) ? ST_MS : (b >= c ? ST_DEL : ST_INS);
    } else if (state == ST_DEL) {
        const int64_t l = (int64_t)row * stride + col - 1;
        return (M[l] & SMASK) >= (D[l] & SMASK) ? ST_MS : ST_DEL;
    } else {
        const int64_t u = (int64_t)(row - 1) * stride + col;
        return (M[u] & SMASK) >= (I[u] & SMASK) ? ST_MS : ST_INS;
    }
}

int orc_msa_score2(orc_msa* m, const int8_t* read, const int8_t* ref, int32_t refStartLoc, int32_t refEndLoc,
                   int32_t maxRow, int32_t maxCol, int32_t maxState, int32_t* out8) {
    /* :537-658 */
    (void)read; (void)ref;
    const int64_t stride = m->maxColumns + 1, plane = (int64_t)(m->maxRows + 1) * stride;
    int32_t row = maxRow, col = maxCol, state = maxState;
    int32_t score = m->packed[maxState * plane + (int64_t)maxRow * stride + maxCol] & SMASK;
    if (row < m->rows) {
        int32_t difR = m->rows - row, difC = m->columns - col;
        while (difR > difC) { score += OFF(P_NOREF); difR--; }
        row += difR; col += difR;
    }
    const int32_t bestRefStop = refStartLoc + col - 1;
    int32_t stateTime = 0;
    while (row > 0 && col > 0) {
        int32_t time;
        const int prev = prev_state(m, state, row, col, &time);
        if (state == ST_MS) { row--; col--; } else if (state == ST_DEL) { col--; } else { row--; }
        if (col < 0) break;
        if (state == prev) stateTime++; else stateTime = 0;
        state = prev;
    }
    if (row > col) col -= row;
    const int32_t bestRefStart = refStartLoc + col;
    score >>= TBITS;
    int32_t padLeft = 0, padRight = 0;
    if (bestRefStart < refStartLoc) padLeft = imax(0, refStartLoc - bestRefStart);
    else if (bestRefStart == refStartLoc && state == ST_INS) padLeft = stateTime;
    if (bestRefStop > refEndLoc) padRight = imax(0, bestRefStop - refEndLoc);
    else if (bestRefStop == refEndLoc && maxState == ST_INS) padRight = m->packed[maxState * plane + (int64_t)maxRow * stride + maxCol] & TMASK;
    out8[0] = score; out8[1] = bestRefStart; out8[2] = bestRefStop; out8[3] = maxRow; out8[4] = maxCol; out8[5] = maxState;
    out8[6] = padLeft; out8[7] = padRight;
    return (padLeft > 0 || padRight > 0) ? 8 : 6;
}

int orc_msa_score(orc_msa* m, const int8_t* read, const int8_t* ref, int32_t refStartLoc, int32_t refEndLoc,
                  int32_t maxRow, int32_t maxCol, int32_t maxState, int gapped, int32_t* out8) {
    /* :499-535 */
    if (!gapped) return orc_msa_score2(m, read, ref, refStartLoc, refEndLoc, maxRow, maxCol, maxState, out8);
    const int32_t gstart = to_gapped(m, refStartLoc), gstop = to_gapped(m, refEndLoc);
    const int n = orc_msa_score2(m, read, m->grefbuffer, gstart, gstop, maxRow, maxCol, maxState, out8);
    out8[1] = from_gapped(m, out8[1]);
    out8[2] = from_gapped(m, out8[2]);
    return n;
}

int32_t orc_msa_traceback2(orc_msa* m, const int8_t* read, const int8_t* ref, int32_t refStartLoc, int32_t refEndLoc,
                           int32_t row, int32_t col, int32_t state, int8_t* out, int32_t outcap) {
    /* :376-495.  Returns the match-string length (or -1 if outcap is too small). */
    (void)refEndLoc;
    const int32_t cap0 = row + col; /* the Java allocates row+col-1 */
    int8_t* tmp = (int8_t*)malloc((size_t)cap0 + 1);
    int32_t n = 0, gaps = 0;
    while (row > 0 && col > 0) {
        int32_t time;
        const int prev = prev_state(m, state, row, col, &time);
        if (state == ST_MS) {
            const int8_t c = read[row - 1], r = ref[refStartLoc + col - 1];
            if (c == r) tmp[n] = 'm';
            else if (!(c >= 0 && m->b2n[(int)c] >= 0) || !(r >= 0 && m->b2n[(int)r] >= 0)) tmp[n] = 'N';
            else tmp[n] = 'S';
            row--; col--;
        } else if (state == ST_DEL) {
            const int8_t r = ref[refStartLoc + col - 1];
            if (r == '-') { tmp[n] = '-'; gaps++; } else tmp[n] = 'D';
            col--;
        } else {
            tmp[n] = col == 0 ? 'X' : (col >= m->columns ? 'Y' : 'I');
            row--;
        }
        state = prev; n++;
    }
    if (col != row) { while (row > 0) { tmp[n++] = 'X'; row--; col--; } }
    const int32_t total = n + gaps * (GAPLEN_ - 1);
    if (total > outcap) { free(tmp); return -1; }
    int32_t j = 0;
    for (int32_t i = n - 1; i >= 0; i--) {
        if (tmp[i] != '-') out[j++] = tmp[i];
        else { for (int k = 0; k < GAPLEN_; k++) out[j++] = 'D'; }
    }
    free(tmp);
    return total;
}

int32_t orc_msa_traceback(orc_msa* m, const int8_t* read, const int8_t* ref, int32_t refStartLoc, int32_t refEndLoc,
                          int32_t row, int32_t col, int32_t state, int gapped, int8_t* out, int32_t outcap) {
    /* :362-373 */
    if (!gapped) return orc_msa_traceback2(m, read, ref, refStartLoc, refEndLoc, row, col, state, out, outcap);
    const int32_t gstart = to_gapped(m, refStartLoc), gstop = to_gapped(m, refEndLoc);
    return orc_msa_traceback2(m, read, m->grefbuffer, gstart, gstop, row, col, state, out, outcap);
}

int orc_msa_fillAndScoreLimited(orc_msa* m, const int8_t* read, int32_t rlen, const int8_t* ref, int32_t reflen,
                                int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* gaps, int32_t ngaps,
                                int32_t* max4, int32_t* out8) {
    /* MSA.java:103-134.  Returns 0 for null, else 6 or 8. */
    int32_t a = imax(0, refStartLoc), b = imin(reflen - 1, refEndLoc);
    const int gapped = (gaps && ngaps > 0);
    if (!gapped && b - a >= m->maxColumns) b = imin(reflen - 1, a + m->maxColumns - 1);
    const int ok = orc_msa_fillLimited(m, read, rlen, ref, reflen, a, b, minScore, gaps, ngaps, max4);
    if (ok <= 0) return 0;
    return orc_msa_score(m, read, ref, a, b, max4[0], max4[1], max4[2], gapped, out8);
}
