/*
 * banded_oracle.c — TEST INFRASTRUCTURE ONLY.
 * Restatement of the reference's banded edit-distance aligner, jni/BandedAlignerJNI.c:97-585 (alignForward,
 * alignForwardRC, alignReverse, alignReverseRC, lastOffsetFunc, penalizeOffCenterFunc), as ONE routine parametrised by
 * the walking directions.  Parity target is the JNI C, not BandedAlignerConcrete.java (they differ: SURVEY §8c).
 * Validated against the reference's own C in oracle/_ref (tests/test_banded_oracle.py) and the Appendix-B KATs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "banded_oracle.h"

#define BIG 999
static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

static void base_tables(int8_t* b2n, int8_t* comp) {
    /* dna/AminoAcid.java:615-624 and :650-664 */
    memset(b2n, -1, 128); memset(comp, -1, 128);
    const char* b = "ACGT";
    for (int i = 0; i < 4; i++) { b2n[(int)b[i]] = (int8_t)i; b2n[(int)b[i] + 32] = (int8_t)i; }
    b2n['U'] = 3; b2n['u'] = 3;
    const char* ext = " ACMGRSVTWYHKDBNX";
    const char* cex = " TGKCYWBASRDMHVNX";
    for (int i = 0; ext[i]; i++) {
        const int x = ext[i], y = cex[i];
        comp[x] = (int8_t)y;
        const int xl = (x >= 'A' && x <= 'Z') ? x + 32 : x, yl = (y >= 'A' && y <= 'Z') ? y + 32 : y;
        comp[xl] = (int8_t)yl;
    }
    comp['U'] = 'A'; comp['u'] = 'a'; comp['?'] = '?'; comp[' '] = ' '; comp['-'] = '-'; comp['*'] = '*'; comp['.'] = '.';
}
void orc_banded_tables(int8_t* b2n, int8_t* comp) { base_tables(b2n, comp); }

static int penalize_off_center(int* arr, int halfWidth) {
    const int center = halfWidth + 1;
    int edits = arr[center];
    for (int i = 1; i <= halfWidth; i++) {
        arr[center + i] = imin(BIG, arr[center + i] + i); edits = imin(edits, arr[center + i]);
        arr[center - i] = imin(BIG, arr[center - i] + i); edits = imin(edits, arr[center - i]);
    }
    return edits;
}
static int last_offset(const int* arr, int halfWidth) {
    const int center = halfWidth + 1;
    int minLoc = center;
    for (int i = 1; i <= halfWidth; i++) {
        if (arr[center + i] < arr[minLoc]) minLoc = center + i;
        if (arr[center - i] < arr[minLoc]) minLoc = center - i;
    }
    return center - minLoc;
}

/* dir: 0 alignForward, 1 alignForwardRC, 2 alignReverse, 3 alignReverseRC */
int orc_banded_align(int dir, const int8_t* query, const int8_t* ref, int qlen, int rlen, int qstart, int rstart, int maxEdits,
                     int exact, int maxWidth, int32_t* rv /* lastQueryLoc,lastRefLoc,lastRow,lastEdits,lastOffset */) {
    static int8_t b2n[128], comp[128]; static int init = 0;
    if (!init) { base_tables(b2n, comp); init = 1; }
    /* the swap rules at the top of each variant (:141-148, :260-267, :375-382, :490-497) */
    int swap = 0, dir2 = dir;
    if (dir == 0) swap = (qlen - qstart > rlen - rstart);
    else if (dir == 1) { swap = (qstart + 1 > rlen - rstart); dir2 = 3; }
    else if (dir == 2) swap = (qstart > rstart);
    else { swap = (qlen - qstart > rstart + 1); dir2 = 1; }
    if (swap) {
        const int x = orc_banded_align(dir2, ref, query, rlen, qlen, rstart, qstart, maxEdits, exact, maxWidth, rv);
        const int t = rv[0]; rv[0] = rv[1]; rv[1] = t;
        return x;
    }
    const int rc = (dir == 1 || dir == 3);
    const int qstep = (dir == 0 || dir == 3) ? 1 : -1;
    const int rfwd = (dir == 0 || dir == 1);
    int edits = 0, row = 0;
    rv[2] = -1; rv[3] = 0; rv[4] = 0;
    const int width = imin(maxWidth, maxEdits * 2 + 1), halfWidth = width / 2, inexact = !exact;
    int qloc = qstart, rsloc = rstart - halfWidth;
    const int xlines = (dir == 0 || dir == 3) ? qlen - qstart : qstart + 1;
    const int ylines = rfwd ? rlen - rstart : rstart + 1;
    const int len = imin(xlines, ylines);
    if (len < 1) return 0;
    int* bufA = (int*)malloc(sizeof(int) * (maxWidth + 2)); int* bufB = (int*)malloc(sizeof(int) * (maxWidth + 2));
    int *cur = bufA, *prev = bufB;
    for (int i = 0; i < maxWidth + 2; i++) { cur[i] = BIG; prev[i] = BIG; }
    for (row = 0; row < len; row++) {
        if (row > 0) { int* t = cur; cur = prev; prev = t; for (int i = 0; i < maxWidth + 2; i++) cur[i] = BIG; }
        const int8_t q = rc ? comp[(int)query[qloc]] : query[qloc];
        const int colStart = imax(0, rsloc), colLimit = imin(rsloc + width, rlen);
        edits = BIG;
        const int forceDiag = (row > 0 && row == len - 1);
        int mloc = rfwd ? 1 + (colStart - rsloc) : 1 + width - (colLimit - rsloc);
        for (int k = 0; k < colLimit - colStart; k++, mloc++) {
            const int col = rfwd ? colStart + k : colLimit - 1 - k;
            const int8_t r = ref[col];
            const int sub = (q == r || (inexact && (!(b2n[(int)q] >= 0) || !(b2n[(int)r] >= 0)))) ? 0 : 1;
            int score;
            if (row == 0) score = sub;
            else {
                const int up = prev[mloc + 1] + 1, diag = prev[mloc] + sub, left = cur[mloc - 1] + 1;
                const int edge = rfwd ? (col == rlen - 1) : (col == 0);
                score = (forceDiag || edge) ? diag : imin(up, imin(diag, left));
            }
            cur[mloc] = score;
            edits = imin(edits, score);
        }
        if (row == 0) edits = penalize_off_center(cur, halfWidth);
        else if (edits > maxEdits) { row++; break; }          /* the for-increment (qloc, rsloc) is skipped on break */
        qloc += qstep; rsloc += rfwd ? 1 : -1;
    }
    edits = penalize_off_center(cur, halfWidth);
    rv[2] = row - 1; rv[3] = edits; rv[4] = last_offset(cur, halfWidth);
    if (dir == 0) { rv[0] = qloc - 1; rv[1] = rsloc + halfWidth - rv[4] - 1; while (rv[1] >= rlen || rv[0] >= qlen) { rv[1]--; rv[0]--; } }
    else if (dir == 1) { rv[0] = qloc + 1; rv[1] = rsloc + halfWidth - rv[4] - 1; while (rv[1] >= rlen || rv[0] < 0) { rv[1]--; rv[0]++; } }
    else if (dir == 2) { rv[0] = qloc + 1; rv[1] = rsloc + halfWidth + rv[4] + 1; while (rv[1] < 0 || rv[0] < 0) { rv[1]++; rv[0]++; } }
    else { rv[0] = qloc - 1; rv[1] = rsloc + halfWidth + rv[4] + 1; while (rv[1] < 0 || rv[0] >= qlen) { rv[1]++; rv[0]--; } }
    free(bufA); free(bufB);
    return edits;
}

/* ---- batch driver (same record layout as include/bbmap_cuda.h: bbm_band_task / bbm_band_out) ---- */
typedef int (*ref_fwd_fn)(int8_t*, int8_t*, int, int, int, int, int, uint8_t, int*, int*, int*, int*, int*, int, int8_t*);
typedef int (*ref_rc_fn)(int8_t*, int8_t*, int, int, int, int, int, uint8_t, int*, int*, int*, int*, int*, int, int8_t*, int8_t*);
static void* g_ref_fns[4] = {0, 0, 0, 0};
void orc_banded_set_reference_fns(void* f, void* frc, void* r, void* rrc) { g_ref_fns[0] = f; g_ref_fns[1] = frc; g_ref_fns[2] = r; g_ref_fns[3] = rrc; }

typedef struct { const int8_t* q; const int8_t* r; const orc_band_task* t; orc_band_out* o; int64_t n; int use_ref; int tid, nth; } bjob;
static void* bworker(void* arg) {
    bjob* J = (bjob*)arg;
    int8_t b2n[128], comp[128]; base_tables(b2n, comp);
    const int64_t lo = J->n * J->tid / J->nth, hi = J->n * (J->tid + 1) / J->nth;
    for (int64_t i = lo; i < hi; i++) {
        const orc_band_task* T = &J->t[i]; orc_band_out* O = &J->o[i];
        int8_t* q = (int8_t*)J->q + T->query_off; int8_t* r = (int8_t*)J->r + T->ref_off;
        int32_t rv[5] = {0, 0, 0, 0, 0}; int e;
        if (J->use_ref) {
            if (T->dir == 0 || T->dir == 2) e = ((ref_fwd_fn)g_ref_fns[T->dir])(q, r, T->query_len, T->ref_len, T->qstart, T->rstart, T->max_edits, (uint8_t)T->exact, &rv[0], &rv[1], &rv[2], &rv[3], &rv[4], T->max_width, b2n);
            else e = ((ref_rc_fn)g_ref_fns[T->dir])(q, r, T->query_len, T->ref_len, T->qstart, T->rstart, T->max_edits, (uint8_t)T->exact, &rv[0], &rv[1], &rv[2], &rv[3], &rv[4], T->max_width, b2n, comp);
        } else e = orc_banded_align(T->dir, q, r, T->query_len, T->ref_len, T->qstart, T->rstart, T->max_edits, T->exact, T->max_width, rv);
        O->edits = e; memcpy(O->rv, rv, sizeof(rv)); O->status = 0; O->pad_ = 0;
    }
    return 0;
}
int orc_banded_batch(const int8_t* queries, const int8_t* refs, const orc_band_task* tasks, orc_band_out* outs, int64_t n, int use_reference, int threads) {
    if (use_reference && !g_ref_fns[0]) return -1;
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256]; bjob jobs[256];
    for (int i = 0; i < threads; i++) { bjob j = { queries, refs, tasks, outs, n, use_reference, i, threads }; jobs[i] = j; pthread_create(&th[i], 0, bworker, &jobs[i]); }
    for (int i = 0; i < threads; i++) pthread_join(th[i], 0);
    return 0;
}
