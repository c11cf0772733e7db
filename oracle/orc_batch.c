/*
 * orc_batch.c — TEST INFRASTRUCTURE ONLY.
 * Batch driver for the CPU oracle: for every task runs the same call sequence the
 * reference's host performs around one alignment (MSA.fillAndScoreLimited,
 * current/align2/MSA.java:103-134; then msa.traceback, BBMapThread.java:341-356),
 * one private aligner per host thread exactly like the reference's one-MSA-per-
 * mapping-thread model (AbstractMapThread.java:133-136).  Used (a) as the checker
 * in tests/ and smoke(), (b) as bench.py's cpu_baseline / --impl reference arm.
 * The fill back-end is either the port (msa_oracle.c) or the reference's own C
 * from oracle/_ref/libbbref.so (pointers supplied by orc_set_reference_fns).
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include "msa_oracle.h"

#define TF_RAW_LIMITED   1
#define TF_RAW_UNLIMITED 2
#define TF_CLAMP         4
#define TF_SCORE         8
#define TF_TRACEBACK     16

static orc_fill_limited_fn g_refL = 0;
static orc_fill_unlimited_fn g_refU = 0;
void orc_set_reference_fns(void* fillLimitedX, void* fillUnlimited) {
    g_refL = (orc_fill_limited_fn)fillLimitedX; g_refU = (orc_fill_unlimited_fn)fillUnlimited;
}

typedef struct {
    const int8_t* reads; const int8_t* refs; const orc_task* tasks; orc_out* outs; int64_t ntasks;
    int8_t* match_buf; const int64_t* match_off; int32_t bandwidth; float ratio; int32_t maxRows, maxColumns;
    int use_ref; int tid, nthreads; int64_t cells;
} job_t;

static int imax_(int a, int b) { return a > b ? a : b; }
static int imin_(int a, int b) { return a < b ? a : b; }

static void* worker(void* arg) {
    job_t* J = (job_t*)arg;
    orc_msa* m = orc_msa_new(J->maxRows, J->maxColumns);
    orc_msa_set_band(m, J->bandwidth, J->ratio);
    if (J->use_ref) orc_msa_set_backend(m, g_refL, g_refU);
    /* raw-mode scratch (the tables/limits a JNI caller would own) */
    int32_t sub[ORC_TABLE_LEN], ins[ORC_TABLE_LEN], insC[ORC_TABLE_LEN]; int8_t b2n[128];
    orc_msa_tables(sub, ins, insC, 0, 0, 0); orc_base_to_number(b2n);
    int32_t* vl = (int32_t*)malloc(sizeof(int32_t) * (J->maxRows + 1));
    int32_t* hl = (int32_t*)malloc(sizeof(int32_t) * (J->maxColumns + 1));
    orc_fill_limited_fn fL = J->use_ref ? g_refL : orc_fill_limitedX;
    orc_fill_unlimited_fn fU = J->use_ref ? g_refU : orc_fill_unlimited;
    int64_t cells = 0;
    /* contiguous chunks per thread: deterministic, and results do not depend on call order */
    const int64_t lo = J->ntasks * J->tid / J->nthreads, hi = J->ntasks * (J->tid + 1) / J->nthreads;
    for (int64_t t = lo; t < hi; t++) {
        const orc_task* T = &J->tasks[t]; orc_out* O = &J->outs[t];
        memset(O, 0, sizeof(*O)); O->match_len = -1;
        const int8_t* read = J->reads + T->read_off; const int8_t* ref = J->refs + T->ref_off;
        int32_t a = T->ref_start, b = T->ref_end;
        if (T->flags & TF_CLAMP) { a = imax_(0, a); b = imin_(T->ref_len - 1, b); }
        const int32_t rows = T->read_len, cols = b - a + 1;
        if (rows < 1 || cols < 1 || rows > J->maxRows || cols > J->maxColumns) { O->status = -2; continue; }
        int32_t max4[4] = {0, 0, 0, 0}; int ok = 1;
        const int64_t itL0 = orc_msa_iterations(m, 0), itU0 = orc_msa_iterations(m, 1);
        if (T->flags & TF_RAW_UNLIMITED) {
            int64_t it = 0; int32_t r4[4];
            fU(read, ref, rows, T->ref_len, a, b, r4, &it, orc_msa_packed(m), sub, ins, J->maxRows, J->maxColumns);
            memcpy(O->result, r4, sizeof(r4)); O->result[4] = 0; O->path = 1; O->iterations = it; memcpy(max4, r4, sizeof(r4));
        } else if (T->flags & TF_RAW_LIMITED) {
            int64_t it = 0; int32_t r5[5];
            fL(read, ref, rows, T->ref_len, a, b, T->min_score, r5, &it, orc_msa_packed(m), sub, ins, J->maxRows, J->maxColumns,
               J->bandwidth, J->ratio, vl, hl, b2n, insC);
            memcpy(O->result, r5, sizeof(r5)); O->path = 0; O->iterations = it; memcpy(max4, r5, 4 * sizeof(int32_t));
            ok = (r5[4] == 0);
        } else {
            ok = orc_msa_fillLimited(m, read, rows, ref, T->ref_len, a, b, T->min_score, 0, 0, max4);
            O->path = orc_msa_last_path(m);
            O->iterations = (orc_msa_iterations(m, 0) - itL0) + (orc_msa_iterations(m, 1) - itU0);
            if (ok > 0) { memcpy(O->result, max4, sizeof(max4)); O->result[4] = 0; }
            else { O->result[0] = rows; O->result[4] = 1; ok = 0; }
        }
        cells += O->iterations;
        if (ok && (T->flags & TF_SCORE)) {
            orc_msa_set_shape(m, rows, cols);
            O->score_len = orc_msa_score2(m, read, ref, a, b, max4[0], max4[1], max4[2], O->score);
        }
        if (ok && (T->flags & TF_TRACEBACK) && J->match_buf) {
            orc_msa_set_shape(m, rows, cols);
            const int64_t off = J->match_off[t], cap = J->match_off[t + 1] - off;
            O->match_len = orc_msa_traceback2(m, read, ref, a, b, max4[0], max4[1], max4[2], J->match_buf + off, (int32_t)cap);
            if (O->match_len < 0) O->status = -3;
        }
    }
    free(vl); free(hl); orc_msa_free(m);
    J->cells = cells;
    return 0;
}

int64_t orc_batch_run(const int8_t* reads, const int8_t* refs, const orc_task* tasks, orc_out* outs, int64_t ntasks,
                      int8_t* match_buf, const int64_t* match_off, int32_t bandwidth, float bandwidthRatio,
                      int32_t maxRows, int32_t maxColumns, int use_reference_fill, int threads) {
    if (use_reference_fill && (!g_refL || !g_refU)) return -1;
    if (threads < 1) threads = 1;
    if (threads > 512) threads = 512;
    pthread_t th[512]; job_t jobs[512];
    for (int i = 0; i < threads; i++) {
        job_t j = { reads, refs, tasks, outs, ntasks, match_buf, match_off, bandwidth, bandwidthRatio, maxRows, maxColumns,
                    use_reference_fill, i, threads, 0 };
        jobs[i] = j;
        pthread_create(&th[i], 0, worker, &jobs[i]);
    }
    int64_t cells = 0;
    for (int i = 0; i < threads; i++) { pthread_join(th[i], 0); cells += jobs[i].cells; }
    return cells;
}
