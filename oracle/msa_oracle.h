/* TEST INFRASTRUCTURE ONLY — declarations for the CPU oracle (see msa_oracle.c). */
#ifndef MSA_ORACLE_H
#define MSA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
#define ORC_TABLE_LEN 604

typedef void (*orc_fill_unlimited_fn)(const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
    int32_t refStartLoc, int32_t refEndLoc, int32_t* result, int64_t* iterations, int32_t* packed,
    const int32_t* SUBA, const int32_t* INSA, int32_t maxRows, int32_t maxColumns);
typedef void (*orc_fill_limited_fn)(const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
    int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* result, int64_t* iterations, int32_t* packed,
    const int32_t* SUBA, const int32_t* INSA, int32_t maxRows, int32_t maxColumns, int32_t bandwidth, float bandwidthRatio,
    int32_t* vertLimit, int32_t* horizLimit, const int8_t* baseToNumber, const int32_t* INSC);

void orc_msa_tables(int32_t* sub_off, int32_t* ins_off, int32_t* insC_off, int32_t* sub_pts, int32_t* ins_pts, int32_t* insC_pts);
void orc_base_to_number(int8_t* t128);

void orc_fill_unlimited(const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
    int32_t refStartLoc, int32_t refEndLoc, int32_t* result, int64_t* iterations, int32_t* packed,
    const int32_t* SUBA, const int32_t* INSA, int32_t maxRows, int32_t maxColumns);
void orc_fill_limitedX(const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
    int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* result, int64_t* iterations, int32_t* packed,
    const int32_t* SUBA, const int32_t* INSA, int32_t maxRows, int32_t maxColumns, int32_t bandwidth, float bandwidthRatio,
    int32_t* vertLimit, int32_t* horizLimit, const int8_t* baseToNumber, const int32_t* INSC);

typedef struct orc_msa orc_msa;
orc_msa* orc_msa_new(int32_t maxRows, int32_t maxColumns);
void orc_msa_free(orc_msa* m);
void orc_msa_set_backend(orc_msa* m, orc_fill_limited_fn l, orc_fill_unlimited_fn u);
void orc_msa_set_band(orc_msa* m, int32_t bandwidth, float ratio);
void orc_msa_set_shape(orc_msa* m, int32_t rows, int32_t columns);
int32_t* orc_msa_packed(orc_msa* m);
int64_t orc_msa_iterations(const orc_msa* m, int which);
int32_t orc_msa_last_path(const orc_msa* m);
int32_t orc_msa_greflimit(const orc_msa* m);
const int8_t* orc_msa_gref(const orc_msa* m);
int32_t orc_msa_to_gapped(const orc_msa* m, int32_t p);
int32_t orc_msa_from_gapped(const orc_msa* m, int32_t p);

int orc_msa_fillLimited(orc_msa* m, const int8_t* read, int32_t rlen, const int8_t* ref, int32_t reflen,
    int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* gaps, int32_t ngaps, int32_t* max4);
int orc_msa_fillUnlimited(orc_msa* m, const int8_t* read, int32_t rlen, const int8_t* ref, int32_t reflen,
    int32_t refStartLoc, int32_t refEndLoc, int32_t* gaps, int32_t ngaps, int32_t* max4);
int orc_msa_score2(orc_msa* m, const int8_t* read, const int8_t* ref, int32_t refStartLoc, int32_t refEndLoc,
    int32_t maxRow, int32_t maxCol, int32_t maxState, int32_t* out8);
int orc_msa_score(orc_msa* m, const int8_t* read, const int8_t* ref, int32_t refStartLoc, int32_t refEndLoc,
    int32_t maxRow, int32_t maxCol, int32_t maxState, int gapped, int32_t* out8);
int32_t orc_msa_traceback2(orc_msa* m, const int8_t* read, const int8_t* ref, int32_t refStartLoc, int32_t refEndLoc,
    int32_t row, int32_t col, int32_t state, int8_t* out, int32_t outcap);
int32_t orc_msa_traceback(orc_msa* m, const int8_t* read, const int8_t* ref, int32_t refStartLoc, int32_t refEndLoc,
    int32_t row, int32_t col, int32_t state, int gapped, int8_t* out, int32_t outcap);
int orc_msa_fillAndScoreLimited(orc_msa* m, const int8_t* read, int32_t rlen, const int8_t* ref, int32_t reflen,
    int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* gaps, int32_t ngaps, int32_t* max4, int32_t* out8);

/* batch driver (orc_batch.c): runs fillLimited -> score -> traceback over a task list on `threads` host threads */
typedef struct {
    int64_t read_off; int64_t ref_off; /* byte offsets into the reads / reference buffers */
    int32_t read_len; int32_t ref_len; /* ref_len = length of the reference array the window lives in */
    int32_t ref_start; int32_t ref_end;
    int32_t min_score; int32_t flags;
} orc_task;
typedef struct {
    int32_t result[5]; int32_t path; int64_t iterations;
    int32_t score[8]; int32_t score_len; int32_t match_len; int32_t status; int32_t pad_;
} orc_out;
int64_t orc_batch_run(const int8_t* reads, const int8_t* refs, const orc_task* tasks, orc_out* outs, int64_t ntasks,
    int8_t* match_buf, const int64_t* match_off, int32_t bandwidth, float bandwidthRatio,
    int32_t maxRows, int32_t maxColumns, int use_reference_fill, int threads);
#ifdef __cplusplus
}
#endif
#endif
