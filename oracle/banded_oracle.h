/* TEST INFRASTRUCTURE ONLY — see banded_oracle.c */
#ifndef BANDED_ORACLE_H
#define BANDED_ORACLE_H
#include <stdint.h>
typedef struct { int64_t query_off, ref_off; int32_t query_len, ref_len, qstart, rstart, max_edits, max_width, exact, dir; } orc_band_task; /* 48 B */
typedef struct { int32_t edits; int32_t rv[5]; int32_t status; int32_t pad_; } orc_band_out;                                           /* 32 B */
void orc_banded_tables(int8_t* b2n, int8_t* comp);
int orc_banded_align(int dir, const int8_t* query, const int8_t* ref, int qlen, int rlen, int qstart, int rstart, int maxEdits,
                     int exact, int maxWidth, int32_t* rv);
void orc_banded_set_reference_fns(void* f, void* frc, void* r, void* rrc);
int orc_banded_batch(const int8_t* queries, const int8_t* refs, const orc_band_task* tasks, orc_band_out* outs, int64_t n, int use_reference, int threads);
#endif
