/* TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the tip-deletion search and the mate-rescue scan.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.  Parity UNPINNED against Java
 * (no JVM in the image): pinned by hand-built cases in tests/test_rescue_oracle.py. */
#pragma once
#include <stdint.h>

typedef struct {            /* == bbm_tipdel_task (include/bbmap_cuda.h), 48 bytes */
    int64_t read_off, ref_off;
    int32_t read_len, ref_len, min_index, start, stop, slow_score, max_imperfect, flags;
} orc_tipdel_task;
typedef struct { int32_t start, stop, right, left; } orc_tipdel_out;
typedef struct { int32_t search_range, max_tiplen, align_columns, slow_rescue_padding; } orc_tipdel_cfg;

typedef struct {            /* == bbm_rescue_task, 56 bytes */
    int64_t read_off, ref_off;
    int32_t read_len, ref_len, min_index, max_index, loc, search_dist, ideal_start, max_mismatches, flags, pad_;
} orc_rescue_task;
typedef struct { int32_t start, stop, mismatches, max_contig, score, perfect, in_bounds, pad_; } orc_rescue_out;
typedef struct { int32_t points_match, points_match2, use_affine, base_hit_score; } orc_rescue_cfg;

int orc_find_tip_deletions_right(const int8_t* bases, int len, const int8_t* ref, int refLen, int minIndex, int originalStop, int searchDist, int tiplen);
int orc_find_tip_deletions_left(const int8_t* bases, int len, const int8_t* ref, int refLen, int minIndex, int originalStart, int searchDist, int tiplen);
void orc_tipdel_batch(const int8_t* reads, const int8_t* refs, const orc_tipdel_task* tasks, int64_t n, const orc_tipdel_cfg* cfg, orc_tipdel_out* outs);
void orc_rescue_batch(const int8_t* reads, const int8_t* refs, const orc_rescue_task* tasks, int64_t n, const orc_rescue_cfg* cfg, orc_rescue_out* outs);
