/* TEST INFRASTRUCTURE ONLY — CPU restatement (oracle) of the per-read site-list policies of the unpaired mapping loop.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.  Parity UNPINNED against Java (no JVM):
 * pinned by hand-built lists in tests/test_sitelist_oracle.py. */
#pragma once
#include <stdint.h>
#include "rescue_oracle.h"

#define ORC_MAX_GAPS 10
typedef struct {            /* == bbm_ss (include/bbmap_cuda.h), 80 bytes */
    int32_t chrom, start, stop, hits, score, quick_score, slow_score, paired_score;
    int8_t strand, perfect, semiperfect, rescued;
    int32_t ngaps;
    int32_t gaps[ORC_MAX_GAPS - 1];
    int32_t has_match;
} orc_ss;
typedef struct {            /* == bbm_policy_cfg, 80 bytes */
    int32_t trim_list, min_trim_sites_to_retain, max_trim_sites_to_retain, quick_match_strings;
    int32_t clearzone1, clearzone1b, clearzone1c, clearzonep, clearzone3, clearzone1e, clearzone_limit1e, print_secondary;
    float min_align_ratio, cz1b_scale, cz1b_flat, cz1c_scale;
    float cz1c_flat; int32_t pad_[3];
} orc_policy_cfg;
typedef struct { int32_t near_perfect, flags, clearzone, best_sites; } orc_read_out;   /* flags: bit0 mapped, bit1 perfect, bit2 ambiguous */

void orc_ss_set_slow_score(orc_ss* s, int x);
int orc_fix_gaps(int a, int b, int32_t* gaps, int n, int minGap);
void orc_ss_set_limits(orc_ss* s, int a, int b);
void orc_ss_set_stop(orc_ss* s, int b);
void orc_ss_set_start(orc_ss* s, int a);
int orc_calc_gref_len(const orc_ss* s);
void orc_sitelist_trim(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int32_t* read_len, const orc_policy_cfg* cfg, orc_read_out* out);
void orc_sitelist_noindel(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int64_t* read_off,
                          const int8_t* refs, const int64_t* chrom_off, const orc_policy_cfg* cfg, orc_read_out* out);
void orc_sitelist_final(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int32_t* read_len, const orc_policy_cfg* cfg, orc_read_out* out);
void orc_sitelist_tipdel(orc_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int8_t* quality,
                         const int64_t* read_off, const int8_t* refs, const int64_t* chrom_off, const int32_t* chrom_min_index, const orc_tipdel_cfg* tc, orc_read_out* out);
void orc_sitelist_bounds(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int32_t* read_len, const int32_t* chrom_max_index,
                         const int32_t* scaf_off, const int32_t* scaf_loc, int32_t inter_scaffold_padding, int32_t sam_out, int32_t expected_len_limit,
                         orc_read_out* out);
void orc_sitelist_clearzone3(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int32_t* read_len, const orc_policy_cfg* cfg,
                             int32_t ambiguous_toss, orc_read_out* out);
void orc_sitelist_tip_penalty(orc_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int64_t* read_off, const int8_t* bases,
                              const int8_t* match, const int64_t* match_off, const orc_read_out* flags, int32_t tiplen, int32_t* penalty, int32_t* status);
void orc_sl_sort(orc_ss* v, int n, int positional);
int orc_sl_trim_below_cutoff(orc_ss* v, int n, int cutoff, int retainPaired, int minS, int maxS);
int orc_sl_trim_list(orc_ss* v, int* n, int retainPaired, int maxScore, int specialCasePerfect, int minS, int maxS);
int orc_sl_merge_duplicates(orc_ss* v, int n);
int orc_sl_count_top_scores(const orc_ss* v, int n, int thresh);
void orc_sl_set_perfect(orc_ss* s, const int8_t* bases, int len, const int8_t* ref, int refLen);
