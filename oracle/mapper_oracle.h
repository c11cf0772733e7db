/* TEST INFRASTRUCTURE ONLY — see mapper_oracle.c (the tail of BBMapThread.processRead / processReadPair, sequential restatement). */
#ifndef MAPPER_ORACLE_H
#define MAPPER_ORACLE_H
#include <stdint.h>
#include "sitelist_oracle.h"
#define ORC_MAP_ST_MATCH_OVERFLOW 1   /* a match string did not fit the caller's slot */
#define ORC_MAP_ST_TIP            2
#define ORC_MAP_ST_SLOW           32
#define ORC_MAP_ST_LIST_OVERFLOW  64  /* a rescued site did not fit the list */   /* calcTipScorePenalty ran off the match string (the reference would throw) */
typedef struct {            /* == bbm_map_cfg (include/bbmap_cuda.h) */
    int32_t paired;                     /* reads 2i / 2i+1 are mates */
    float min_ratio, min_ratio_paired, min_ratio_pre_rescue, secondary_site_score_ratio;
    int32_t slow_align_padding, max_indel, ambiguous_toss, penalize_ambig;
    int32_t average_pair_dist, max_pair_dist, max_rescue_dist, max_rescue_mismatches;
    int32_t do_rescue, kill_bad_pairs, require_correct_strands, same_strand_pairs;
    int32_t pad_[3];
} orc_map_cfg;
typedef struct {            /* == bbm_map_rec: the Read fields SamLine(Read,int) reads, 48 bytes */
    int32_t chrom, start, stop, strand, map_score, flags;     /* flags: bit0 mapped, bit1 perfect, bit2 ambiguous, bit3 paired, bit4 rescued, bit5 discarded */
    int32_t match_len, cz3_sub, tip_penalty, status, pad_[2];
} orc_map_rec;
int orc_score_match(const int8_t* match, int n);
int64_t orc_map_finish_single(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int64_t* read_off,
                              const int8_t* refs, const int64_t* chrom_off, const orc_policy_cfg* pc, const orc_map_cfg* cfg, const orc_read_out* flags_in,
                              orc_map_rec* recs, int8_t* match_buf, int64_t match_stride);
int64_t orc_map_pairs(orc_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int8_t* basesP, const int8_t* basesM, const int8_t* quality,
                      const int64_t* read_off, const int8_t* refs, const int64_t* chrom_off, const int32_t* nkeys, const orc_policy_cfg* pc,
                      const orc_map_cfg* cfg, const void* slow_cfg, const orc_tipdel_cfg* tc, orc_map_rec* recs, int8_t* match_buf, int64_t match_stride,
                      int64_t* stats);
int orc_test_pair_initial(orc_ss* a, int32_t* na, int len1, orc_ss* b, int32_t* nb, int len2, const orc_map_cfg* cfg, int maxTrim);
void orc_test_pair_final(orc_ss* a, int32_t* na, int len1, orc_ss* b, int32_t* nb, int len2, const orc_map_cfg* cfg, int maxTrim);
int orc_test_can_pair(const orc_ss* ss1, const orc_ss* ss2, int len1, int len2, const orc_map_cfg* cfg);
int orc_test_remove_low_quality_paired(orc_ss* v, int n, int maxSw, float multSingle, float multPaired);
int orc_test_is_bad_pair(const orc_map_rec* r, const orc_map_rec* m, const orc_map_cfg* cfg);
int orc_test_site_op(int op, orc_ss* s, int8_t* match, int32_t* mlen, int32_t mcap, const int8_t* bases, int len, const int8_t* refs, const int64_t* chrom_off,
                     int tiplen, int maxIndel);
int orc_test_realign_new(orc_ss* s, int8_t* match, int32_t* mlen, int32_t mcap, const int8_t* bases, int len, const int8_t* refs, const int64_t* chrom_off,
                         int padding, int recur, int minValidScore, int forbidIndels, int fixXY);
int orc_test_rescue(orc_ss* A, int nA, int lenA, orc_ss* L, int32_t* nL, int cap, const int8_t* basesP, const int8_t* basesM, const int8_t* qualL, int lenL,
                    int searchDist, const int8_t* refs, const int64_t* chrom_off, const orc_map_cfg* cfg, const orc_tipdel_cfg* tc, int clearzone1e, int64_t* counts);
int orc_test_gen_match_string(orc_ss* lists, int32_t* n, int cap, const int8_t* basesP, const int8_t* basesM, int len, const int8_t* refs, const int64_t* chrom_off,
                              const orc_map_cfg* cfg, int maxSwScore, int setSSScore, int32_t* paired, int8_t* top_match, int32_t* top_mlen, int32_t mcap);
#endif
