/* TEST INFRASTRUCTURE ONLY — see sam_oracle.c */
#ifndef SAM_ORACLE_H
#define SAM_ORACLE_H
#include <stdint.h>
typedef struct {
    int64_t match_off; int32_t match_len, chrom, start, stop, read_len, score, mate, flags; int32_t pad_;
} orc_sam_task;     /* 48 bytes, mirrors bbm_sam_task */
typedef struct { int32_t flag, pos, mapq, scaffold, rnext, pnext, tlen, cigar_len; } orc_sam_out;   /* mirrors bbm_sam_out */
typedef struct { int32_t version14, soft_clip, intron_limit, penalize_ambig, inter_scaffold_padding, pad_[3]; } orc_sam_cfg;
void orc_sam_batch(const orc_sam_task* tasks, int64_t n, const int8_t* match_buf, const int32_t* scaf_off, const int32_t* scaf_loc,
                   const int32_t* scaf_len, int32_t nchroms, const orc_sam_cfg* cfg, orc_sam_out* outs, int8_t* cigar_buf, const int64_t* cigar_off);
#endif
