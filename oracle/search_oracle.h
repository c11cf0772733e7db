/* TEST INFRASTRUCTURE ONLY — see search_oracle.c */
#ifndef SEARCH_ORACLE_H
#define SEARCH_ORACLE_H
#include <stdint.h>
#include "index_oracle.h"
#define ORC_MAX_KEYS   96
#define ORC_MAX_READ   608
#define ORC_MAX_SITES  48
#define ORC_MAX_GAPS   10
#define ORC_MAX_BLOCKS 64
#define ORC_ST_ANOMALY        1   /* extendScore left no located base (reference prints an anomaly, score=-99999) */
#define ORC_ST_SITE_OVERFLOW  2   /* more than ORC_MAX_SITES sites for one read */
#define ORC_ST_GAP_OVERFLOW   4   /* gap array longer than ORC_MAX_GAPS */
#define ORC_ST_GAPFIX         8   /* subsumption into a site that carries gaps (GapTools.fixGaps not restated) */
#define ORC_ST_BADARG        16
typedef struct { const int32_t* starts; const int32_t* sites; } orc_search_block;
typedef struct {
    const orc_index_cfg* cfg; const orc_search_block* blocks; int32_t nblocks; int32_t nchroms;
    const int32_t* counts; const int32_t* hist; const int8_t* chroms; const int64_t* chrom_off;
} orc_search_index;
typedef struct {            /* 64 bytes: one SiteScore as BBIndex emits it */
    int32_t chrom, start, stop, hits, score, ngaps;
    int8_t strand, perfect, semiperfect, pad_;
    int32_t gaps[ORC_MAX_GAPS - 1];      /* 9 ints: start/stop pairs; longer arrays set ORC_ST_GAP_OVERFLOW */
} orc_site;
typedef struct {
    int32_t nsites, status, num_hits, max_score, max_quick_score, pad_;
    int32_t best_scores[6];
    orc_site sites[ORC_MAX_SITES];
} orc_read_result;
void orc_search_read(const orc_search_index* X, const int8_t* basesP, int len, const int8_t* baseScoresP, const int32_t* offsetsIn,
                     const int32_t* keyScoresIn, int nkeys, int quitAfterTwoPerfects, orc_read_result* R);
void orc_search_batch(const orc_search_index* X, const int8_t* bases, const int8_t* baseScores, const int64_t* read_off, int64_t nreads,
                      const int32_t* nkeys, const int32_t* offsets, const int32_t* keyScores, int32_t maxKeys, int quitAfterTwoPerfects,
                      orc_read_result* results);
#endif
