/* TEST INFRASTRUCTURE ONLY — see index_oracle.c */
#ifndef INDEX_ORACLE_H
#define INDEX_ORACLE_H
#include <stdint.h>
typedef struct {        /* mirrors bbm_index_cfg (include/bbmap_cuda.h): BBIndex statics after BBMap.loadIndex */
    int32_t keylen, chrombits, shift_length, chroms_per_block;
    int32_t max_hits_reduction2, maximum_max_hits_reduction, hit_reduction_div, points_per_site;
    int32_t min_index_to_drop_long_hit_list, max_average_list_to_search, max_average_list_to_search2, max_single_list_to_search;
    int32_t max_shortest_list_to_search, max_usable_length, max_usable_length2, pad_;
    float fraction_to_exclude, padf_[3];
} orc_index_cfg;        /* 80 bytes */
void orc_index_cfg_init(orc_index_cfg* c, int k, int chrombits, int64_t numDefinedBases);
int orc_auto_chrombits(const int64_t* chrom_off, int nchroms);
int64_t orc_index_build_block(const int8_t* chroms, const int64_t* chrom_off, int minChrom, int maxChrom, const orc_index_cfg* c,
                              int32_t* starts, int32_t** sites_out);
void orc_index_analyze(int nblocks, int32_t* const* starts, int32_t* const* sites, orc_index_cfg* c, int32_t* COUNTS, int32_t* hist1001);
#endif
