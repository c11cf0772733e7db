/* TEST INFRASTRUCTURE ONLY — see host_oracle.c */
#ifndef HOST_ORACLE_H
#define HOST_ORACLE_H
#include <stdint.h>
typedef struct { int32_t keylen, maxDesiredKeys, baseKeyHitScore, minApproxHitsToKeep; float keyDensity, maxKeyDensity, minKeyDensity, pad_; } orc_seed_cfg; /* 32 B */
void orc_quality_tables(float* prob_correct, float* prob_correct_inverse);
void orc_make_key_probs(const int8_t* quality, const int8_t* bases, int len, int keylen, float* out);
int orc_make_offsets3(const float* kep, int readlenOriginal, int blocksize, float density, float maxDensity, int minKeysDesired, int semiperfect, int* offsets);
int orc_rcomp_key_fast(int kmer, int k);
int orc_quickmap_seed(const int8_t* bases, const int8_t* quality, int len, const orc_seed_cfg* cfg, int32_t* offsets, int32_t* keys, int32_t* keyScores, int8_t* baseScores, float* keyProbsScratch);
void orc_seed_batch(const int8_t* bases, const int8_t* quality, const int64_t* read_off, int64_t nreads, const orc_seed_cfg* cfg,
                    int32_t maxKeys, int32_t* nkeys, int32_t* offsets, int32_t* keys, int32_t* keyScores, int8_t* baseScores);
typedef struct { int64_t read_off, ref_off; int32_t read_len, ref_len, ref_start, flags; } orc_noindel_task; /* 32 B */
int orc_score_no_indels(const int8_t* read, int len, const int8_t* ref, int refLen, int refStart, int8_t* match);
void orc_noindel_batch(const int8_t* reads, const int8_t* refs, const orc_noindel_task* tasks, int32_t* scores, int8_t* match_buf,
                       const int64_t* match_off, int64_t n);
int orc_ingest_read(int8_t* bases, int8_t* quality, int len, int flags, int8_t* basesM);
void orc_ingest_batch(int8_t* bases, int8_t* quality, const int64_t* read_off, int64_t nreads, int flags, int8_t* basesM, int32_t* readFlags);
#endif
