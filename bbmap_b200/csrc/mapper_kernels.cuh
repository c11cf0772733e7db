// mapper_kernels.cuh — parameter blocks of the mapper kernels (genmatch.cu), shared with the host glue (capi_mapper.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/bbmap_cuda.h"

namespace bbm {

constexpr int GM_SLOTS = 3;            // sites of one read that may hold a match string at the same time
constexpr int GM_STATE = 40;           // ints of coroutine state per read

struct GmParams {
    bbm_ss* lists; int* nss; long long nreads; int cap; const long long* read_off;
    const int8_t* basesP; const int8_t* basesM; const int8_t* refs; const long long* chrom_off;
    bbm_map_cfg cfg; int setSSScore;
    const int* rflags;                                   // per read: bit3 = r.paired() (NULL = unpaired run)
    int* state; int8_t* mslots; long long ms; int* mlen; // [nreads][GM_SLOTS] slots of ms bytes + their lengths
    bbm_msa_task* tasks; const bbm_msa_out* outs; const int8_t* rmatch; long long rstride;
    bbm_gapped_task* gtasks; int* gaps; const bbm_msa_out* gouts; const int8_t* gmatch; long long gstride;
    int* counters;                                       // [0] reads parked, [1] plain requests, [2] gapped requests, [3] largest match capacity asked for (plain), [4] (gapped)
    int first;                                           // first launch: initialise the state
};

struct FinParams {
    bbm_ss* lists; int* nss; long long nreads; int cap; const long long* read_off;
    const int8_t* basesP; const int8_t* basesM; const int8_t* refs; const long long* chrom_off;
    bbm_policy_cfg pc; bbm_map_cfg cfg;
    const bbm_read_out* flags; const int* state; int8_t* mslots; long long ms; int* mlen; bbm_map_rec* recs;
};

struct SamTextParams {
    const bbm_map_rec* recs; const bbm_sam_out* sam; long long nreads; const long long* read_off;
    const int8_t* bases; const int8_t* basesM; const int8_t* quality;           // validated reads as sequenced, their reverse complements, phred values or NULL
    const int8_t* names; const long long* name_off;                            // read names or NULL ('*')
    const int8_t* scaf_names; const long long* scaf_name_off;                  // scaffold names or NULL ('*')
    const int8_t* cigar; const long long* cigar_off;
    const int8_t* mslots; long long ms;
    int paired; int intron_limit;
    int* lens; int* nm;                                                        // [nreads] pass-1 results
    const int* line_off;                                                       // [nreads+1] exclusive scan of lens (ints: a batch's text stays below 2 GiB)
    int8_t* text; long long* text_off;                                         // output
};

}  // namespace bbm

extern "C" int bbm_launch_genmatch(const bbm::GmParams* P, cudaStream_t st);
extern "C" int bbm_launch_map_finish(const bbm::FinParams* P, cudaStream_t st);
extern "C" int bbm_launch_map_sam_tasks(const bbm_map_rec* recs, long long nreads, const long long* read_off, long long ms, int paired, bbm_sam_task* tasks, cudaStream_t st);
extern "C" int bbm_launch_map_runmask(const bbm_read_out* out, const int* nss, long long n, int paired, int* run, int* masked, cudaStream_t st);
extern "C" int bbm_launch_map_arange(long long* off, long long n, long long stride, cudaStream_t st);
extern "C" int bbm_launch_map_overflow(const bbm_search_head* heads, long long n, int maxSites, int* counter, cudaStream_t st);
extern "C" int bbm_launch_map_status(const bbm_search_head* heads, int maxSites, const int* slowStatus, const int* nkeys, bbm_map_rec* recs, long long n,
                                     unsigned long long* counters, cudaStream_t st);
extern "C" int bbm_launch_map_copy_match(const bbm_map_rec* recs, const int8_t* mslots, long long ms, int8_t* out, long long stride, long long n, cudaStream_t st);
extern "C" int bbm_genmatch_state_ints();
extern "C" int bbm_genmatch_slots();
extern "C" int bbm_launch_samtext_len(const bbm::SamTextParams* P, cudaStream_t st);
extern "C" int bbm_launch_samtext_write(const bbm::SamTextParams* P, cudaStream_t st);
