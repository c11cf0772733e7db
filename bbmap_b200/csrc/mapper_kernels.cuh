// mapper_kernels.cuh — parameter blocks of the mapper kernels (genmatch.cu), shared with the host glue (capi_mapper.cu).
#pragma once
#include <cuda_runtime.h>
#include "../../include/bbmap_cuda.h"

namespace bbm {

constexpr int GM_SLOTS = 3;            // sites of one read that may hold a match string at the same time
constexpr int GM_STATE = 40;           // ints of coroutine state per read

// coroutine state words of genmatch.cu (per read) and its resume labels
enum { S_PC = 0, S_SITE, S_BEST, S_CHANGED, S_RET_SITE, S_RET_RA, S_FIRST, S_TOPSERIAL, S_OLDSLOW, S_OLDSCORE,      // genMatchString / processRead loop
       S_SITE_OLDSCORE, S_MINMSA, S_RECUR, S_PADDING, S_FIXXY, S_FORBID, S_NOINDEL, S_MINLOC, S_MAXLOC, S_EPL, S_EPR, S_OLD0, S_LIM,   // genMatchStringForSite / realign_new
       S_REQ, S_REQ_GAPPED, S_SLOTMASK, S_STATUS, S_TOPCHANGED, S_FILLS, S_SERIALS, S_SETSS };
enum { PC_DONE = 0, PC_BEGIN, PC_GMS_LOOP, PC_GMS_AFTER_SITE, PC_GMS_SORT, PC_GMS_AFTER_TOP, PC_GMS_FINISH,
       PC_SITE_BEGIN, PC_SITE_AFTER_R1, PC_SITE_AFTER_R2, PC_SITE_END,
       PC_RA_BEGIN, PC_RA_FILL1, PC_RA_FILL2, PC_RA_FILL3, PC_RA_FILL4, PC_RA_AFTER };

// ---- pairing.cu ----
constexpr int PAIR_STATE = 16;
enum { PS_UNPAIRED0 = 0, PS_UNPAIRED1, PS_ACTIVE0, PS_ACTIVE1, PS_RETAIN1, PS_RETAIN2, PS_MAXMM, PS_FIND, PS_STATUS0, PS_STATUS1, PS_DISCARDED };
enum { PAIR_OP_INIT = 0, PAIR_OP_RESCUE_PREP0, PAIR_OP_RESCUE_PREP1, PAIR_OP_RESCUE_APPLY0, PAIR_OP_RESCUE_APPLY1, PAIR_OP_FINAL };
struct RescueAux { int pair, dir, anchor, chrom, strand, minus, msa_req, valid, sw, old_start, pad_[2]; };      // one quickRescue task: who asked, and slowRescue's state
struct PairParams {
    bbm_ss* lists; int* nss; long long npairs; int cap; const long long* read_off;
    const int8_t* basesP; const int8_t* basesM; const int8_t* quality; const int8_t* refs; const long long* chrom_off; const int* nkeys;
    bbm_map_cfg cfg; bbm_policy_cfg pc; bbm_tipdel_cfg tc; int clearzone1e;
    int* pstate; int* rflags;                            // [npairs][PAIR_STATE]; per read: bit0 mapped, bit1 perfect, bit2 ambiguous, bit3 paired
    bbm_rescue_task* rtasks; const bbm_rescue_out* routs; RescueAux* raux; bbm_ss* rsites; int* rtask_of; int maxTasks;
    bbm_msa_task* mtasks; const bbm_msa_out* mouts;
    int* counters;                                       // [1] quickRescue tasks, [2] alignment requests
    unsigned long long* stats;                           // [0] mated pairs, [1] sum of inner lengths
};

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
extern "C" int bbm_pair_state_ints();
extern "C" int bbm_launch_pair(const bbm::PairParams* P, int op, cudaStream_t st);
extern "C" int bbm_launch_rescue_mid(const bbm::PairParams* P, int ntasks, cudaStream_t st);
extern "C" int bbm_launch_pair_finish(const bbm::PairParams* P, const bbm::FinParams* F, cudaStream_t st);
