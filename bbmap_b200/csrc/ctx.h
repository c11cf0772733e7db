// ctx.h — internal header shared by the capi_*.cu translation units: the context behind the C ABI (include/bbmap_cuda.h), its device/pinned
// buffers, the launch wrappers exported by the kernel translation units and the run_* helpers the batched mapper chains.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <chrono>
#include <mutex>
#include <cmath>
#include <algorithm>
#include <cuda_runtime.h>
#include "msa_common.cuh"

using namespace bbm;

#define DECL_W(W) extern "C" int bbm_launch_msa_tiled_w##W(const MsaParams* P, const int* list, int nlist, const unsigned int* endPtr, unsigned int base, unsigned int* counter, int blocks, int dump, cudaStream_t stream);
DECL_W(4) DECL_W(5) DECL_W(6) DECL_W(8) DECL_W(9) DECL_W(12) DECL_W(16)
extern "C" int bbm_launch_msa_classify(const MsaParams* P, unsigned char* cls, unsigned int* cb, int useNarrow, int useStrip, int useBand, cudaStream_t stream);
extern "C" int bbm_msa_class_strip();
extern "C" int bbm_msa_class_band();
extern "C" int bbm_msa_band_threads();
extern "C" size_t bbm_msa_band_thread_bytes(int maxRows, int maxCols, int nd);
extern "C" int bbm_launch_msa_band(const MsaParams* P, const int* list, int nlist, const unsigned int* endPtr, unsigned int base, int nd, int maxRows, int maxCols,
                                   void* scratch, unsigned int* counter, int blocks, cudaStream_t st);
extern "C" int bbm_msa_strip_blocks_per_sm();
extern "C" int bbm_msa_strip_max_cols();
extern "C" size_t bbm_msa_band_smem_bytes(int maxRows, int maxCols, int nd);
extern "C" unsigned long long bbm_msa_strip_task_bytes(int rows, int cols);
extern "C" size_t bbm_msa_strip_fixed_bytes(int chunkCount, int maxRows, int blocks);
extern "C" int bbm_launch_msa_strip(const MsaParams* P, const int* list, const unsigned int* endPtr, unsigned int base, int chunkStart, int chunkCount,
                                    int maxRows, void* scratch, size_t scratchBytes, unsigned int* counter, unsigned long long* poolCursor,
                                    int blocks, int debug, unsigned long long* stats, cudaStream_t st);
extern "C" int bbm_launch_msa_scatter(const MsaParams* P, const unsigned char* cls, unsigned int* cb, int* lists, int* nlist, cudaStream_t stream);
extern "C" int bbm_launch_msa_narrow(const MsaParams* P, const int* nlist, int n, unsigned int* cb, unsigned long long* tb, long long tbWordsPerWarp,
                                     int* lists, int blocks, int useStrip, cudaStream_t stream);
extern "C" int bbm_msa_narrow_threads();
extern "C" int bbm_msa_narrow_buckets();
extern "C" int bbm_launch_banded(const int8_t* q, const int8_t* r, const bbm_band_task* t, bbm_band_out* o, long long n,
                                 unsigned int* counter, int blocks, cudaStream_t st, const int* list, const unsigned int* listCount);
extern "C" int bbm_launch_banded_maxwidth(const bbm_band_task* t, long long n, unsigned int* outMax, int blocks, cudaStream_t st);
extern "C" int bbm_launch_banded_thread(const int8_t* q, const int8_t* r, const bbm_band_task* t, bbm_band_out* o, long long n,
                                        unsigned int* counter, int* wideList, unsigned int* wideCount, int maxBand, int blocks, cudaStream_t st);
extern "C" int bbm_seed_upload_tables(const float* pc, const float* pci);
extern "C" int bbm_launch_seed(const int8_t* bases, const int8_t* quality, const long long* read_off, long long nreads, const bbm_seed_cfg* cfg,
                               int maxKeys, int* nkeys, int* offsets, int* keys, int* keyScores, int8_t* baseScores,
                               float* probScratch, int blocks, int maxProbLen, unsigned int* counter, cudaStream_t st);
extern "C" int bbm_launch_seed_reverse(const int* nkeys, const int* offsets, const int* keys, const long long* read_off, long long nreads,
                                       int maxKeys, int keylen, int* offsetsM, int* keysM, cudaStream_t st);
extern "C" int bbm_seed_threads();
extern "C" int bbm_launch_noindel(const int8_t* reads, const int8_t* refs, const bbm_noindel_task* tasks, int* scores,
                                  int8_t* match_buf, const long long* match_off, long long n, cudaStream_t st);
extern "C" int bbm_index_emit(const int8_t* chrom, int chromLen, int k, int siteHigh, unsigned* keys, int* vals, long long outBase, int* sizes,
                              unsigned invalidKey, cudaStream_t st);
extern "C" int bbm_index_sort_pairs(void* temp, size_t* tempBytes, const unsigned* keysIn, unsigned* keysOut, const int* valsIn, int* valsOut,
                                    long long n, int endBit, cudaStream_t st);
extern "C" int bbm_index_scan(void* temp, size_t* tempBytes, const int* in, int* out, long long n, cudaStream_t st);
extern "C" int bbm_index_count_defined(const int8_t* bytes, long long n, unsigned long long* out, cudaStream_t st);
extern "C" int bbm_index_analyze_block(const int* starts, const int* sites, int k, int* COUNTS, unsigned long long* clump, cudaStream_t st);
extern "C" int bbm_index_finish_counts(int k, int* COUNTS, const unsigned long long* clump, int* maxOut, cudaStream_t st);
extern "C" int bbm_index_lenhist(int k, const int* COUNTS, int* lenCounts, cudaStream_t st);
extern "C" int bbm_launch_gref_build(const int8_t* refs, const bbm_gapped_task* gt, const int* gaps, long long n, int8_t* pool, int stride,
                                     int greflen, bbm_gref_info* info, bbm_msa_task* tasksOut, cudaStream_t st);
extern "C" int bbm_launch_gref_translate(const bbm_gapped_task* gt, long long n, const int8_t* pool, int stride, const bbm_gref_info* info,
                                         bbm_msa_out* outs, cudaStream_t st);
extern "C" int bbm_ingest_threads();
extern "C" int bbm_launch_ingest(int8_t* bases, int8_t* quality, const long long* read_off, long long nreads, int8_t* basesM, int* readFlags,
                                 int flags, int readsPerBlock, int stageBytes, int blocks, cudaStream_t st);
extern "C" int bbm_sam_upload_table(const float* log2tab);
extern "C" int bbm_sam_log2_tab();
extern "C" int bbm_launch_sam(const bbm_sam_task* tasks, long long n, const int8_t* match_buf, const int* scaf_off, const int* scaf_loc, const int* scaf_len,
                              int nchroms, const bbm_sam_cfg* cfg, bbm_sam_out* outs, int8_t* cigar_buf, const long long* cigar_off, cudaStream_t st);
extern "C" int bbm_search_threads();
extern "C" size_t bbm_search_pool_bytes();
extern "C" int bbm_search_mid_stride(int maxKeys, int nblocks);
extern "C" int bbm_launch_search_walk_warp(const bbm_index_cfg* d_cfg, const void* d_blocks, int nblocks, int nchroms, const int* d_counts, const int8_t* d_chroms,
                                           const long long* d_chrom_off, const int8_t* bases, const int8_t* baseScores, const long long* read_off, long long nreads,
                                           const int* nkeys, int maxKeys, int quit2, bbm_search_head* heads, bbm_site* sites, int maxSites, unsigned int* counter, int blocks,
                                           int* mid, int midStride, cudaStream_t st);
extern "C" int bbm_launch_search_prescan_warp(const bbm_index_cfg* d_cfg, const void* d_blocks, int nblocks, int nchroms, const int* d_counts,
                                              const long long* read_off, long long nreads, const int* nkeys, int maxKeys, bbm_search_head* heads,
                                              unsigned int* counter, int blocks, int* mid, int midStride, cudaStream_t st);
extern "C" int bbm_launch_search(const bbm_index_cfg* d_cfg, const void* d_blocks, int nblocks, int nchroms, const int* d_counts, const int* d_hist,
                                 const int8_t* d_chroms, const long long* d_chrom_off, const int8_t* bases, const int8_t* baseScores,
                                 const long long* read_off, long long nreads, const int* nkeys, const int* offsets, const int* keyScores, int maxKeys,
                                 int quitAfterTwoPerfects, bbm_search_head* heads, bbm_site* sites, int maxSites, void* pool,
                                 unsigned int* counter, unsigned long long* prof, int blocks, int forcePool, int phases, int* mid, int midStride,
                                 cudaStream_t st);
extern "C" int bbm_launch_msa_sum_iterations(const bbm_msa_out* outs, long long n, unsigned long long* sum, cudaStream_t st);
extern "C" int bbm_launch_peak(int kind, int blocks, int iters, int* d_out, cudaStream_t st);
extern "C" int bbm_launch_msa_generic(const MsaParams* P, const int* list, int nlist, int* gscratch, long long gstride, cudaStream_t stream, int max_rows, int max_cols, const unsigned int* endPtr, unsigned int base);
extern "C" int bbm_msa_warps_per_block();
extern "C" int bbm_msa_num_wclass();
extern "C" long long bbm_generic_scratch_ints(int rows, int cols);

extern thread_local std::string bbm_g_err;
int fail(int code, const char* what, cudaError_t e = cudaSuccess);
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(BBM_E_CUDA, #call, e_); } while (0)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    DevBuf() = default; DevBuf(const DevBuf&) = delete; DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }            // function-local staging buffers are freed on every return path
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { p = nullptr; return -1; }
        cap = want; return 0;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    int ensure(size_t n) {
        if (n <= cap) return 0;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) { p = nullptr; return -1; }
        cap = want; return 0;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

struct bbm_ctx {
    int device = 0;
    int sms = 0;
    int blocks = 0;             // persistent grid of the tiled kernel
    int bandwidth = 0; float ratio = 0.f;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t gstream = nullptr; cudaEvent_t gev0 = nullptr, gev1 = nullptr;   // side stream for the row-sequential kernel (a few long alignments: pure latency)
    DevBuf scratch, nscratch, counters, overflow, gscratch, lists, nlist, cls;
    int use_narrow = 1000, use_strip = 16, strip_debug = 0, search_shared = 0, search_split = 3;
    long long strip_min_tasks = 8192;
    int use_band = 1;                          // banded limited fills go to the thread-per-alignment band kernel (0: register-tiled kernel + row-sequential re-runs)
    DevBuf bandScratch;
    DevBuf bandedWide;                         // BandedAligner: pairs whose band is too wide for the thread-per-pair kernel
    int banded_thread = 1;                     // option "banded_thread": 0 = warp-per-pair kernel for every pair (A/B, tests)
    int slow_lookahead = 16;                   // scoreSlow: sites of one read taken per round after its first (1 = one site per round, the round-1 schedule)
    DevBuf slowBuf[9];                         // scoreSlow rounds: per-read state, packed requests, their results, counters, gapped requests / gap arrays / results
    size_t strip_budget = (size_t)32 << 30;    // device scratch the strip kernel may use per chunk (raised or lowered with "strip_budget_mb")
    DevBuf stripScratch;
    long long strip_tasks = 0, index_build_us = 0;
    unsigned long long strip_units = 0, strip_lane_iters = 0;
    int search_prof = 0; unsigned long long search_cycles[5] = {0, 0, 0, 0, 0};
    long long band_misses = 0, narrow_tried = 0, narrow_handed_over = 0, tasks_total = 0;
    DevBuf d_reads, d_tasks, d_outs, d_match, d_moff, d_dump, d_refs2, seedScratch, d_seed[8];
    bool seed_tables = false;
    struct IndexBlock { int* starts = nullptr; int* sites = nullptr; long long nsites = 0; int minChrom = 0, maxChrom = 0; };
    std::vector<IndexBlock> iblocks;
    int* d_counts = nullptr; int ihist[1001]; bbm_index_cfg icfg; bool has_index = false; bool index_shared = false;
    const int8_t* d_chroms = nullptr; std::vector<long long> chrom_off;
    void* d_icfg = nullptr; void* d_iblocks = nullptr; int* d_ihist = nullptr; long long* d_chrom_off = nullptr;
    DevBuf searchCtx, searchRev, d_srch[8];
    DevBuf d_sam[8]; bool sam_table = false;   // staging for bbm_sam_batch_host
    DevBuf d_ing[5];   // staging for bbm_ingest_batch_host
    DevBuf grefPool, grefInfo, grefTasks, d_gtasks, d_gaps;   // gapped references (a15)   // staging for the host-buffer entry point
    PinBuf h_stage;
    // the batched mapper (capi_mapper.cu): working buffers of the chain, scaffold table, staging of the host entry point
    DevBuf mapBuf[64], mapScaf[6], mapHost[8];
    int map_nchroms = 0, map_nscaf = 0, map_maxidx_for = -1; bool map_has_names = false; long long map_last_cs = 0, map_last_ms = 0; int map_sites_hint = 0;   // site slots per read the previous batch ended up needing
    std::vector<void*> uploads;
    long long launches = 0;
    double msa_ms = 0.0; long long msa_cells = 0; int msa_count = 0; DevBuf msaCells;     // device time of every run_msa so far; reference cells (with "msa_count")
    std::mutex mu;
};

// helpers defined in the capi_*.cu units and chained by the batched mapper (capi_mapper.cu)
void index_free(bbm_ctx* c);
int run_search(bbm_ctx* c, const int8_t* db, const int8_t* dbs, const int64_t* doff, int64_t nreads, const int* dn, const int* dof, const int* dks, int maxKeys, int quit2, bbm_search_head* dh, bbm_site* ds, int maxSites, int maxReadLen, cudaStream_t st, float* ms_out);
int scoreslow_locked(bbm_ctx* c, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off, const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_refs, const int64_t* d_chrom_off, const int32_t* d_run, const bbm_slow_cfg* cfg, int32_t* d_status, int32_t max_read_len, cudaStream_t st, int64_t* alignments_out, float* ms_out);
int run_msa(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_msa_task* d_tasks, bbm_msa_out* d_outs, int64_t ntasks, int8_t* d_match, const int64_t* d_moff, int max_rows, int max_cols, cudaStream_t st, float* ms_out, int* d_dump);
int run_msa_gapped(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_gapped_task* d_gt, const int32_t* d_gaps, bbm_msa_out* d_outs, int64_t ntasks, int8_t* d_match, const int64_t* d_moff, cudaStream_t st, float* ms_out);
int run_banded(bbm_ctx* c, const int8_t* dq, const int8_t* dr, const bbm_band_task* dt, bbm_band_out* dout, int64_t n, cudaStream_t st, float* ms_out);
int run_seed(bbm_ctx* c, const int8_t* db, const int8_t* dq, const int64_t* doff, int64_t nreads, int max_len, const bbm_seed_cfg* cfg, int maxKeys, int* dn, int* dof, int* dk, int* dks, int8_t* dbs, int* dofM, int* dkM, cudaStream_t st, float* ms_out);
int run_noindel(bbm_ctx* c, const int8_t* dr, const int8_t* dref, const bbm_noindel_task* dt, int* ds, int8_t* dm, const int64_t* dmo, int64_t n, cudaStream_t st, float* ms_out);
int run_ingest(bbm_ctx* c, int8_t* db, int8_t* dq, const int64_t* doff, int64_t nreads, int max_len, int flags, int8_t* dm, int* df, cudaStream_t st, float* ms_out);
int run_sam(bbm_ctx* c, const bbm_sam_task* dt, int64_t n, const int8_t* dm, const int* dso, const int* dsl, const int* dsn, int nchroms, const bbm_sam_cfg* cfg, bbm_sam_out* dout, int8_t* dcb, const int64_t* dco, cudaStream_t st, float* ms_out);

