// capi.cu — the C ABI of libbbmapcuda.so (see include/bbmap_cuda.h).  Host-side glue only: device buffers, streams,
// launches.  No CPU implementation of any compute path lives here: without a device every call fails loudly.
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
extern "C" int bbm_launch_msa_classify(const MsaParams* P, unsigned char* cls, unsigned int* cb, int useNarrow, int useStrip, cudaStream_t stream);
extern "C" int bbm_msa_class_strip();
extern "C" int bbm_msa_strip_blocks_per_sm();
extern "C" int bbm_msa_strip_max_cols();
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
                                 unsigned int* counter, int blocks, cudaStream_t st);
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
extern "C" int bbm_launch_search_prescan_warp(const bbm_index_cfg* d_cfg, const void* d_blocks, int nblocks, int nchroms, const int* d_counts,
                                              const long long* read_off, long long nreads, const int* nkeys, int maxKeys, bbm_search_head* heads,
                                              unsigned int* counter, int blocks, int* mid, int midStride, cudaStream_t st);
extern "C" int bbm_launch_search(const bbm_index_cfg* d_cfg, const void* d_blocks, int nblocks, int nchroms, const int* d_counts, const int* d_hist,
                                 const int8_t* d_chroms, const long long* d_chrom_off, const int8_t* bases, const int8_t* baseScores,
                                 const long long* read_off, long long nreads, const int* nkeys, const int* offsets, const int* keyScores, int maxKeys,
                                 int quitAfterTwoPerfects, bbm_search_head* heads, bbm_site* sites, int maxSites, void* pool,
                                 unsigned int* counter, unsigned long long* prof, int blocks, int forcePool, int phases, int* mid, int midStride,
                                 cudaStream_t st);
extern "C" int bbm_launch_peak(int kind, int blocks, int iters, int* d_out, cudaStream_t st);
extern "C" int bbm_launch_msa_generic(const MsaParams* P, const int* list, int nlist, int* gscratch, long long gstride, cudaStream_t stream, int max_rows, int max_cols, const unsigned int* endPtr, unsigned int base);
extern "C" int bbm_msa_warps_per_block();
extern "C" int bbm_msa_num_wclass();
extern "C" long long bbm_generic_scratch_ints(int rows, int cols);

static thread_local std::string g_err;
static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    g_err = what;
    if (e != cudaSuccess) { g_err += ": "; g_err += cudaGetErrorString(e); }
    return code;
}
#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(BBM_E_CUDA, #call, e_); } while (0)

struct DevBuf {
    void* p = nullptr; size_t cap = 0;
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
    int use_narrow = 1000, use_strip = 16, strip_debug = 0, search_shared = 0, search_split = 2;
    long long strip_min_tasks = 8192;
    DevBuf slowBuf[7];                         // scoreSlow rounds: per-read state, packed requests, their results, counters, gapped requests / gap arrays / results
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
    int* d_counts = nullptr; int ihist[1001]; bbm_index_cfg icfg; bool has_index = false;
    const int8_t* d_chroms = nullptr; std::vector<long long> chrom_off;
    void* d_icfg = nullptr; void* d_iblocks = nullptr; int* d_ihist = nullptr; long long* d_chrom_off = nullptr;
    DevBuf searchCtx, searchRev, d_srch[8];
    DevBuf d_sam[8]; bool sam_table = false;   // staging for bbm_sam_batch_host
    DevBuf d_ing[5];   // staging for bbm_ingest_batch_host
    DevBuf grefPool, grefInfo, grefTasks, d_gtasks, d_gaps;   // gapped references (a15)   // staging for the host-buffer entry point
    PinBuf h_stage;
    std::vector<void*> uploads;
    long long launches = 0;
    std::mutex mu;
};

static void index_free(bbm_ctx* c);

extern "C" const char* bbm_last_error(void) { return g_err.c_str(); }

extern "C" int bbm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int bbm_init(int device, bbm_ctx** out) {
    if (!out) return fail(BBM_E_ARG, "bbm_init: out is null");
    int n = bbm_device_count();
    if (n <= 0) return fail(BBM_E_NODEVICE, "no CUDA device: libbbmapcuda has no CPU fallback");
    if (device < 0 || device >= n) return fail(BBM_E_ARG, "bbm_init: bad device ordinal");
    CK(cudaSetDevice(device));
    bbm_ctx* c = new bbm_ctx();
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->sms = prop.multiProcessorCount;
    c->blocks = c->sms * 4;     // 4 blocks x 4 warps per SM (register-bound); the kernel is persistent over a task counter
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaStreamCreateWithFlags(&c->gstream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->gev0, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->gev1, cudaEventDisableTiming));
    if (c->counters.ensure(256 * 4)) return fail(BBM_E_CUDA, "cudaMalloc counters");
    {   // strip-kernel scratch budget: a third of what is free now, at most 32 GB (B200: 180 GB of HBM3e)
        size_t freeB = 0, totalB = 0;
        if (cudaMemGetInfo(&freeB, &totalB) == cudaSuccess && freeB / 3 < c->strip_budget) c->strip_budget = freeB / 3;
    }
    *out = c;
    return BBM_OK;
}

extern "C" void bbm_destroy(bbm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (void* p : c->uploads) cudaFree(p);
    c->scratch.release(); c->counters.release(); c->overflow.release(); c->gscratch.release(); c->lists.release(); c->cls.release(); c->nscratch.release(); c->nlist.release();
    c->d_reads.release(); c->d_tasks.release(); c->d_outs.release(); c->d_match.release(); c->d_moff.release(); c->d_dump.release(); c->d_refs2.release(); c->seedScratch.release(); for (auto& b : c->d_seed) b.release(); c->stripScratch.release(); for (auto& b : c->slowBuf) b.release(); c->searchCtx.release(); c->searchRev.release(); for (auto& b : c->d_ing) b.release(); for (auto& b : c->d_sam) b.release(); c->grefPool.release(); c->grefInfo.release(); c->grefTasks.release(); c->d_gtasks.release(); c->d_gaps.release(); for (auto& b : c->d_srch) b.release();
    c->h_stage.release();
    index_free(c);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->gev0) cudaEventDestroy(c->gev0);
    if (c->gev1) cudaEventDestroy(c->gev1);
    if (c->gstream) cudaStreamDestroy(c->gstream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int bbm_set_band(bbm_ctx* c, int32_t bandwidth, float ratio) {
    if (!c) return fail(BBM_E_ARG, "null ctx");
    c->bandwidth = bandwidth; c->ratio = ratio;
    return BBM_OK;
}

extern "C" int bbm_upload(bbm_ctx* c, const void* host, int64_t nbytes, void** dev_out) {
    if (!c || !host || nbytes < 0 || !dev_out) return fail(BBM_E_ARG, "bbm_upload: bad argument");
    CK(cudaSetDevice(c->device));
    void* p = nullptr;
    CK(cudaMalloc(&p, (size_t)nbytes + 256));
    CK(cudaMemcpy(p, host, (size_t)nbytes, cudaMemcpyHostToDevice));
    CK(cudaMemset((char*)p + nbytes, 'N', 256));
    c->uploads.push_back(p);
    *dev_out = p;
    return BBM_OK;
}

extern "C" int bbm_free_dev(bbm_ctx* c, void* dev) {
    if (!c) return fail(BBM_E_ARG, "null ctx");
    for (size_t i = 0; i < c->uploads.size(); ++i)
        if (c->uploads[i] == dev) { cudaFree(dev); c->uploads.erase(c->uploads.begin() + i); return BBM_OK; }
    return fail(BBM_E_ARG, "bbm_free_dev: unknown pointer");
}

extern "C" int64_t bbm_launch_count(const bbm_ctx* c) { return c ? c->launches : 0; }

// Counter block layout: see CB_* in msa_kernels.cuh.
static int run_msa(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_msa_task* d_tasks, bbm_msa_out* d_outs,
                   int64_t ntasks, int8_t* d_match, const int64_t* d_moff, int max_rows, int max_cols, cudaStream_t st,
                   float* ms_out, int* d_dump) {
    if (ntasks <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    if (ntasks > 0x7fffffffLL) return fail(BBM_E_ARG, "too many tasks in one batch (max 2^31-1)");
    if (max_rows < 1) max_rows = MAXR;
    if (max_cols < 1) max_cols = 3000;
    const int wpb = bbm_msa_warps_per_block();
    const int nw = bbm_msa_num_wclass();
    const int nb = bbm_msa_narrow_buckets();
    const int tiledRows = max_rows < MAXR ? max_rows : MAXR;
    const long long words = (long long)(tiledRows + 40) * 32;           // one 64-bit code word per (step,lane)
    const int narrowBlocks = c->sms * 4;
    const int narrowWarps = narrowBlocks * (bbm_msa_narrow_threads() / 32);
    // 0 = off, 1 = try every shape-eligible alignment, n>1 = only those with (best possible score - minScore) <= n points
    const int useNarrow = (c->use_narrow && d_dump == nullptr && ntasks >= c->strip_min_tasks) ? c->use_narrow : 0;   // thread-per-alignment as well: not for small batches (see useStrip)
    // 0 = off; n>0: limited un-banded fills whose work estimate falls in buckets < n go to the strip kernel, larger ones to the tiled kernel
    // The strip kernel is thread-per-alignment: it needs tens of thousands of alignments to fill 148 SMs, and a batch of a few dozen wide
    // alignments would run as a few dozen single threads (measured: 14 alignments = 48 ms).  Small batches (scoreSlow's later rounds and
    // padding retries) go to the warp-per-alignment tiled kernel instead; results are identical by construction and by test.
    const int useStrip = (c->use_strip && d_dump == nullptr && ntasks >= c->strip_min_tasks) ? c->use_strip : 0;
    const int CS = bbm_msa_class_strip();
    if (c->scratch.ensure((size_t)c->blocks * wpb * words * 8)) return fail(BBM_E_CUDA, "cudaMalloc traceback scratch");
    if (useNarrow && c->nscratch.ensure((size_t)narrowWarps * words * 8)) return fail(BBM_E_CUDA, "cudaMalloc narrow traceback scratch");
    if (c->overflow.ensure((size_t)ntasks * 4 + 16)) return fail(BBM_E_CUDA, "cudaMalloc overflow list");
    if (c->lists.ensure((size_t)ntasks * 4 + 16)) return fail(BBM_E_CUDA, "cudaMalloc class lists");
    if (useNarrow && c->nlist.ensure((size_t)ntasks * 4 + 16)) return fail(BBM_E_CUDA, "cudaMalloc narrow list");
    if (c->cls.ensure((size_t)ntasks + 16)) return fail(BBM_E_CUDA, "cudaMalloc class ids");
    unsigned int* cb = (unsigned int*)c->counters.p;
    MsaParams P;
    P.reads = d_reads; P.refs = d_refs; P.tasks = d_tasks; P.outs = d_outs; P.ntasks = ntasks;
    P.match_buf = d_match; P.match_off = (const long long*)d_moff;
    P.bandwidth = c->bandwidth; P.ratio = c->ratio;
    P.scratch = (unsigned long long*)c->scratch.p; P.scratch_words = words;
    P.counter = nullptr; P.overflow_count = cb + 48;
    P.overflow_list = (int*)c->overflow.p;
    P.dump = d_dump;
    CK(cudaMemsetAsync(c->counters.p, 0, 192 * 4, st));
    CK(cudaEventRecord(c->ev0, st));
    int e = bbm_launch_msa_classify(&P, (unsigned char*)c->cls.p, cb, useNarrow, useStrip, st);
    if (e) return fail(BBM_E_CUDA, "msa_classify_kernel launch", (cudaError_t)e);
    c->launches++;
    unsigned int h[192];
    CK(cudaMemcpyAsync(h, cb, 192 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    unsigned int base[16]; unsigned int acc = 0;
    for (int k = 0; k < 16; ++k) { base[k] = acc; if (k <= nw || k == CS) acc += h[k]; }
    unsigned int nbase[64]; unsigned int nacc = 0;
    for (int k = 0; k < 64; ++k) { nbase[k] = nacc; if (k < nb) nacc += h[64 + k]; }
    unsigned int curs[16]; memcpy(curs, base, sizeof(curs));
    unsigned int sbBase[16];
    {   // strip list: direct tasks ordered by estimated work (largest bucket first), narrow-kernel hand-overs appended after them
        unsigned int cur = base[CS];
        for (int b = 15; b >= 0; --b) { sbBase[b] = cur; cur += h[104 + b]; }
        curs[CS] = cur;
    }
    CK(cudaMemcpyAsync(cb + 16, curs, 16 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cb + 168, sbBase, 16 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(cb + 128, nbase, (size_t)nb * 4, cudaMemcpyHostToDevice, st));       // narrow bucket cursors only: 168.. are the strip buckets
    e = bbm_launch_msa_scatter(&P, (const unsigned char*)c->cls.p, cb, (int*)c->lists.p, (int*)c->nlist.p, st);
    if (e) return fail(BBM_E_CUDA, "msa_scatter_kernel launch", (cudaError_t)e);
    c->launches++;
    const int gRows0 = (h[51] && h[52] && !(c->bandwidth > 0 || c->ratio > 0.f)) ? (int)h[52] : max_rows, gCols0 = (h[51] && h[52] && !(c->bandwidth > 0 || c->ratio > 0.f)) ? (int)h[51] : max_cols;
    const long long gstride = bbm_generic_scratch_ints(gRows0, gCols0);       // per-task scratch (predecessor codes) from the class's own largest shape when no banded re-runs can follow
    long long chunk = (long long)((1ULL << 31) / ((size_t)gstride * 4));     // <= 2 GiB of row scratch at a time
    if (chunk < 1) chunk = 1;
    // shared-memory rows of the row-sequential kernel are sized from the largest shape actually in the class (the classifier tracks it):
    // a gapped reference is 500-700 columns, the upper bound 3002, and the difference is 2 versus 13 alignments resident per SM
    const int gRows = (h[51] && h[52]) ? (int)h[52] : max_rows, gCols = (h[51] && h[52]) ? (int)h[51] : max_cols;
    auto run_generic = [&](const int* list, long long n, cudaStream_t gs, const unsigned int* endPtr, unsigned int lbase) -> int {
        if (n <= 0) return BBM_OK;
        const bool classList = endPtr != nullptr;
        const long long ch = chunk > n ? n : chunk;
        if (c->gscratch.ensure((size_t)ch * (size_t)gstride * 4)) return fail(BBM_E_CUDA, "cudaMalloc generic scratch");
        for (long long done = 0; done < n; done += ch) {
            const int m = (int)((n - done) < ch ? (n - done) : ch);
            int e2 = bbm_launch_msa_generic(&P, list + done, m, (int*)c->gscratch.p, gstride, gs, classList ? gRows : max_rows, classList ? gCols : max_cols, endPtr, lbase + (unsigned int)done);
            if (e2) return fail(BBM_E_CUDA, "msa_generic_kernel launch", (cudaError_t)e2);
            c->launches++;
        }
        return BBM_OK;
    };
    if (nacc > 0) {
        int blocks = narrowBlocks;
        const long long need = ((long long)nacc + bbm_msa_narrow_threads() - 1) / bbm_msa_narrow_threads();
        if (need < blocks) blocks = (int)need;
        e = bbm_launch_msa_narrow(&P, (const int*)c->nlist.p, (int)nacc, cb, (unsigned long long*)c->nscratch.p, words, (int*)c->lists.p, blocks, useStrip, st);
        if (e) return fail(BBM_E_CUDA, "msa_narrow_kernel launch", (cudaError_t)e);
        c->launches++;
    }
    // shapes outside the tiled kernels (windows wider than 512 columns, reads longer than 606): a handful of long fills, one warp each.  Their
    // list is complete once the narrow kernel has handed its failures over, so they start here, on a side stream, beside the tiled and strip
    // kernels of this batch; the main stream joins them before anything reads the results.
    bool genericAside = false;
    if (h[nw] > 0 && d_dump == nullptr) {
        CK(cudaEventRecord(c->gev0, st));
        CK(cudaStreamWaitEvent(c->gstream, c->gev0, 0));
        int rcg = run_generic((const int*)c->lists.p + base[nw], h[nw], c->gstream, cb + 16 + nw, base[nw]);
        if (rcg) return rcg;
        CK(cudaEventRecord(c->gev1, c->gstream));
        genericAside = true;
    }
    typedef int (*launch_fn)(const MsaParams*, const int*, int, const unsigned int*, unsigned int, unsigned int*, int, int, cudaStream_t);
    static const launch_fn fns[7] = { bbm_launch_msa_tiled_w4, bbm_launch_msa_tiled_w5, bbm_launch_msa_tiled_w6, bbm_launch_msa_tiled_w8,
                                      bbm_launch_msa_tiled_w9, bbm_launch_msa_tiled_w12, bbm_launch_msa_tiled_w16 };
    for (int k = 0; k < nw; ++k) {
        if (!h[k]) continue;                       // no task of this width at all (narrow hand-overs included in h[k])
        int blocks = c->blocks;
        const long long needBlocks = ((long long)h[k] + wpb - 1) / wpb;
        if (needBlocks < blocks) blocks = (int)needBlocks;
        e = fns[k](&P, (const int*)c->lists.p + base[k], 0, cb + 16 + k, base[k], cb + 32 + k, blocks, d_dump != nullptr, st);
        if (e) return fail(BBM_E_CUDA, "msa_tiled_kernel launch", (cudaError_t)e);
        c->launches++;
    }
    if (useStrip && h[CS]) {
        // limited, un-banded fills (narrow-kernel hand-overs included): thread-per-alignment strip kernel.  Scratch = fixed part + one
        // block per alignment sized from its own rows/columns; the classifier summed those sizes (an upper bound: it includes the
        // alignments the narrow kernel has finished meanwhile).  If that does not fit the budget the list is processed in chunks.
        unsigned int cur = 0;
        CK(cudaMemcpyAsync(&cur, cb + 16 + CS, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const long long nstrip = (long long)cur - base[CS];
        const int sRows = max_rows < MAXR ? max_rows : MAXR;
        const int sCols = max_cols < bbm_msa_strip_max_cols() ? max_cols : bbm_msa_strip_max_cols();
        unsigned long long totalBytes = 0; memcpy(&totalBytes, &h[184], 8);
        const unsigned long long perMax = bbm_msa_strip_task_bytes(sRows, sCols);
        if (totalBytes > (unsigned long long)nstrip * perMax) totalBytes = (unsigned long long)nstrip * perMax;
        const int blocksMax = c->sms * bbm_msa_strip_blocks_per_sm();
        long long chunk = nstrip;
        if (totalBytes > c->strip_budget) { chunk = (long long)(c->strip_budget / perMax); if (chunk < 1024) chunk = 1024; if (chunk > nstrip) chunk = nstrip; }
        const unsigned long long poolBytes = (chunk == nstrip) ? totalBytes : (unsigned long long)chunk * perMax;
        for (long long start = 0; start < nstrip; start += chunk) {
            const int cnt = (int)((nstrip - start) < chunk ? (nstrip - start) : chunk);
            // the list is ordered longest-first: the first wave takes the expensive alignments, the cheap ones fill in behind them
            long long blocks = ((long long)cnt + 127) / 128;
            if (blocks > blocksMax) blocks = blocksMax;
            if (blocks < 1) blocks = 1;
            const size_t need = bbm_msa_strip_fixed_bytes(cnt, sRows, (int)blocks) + (size_t)poolBytes + 256;
            if (c->stripScratch.ensure(need)) return fail(BBM_E_CUDA, "cudaMalloc strip scratch");
            CK(cudaMemsetAsync(cb + 50, 0, 4, st));
            CK(cudaMemsetAsync(cb + 186, 0, 8, st));
            e = bbm_launch_msa_strip(&P, (const int*)c->lists.p + base[CS], cb + 16 + CS, base[CS], (int)start, cnt, sRows, c->stripScratch.p, c->stripScratch.cap,
                                     cb + 50, (unsigned long long*)(cb + 186), (int)blocks, c->strip_debug, (unsigned long long*)(cb + 220), st);
            if (e) return fail(BBM_E_CUDA, "msa_strip kernels launch", (cudaError_t)e);
            c->launches += 3;
        }
        c->strip_tasks += nstrip;
        if (c->strip_debug & 4) {
            unsigned long long z[2];
            CK(cudaMemcpyAsync(z, cb + 220, 16, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
            c->strip_units += z[0]; c->strip_lane_iters += z[1];
            CK(cudaMemsetAsync(cb + 220, 0, 16, st));
        }
    }
    int rc = BBM_OK;
    if (genericAside) CK(cudaStreamWaitEvent(st, c->gev1, 0));
    else { rc = run_generic((const int*)c->lists.p + base[nw], h[nw], st, cb + 16 + nw, base[nw]); if (rc) return rc; }
    if (c->bandwidth > 0 || c->ratio > 0.f) {
        unsigned int nover = 0;
        CK(cudaMemcpyAsync(&nover, cb + 48, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        rc = run_generic((const int*)c->overflow.p, nover, st, nullptr, 0);                   // banded right-edge misses
        if (rc) return rc;
        c->band_misses += nover;
    }
    CK(cudaEventRecord(c->ev1, st));
    unsigned int hend[16];
    CK(cudaMemcpyAsync(hend, cb + 16, 16 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    {   // bookkeeping: how many alignments the narrow kernel tried / handed over to the tiled kernels
        long long tiledTotal = 0;
        for (int k = 0; k < nw; ++k) tiledTotal += (long long)hend[k] - base[k];
        if (useStrip) tiledTotal += (long long)hend[CS] - base[CS];
        long long direct = 0;
        for (int k = 0; k < nw; ++k) direct += h[k];
        if (useStrip) direct += h[CS];
        direct -= nacc;                                   // tasks that went straight to a tiled list
        c->narrow_tried += nacc;
        c->narrow_handed_over += tiledTotal - direct;
        c->tasks_total += ntasks;
    }
    return BBM_OK;
}

extern "C" int bbm_set_option(bbm_ctx* c, const char* key, int value) {
    if (!c || !key) return fail(BBM_E_ARG, "bbm_set_option: null");
    if (!strcmp(key, "narrow")) { c->use_narrow = value; return BBM_OK; }
    if (!strcmp(key, "strip_min_tasks")) { c->strip_min_tasks = value; return BBM_OK; }
    if (!strcmp(key, "strip")) { c->use_strip = value; return BBM_OK; }
    if (!strcmp(key, "strip_debug")) { c->strip_debug = value; return BBM_OK; }
    if (!strcmp(key, "search_split")) { c->search_split = value; return BBM_OK; }
    if (!strcmp(key, "search_profile")) { c->search_prof = value; return BBM_OK; }
    if (!strcmp(key, "search_shared")) { c->search_shared = value; return BBM_OK; }     // 1 = walk arrays in shared memory when a batch has <=32 keys per read (A/B: measured slower)
    if (!strcmp(key, "strip_budget_mb")) { c->strip_budget = (size_t)value << 20; return BBM_OK; }
    return fail(BBM_E_ARG, "bbm_set_option: unknown key");
}
extern "C" int64_t bbm_get_stat(const bbm_ctx* c, const char* key) {
    if (!c || !key) return -1;
    if (!strcmp(key, "band_misses")) return c->band_misses;
    if (!strcmp(key, "launches")) return c->launches;
    if (!strcmp(key, "narrow_tried")) return c->narrow_tried;
    if (!strcmp(key, "narrow_handed_over")) return c->narrow_handed_over;
    if (!strcmp(key, "tasks_total")) return c->tasks_total;
    if (!strcmp(key, "strip_tasks")) return c->strip_tasks;
    if (!strcmp(key, "strip_units")) return (int64_t)c->strip_units;            // rows of 8 cells evaluated (with strip_debug bit 2)
    if (!strcmp(key, "strip_lane_iters")) return (int64_t)c->strip_lane_iters;  // lane-iterations of the evaluation phase: units/iters = lane utilisation
    if (!strncmp(key, "search_cycles_", 14) && key[14] >= '0' && key[14] <= '4') return (int64_t)c->search_cycles[key[14] - '0'];   // thread-cycles: total, filter, prescan, walk, extend
    if (!strcmp(key, "index_build_us")) return c->index_build_us;        // host wall time of the last bbm_index_build (reference already resident)
    return -1;
}

extern "C" int bbm_msa_batch_dev(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_msa_task* d_tasks,
                                 bbm_msa_out* d_outs, int64_t ntasks, int8_t* d_match_buf, const int64_t* d_match_off,
                                 int32_t max_rows, int32_t max_cols, void* stream, float* kernel_ms_out) {
    if (!c || !d_reads || !d_refs || !d_tasks || !d_outs) return fail(BBM_E_ARG, "bbm_msa_batch_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return run_msa(c, d_reads, d_refs, d_tasks, d_outs, ntasks, d_match_buf, d_match_off, max_rows, max_cols, st, kernel_ms_out, nullptr);
}

extern "C" int bbm_msa_batch_host(bbm_ctx* c, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs,
                                  const bbm_msa_task* tasks, bbm_msa_out* outs, int64_t ntasks,
                                  int8_t* match_buf, const int64_t* match_off) {
    if (!c || !reads || !d_refs || !tasks || !outs || reads_bytes < 0) return fail(BBM_E_ARG, "bbm_msa_batch_host: bad argument");
    if (ntasks <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    int max_rows = 1, max_cols = 1;
    for (int64_t i = 0; i < ntasks; ++i) {
        if (tasks[i].read_len > max_rows) max_rows = tasks[i].read_len;
        const int cols = tasks[i].ref_end - tasks[i].ref_start + 1;
        if (cols > max_cols) max_cols = cols;
    }
    const size_t tb = (size_t)ntasks * sizeof(bbm_msa_task), ob = (size_t)ntasks * sizeof(bbm_msa_out);
    const size_t mb = match_buf && match_off ? (size_t)match_off[ntasks] : 0, fb = (size_t)(ntasks + 1) * 8;
    if (c->d_reads.ensure((size_t)reads_bytes + 16) || c->d_tasks.ensure(tb) || c->d_outs.ensure(ob) ||
        c->d_match.ensure(mb + 16) || c->d_moff.ensure(fb))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(c->d_reads.p, reads, (size_t)reads_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_tasks.p, tasks, tb, cudaMemcpyHostToDevice, st));
    if (mb) CK(cudaMemcpyAsync(c->d_moff.p, match_off, fb, cudaMemcpyHostToDevice, st));
    int rc = run_msa(c, (const int8_t*)c->d_reads.p, d_refs, (const bbm_msa_task*)c->d_tasks.p, (bbm_msa_out*)c->d_outs.p, ntasks,
                     mb ? (int8_t*)c->d_match.p : nullptr, mb ? (const int64_t*)c->d_moff.p : nullptr, max_rows, max_cols, st, nullptr, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(outs, c->d_outs.p, ob, cudaMemcpyDeviceToHost, st));
    if (mb) CK(cudaMemcpyAsync(match_buf, c->d_match.p, mb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}

// =====================  gapped references (makeGref + coordinate translation, a15)  =====================
static const int GREF_LEN = 3002, GREF_STRIDE = 3008;      // grefbuffer = new byte[maxColumns+2] (MultiStateAligner11tsJNI.java:88)

static int run_msa_gapped(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_gapped_task* d_gt, const int32_t* d_gaps,
                          bbm_msa_out* d_outs, int64_t ntasks, int8_t* d_match, const int64_t* d_moff, cudaStream_t st, float* ms_out) {
    if (ntasks <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    if (c->grefPool.ensure((size_t)ntasks * GREF_STRIDE) || c->grefInfo.ensure((size_t)ntasks * sizeof(bbm_gref_info)) ||
        c->grefTasks.ensure((size_t)ntasks * sizeof(bbm_msa_task)))
        return fail(BBM_E_CUDA, "cudaMalloc gref pool");
    int e = bbm_launch_gref_build(d_refs, d_gt, d_gaps, ntasks, (int8_t*)c->grefPool.p, GREF_STRIDE, GREF_LEN, (bbm_gref_info*)c->grefInfo.p,
                                  (bbm_msa_task*)c->grefTasks.p, st);
    if (e) return fail(BBM_E_CUDA, "gref_build_kernel launch", (cudaError_t)e);
    c->launches++;
    int rc = run_msa(c, d_reads, d_refs, (const bbm_msa_task*)c->grefTasks.p, d_outs, ntasks, d_match, d_moff, MAXR, GREF_LEN, st, ms_out, nullptr);
    if (rc) return rc;
    e = bbm_launch_gref_translate(d_gt, ntasks, (const int8_t*)c->grefPool.p, GREF_STRIDE, (const bbm_gref_info*)c->grefInfo.p, d_outs, st);
    if (e) return fail(BBM_E_CUDA, "gref_translate_kernel launch", (cudaError_t)e);
    c->launches++;
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}

extern "C" int bbm_msa_gapped_batch_dev(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_gapped_task* d_tasks,
                                        const int32_t* d_gaps, bbm_msa_out* d_outs, int64_t ntasks, int8_t* d_match_buf,
                                        const int64_t* d_match_off, void* stream, float* kernel_ms_out) {
    if (!c || !d_reads || !d_refs || !d_tasks || !d_gaps || !d_outs) return fail(BBM_E_ARG, "bbm_msa_gapped_batch_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_msa_gapped(c, d_reads, d_refs, d_tasks, d_gaps, d_outs, ntasks, d_match_buf, d_match_off, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}

extern "C" int bbm_msa_gapped_batch_host(bbm_ctx* c, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs,
                                         const bbm_gapped_task* tasks, const int32_t* gaps, int64_t ngap_ints, bbm_msa_out* outs,
                                         int64_t ntasks, int8_t* match_buf, const int64_t* match_off) {
    if (!c || !reads || !d_refs || !tasks || !outs || reads_bytes < 0 || ngap_ints < 0 || (ngap_ints > 0 && !gaps))
        return fail(BBM_E_ARG, "bbm_msa_gapped_batch_host: bad argument");
    if (ntasks <= 0) return BBM_OK;
    for (int64_t i = 0; i < ntasks; ++i)
        if (tasks[i].ngaps < 0 || (tasks[i].ngaps > 0 && (tasks[i].gaps_off < 0 || (int64_t)tasks[i].gaps_off + tasks[i].ngaps > ngap_ints)))
            return fail(BBM_E_ARG, "bbm_msa_gapped_batch_host: gap array outside the gaps buffer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t tb = (size_t)ntasks * sizeof(bbm_gapped_task), ob = (size_t)ntasks * sizeof(bbm_msa_out);
    const size_t mb = match_buf && match_off ? (size_t)match_off[ntasks] : 0, fb = (size_t)(ntasks + 1) * 8;
    if (c->d_reads.ensure((size_t)reads_bytes + 16) || c->d_gtasks.ensure(tb) || c->d_outs.ensure(ob) || c->d_gaps.ensure((size_t)ngap_ints * 4 + 16) ||
        c->d_match.ensure(mb + 16) || c->d_moff.ensure(fb))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(c->d_reads.p, reads, (size_t)reads_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_gtasks.p, tasks, tb, cudaMemcpyHostToDevice, st));
    if (ngap_ints) CK(cudaMemcpyAsync(c->d_gaps.p, gaps, (size_t)ngap_ints * 4, cudaMemcpyHostToDevice, st));
    if (mb) CK(cudaMemcpyAsync(c->d_moff.p, match_off, fb, cudaMemcpyHostToDevice, st));
    int rc = run_msa_gapped(c, (const int8_t*)c->d_reads.p, d_refs, (const bbm_gapped_task*)c->d_gtasks.p, (const int32_t*)c->d_gaps.p,
                            (bbm_msa_out*)c->d_outs.p, ntasks, mb ? (int8_t*)c->d_match.p : nullptr, mb ? (const int64_t*)c->d_moff.p : nullptr, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(outs, c->d_outs.p, ob, cudaMemcpyDeviceToHost, st));
    if (mb) CK(cudaMemcpyAsync(match_buf, c->d_match.p, mb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}

// =====================  single-alignment twins of the reference's C entry points  =====================
// The kernels dump every evaluated cell into a dense [3][rows+1][cols+2] buffer; the host then replays the reference's
// *write pattern* (which cells fillLimitedX touches, its explicit subfloor writes and the BADoff reset of the last row:
// jni/MultiStateAligner11tsJNI.c:398-403, 451-456, 660-668) into the caller's `packed`, so the Java side's
// score2/traceback2 read exactly what the C would have left there.

static int single_fill(bbm_ctx* c, const int8_t* read, const int8_t* ref, int rows, int ref_length, int a, int b, int minScore,
                       bool limitedMode, int bandwidth, float ratio, std::vector<int>& dump, bbm_msa_out& out) {
    const int cols = b - a + 1;
    if (rows < 1 || cols < 1 || a < 0 || b >= ref_length) return fail(BBM_E_ARG, "fill: window outside the reference array");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t dumpInts = (size_t)3 * (rows + 1) * (cols + 2);
    if (c->d_reads.ensure((size_t)rows + cols + 64) || c->d_tasks.ensure(sizeof(bbm_msa_task)) || c->d_outs.ensure(sizeof(bbm_msa_out)) ||
        c->d_dump.ensure(dumpInts * 4))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    bbm_msa_task t;
    t.read_off = 0; t.ref_off = rows; t.read_len = rows; t.ref_len = cols; t.ref_start = 0; t.ref_end = cols - 1;
    t.min_score = minScore; t.flags = limitedMode ? BBM_TF_RAW_LIMITED : BBM_TF_RAW_UNLIMITED;
    CK(cudaMemcpyAsync(c->d_reads.p, read, (size_t)rows, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((char*)c->d_reads.p + rows, ref + a, (size_t)cols, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_tasks.p, &t, sizeof(t), cudaMemcpyHostToDevice, st));
    const int bw0 = c->bandwidth; const float r0 = c->ratio;
    c->bandwidth = bandwidth; c->ratio = ratio;
    int rc = run_msa(c, (const int8_t*)c->d_reads.p, (const int8_t*)c->d_reads.p, (const bbm_msa_task*)c->d_tasks.p, (bbm_msa_out*)c->d_outs.p, 1,
                     nullptr, nullptr, rows, cols, st, nullptr, (int*)c->d_dump.p);
    c->bandwidth = bw0; c->ratio = r0;
    if (rc) return rc;
    dump.resize(dumpInts);
    CK(cudaMemcpyAsync(dump.data(), c->d_dump.p, dumpInts * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&out, c->d_outs.p, sizeof(out), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (out.status != 0) return fail(out.status, "fill: kernel reported an error status");
    return BBM_OK;
}

extern "C" int bbm_fillUnlimited(bbm_ctx* c, const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
                                 int32_t refStartLoc, int32_t refEndLoc, int32_t* result4, int64_t* iterationsUnlimited,
                                 int32_t* packed, int32_t maxRows, int32_t maxColumns) {
    if (!c || !read || !ref || !result4 || !packed) return fail(BBM_E_ARG, "bbm_fillUnlimited: null pointer");
    const int rows = read_length, cols = refEndLoc - refStartLoc + 1;
    if (rows > maxRows || cols > maxColumns) return fail(BBM_E_SHAPE, "bbm_fillUnlimited: rows>maxRows or columns>maxColumns (the reference exit()s here)");
    std::vector<int> dump; bbm_msa_out out;
    int rc = single_fill(c, read, ref, rows, ref_length, refStartLoc, refEndLoc, 0, false, 0, 0.f, dump, out);
    if (rc) return rc;
    const long long stride = (long long)maxColumns + 1, plane = (long long)(maxRows + 1) * stride;
    const long long dplane = (long long)(rows + 1) * (cols + 2);
    for (int s = 0; s < 3; ++s)
        for (int r = 1; r <= rows; ++r)
            memcpy(packed + s * plane + r * stride + 1, dump.data() + s * dplane + (long long)r * (cols + 2) + 1, (size_t)cols * 4);
    for (int k = 0; k < 4; ++k) result4[k] = out.result[k];
    if (iterationsUnlimited) *iterationsUnlimited += out.iterations;
    return BBM_OK;
}

static inline bool host_defined(int ch) { return ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T' || ch == 'U' || ch == 'a' || ch == 'c' || ch == 'g' || ch == 't' || ch == 'u'; }

extern "C" int bbm_fillLimitedX(bbm_ctx* c, const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
                                int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* result5, int64_t* iterationsLimited,
                                int32_t* packed, int32_t maxRows, int32_t maxColumns, int32_t bandwidth, float bandwidthRatio,
                                int32_t* vertLimit, int32_t* horizLimit) {
    if (!c || !read || !ref || !result5 || !packed) return fail(BBM_E_ARG, "bbm_fillLimitedX: null pointer");
    const int rows = read_length, cols = refEndLoc - refStartLoc + 1;
    if (rows > maxRows || cols > maxColumns) return fail(BBM_E_SHAPE, "bbm_fillLimitedX: rows>maxRows or columns>maxColumns");
    std::vector<int> dump; bbm_msa_out out;
    int rc = single_fill(c, read, ref, rows, ref_length, refStartLoc, refEndLoc, minScore, true, bandwidth, bandwidthRatio, dump, out);
    if (rc) return rc;
    const long long stride = (long long)maxColumns + 1, plane = (long long)(maxRows + 1) * stride;
    const long long dstride = cols + 2, dplane = (long long)(rows + 1) * dstride;
    const int minScore_off = (int)((unsigned)minScore << TBITS);
    const int maxGain = (rows - 1) * P_MATCH2 + P_MATCH;
    const int floor_ = minScore_off - maxGain, subfloor = floor_ - 5 * P_MATCH2;
    int halfband = 0;
    if (!(bandwidth < 1 && bandwidthRatio <= 0.f)) {
        const int x = bandwidth < 1 ? 9999999 : bandwidth, y = bandwidthRatio <= 0.f ? 9999999 : 8 + (int)(rows * bandwidthRatio);
        const int m = x < y ? x : y, n = cols - rows + 8;
        halfband = (m > n ? m : n) / 2;
    }
    // vertLimit / horizLimit are outputs of the reference call too (jni/...JNI.c:413-438)
    if (vertLimit) {
        vertLimit[rows] = minScore_off; bool pd = false;
        for (int i = rows - 1; i >= 0; --i) { const bool d = host_defined(read[i]); const int v = vertLimit[i + 1] - (d ? (pd ? P_MATCH2 : P_MATCH) : 0); vertLimit[i] = v > floor_ ? v : floor_; pd = d; }
    }
    if (horizLimit) {
        horizLimit[cols] = minScore_off; bool pd = false;
        for (int i = cols - 1; i >= 0; --i) {
            const int ch = ref[refStartLoc + i]; const bool d = host_defined(ch);
            const int v = horizLimit[i + 1] - (d ? (pd ? P_MATCH2 : P_MATCH) : ((pd && ch == '-') ? P_DEL : 0));
            horizLimit[i] = v > floor_ ? v : floor_; pd = d;
        }
    }
    // replay the write pattern
    for (int s = 0; s < 3; ++s) for (int i = 1; i <= cols; ++i) packed[s * plane + (long long)rows * stride + i] = BADOFF;
    auto cellGood = [&](int r, int col) -> bool {
        const long long idx = (long long)r * dstride + col;
        return (dump[idx] & SMASK) != subfloor || (dump[dplane + idx] & SMASK) != subfloor || (dump[2 * dplane + idx] & SMASK) != subfloor;
    };
    int minGood = 1, maxGood = cols;
    for (int row = 1; row <= rows; ++row) {
        const int colStart = halfband < 1 ? minGood : (minGood > row - halfband ? minGood : row - halfband);
        const int colStop = halfband < 1 ? maxGood : (maxGood < row + halfband * 2 - 1 ? maxGood : row + halfband * 2 - 1);
        minGood = -1; maxGood = -2;
        if (colStart < 0 || colStop < colStart) break;
        if (colStart > 1) for (int s = 0; s < 3; ++s) packed[s * plane + (long long)row * stride + colStart - 1] = subfloor;
        for (int col = colStart; col <= cols; ++col) {
            for (int s = 0; s < 3; ++s) packed[s * plane + (long long)row * stride + col] = dump[s * dplane + (long long)row * dstride + col];
            if (cellGood(row, col)) { maxGood = col; if (minGood < 0) minGood = col; }
            if (col >= colStop) {
                if (col > colStop && (maxGood < col || halfband > 0)) break;
                if (row > 1) for (int s = 0; s < 3; ++s) packed[s * plane + (long long)(row - 1) * stride + col + 1] = subfloor;
            }
        }
    }
    for (int k = 0; k < 5; ++k) result5[k] = out.result[k];
    if (iterationsLimited) *iterationsLimited += out.iterations;
    return BBM_OK;
}

// Integer / DPX pipe peak: lane-ops per second of instruction kind `kind` (0 IADD3, 1 LOP3, 2 VIMNMX3, 3 VIADDMNMX, 4 IMAD,
// 5 half IMAD + half LOP3, 6 compare+select).  8 independent chains x 256 threads x 8 blocks per SM.
extern "C" int bbm_int_peak(bbm_ctx* c, int kind, double* gops_out) {
    if (!c || !gops_out) return fail(BBM_E_ARG, "bbm_int_peak: null");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    if (c->d_dump.ensure(64)) return fail(BBM_E_CUDA, "cudaMalloc");
    const int blocks = c->sms * 8, iters = 1024;
    cudaStream_t st = c->stream;
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        CK(cudaEventRecord(c->ev0, st));
        int e = bbm_launch_peak(kind, blocks, iters, (int*)c->d_dump.p, st);
        if (e) return fail(BBM_E_CUDA, "peak kernel launch", (cudaError_t)e);
        c->launches++;
        CK(cudaEventRecord(c->ev1, st));
        CK(cudaStreamSynchronize(st));
        float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    const double ops = (double)blocks * 256.0 * iters * 8.0;
    *gops_out = ops / (best * 1e-3) / 1e9;
    return BBM_OK;
}

// =====================  BandedAligner  =====================
static int run_banded(bbm_ctx* c, const int8_t* dq, const int8_t* dr, const bbm_band_task* dt, bbm_band_out* dout, int64_t n,
                      cudaStream_t st, float* ms_out) {
    if (n <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    unsigned int* cb = (unsigned int*)c->counters.p;
    CK(cudaMemsetAsync(cb + 200, 0, 4, st));
    CK(cudaEventRecord(c->ev0, st));
    int blocks = c->sms * 8;
    const long long need = (n + 3) / 4;
    if (need < blocks) blocks = (int)need;
    int e = bbm_launch_banded(dq, dr, dt, dout, n, cb + 200, blocks, st);
    if (e) return fail(BBM_E_CUDA, "banded_kernel launch", (cudaError_t)e);
    c->launches++;
    CK(cudaEventRecord(c->ev1, st));
    CK(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    return BBM_OK;
}

extern "C" int bbm_banded_batch_dev(bbm_ctx* c, const int8_t* d_queries, const int8_t* d_refs, const bbm_band_task* d_tasks,
                                    bbm_band_out* d_outs, int64_t ntasks, void* stream, float* kernel_ms_out) {
    if (!c || !d_queries || !d_refs || !d_tasks || !d_outs) return fail(BBM_E_ARG, "bbm_banded_batch_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_banded(c, d_queries, d_refs, d_tasks, d_outs, ntasks, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}

extern "C" int bbm_banded_batch_host(bbm_ctx* c, const int8_t* queries, int64_t query_bytes, const int8_t* refs, int64_t ref_bytes,
                                     const bbm_band_task* tasks, bbm_band_out* outs, int64_t ntasks) {
    if (!c || !queries || !refs || !tasks || !outs) return fail(BBM_E_ARG, "bbm_banded_batch_host: null pointer");
    if (ntasks <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t tb = (size_t)ntasks * sizeof(bbm_band_task), ob = (size_t)ntasks * sizeof(bbm_band_out);
    if (c->d_reads.ensure((size_t)query_bytes + 16) || c->d_refs2.ensure((size_t)ref_bytes + 16) || c->d_tasks.ensure(tb) || c->d_outs.ensure(ob))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(c->d_reads.p, queries, (size_t)query_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_refs2.p, refs, (size_t)ref_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_tasks.p, tasks, tb, cudaMemcpyHostToDevice, st));
    int rc = run_banded(c, (const int8_t*)c->d_reads.p, (const int8_t*)c->d_refs2.p, (const bbm_band_task*)c->d_tasks.p, (bbm_band_out*)c->d_outs.p, ntasks, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(outs, c->d_outs.p, ob, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}

// =====================  KeyRing seeding  =====================
static int run_seed(bbm_ctx* c, const int8_t* db, const int8_t* dq, const int64_t* doff, int64_t nreads, int max_len, const bbm_seed_cfg* cfg,
                    int maxKeys, int* dn, int* dof, int* dk, int* dks, int8_t* dbs, int* dofM, int* dkM, cudaStream_t st, float* ms_out) {
    if (nreads <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    if (!c->seed_tables) {
        // QualityTools.PROB_ERROR / PROB_CORRECT / PROB_CORRECT_INVERSE (current/align2/QualityTools.java:475-480, 519-539)
        float pc[127], pci[127];
        for (int i = 0; i < 127; ++i) { float pe = (float)pow(10.0, 0 - .1 * i); if (i == 0) pe = .8f; pc[i] = 1 - pe; pci[i] = 1 / pc[i]; }
        int e = bbm_seed_upload_tables(pc, pci);
        if (e) return fail(BBM_E_CUDA, "seed tables upload", (cudaError_t)e);
        c->seed_tables = true;
    }
    const int T = bbm_seed_threads();
    int blocks = c->sms * 4;
    const long long need = (nreads + T - 1) / T;
    if (need < blocks) blocks = (int)need;
    const int maxProbLen = max_len - cfg->keylen + 1 > 1 ? max_len - cfg->keylen + 1 : 1;
    if (c->seedScratch.ensure((size_t)blocks * T * (size_t)maxProbLen * 4)) return fail(BBM_E_CUDA, "cudaMalloc seed scratch");
    unsigned int* cb = (unsigned int*)c->counters.p;
    CK(cudaMemsetAsync(cb + 201, 0, 4, st));
    CK(cudaEventRecord(c->ev0, st));
    int e = bbm_launch_seed(db, dq, (const long long*)doff, nreads, cfg, maxKeys, dn, dof, dk, dks, dbs, (float*)c->seedScratch.p, blocks, maxProbLen, cb + 201, st);
    if (e) return fail(BBM_E_CUDA, "seed_kernel launch", (cudaError_t)e);
    c->launches++;
    if (dofM && dkM) {
        e = bbm_launch_seed_reverse(dn, dof, dk, (const long long*)doff, nreads, maxKeys, cfg->keylen, dofM, dkM, st);
        if (e) return fail(BBM_E_CUDA, "seed_reverse_kernel launch", (cudaError_t)e);
        c->launches++;
    }
    CK(cudaEventRecord(c->ev1, st));
    CK(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    return BBM_OK;
}

extern "C" int bbm_seed_batch_dev(bbm_ctx* c, const int8_t* d_bases, const int8_t* d_quality, const int64_t* d_read_off, int64_t nreads,
                                  int32_t max_read_len, const bbm_seed_cfg* cfg, int32_t maxKeys, int32_t* d_nkeys, int32_t* d_offsets,
                                  int32_t* d_keys, int32_t* d_keyScores, int8_t* d_baseScores, int32_t* d_offsetsM, int32_t* d_keysM,
                                  void* stream, float* kernel_ms_out) {
    if (!c || !d_bases || !d_read_off || !cfg || !d_nkeys || !d_offsets || !d_keys || !d_keyScores || !d_baseScores || maxKeys < 1)
        return fail(BBM_E_ARG, "bbm_seed_batch_dev: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_seed(c, d_bases, d_quality, d_read_off, nreads, max_read_len, cfg, maxKeys, d_nkeys, d_offsets, d_keys, d_keyScores, d_baseScores,
                    d_offsetsM, d_keysM, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}

extern "C" int bbm_seed_batch_host(bbm_ctx* c, const int8_t* bases, const int8_t* quality, const int64_t* read_off, int64_t nreads,
                                   const bbm_seed_cfg* cfg, int32_t maxKeys, int32_t* nkeys, int32_t* offsets, int32_t* keys,
                                   int32_t* keyScores, int8_t* baseScores, int32_t* offsetsM, int32_t* keysM) {
    if (!c || !bases || !read_off || !cfg || !nkeys || !offsets || !keys || !keyScores || !baseScores || maxKeys < 1)
        return fail(BBM_E_ARG, "bbm_seed_batch_host: bad argument");
    if (nreads <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t nb = (size_t)read_off[nreads], kb = (size_t)nreads * maxKeys * 4;
    int max_len = 1;
    for (int64_t i = 0; i < nreads; ++i) { const int l = (int)(read_off[i + 1] - read_off[i]); if (l > max_len) max_len = l; }
    DevBuf* B = c->d_seed;   // 0 bases, 1 qual, 2 off, 3 nkeys, 4 offsets|keys|scores, 5 baseScores, 6 offsetsM|keysM
    if (B[0].ensure(nb + 32) || (quality && B[1].ensure(nb + 32)) || B[2].ensure((size_t)(nreads + 1) * 8) || B[3].ensure((size_t)nreads * 4) ||
        B[4].ensure(3 * kb) || B[5].ensure(nb + 32) || B[6].ensure(2 * kb))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(B[0].p, bases, nb, cudaMemcpyHostToDevice, st));
    if (quality) CK(cudaMemcpyAsync(B[1].p, quality, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[2].p, read_off, (size_t)(nreads + 1) * 8, cudaMemcpyHostToDevice, st));
    int* d4 = (int*)B[4].p; int* d6 = (int*)B[6].p;
    const bool rev = offsetsM && keysM;
    int rc = run_seed(c, (const int8_t*)B[0].p, quality ? (const int8_t*)B[1].p : nullptr, (const int64_t*)B[2].p, nreads, max_len, cfg, maxKeys,
                      (int*)B[3].p, d4, d4 + (size_t)nreads * maxKeys, d4 + 2 * (size_t)nreads * maxKeys, (int8_t*)B[5].p,
                      rev ? d6 : nullptr, rev ? d6 + (size_t)nreads * maxKeys : nullptr, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(nkeys, B[3].p, (size_t)nreads * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(offsets, d4, kb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(keys, d4 + (size_t)nreads * maxKeys, kb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(keyScores, d4 + 2 * (size_t)nreads * maxKeys, kb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(baseScores, B[5].p, nb, cudaMemcpyDeviceToHost, st));
    if (rev) { CK(cudaMemcpyAsync(offsetsM, d6, kb, cudaMemcpyDeviceToHost, st)); CK(cudaMemcpyAsync(keysM, d6 + (size_t)nreads * maxKeys, kb, cudaMemcpyDeviceToHost, st)); }
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}

// =====================  ungapped site scoring  =====================
static int run_noindel(bbm_ctx* c, const int8_t* dr, const int8_t* dref, const bbm_noindel_task* dt, int* ds, int8_t* dm, const int64_t* dmo,
                       int64_t n, cudaStream_t st, float* ms_out) {
    if (n <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    CK(cudaEventRecord(c->ev0, st));
    int e = bbm_launch_noindel(dr, dref, dt, ds, dm, (const long long*)dmo, n, st);
    if (e) return fail(BBM_E_CUDA, "noindel_kernel launch", (cudaError_t)e);
    c->launches++;
    CK(cudaEventRecord(c->ev1, st));
    CK(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    return BBM_OK;
}
extern "C" int bbm_noindel_batch_dev(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_noindel_task* d_tasks, int32_t* d_scores,
                                     int8_t* d_match_buf, const int64_t* d_match_off, int64_t ntasks, void* stream, float* kernel_ms_out) {
    if (!c || !d_reads || !d_refs || !d_tasks || !d_scores) return fail(BBM_E_ARG, "bbm_noindel_batch_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_noindel(c, d_reads, d_refs, d_tasks, d_scores, d_match_buf, d_match_off, ntasks, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}
extern "C" int bbm_noindel_batch_host(bbm_ctx* c, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs, const bbm_noindel_task* tasks,
                                      int32_t* scores, int8_t* match_buf, const int64_t* match_off, int64_t ntasks) {
    if (!c || !reads || !d_refs || !tasks || !scores) return fail(BBM_E_ARG, "bbm_noindel_batch_host: null pointer");
    if (ntasks <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t tb = (size_t)ntasks * sizeof(bbm_noindel_task), sb = (size_t)ntasks * 4;
    const size_t mb = match_buf && match_off ? (size_t)match_off[ntasks] : 0, fb = (size_t)(ntasks + 1) * 8;
    if (c->d_reads.ensure((size_t)reads_bytes + 16) || c->d_tasks.ensure(tb) || c->d_outs.ensure(sb) || c->d_match.ensure(mb + 16) || c->d_moff.ensure(fb))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(c->d_reads.p, reads, (size_t)reads_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_tasks.p, tasks, tb, cudaMemcpyHostToDevice, st));
    if (mb) { CK(cudaMemcpyAsync(c->d_moff.p, match_off, fb, cudaMemcpyHostToDevice, st)); CK(cudaMemsetAsync(c->d_match.p, 0, mb, st)); }
    int rc = run_noindel(c, (const int8_t*)c->d_reads.p, d_refs, (const bbm_noindel_task*)c->d_tasks.p, (int*)c->d_outs.p,
                         mb ? (int8_t*)c->d_match.p : nullptr, mb ? (const int64_t*)c->d_moff.p : nullptr, ntasks, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(scores, c->d_outs.p, sb, cudaMemcpyDeviceToHost, st));
    if (mb) CK(cudaMemcpyAsync(match_buf, c->d_match.p, mb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}

// =====================  tip-deletion search and mate rescue scans (rescue.cu)  =====================
extern "C" int bbm_launch_tipdel(const int8_t* reads, const int8_t* refs, const bbm_tipdel_task* tasks, long long n, const bbm_tipdel_cfg* cfg,
                                 bbm_tipdel_out* outs, cudaStream_t st);
extern "C" int bbm_launch_rescue(const int8_t* reads, const int8_t* refs, const bbm_rescue_task* tasks, long long n, const bbm_rescue_cfg* cfg,
                                 bbm_rescue_out* outs, cudaStream_t st);
template <class Task, class Cfg, class Out, class Launch>
static int run_scan(bbm_ctx* c, const char* what, Launch launch, const int8_t* dr, const int8_t* dref, const Task* dt, int64_t n, const Cfg* cfg,
                    Out* dout, cudaStream_t st, float* ms_out) {
    if (n <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    if (ms_out) CK(cudaEventRecord(c->ev0, st));
    int e = launch(dr, dref, dt, (long long)n, cfg, dout, st);
    if (e) return fail(BBM_E_CUDA, what, (cudaError_t)e);
    c->launches += 1;
    if (ms_out) { CK(cudaEventRecord(c->ev1, st)); CK(cudaEventSynchronize(c->ev1)); float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    return BBM_OK;
}
template <class Task, class Cfg, class Out, class Launch>
static int run_scan_host(bbm_ctx* c, const char* what, Launch launch, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs, const Task* tasks,
                         int64_t n, const Cfg* cfg, Out* outs) {
    if (n <= 0) return BBM_OK;
    for (int64_t i = 0; i < n; ++i)
        if (tasks[i].read_len < 0 || tasks[i].read_off < 0 || tasks[i].read_off + tasks[i].read_len > reads_bytes || tasks[i].ref_off < 0 || tasks[i].ref_len < 0)
            return fail(BBM_E_ARG, "scan task outside the read buffer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t tb = (size_t)n * sizeof(Task), ob = (size_t)n * sizeof(Out);
    if (c->d_reads.ensure((size_t)reads_bytes + 16) || c->d_tasks.ensure(tb) || c->d_outs.ensure(ob)) return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(c->d_reads.p, reads, (size_t)reads_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->d_tasks.p, tasks, tb, cudaMemcpyHostToDevice, st));
    int rc = run_scan(c, what, launch, (const int8_t*)c->d_reads.p, d_refs, (const Task*)c->d_tasks.p, n, cfg, (Out*)c->d_outs.p, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(outs, c->d_outs.p, ob, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}
extern "C" int bbm_tipdel_batch_dev(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_tipdel_task* d_tasks, int64_t n,
                                    const bbm_tipdel_cfg* cfg, bbm_tipdel_out* d_outs, void* stream, float* kernel_ms_out) {
    if (!c || !d_reads || !d_refs || !d_tasks || !cfg || !d_outs) return fail(BBM_E_ARG, "bbm_tipdel_batch_dev: null pointer");
    if (cfg->max_tiplen < 3 || cfg->max_tiplen > 32) return fail(BBM_E_ARG, "bbm_tipdel: max_tiplen must be in 3..32");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_scan(c, "tipdel_kernel launch", bbm_launch_tipdel, d_reads, d_refs, d_tasks, n, cfg, d_outs, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}
extern "C" int bbm_tipdel_batch_host(bbm_ctx* c, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs, const bbm_tipdel_task* tasks, int64_t n,
                                     const bbm_tipdel_cfg* cfg, bbm_tipdel_out* outs) {
    if (!c || !reads || !d_refs || !tasks || !cfg || !outs) return fail(BBM_E_ARG, "bbm_tipdel_batch_host: null pointer");
    if (cfg->max_tiplen < 3 || cfg->max_tiplen > 32) return fail(BBM_E_ARG, "bbm_tipdel: max_tiplen must be in 3..32");
    return run_scan_host(c, "tipdel_kernel launch", bbm_launch_tipdel, reads, reads_bytes, d_refs, tasks, n, cfg, outs);
}
extern "C" int bbm_rescue_batch_dev(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_rescue_task* d_tasks, int64_t n,
                                    const bbm_rescue_cfg* cfg, bbm_rescue_out* d_outs, void* stream, float* kernel_ms_out) {
    if (!c || !d_reads || !d_refs || !d_tasks || !cfg || !d_outs) return fail(BBM_E_ARG, "bbm_rescue_batch_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_scan(c, "rescue_kernel launch", bbm_launch_rescue, d_reads, d_refs, d_tasks, n, cfg, d_outs, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}
extern "C" int bbm_rescue_batch_host(bbm_ctx* c, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs, const bbm_rescue_task* tasks, int64_t n,
                                     const bbm_rescue_cfg* cfg, bbm_rescue_out* outs) {
    if (!c || !reads || !d_refs || !tasks || !cfg || !outs) return fail(BBM_E_ARG, "bbm_rescue_batch_host: null pointer");
    return run_scan_host(c, "rescue_kernel launch", bbm_launch_rescue, reads, reads_bytes, d_refs, tasks, n, cfg, outs);
}

// =====================  per-read site-list policies (sitelist.cu)  =====================
extern "C" int bbm_sitelist_max_cap();
extern "C" int bbm_launch_sitelist(int op, bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const int8_t* basesP,
                                   const int8_t* basesM, const int8_t* refs, const long long* chrom_off, const bbm_policy_cfg* cfg, bbm_read_out* out,
                                   cudaStream_t st);
extern "C" int bbm_launch_sitelist_from_search(const bbm_search_head* heads, const bbm_site* sites, long long nreads, int maxSites, bbm_ss* lists,
                                               int* nss, int cap, cudaStream_t st);
static int sitelist_args(int op, int cap, const bbm_policy_cfg* cfg) {
    if (op != BBM_SL_TRIM && op != BBM_SL_NOINDEL && op != BBM_SL_FINAL) return fail(BBM_E_ARG, "bbm_sitelist: unknown op");
    if (cap < 1 || cap > bbm_sitelist_max_cap()) return fail(BBM_E_ARG, "bbm_sitelist: cap must be in 1..64");
    if (!cfg || cfg->min_trim_sites_to_retain < 1 || cfg->max_trim_sites_to_retain <= cfg->min_trim_sites_to_retain) return fail(BBM_E_ARG, "bbm_sitelist: bad policy cfg");
    return BBM_OK;
}
extern "C" int bbm_sitelist_from_search_dev(bbm_ctx* c, const bbm_search_head* d_heads, const bbm_site* d_sites, int64_t nreads, int32_t max_sites,
                                            bbm_ss* d_lists, int32_t* d_nss, int32_t cap, void* stream) {
    if (!c || !d_heads || !d_sites || !d_lists || !d_nss || max_sites < 1 || cap < 1) return fail(BBM_E_ARG, "bbm_sitelist_from_search_dev: bad argument");
    if (nreads <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    int e = bbm_launch_sitelist_from_search(d_heads, d_sites, nreads, max_sites, d_lists, d_nss, cap, stream ? (cudaStream_t)stream : c->stream);
    if (e) return fail(BBM_E_CUDA, "sitelist_from_search_kernel launch", (cudaError_t)e);
    c->launches++;
    return BBM_OK;
}
extern "C" int bbm_sitelist_batch_dev(bbm_ctx* c, int32_t op, bbm_ss* d_lists, int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                      const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_refs, const int64_t* d_chrom_off,
                                      const bbm_policy_cfg* cfg, bbm_read_out* d_out, void* stream, float* kernel_ms_out) {
    if (!c || !d_lists || !d_nss || !d_read_off || !d_out) return fail(BBM_E_ARG, "bbm_sitelist_batch_dev: null pointer");
    if (int rc = sitelist_args(op, cap, cfg)) return rc;
    if (op == BBM_SL_NOINDEL && (!d_basesP || !d_basesM || !d_refs || !d_chrom_off)) return fail(BBM_E_ARG, "bbm_sitelist_batch_dev: BBM_SL_NOINDEL needs reads and reference");
    if (nreads <= 0) { if (kernel_ms_out) *kernel_ms_out = 0.f; return BBM_OK; }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (kernel_ms_out) CK(cudaEventRecord(c->ev0, st));
    int e = bbm_launch_sitelist(op, d_lists, d_nss, nreads, cap, (const long long*)d_read_off, d_basesP, d_basesM, d_refs, (const long long*)d_chrom_off, cfg, d_out, st);
    if (e) return fail(BBM_E_CUDA, "sitelist_kernel launch", (cudaError_t)e);
    c->launches++;
    if (kernel_ms_out) { CK(cudaEventRecord(c->ev1, st)); CK(cudaEventSynchronize(c->ev1)); float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *kernel_ms_out = ms; }
    return BBM_OK;
}
extern "C" int bbm_sitelist_batch_host(bbm_ctx* c, int32_t op, bbm_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int64_t* read_off,
                                       const int8_t* basesP, const int8_t* basesM, const int8_t* d_refs, const int64_t* chrom_off, int32_t nchroms,
                                       const bbm_policy_cfg* cfg, bbm_read_out* out) {
    if (!c || !lists || !nss || !read_off || !out) return fail(BBM_E_ARG, "bbm_sitelist_batch_host: null pointer");
    if (int rc = sitelist_args(op, cap, cfg)) return rc;
    if (op == BBM_SL_NOINDEL && (!basesP || !basesM || !d_refs || !chrom_off || nchroms < 1)) return fail(BBM_E_ARG, "bbm_sitelist_batch_host: BBM_SL_NOINDEL needs reads and reference");
    if (nreads <= 0) return BBM_OK;
    for (int64_t r = 0; r < nreads; ++r) {
        if (nss[r] < 0 || nss[r] > cap || read_off[r + 1] < read_off[r]) return fail(BBM_E_ARG, "bbm_sitelist_batch_host: list length outside 0..cap");
        if (op == BBM_SL_NOINDEL) for (int i = 0; i < nss[r]; ++i) { const bbm_ss& s = lists[r * cap + i]; if (s.chrom < 1 || s.chrom > nchroms) return fail(BBM_E_ARG, "bbm_sitelist_batch_host: chromosome out of range"); }
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t lb = (size_t)nreads * cap * sizeof(bbm_ss), nb = (size_t)nreads * 4, ob = (size_t)nreads * sizeof(bbm_read_out), fb = (size_t)(nreads + 1) * 8;
    const size_t rb = (size_t)read_off[nreads], cb = (size_t)(nchroms + 1) * 8;
    DevBuf L_, N_, O_, F_, P_, M_, C_;
    const bool need = op == BBM_SL_NOINDEL;
    if (L_.ensure(lb) || N_.ensure(nb) || O_.ensure(ob) || F_.ensure(fb) || (need && (P_.ensure(rb + 16) || M_.ensure(rb + 16) || C_.ensure(cb)))) return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(L_.p, lists, lb, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(N_.p, nss, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(F_.p, read_off, fb, cudaMemcpyHostToDevice, st));
    if (need) {
        CK(cudaMemcpyAsync(P_.p, basesP, rb, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(M_.p, basesM, rb, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(C_.p, chrom_off, cb, cudaMemcpyHostToDevice, st));
    }
    int e = bbm_launch_sitelist(op, (bbm_ss*)L_.p, (int*)N_.p, nreads, cap, (const long long*)F_.p, (const int8_t*)P_.p, (const int8_t*)M_.p, d_refs, (const long long*)C_.p, cfg, (bbm_read_out*)O_.p, st);
    int rc = BBM_OK;
    if (e) rc = fail(BBM_E_CUDA, "sitelist_kernel launch", (cudaError_t)e);
    else {
        c->launches++;
        cudaError_t ce = cudaMemcpyAsync(lists, L_.p, lb, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(nss, N_.p, nb, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(out, O_.p, ob, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) rc = fail(BBM_E_CUDA, "sitelist copy back", ce);
    }
    L_.release(); N_.release(); O_.release(); F_.release(); P_.release(); M_.release(); C_.release();
    return rc;
}

extern "C" int bbm_launch_sitelist_tipdel(bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off, const int8_t* basesP,
                                          const int8_t* basesM, const int8_t* quality, const int8_t* refs, const long long* chrom_off,
                                          const int* chrom_min_index, const bbm_tipdel_cfg* tc, bbm_read_out* out, cudaStream_t st);
extern "C" int bbm_sitelist_tipdel_dev(bbm_ctx* c, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                       const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_quality, const int8_t* d_refs, const int64_t* d_chrom_off,
                                       const int32_t* d_chrom_min_index, const bbm_tipdel_cfg* cfg, bbm_read_out* d_out, void* stream, float* kernel_ms_out) {
    if (!c || !d_lists || !d_nss || !d_read_off || !d_basesP || !d_basesM || !d_refs || !d_chrom_off || !cfg || !d_out) return fail(BBM_E_ARG, "bbm_sitelist_tipdel_dev: null pointer");
    if (cfg->max_tiplen < 3 || cfg->max_tiplen > 32 || cap < 1 || cap > bbm_sitelist_max_cap()) return fail(BBM_E_ARG, "bbm_sitelist_tipdel_dev: bad cfg or cap");
    if (nreads <= 0) { if (kernel_ms_out) *kernel_ms_out = 0.f; return BBM_OK; }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    if (kernel_ms_out) CK(cudaEventRecord(c->ev0, st));
    int e = bbm_launch_sitelist_tipdel(d_lists, d_nss, nreads, cap, (const long long*)d_read_off, d_basesP, d_basesM, d_quality, d_refs, (const long long*)d_chrom_off,
                                       d_chrom_min_index, cfg, d_out, st);
    if (e) return fail(BBM_E_CUDA, "sitelist_tipdel_kernel launch", (cudaError_t)e);
    c->launches++;
    if (kernel_ms_out) { CK(cudaEventRecord(c->ev1, st)); CK(cudaEventSynchronize(c->ev1)); float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *kernel_ms_out = ms; }
    return BBM_OK;
}

extern "C" int bbm_launch_sitelist_bounds(bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const int* chrom_max_index,
                                          const int* scaf_off, const int* scaf_loc, int pad, int sam_out, int expected_len_limit, bbm_read_out* out, cudaStream_t st);
extern "C" int bbm_sitelist_bounds_dev(bbm_ctx* c, bbm_ss* d_lists, int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                       const int32_t* d_chrom_max_index, const int32_t* d_scaf_off, const int32_t* d_scaf_loc, int32_t inter_scaffold_padding,
                                       int32_t sam_out, int32_t expected_len_limit, bbm_read_out* d_out, void* stream) {
    if (!c || !d_lists || !d_nss || !d_read_off || !d_chrom_max_index || !d_out) return fail(BBM_E_ARG, "bbm_sitelist_bounds_dev: null pointer");
    if ((d_scaf_off == nullptr) != (d_scaf_loc == nullptr) || cap < 1 || cap > bbm_sitelist_max_cap() || expected_len_limit < 1) return fail(BBM_E_ARG, "bbm_sitelist_bounds_dev: bad argument");
    if (nreads <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    int e = bbm_launch_sitelist_bounds(d_lists, d_nss, nreads, cap, (const long long*)d_read_off, d_chrom_max_index, d_scaf_off, d_scaf_loc, inter_scaffold_padding,
                                       sam_out, expected_len_limit, d_out, stream ? (cudaStream_t)stream : c->stream);
    if (e) return fail(BBM_E_CUDA, "sitelist_bounds_kernel launch", (cudaError_t)e);
    c->launches++;
    return BBM_OK;
}

extern "C" int bbm_launch_sitelist_cz3(bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const bbm_policy_cfg* cfg,
                                       int ambiguous_toss, bbm_read_out* io, cudaStream_t st);
extern "C" int bbm_launch_sitelist_tip_penalty(bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off, const int8_t* bases,
                                               const int8_t* match, const long long* match_off, const bbm_read_out* flags, int tiplen, int* penalty,
                                               int* status, cudaStream_t st);
extern "C" int bbm_sitelist_clearzone3_dev(bbm_ctx* c, bbm_ss* d_lists, int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                           const bbm_policy_cfg* cfg, int32_t ambiguous_toss, bbm_read_out* d_io, void* stream) {
    if (!c || !cfg) return fail(BBM_E_ARG, "bbm_sitelist_clearzone3_dev: null pointer");
    if (cap < 1 || cap > bbm_sitelist_max_cap()) return fail(BBM_E_ARG, "bbm_sitelist_clearzone3_dev: cap must be in 1..64");
    if (nreads <= 0) return BBM_OK;                                                  // an empty batch carries no buffers
    if (!d_lists || !d_nss || !d_read_off || !d_io) return fail(BBM_E_ARG, "bbm_sitelist_clearzone3_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    int e = bbm_launch_sitelist_cz3(d_lists, d_nss, nreads, cap, (const long long*)d_read_off, cfg, ambiguous_toss, d_io, stream ? (cudaStream_t)stream : c->stream);
    if (e) return fail(BBM_E_CUDA, "sitelist_cz3_kernel launch", (cudaError_t)e);
    c->launches++;
    return BBM_OK;
}
extern "C" int bbm_sitelist_tip_penalty_dev(bbm_ctx* c, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                            const int8_t* d_bases, const int8_t* d_match, const int64_t* d_match_off, const bbm_read_out* d_flags,
                                            int32_t tiplen, int32_t* d_penalty, int32_t* d_status, void* stream) {
    if (!c) return fail(BBM_E_ARG, "bbm_sitelist_tip_penalty_dev: null pointer");
    if (cap < 1 || cap > bbm_sitelist_max_cap() || tiplen < 1 || tiplen > 64) return fail(BBM_E_ARG, "bbm_sitelist_tip_penalty_dev: bad cap or tiplen");
    if (nreads <= 0) return BBM_OK;
    if (!d_lists || !d_nss || !d_read_off || !d_bases || !d_match || !d_match_off || !d_flags || !d_penalty) return fail(BBM_E_ARG, "bbm_sitelist_tip_penalty_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    int e = bbm_launch_sitelist_tip_penalty(d_lists, d_nss, nreads, cap, (const long long*)d_read_off, d_bases, d_match, (const long long*)d_match_off, d_flags,
                                            tiplen, d_penalty, d_status, stream ? (cudaStream_t)stream : c->stream);
    if (e) return fail(BBM_E_CUDA, "sitelist_tip_penalty_kernel launch", (cudaError_t)e);
    c->launches++;
    return BBM_OK;
}

extern "C" int bbm_launch_sam_tasks_from_lists(const bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off, const bbm_read_out* flags,
                                               const long long* match_off, bbm_sam_task* tasks, cudaStream_t st);
extern "C" int bbm_sam_tasks_from_lists_dev(bbm_ctx* c, const bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                            const bbm_read_out* d_flags, const int64_t* d_match_off, bbm_sam_task* d_tasks, void* stream) {
    if (!c) return fail(BBM_E_ARG, "bbm_sam_tasks_from_lists_dev: null pointer");
    if (cap < 1 || cap > bbm_sitelist_max_cap()) return fail(BBM_E_ARG, "bbm_sam_tasks_from_lists_dev: cap must be in 1..64");
    if (nreads <= 0) return BBM_OK;
    if (!d_lists || !d_nss || !d_read_off || !d_flags || !d_tasks) return fail(BBM_E_ARG, "bbm_sam_tasks_from_lists_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    int e = bbm_launch_sam_tasks_from_lists(d_lists, d_nss, nreads, cap, (const long long*)d_read_off, d_flags, (const long long*)d_match_off, d_tasks,
                                            stream ? (cudaStream_t)stream : c->stream);
    if (e) return fail(BBM_E_CUDA, "sam_tasks_from_lists_kernel launch", (cudaError_t)e);
    c->launches++;
    return BBM_OK;
}

// =====================  scoreSlow in rounds (sitelist.cu kernels + the aligner)  =====================
extern "C" int bbm_launch_scoreslow(int phase, int round, bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off,
                                    const int8_t* basesP, const int8_t* basesM, const int8_t* refs, const long long* chrom_off, const int* run,
                                    const bbm_slow_cfg* cfg, int* state, bbm_msa_task* tasks, const bbm_msa_out* outs, bbm_gapped_task* gtasks, int* gaps,
                                    const bbm_msa_out* gouts, int* counters, cudaStream_t st);
static int run_msa_gapped(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_gapped_task* d_gt, const int32_t* d_gaps,
                          bbm_msa_out* d_outs, int64_t ntasks, int8_t* d_match, const int64_t* d_moff, cudaStream_t st, float* ms_out);
extern "C" int bbm_scoreslow_state_ints();
static int scoreslow_locked(bbm_ctx* c, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                            const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_refs, const int64_t* d_chrom_off, const int32_t* d_run,
                            const bbm_slow_cfg* cfg, int32_t* d_status, int32_t max_read_len, cudaStream_t st, int64_t* alignments_out, float* ms_out) {
    const int SI = bbm_scoreslow_state_ints();
    DevBuf &state = c->slowBuf[0], &tasks = c->slowBuf[1], &outs = c->slowBuf[2], &counters = c->slowBuf[3];
    DevBuf &gtasks = c->slowBuf[4], &gaps = c->slowBuf[5], &gouts = c->slowBuf[6];
    if (state.ensure((size_t)nreads * SI * 4) || tasks.ensure((size_t)nreads * sizeof(bbm_msa_task)) || outs.ensure((size_t)nreads * sizeof(bbm_msa_out)) || counters.ensure(16) ||
        gtasks.ensure((size_t)nreads * sizeof(bbm_gapped_task)) || gaps.ensure((size_t)nreads * BBM_MAX_GAPS * 4) || gouts.ensure((size_t)nreads * sizeof(bbm_msa_out)))
        return fail(BBM_E_CUDA, "cudaMalloc scoreSlow scratch");
    int rc = BBM_OK; int64_t aligned = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms_out) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    auto launch = [&](int phase, int k) -> int {
        int e = bbm_launch_scoreslow(phase, k, d_lists, d_nss, nreads, cap, (const long long*)d_read_off, d_basesP, d_basesM, d_refs, (const long long*)d_chrom_off,
                                     d_run, cfg, (int*)state.p, (bbm_msa_task*)tasks.p, (const bbm_msa_out*)outs.p, (bbm_gapped_task*)gtasks.p, (int*)gaps.p,
                                     (const bbm_msa_out*)gouts.p, (int*)counters.p, st);
        if (e) return fail(BBM_E_CUDA, "scoreslow_kernel launch", (cudaError_t)e);
        c->launches++;
        return BBM_OK;
    };
    auto counts = [&](int* h) -> int {
        cudaError_t ce = cudaMemcpyAsync(h, counters.p, 12, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        return ce == cudaSuccess ? BBM_OK : fail(BBM_E_CUDA, "scoreSlow counters", ce);
    };
    const bool trace = getenv("BBM_SLOW_TRACE") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (int k = 0; k < cap && rc == BBM_OK; ++k) {
        int h[3] = {0, 0, 0};
        const double t0 = now();
        if (cudaMemsetAsync(counters.p, 0, 12, st) != cudaSuccess) { rc = fail(BBM_E_CUDA, "memset"); break; }
        if ((rc = launch(0, k)) || (rc = counts(h))) break;
        if (trace) fprintf(stderr, "[scoreSlow] round %d: %d reads active, %d + %d (gapped) alignments requested (prep %.2f ms)\n", k, h[0], h[1], h[2], now() - t0);
        if (h[0] == 0) break;                                   // no read has a k-th site
        if (h[1] > 0) {
            aligned += h[1];
            if ((rc = run_msa(c, d_basesP, d_refs, (const bbm_msa_task*)tasks.p, (bbm_msa_out*)outs.p, h[1], nullptr, nullptr, max_read_len, 0, st, nullptr, nullptr))) break;
        }
        if (h[2] > 0) {
            aligned += h[2];
            if ((rc = run_msa_gapped(c, d_basesP, d_refs, (const bbm_gapped_task*)gtasks.p, (const int32_t*)gaps.p, (bbm_msa_out*)gouts.p, h[2], nullptr, nullptr, st, nullptr))) break;
        }
        if (cudaMemsetAsync(counters.p, 0, 12, st) != cudaSuccess) { rc = fail(BBM_E_CUDA, "memset"); break; }
        if (trace) { cudaStreamSynchronize(st); fprintf(stderr, "[scoreSlow]   first pass done at %.2f ms\n", now() - t0); }
        if ((rc = launch(1, k)) || (rc = counts(h))) break;
        if (trace) fprintf(stderr, "[scoreSlow]   %d + %d (gapped) padding retries\n", h[1], h[2]);
        if (h[1] > 0) {
            aligned += h[1];
            if ((rc = run_msa(c, d_basesP, d_refs, (const bbm_msa_task*)tasks.p, (bbm_msa_out*)outs.p, h[1], nullptr, nullptr, max_read_len, 0, st, nullptr, nullptr))) break;
        }
        if (h[2] > 0) {
            aligned += h[2];
            if ((rc = run_msa_gapped(c, d_basesP, d_refs, (const bbm_gapped_task*)gtasks.p, (const int32_t*)gaps.p, (bbm_msa_out*)gouts.p, h[2], nullptr, nullptr, st, nullptr))) break;
        }
        if ((rc = launch(2, k))) break;
        if (trace) { cudaStreamSynchronize(st); fprintf(stderr, "[scoreSlow]   round done at %.2f ms\n", now() - t0); }
    }
    if (rc == BBM_OK && d_status) {
        cudaError_t ce = cudaMemcpy2DAsync(d_status, 4, (const int*)state.p + 14, (size_t)SI * 4, 4, (size_t)nreads, cudaMemcpyDeviceToDevice, st);
        if (ce != cudaSuccess) rc = fail(BBM_E_CUDA, "scoreSlow status", ce);
    }
    if (ms_out) {
        cudaEventRecord(e1, st); cudaEventSynchronize(e1); float ms = 0.f; cudaEventElapsedTime(&ms, e0, e1); *ms_out = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
    } else cudaStreamSynchronize(st);
    if (alignments_out) *alignments_out = aligned;
    return rc;
}
static int scoreslow_args(const bbm_slow_cfg* cfg, int cap) {
    if (!cfg || cfg->slow_align_padding < 0 || cfg->extra_padding < 0 || cfg->expected_len_limit < 1) return fail(BBM_E_ARG, "bbm_scoreslow: bad cfg");
    if (cap < 1 || cap > bbm_sitelist_max_cap()) return fail(BBM_E_ARG, "bbm_scoreslow: cap must be in 1..64");
    return BBM_OK;
}
extern "C" int bbm_scoreslow_dev(bbm_ctx* c, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                 const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_refs, const int64_t* d_chrom_off, const int32_t* d_run,
                                 const bbm_slow_cfg* cfg, int32_t* d_status, int32_t max_read_len, void* stream, int64_t* alignments_out, float* ms_out) {
    if (!c || !d_lists || !d_nss || !d_read_off || !d_basesP || !d_basesM || !d_refs || !d_chrom_off || !d_run) return fail(BBM_E_ARG, "bbm_scoreslow_dev: null pointer");
    if (int rc = scoreslow_args(cfg, cap)) return rc;
    if (alignments_out) *alignments_out = 0;
    if (nreads <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return scoreslow_locked(c, d_lists, d_nss, nreads, cap, d_read_off, d_basesP, d_basesM, d_refs, d_chrom_off, d_run, cfg, d_status, max_read_len,
                            stream ? (cudaStream_t)stream : c->stream, alignments_out, ms_out);
}
extern "C" int bbm_scoreslow_host(bbm_ctx* c, bbm_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int64_t* read_off,
                                  const int8_t* basesP, const int8_t* basesM, const int8_t* d_refs, const int64_t* chrom_off, int32_t nchroms,
                                  const int32_t* run, const bbm_slow_cfg* cfg, int32_t* status, int64_t* alignments_out) {
    if (!c || !lists || !nss || !read_off || !basesP || !basesM || !d_refs || !chrom_off || !run || nchroms < 1) return fail(BBM_E_ARG, "bbm_scoreslow_host: bad argument");
    if (int rc = scoreslow_args(cfg, cap)) return rc;
    if (alignments_out) *alignments_out = 0;
    if (nreads <= 0) return BBM_OK;
    int maxLen = 1;
    for (int64_t r = 0; r < nreads; ++r) {
        if (nss[r] < 0 || nss[r] > cap || read_off[r + 1] < read_off[r]) return fail(BBM_E_ARG, "bbm_scoreslow_host: list length outside 0..cap");
        for (int i = 0; i < nss[r]; ++i) { const bbm_ss& s = lists[r * cap + i]; if (s.chrom < 1 || s.chrom > nchroms) return fail(BBM_E_ARG, "bbm_scoreslow_host: chromosome out of range"); }
        if (read_off[r + 1] - read_off[r] > maxLen) maxLen = (int)(read_off[r + 1] - read_off[r]);
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t lb = (size_t)nreads * cap * sizeof(bbm_ss), nb = (size_t)nreads * 4, fb = (size_t)(nreads + 1) * 8, rb = (size_t)read_off[nreads], cb = (size_t)(nchroms + 1) * 8;
    DevBuf L_, N_, F_, PM_, C_, R_, S_;
    const size_t half = (rb + 31) & ~(size_t)15;            // both strands in one allocation: the aligner addresses the minus strand as an offset from the plus strand
    if (L_.ensure(lb) || N_.ensure(nb) || F_.ensure(fb) || PM_.ensure(2 * half + 16) || C_.ensure(cb) || R_.ensure(nb) || S_.ensure(nb)) return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(L_.p, lists, lb, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(N_.p, nss, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(F_.p, read_off, fb, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(PM_.p, basesP, rb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync((char*)PM_.p + half, basesM, rb, cudaMemcpyHostToDevice, st)); CK(cudaMemcpyAsync(C_.p, chrom_off, cb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(R_.p, run, nb, cudaMemcpyHostToDevice, st));
    int rc = scoreslow_locked(c, (bbm_ss*)L_.p, (const int32_t*)N_.p, nreads, cap, (const int64_t*)F_.p, (const int8_t*)PM_.p, (const int8_t*)PM_.p + half, d_refs,
                              (const int64_t*)C_.p, (const int32_t*)R_.p, cfg, (int32_t*)S_.p, maxLen, st, alignments_out, nullptr);
    if (rc == BBM_OK) {
        cudaError_t ce = cudaMemcpyAsync(lists, L_.p, lb, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess && status) ce = cudaMemcpyAsync(status, S_.p, nb, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) rc = fail(BBM_E_CUDA, "scoreSlow copy back", ce);
    }
    L_.release(); N_.release(); F_.release(); PM_.release(); C_.release(); R_.release(); S_.release();
    return rc;
}

// =====================  k-mer index build + analysis  =====================
static void index_free(bbm_ctx* c) {
    for (auto& b : c->iblocks) { if (b.starts) cudaFree(b.starts); if (b.sites) cudaFree(b.sites); }
    c->iblocks.clear();
    if (c->d_counts) cudaFree(c->d_counts);
    c->d_counts = nullptr; c->has_index = false;
    if (c->d_icfg) cudaFree(c->d_icfg); if (c->d_iblocks) cudaFree(c->d_iblocks); if (c->d_ihist) cudaFree(c->d_ihist); if (c->d_chrom_off) cudaFree(c->d_chrom_off);
    c->d_icfg = c->d_iblocks = nullptr; c->d_ihist = nullptr; c->d_chrom_off = nullptr;
}

static void index_cfg_init(bbm_index_cfg* c, int k, int chrombits, long long numDefinedBases) {
    // BBIndex statics (current/align2/BBIndex.java:3169-3262) + the small-genome retune of BBMap.loadIndex (BBMap.java:367-382)
    memset(c, 0, sizeof(*c));
    c->keylen = k; c->chrombits = chrombits;
    c->max_hits_reduction2 = 2; c->maximum_max_hits_reduction = 3; c->hit_reduction_div = 5;
    float f = 0.03f;
    if (numDefinedBases < 300000000LL) {
        c->max_hits_reduction2 += 1; c->maximum_max_hits_reduction += 1;
        if (numDefinedBases < 30000000LL) { f = f * 0.5f; c->maximum_max_hits_reduction += 1; c->hit_reduction_div = std::max(c->hit_reduction_div - 1, 3); }
        else if (numDefinedBases < 100000000LL) f = f * 0.6f;
        else f = f * 0.75f;
    }
    c->fraction_to_exclude = f;
    c->min_index_to_drop_long_hit_list = (int)(1000 * (1 - 3.5 * f));      // setFractionToExclude: double arithmetic
    c->max_average_list_to_search = (int)(1000 * (1 - 2.3 * f));
    c->max_average_list_to_search2 = (int)(1000 * (1 - 1.4 * f));
    c->max_single_list_to_search = (int)(1000 * (1 - 1.0 * f));
    c->max_shortest_list_to_search = (int)(1000 * (1 - 2.8 * f));
    c->shift_length = 32 - 1 - chrombits;
    c->chroms_per_block = 1 << chrombits;
}

extern "C" int bbm_index_build(bbm_ctx* c, const int8_t* d_chroms, const int64_t* chrom_off, int32_t nchroms, int32_t keylen, int32_t chrombits,
                               bbm_index_cfg* cfg_out, int32_t* nblocks_out) {
    if (!c || !d_chroms || !chrom_off || nchroms < 1 || keylen < 8 || keylen > 15) return fail(BBM_E_ARG, "bbm_index_build: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    index_free(c);
    const auto t_build0 = std::chrono::steady_clock::now();
    const int k = keylen;
    const long long keyspace = 1LL << (2 * k);
    long long maxLen = 0, total = chrom_off[nchroms] - chrom_off[0];
    for (int i = 0; i < nchroms; ++i) maxLen = std::max<long long>(maxLen, chrom_off[i + 1] - chrom_off[i]);
    if (chrombits < 0) { int nlz = maxLen == 0 ? 32 : __builtin_clz((unsigned)maxLen); chrombits = std::min(nlz - 1, 16); }
    if (maxLen - 1 > (long long)(~((-1) << (32 - 1 - chrombits)))) return fail(BBM_E_ARG, "bbm_index_build: chromosome longer than MAX_ALLOWED_CHROM_INDEX for these chrombits");
    // numDefinedBases
    unsigned long long* d_def = nullptr; CK(cudaMalloc(&d_def, 8)); CK(cudaMemsetAsync(d_def, 0, 8, st));
    int e = bbm_index_count_defined(d_chroms + chrom_off[0], total, d_def, st);
    if (e) return fail(BBM_E_CUDA, "count_defined", (cudaError_t)e);
    unsigned long long nDefined = 0; CK(cudaMemcpyAsync(&nDefined, d_def, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); cudaFree(d_def);
    c->launches++;
    index_cfg_init(&c->icfg, k, chrombits, (long long)nDefined);
    const int cpb = c->icfg.chroms_per_block, low = cpb - 1, shift = c->icfg.shift_length;
    // blocks: chromosomes sharing (chrom & ~low); chrom numbers are 1-based (IndexMaker4.makeIndex :44-62)
    for (int chrom = 1; chrom <= nchroms;) {
        const int a = std::max(1, chrom & ~low), b = std::min(nchroms, (chrom & ~low) + cpb - 1);
        bbm_ctx::IndexBlock B; B.minChrom = a; B.maxChrom = b;
        long long n = 0;
        for (int ch = a; ch <= b; ++ch) n += chrom_off[ch] - chrom_off[ch - 1];
        unsigned *k0 = nullptr, *k1 = nullptr; int *v0 = nullptr, *v1 = nullptr, *sizes = nullptr;
        CK(cudaMalloc(&k0, (size_t)n * 4 + 16)); CK(cudaMalloc(&k1, (size_t)n * 4 + 16)); CK(cudaMalloc(&v0, (size_t)n * 4 + 16)); CK(cudaMalloc(&v1, (size_t)n * 4 + 16));
        CK(cudaMalloc(&sizes, (size_t)(keyspace + 1) * 4)); CK(cudaMemsetAsync(sizes, 0, (size_t)(keyspace + 1) * 4, st));
        CK(cudaMalloc(&B.starts, (size_t)(keyspace + 1) * 4));
        const unsigned invalid = 1u << (2 * k);
        long long base = 0;
        for (int ch = a; ch <= b; ++ch) {
            const int len = (int)(chrom_off[ch] - chrom_off[ch - 1]);
            e = bbm_index_emit(d_chroms + chrom_off[ch - 1], len, k, (ch & low) << shift, k0, v0, base, sizes, invalid, st);
            if (e) return fail(BBM_E_CUDA, "index_emit_kernel", (cudaError_t)e);
            c->launches++;
            base += len;
        }
        size_t tb1 = 0, tb2 = 0;
        bbm_index_sort_pairs(nullptr, &tb1, k0, k1, v0, v1, n, 2 * k + 1, st);
        bbm_index_scan(nullptr, &tb2, sizes, B.starts, keyspace + 1, st);
        void* temp = nullptr; CK(cudaMalloc(&temp, std::max(tb1, tb2) + 16));
        e = bbm_index_sort_pairs(temp, &tb1, k0, k1, v0, v1, n, 2 * k + 1, st);
        if (e) return fail(BBM_E_CUDA, "radix sort", (cudaError_t)e);
        e = bbm_index_scan(temp, &tb2, sizes, B.starts, keyspace + 1, st);
        if (e) return fail(BBM_E_CUDA, "scan", (cudaError_t)e);
        c->launches += 2;
        int nsites = 0;
        CK(cudaMemcpyAsync(&nsites, B.starts + keyspace, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        B.nsites = nsites;
        CK(cudaMalloc(&B.sites, (size_t)std::max(nsites, 1) * 4));
        CK(cudaMemcpyAsync(B.sites, v1, (size_t)nsites * 4, cudaMemcpyDeviceToDevice, st));      // valid pairs sort before the invalid key
        CK(cudaStreamSynchronize(st));
        cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(sizes); cudaFree(temp);
        c->iblocks.push_back(B);
        chrom = b + 1;
    }
    // analyzeIndex
    unsigned long long* d_clump = nullptr; int* d_max = nullptr;
    CK(cudaMalloc(&c->d_counts, (size_t)keyspace * 4)); CK(cudaMemsetAsync(c->d_counts, 0, (size_t)keyspace * 4, st));
    CK(cudaMalloc(&d_clump, (size_t)keyspace * 8)); CK(cudaMemsetAsync(d_clump, 0, (size_t)keyspace * 8, st));
    CK(cudaMalloc(&d_max, 4)); CK(cudaMemsetAsync(d_max, 0, 4, st));
    for (auto& B : c->iblocks) { e = bbm_index_analyze_block(B.starts, B.sites, k, c->d_counts, d_clump, st); if (e) return fail(BBM_E_CUDA, "analyze_block", (cudaError_t)e); c->launches++; }
    e = bbm_index_finish_counts(k, c->d_counts, d_clump, d_max, st);
    if (e) return fail(BBM_E_CUDA, "finish_counts", (cudaError_t)e);
    c->launches += 3;
    int maxv = 0; CK(cudaMemcpyAsync(&maxv, d_max, 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    int* d_len = nullptr; CK(cudaMalloc(&d_len, (size_t)(maxv + 1) * 4)); CK(cudaMemsetAsync(d_len, 0, (size_t)(maxv + 1) * 4, st));
    e = bbm_index_lenhist(k, c->d_counts, d_len, st);
    if (e) return fail(BBM_E_CUDA, "lenhist", (cudaError_t)e);
    c->launches++;
    std::vector<int> lenCounts((size_t)maxv + 1);
    CK(cudaMemcpyAsync(lenCounts.data(), d_len, (size_t)(maxv + 1) * 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    cudaFree(d_clump); cudaFree(d_max); cudaFree(d_len);
    {   // Tools.makeLengthHistogram4 (Tools.java:1817-1850), buckets=1000, on the (small) histogram of list lengths
        long long tot = 0;
        for (int i = 1; i <= maxv; ++i) tot += (long long)i * lenCounts[i];
        long long sum = 0; int ptr = 0; const int buckets = 1000;
        for (int i = 0; i < buckets; ++i) {
            const long long nextLimit = ((tot * i) + buckets / 2) / buckets;
            while (ptr < maxv + 1 && sum < nextLimit) { sum += (int)(lenCounts[ptr] * ptr); ptr++; }
            c->ihist[i] = std::max(0, ptr - 1);
        }
        c->ihist[buckets] = maxv;
        const float f = c->icfg.fraction_to_exclude;
        c->icfg.max_usable_length = std::max(2 * 20, c->ihist[(int)((1 - f) * (1001 - 1))]);
        c->icfg.max_usable_length2 = std::max(6 * 20, c->ihist[(int)((1 - f * 0.25f) * (1001 - 1))]);
        int pps = (int)floor((double)((-50 * 4000.f) / std::max(2 * 20, c->ihist[c->icfg.max_average_list_to_search])));
        if (pps == 0) pps = -1;
        c->icfg.points_per_site = pps;
    }
    c->d_chroms = d_chroms; c->chrom_off.assign(chrom_off, chrom_off + nchroms + 1);
    {   // device-side descriptors for the search kernel
        struct Blk { const int* starts; const int* sites; };
        std::vector<Blk> hb;
        for (auto& B : c->iblocks) hb.push_back(Blk{B.starts, B.sites});
        CK(cudaMalloc(&c->d_icfg, sizeof(bbm_index_cfg))); CK(cudaMemcpy(c->d_icfg, &c->icfg, sizeof(bbm_index_cfg), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&c->d_iblocks, hb.size() * sizeof(Blk))); CK(cudaMemcpy(c->d_iblocks, hb.data(), hb.size() * sizeof(Blk), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&c->d_ihist, sizeof(c->ihist))); CK(cudaMemcpy(c->d_ihist, c->ihist, sizeof(c->ihist), cudaMemcpyHostToDevice));
        std::vector<long long> rel(nchroms + 1);
        for (int i = 0; i <= nchroms; ++i) rel[i] = chrom_off[i];
        CK(cudaMalloc(&c->d_chrom_off, rel.size() * 8)); CK(cudaMemcpy(c->d_chrom_off, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice));
    }
    CK(cudaStreamSynchronize(st));
    c->index_build_us = (long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_build0).count();
    c->has_index = true;
    if (cfg_out) *cfg_out = c->icfg;
    if (nblocks_out) *nblocks_out = (int)c->iblocks.size();
    return BBM_OK;
}

extern "C" int bbm_index_block_sites(bbm_ctx* c, int32_t block, int64_t* nsites_out) {
    if (!c || !c->has_index || block < 0 || block >= (int)c->iblocks.size() || !nsites_out) return fail(BBM_E_ARG, "bbm_index_block_sites: bad argument");
    *nsites_out = c->iblocks[block].nsites;
    return BBM_OK;
}

extern "C" int bbm_index_download(bbm_ctx* c, int32_t block, int32_t* starts, int32_t* sites, int32_t* counts, int32_t* hist1001) {
    if (!c || !c->has_index || block < 0 || block >= (int)c->iblocks.size()) return fail(BBM_E_ARG, "bbm_index_download: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const long long keyspace = 1LL << (2 * c->icfg.keylen);
    const auto& B = c->iblocks[block];
    if (starts) CK(cudaMemcpy(starts, B.starts, (size_t)(keyspace + 1) * 4, cudaMemcpyDeviceToHost));
    if (sites && B.nsites) CK(cudaMemcpy(sites, B.sites, (size_t)B.nsites * 4, cudaMemcpyDeviceToHost));
    if (counts) CK(cudaMemcpy(counts, c->d_counts, (size_t)keyspace * 4, cudaMemcpyDeviceToHost));
    if (hist1001) memcpy(hist1001, c->ihist, sizeof(c->ihist));
    return BBM_OK;
}

// =====================  read ingest (Read.validate + reverse complement, a0)  =====================
static int run_ingest(bbm_ctx* c, int8_t* db, int8_t* dq, const int64_t* doff, int64_t nreads, int max_len, int flags, int8_t* dm, int* df,
                      cudaStream_t st, float* ms_out) {
    if (nreads <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    if (max_len < 1) return fail(BBM_E_ARG, "bbm_ingest: max_len < 1");
    int rpb = 64;                                             // reads per block; three staged arrays must fit 192 KB of shared memory
    while (rpb > 1 && (long long)rpb * max_len + 48 > 64 * 1024) rpb >>= 1;
    if ((long long)rpb * max_len + 48 > 64 * 1024) return fail(BBM_E_SHAPE, "bbm_ingest: read longer than 65488 bases");
    const int stage = (int)((((long long)rpb * max_len + 32) + 15) & ~15LL);
    long long blocks = (nreads + rpb - 1) / rpb;
    const long long cap = (long long)c->sms * 4;
    if (blocks > cap) blocks = cap;
    CK(cudaEventRecord(c->ev0, st));
    int e = bbm_launch_ingest(db, dq, (const long long*)doff, nreads, dm, df, flags, rpb, stage, (int)blocks, st);
    if (e) return fail(BBM_E_CUDA, "ingest_kernel launch", (cudaError_t)e);
    c->launches++;
    CK(cudaEventRecord(c->ev1, st));
    CK(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    return BBM_OK;
}

extern "C" int bbm_ingest_batch_dev(bbm_ctx* c, int8_t* d_bases, int8_t* d_quality, const int64_t* d_read_off, int64_t nreads, int32_t max_len,
                                    int32_t flags, int8_t* d_basesM, int32_t* d_read_flags, void* stream, float* kernel_ms_out) {
    if (!c || !d_bases || !d_read_off) return fail(BBM_E_ARG, "bbm_ingest_batch_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_ingest(c, d_bases, d_quality, d_read_off, nreads, max_len, flags, d_basesM, d_read_flags, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}

extern "C" int bbm_ingest_batch_host(bbm_ctx* c, int8_t* bases, int8_t* quality, const int64_t* read_off, int64_t nreads, int32_t flags,
                                     int8_t* basesM, int32_t* read_flags) {
    if (!c || !bases || !read_off) return fail(BBM_E_ARG, "bbm_ingest_batch_host: null pointer");
    if (nreads <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t nb = (size_t)(read_off[nreads] - read_off[0]);
    int max_len = 1;
    for (int64_t i = 0; i < nreads; ++i) { const int64_t l = read_off[i + 1] - read_off[i]; if (l < 0) return fail(BBM_E_ARG, "read_off not ascending"); if (l > max_len) max_len = (int)l; }
    if (read_off[0] != 0) return fail(BBM_E_ARG, "bbm_ingest_batch_host: read_off[0] must be 0");
    DevBuf* B = c->d_ing;   // 0 bases, 1 quality, 2 offsets, 3 basesM, 4 flags
    if (B[0].ensure(nb + 32) || B[1].ensure(nb + 32) || B[2].ensure((size_t)(nreads + 1) * 8) || B[3].ensure(nb + 32) || B[4].ensure((size_t)nreads * 4))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(B[0].p, bases, nb, cudaMemcpyHostToDevice, st));
    if (quality) CK(cudaMemcpyAsync(B[1].p, quality, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[2].p, read_off, (size_t)(nreads + 1) * 8, cudaMemcpyHostToDevice, st));
    int rc = run_ingest(c, (int8_t*)B[0].p, quality ? (int8_t*)B[1].p : nullptr, (const int64_t*)B[2].p, nreads, max_len, flags,
                        basesM ? (int8_t*)B[3].p : nullptr, read_flags ? (int*)B[4].p : nullptr, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(bases, B[0].p, nb, cudaMemcpyDeviceToHost, st));
    if (quality) CK(cudaMemcpyAsync(quality, B[1].p, nb, cudaMemcpyDeviceToHost, st));
    if (basesM) CK(cudaMemcpyAsync(basesM, B[3].p, nb, cudaMemcpyDeviceToHost, st));
    if (read_flags) CK(cudaMemcpyAsync(read_flags, B[4].p, (size_t)nreads * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}

// =====================  SAM record fields (SamLine)  =====================
static int run_sam(bbm_ctx* c, const bbm_sam_task* dt, int64_t n, const int8_t* dm, const int* dso, const int* dsl, const int* dsn, int nchroms,
                   const bbm_sam_cfg* cfg, bbm_sam_out* dout, int8_t* dcb, const int64_t* dco, cudaStream_t st, float* ms_out) {
    if (n <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    if (!c->sam_table) {
        // (float)Tools.log2(length) of SamLine.toMapq, computed with the host libm (the same call the oracle makes)
        std::vector<float> tab(bbm_sam_log2_tab());
        for (size_t i = 0; i < tab.size(); ++i) tab[i] = i == 0 ? 0.f : (float)(log((double)i) * (1 / log(2.0)));
        int e0 = bbm_sam_upload_table(tab.data());
        if (e0) return fail(BBM_E_CUDA, "sam log2 table", (cudaError_t)e0);
        c->sam_table = true;
    }
    CK(cudaEventRecord(c->ev0, st));
    int e = bbm_launch_sam(dt, n, dm, dso, dsl, dsn, nchroms, cfg, dout, dcb, (const long long*)dco, st);
    if (e) return fail(BBM_E_CUDA, "sam_kernel launch", (cudaError_t)e);
    c->launches++;
    CK(cudaEventRecord(c->ev1, st));
    CK(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    return BBM_OK;
}

extern "C" int bbm_sam_batch_dev(bbm_ctx* c, const bbm_sam_task* d_tasks, int64_t n, const int8_t* d_match_buf, const int32_t* d_scaf_off,
                                 const int32_t* d_scaf_loc, const int32_t* d_scaf_len, int32_t nchroms, const bbm_sam_cfg* cfg, bbm_sam_out* d_outs,
                                 int8_t* d_cigar_buf, const int64_t* d_cigar_off, void* stream, float* kernel_ms_out) {
    if (!c || !d_tasks || !d_match_buf || !d_scaf_off || !d_scaf_loc || !d_scaf_len || !cfg || !d_outs || !d_cigar_buf || !d_cigar_off || nchroms < 1)
        return fail(BBM_E_ARG, "bbm_sam_batch_dev: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_sam(c, d_tasks, n, d_match_buf, d_scaf_off, d_scaf_loc, d_scaf_len, nchroms, cfg, d_outs, d_cigar_buf, d_cigar_off,
                   stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}

extern "C" int bbm_sam_batch_host(bbm_ctx* c, const bbm_sam_task* tasks, int64_t n, const int8_t* match_buf, int64_t match_bytes, const int32_t* scaf_off,
                                  const int32_t* scaf_loc, const int32_t* scaf_len, int32_t nchroms, const bbm_sam_cfg* cfg, bbm_sam_out* outs,
                                  int8_t* cigar_buf, const int64_t* cigar_off) {
    if (!c || !tasks || !match_buf || !scaf_off || !scaf_loc || !scaf_len || !cfg || !outs || !cigar_buf || !cigar_off || nchroms < 1 || match_bytes < 0)
        return fail(BBM_E_ARG, "bbm_sam_batch_host: bad argument");
    if (n <= 0) return BBM_OK;
    for (int64_t i = 0; i < n; ++i) {
        const bbm_sam_task& t = tasks[i];
        if (t.mate >= n || t.match_len < 0 || t.match_off < 0 || t.match_off + t.match_len > match_bytes) return fail(BBM_E_ARG, "bbm_sam_batch_host: record outside the buffers");
        if ((t.flags & BBM_RF_MAPPED) && (t.chrom < 1 || t.chrom > nchroms)) return fail(BBM_E_ARG, "bbm_sam_batch_host: chromosome out of range");
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int nscaf = scaf_off[nchroms];
    const size_t cb = (size_t)cigar_off[n];
    DevBuf* B = c->d_sam;   // 0 tasks, 1 match, 2 scaf_off, 3 scaf_loc, 4 scaf_len, 5 outs, 6 cigar, 7 cigar_off
    if (B[0].ensure((size_t)n * sizeof(bbm_sam_task)) || B[1].ensure((size_t)match_bytes + 16) || B[2].ensure((size_t)(nchroms + 1) * 4) ||
        B[3].ensure((size_t)nscaf * 4 + 16) || B[4].ensure((size_t)nscaf * 4 + 16) || B[5].ensure((size_t)n * sizeof(bbm_sam_out)) || B[6].ensure(cb + 16) ||
        B[7].ensure((size_t)(n + 1) * 8))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(B[0].p, tasks, (size_t)n * sizeof(bbm_sam_task), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[1].p, match_buf, (size_t)match_bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[2].p, scaf_off, (size_t)(nchroms + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[3].p, scaf_loc, (size_t)nscaf * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[4].p, scaf_len, (size_t)nscaf * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[7].p, cigar_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(B[6].p, 0, cb, st));
    int rc = run_sam(c, (const bbm_sam_task*)B[0].p, n, (const int8_t*)B[1].p, (const int*)B[2].p, (const int*)B[3].p, (const int*)B[4].p, nchroms, cfg,
                     (bbm_sam_out*)B[5].p, (int8_t*)B[6].p, (const int64_t*)B[7].p, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(outs, B[5].p, (size_t)n * sizeof(bbm_sam_out), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(cigar_buf, B[6].p, cb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}

// =====================  index search (BBIndex.find)  =====================
static int run_search(bbm_ctx* c, const int8_t* db, const int8_t* dbs, const int64_t* doff, int64_t nreads, const int* dn, const int* dof,
                      const int* dks, int maxKeys, int quit2, bbm_search_head* dh, bbm_site* ds, int maxSites, int maxReadLen, cudaStream_t st, float* ms_out) {
    if (!c->has_index) return fail(BBM_E_ARG, "bbm_search: no index in this context (call bbm_index_build first)");
    if (nreads <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    if ((int)c->iblocks.size() > 64) return fail(BBM_E_SHAPE, "bbm_search: more than 64 index blocks");
    const int T = bbm_search_threads();
    int blocks = c->sms * 8;
    const long long need = (nreads + T - 1) / T;
    if (need < blocks) blocks = (int)need;
    (void)maxReadLen;     // reserved: lets a later kernel size its per-read working set to the batch
    if (c->searchCtx.ensure((size_t)c->sms * 8 * T * bbm_search_pool_bytes())) return fail(BBM_E_CUDA, "cudaMalloc search scratch");
    unsigned int* cb = (unsigned int*)c->counters.p;
    if (c->search_prof) CK(cudaMemsetAsync(cb + 208, 0, 40, st));
    unsigned long long* prof = c->search_prof ? (unsigned long long*)(cb + 208) : nullptr;
    const int nblk = (int)c->iblocks.size(), nchr = (int)c->chrom_off.size() - 1;
    CK(cudaEventRecord(c->ev0, st));
    if (c->search_split) {
        // one phase of BBIndex.find per launch (key filtering, prescan, walk): all lanes of a warp run the same phase
        const int stride = bbm_search_mid_stride(maxKeys, nblk);
        if (c->searchRev.ensure((size_t)nreads * stride * 4)) return fail(BBM_E_CUDA, "cudaMalloc search phase state");
        for (int ph = 1; ph <= 4; ph <<= 1) {
            if (ph == 2 && c->search_split >= 2) {
                // prescan with one warp per read; reads with more than 32 keys are left to the thread-per-read launch that follows
                CK(cudaMemsetAsync(cb + 202, 0, 4, st));
                int wblocks = c->sms * 8; const long long wneed = (nreads + 3) / 4; if (wneed < wblocks) wblocks = (int)wneed;
                int e = bbm_launch_search_prescan_warp((const bbm_index_cfg*)c->d_icfg, c->d_iblocks, nblk, nchr, c->d_counts, (const long long*)doff, nreads, dn,
                                                       maxKeys, dh, cb + 202, wblocks, (int*)c->searchRev.p, stride, st);
                if (e) return fail(BBM_E_CUDA, "prescan_warp_kernel launch", (cudaError_t)e);
                c->launches++;
                if (maxKeys <= 32) continue;
            }
            CK(cudaMemsetAsync(cb + 202, 0, 4, st));
            int e = bbm_launch_search((const bbm_index_cfg*)c->d_icfg, c->d_iblocks, nblk, nchr, c->d_counts, c->d_ihist, c->d_chroms, c->d_chrom_off, db, dbs,
                                      (const long long*)doff, nreads, dn, dof, dks, maxKeys, quit2, dh, ds, maxSites, c->searchCtx.p, cb + 202, prof, blocks,
                                      c->search_shared ? 0 : 1, ph, (int*)c->searchRev.p, stride, st);
            if (e) return fail(BBM_E_CUDA, "search_kernel launch", (cudaError_t)e);
            c->launches++;
        }
    } else {
        CK(cudaMemsetAsync(cb + 202, 0, 4, st));
        int e = bbm_launch_search((const bbm_index_cfg*)c->d_icfg, c->d_iblocks, nblk, nchr, c->d_counts, c->d_ihist, c->d_chroms, c->d_chrom_off, db, dbs,
                                  (const long long*)doff, nreads, dn, dof, dks, maxKeys, quit2, dh, ds, maxSites, c->searchCtx.p, cb + 202, prof, blocks,
                                  c->search_shared ? 0 : 1, 7, nullptr, 0, st);
        if (e) return fail(BBM_E_CUDA, "search_kernel launch", (cudaError_t)e);
        c->launches++;
    }
    CK(cudaEventRecord(c->ev1, st));
    CK(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    if (c->search_prof) CK(cudaMemcpy(c->search_cycles, cb + 208, 40, cudaMemcpyDeviceToHost));
    return BBM_OK;
}

extern "C" int bbm_search_batch_dev(bbm_ctx* c, const int8_t* d_bases, const int8_t* d_baseScores, const int64_t* d_read_off, int64_t nreads,
                                    const int32_t* d_nkeys, const int32_t* d_offsets, const int32_t* d_keyScores, int32_t maxKeys,
                                    int32_t quit2, bbm_search_head* d_heads, bbm_site* d_sites, int32_t max_sites, int32_t max_read_len, void* stream, float* kernel_ms_out) {
    if (!c || !d_bases || !d_baseScores || !d_read_off || !d_nkeys || !d_offsets || !d_keyScores || !d_heads || !d_sites || max_sites < 1 || maxKeys < 1 || maxKeys > 96)
        return fail(BBM_E_ARG, "bbm_search_batch_dev: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_search(c, d_bases, d_baseScores, d_read_off, nreads, d_nkeys, d_offsets, d_keyScores, maxKeys, quit2, d_heads, d_sites, max_sites,
                      max_read_len, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}

extern "C" int bbm_search_batch_host(bbm_ctx* c, const int8_t* bases, const int8_t* baseScores, const int64_t* read_off, int64_t nreads,
                                     const int32_t* nkeys, const int32_t* offsets, const int32_t* keyScores, int32_t maxKeys,
                                     int32_t quit2, bbm_search_head* heads, bbm_site* sites, int32_t max_sites) {
    if (!c || !bases || !baseScores || !read_off || !nkeys || !offsets || !keyScores || !heads || !sites || max_sites < 1 || maxKeys < 1 || maxKeys > 96)
        return fail(BBM_E_ARG, "bbm_search_batch_host: bad argument");
    if (nreads <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t nb = (size_t)read_off[nreads], kb = (size_t)nreads * maxKeys * 4;
    int max_len = 1;
    for (int64_t i = 0; i < nreads; ++i) { const int64_t l = read_off[i + 1] - read_off[i]; if (l > max_len) max_len = (int)(l > 100000 ? 100000 : l); }
    const size_t hb = (size_t)nreads * sizeof(bbm_search_head), sb = (size_t)nreads * max_sites * sizeof(bbm_site);
    DevBuf* B = c->d_srch;   // 0 bases, 1 baseScores, 2 off, 3 nkeys, 4 offsets, 5 keyScores, 6 heads, 7 sites
    if (B[0].ensure(nb + 32) || B[1].ensure(nb + 32) || B[2].ensure((size_t)(nreads + 1) * 8) || B[3].ensure((size_t)nreads * 4) || B[4].ensure(kb) ||
        B[5].ensure(kb) || B[6].ensure(hb) || B[7].ensure(sb))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(B[0].p, bases, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[1].p, baseScores, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[2].p, read_off, (size_t)(nreads + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[3].p, nkeys, (size_t)nreads * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[4].p, offsets, kb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[5].p, keyScores, kb, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(B[7].p, 0, sb, st));
    int rc = run_search(c, (const int8_t*)B[0].p, (const int8_t*)B[1].p, (const int64_t*)B[2].p, nreads, (const int*)B[3].p, (const int*)B[4].p,
                        (const int*)B[5].p, maxKeys, quit2, (bbm_search_head*)B[6].p, (bbm_site*)B[7].p, max_sites, max_len, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(heads, B[6].p, hb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sites, B[7].p, sb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}
