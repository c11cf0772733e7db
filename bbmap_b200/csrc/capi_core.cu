// capi_core.cu — context lifecycle, options, statistics, the integer-peak probe
// Part of the C ABI of libbbmapcuda.so (include/bbmap_cuda.h): host-side glue only (device buffers, streams, launches).
// No CPU implementation of any compute path lives here: without a device every call fails loudly.
#include "ctx.h"

thread_local std::string bbm_g_err;
int fail(int code, const char* what, cudaError_t e) {
    bbm_g_err = what;
    if (e != cudaSuccess) { bbm_g_err += ": "; bbm_g_err += cudaGetErrorString(e); }
    return code;
}

extern "C" const char* bbm_last_error(void) { return bbm_g_err.c_str(); }

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
    if (c->counters.ensure(512 * 4)) return fail(BBM_E_CUDA, "cudaMalloc counters");
    CK(cudaMemset(c->counters.p, 0, 512 * 4));          // run_msa re-zeroes the first 192 words per batch; the debug counters behind them start at 0 too
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
    c->d_reads.release(); c->d_tasks.release(); c->d_outs.release(); c->d_match.release(); c->d_moff.release(); c->d_dump.release(); c->d_refs2.release(); c->seedScratch.release(); for (auto& b : c->d_seed) b.release(); c->stripScratch.release(); c->bandScratch.release(); for (auto& b : c->slowBuf) b.release(); c->searchCtx.release(); c->searchRev.release(); for (auto& b : c->d_ing) b.release(); for (auto& b : c->d_sam) b.release(); c->grefPool.release(); c->grefInfo.release(); c->grefTasks.release(); c->d_gtasks.release(); c->d_gaps.release(); for (auto& b : c->d_srch) b.release();
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
    cudaError_t ce = cudaMemcpy(p, host, (size_t)nbytes, cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = cudaMemset((char*)p + nbytes, 'N', 256);
    if (ce != cudaSuccess) { cudaFree(p); return fail(BBM_E_CUDA, "bbm_upload: copy to device", ce); }
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

extern "C" int bbm_set_option(bbm_ctx* c, const char* key, int value) {
    if (!c || !key) return fail(BBM_E_ARG, "bbm_set_option: null");
    if (!strcmp(key, "narrow")) { c->use_narrow = value; return BBM_OK; }
    if (!strcmp(key, "strip_min_tasks")) { c->strip_min_tasks = value; return BBM_OK; }
    if (!strcmp(key, "slow_lookahead")) { c->slow_lookahead = value; return BBM_OK; }
    if (!strcmp(key, "band")) { c->use_band = value; return BBM_OK; }
    if (!strcmp(key, "banded_thread")) { c->banded_thread = value; return BBM_OK; }
    if (!strcmp(key, "strip")) { c->use_strip = value; return BBM_OK; }
    if (!strcmp(key, "msa_count")) { c->msa_count = value; return BBM_OK; }
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
    if (!strcmp(key, "msa_us")) return (int64_t)(c->msa_ms * 1000.0);          // device time of all aligner batches so far (CUDA events)
    if (!strcmp(key, "msa_cells")) return c->msa_cells;                         // reference cell visits summed while "msa_count" is on
    if (!strcmp(key, "strip_units")) return (int64_t)c->strip_units;            // rows of 8 cells evaluated (with strip_debug bit 2)
    if (!strcmp(key, "strip_lane_iters")) return (int64_t)c->strip_lane_iters;  // lane-iterations of the evaluation phase: units/iters = lane utilisation
    if (!strncmp(key, "search_cycles_", 14) && key[14] >= '0' && key[14] <= '4') return (int64_t)c->search_cycles[key[14] - '0'];   // thread-cycles: total, filter, prescan, walk, extend
    if (!strcmp(key, "index_build_us")) return c->index_build_us;        // host wall time of the last bbm_index_build (reference already resident)
    return -1;
}

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

