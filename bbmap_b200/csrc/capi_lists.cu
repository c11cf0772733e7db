// capi_lists.cu — per-read site-list policies and scoreSlow in rounds
// Part of the C ABI of libbbmapcuda.so (include/bbmap_cuda.h): host-side glue only (device buffers, streams, launches).
// No CPU implementation of any compute path lives here: without a device every call fails loudly.
#include "ctx.h"

// =====================  per-read site-list policies (sitelist.cu)  =====================
extern "C" int bbm_sitelist_max_cap();
extern "C" int bbm_launch_sitelist(int op, bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const int8_t* basesP,
                                   const int8_t* basesM, const int8_t* refs, const long long* chrom_off, const bbm_policy_cfg* cfg, bbm_read_out* out,
                                   cudaStream_t st);
extern "C" int bbm_launch_sitelist_from_search(const bbm_search_head* heads, const bbm_site* sites, long long nreads, int maxSites, bbm_ss* lists,
                                               int* nss, int cap, cudaStream_t st);
static int sitelist_args(int op, int cap, const bbm_policy_cfg* cfg) {
    if (op != BBM_SL_TRIM && op != BBM_SL_NOINDEL && op != BBM_SL_FINAL && op != BBM_SL_MERGE) return fail(BBM_E_ARG, "bbm_sitelist: unknown op");
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
                                    const bbm_msa_out* gouts, int* counters, int window, int* slots, bbm_ss* backup, cudaStream_t st);
extern "C" int bbm_scoreslow_state_ints();
int scoreslow_locked(bbm_ctx* c, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                            const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_refs, const int64_t* d_chrom_off, const int32_t* d_run,
                            const bbm_slow_cfg* cfg, int32_t* d_status, int32_t max_read_len, cudaStream_t st, int64_t* alignments_out, float* ms_out) {
    const int SI = bbm_scoreslow_state_ints();
    DevBuf &state = c->slowBuf[0], &tasks = c->slowBuf[1], &outs = c->slowBuf[2], &counters = c->slowBuf[3];
    DevBuf &gtasks = c->slowBuf[4], &gaps = c->slowBuf[5], &gouts = c->slowBuf[6], &slots = c->slowBuf[7], &backup = c->slowBuf[8];
    // `nreads` sites can be in flight in one round (one per read in the first round, a window of several for the few reads still active later)
    if (state.ensure((size_t)nreads * SI * 4) || tasks.ensure((size_t)nreads * sizeof(bbm_msa_task)) || outs.ensure((size_t)nreads * sizeof(bbm_msa_out)) || counters.ensure(32) ||
        slots.ensure((size_t)nreads * SI * 4) || backup.ensure((size_t)nreads * sizeof(bbm_ss)) ||
        gtasks.ensure((size_t)nreads * sizeof(bbm_gapped_task)) || gaps.ensure((size_t)nreads * BBM_MAX_GAPS * 4) || gouts.ensure((size_t)nreads * sizeof(bbm_msa_out)))
        return fail(BBM_E_CUDA, "cudaMalloc scoreSlow scratch");
    int rc = BBM_OK; int64_t aligned = 0;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ms_out) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, st); }
    int window = 1;
    auto launch = [&](int phase, int k) -> int {
        int e = bbm_launch_scoreslow(phase, k, d_lists, d_nss, nreads, cap, (const long long*)d_read_off, d_basesP, d_basesM, d_refs, (const long long*)d_chrom_off,
                                     d_run, cfg, (int*)state.p, (bbm_msa_task*)tasks.p, (const bbm_msa_out*)outs.p, (bbm_gapped_task*)gtasks.p, (int*)gaps.p,
                                     (const bbm_msa_out*)gouts.p, (int*)counters.p, window, (int*)slots.p, (bbm_ss*)backup.p, st);
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
    if (cudaMemsetAsync(counters.p, 0, 32, st) != cudaSuccess) return fail(BBM_E_CUDA, "memset");
    const int lookahead = c->slow_lookahead > 0 ? (c->slow_lookahead > 64 ? 64 : c->slow_lookahead) : 1;
    long long activePrev = nreads;
    for (int k = 0; k <= cap && rc == BBM_OK; ++k) {
        int h[3] = {0, 0, 0};
        const double t0 = now();
        // sites in flight <= reads still active x window <= nreads (the size of the slot pool and of the request lists)
        window = (int)std::max<long long>(1, std::min<long long>(lookahead, nreads / std::max<long long>(1, activePrev)));
        if (cudaMemsetAsync(counters.p, 0, 16, st) != cudaSuccess) { rc = fail(BBM_E_CUDA, "memset"); break; }
        if ((rc = launch(0, k)) || (rc = counts(h))) break;
        if (trace) fprintf(stderr, "[scoreSlow] round %d (window %d): %d reads active, %d + %d (gapped) alignments requested (prep %.2f ms)\n", k, window, h[0], h[1], h[2], now() - t0);
        if (h[0] == 0) break;                                   // no read has a site left
        activePrev = h[0];
        if (h[1] > 0) {
            if ((rc = run_msa(c, d_basesP, d_refs, (const bbm_msa_task*)tasks.p, (bbm_msa_out*)outs.p, h[1], nullptr, nullptr, max_read_len, 0, st, nullptr, nullptr))) break;
        }
        if (h[2] > 0) {
            if ((rc = run_msa_gapped(c, d_basesP, d_refs, (const bbm_gapped_task*)gtasks.p, (const int32_t*)gaps.p, (bbm_msa_out*)gouts.p, h[2], nullptr, nullptr, st, nullptr))) break;
        }
        if (cudaMemsetAsync(counters.p, 0, 12, st) != cudaSuccess) { rc = fail(BBM_E_CUDA, "memset"); break; }
        if (trace) { cudaStreamSynchronize(st); fprintf(stderr, "[scoreSlow]   first pass done at %.2f ms\n", now() - t0); }
        if ((rc = launch(1, k)) || (rc = counts(h))) break;
        if (trace) fprintf(stderr, "[scoreSlow]   %d + %d (gapped) padding retries\n", h[1], h[2]);
        if (h[1] > 0) {
            if ((rc = run_msa(c, d_basesP, d_refs, (const bbm_msa_task*)tasks.p, (bbm_msa_out*)outs.p, h[1], nullptr, nullptr, max_read_len, 0, st, nullptr, nullptr))) break;
        }
        if (h[2] > 0) {
            if ((rc = run_msa_gapped(c, d_basesP, d_refs, (const bbm_gapped_task*)gtasks.p, (const int32_t*)gaps.p, (bbm_msa_out*)gouts.p, h[2], nullptr, nullptr, st, nullptr))) break;
        }
        if ((rc = launch(2, k))) break;
        if (trace) { cudaStreamSynchronize(st); fprintf(stderr, "[scoreSlow]   round done at %.2f ms\n", now() - t0); }
    }
    if (rc == BBM_OK) {          // alignments behind the results that were applied (requests thrown away by the look-ahead are not the reference's work)
        int na = 0;
        cudaError_t ce = cudaMemcpyAsync(&na, (const int*)counters.p + 4, 4, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) rc = fail(BBM_E_CUDA, "scoreSlow counters", ce);
        aligned = na;
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
