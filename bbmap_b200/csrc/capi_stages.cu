// capi_stages.cu — BandedAligner, KeyRing seeding, scoreNoIndels, tip/rescue scans, read ingest, SAM record fields
// Part of the C ABI of libbbmapcuda.so (include/bbmap_cuda.h): host-side glue only (device buffers, streams, launches).
// No CPU implementation of any compute path lives here: without a device every call fails loudly.
#include "ctx.h"

// =====================  BandedAligner  =====================
int run_banded(bbm_ctx* c, const int8_t* dq, const int8_t* dr, const bbm_band_task* dt, bbm_band_out* dout, int64_t n,
                      cudaStream_t st, float* ms_out) {
    if (n <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    unsigned int* cb = (unsigned int*)c->counters.p;
    CK(cudaMemsetAsync(cb + 200, 0, 16, st));          // [200] task counter, [201] wide-pair count, [202] widest band of the batch, [203] counter of the wide pass
    CK(cudaEventRecord(c->ev0, st));
    int e;
    if (c->banded_thread) {
        // narrow bands (Dedupe: 3-9 cells): one thread per pair (banded_thread_kernel); the widest band of the batch picks the instantiation and tells
        // whether any pair needs the warp-per-pair kernel at all
        int sb = c->sms * 4; if ((n + 255) / 256 < sb) sb = (int)((n + 255) / 256);
        e = bbm_launch_banded_maxwidth(dt, n, cb + 202, sb, st);
        if (e) return fail(BBM_E_CUDA, "banded_maxwidth_kernel launch", (cudaError_t)e);
        unsigned int maxBand = 0;
        CK(cudaMemcpyAsync(&maxBand, cb + 202, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const bool anyWide = maxBand > 15;
        if (anyWide && c->bandedWide.ensure((size_t)n * 4 + 16)) return fail(BBM_E_CUDA, "cudaMalloc banded wide list");
        int blocks = c->sms * 8;
        const long long need = (n + 127) / 128;
        if (need < blocks) blocks = (int)need;
        e = bbm_launch_banded_thread(dq, dr, dt, dout, n, cb + 200, (int*)c->bandedWide.p, cb + 201, (int)maxBand, blocks, st);
        if (e) return fail(BBM_E_CUDA, "banded_thread_kernel launch", (cudaError_t)e);
        c->launches += 2;
        if (anyWide) {
            e = bbm_launch_banded(dq, dr, dt, dout, n, cb + 203, c->sms * 8, st, (const int*)c->bandedWide.p, cb + 201);
            if (e) return fail(BBM_E_CUDA, "banded_kernel launch", (cudaError_t)e);
            c->launches++;
        }
    } else {
        int blocks = c->sms * 8;
        const long long need = (n + 3) / 4;
        if (need < blocks) blocks = (int)need;
        e = bbm_launch_banded(dq, dr, dt, dout, n, cb + 200, blocks, st, nullptr, nullptr);
        if (e) return fail(BBM_E_CUDA, "banded_kernel launch", (cudaError_t)e);
        c->launches++;
    }
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
int run_seed(bbm_ctx* c, const int8_t* db, const int8_t* dq, const int64_t* doff, int64_t nreads, int max_len, const bbm_seed_cfg* cfg,
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
int run_noindel(bbm_ctx* c, const int8_t* dr, const int8_t* dref, const bbm_noindel_task* dt, int* ds, int8_t* dm, const int64_t* dmo,
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


// =====================  read ingest (Read.validate + reverse complement, a0)  =====================
int run_ingest(bbm_ctx* c, int8_t* db, int8_t* dq, const int64_t* doff, int64_t nreads, int max_len, int flags, int8_t* dm, int* df,
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
int run_sam(bbm_ctx* c, const bbm_sam_task* dt, int64_t n, const int8_t* dm, const int* dso, const int* dsl, const int* dsn, int nchroms,
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

