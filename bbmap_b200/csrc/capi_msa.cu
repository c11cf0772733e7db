// capi_msa.cu — MultiStateAligner11ts: batched fills (run_msa), gapped references, single-alignment twins
// Part of the C ABI of libbbmapcuda.so (include/bbmap_cuda.h): host-side glue only (device buffers, streams, launches).
// No CPU implementation of any compute path lives here: without a device every call fails loudly.
#include "ctx.h"

// Counter block layout: see CB_* in msa_kernels.cuh.
int run_msa(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_msa_task* d_tasks, bbm_msa_out* d_outs,
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
    const int CS = bbm_msa_class_strip(), CB = bbm_msa_class_band();
    const int useBand = (c->use_band && d_dump == nullptr && (c->bandwidth > 0 || c->ratio > 0.f)) ? 1 : 0;
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
    if (useBand) CK(cudaMemsetAsync(cb + 256, 0, 256 * 4, st));
    CK(cudaEventRecord(c->ev0, st));
    int e = bbm_launch_msa_classify(&P, (unsigned char*)c->cls.p, cb, useNarrow, useStrip, useBand, st);
    if (e) return fail(BBM_E_CUDA, "msa_classify_kernel launch", (cudaError_t)e);
    c->launches++;
    unsigned int h[512];
    CK(cudaMemcpyAsync(h, cb, (useBand ? 512 : 192) * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    unsigned int base[16]; unsigned int acc = 0;
    for (int k = 0; k < 16; ++k) { base[k] = acc; if (k <= nw || k == CS || k == CB) acc += h[k]; }
    unsigned int nbase[64]; unsigned int nacc = 0;
    for (int k = 0; k < 64; ++k) { nbase[k] = nacc; if (k < nb) nacc += h[64 + k]; }
    unsigned int curs[16]; memcpy(curs, base, sizeof(curs));
    unsigned int sbBase[16];
    {   // strip list: direct tasks ordered by estimated work (largest bucket first), narrow-kernel hand-overs appended after them
        unsigned int cur = base[CS];
        for (int b = 15; b >= 0; --b) { sbBase[b] = cur; cur += h[104 + b]; }
        curs[CS] = cur;
    }
    unsigned int bdBase[120], bdClass[4] = {0, 0, 0, 0};
    if (useBand && h[CB]) {       // band list: slot class, then read length (the 32 alignments of a warp then have the same shape)
        unsigned int cur = base[CB];
        for (int b = 0; b < 120; ++b) { bdBase[b] = cur; cur += h[256 + b]; bdClass[b / 40] += h[256 + b]; }
        CK(cudaMemcpyAsync(cb + 384, bdBase, 120 * 4, cudaMemcpyHostToDevice, st));
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
    if (useBand && h[CB]) {
        // banded limited fills: one thread per alignment, the band by diagonal in shared memory (msa_band.cu); no right-edge assumption, no re-runs.
        // One launch per slot class present (32 / 64 / 128 slots of shared memory per thread and state).
        const int bRows = (int)h[53], bCols = (int)h[54];
        const int BT = bbm_msa_band_threads();
        unsigned int segBase = base[CB];
        for (int cls = 0; cls < 3; ++cls) {
            const unsigned int cnt = bdClass[cls];
            if (!cnt) continue;
            const int nd = 32 << cls;
            const size_t smemPerBlock = bbm_msa_band_smem_bytes(bRows, bCols, nd) + 1024;       // + the per-block reservation of the runtime
            int perSm = (int)((227 * 1024) / smemPerBlock); if (perSm < 1) perSm = 1; if (perSm > 12) perSm = 12;
            long long blocks = (long long)c->sms * perSm;
            const long long needB = ((long long)cnt + BT - 1) / BT;
            if (needB < blocks) blocks = needB;
            if (c->bandScratch.ensure((size_t)blocks * BT * bbm_msa_band_thread_bytes(bRows, bCols, nd) + 256)) return fail(BBM_E_CUDA, "cudaMalloc band scratch");
            e = bbm_launch_msa_band(&P, (const int*)c->lists.p + segBase, (int)cnt, nullptr, segBase, nd, bRows, bCols, c->bandScratch.p, cb + 56 + cls, (int)blocks, st);
            if (e) return fail(BBM_E_CUDA, "msa_band_kernel launch", (cudaError_t)e);
            c->launches++;
            segBase += cnt;
        }
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
        // the budget also follows what the device has left NOW (other contexts of the process hold their own scratch): what this context already owns
        // plus 70 % of the free memory, so that a later context works in more, smaller chunks instead of failing its allocation
        unsigned long long budget = c->strip_budget;
        {
            size_t freeB = 0, totalB = 0;
            static const bool noFollow = getenv("BBM_NO_ADAPTIVE_BUDGET") != nullptr;
            // only when the scratch would have to grow: cudaMemGetInfo is a synchronous driver query, not something for every launch of a steady-state step
            const bool wouldGrow = totalBytes + bbm_msa_strip_fixed_bytes((int)std::min<long long>(nstrip, 0x7fffffff), sRows, blocksMax) + 4096 > (unsigned long long)c->stripScratch.cap;
            if (!noFollow && wouldGrow && cudaMemGetInfo(&freeB, &totalB) == cudaSuccess) {
                const unsigned long long avail = (unsigned long long)c->stripScratch.cap + (unsigned long long)(freeB * 0.7);
                if (avail < budget) budget = avail;
            }
            const unsigned long long fixedMax = bbm_msa_strip_fixed_bytes((int)std::min<long long>(nstrip, 0x7fffffff), sRows, blocksMax) + 4096;
            budget = budget > 2 * fixedMax ? budget - fixedMax : fixedMax;
        }
        if (totalBytes > budget) { chunk = (long long)(budget / perMax); if (chunk < 1024) chunk = 1024; if (chunk > nstrip) chunk = nstrip; }
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
    { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); if (ms_out) *ms_out = ms; c->msa_ms += ms; }
    if (c->msa_count) {             // diagnostics ("msa_count" option): the reference's cell counter summed over the batch, for the roofline of a chained step
        if (c->msaCells.ensure(8)) return fail(BBM_E_CUDA, "cudaMalloc cell counter");
        CK(cudaMemsetAsync(c->msaCells.p, 0, 8, st));
        int e3 = bbm_launch_msa_sum_iterations(d_outs, ntasks, (unsigned long long*)c->msaCells.p, st);
        if (e3) return fail(BBM_E_CUDA, "msa_sum_iterations launch", (cudaError_t)e3);
        unsigned long long cells = 0;
        CK(cudaMemcpyAsync(&cells, c->msaCells.p, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        c->msa_cells += (long long)cells;
    }
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

int run_msa_gapped(bbm_ctx* c, const int8_t* d_reads, const int8_t* d_refs, const bbm_gapped_task* d_gt, const int32_t* d_gaps,
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
