// msa_kernels.cu — task classification, the row-sequential generic kernel, and launch glue.
#include <cstdio>
#include <cstdlib>
#include <climits>
#include "msa_kernels.cuh"
#include "msa_generic.cuh"
#include "msa_narrow.cuh"

namespace bbm {

// pass 1: class of every task + per-class counts; pass 2: scatter ids into per-class lists
__global__ void msa_classify_kernel(MsaParams P, unsigned char* cls, unsigned int* cb, int useNarrow, int useStrip, int useBand) {
    __shared__ unsigned int local[NUM_CLASS];
    __shared__ unsigned int localNb[NARROW_BUCKETS];
    __shared__ unsigned int localSb[STRIP_BUCKETS];
    __shared__ unsigned long long localBytes;
    __shared__ unsigned int localBd[BAND_BUCKETS];
    if (threadIdx.x < BAND_BUCKETS) localBd[threadIdx.x] = 0;
    if (threadIdx.x < STRIP_BUCKETS) localSb[threadIdx.x] = 0;
    if (threadIdx.x == 0) localBytes = 0;
    if (threadIdx.x < NUM_CLASS) local[threadIdx.x] = 0;
    if (threadIdx.x < NARROW_BUCKETS) localNb[threadIdx.x] = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.ntasks) {
        TaskCtx T;
        const bbm_msa_task task = P.tasks[i];
        int k = CLASS_BAD;
        if (resolve_task(task, P.bandwidth, P.ratio, T)) {
            k = (useBand && band_eligible(T)) ? CLASS_BAND : ((useStrip && strip_eligible(T) && strip_bucket(T) < useStrip) ? CLASS_STRIP : classify(T));
            atomicAdd(&local[k], 1u);
            if (k == CLASS_BAND) { atomicMax(&cb[CB_BAND_MAXROWS], (unsigned)T.rows); atomicMax(&cb[CB_BAND_MAXCOLS], (unsigned)T.cols); atomicMax(&cb[CB_BAND_MAXHB], (unsigned)T.halfband); atomicAdd(&localBd[band_bucket(T)], 1u); }
            if (k == CLASS_GENERIC) { atomicMax(&cb[CB_GENERIC_MAXCOLS], (unsigned)T.cols); atomicMax(&cb[CB_GENERIC_MAXROWS], (unsigned)T.rows); }
            if (k == CLASS_STRIP) atomicAdd(&localBytes, strip_task_bytes(T.rows, T.cols));
            if (useNarrow && narrow_eligible(T, useNarrow > 1 ? useNarrow : 0)) { atomicAdd(&localNb[narrow_bucket(T.rows)], 1u); k |= CLS_NARROW_BIT; }
            else if (k == CLASS_STRIP) atomicAdd(&localSb[strip_bucket(T)], 1u);
        } else { bbm_msa_out o = {}; o.status = BBM_E_ARG; o.match_len = -1; P.outs[i] = o; }
        cls[i] = (unsigned char)k;
    }
    __syncthreads();
    if (threadIdx.x < NUM_CLASS && local[threadIdx.x]) atomicAdd(&cb[CB_COUNTS + threadIdx.x], local[threadIdx.x]);
    if (threadIdx.x < NARROW_BUCKETS && localNb[threadIdx.x]) atomicAdd(&cb[CB_NB_COUNTS + threadIdx.x], localNb[threadIdx.x]);
    if (threadIdx.x < STRIP_BUCKETS && localSb[threadIdx.x]) atomicAdd(&cb[CB_SB_COUNTS + threadIdx.x], localSb[threadIdx.x]);
    if (threadIdx.x == 0 && localBytes) atomicAdd(reinterpret_cast<unsigned long long*>(cb + CB_STRIP_BYTES), localBytes);
    if (threadIdx.x < BAND_BUCKETS && localBd[threadIdx.x]) atomicAdd(&cb[CB_BD_COUNTS + threadIdx.x], localBd[threadIdx.x]);
}

__global__ void msa_scatter_kernel(MsaParams P, const unsigned char* cls, unsigned int* cb, int* lists, int* nlist) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.ntasks) return;
    const int k = cls[i];
    if (k & CLS_NARROW_BIT) {
        const unsigned pos = atomicAdd(&cb[CB_NB_CURSORS + narrow_bucket(P.tasks[i].read_len)], 1u);
        nlist[pos] = (int)i;
        return;
    }
    if (k == CLASS_BAD) return;
    if (k == CLASS_STRIP) {
        TaskCtx T;
        resolve_task(P.tasks[i], P.bandwidth, P.ratio, T);
        const unsigned pos = atomicAdd(&cb[CB_SB_CURSORS + strip_bucket(T)], 1u);
        lists[pos] = (int)i;
        return;
    }
    if (k == CLASS_BAND) {
        TaskCtx T;
        resolve_task(P.tasks[i], P.bandwidth, P.ratio, T);
        const unsigned pos = atomicAdd(&cb[CB_BD_CURSORS + band_bucket(T)], 1u);
        lists[pos] = (int)i;
        return;
    }
    const unsigned pos = atomicAdd(&cb[CB_CURSORS + k], 1u);
    lists[pos] = (int)i;
}

// `endPtr` (optional): device cursor of the class list.  The narrow kernel hands alignments it cannot finish over to their class list on the
// device, and a narrow candidate that it does finish leaves its reserved slot unused: the true length is cursor - base, never more than nlist.
__global__ void msa_generic_kernel(MsaParams P, const int* list, int nlist, int* gscratch, long long gstride, const unsigned int* endPtr, unsigned int base, int skipWide, int wideRows) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (endPtr) nlist = min(nlist, (int)(*endPtr - base));
    if (i >= nlist) return;
    const int id = list[i];
    const bbm_msa_task task = P.tasks[id];
    TaskCtx T;
    if (!resolve_task(task, P.bandwidth, P.ratio, T)) return;
    if (skipWide && !T.limited && T.rows <= wideRows && P.dump == nullptr) return;      // taken by msa_wide_unlimited_kernel
    msa_generic_task(P, T, task, id, gscratch + (long long)i * gstride, gstride, P.outs + id);
}

// One alignment per block (one warp), rolling rows and limit vectors in dynamic shared memory.  Un-banded limited fills — the wide
// windows scoreSlow sends — are evaluated 32 columns at a time (msa_generic.cuh); anything else runs on lane 0 in the reference's own
// order, where shared memory still takes ~10 dependent L2 round trips per cell out of the chain.
__global__ void __launch_bounds__(32) msa_generic_smem_kernel(MsaParams P, const int* list, int nlist, int* gscratch, long long gstride, int smemInts,
                                                              const unsigned int* endPtr, unsigned int base, int skipWide, int wideRows) {
    extern __shared__ int fastbuf[];
    const int i = blockIdx.x;
    if (endPtr) nlist = min(nlist, (int)(*endPtr - base));
    if (i >= nlist) return;
    const int id = list[i];
    const bbm_msa_task task = P.tasks[id];
    TaskCtx T;
    if (!resolve_task(task, P.bandwidth, P.ratio, T)) return;
    if (skipWide && !T.limited && T.rows <= wideRows && P.dump == nullptr) return;      // taken by msa_wide_unlimited_kernel
    const bool fits = msa_generic_fast_ints(T.rows, T.cols) <= smemInts;
    msa_generic_task(P, T, task, id, gscratch + (long long)i * gstride, gstride, P.outs + id, fits ? fastbuf : nullptr, (int)threadIdx.x);
}


// ---------------- wide unlimited fills: one block per alignment, one THREAD PER READ ROW, skewed wavefront ----------------
// fillUnlimited (jni/MultiStateAligner11tsJNI.c:100-314) visits every cell and has no pruning state, so the only order that matters is the data
// dependence: (row, col) needs (row-1, col-1), (row, col-1), (row-1, col).  Thread t owns read row t+1 and walks its columns left to right; at step s it
// is on column s-t+1, so its left neighbour is its own previous value (registers), its upper neighbour is what thread t-1 wrote one step ago
// (shared memory, double-buffered) and its diagonal neighbour is the upper neighbour it read the step before.  rows+cols-1 steps, one barrier each;
// every gapped reference (500-700 columns, unlimited by the reference's own rule, MultiStateAligner11tsJNI.java:132-144) and every plain window
// wider than rows+170 takes this path instead of the row-sequential kernel.  Predecessor codes: 4 bits per cell, 8 cells per word, in the task's
// scratch block; score2 / traceback2 (…JNI.java:376-495, 537-658) walk them on one thread afterwards, exactly as in msa_generic.cuh.
constexpr int WIDE_THREADS = 640;          // >= MAXR rounded up to a warp multiple
__device__ __forceinline__ bool wide_takes(const TaskCtx& T, const MsaParams& P) { return !T.limited && T.rows <= WIDE_THREADS && T.rows >= 1 && P.dump == nullptr; }

__global__ void __launch_bounds__(WIDE_THREADS) msa_wide_unlimited_kernel(MsaParams P, const int* list, int nlist, int* gscratch, long long gstride,
                                                                          const unsigned int* endPtr, unsigned int base) {
    __shared__ int X[2][WIDE_THREADS][3];
    __shared__ int best[4];
    const int i = blockIdx.x;
    if (endPtr) nlist = min(nlist, (int)(*endPtr - base));
    if (i >= nlist) return;
    const int id = list[i];
    const bbm_msa_task task = P.tasks[id];
    TaskCtx T;
    if (!resolve_task(task, P.bandwidth, P.ratio, T)) return;
    if (!wide_takes(T, P) || T.rows > (int)blockDim.x) return;          // left to the row-sequential kernel
    const int rows = T.rows, cols = T.cols;
    const int wstride = (cols + 7) >> 3;
    if ((long long)(rows + 1) * wstride > gstride) { if (threadIdx.x == 0) { bbm_msa_out o = {}; o.status = BBM_E_SHAPE; o.match_len = -1; P.outs[id] = o; } return; }
    unsigned int* tbw = reinterpret_cast<unsigned int*>(gscratch + (long long)i * gstride);
    const int8_t* __restrict__ read = P.reads + task.read_off;
    const int8_t* __restrict__ ref = P.refs + task.ref_off + T.a;
    const int t = threadIdx.x, row = t + 1;
    const bool active = row <= rows;
    const int maxGain = (rows - 1) * P_MATCH2 + P_MATCH;
    const int subfloor = 0 - 2 * maxGain;
    const int call1 = active ? read[row - 1] : 0, call0 = (active && row >= 2) ? read[row - 2] : '?';
    const bool delBar = (row < 3) || (row > rows - 3);
    const bool insTop = row < 2, insBot = row > rows - 2;
    // column 0 of the matrix (MultiStateAligner11tsJNI.java:105-111): cumulative leading-insertion cost, all three states
    int lM = ins_score_offset(row), lD = lM, lI = lM;
    int dM = row == 1 ? 0 : ins_score_offset(row - 1), dD = dM, dI = dM;
    unsigned word = 0;
    int bS[3] = {INT_MIN, INT_MIN, INT_MIN}, bC[3] = {-1, -1, -1}, bP[3] = {0, 0, 0};      // last row: first maximum per state
    const int steps = rows + cols - 1;
    for (int s = 0; s < steps; ++s) {
        const int col = s - t + 1;
        if (active && col >= 1 && col <= cols) {
            int uM = 0, uD = 0, uI = 0;                                   // row 0 of the matrix is all zero
            if (row > 1) { const int* x = X[(s - 1) & 1][t - 1]; uM = x[0]; uD = x[1]; uI = x[2]; }
            const int r1 = ref[col - 1];
            const int r0 = col < 2 ? '!' : ref[col - 2];
            const bool gap = (r1 == '-'), match = (call1 == r1 && r1 != 'N'), prevMatch = (call0 == r0 && r0 != 'N');
            unsigned code = 0;
            int nM, nD, nI;
            {   // MS (jni/...JNI.c:137-213)
                const int sM = dM & SMASK, sD = dD & SMASK, sI = dI & SMASK, streak = dM & TMASK;
                if (gap) nM = subfloor;
                else {
                    int a_, o;
                    if (match) { a_ = sM + (prevMatch ? P_MATCH2 : P_MATCH); o = P_MATCH; }
                    else {
                        a_ = sM + ((r1 != 'N' && call1 != 'N') ? (prevMatch ? (streak <= 1 ? P_SUBR : P_SUB) : (streak == 0 ? P_SUB : (streak < 5 ? P_SUB2 : P_SUB3))) : 0);
                        o = P_SUB;
                    }
                    const int b_ = sD + o, c_ = sI + o;
                    int score, time;
                    if (a_ >= b_ && a_ >= c_) { score = a_; time = (match == prevMatch) ? streak + 1 : 1; }
                    else if (b_ >= c_) { score = b_; time = 1; }
                    else { score = c_; time = 1; }
                    if (time > MAX_TIME) time = TIME_WRAP;
                    nM = score | time;
                    code |= (time > 1) ? 0u : ((sM >= sD && sM >= sI) ? 0u : (sD >= sI ? 1u : 2u));
                }
            }
            {   // DEL (:215-256)
                const int sM = lM & SMASK, sD = lD & SMASK, streak = lD & TMASK;
                if (delBar) nD = subfloor;
                else {
                    int a_ = sM + P_DEL;
                    int b_ = sD + (streak == 0 ? P_DEL : (streak < LIM3 ? P_DEL2 : (streak < LIM4 ? P_DEL3 : (streak < LIM5 ? P_DEL4 : (((streak & 3) == 0) ? P_DEL5 : 0)))));
                    if (r1 == 'N') { a_ += P_DEL_REF_N; b_ += P_DEL_REF_N; } else if (gap) { a_ += P_GAP; b_ += P_GAP; }
                    int score, time;
                    if (a_ >= b_) { score = a_; time = 1; } else { score = b_; time = streak + 1; }
                    if (time > MAX_TIME) time = TIME_WRAP;
                    nD = score | time;
                    code |= ((time > 1) ? 1u : (sM >= sD ? 0u : 1u)) << 2;
                }
            }
            {   // INS (:258-288)
                const int sM = uM & SMASK, sI = uI & SMASK, streak = uI & TMASK;
                if (gap || (insTop && col > 1) || (insBot && col < cols - 1)) nI = subfloor;
                else {
                    const int a_ = sM + P_INS;
                    const int b_ = sI + (streak == 0 ? P_INS : (streak < LIM3 ? P_INS2 : (streak < LIM4 ? P_INS3 : P_INS4)));
                    int score, time;
                    if (a_ >= b_) { score = a_; time = 1; } else { score = b_; time = streak + 1; }
                    if (time > MAX_TIME) time = TIME_WRAP;
                    nI = score | time;
                    code |= ((time > 1) ? 1u : (sM >= sI ? 0u : 1u)) << 3;
                }
            }
            int* xo = X[s & 1][t]; xo[0] = nM; xo[1] = nD; xo[2] = nI;
            word |= code << (4 * ((col - 1) & 7));
            if (((col - 1) & 7) == 7 || col == cols) { tbw[(long long)row * wstride + ((col - 1) >> 3)] = word; word = 0; }
            if (row == rows) {
                const int v[3] = {nM, nD, nI};
#pragma unroll
                for (int st = 0; st < 3; ++st) { const int x = v[st] & SMASK; if (x > bS[st]) { bS[st] = x; bC[st] = col; bP[st] = v[st]; } }
            }
            dM = uM; dD = uD; dI = uI;
            lM = nM; lD = nD; lI = nI;
        }
        __syncthreads();
    }
    if (active && row == rows) {          // final scan order (jni/...JNI.c:297-311): states 0,1,2, columns ascending, strict >
        int st = 0;
        if (bS[1] > bS[st]) st = 1;
        if (bS[2] > bS[st]) st = 2;
        best[0] = bS[st]; best[1] = bC[st]; best[2] = st; best[3] = bP[st];
    }
    __syncthreads();
    if (t != 0) return;
    const int maxScore = best[0], maxCol = best[1], maxState = best[2], maxPacked = best[3];
    bbm_msa_out* out = P.outs + id;
    out->path = 1; out->iterations = (long long)rows * cols; out->status = 0; out->score_len = 0; out->match_len = -1; out->pad_ = 0;
    for (int k = 0; k < 8; ++k) out->score[k] = 0;
    out->result[0] = rows; out->result[1] = maxCol; out->result[2] = maxState; out->result[3] = maxScore >> TBITS; out->result[4] = 0;
    if ((T.flags & (BBM_TF_SCORE | BBM_TF_TRACEBACK)) == 0) return;
    const bool wantTb = (T.flags & BBM_TF_TRACEBACK) != 0 && P.match_buf != nullptr;
    int8_t* mslot = nullptr; long long mcap = 0;
    if (wantTb) { mslot = P.match_buf + P.match_off[id]; mcap = P.match_off[id + 1] - P.match_off[id]; }
    int r = rows, col = maxCol, state = maxState, stateTime = 0, nOps = 0, gapsSeen = 0;
    const int bestRefStop = T.a + col - 1;
    while (r > 0 && col > 0) {
        const unsigned code = (tbw[(long long)r * wstride + ((col - 1) >> 3)] >> (4 * ((col - 1) & 7))) & 15u;
        int prev; char op = 0;
        if (state == ST_MS) {
            prev = code & 3u;
            const int c = read[r - 1], rf = ref[col - 1];
            op = (c == rf) ? 'm' : ((!base_defined(c) || !base_defined(rf)) ? 'N' : 'S');
            r--; col--;
        } else if (state == ST_DEL) {
            prev = ((code >> 2) & 1u) ? ST_DEL : ST_MS;
            if (ref[col - 1] == '-') { op = '-'; gapsSeen++; } else op = 'D';
            col--;
        } else {
            prev = ((code >> 3) & 1u) ? ST_INS : ST_MS;
            op = (col == 0) ? 'X' : ((col >= cols) ? 'Y' : 'I');
            r--;
        }
        if (wantTb && nOps < mcap) mslot[mcap - 1 - nOps] = op;
        nOps++;
        stateTime = (state == prev) ? stateTime + 1 : 0;
        state = prev;
    }
    const int rowEnd = r, colEnd = col;
    if (wantTb && colEnd != rowEnd) { int rr = rowEnd; while (rr > 0) { if (nOps < mcap) mslot[mcap - 1 - nOps] = 'X'; nOps++; rr--; } }
    if (T.flags & BBM_TF_SCORE) {
        int colf = colEnd; if (rowEnd > colEnd) colf -= rowEnd;
        const int bestRefStart = T.a + colf;
        int padLeft = 0, padRight = 0;
        if (bestRefStart < T.a) padLeft = max(0, T.a - bestRefStart);
        else if (bestRefStart == T.a && state == ST_INS) padLeft = stateTime;
        if (bestRefStop > score_ref_end(T)) padRight = max(0, bestRefStop - score_ref_end(T));
        else if (bestRefStop == score_ref_end(T) && maxState == ST_INS) padRight = maxPacked & TMASK;
        out->score[0] = maxScore >> TBITS; out->score[1] = bestRefStart; out->score[2] = bestRefStop;
        out->score[3] = rows; out->score[4] = maxCol; out->score[5] = maxState; out->score[6] = padLeft; out->score[7] = padRight;
        out->score_len = (padLeft > 0 || padRight > 0) ? 8 : 6;
    }
    if (!wantTb) return;
    const long long total = (long long)nOps + (long long)gapsSeen * 127;
    if (nOps > mcap || total > mcap) { out->status = BBM_E_CAPACITY; out->match_len = -1; return; }
    const long long src = mcap - nOps;
    long long j = 0;
    for (int k2 = 0; k2 < nOps; ++k2) {
        const int8_t c = mslot[src + k2];
        if (c != '-') mslot[j++] = c; else for (int k = 0; k < 128; ++k) mslot[j++] = 'D';
    }
    out->match_len = (int)total;
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_launch_msa_classify(const MsaParams* P, unsigned char* cls, unsigned int* cb, int useNarrow, int useStrip, int useBand, cudaStream_t stream) {
    const int threads = 256;
    msa_classify_kernel<<<(unsigned)((P->ntasks + threads - 1) / threads), threads, 0, stream>>>(*P, cls, cb, useNarrow, useStrip, useBand);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_msa_scatter(const MsaParams* P, const unsigned char* cls, unsigned int* cb, int* lists, int* nlist, cudaStream_t stream) {
    const int threads = 256;
    msa_scatter_kernel<<<(unsigned)((P->ntasks + threads - 1) / threads), threads, 0, stream>>>(*P, cls, cb, lists, nlist);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_msa_narrow(const MsaParams* P, const int* nlist, int n, unsigned int* cb, unsigned long long* tb, long long tbWordsPerWarp,
                                     int* lists, int blocks, int useStrip, cudaStream_t stream) {
    msa_narrow_kernel<<<blocks, NARROW_THREADS, 0, stream>>>(*P, nlist, n, cb + CB_NARROW_WORK, tb, tbWordsPerWarp, cb + CB_CURSORS, lists, useStrip);
    return (int)cudaGetLastError();
}
extern "C" int bbm_msa_narrow_threads() { return NARROW_THREADS; }
extern "C" int bbm_msa_narrow_buckets() { return NARROW_BUCKETS; }
extern "C" int bbm_launch_msa_generic(const MsaParams* P, const int* list, int nlist, int* gscratch, long long gstride, cudaStream_t stream,
                                      int max_rows, int max_cols, const unsigned int* endPtr, unsigned int base) {
    // few, long alignments (the usual case: a handful of wide windows per batch): one per block with its rows in shared memory;
    // many alignments: the thread-per-alignment form keeps more of them in flight
    // unlimited fills first: one block per alignment, one thread per read row (msa_wide_unlimited_kernel); what is left (limited fills wider than
    // 512 columns, reads longer than 640, banded re-runs, the dump path) goes through the row-sequential kernels below
    int skipWide = 0, wideRows = 0;
    if (P->dump == nullptr && getenv("BBM_NO_WIDE") == nullptr) {
        wideRows = max_rows < WIDE_THREADS ? ((max_rows + 31) / 32) * 32 : WIDE_THREADS;
        msa_wide_unlimited_kernel<<<nlist, wideRows, 0, stream>>>(*P, list, nlist, gscratch, gstride, endPtr, base);
        cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return (int)e;
        skipWide = 1;
    }
    const long long fastInts = msa_generic_fast_ints(max_rows, max_cols);
    const size_t smem = (size_t)fastInts * 4;
    if (nlist <= 4096 && smem <= 200 * 1024) {
        // per device, and cheap: set on every launch rather than once per process (one context per GPU may live in the same process)
        cudaFuncSetAttribute(msa_generic_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        msa_generic_smem_kernel<<<nlist, 32, smem, stream>>>(*P, list, nlist, gscratch, gstride, (int)fastInts, endPtr, base, skipWide, wideRows);
        return (int)cudaGetLastError();
    }
    const int threads = 64;
    msa_generic_kernel<<<(nlist + threads - 1) / threads, threads, 0, stream>>>(*P, list, nlist, gscratch, gstride, endPtr, base, skipWide, wideRows);
    return (int)cudaGetLastError();
}
extern "C" int bbm_msa_warps_per_block() { return WARPS_PER_BLOCK; }
extern "C" int bbm_msa_num_wclass() { return NUM_WCLASS; }
extern "C" int bbm_msa_class_strip() { return CLASS_STRIP; }
extern "C" int bbm_msa_class_band() { return CLASS_BAND; }
extern "C" int bbm_msa_wclass_width(int k) { return wclass_width(k); }
extern "C" long long bbm_generic_scratch_ints(int rows, int cols) { return msa_generic_scratch_ints(rows, cols); }

// sum of the reference's cell counter over a batch (diagnostics for the roofline of a chained step)
__global__ void __launch_bounds__(256) msa_sum_iterations_kernel(const bbm_msa_out* __restrict__ outs, long long n, unsigned long long* __restrict__ sum) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    for (; i < n; i += (long long)gridDim.x * blockDim.x) if (outs[i].status == 0) v += (unsigned long long)outs[i].iterations;
    for (int o = 16; o; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(sum, v);
}
extern "C" int bbm_launch_msa_sum_iterations(const bbm_msa_out* outs, long long n, unsigned long long* sum, cudaStream_t st) {
    const int blocks = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    msa_sum_iterations_kernel<<<blocks, 256, 0, st>>>(outs, n, sum);
    return (int)cudaGetLastError();
}
