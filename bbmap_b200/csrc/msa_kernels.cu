// msa_kernels.cu — task classification, the row-sequential generic kernel, and launch glue.
#include <cstdio>
#include "msa_kernels.cuh"
#include "msa_generic.cuh"
#include "msa_narrow.cuh"

namespace bbm {

// pass 1: class of every task + per-class counts; pass 2: scatter ids into per-class lists
__global__ void msa_classify_kernel(MsaParams P, unsigned char* cls, unsigned int* cb, int useNarrow, int useStrip) {
    __shared__ unsigned int local[NUM_CLASS];
    __shared__ unsigned int localNb[NARROW_BUCKETS];
    __shared__ unsigned int localSb[STRIP_BUCKETS];
    __shared__ unsigned long long localBytes;
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
            k = (useStrip && strip_eligible(T) && strip_bucket(T) < useStrip) ? CLASS_STRIP : classify(T);
            atomicAdd(&local[k], 1u);
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
    const unsigned pos = atomicAdd(&cb[CB_CURSORS + k], 1u);
    lists[pos] = (int)i;
}

// `endPtr` (optional): device cursor of the class list.  The narrow kernel hands alignments it cannot finish over to their class list on the
// device, and a narrow candidate that it does finish leaves its reserved slot unused: the true length is cursor - base, never more than nlist.
__global__ void msa_generic_kernel(MsaParams P, const int* list, int nlist, int* gscratch, long long gstride, const unsigned int* endPtr, unsigned int base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (endPtr) nlist = min(nlist, (int)(*endPtr - base));
    if (i >= nlist) return;
    const int id = list[i];
    const bbm_msa_task task = P.tasks[id];
    TaskCtx T;
    if (!resolve_task(task, P.bandwidth, P.ratio, T)) return;
    msa_generic_task(P, T, task, id, gscratch + (long long)i * gstride, gstride, P.outs + id);
}

// One alignment per block (one warp), rolling rows and limit vectors in dynamic shared memory.  Un-banded limited fills — the wide
// windows scoreSlow sends — are evaluated 32 columns at a time (msa_generic.cuh); anything else runs on lane 0 in the reference's own
// order, where shared memory still takes ~10 dependent L2 round trips per cell out of the chain.
__global__ void __launch_bounds__(32) msa_generic_smem_kernel(MsaParams P, const int* list, int nlist, int* gscratch, long long gstride, int smemInts,
                                                              const unsigned int* endPtr, unsigned int base) {
    extern __shared__ int fastbuf[];
    const int i = blockIdx.x;
    if (endPtr) nlist = min(nlist, (int)(*endPtr - base));
    if (i >= nlist) return;
    const int id = list[i];
    const bbm_msa_task task = P.tasks[id];
    TaskCtx T;
    if (!resolve_task(task, P.bandwidth, P.ratio, T)) return;
    const bool fits = msa_generic_fast_ints(T.rows, T.cols) <= smemInts;
    msa_generic_task(P, T, task, id, gscratch + (long long)i * gstride, gstride, P.outs + id, fits ? fastbuf : nullptr, (int)threadIdx.x);
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_launch_msa_classify(const MsaParams* P, unsigned char* cls, unsigned int* cb, int useNarrow, int useStrip, cudaStream_t stream) {
    const int threads = 256;
    msa_classify_kernel<<<(unsigned)((P->ntasks + threads - 1) / threads), threads, 0, stream>>>(*P, cls, cb, useNarrow, useStrip);
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
    const long long fastInts = msa_generic_fast_ints(max_rows, max_cols);
    const size_t smem = (size_t)fastInts * 4;
    if (nlist <= 4096 && smem <= 200 * 1024) {
        // per device, and cheap: set on every launch rather than once per process (one context per GPU may live in the same process)
        cudaFuncSetAttribute(msa_generic_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        msa_generic_smem_kernel<<<nlist, 32, smem, stream>>>(*P, list, nlist, gscratch, gstride, (int)fastInts, endPtr, base);
        return (int)cudaGetLastError();
    }
    const int threads = 64;
    msa_generic_kernel<<<(nlist + threads - 1) / threads, threads, 0, stream>>>(*P, list, nlist, gscratch, gstride, endPtr, base);
    return (int)cudaGetLastError();
}
extern "C" int bbm_msa_warps_per_block() { return WARPS_PER_BLOCK; }
extern "C" int bbm_msa_num_wclass() { return NUM_WCLASS; }
extern "C" int bbm_msa_class_strip() { return CLASS_STRIP; }
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
