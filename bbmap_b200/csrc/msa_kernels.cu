// msa_kernels.cu — task classification, the row-sequential generic kernel, and launch glue.
#include <cstdio>
#include "msa_kernels.cuh"
#include "msa_generic.cuh"

namespace bbm {

// pass 1: class of every task + per-class counts; pass 2: scatter ids into per-class lists
__global__ void msa_classify_kernel(MsaParams P, unsigned char* cls, unsigned int* counts) {
    __shared__ unsigned int local[NUM_CLASS];
    if (threadIdx.x < NUM_CLASS) local[threadIdx.x] = 0;
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.ntasks) {
        TaskCtx T;
        const bbm_msa_task task = P.tasks[i];
        int k = CLASS_BAD;
        if (resolve_task(task, P.bandwidth, P.ratio, T)) k = classify(T);
        else { bbm_msa_out o = {}; o.status = BBM_E_ARG; o.match_len = -1; P.outs[i] = o; }
        cls[i] = (unsigned char)k;
        atomicAdd(&local[k], 1u);
    }
    __syncthreads();
    if (threadIdx.x < NUM_CLASS && local[threadIdx.x]) atomicAdd(&counts[threadIdx.x], local[threadIdx.x]);
}

__global__ void msa_scatter_kernel(long long ntasks, const unsigned char* cls, unsigned int* cursors, int* lists) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ntasks) return;
    const int k = cls[i];
    if (k >= CLASS_BAD) return;
    const unsigned pos = atomicAdd(&cursors[k], 1u);
    lists[pos] = (int)i;
}

__global__ void msa_generic_kernel(MsaParams P, const int* list, int nlist, int* gscratch, long long gstride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlist) return;
    const int id = list[i];
    const bbm_msa_task task = P.tasks[id];
    TaskCtx T;
    if (!resolve_task(task, P.bandwidth, P.ratio, T)) return;
    msa_generic_task(P, T, task, id, gscratch + (long long)i * gstride, gstride, P.outs + id);
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_launch_msa_classify(const MsaParams* P, unsigned char* cls, unsigned int* counts, cudaStream_t stream) {
    const int threads = 256;
    msa_classify_kernel<<<(unsigned)((P->ntasks + threads - 1) / threads), threads, 0, stream>>>(*P, cls, counts);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_msa_scatter(long long ntasks, const unsigned char* cls, unsigned int* cursors, int* lists, cudaStream_t stream) {
    const int threads = 256;
    msa_scatter_kernel<<<(unsigned)((ntasks + threads - 1) / threads), threads, 0, stream>>>(ntasks, cls, cursors, lists);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_msa_generic(const MsaParams* P, const int* list, int nlist, int* gscratch, long long gstride, cudaStream_t stream) {
    const int threads = 64;
    msa_generic_kernel<<<(nlist + threads - 1) / threads, threads, 0, stream>>>(*P, list, nlist, gscratch, gstride);
    return (int)cudaGetLastError();
}
extern "C" int bbm_msa_warps_per_block() { return WARPS_PER_BLOCK; }
extern "C" int bbm_msa_num_wclass() { return NUM_WCLASS; }
extern "C" int bbm_msa_wclass_width(int k) { return wclass_width(k); }
extern "C" long long bbm_generic_scratch_ints(int rows, int cols) { return msa_generic_scratch_ints(rows, cols); }
