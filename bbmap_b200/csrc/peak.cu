// peak.cu — integer / DPX issue-rate microbenchmarks: the roofline denominators for the DP kernels.
// MEASURED_PEAKS.json has no integer peak, so bench.py measures one in the same run.
// Each kernel runs ITER iterations of 8 independent dependency chains per thread, so the pipes — not latency — bound it.
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

enum { PK_IADD3 = 0, PK_LOP3 = 1, PK_VIMNMX3 = 2, PK_VIADDMAX = 3, PK_IMAD = 4, PK_MIX_ALU_FMA = 5, PK_SEL = 6 };

template <int KIND>
__global__ void __launch_bounds__(256) peak_kernel(int iters, int seed, int* out) {
    int a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * (i + 1);
    int b = seed ^ 0x5bd1e995, c = seed * 31 + 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == PK_IADD3) a[i] = a[i] + b + c;                                        // IADD3
            else if (KIND == PK_LOP3) a[i] = (a[i] & b) ^ c;                                  // LOP3
            else if (KIND == PK_VIMNMX3) a[i] = __vimax3_s32(a[i], b, c ^ a[(i + 1) & 7]);    // VIMNMX3 (DPX)
            else if (KIND == PK_VIADDMAX) a[i] = __viaddmax_s32(a[i], b, c);                  // VIADDMNMX (DPX)
            else if (KIND == PK_IMAD) a[i] = a[i] * b + c;                                    // IMAD (fma pipe)
            else if (KIND == PK_SEL) a[i] = (a[i] > b) ? (a[i] - c) : (a[i] + c);             // ISETP + SEL-ish
            else { a[i] = (i & 1) ? (a[i] * b + c) : ((a[i] & b) ^ c); }                      // half fma pipe, half alu pipe
        }
        b += it; c ^= it;
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x7fffffff) out[0] = s;     // keep the work alive
}

}  // namespace bbm

using namespace bbm;

// Returns lane-operations per second (one "operation" = the per-lane result of one instruction of the measured kind).
extern "C" int bbm_launch_peak(int kind, int blocks, int iters, int* d_out, cudaStream_t st) {
    switch (kind) {
        case PK_IADD3: peak_kernel<PK_IADD3><<<blocks, 256, 0, st>>>(iters, 12345, d_out); break;
        case PK_LOP3: peak_kernel<PK_LOP3><<<blocks, 256, 0, st>>>(iters, 12345, d_out); break;
        case PK_VIMNMX3: peak_kernel<PK_VIMNMX3><<<blocks, 256, 0, st>>>(iters, 12345, d_out); break;
        case PK_VIADDMAX: peak_kernel<PK_VIADDMAX><<<blocks, 256, 0, st>>>(iters, 12345, d_out); break;
        case PK_IMAD: peak_kernel<PK_IMAD><<<blocks, 256, 0, st>>>(iters, 12345, d_out); break;
        case PK_SEL: peak_kernel<PK_SEL><<<blocks, 256, 0, st>>>(iters, 12345, d_out); break;
        default: peak_kernel<PK_MIX_ALU_FMA><<<blocks, 256, 0, st>>>(iters, 12345, d_out); break;
    }
    return (int)cudaGetLastError();
}
