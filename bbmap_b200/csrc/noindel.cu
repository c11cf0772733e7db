// noindel.cu — ungapped scoring of a read against a candidate site.
// MSA.scoreNoIndels                       current/align2/MultiStateAligner11tsJNI.java:1033-1089
// MSA.scoreNoIndelsAndMakeMatchString     :1243-1318   (returns -99999 when the read runs outside the reference array)
// One thread per (read, site): a single linear scan, MATCH 70 for the first match of a run and MATCH2 100 afterwards,
// POINTS_SUB_ARRAY[timeInMode+1] for substitutions, no-calls / no-refs score 0 and do not change the mode.
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

__global__ void __launch_bounds__(128) noindel_kernel(const int8_t* __restrict__ reads, const int8_t* __restrict__ refs,
                                                      const bbm_noindel_task* __restrict__ tasks, int* __restrict__ scores,
                                                      int8_t* __restrict__ match_buf, const long long* __restrict__ match_off, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bbm_noindel_task T = tasks[i];
    const int8_t* read = reads + T.read_off;
    const int8_t* ref = refs + T.ref_off;
    const int len = T.read_len, refStart = T.ref_start, refLen = T.ref_len;
    const bool wantMatch = (T.flags & 1) != 0 && match_buf != nullptr;
    int8_t* match = wantMatch ? match_buf + match_off[i] : nullptr;
    int readStart = 0, readStop = len;
    const long long refStop = (long long)refStart + len;
    if (wantMatch && (refStart < 0 || refStop > refLen)) { scores[i] = -99999; return; }
    if (refStart < 0) readStart = -refStart;                      // POINTS_NOREF == 0
    if (refStop > refLen) readStop -= (int)(refStop - refLen);
    int score = 0, mode = -1, timeInMode = 0;
    for (int k = readStart; k < readStop; ++k) {
        const int c = read[k], r = ref[refStart + k];
        char m;
        if (c == r && c != 'N') {
            if (mode == 0) { timeInMode++; score += 100; } else { timeInMode = 0; score += 70; }
            mode = 0; m = 'm';
        } else if (c < 0 || c == 'N') { m = 'N'; }
        else if (r < 0 || r == 'N') { m = 'N'; }
        else {
            if (mode == 3) timeInMode++; else timeInMode = 0;
            score += timeInMode == 0 ? -127 : (timeInMode < 5 ? -51 : -25);       // POINTS_SUB_ARRAY[timeInMode+1]
            mode = 3; m = 'S';
        }
        if (wantMatch) match[k] = m;
    }
    scores[i] = score;
}

}  // namespace bbm

extern "C" int bbm_launch_noindel(const int8_t* reads, const int8_t* refs, const bbm_noindel_task* tasks, int* scores,
                                  int8_t* match_buf, const long long* match_off, long long n, cudaStream_t st) {
    bbm::noindel_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(reads, refs, tasks, scores, match_buf, match_off, n);
    return (int)cudaGetLastError();
}
