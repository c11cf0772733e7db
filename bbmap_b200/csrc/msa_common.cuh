// msa_common.cuh — constants of the MultiStateAligner11ts scoring scheme and small device helpers.
// Values: reference current/align2/MultiStateAligner11tsJNI.java:1489-1563 (== jni/MultiStateAligner11tsJNI.c:18-98).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "../../include/bbmap_cuda.h"

namespace bbm {

constexpr int TBITS = 11;
constexpr int TMASK = (1 << TBITS) - 1;            // TIMEMASK
constexpr int SMASK = (int)0xFFFFF800;             // SCOREMASK
constexpr int MAX_TIME = TMASK;
constexpr int TIME_WRAP = MAX_TIME - 3;            // MAX_TIME-MASK5
constexpr int MAX_SCORE = ((1 << 20) - 1) - 2000;
constexpr int MIN_SCORE = -MAX_SCORE;
constexpr int BAD = MIN_SCORE - 1;

__host__ __device__ constexpr int OFF(int pts) { return pts * 2048; }

constexpr int P_MATCH = OFF(70), P_MATCH2 = OFF(100);
constexpr int P_SUB = OFF(-127), P_SUBR = OFF(-147), P_SUB2 = OFF(-51), P_SUB3 = OFF(-25);
constexpr int P_INS = OFF(-395), P_INS2 = OFF(-39), P_INS3 = OFF(-23), P_INS4 = OFF(-8);
constexpr int P_DEL = OFF(-472), P_DEL2 = OFF(-33), P_DEL3 = OFF(-9), P_DEL4 = OFF(-1), P_DEL5 = OFF(-1);
constexpr int P_DEL_REF_N = OFF(-10), P_GAP = OFF(-2);
constexpr int BADOFF = OFF(BAD);
constexpr int MINOFF_SCORE = OFF(MIN_SCORE);
constexpr int LIM3 = 5, LIM4 = 20, LIM5 = 80;
constexpr int MIN_SCORE_ADJUST = 120;              // current/align2/MSA.java:868

constexpr int ST_MS = 0, ST_DEL = 1, ST_INS = 2;

// Largest read the register-tiled kernels take (reference ALIGN_ROWS=601, BBMapThread.java:28)
constexpr int MAXR = 608;
constexpr int TAB_MAX_COLS = 768;              // widest window of the kernels that read the penalty tables (strip / band kernels; the tiled ones stop at 512)
constexpr int PEN_TAB = TAB_MAX_COLS + 8;      // streak-indexed tables: a DEL run is at most `columns` long, INS / SUB runs at most `rows` (<= MAXR)
constexpr int DELC_TAB = MAXR + TAB_MAX_COLS + 32;   // DEL run length (<= columns) + delNeeded (<= rows)

__device__ __forceinline__ int imax(int a, int b) { return max(a, b); }
__device__ __forceinline__ int imin(int a, int b) { return min(a, b); }
__device__ __forceinline__ int imax3(int a, int b, int c) { return __vimax3_s32(a, b, c); }   // DPX VIMNMX3

// baseToNumber[c]>=0  (dna/AminoAcid.java:615-624): ACGTU in either case
__device__ __forceinline__ bool base_defined(int c) {
    const int u = c & 0xDF;   // fold case for letters
    return (c >= 'A') && (c <= 'u') && (u == 'A' || u == 'C' || u == 'G' || u == 'T' || u == 'U') && ((c & 0x80) == 0);
}

// calcDelScoreOffset (jni/MultiStateAligner11tsJNI.c:316-336)
__host__ __device__ inline int del_score_offset(int len) {
    if (len <= 0) return 0;
    int s = P_DEL;
    if (len > LIM5) { s += ((len - LIM5 + 3) / 4) * P_DEL5; len = LIM5; }
    if (len > LIM4) { s += (len - LIM4) * P_DEL4; len = LIM4; }
    if (len > LIM3) { s += (len - LIM3) * P_DEL3; len = LIM3; }
    if (len > 1) s += (len - 1) * P_DEL2;
    return s;
}
// POINTSoff_INS_ARRAY_C[len] (MultiStateAligner11tsJNI.java:1583-1603); len>=0
__host__ __device__ inline int ins_score_offset(int len) {
    if (len <= 0) return 0;
    int s = P_INS;
    if (len > LIM4) { s += (len - LIM4) * P_INS4; len = LIM4; }
    if (len > LIM3) { s += (len - LIM3) * P_INS3; len = LIM3; }
    if (len > 1) s += (len - 1) * P_INS2;
    return s < MINOFF_SCORE ? MINOFF_SCORE : s;
}

struct MsaParams {
    const int8_t* reads;
    const int8_t* refs;
    const bbm_msa_task* tasks;
    bbm_msa_out* outs;
    long long ntasks;
    int8_t* match_buf;
    const long long* match_off;
    int bandwidth;
    float ratio;
    unsigned long long* scratch;     // per-warp traceback scratch
    long long scratch_words;         // 64-bit words per warp
    unsigned int* counter;           // dynamic task counter
    int* overflow_list;              // tasks the tiled kernel could not take (shape), for the generic kernel
    unsigned int* overflow_count;
    int* dump;                       // optional packed dump (single-task latency path)
};

}  // namespace bbm
