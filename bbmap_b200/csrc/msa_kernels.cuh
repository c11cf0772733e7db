// msa_kernels.cuh — kernel templates shared by the per-width translation units (msa_w*.cu) and msa_kernels.cu.
#pragma once
#include <climits>
#include "msa_tiled.cuh"

namespace bbm {

// MultiStateAligner11tsJNI.fillLimitedX's halfband (…JNI.java:137-138 == jni/...JNI.c:392-393)
__device__ __forceinline__ int calc_halfband(int bandwidth, float ratio, int rows, int cols) {
    if (bandwidth < 1 && ratio <= 0.f) return 0;
    const int x = bandwidth < 1 ? 9999999 : bandwidth;
    const int y = ratio <= 0.f ? 9999999 : 8 + (int)(rows * ratio);   // float multiply then truncation, like the C
    return imax(imin(x, y), cols - rows + 8) / 2;
}

// Resolve what the reference would run for this task (…JNI.java:132-144; MSA.java:104-105)
__device__ __forceinline__ bool resolve_task(const bbm_msa_task& t, int bandwidth, float ratio, TaskCtx& T) {
    int a = t.ref_start, b = t.ref_end;
    if (t.flags & BBM_TF_CLAMP) { a = imax(0, a); b = imin(t.ref_len - 1, b); }
    T.a = a; T.b = b; T.rows = t.read_len; T.cols = b - a + 1; T.flags = t.flags;
    T.minScore = t.min_score;
    T.limited = 0; T.halfband = 0;
    if (T.rows < 1 || T.cols < 1 || a < 0 || b >= t.ref_len) return false;
    T.halfband = calc_halfband(bandwidth, ratio, T.rows, T.cols);
    if (t.flags & BBM_TF_RAW_UNLIMITED) { T.limited = 0; }
    else if (t.flags & BBM_TF_RAW_LIMITED) { T.limited = 1; }
    else {
        const int hb = T.halfband;
        const bool unl = (T.minScore < 1) || (T.cols + T.rows < 90) ||
                         ((hb < 1 || hb * 3 > T.cols) && (T.cols > T.rows + imin(170, T.rows + 20)));
        T.limited = unl ? 0 : 1;
        if (!unl) T.minScore -= MIN_SCORE_ADJUST;
    }
    if (!T.limited) T.halfband = 0;
    return true;
}

// width classes of the register-tiled kernel: columns per lane
constexpr int NUM_WCLASS = 7;
__host__ __device__ constexpr int wclass_width(int k) { return k == 0 ? 4 : k == 1 ? 5 : k == 2 ? 6 : k == 3 ? 8 : k == 4 ? 9 : k == 5 ? 12 : 16; }
constexpr int CLASS_GENERIC = NUM_WCLASS;      // row-sequential kernel
constexpr int CLASS_BAD = NUM_WCLASS + 1;      // invalid task
constexpr int CLASS_STRIP = NUM_WCLASS + 2;    // limited, un-banded fills: thread-per-alignment strip kernel (msa_strip.cu)
constexpr int CLASS_BAND = NUM_WCLASS + 3;     // limited, banded fills: thread-per-alignment band kernel (msa_band.cu)
constexpr int NUM_CLASS = NUM_WCLASS + 4;
constexpr int STRIP_MAX_COLS = TAB_MAX_COLS;   // CellTables cover DEL/INS runs up to this (msa_cell.cuh): 600-row reads (configs[4]'s pieces, `maxlen=500`) in padded windows
constexpr int CLS_NARROW_BIT = 0x80;           // class byte flag: first try the thread-per-alignment narrow kernel
constexpr int NARROW_BUCKETS_C = 40;
// counter block (32-bit words)
constexpr int CB_COUNTS = 0, CB_CURSORS = 16, CB_WORK = 32, CB_OVERFLOW = 48, CB_NARROW_WORK = 49;
constexpr int CB_BAND_MAXROWS = 53, CB_BAND_MAXCOLS = 54, CB_BAND_MAXHB = 55, CB_BAND_WORK = 56;      // largest shape / halfband in the band class, its work counters (56, 57, 58: one per slot class)
constexpr int BAND_ND_CLASSES = 3, BAND_BUCKETS = BAND_ND_CLASSES * NARROW_BUCKETS_C;                  // band list order: slot class (32 / 64 / 128 slots), then read length, so that the 32 alignments of a warp have the same shape
constexpr int CB_BD_COUNTS = 256, CB_BD_CURSORS = 384, CB_WORDS_ALL = 512;
constexpr int CB_GENERIC_MAXCOLS = 51, CB_GENERIC_MAXROWS = 52;   // largest shape in the row-sequential class: sizes its shared-memory rows
constexpr int CB_NB_COUNTS = 64, CB_NB_CURSORS = 128, CB_WORDS = 192;
constexpr int STRIP_BUCKETS = 16;              // strip list is ordered by estimated work, largest first (longest-processing-time-first)
constexpr int CB_SB_COUNTS = 104, CB_SB_CURSORS = 168, CB_STRIP_BYTES = 184 /* 64-bit */, CB_STRIP_WORK = 50, CB_STRIP_POOL = 186 /* 64-bit */;
constexpr int NARROW_BUCKETS = 40;             // narrow list is ordered by read length (rows-1)/16 so warps are uniform
__host__ __device__ constexpr int narrow_bucket(int rows) { return (rows - 1) / 16 < NARROW_BUCKETS - 1 ? (rows - 1) / 16 : NARROW_BUCKETS - 1; }

__device__ __forceinline__ int classify(const TaskCtx& T) {
    const int w = (T.cols + 31) >> 5;
    if (T.rows > MAXR - 2 || w > 16) return CLASS_GENERIC;
    return w <= 4 ? 0 : w == 5 ? 1 : w == 6 ? 2 : w <= 8 ? 3 : w == 9 ? 4 : w <= 12 ? 5 : 6;
}

__device__ __forceinline__ bool strip_eligible(const TaskCtx& T) {
    return T.limited && T.halfband < 1 && T.rows >= 2 && T.rows <= MAXR - 2 && T.cols <= STRIP_MAX_COLS;
}

// banded limited fills whose band (3*halfband+3 slots) fits the band kernel's shared-memory row (msa_band.cu); CellTables bound the columns as for the strips
__device__ __forceinline__ bool band_eligible(const TaskCtx& T) {
    return T.limited && T.halfband >= 1 && 3 * T.halfband + 3 <= 128 && T.rows >= 1 && T.rows <= MAXR - 2 && T.cols <= STRIP_MAX_COLS;
}

__device__ __forceinline__ int band_nd_class(int hb) { const int nd = 3 * hb + 3; return nd <= 32 ? 0 : (nd <= 64 ? 1 : 2); }
// band list order within a slot class: estimated work (rows x cells a path may wander over given the slack between the best possible score and
// minScore, at most the band), largest first — the 32 alignments a warp starts together should take about equally long
__device__ __forceinline__ int band_bucket(const TaskCtx& T) {
    const int maxQ = (T.rows - 1) * 100 + 70;
    const int slack = imax(0, maxQ - T.minScore);
    const int w = (T.rows * imin(3 * T.halfband + 2, slack / 64 + 5)) >> 8;         // buckets of 256 cells
    return band_nd_class(T.halfband) * NARROW_BUCKETS_C + (NARROW_BUCKETS_C - 1 - imin(w, NARROW_BUCKETS_C - 1));
}

constexpr int SW = 8;                           // columns per strip (msa_strip.cu)
// estimated work of a limited fill: rows x (width of the band a path may wander given the slack between the best possible score
// and minScore); only used to order the strip list
__device__ __forceinline__ int strip_bucket(const TaskCtx& T) {
    const int maxQ = (T.rows - 1) * 100 + 70;
    const int slack = imax(0, maxQ - T.minScore);
    const long long w = (long long)T.rows * imin(T.cols, slack / 32 + 16);
    return (int)(w >> 12) < STRIP_BUCKETS - 1 ? (int)(w >> 12) : STRIP_BUCKETS - 1;      // buckets of 4096 cells
}
__host__ __device__ inline unsigned long long strip_task_bytes(int rows, int cols) {
    const unsigned long long ns = (unsigned long long)(cols + SW - 1) / SW, rs = (unsigned long long)rows + 2;
    return ((8 * rs + 4 * SW * ns + 4 * ns * rs) + 15ull) & ~15ull;
}

constexpr int WARPS_PER_BLOCK = 4;

template <int W, bool DUMP>
__global__ void __launch_bounds__(WARPS_PER_BLOCK * 32, (W <= 9 ? 4 : (W <= 12 ? 3 : 2))) msa_tiled_kernel(MsaParams P, const int* __restrict__ list, int nlist, const unsigned int* __restrict__ endPtr, unsigned int base, unsigned int* counter) {
    __shared__ BlockShared bs;
    __shared__ WarpShared wsAll[WARPS_PER_BLOCK];
    cell_tables_init(bs);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WarpShared& ws = wsAll[warp];
    const long long gwarp = (long long)blockIdx.x * WARPS_PER_BLOCK + warp;
    unsigned long long* scratch = P.scratch + gwarp * P.scratch_words;
    if (endPtr) nlist = (int)(*endPtr - base);        // list length decided on the device (narrow-kernel hand-overs included)
    for (;;) {
        unsigned k = 0;
        if (lane == 0) k = atomicAdd(counter, 1u);
        k = __shfl_sync(FULL, k, 0);
        if (k >= (unsigned)nlist) break;
        const int id = list ? list[k] : (int)k;
        const bbm_msa_task task = P.tasks[id];
        bbm_msa_out* out = P.outs + id;
        TaskCtx T;
        resolve_task(task, P.bandwidth, P.ratio, T);
        if (!T.limited) msa_fill_task<W, false, false, DUMP>(P, T, task, id, ws, bs, scratch, out);
        else if (T.halfband < 1) msa_fill_task<W, true, false, DUMP>(P, T, task, id, ws, bs, scratch, out);
        else msa_fill_task<W, true, true, DUMP>(P, T, task, id, ws, bs, scratch, out);
        __syncwarp();
    }
}

#define BBM_DECLARE_TILED_LAUNCH(W) \
    extern "C" int bbm_launch_msa_tiled_w##W(const bbm::MsaParams* P, const int* list, int nlist, const unsigned int* endPtr, unsigned int base, unsigned int* counter, int blocks, int dump, cudaStream_t stream);

#define BBM_DEFINE_TILED_LAUNCH(W) \
    extern "C" int bbm_launch_msa_tiled_w##W(const bbm::MsaParams* P, const int* list, int nlist, const unsigned int* endPtr, unsigned int base, unsigned int* counter, int blocks, int dump, cudaStream_t stream) { \
        if (dump) bbm::msa_tiled_kernel<W, true><<<1, bbm::WARPS_PER_BLOCK * 32, 0, stream>>>(*P, list, nlist, endPtr, base, counter); \
        else bbm::msa_tiled_kernel<W, false><<<blocks, bbm::WARPS_PER_BLOCK * 32, 0, stream>>>(*P, list, nlist, endPtr, base, counter); \
        return (int)cudaGetLastError(); \
    }

}  // namespace bbm
