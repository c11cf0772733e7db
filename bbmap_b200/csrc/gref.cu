// gref.cu — gapped references on the device.
//
// Reference: MSA.fillAndScoreLimited with a gap array (current/align2/MSA.java:103-134) ->
//   MultiStateAligner11tsJNI.fillLimited(…, gaps) (…JNI.java:116-130) -> makeGref (:668-757): the window
//   [min(gaps[0],a), max(gaps[n-1],b)] is copied exon by exon; every intron of `gap` bases is replaced by
//   GAPBUFFER+gap%GAPLEN bases, (gap-GAPBUFFER2)/GAPLEN '-' symbols and GAPBUFFER bases, then GREFLIMIT2_CUSHION
//   bases follow;  the fill runs on gref[0..greflimit] and score() translates bestRefStart/bestRefStop back with
//   translateFromGappedCoordinate (:759-779, :499-535).
//
// gref_build_kernel: one warp per task writes the gapped reference into a pool slot and rewrites the task so the
// ordinary MSA kernels (which already treat '-' columns exactly) run on it; gref_translate_kernel maps the two
// coordinates of score2 back.  Tasks without gaps pass through untouched.
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

constexpr int GAPBUFFER = 64, GAPBUFFER2 = 128, GAPLEN = 128, GREF_CUSHION = 128;   // Shared.java:20-24, …JNI.java:1565-1567

__global__ void __launch_bounds__(128) gref_build_kernel(const int8_t* __restrict__ refs, const bbm_gapped_task* __restrict__ gt,
                                                         const int* __restrict__ gapsAll, long long n, int8_t* pool, long long poolBase,
                                                         int stride, int greflen, bbm_gref_info* info, bbm_msa_task* tasksOut) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n) return;
    const bbm_gapped_task g = gt[w];
    bbm_msa_task t = g.t;
    bbm_gref_info I; I.origin = 0; I.greflimit = 0; I.greflimit2 = 0; I.status = 0;
    if (g.ngaps <= 0) { if (lane == 0) { tasksOut[w] = t; info[w] = I; } return; }
    const int8_t* ref = refs + t.ref_off;
    const int* gaps = gapsAll + g.gaps_off;
    const int ng = g.ngaps;
    const int a = imax(0, t.ref_start), b = imin(t.ref_len - 1, t.ref_end);          // MSA.java:104-105
    int8_t* gref = pool + w * (long long)stride;
    const int first = imin(gaps[0], a), last = imax(gaps[ng - 1], b);                // …JNI.java:676-679
    I.origin = first;
    int gpos = 0; bool over = ((ng & 1) != 0) || b < a;
    for (int i = 0; i < ng && !over; i += 2) {
        const int x = (i == 0 ? first : gaps[i]), y = (i + 1 == ng - 1 ? last : gaps[i + 1]);
        int len = y - x + 1;
        if (len < 0 || gpos + len > greflen) { over = true; break; }
        for (int j = lane; j < len; j += 32) { const int r = x + j; gref[gpos + j] = (r >= 0 && r < t.ref_len) ? ref[r] : (int8_t)'N'; }
        gpos += len;
        if (i + 2 < ng) {
            const int z = gaps[i + 2], gap = z - y - 1;
            if (gap < GAPBUFFER2) { over = true; break; }                            // the reference asserts gap>=MINGAP (:716)
            const int rem = gap % GAPLEN, div = (gap - GAPBUFFER2) / GAPLEN;
            const int n1 = GAPBUFFER + rem;
            if (gpos + n1 + div + GAPBUFFER > greflen) { over = true; break; }
            for (int j = lane; j < n1; j += 32) { const int r = y + 1 + j; gref[gpos + j] = (r >= 0 && r < t.ref_len) ? ref[r] : (int8_t)'N'; }   // a gap array hanging over the array would make the reference throw; never read outside
            gpos += n1;
            for (int j = lane; j < div; j += 32) gref[gpos + j] = (int8_t)'-';
            gpos += div;
            for (int j = lane; j < GAPBUFFER; j += 32) { const int r = z - GAPBUFFER + j; gref[gpos + j] = (r >= 0 && r < t.ref_len) ? ref[r] : (int8_t)'N'; }
            gpos += GAPBUFFER;
        }
    }
    if (!over && gpos >= greflen) over = true;        // no room for the column fillLimitedX(gref, 0, greflimit) adds
    if (over) {
        I.status = BBM_E_SHAPE;
        t.read_len = 0;                                // classified as an invalid task; outs[].status = BBM_E_SHAPE
    } else {
        I.greflimit = gpos;
        const int lim = imin(greflen, gpos + GREF_CUSHION);
        for (int i = gpos + lane; i < lim; i += 32) { const int r = b + 1 + (i - gpos); gref[i] = r < t.ref_len ? ref[r] : (int8_t)'N'; }
        I.greflimit2 = lim - 1;
        t.ref_off = poolBase + w * (long long)stride;
        t.ref_len = greflen;
        t.ref_start = 0; t.ref_end = gpos;            // …JNI.java:127: fillLimitedX(read, gref, 0, greflimit, minScore)
        t.flags = (t.flags & ~BBM_TF_CLAMP) | BBM_TF_GAPPED;
    }
    if (lane == 0) { tasksOut[w] = t; info[w] = I; }
}

// translateFromGappedCoordinate (…JNI.java:759-779): j advances by GAPLEN over every '-' left of `point`.
__device__ int from_gapped(const int8_t* gref, const bbm_gref_info& I, int point) {
    if (point <= 0) return I.origin + point;
    if (point >= I.greflimit2) return (-0x7fffffff - 1);     // the reference throws "Out of bounds."
    int j = I.origin;
    for (int i = 0; i < point; ++i) j += (gref[i] == '-' ? GAPLEN : 1);
    return j;
}

__global__ void __launch_bounds__(128) gref_translate_kernel(const bbm_gapped_task* __restrict__ gt, long long n, const int8_t* __restrict__ pool,
                                                             int stride, const bbm_gref_info* __restrict__ info, bbm_msa_out* outs) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || gt[i].ngaps <= 0) return;
    const bbm_gref_info I = info[i];
    if (I.status) { outs[i].status = I.status; return; }
    if (outs[i].score_len > 0) {
        const int8_t* gref = pool + i * (long long)stride;
        outs[i].score[1] = from_gapped(gref, I, outs[i].score[1]);
        outs[i].score[2] = from_gapped(gref, I, outs[i].score[2]);
    }
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_launch_gref_build(const int8_t* refs, const bbm_gapped_task* gt, const int* gaps, long long n, int8_t* pool, int stride,
                                     int greflen, bbm_gref_info* info, bbm_msa_task* tasksOut, cudaStream_t st) {
    const long long threads = n * 32;
    gref_build_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(refs, gt, gaps, n, pool, (long long)(pool - refs), stride, greflen, info, tasksOut);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_gref_translate(const bbm_gapped_task* gt, long long n, const int8_t* pool, int stride, const bbm_gref_info* info,
                                         bbm_msa_out* outs, cudaStream_t st) {
    gref_translate_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(gt, n, pool, stride, info, outs);
    return (int)cudaGetLastError();
}
