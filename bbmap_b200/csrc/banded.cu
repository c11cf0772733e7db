// banded.cu — BandedAligner on the GPU: one warp per (query, ref) pair, lanes across the band.
//
// Reproduces the reference's JNI C exactly (jni/BandedAlignerJNI.c:97-585; the C and BandedAlignerConcrete.java differ and
// the north star names the C): unit-cost edit distance inside a band of `width=min(maxWidth,2*maxEdits+1)` cells that slides
// one reference column per query row, early exit when every cell exceeds maxEdits, forced-diagonal last row / last column,
// off-centre penalty, and the five "last*" outputs.  The row recurrence  cur[m]=min(up+1, diag+sub, cur[m-1]+1)  has a
// left-to-right chain; since the chain cost is a constant +1 per step it is an ordinary prefix-min of (a[m]-m), done with
// warp shuffles — exact, no iteration.
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

constexpr int BIGV = 999;
constexpr unsigned FULLM = 0xffffffffu;

__device__ __forceinline__ int comp_ext(int c, const signed char* tab) { return (c >= 0 && c < 128) ? tab[c] : -1; }

struct BandParams {
    const int8_t* queries; const int8_t* refs; const bbm_band_task* tasks; bbm_band_out* outs; long long ntasks;
    unsigned int* counter;
    const int* list; const unsigned int* listCount;      // warp kernel: when set, work on tasks[list[0 .. *listCount)] (pairs too wide for the thread kernel)
    int* wideList; unsigned int* wideCount;              // thread kernel: where it leaves those pairs
};

// K = band cells per lane; lane L owns band indices m in [L*K+1, L*K+K] (the reference's array index, 1-based)
template <int K>
__device__ void banded_task(const BandParams& P, const bbm_band_task& T0, bbm_band_out* out, const signed char* comp) {
    const int lane = threadIdx.x & 31;
    // ---- the swap rules at the top of each variant (jni/BandedAlignerJNI.c:141-148, 260-267, 375-382, 490-497) ----
    int dir = T0.dir;
    const int8_t* query = P.queries + T0.query_off; const int8_t* ref = P.refs + T0.ref_off;
    int qlen = T0.query_len, rlen = T0.ref_len, qstart = T0.qstart, rstart = T0.rstart;
    bool swapped = false;
    {
        bool sw; int dir2 = dir;
        if (dir == 0) sw = (qlen - qstart > rlen - rstart);
        else if (dir == 1) { sw = (qstart + 1 > rlen - rstart); dir2 = 3; }
        else if (dir == 2) sw = (qstart > rstart);
        else { sw = (qlen - qstart > rstart + 1); dir2 = 1; }
        if (sw) {
            const int8_t* tp = query; query = ref; ref = tp;
            int t = qlen; qlen = rlen; rlen = t;
            t = qstart; qstart = rstart; rstart = t;
            dir = dir2; swapped = true;
            // the callee may swap back only if its own rule fires; the rules are mutually exclusive except on ties,
            // where the callee's rule is evaluated on the swapped arguments exactly like the recursive C call
            bool sw2;
            if (dir == 0) sw2 = (qlen - qstart > rlen - rstart);
            else if (dir == 1) sw2 = (qstart + 1 > rlen - rstart);
            else if (dir == 2) sw2 = (qstart > rstart);
            else sw2 = (qlen - qstart > rstart + 1);
            if (sw2) {   // cannot happen (strict inequalities are antisymmetric); keep the semantics anyway
                tp = query; query = ref; ref = tp; t = qlen; qlen = rlen; rlen = t; t = qstart; qstart = rstart; rstart = t;
                dir = (dir == 1) ? 3 : (dir == 3 ? 1 : dir); swapped = false;
            }
        }
    }
    const bool rc = (dir == 1 || dir == 3);
    const int qstep = (dir == 0 || dir == 3) ? 1 : -1;
    const bool rfwd = (dir == 0 || dir == 1);
    const int maxEdits = T0.max_edits, maxWidth = T0.max_width;
    const bool inexact = !T0.exact;
    const int width = imin(maxWidth, maxEdits * 2 + 1), halfWidth = width / 2;
    int qloc = qstart, rsloc = rstart - halfWidth;
    const int xlines = (dir == 0 || dir == 3) ? qlen - qstart : qstart + 1;
    const int ylines = rfwd ? rlen - rstart : rstart + 1;
    const int len = imin(xlines, ylines);
    int rv0 = 0, rv1 = 0, lastRow = -1, lastEdits = 0, lastOffset = 0, edits = 0;
    if (len >= 1 && width >= 1) {
        int cur[K], prev[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { cur[k] = BIGV; prev[k] = BIGV; }
        const int center = halfWidth + 1;
        // off-centre penalty (…JNI.c:111-121): arr[center±i]=min(big, arr[center±i]+i), returns the minimum over 1..2*halfWidth+1
        auto penalize = [&](int* arr) -> int {
            int mn = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int m = lane * K + k + 1;
                const int off = m > center ? m - center : center - m;
                if (off >= 1 && off <= halfWidth) arr[k] = imin(BIGV, arr[k] + off);
                if (off <= halfWidth) mn = imin(mn, arr[k]);
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) mn = imin(mn, __shfl_xor_sync(FULLM, mn, o));
            return mn;
        };
        int row = 0;
        for (row = 0; row < len; ++row) {
            if (row > 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) { prev[k] = cur[k]; cur[k] = BIGV; }
            }
            int q = query[qloc];
            if (rc) q = comp_ext(q, comp);
            const bool qdef = base_defined(q);
            const int colStart = imax(0, rsloc), colLimit = imin(rsloc + width, rlen);
            const int ncols = colLimit - colStart;
            const int mstart = rfwd ? 1 + (colStart - rsloc) : 1 + width - (colLimit - rsloc);
            const bool forceDiag = (row > 0 && row == len - 1);
            // neighbour values of the previous row: prev[m+1] lives in the next register / next lane
            const int nextFirst = __shfl_down_sync(FULLM, prev[0], 1);
            int a[K]; bool valid[K], noscan[K];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int m = lane * K + k + 1;
                const int idx = m - mstart;
                valid[k] = (idx >= 0 && idx < ncols);
                const int col = rfwd ? colStart + idx : colLimit - 1 - idx;
                int r = 'N';
                if (valid[k]) r = ref[col];
                const int sub = (q == r || (inexact && (!qdef || !base_defined(r)))) ? 0 : 1;
                int up = (k + 1 < K) ? prev[k + 1] : ((lane < 31) ? nextFirst : BIGV);
                if (m + 1 > maxWidth + 1) up = BIGV;        // beyond the reference's array (never read there)
                const int diag = prev[k] + sub;
                const bool edge = rfwd ? (col == rlen - 1) : (col == 0);
                noscan[k] = forceDiag || edge;
                a[k] = (row == 0) ? sub : (noscan[k] ? diag : imin(up + 1, diag));
            }
            int rowMin = BIGV;
            if (row == 0 || forceDiag) {
#pragma unroll
                for (int k = 0; k < K; ++k) { cur[k] = valid[k] ? a[k] : BIGV; if (valid[k]) rowMin = imin(rowMin, cur[k]); }
            } else {
                // cur[m] = min(a[m], cur[m-1]+1) with cur[mstart-1]=big  ==>  cur[m] = m + prefixmin_{j<=m}(a[j]-j), seeded with big-(mstart-1).
                // An edge cell takes its diagonal value and is the last valid cell of the row, so it never feeds the chain.
                int v[K];
                int run = 0x3fffffff;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int m = lane * K + k + 1;
                    v[k] = (valid[k] && !noscan[k]) ? a[k] - m : 0x3fffffff;
                    run = imin(run, v[k]);
                    v[k] = run;                               // lane-local inclusive prefix min
                }
                int incl = run;                              // inclusive prefix-min over lanes
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULLM, incl, o); if (lane >= o) incl = imin(incl, t); }
                int excl = __shfl_up_sync(FULLM, incl, 1);
                if (lane == 0) excl = 0x3fffffff;
                const int seed = BIGV - (mstart - 1);
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int m = lane * K + k + 1;
                    const int pm = imin(imin(v[k], excl), seed);
                    int val = BIGV;
                    if (valid[k]) val = noscan[k] ? a[k] : pm + m;
                    cur[k] = val;
                    if (valid[k]) rowMin = imin(rowMin, val);
                }
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) rowMin = imin(rowMin, __shfl_xor_sync(FULLM, rowMin, o));
            edits = rowMin;
            if (row == 0) edits = penalize(cur);
            else if (edits > maxEdits) { row++; break; }       // the for-increment (qloc, rsloc) is skipped on break
            qloc += qstep; rsloc += rfwd ? 1 : -1;
        }
        edits = penalize(cur);
        lastRow = row - 1; lastEdits = edits;
        // lastOffsetFunc (…JNI.c:97-109): first strict minimum in the order center, +1, -1, +2, -2, ...
        {
            int bestV = 0x7fffffff, bestRank = 0x7fffffff, bestM = center;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int m = lane * K + k + 1;
                const int off = m - center;
                const int aoff = off < 0 ? -off : off;
                if (aoff <= halfWidth) {
                    const int rank = off == 0 ? 0 : (off > 0 ? 2 * off - 1 : 2 * aoff);
                    if (cur[k] < bestV || (cur[k] == bestV && rank < bestRank)) { bestV = cur[k]; bestRank = rank; bestM = m; }
                }
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                const int ov = __shfl_xor_sync(FULLM, bestV, o), orank = __shfl_xor_sync(FULLM, bestRank, o), om = __shfl_xor_sync(FULLM, bestM, o);
                if (ov < bestV || (ov == bestV && orank < bestRank)) { bestV = ov; bestRank = orank; bestM = om; }
            }
            lastOffset = center - bestM;
        }
        if (dir == 0) { rv0 = qloc - 1; rv1 = rsloc + halfWidth - lastOffset - 1; while (rv1 >= rlen || rv0 >= qlen) { rv1--; rv0--; } }
        else if (dir == 1) { rv0 = qloc + 1; rv1 = rsloc + halfWidth - lastOffset - 1; while (rv1 >= rlen || rv0 < 0) { rv1--; rv0++; } }
        else if (dir == 2) { rv0 = qloc + 1; rv1 = rsloc + halfWidth + lastOffset + 1; while (rv1 < 0 || rv0 < 0) { rv1++; rv0++; } }
        else { rv0 = qloc - 1; rv1 = rsloc + halfWidth + lastOffset + 1; while (rv1 < 0 || rv0 >= qlen) { rv1++; rv0--; } }
    }
    if (lane == 0) {
        out->edits = edits;
        out->rv[0] = swapped ? rv1 : rv0; out->rv[1] = swapped ? rv0 : rv1;
        out->rv[2] = lastRow; out->rv[3] = lastEdits; out->rv[4] = lastOffset;
        out->status = 0; out->pad_ = 0;
    }
}

__global__ void __launch_bounds__(128) banded_kernel(BandParams P) {
    __shared__ signed char comp[128];
    {   // baseToComplementExtended (dna/AminoAcid.java:650-664)
        const char* ext = " ACMGRSVTWYHKDBNX"; const char* cex = " TGKCYWBASRDMHVNX";
        if (threadIdx.x < 128) comp[threadIdx.x] = -1;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < 17; ++i) {
                const int x = ext[i], y = cex[i];
                comp[x] = (signed char)y;
                const int xl = (x >= 'A' && x <= 'Z') ? x + 32 : x, yl = (y >= 'A' && y <= 'Z') ? y + 32 : y;
                comp[xl] = (signed char)yl;
            }
            comp['U'] = 'A'; comp['u'] = 'a'; comp['?'] = '?'; comp[' '] = ' '; comp['-'] = '-'; comp['*'] = '*'; comp['.'] = '.';
        }
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const long long nwork = P.list ? (long long)*P.listCount : P.ntasks;
    for (;;) {
        unsigned id = 0;
        if (lane == 0) id = atomicAdd(P.counter, 1u);
        id = __shfl_sync(FULLM, id, 0);
        if ((long long)id >= nwork) break;
        if (P.list) id = (unsigned)P.list[id];
        const bbm_band_task T = P.tasks[id];
        bbm_band_out* out = P.outs + id;
        const int width = imin(T.max_width, T.max_edits * 2 + 1);
        const bool bad = T.query_len < 0 || T.ref_len < 0 || T.dir < 0 || T.dir > 3 || T.max_width < 1;
        if (bad || width + 1 > 128) {
            if (lane == 0) { out->edits = 0; for (int k = 0; k < 5; ++k) out->rv[k] = 0; out->status = bad ? BBM_E_ARG : BBM_E_SHAPE; out->pad_ = 0; }
            continue;
        }
        if (width + 1 <= 32) banded_task<1>(P, T, out, comp);
        else if (width + 1 <= 64) banded_task<2>(P, T, out, comp);
        else banded_task<4>(P, T, out, comp);
        __syncwarp();
    }
}


// =====================  one THREAD per pair (narrow bands)  =====================
// Dedupe's bands are 3-9 cells wide (maxWidth = max(min(bw, 2*maxEdits+1), 3)|1 with bw = 9, jgi/Dedupe.java:423-426): a warp per pair keeps 9 of 32
// lanes busy and pays a shuffle scan per row.  Here every lane runs its own pair with the whole band in registers: b[1..width] is updated in place
// (the upper neighbour of cell m is the old b[m+1], the diagonal one the old b[m], the left one the value just written), the reference bytes under
// the band are a shift register that takes ONE new byte per row, and the left-to-right chain is the running minimum of (a[m] - m) exactly as in the
// scan formulation above.  Lanes fetch a new pair as soon as theirs is finished (pairs are 150-5000 rows long), so the row step always runs converged.
// W - 1 = widest band taken; wider pairs are left in `wideList` for the warp kernel.
template <int W>
__global__ void __launch_bounds__(128) banded_thread_kernel(BandParams P) {
    __shared__ signed char comp[128];
    {
        const char* ext = " ACMGRSVTWYHKDBNX"; const char* cex = " TGKCYWBASRDMHVNX";
        if (threadIdx.x < 128) comp[threadIdx.x] = -1;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int i = 0; i < 17; ++i) {
                const int x = ext[i], y = cex[i];
                comp[x] = (signed char)y;
                const int xl = (x >= 'A' && x <= 'Z') ? x + 32 : x, yl = (y >= 'A' && y <= 'Z') ? y + 32 : y;
                comp[xl] = (signed char)yl;
            }
            comp['U'] = 'A'; comp['u'] = 'a'; comp['?'] = '?'; comp[' '] = ' '; comp['-'] = '-'; comp['*'] = '*'; comp['.'] = '.';
        }
        __syncthreads();
    }
    bool have = false, done = false;
    // the pair
    bbm_band_out* out = nullptr;
    const int8_t* query = nullptr; const int8_t* ref = nullptr;
    int qlen = 0, rlen = 0, dir = 0, qstep = 1, maxEdits = 0, width = 0, halfWidth = 0, center = 1, len = 0;
    bool rc = false, rfwd = true, inexact = false, swapped = false;
    // the walk
    int row = 0, qloc = 0, rsloc = 0, edits = 0;
    int b[W + 1], r[W + 1];
#pragma unroll
    for (int m = 0; m <= W; ++m) { b[m] = BIGV; r[m] = 'N'; }

    auto penalize = [&]() -> int {                              // …JNI.c:111-121
        int mn = 0x7fffffff;
#pragma unroll
        for (int m = 1; m <= W; ++m) {
            const int off = m > center ? m - center : center - m;
            if (off >= 1 && off <= halfWidth) b[m] = imin(BIGV, b[m] + off);
            if (off <= halfWidth) mn = imin(mn, b[m]);
        }
        return mn;
    };

    for (;;) {
        if (!have && !done) {
            const unsigned id = atomicAdd(P.counter, 1u);
            if ((long long)id >= P.ntasks) done = true;
            else {
                const bbm_band_task T0 = P.tasks[id];
                out = P.outs + id;
                const int w0 = imin(T0.max_width, T0.max_edits * 2 + 1);
                const bool bad = T0.query_len < 0 || T0.ref_len < 0 || T0.dir < 0 || T0.dir > 3 || T0.max_width < 1;
                if (bad || w0 + 1 > 128) {
                    out->edits = 0; for (int k = 0; k < 5; ++k) out->rv[k] = 0; out->status = bad ? BBM_E_ARG : BBM_E_SHAPE; out->pad_ = 0;
                } else if (w0 + 1 > W) {
                    P.wideList[atomicAdd(P.wideCount, 1u)] = (int)id;
                } else {
                    // ---- the swap rules at the top of each variant (jni/BandedAlignerJNI.c:141-148, 260-267, 375-382, 490-497) ----
                    dir = T0.dir;
                    query = P.queries + T0.query_off; ref = P.refs + T0.ref_off;
                    qlen = T0.query_len; rlen = T0.ref_len;
                    int qstart = T0.qstart, rstart = T0.rstart;
                    swapped = false;
                    bool sw; int dir2 = dir;
                    if (dir == 0) sw = (qlen - qstart > rlen - rstart);
                    else if (dir == 1) { sw = (qstart + 1 > rlen - rstart); dir2 = 3; }
                    else if (dir == 2) sw = (qstart > rstart);
                    else { sw = (qlen - qstart > rstart + 1); dir2 = 1; }
                    if (sw) {
                        const int8_t* tp = query; query = ref; ref = tp;
                        int t = qlen; qlen = rlen; rlen = t;
                        t = qstart; qstart = rstart; rstart = t;
                        dir = dir2; swapped = true;
                        bool sw2;
                        if (dir == 0) sw2 = (qlen - qstart > rlen - rstart);
                        else if (dir == 1) sw2 = (qstart + 1 > rlen - rstart);
                        else if (dir == 2) sw2 = (qstart > rstart);
                        else sw2 = (qlen - qstart > rstart + 1);
                        if (sw2) {   // cannot happen (strict inequalities are antisymmetric); same semantics as the warp kernel
                            tp = query; query = ref; ref = tp; t = qlen; qlen = rlen; rlen = t; t = qstart; qstart = rstart; rstart = t;
                            dir = (dir == 1) ? 3 : (dir == 3 ? 1 : dir); swapped = false;
                        }
                    }
                    rc = (dir == 1 || dir == 3);
                    qstep = (dir == 0 || dir == 3) ? 1 : -1;
                    rfwd = (dir == 0 || dir == 1);
                    maxEdits = T0.max_edits; inexact = !T0.exact;
                    width = w0; halfWidth = width / 2; center = halfWidth + 1;
                    qloc = qstart; rsloc = rstart - halfWidth;
                    const int xlines = (dir == 0 || dir == 3) ? qlen - qstart : qstart + 1;
                    const int ylines = rfwd ? rlen - rstart : rstart + 1;
                    len = imin(xlines, ylines);
                    row = 0; edits = 0;
                    if (len >= 1 && width >= 1) {
#pragma unroll
                        for (int m = 1; m <= W; ++m) {
                            b[m] = BIGV;
                            const int col = rfwd ? rsloc + m - 1 : rsloc + width - m;
                            r[m] = (m <= width && col >= 0 && col < rlen) ? (int)ref[col] : (int)'N';
                        }
                        have = true;
                    } else {
                        out->edits = 0; out->rv[0] = 0; out->rv[1] = 0; out->rv[2] = -1; out->rv[3] = 0; out->rv[4] = 0; out->status = 0; out->pad_ = 0;
                    }
                }
            }
        }
        if (__all_sync(FULLM, done)) break;
        if (!have) continue;

        // ---- one row (converged across the warp's 32 pairs) ----
        bool finished = false;
        {
            int q = query[qloc];
            if (rc) q = comp_ext(q, comp);
            const bool qdef = base_defined(q);
            const int colStart = imax(0, rsloc), colLimit = imin(rsloc + width, rlen);
            const int ncols = colLimit - colStart;
            const int mstart = rfwd ? 1 + (colStart - rsloc) : 1 + width - (colLimit - rsloc);
            const bool first = (row == 0);
            const bool forceDiag = (row > 0 && row == len - 1);
            const int edgeCol = rfwd ? rlen - 1 : 0;
            int pmin = BIGV - (mstart - 1);                       // the chain enters the first valid cell as big + 1
            int rowMin = BIGV;
#pragma unroll
            for (int m = 1; m < W; ++m) {
                const int idx = m - mstart;
                const bool valid = (idx >= 0 && idx < ncols);
                const int col = rfwd ? colStart + idx : colLimit - 1 - idx;
                const int rb = r[m];
                const int sub = (q == rb || (inexact && (!qdef || !base_defined(rb)))) ? 0 : 1;
                const int up = b[m + 1], diag = b[m] + sub;
                const bool noscan = forceDiag || (col == edgeCol);
                const int a = first ? sub : (noscan ? diag : imin(up + 1, diag));
                int val = a;
                if (!first && !forceDiag && !noscan) {
                    if (valid) pmin = imin(pmin, a - m);
                    val = pmin + m;
                }
                val = valid ? val : BIGV;
                b[m] = val;
                rowMin = imin(rowMin, val);
            }
            edits = rowMin;
            if (first) edits = penalize();
            else if (edits > maxEdits) { row++; finished = true; }       // qloc / rsloc are not advanced on this exit
            if (!finished) {
                qloc += qstep; rsloc += rfwd ? 1 : -1;
                row++;
                if (row >= len) finished = true;
                else {
                    // the band slides one reference column: every cell takes its right neighbour's byte, the last one a new byte
                    const int col = rfwd ? rsloc + width - 1 : rsloc;
                    const int nb = (col >= 0 && col < rlen) ? (int)ref[col] : (int)'N';
#pragma unroll
                    for (int m = 1; m < W; ++m) r[m] = (m == width) ? nb : r[m + 1];
                }
            }
        }
        if (!finished) continue;

        // ---- the pair is done (…JNI.c:97-109 lastOffsetFunc, the return values of each variant) ----
        {
            edits = penalize();
            const int lastRow = row - 1, lastEdits = edits;
            int bestV = 0x7fffffff, bestRank = 0x7fffffff, bestM = center;
#pragma unroll
            for (int m = 1; m <= W; ++m) {
                const int off = m - center;
                const int aoff = off < 0 ? -off : off;
                if (aoff <= halfWidth) {
                    const int rank = off == 0 ? 0 : (off > 0 ? 2 * off - 1 : 2 * aoff);
                    if (b[m] < bestV || (b[m] == bestV && rank < bestRank)) { bestV = b[m]; bestRank = rank; bestM = m; }
                }
            }
            const int lastOffset = center - bestM;
            int rv0, rv1;
            if (dir == 0) { rv0 = qloc - 1; rv1 = rsloc + halfWidth - lastOffset - 1; while (rv1 >= rlen || rv0 >= qlen) { rv1--; rv0--; } }
            else if (dir == 1) { rv0 = qloc + 1; rv1 = rsloc + halfWidth - lastOffset - 1; while (rv1 >= rlen || rv0 < 0) { rv1--; rv0++; } }
            else if (dir == 2) { rv0 = qloc + 1; rv1 = rsloc + halfWidth + lastOffset + 1; while (rv1 < 0 || rv0 < 0) { rv1++; rv0++; } }
            else { rv0 = qloc - 1; rv1 = rsloc + halfWidth + lastOffset + 1; while (rv1 < 0 || rv0 >= qlen) { rv1++; rv0--; } }
            out->edits = edits;
            out->rv[0] = swapped ? rv1 : rv0; out->rv[1] = swapped ? rv0 : rv1;
            out->rv[2] = lastRow; out->rv[3] = lastEdits; out->rv[4] = lastOffset;
            out->status = 0; out->pad_ = 0;
            have = false;
        }
    }
}

// largest band width of a batch (picks the thread kernel's instantiation)
__global__ void banded_maxwidth_kernel(const bbm_band_task* tasks, long long n, unsigned int* outMax) {
    unsigned int mx = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const bbm_band_task T = tasks[i];
        if (T.max_width >= 1 && T.max_edits >= 0) mx = max(mx, (unsigned)imin(T.max_width, T.max_edits * 2 + 1));
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) mx = max(mx, __shfl_xor_sync(FULLM, mx, o));
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(outMax, mx);
}

}  // namespace bbm

using namespace bbm;
extern "C" int bbm_launch_banded(const int8_t* q, const int8_t* r, const bbm_band_task* t, bbm_band_out* o, long long n,
                                 unsigned int* counter, int blocks, cudaStream_t st, const int* list, const unsigned int* listCount) {
    BandParams P; P.queries = q; P.refs = r; P.tasks = t; P.outs = o; P.ntasks = n; P.counter = counter;
    P.list = list; P.listCount = listCount; P.wideList = nullptr; P.wideCount = nullptr;
    banded_kernel<<<blocks, 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_banded_maxwidth(const bbm_band_task* t, long long n, unsigned int* outMax, int blocks, cudaStream_t st) {
    banded_maxwidth_kernel<<<blocks, 256, 0, st>>>(t, n, outMax);
    return (int)cudaGetLastError();
}
// thread-per-pair kernel for bands of up to `maxBand` (<= 15) cells; pairs with wider bands are appended to wideList
extern "C" int bbm_launch_banded_thread(const int8_t* q, const int8_t* r, const bbm_band_task* t, bbm_band_out* o, long long n,
                                        unsigned int* counter, int* wideList, unsigned int* wideCount, int maxBand, int blocks, cudaStream_t st) {
    BandParams P; P.queries = q; P.refs = r; P.tasks = t; P.outs = o; P.ntasks = n; P.counter = counter;
    P.list = nullptr; P.listCount = nullptr; P.wideList = wideList; P.wideCount = wideCount;
    if (maxBand <= 9) banded_thread_kernel<10><<<blocks, 128, 0, st>>>(P);
    else banded_thread_kernel<16><<<blocks, 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}
