// msa_strip.cu — thread-per-alignment MultiStateAligner11ts fill in column strips, for limited (pruned) fills of any width.
//
// fillLimitedX (jni/MultiStateAligner11tsJNI.c:361-704) visits, per read row, only the columns between the previous row's
// first and last "good" cell, so the useful work of one alignment is a ragged band around its path(s): 3-5 % of the rectangle
// for a read that matches its window, 50-80 % for a read with a long indel under a loose minScore.  The register-tiled kernel
// (msa_tiled.cuh) evaluates the whole rectangle with one warp per alignment; this kernel evaluates only what can matter:
//
//   * one THREAD per alignment, the matrix is swept in strips of SW columns (all three states of the previous row of the strip
//     live in registers, exactly like msa_narrow.cuh keeps diagonals);
//   * between strips only the last column of every row crosses (one 16-byte record per row, in place);
//   * a row of a strip is skipped when none of its predecessors can pass the reference's own limit tests
//     (jni/...JNI.c:478-486,566-570,619-623): every predecessor score <= max(vertLimit[row], min horizLimit of the strip) - MATCH2
//     makes the reference skip all three states of all SW cells, i.e. they hold `subfloor`; rows the previous strip never reached
//     are not even looked at;
//   * lanes of a warp work on different alignments and fetch a new one as soon as theirs is finished; every lane first advances
//     (cheaply, divergently) to its next row that needs evaluation, then all lanes evaluate one row of SW cells together, so the
//     expensive code always runs with a full warp whatever the shapes of the 32 alignments;
//   * vertLimit/horizLimit are prepared by a small kernel before, and the per-row control flow of the reference (iteration
//     counter, early break, final scan), score2 and traceback2 (MultiStateAligner11tsJNI.java:376-495, 537-658) run in a
//     small kernel after, both thread-per-alignment and convergent.
//
// Exactness (same argument as msa_tiled.cuh / msa_narrow.cuh): a stored state value is either a score that passed its limit
// ("good") or `subfloor`; a cell can only be good if one of its predecessors is good or lies in row 0 / column 0; cells the
// reference does not visit are read by it as `subfloor`.  Evaluating a cell the reference skips therefore yields subfloor, and
// not evaluating a cell whose predecessors are all below the limits is the reference's own skip.  Column 1 depends on column 0
// (real scores), so strip 0 applies the reference's colStart rule explicitly.
#include <climits>
#include "msa_kernels.cuh"

namespace bbm {

constexpr int STRIP_THREADS = 128;
#ifndef STRIP_BLOCKS_PER_SM
#define STRIP_BLOCKS_PER_SM 4
#endif

struct StripParams {
    MsaParams P;
    const int* list;                          // strip list segment (task ids), most expensive alignments first
    const unsigned int* endPtr; unsigned int base;   // segment length = *endPtr - base (narrow-kernel hand-overs included)
    int chunkStart, chunkCount;               // this launch handles list[chunkStart .. chunkStart+chunkCount)
    int rowStride;                            // maxRows + 2 (lane-private record array)
    long long* hdr;                           // [chunkCount]  byte offset of the task's block in `pool`, or -1
    int4* fin;                                // [chunkCount]  {bestScore, bestCol, bestState, bestPacked} of the last row
    int4* rec;                                // [threads][rowStride]  last column of the previous strip: {MS, DEL, INS, -}; lane-private, in place
    char* pool; unsigned long long poolBytes; unsigned long long* poolCursor;
    // per-task block in the pool:  int2 A[rows+2] {vertLimit[row], (minGood<<16|maxGood) of the row};  int hl[ns*SW] (hl[c-1] = horizLimit[c]);
    //                              unsigned tb[ns][rows+2]  4-bit predecessor codes, SW cells per word
    unsigned int* counter;
    int debug;                                // bit 0: never skip a row, bit 1: never jump over rows (A/B and bisecting), bit 2: count work
    unsigned long long* stats;                // [2] with debug bit 2: row-units evaluated, 32 x warp iterations of the evaluation phase
};

struct StripBlock { int2* A; int* hl; unsigned int* tb; int rs; };
__device__ __forceinline__ StripBlock strip_block(const StripParams& S, long long off, int rows, int cols) {
    StripBlock b; const int ns = (cols + SW - 1) / SW; b.rs = rows + 2;
    char* p = S.pool + off;
    b.A = (int2*)p; p += 8ll * b.rs;
    b.hl = (int*)p; p += 4ll * SW * ns;
    b.tb = (unsigned int*)p;
    return b;
}

__device__ __forceinline__ int strip_count(const StripParams& S) {
    const long long total = (long long)*S.endPtr - S.base - S.chunkStart;
    return (int)(total < 0 ? 0 : (total < S.chunkCount ? total : S.chunkCount));
}

// ---------------- K0: vertLimit / horizLimit (jni/...JNI.c:405-438) ----------------
__global__ void __launch_bounds__(128) msa_strip_prep_kernel(StripParams S) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= strip_count(S)) return;
    const int id = S.list[S.chunkStart + slot];
    const bbm_msa_task task = S.P.tasks[id];
    TaskCtx T;
    resolve_task(task, S.P.bandwidth, S.P.ratio, T);
    const int rows = T.rows, cols = T.cols;
    const int8_t* __restrict__ read = S.P.reads + task.read_off;
    const int8_t* __restrict__ ref = S.P.refs + task.ref_off + T.a;
    const int maxGain = (rows - 1) * P_MATCH2 + P_MATCH;
    const int minScore_off = (int)((unsigned)T.minScore << TBITS);
    const int floor_ = minScore_off - maxGain;
    const unsigned long long need = strip_task_bytes(rows, cols);
    const unsigned long long off = atomicAdd(S.poolCursor, need);
    S.fin[slot] = make_int4(INT_MIN, -1, -1, 0);
    if (off + need > S.poolBytes || rows + 2 > S.rowStride) {       // cannot happen when the host sized the pool from the classifier's total
        S.hdr[slot] = -1;
        bbm_msa_out o = {}; o.status = BBM_E_CAPACITY; o.match_len = -1; S.P.outs[id] = o;
        return;
    }
    S.hdr[slot] = (long long)off;
    const StripBlock B = strip_block(S, (long long)off, rows, cols);
    int2* A = B.A;
    int x = minScore_off;
    A[rows] = make_int2(x, MM_NONE);
    bool pd = false;                                   // defined(read[i+1]) && i+1 < rows
    for (int i = rows - 1; i >= 1; --i) {
        const bool d = base_defined(read[i]);
        x = imax(x - (d ? (pd ? P_MATCH2 : P_MATCH) : 0), floor_);
        A[i] = make_int2(x, MM_NONE);
        pd = d;
    }
    int* hl = B.hl;
    const int ncolPad = ((cols + SW - 1) / SW) * SW;
    for (int c = cols + 1; c <= ncolPad; ++c) hl[c - 1] = 0;
    x = minScore_off;
    hl[cols - 1] = x;
    pd = false;                                        // defined(ref[i+1]) && i+1 <= cols-1
    for (int i = cols - 1; i >= 1; --i) {
        bool d; const int cost = hcost(ref[i], pd, d);
        x = imax(x - cost, floor_);
        hl[i - 1] = x;
        pd = d;
    }
}


// ---------------- K1: the fill ----------------
template <int W>
__global__ void __launch_bounds__(STRIP_THREADS, STRIP_BLOCKS_PER_SM) msa_strip_fill_kernel(StripParams S) {
    __shared__ CellTables tab;
    cell_tables_init(tab);
    __syncthreads();
    const int n = strip_count(S);
    int4* const rec = S.rec + ((long long)blockIdx.x * STRIP_THREADS + threadIdx.x) * S.rowStride;     // lane-private
    int2* A = nullptr; const int* hlT = nullptr; unsigned int* tbT = nullptr; int rs = 0;           // this task's block in the pool
    bool have = false, done = false, newStrip = false;
    // task
    int slot = 0, rows = 0, cols = 0, nstrips = 0, minScore_off = 0;
    const int8_t* read = nullptr; const int8_t* ref = nullptr;
    CellConst K; K.floor_ = 0; K.subfloor = 0;
    // strip
    int s = 0, c0 = 1, r = 1, refLeft = '!', hlMin = 0;
    int rf[W], hlr[W], MS[W], DL[W], IN[W];
#pragma unroll
    for (int j = 0; j < W; ++j) { rf[j] = 0x200; hlr[j] = 0; MS[j] = 0; DL[j] = 0; IN[j] = 0; }
    unsigned gPrev = 0;
    int dM = 0, dD = 0, dI = 0;               // left neighbour column at row r-1
    int lM = 0, lD = 0, lI = 0, vlim = 0;     // left neighbour column at row r; vertLimit[r]
    int mmRow = MM_NONE, call1 = 0, call0 = 0; // (minGood,maxGood) of row r so far; read[r-1], read[r-2]
    int loP = 1, hiP = 0, loC = INT_MAX, hiC = -1;
    int bestScore = INT_MIN, bestCol = -1, bestState = -1, bestPacked = 0;
    unsigned long long nUnits = 0, nIters = 0;

    for (;;) {
        // ---- phase A: every lane advances to its next row that needs evaluation (cheap, divergent) ----
        bool pending = false;
        while (!pending && !done) {
            if (!have) {
                const unsigned k = atomicAdd(S.counter, 1u);
                if (k >= (unsigned)n) { done = true; break; }
                slot = (int)k;
                const long long off = S.hdr[slot];
                if (off < 0) continue;                 // no room in the pool (status already set by the prep kernel)
                const int id = S.list[S.chunkStart + slot];
                const bbm_msa_task task = S.P.tasks[id];
                TaskCtx T;
                resolve_task(task, S.P.bandwidth, S.P.ratio, T);
                rows = T.rows; cols = T.cols; nstrips = (cols + W - 1) / W;
                { const StripBlock B = strip_block(S, off, rows, cols); A = B.A; hlT = B.hl; tbT = B.tb; rs = B.rs; }
                read = S.P.reads + task.read_off; ref = S.P.refs + task.ref_off + T.a;
                minScore_off = (int)((unsigned)T.minScore << TBITS);
                K.floor_ = minScore_off - ((rows - 1) * P_MATCH2 + P_MATCH);
                K.subfloor = K.floor_ - 5 * P_MATCH2;
                bestScore = INT_MIN; bestCol = -1; bestState = -1; bestPacked = 0;
                s = -1; loC = 1; hiC = rows;           // "strip -1" is column 0: real scores in every row
                have = true; newStrip = true;
            }
            if (newStrip) {
                ++s;
                if (s == nstrips) { S.fin[slot] = make_int4(bestScore, bestCol, bestState, bestPacked); have = false; continue; }
                loP = loC; hiP = hiC; loC = INT_MAX; hiC = -1;
                c0 = s * W + 1;
                const int* hl = hlT + (c0 - 1);
                hlMin = INT_MAX;
#pragma unroll
                for (int j = 0; j < W; ++j) {
                    const int c = c0 + j;
                    int v = 0x200, h = 0;
                    if (c <= cols) { v = ref[c - 1]; if (v == 'N') v = 0x100; h = hl[j]; hlMin = imin(hlMin, h); }
                    rf[j] = v; hlr[j] = h;
                    MS[j] = 0; DL[j] = 0; IN[j] = 0;                 // row 0 of the matrix is all zero
                }
                refLeft = (c0 >= 2) ? (ref[c0 - 2] == 'N' ? 0x100 : (int)ref[c0 - 2]) : '!';
                gPrev = 0xffffffffu;
                dM = 0; dD = 0; dI = 0;                              // (0, c0-1)
                r = 1;
                newStrip = false;
            }
            // ---- examine row r ----
            if (s == 0) { const int v = tab.insc[r]; lM = v; lD = v; lI = v; }           // column 0 (…JNI.java:105-111)
            else if (r >= loP && r <= hiP) { const int4 q = rec[r]; lM = q.x; lD = q.y; lI = q.z; }
            else { lM = K.subfloor; lD = K.subfloor; lI = K.subfloor; }
            { const int2 a = A[r]; vlim = a.x; mmRow = a.y; }
            call1 = read[r - 1]; call0 = r < 2 ? '?' : read[r - 2];        // issued with the loads above: one memory round trip per row
            const int Lmin = imax(imax(vlim, hlMin) - P_MATCH2, K.floor_);
            const int prevTop = (r == 1) ? 0 : (gPrev ? INT_MAX : K.subfloor);
            const int bmax = imax(imax3(dM & SMASK, dD & SMASK, dI & SMASK), imax(lM & SMASK, lD & SMASK));
            if (imax(prevTop, bmax) > Lmin || (S.debug & 1)) { pending = true; break; }
            // the reference skips every state of every cell of this row of the strip: all subfloor
            if (gPrev) {
#pragma unroll
                for (int j = 0; j < W; ++j) { MS[j] = K.subfloor; DL[j] = K.subfloor; IN[j] = K.subfloor; }
            }
            gPrev = 0;
            if (loC != INT_MAX && s + 1 < nstrips) rec[r] = make_int4(K.subfloor, K.subfloor, K.subfloor, 0);
            dM = lM; dD = lD; dI = lI;
            if (s == 0 && !(S.debug & 2)) { newStrip = true; }          // vertLimit only grows and column 0 only falls: every later row is skipped too
            else {
                const int r2 = r + 1;
                if (r2 > rows) newStrip = true;
                else if (r2 >= loP && r2 <= hiP + 1) r = r2;        // row hiP+1 still sees (hiP, c0-1) on its diagonal
                else if (r2 < loP && loP <= hiP && !(S.debug & 2)) {
                    if (loC != INT_MAX && s + 1 < nstrips) for (int q = r2; q < loP; ++q) rec[q] = make_int4(K.subfloor, K.subfloor, K.subfloor, 0);
                    r = loP; dM = K.subfloor; dD = K.subfloor; dI = K.subfloor;
                }
                else if (S.debug & 2) r = r2;
                else newStrip = true;
            }
        }
        if (__all_sync(FULL, done)) break;
        nIters++;
        if (!pending) continue;
        nUnits++;          // (done lanes idle here until the warp has drained)

        // ---- phase B: one row of W cells (convergent) ----
        {
            CellRow R;
            R.call1 = call1;
            R.call0 = call0;
            R.callN = (R.call1 == 'N');
            R.vlimit = vlim;
            R.delBar = (r < 3) || (r > rows - 3);
            const bool insTop = (r < 2), insBot = (r > rows - 2);
            const bool allVis = (s > 0) || (r == 1);
            const int dn0 = r - c0 - 1, in0 = (rows - r) - (cols - c0) - 1;
            unsigned gCur = 0, word = 0;
            int ref0 = refLeft;
            int xM = dM, xD = dD, xI = dI, yM = lM, yD = lD;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int c = c0 + j;
                const bool inRange = (c <= cols);
                const bool visit = inRange && (allVis || ((gPrev & ((2u << j) - 1u)) != 0u));
                const int delNeeded = imax(0, dn0 - j), insNeeded = imax(0, in0 + j);
                const bool insBar = (insTop && c > 1) || (insBot && c < cols - 1);
                const CellOut o = msa_cell<true>(K, R, xM, xD, xI, yM, yD, MS[j], IN[j], rf[j], ref0, insBar, hlr[j], delNeeded, insNeeded, tab);
                const int nM = visit ? o.ms : K.subfloor, nD = visit ? o.del : K.subfloor, nI = visit ? o.ins : K.subfloor;
                const bool good = visit && o.good;
                word |= o.code << (4 * j);
                gCur |= (good ? 1u : 0u) << j;
                if (r == rows && visit) {
                    // candidates of the final scan (jni/...JNI.c:672-686): state-major, first max wins
                    const int s0 = nM & SMASK, s1 = nD & SMASK, s2 = nI & SMASK;
                    if (s0 > bestScore || (s0 == bestScore && 0 < bestState)) { bestScore = s0; bestCol = c; bestState = 0; bestPacked = nM; }
                    if (s1 > bestScore || (s1 == bestScore && 1 < bestState)) { bestScore = s1; bestCol = c; bestState = 1; bestPacked = nD; }
                    if (s2 > bestScore) { bestScore = s2; bestCol = c; bestState = 2; bestPacked = nI; }
                }
                xM = MS[j]; xD = DL[j]; xI = IN[j];
                MS[j] = nM; DL[j] = nD; IN[j] = nI;
                yM = nM; yD = nD;
                ref0 = rf[j];
            }
            tbT[(long long)s * rs + r] = word;
            if (gCur) {
                const int old = mmRow;
                const int first = c0 + __ffs(gCur) - 1, lastc = c0 + 31 - __clz(gCur);
                const int mn = (old == MM_NONE) ? first : (old >> 16);
                A[r].y = (mn << 16) | lastc;
            }
            if (s + 1 < nstrips) {
                const int sub = K.subfloor;
                const bool real = ((MS[W - 1] & SMASK) != sub) || ((DL[W - 1] & SMASK) != sub) || ((IN[W - 1] & SMASK) != sub);
                if (real) { if (loC == INT_MAX) loC = r; hiC = r; }
                if (loC != INT_MAX) rec[r] = make_int4(MS[W - 1], DL[W - 1], IN[W - 1], 0);
            }
            gPrev = gCur;
            dM = lM; dD = lD; dI = lI;
            const int r2 = r + 1;
            if (r2 > rows) newStrip = true;
            else if (gCur != 0u || s == 0 || (r2 >= loP && r2 <= hiP + 1)) r = r2;
            else if (r2 < loP && loP <= hiP && !(S.debug & 2)) {
                if (loC != INT_MAX && s + 1 < nstrips) for (int q = r2; q < loP; ++q) rec[q] = make_int4(K.subfloor, K.subfloor, K.subfloor, 0);
                r = loP; dM = K.subfloor; dD = K.subfloor; dI = K.subfloor;
#pragma unroll
                for (int j = 0; j < W; ++j) { MS[j] = K.subfloor; DL[j] = K.subfloor; IN[j] = K.subfloor; }
            }
            else if (S.debug & 2) r = r2;
            else newStrip = true;
        }
    }
    if (S.debug & 4) { atomicAdd(S.stats, nUnits); atomicAdd(S.stats + 1, nIters); }
}

// ---------------- K2: per-row control flow of the reference, result, score2 + traceback2 ----------------
template <int W>
__global__ void __launch_bounds__(128) msa_strip_finish_kernel(StripParams S) {
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= strip_count(S)) return;
    const MsaParams& P = S.P;
    const long long off = S.hdr[slot];
    if (off < 0) return;
    const int id = S.list[S.chunkStart + slot];
    const bbm_msa_task task = P.tasks[id];
    TaskCtx T;
    resolve_task(task, P.bandwidth, P.ratio, T);
    const int rows = T.rows, cols = T.cols;
    const int8_t* __restrict__ read = P.reads + task.read_off;
    const int8_t* __restrict__ ref = P.refs + task.ref_off + T.a;
    bbm_msa_out* out = P.outs + id;
    const int minScore_off = (int)((unsigned)T.minScore << TBITS);
    const int floor_ = minScore_off - ((rows - 1) * P_MATCH2 + P_MATCH);
    const int subfloor = floor_ - 5 * P_MATCH2;
    const StripBlock B = strip_block(S, off, rows, cols);
    const int2* A = B.A;
    // row bookkeeping (jni/...JNI.c:440-449, 554-556, 660-668 with halfband==0)
    int prevMin = 1, prevMax = cols, lastColStart = 1;
    long long iters = 0;
    bool broke = false;
    for (int r = 1; r <= rows; ++r) {
        const int colStart = prevMin, colStop = prevMax;
        if (colStart < 0 || colStop < colStart) { broke = true; break; }
        const int mm = A[r].y;
        const int curMin = (mm == MM_NONE) ? -1 : (mm >> 16), curMax = (mm == MM_NONE) ? -2 : (mm & 0xffff);
        const int lastc = imin(cols, imax(colStop, curMax) + 1);
        iters += lastc - colStart + 1;
        if (r == rows) lastColStart = colStart;
        prevMin = curMin; prevMax = curMax;
        if (curMin < 0 && r < rows) { broke = true; break; }
    }
    const int4 f = S.fin[slot];
    int maxCol = f.y, maxState = f.z, maxScoreOff = f.x, maxPacked = f.w;
    if (broke) { maxCol = 1; maxState = 0; maxScoreOff = BADOFF; maxPacked = BADOFF; }
    else if (f.y < 0 || f.x <= subfloor) {
        // every visited cell of the last row holds subfloor: the scan stops at the first one, (rows, colStart-1) included
        maxCol = lastColStart > 1 ? lastColStart - 1 : 1; maxState = 0; maxScoreOff = subfloor; maxPacked = subfloor;
    }
    const int fail = (maxScoreOff < minScore_off) ? 1 : 0;
    const bool javaMode = (T.flags & (BBM_TF_RAW_LIMITED | BBM_TF_RAW_UNLIMITED)) == 0;
    out->path = 0; out->iterations = iters; out->status = 0; out->score_len = 0; out->match_len = -1; out->pad_ = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) out->score[q] = 0;
    if (fail && javaMode) { out->result[0] = rows; out->result[1] = 0; out->result[2] = 0; out->result[3] = 0; out->result[4] = 1; }
    else {
        out->result[0] = rows; out->result[1] = maxCol; out->result[2] = maxState;
        out->result[3] = fail ? maxScoreOff : (maxScoreOff >> TBITS); out->result[4] = fail;
    }
    if (fail || (T.flags & (BBM_TF_SCORE | BBM_TF_TRACEBACK)) == 0) return;

    const bool wantTb = (T.flags & BBM_TF_TRACEBACK) != 0 && P.match_buf != nullptr;
    int8_t* mslot = nullptr; long long mcap = 0;
    if (wantTb) { mslot = P.match_buf + P.match_off[id]; mcap = P.match_off[id + 1] - P.match_off[id]; }
    const unsigned int* tb = B.tb;
    int row = rows, col = maxCol, state = maxState, stateTime = 0, nOps = 0, gapsSeen = 0;
    const int bestRefStop = T.a + col - 1;
    while (row > 0 && col > 0) {
        const int st = (col - 1) / W, j = (col - 1) - st * W;
        const unsigned code = (tb[(long long)st * B.rs + row] >> (4 * j)) & 15u;
        int prev; char op = 0;
        if (state == ST_MS) {
            prev = code & 3u;
            if (wantTb) { const int c = read[row - 1], rfb = ref[col - 1]; op = (c == rfb) ? 'm' : ((!base_defined(c) || !base_defined(rfb)) ? 'N' : 'S'); }
            row--; col--;
        } else if (state == ST_DEL) {
            prev = ((code >> 2) & 1u) ? ST_DEL : ST_MS;
            if (wantTb) { if (ref[col - 1] == '-') { op = '-'; gapsSeen++; } else op = 'D'; }
            col--;
        } else {
            prev = ((code >> 3) & 1u) ? ST_INS : ST_MS;
            op = (col == 0) ? 'X' : ((col >= cols) ? 'Y' : 'I');
            row--;
        }
        if (wantTb && nOps < mcap) mslot[mcap - 1 - nOps] = op;
        nOps++;
        stateTime = (state == prev) ? stateTime + 1 : 0;
        state = prev;
    }
    const int rowEnd = row, colEnd = col;
    if (wantTb && colEnd != rowEnd) { int rr = rowEnd; while (rr > 0) { if (nOps < mcap) mslot[mcap - 1 - nOps] = 'X'; nOps++; rr--; } }
    if (T.flags & BBM_TF_SCORE) {
        int colf = colEnd; if (rowEnd > colEnd) colf -= rowEnd;
        const int bestRefStart = T.a + colf;
        int padLeft = 0, padRight = 0;
        if (bestRefStart < T.a) padLeft = imax(0, T.a - bestRefStart);
        else if (bestRefStart == T.a && state == ST_INS) padLeft = stateTime;
        if (bestRefStop > score_ref_end(T)) padRight = imax(0, bestRefStop - score_ref_end(T));
        else if (bestRefStop == score_ref_end(T) && maxState == ST_INS) padRight = maxPacked & TMASK;
        out->score[0] = maxScoreOff >> TBITS; out->score[1] = bestRefStart; out->score[2] = bestRefStop;
        out->score[3] = rows; out->score[4] = maxCol; out->score[5] = maxState; out->score[6] = padLeft; out->score[7] = padRight;
        out->score_len = (padLeft > 0 || padRight > 0) ? 8 : 6;
    }
    if (!wantTb) return;
    const long long total = (long long)nOps + (long long)gapsSeen * 127;
    if (nOps > mcap || total > mcap) { out->status = BBM_E_CAPACITY; out->match_len = -1; return; }
    const long long shift = mcap - nOps;
    if (gapsSeen == 0) {
        if (shift > 0) for (int i = 0; i < nOps; ++i) mslot[i] = mslot[shift + i];
    } else {
        // gapped reference: every '-' expands to GAPLEN 'D' (…JNI.java:478-493); the staged ops sit at the end of the slot
        long long j = 0;
        for (int i = 0; i < nOps; ++i) {
            const int8_t c = mslot[shift + i];
            if (c != '-') mslot[j++] = c;
            else { if (j + 128 > shift + i + 1) { out->status = BBM_E_CAPACITY; out->match_len = -1; return; } for (int k = 0; k < 128; ++k) mslot[j++] = 'D'; }
        }
    }
    out->match_len = (int)total;
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_msa_strip_width() { return SW; }
extern "C" int bbm_msa_strip_blocks_per_sm() { return STRIP_BLOCKS_PER_SM; }
extern "C" int bbm_msa_strip_max_cols() { return STRIP_MAX_COLS; }
extern "C" unsigned long long bbm_msa_strip_task_bytes(int rows, int cols) { return strip_task_bytes(rows, cols); }
// fixed part of the scratch: per-slot header + final-row candidates, lane-private record arrays
extern "C" size_t bbm_msa_strip_fixed_bytes(int chunkCount, int maxRows, int blocks) {
    return (size_t)chunkCount * (8 + 16) + (size_t)blocks * STRIP_THREADS * ((size_t)maxRows + 2) * 16 + 256;
}
extern "C" int bbm_launch_msa_strip(const MsaParams* P, const int* list, const unsigned int* endPtr, unsigned int base, int chunkStart, int chunkCount,
                                    int maxRows, void* scratch, size_t scratchBytes, unsigned int* counter, unsigned long long* poolCursor,
                                    int blocks, int debug, unsigned long long* stats, cudaStream_t st) {
    StripParams S; S.stats = stats;
    S.P = *P; S.list = list; S.endPtr = endPtr; S.base = base; S.chunkStart = chunkStart; S.chunkCount = chunkCount;
    S.rowStride = maxRows + 2;
    char* p = (char*)scratch;
    S.fin = (int4*)p; p += (size_t)chunkCount * 16;
    S.hdr = (long long*)p; p += (((size_t)chunkCount * 8) + 15) & ~(size_t)15;
    S.rec = (int4*)p; p += (size_t)blocks * STRIP_THREADS * S.rowStride * 16;
    S.pool = p; S.poolBytes = scratchBytes - (size_t)(p - (char*)scratch); S.poolCursor = poolCursor;
    S.counter = counter; S.debug = debug;
    const int pb = (chunkCount + 127) / 128;
    msa_strip_prep_kernel<<<pb, 128, 0, st>>>(S);
    cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) return (int)e;
    msa_strip_fill_kernel<SW><<<blocks, STRIP_THREADS, 0, st>>>(S);
    e = cudaGetLastError(); if (e != cudaSuccess) return (int)e;
    msa_strip_finish_kernel<SW><<<pb, 128, 0, st>>>(S);
    return (int)cudaGetLastError();
}
