// msa_narrow.cuh — thread-per-alignment MultiStateAligner11ts for alignments whose pruned fill stays near the diagonal.
//
// fillLimitedX (jni/MultiStateAligner11tsJNI.c:361-704) only visits, per read row, the columns between the previous
// row's first and last "good" column.  For the common case of a read that matches its window with few edits and a
// tight minScore (BBMapThread.scoreSlow passes max(scoreNoIndels, limit), current/align2/BBMapThread.java:306) that
// interval is ~10 columns wide and hugs the diagonal.  This kernel keeps a fixed window of DW=16 diagonals
// (d = col-row in [DLO, DLO+DW)) of all three states in registers of ONE thread, so 32 alignments run per warp with no
// inter-lane traffic and no wavefront fill/drain.  In diagonal coordinates the MS predecessor is the same register, the
// INS predecessor the register to the right (previous row) and the DEL predecessor the register to the left (this row).
//
// Exactness: cells outside the window are never evaluated, which is only valid while they cannot be "good".  A cell can
// only be good if one of its three predecessors is, so by induction it suffices that (a) row 1 has no good column
// outside the window interior and (b) the two edge diagonals of the window never hold a good cell.  Both are checked;
// on violation (and for gapped references, bands, or unlimited fills) the alignment is appended to the list of the
// register-tiled kernel (msa_tiled.cuh), which computes the full rectangle.  Everything observable — result vector,
// iteration counter (from per-row minGoodCol/maxGoodCol), score2, traceback2 — is the same as the reference's.
#pragma once
#include "msa_kernels.cuh"
#include "msa_cell.cuh"

namespace bbm {

constexpr int NDW = 16;        // diagonals per thread
constexpr int NDLO = -3;       // first diagonal of the window (col - row)
constexpr int NARROW_THREADS = 128;

typedef CellTables NarrowShared;

// maxSlack: alignments whose minScore leaves more than this many points below the best possible score almost always wander out
// of the 16-diagonal window (an indel read scored against a ratio-based limit): they skip the attempt (0 = try everything)
__device__ __forceinline__ bool narrow_eligible(const TaskCtx& T, int maxSlack) {
    const int D = T.cols - T.rows;
    if (!(T.limited && T.halfband < 1 && D >= NDLO + 1 && D <= NDLO + NDW - 2 && T.rows <= MAXR - 2 && T.rows >= 2)) return false;
    return maxSlack <= 0 || ((T.rows - 1) * 100 + 70 - T.minScore) <= maxSlack;
}

// map a reference byte for comparison with a call: 'N' never matches (jni/...JNI.c:468-469)
__device__ __forceinline__ int map_ref(int v) { return v == 'N' ? 0x100 : v; }

__device__ void msa_narrow_warp(const MsaParams& P, const int* __restrict__ list, int first, int nlist,
                                const NarrowShared& sh, unsigned long long* tbWarp, unsigned int* classCursors, int* classLists, int useStrip) {
    const int lane = threadIdx.x & 31;
    const int k = first + lane;
    bool alive = k < nlist;
    int id = 0;
    bbm_msa_task task = {};
    TaskCtx T = {};
    if (alive) {
        id = list[k];
        task = P.tasks[id];
        resolve_task(task, P.bandwidth, P.ratio, T);
    }
    const int rows = alive ? T.rows : 0, cols = T.cols;
    const int8_t* __restrict__ read = P.reads + task.read_off;
    const int8_t* __restrict__ ref = P.refs + task.ref_off + T.a;      // ref[c-1] is column c
    bbm_msa_out* out = P.outs + id;

    const int maxGain = (rows - 1) * P_MATCH2 + P_MATCH;
    const int minScore_off = (int)((unsigned)T.minScore << TBITS);
    const int floor_ = minScore_off - maxGain;
    const int subfloor = floor_ - 5 * P_MATCH2;
    const int D = cols - rows;
    CellConst K; K.floor_ = floor_; K.subfloor = subfloor;

    bool bail = false;
    // ---- suffix costs of the limits (jni/...JNI.c:413-438).  No '-' in the window => all costs >= 0, so
    //      limit[i] = max(minScore_off - S(i), floor) with S the plain suffix sum. ----
    int Sv = 0;     // becomes S_v(1): cost of read indices 1..rows-1
    int Sh = 0;     // becomes S_h(c) for the running column
    int signAcc = 0;
    if (alive) {
        bool pd = false;
        for (int i = rows - 1; i >= 1; --i) {
            const int c = read[i];
            signAcc |= c;
            const bool d = base_defined(c);
            Sv += d ? (pd ? P_MATCH2 : P_MATCH) : 0;
            pd = d;
        }
        const int call1 = read[0];
        signAcc |= call1;
        const int vl1 = imax(minScore_off - Sv, floor_);
        // backward over the window: S_h(c) for c=cols..1 and the goodness of row 1 outside the register window.
        // Row 1 (jni/...JNI.c:491-563 with row 0 all zero): MS = {MATCH | SUB | NOCALL}, DEL is barred (row<3),
        // INS only exists at column 1.  Columns inside the window are evaluated by the main loop.
        pd = false;
        const int winLo = 1 + NDLO + 1, winHi = 1 + NDLO + NDW - 2;     // interior columns of the window at row 1
        for (int c = cols; c >= 1; --c) {
            // here Sh == S_h(c)
            const int rb = ref[c - 1];
            signAcc |= rb;
            if (rb == '-') bail = true;
            if (c > winHi || c < winLo) {
                const int ref1 = map_ref(rb);
                const int ref0 = c < 2 ? '!' : map_ref(ref[c - 2]);
                if (ref0 == '?') bail = true;                               // prevMatch against the row-0 sentinel: leave it to the tiled kernel
                const bool match = (call1 == ref1);
                const int score = match ? P_MATCH : ((rb == 'N' || call1 == 'N') ? 0 : P_SUB);
                const int limit = imax(vl1, imax(minScore_off - Sh, floor_));
                const int insNeeded = imax(0, (rows - 1) - (cols - c) - 1);
                const int limit3 = imax(floor_, match ? limit - P_MATCH2 : limit - P_SUB3);
                const bool skip = (0 <= limit3);                           // all three diagonal inputs are 0
                const int lim2 = insNeeded > 0 ? limit - sh.insc[imin(insNeeded, PEN_TAB - 1)] : limit;
                if (!skip && score >= lim2) bail = true;                    // a good cell outside the window interior
            }
            // advance to S_h(c-1): add the cost of reference index c-1 (jni/...JNI.c:429-438)
            const bool d = base_defined(rb);
            Sh += d ? (pd ? P_MATCH2 : P_MATCH) : 0;
            pd = d;
        }
        // now Sh == S_h(0)
        if (signAcc & 0x80) bail = true;       // bytes >= 0x80: keep the byte-exact path in the tiled kernel
    }

    // ---- register window ----
    int MS[NDW], DL[NDW], IN[NDW], rf[NDW], hlr[NDW];
#pragma unroll
    for (int j = 0; j < NDW; ++j) { MS[j] = 0; DL[j] = 0; IN[j] = 0; rf[j] = 0x200; hlr[j] = 0; }
    // columns of the window at row 1: c = 1 + NDLO + j.  Preload ref bytes / horizLimit for them (slot j holds column cr+j).
    // S_h is walked forward again from S_h(0): S_h(c) = S_h(c-1) - cost(index c-1).
    int ShRun = Sh;              // S_h of the last column that entered the window (starts at S_h(0))
    // helper state for forward evaluation of S_h: S_h(c) = S_h(c-1) - cost(c-1); cost(i) needs defined(ref[i]) and defined(ref[i+1]) (i+1<=cols-1)
    auto hcost_at = [&](int i) -> int {      // i = reference index inside the window, 0 <= i <= cols-1
        const bool d = base_defined(ref[i]);
        const bool pdn = (i + 1 <= cols - 1) && base_defined(ref[i + 1]);
        return d ? (pdn ? P_MATCH2 : P_MATCH) : 0;
    };
    if (alive && !bail) {
        // fill slots for row 1
#pragma unroll
        for (int j = 0; j < NDW; ++j) {
            const int c = 1 + NDLO + j;
            if (c >= 1) {
                // advance ShRun from S_h(nextCol-1) to S_h(c)
                if (c <= cols) {
                    ShRun -= hcost_at(c - 1);
                    rf[j] = map_ref(ref[c - 1]);
                    hlr[j] = (c == cols) ? minScore_off : imax(minScore_off - ShRun, floor_);
                }
            }
        }
    }

    int minGoodPrev = 1, maxGoodPrev = cols;
    long long iters = 0;
    bool broke = false;
    int call0 = '?';
    int SvRun = Sv;                 // S_v(r) for the current row r (starts at S_v(1))
    int nIters = rows;
    // warp-uniform trip count
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) nIters = imax(nIters, __shfl_xor_sync(FULL, nIters, o));

    int bestScore = INT_MIN, bestCol = -1, bestState = -1, bestPacked = 0;
    int lastColStart = 1;

    for (int r = 1; r <= nIters; ++r) {
        const bool run = alive && !bail && !broke && r <= rows;
        if (!__any_sync(FULL, run)) break;
        unsigned long long word = 0;
        if (run) {
            const int cr = r + NDLO;                          // column of slot 0
            const int call1 = read[r - 1];
            CellRow R;
            R.call1 = call1; R.call0 = call0; R.callN = (call1 == 'N');
            R.vlimit = (r == rows) ? minScore_off : imax(minScore_off - SvRun, floor_);
            R.delBar = (r < 3) || (r > rows - 3);
            const bool insTop = (r < 2), insBot = (r > rows - 2);
            const int colStart = minGoodPrev;                 // halfband == 0 here
            const int col0 = sh.insc[imin(r, PEN_TAB - 1)];
            if (r == rows) lastColStart = colStart;
            unsigned gmask = 0;
            int lMS = subfloor, lDL = subfloor;               // left neighbour of slot 0 is outside the window
            int ref0 = (cr - 1 >= 1 && cr - 1 <= cols) ? map_ref(ref[cr - 2]) : '!';
#pragma unroll
            for (int j = 0; j < NDW; ++j) {
                const int c = cr + j;
                const int dMS = MS[j], dDL = DL[j], dIN = IN[j];           // (r-1, c-1): same diagonal
                int uMS = subfloor, uIN = subfloor;                         // (r-1, c): diagonal to the right
                if (j + 1 < NDW) { uMS = MS[j + 1]; uIN = IN[j + 1]; }
                const int ref1 = rf[j];
                const bool visit = (c >= colStart) && (c >= 1) && (c <= cols);
                const int delNeeded = (-(NDLO + j) - 1) > 0 ? (-(NDLO + j) - 1) : 0;      // max(0,row-col-1): fixed per diagonal
                const int insNeeded = imax(0, NDLO + j - D - 1);                            // max(0,(rows-row)-(cols-col)-1)
                const bool insBar = (insTop && c > 1) || (insBot && c < cols - 1);
                const CellOut o = msa_cell<true>(K, R, dMS, dDL, dIN, lMS, lDL, uMS, uIN, ref1, ref0,
                                                 insBar, hlr[j], delNeeded, insNeeded, sh);
                int nMS = visit ? o.ms : subfloor, nDL = visit ? o.del : subfloor, nIN = visit ? o.ins : subfloor;
                const unsigned code = o.code;
                const bool good = visit && o.good;
                if (NDLO + j < 0) { if (c == 0) { nMS = col0; nDL = col0; nIN = col0; } }   // column 0 of the matrix (…JNI.java:105-111)
                word |= (unsigned long long)code << (4 * j);
                gmask |= (good ? 1u : 0u) << j;
                MS[j] = nMS; DL[j] = nDL; IN[j] = nIN;
                lMS = nMS; lDL = nDL;
                ref0 = ref1;
            }
            // row bookkeeping (jni/...JNI.c:440-449, 554-556, 660-668 with halfband==0)
            if (gmask & ((1u << 0) | (1u << (NDW - 1)))) bail = true;               // a good cell on a window edge
            const int curMin = gmask ? cr + __ffs(gmask) - 1 : -1;
            const int curMax = gmask ? cr + 31 - __clz(gmask) : -2;
            {
                const int colStop = maxGoodPrev;
                const int lastc = imin(cols, imax(colStop, curMax) + 1);
                iters += lastc - colStart + 1;
            }
            minGoodPrev = curMin; maxGoodPrev = curMax;
            if (curMin < 0 && r < rows) broke = true;           // next row would break out of the fill
            // slide the column-indexed registers one column to the right
#pragma unroll
            for (int j = 0; j + 1 < NDW; ++j) { rf[j] = rf[j + 1]; hlr[j] = hlr[j + 1]; }
            {
                const int c = cr + NDW;                          // column entering at the right edge for row r+1
                int v = 0x200, h = 0;
                if (c >= 1 && c <= cols) {
                    ShRun -= hcost_at(c - 1);
                    v = map_ref(ref[c - 1]);
                    h = (c == cols) ? minScore_off : imax(minScore_off - ShRun, floor_);
                }
                rf[NDW - 1] = v; hlr[NDW - 1] = h;
            }
            // S_v(r+1) = S_v(r) - cost(read index r)
            if (r < rows) {
                const bool d = base_defined(read[r]);
                const bool pdn = (r + 1 < rows) && base_defined(read[r + 1]);
                SvRun -= d ? (pdn ? P_MATCH2 : P_MATCH) : 0;
            }
            call0 = call1;
        }
        if (r <= MAXR) tbWarp[(size_t)r * 32 + lane] = word;
    }

    if (!alive) return;
    if (!bail && !broke) {
        // final scan candidates (jni/...JNI.c:672-686): state-major, first max wins; registers hold the last row
#pragma unroll
        for (int st = 0; st < 3; ++st) {
#pragma unroll
            for (int j = 0; j < NDW; ++j) {
                const int c = rows + NDLO + j;
                const bool visit = (c >= lastColStart) && (c >= 1) && (c <= cols);
                const int v = st == 0 ? MS[j] : (st == 1 ? DL[j] : IN[j]);
                const int x = v & SMASK;
                if (visit && x > bestScore) { bestScore = x; bestCol = c; bestState = st; bestPacked = v; }
            }
        }
    }
    if (bail) {
        // hand over to the register-tiled kernel of the right width
        const int kcls = (useStrip && strip_eligible(T) && strip_bucket(T) < useStrip) ? CLASS_STRIP : classify(T);
        const unsigned pos = atomicAdd(&classCursors[kcls], 1u);
        classLists[pos] = id;
        return;
    }

    // ---- result (jni/...JNI.c:672-703) ----
    int maxCol = bestCol, maxState = bestState, maxScoreOff = bestScore, maxPacked = bestPacked;
    if (broke || bestCol < 0) { maxCol = 1; maxState = 0; maxScoreOff = BADOFF; maxPacked = BADOFF; }
    else if (bestScore == subfloor && lastColStart > 1) { maxCol = lastColStart - 1; maxState = 0; maxPacked = subfloor; }  // (rows,colStart-1) is scanned first
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

    // ---- score2 + traceback2 over the per-row code words (…JNI.java:376-495, 537-658) ----
    const bool wantTb = (T.flags & BBM_TF_TRACEBACK) != 0 && P.match_buf != nullptr;
    int8_t* mslot = nullptr; long long mcap = 0;
    if (wantTb) { mslot = P.match_buf + P.match_off[id]; mcap = P.match_off[id + 1] - P.match_off[id]; }
    int row = rows, col = maxCol, state = maxState, stateTime = 0, nOps = 0;
    const int bestRefStop = T.a + col - 1;
    while (row > 0 && col > 0) {
        const int j = col - (row + NDLO);
        unsigned code = 0;
        if (j >= 0 && j < NDW) code = (unsigned)(tbWarp[(size_t)row * 32 + lane] >> (4 * j)) & 15u;
        int prev; char op = 0;
        if (state == ST_MS) {
            prev = code & 3u;
            if (wantTb) { const int c = read[row - 1], rfb = ref[col - 1]; op = (c == rfb) ? 'm' : ((!base_defined(c) || !base_defined(rfb)) ? 'N' : 'S'); }
            row--; col--;
        } else if (state == ST_DEL) {
            prev = ((code >> 2) & 1u) ? ST_DEL : ST_MS;
            op = 'D';
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
    if (nOps > mcap) { out->status = BBM_E_CAPACITY; out->match_len = -1; return; }
    const long long shift = mcap - nOps;
    if (shift > 0) for (int i = 0; i < nOps; ++i) mslot[i] = mslot[shift + i];
    out->match_len = nOps;
}

__global__ void __launch_bounds__(NARROW_THREADS, 4) msa_narrow_kernel(MsaParams P, const int* __restrict__ list, int nlist, unsigned int* counter,
                                                                     unsigned long long* tbAll, long long tbWordsPerWarp,
                                                                     unsigned int* classCursors, int* classLists, int useStrip) {
    __shared__ NarrowShared sh;
    cell_tables_init(sh);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long long gwarp = (long long)blockIdx.x * (NARROW_THREADS / 32) + (threadIdx.x >> 5);
    unsigned long long* tbWarp = tbAll + gwarp * tbWordsPerWarp;
    for (;;) {
        unsigned first = 0;
        if (lane == 0) first = atomicAdd(counter, 32u);
        first = __shfl_sync(FULL, first, 0);
        if (first >= (unsigned)nlist) break;
        msa_narrow_warp(P, list, (int)first, nlist, sh, tbWarp, classCursors, classLists, useStrip);
        __syncwarp();
    }
}

}  // namespace bbm
