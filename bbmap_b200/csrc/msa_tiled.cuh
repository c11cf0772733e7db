// msa_tiled.cuh — warp-per-alignment, register-tiled MultiStateAligner11ts fill + score2 + traceback2.
//
// Reproduces, bit for bit, what the reference computes for one alignment:
//   fillLimitedX   jni/MultiStateAligner11tsJNI.c:361-704   (≡ current/align2/MultiStateAligner11ts.java:131-607)
//   fillUnlimited  jni/MultiStateAligner11tsJNI.c:100-314   (≡ MultiStateAligner11ts.java:624-878)
//   score2         current/align2/MultiStateAligner11tsJNI.java:537-658
//   traceback2     current/align2/MultiStateAligner11tsJNI.java:376-495
// without ever materialising the reference's 12-byte-per-cell `packed` matrix.
//
// Layout: one warp owns one alignment.  Lane L owns the W consecutive columns [L*W+1, L*W+W]; the three state rows
// of the previous read row live in registers.  The warp sweeps the rectangle as a skewed wavefront: at step t lane L
// processes read row t-L, so the left-neighbour dependency of the DEL state and the diagonal/upper dependencies of the
// MS / INS states are one __shfl_up away.  Per (row,lane) the kernel stores one word of 4-bit predecessor codes
// (2 bits MS, 1 bit DEL, 1 bit INS) — exactly the decisions score2/traceback2 would take from the full matrix.
//
// Pruning (fillLimitedX): the reference visits, per row, the interval [colStart,last] derived from the previous row's
// good columns.  Here every lane evaluates its cells; a cell left of colStart is forced to the reference's
// `subfloor`, and cells right of the reference's interval evaluate to pure subfloor by construction (all their inputs
// are sub-limit), so values of every cell the reference can ever read are identical.  The reference's iteration counter
// and its break conditions are reproduced from the per-row (minGoodCol,maxGoodCol) pair, which travels with the row
// from lane to lane.  With a band (halfband>0) the right edge depends on the *whole* previous row; a violated
// assumption is detected exactly and the alignment is re-run by the row-sequential generic kernel (msa_generic.cuh).
#pragma once
#include "msa_common.cuh"
#include "msa_cell.cuh"

namespace bbm {

constexpr unsigned FULL = 0xffffffffu;
constexpr int MM_NONE = 0x7fff0000;   // packed (minGood<<16 | maxGood): empty row

template <int W> struct TbWord { using type = unsigned long long; };
template <> struct TbWord<1> { using type = unsigned int; };
template <> struct TbWord<2> { using type = unsigned int; };
template <> struct TbWord<3> { using type = unsigned int; };
template <> struct TbWord<4> { using type = unsigned int; };
template <> struct TbWord<5> { using type = unsigned int; };
template <> struct TbWord<6> { using type = unsigned int; };
template <> struct TbWord<7> { using type = unsigned int; };
template <> struct TbWord<8> { using type = unsigned int; };

struct WarpShared {
    int vl[MAXR + 2];          // vertLimit
    signed char read[MAXR + 8];
};

typedef CellTables BlockShared;

struct TaskCtx {
    int rows, cols, a, b;       // window [a,b] inside the reference array
    int minScore;               // as passed to the fill (after -120 when Java semantics)
    int limited;                // 1 limited, 0 unlimited
    int halfband;
    int flags;
};
// refEndLoc as score2 sees it: on a gapped reference the fill ran on [0, greflimit] but score() passes gstop = greflimit-1
// (MultiStateAligner11tsJNI.java:127 vs :508-516)
__device__ __forceinline__ int score_ref_end(const TaskCtx& T) { return T.b - ((T.flags & BBM_TF_GAPPED) ? 1 : 0); }

// ---- suffix "limit" recurrences  h[i] = max(h[i+1] - cost_i, floor)  as a scan over g(x)=max(x-A,B) ----
struct GFun { int A, B; };
__device__ __forceinline__ GFun g_compose(GFun first, GFun second) {   // second ∘ first
    GFun r; r.A = first.A + second.A; r.B = imax(first.B - second.A, second.B); return r;
}
__device__ __forceinline__ int g_apply(GFun g, int x) { return imax(x - g.A, g.B); }
constexpr int G_NEG = -(1 << 30);

// cost of reference base `c` in the horizLimit recurrence (jni/...JNI.c:429-438)
__device__ __forceinline__ int hcost(int c, bool prevDefined, bool& definedOut) {
    const bool d = base_defined(c);
    definedOut = d;
    return d ? (prevDefined ? P_MATCH2 : P_MATCH) : ((prevDefined && c == '-') ? P_DEL : 0);
}

template <int W, bool LIMITED, bool BAND, bool DUMP>
__device__ void msa_fill_task(const MsaParams& P, const TaskCtx& T, const bbm_msa_task& task, long long taskId,
                              WarpShared& ws, const BlockShared& bs, unsigned long long* scratch, bbm_msa_out* out) {
    using tbw = typename TbWord<W>::type;
    const int lane = threadIdx.x & 31;
    const int rows = T.rows, cols = T.cols;
    const int8_t* __restrict__ read = P.reads + task.read_off;
    const int8_t* __restrict__ ref = P.refs + task.ref_off + T.a;     // ref[0] is column 1
    tbw* tb = reinterpret_cast<tbw*>(scratch);
    const int nAct = (cols + W - 1) / W;                 // lanes that own at least one column
    const int c0 = lane * W + 1;                         // first column of this lane

    const int maxGain = (rows - 1) * P_MATCH2 + P_MATCH;
    const int minScore_off = (int)((unsigned)T.minScore << TBITS);
    const int floor_ = LIMITED ? minScore_off - maxGain : 0;
    const int subfloor = LIMITED ? floor_ - 5 * P_MATCH2 : 0 - 2 * maxGain;
    const int hb = T.halfband;
    CellConst K; K.floor_ = floor_; K.subfloor = subfloor;

    // ---- stage the read; vertLimit (jni/...JNI.c:413-425) ----
    for (int i = lane; i < rows; i += 32) ws.read[i] = read[i];
    __syncwarp();
    if (LIMITED) {
        const int R = (rows + 31) >> 5;
        const int lo = lane * R, hi = imin(rows, lo + R);     // this lane's rows [lo,hi)
        GFun g; g.A = 0; g.B = G_NEG;
        for (int i = hi - 1; i >= lo; --i) {
            const bool prevDef = (i + 1 < rows) && base_defined(ws.read[i + 1]);
            const int cost = base_defined(ws.read[i]) ? (prevDef ? P_MATCH2 : P_MATCH) : 0;
            GFun s; s.A = cost; s.B = floor_;
            g = g_compose(g, s);
        }
        GFun inc = g;   // inclusive suffix composition over lanes >= lane (higher lanes applied first)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            GFun other; other.A = __shfl_down_sync(FULL, inc.A, o); other.B = __shfl_down_sync(FULL, inc.B, o);
            if (lane + o < 32) inc = g_compose(other, inc);
        }
        int entryA = __shfl_down_sync(FULL, inc.A, 1), entryB = __shfl_down_sync(FULL, inc.B, 1);
        int x = minScore_off;
        if (lane < 31) { GFun e; e.A = entryA; e.B = entryB; x = g_apply(e, minScore_off); }
        for (int i = hi - 1; i >= lo; --i) {
            const bool prevDef = (i + 1 < rows) && base_defined(ws.read[i + 1]);
            const int cost = base_defined(ws.read[i]) ? (prevDef ? P_MATCH2 : P_MATCH) : 0;
            x = imax(x - cost, floor_);
            ws.vl[i] = x;
        }
        if (lane == 0) ws.vl[rows] = minScore_off;
        __syncwarp();
    }

    // ---- per-column constants: reference bases and horizLimit (jni/...JNI.c:427-438) ----
    int refc[W];          // reference byte of column c0+j; 0x100 for 'N' (never equals a call), 0x200 beyond the window
    int hl[W];
    unsigned nmask = 0, gapmask = 0;
#pragma unroll
    for (int j = 0; j < W; ++j) {
        const int c = c0 + j;
        int v = 0x200;
        if (c <= cols) {
            v = ref[c - 1];
            if (v == 'N') { v = 0x100; nmask |= 1u << j; }
            if (v == '-') gapmask |= 1u << j;
        }
        refc[j] = v;
        hl[j] = 0;
    }
    const int refLeft = (c0 >= 2 && c0 - 1 <= cols) ? (ref[c0 - 2] == 'N' ? 0x100 : (int)ref[c0 - 2]) : '!';  // ref0 of column c0
    if (LIMITED) {
        // hl[c] depends on the bases of columns c+1..cols, i.e. ref indices c..cols-1
        GFun g; g.A = 0; g.B = G_NEG;
        // this lane's ref indices: i = c0+j for j=W-1..0  (index i ↔ horizLimit[i], i in [c0, c0+W-1]), valid when i<=cols-1
#pragma unroll
        for (int j = W - 1; j >= 0; --j) {
            const int i = c0 + j;
            if (i <= cols - 1) {
                const int cb = ref[i];
                const bool prevDef = (i + 1 <= cols - 1) && base_defined(ref[i + 1]);
                bool d; const int cost = hcost(cb, prevDef, d);
                GFun s; s.A = cost; s.B = floor_;
                g = g_compose(g, s);
            }
        }
        GFun inc = g;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            GFun other; other.A = __shfl_down_sync(FULL, inc.A, o); other.B = __shfl_down_sync(FULL, inc.B, o);
            if (lane + o < 32) inc = g_compose(other, inc);
        }
        const int entryA = __shfl_down_sync(FULL, inc.A, 1), entryB = __shfl_down_sync(FULL, inc.B, 1);
        int x = minScore_off;
        if (lane < 31) { GFun e; e.A = entryA; e.B = entryB; x = g_apply(e, minScore_off); }
#pragma unroll
        for (int j = W - 1; j >= 0; --j) {
            const int i = c0 + j;
            if (i <= cols - 1) {
                const int cb = ref[i];
                const bool prevDef = (i + 1 <= cols - 1) && base_defined(ref[i + 1]);
                bool d; const int cost = hcost(cb, prevDef, d);
                x = imax(x - cost, floor_);
            } else if (i == cols) {
                x = minScore_off;
            }
            hl[j] = x;
        }
    }

    // ---- wavefront state ----
    int pMS[W], pDEL[W], pINS[W];
#pragma unroll
    for (int j = 0; j < W; ++j) { pMS[j] = 0; pDEL[j] = 0; pINS[j] = 0; }   // row 0 of the matrix is all zero
    int lastOldMS = 0, lastOldDEL = 0, lastOldINS = 0;     // my last column, row r-1
    int lastNewMS = 0, lastNewDEL = 0;                      // my last column, row r
    int mmNew = MM_NONE, mmOld = MM_NONE;                   // running (min,max) good column of my last two rows
    unsigned gPrev = 0xffffffffu;                           // good mask of my strip in the previous row
    // per-row bookkeeping (meaningful in the last active lane)
    int prevMin = 1, prevMax = cols;
    long long iters = 0;
    bool broke = false, misspec = false;
    // last-row candidates
    int bestScore = INT_MIN, bestCol = -1, bestState = -1, bestPacked = 0;
    int firstVisited = 0x7fffffff;

    const int steps = rows + nAct - 1;
    for (int t = 1; t <= steps; ++t) {
        const int r = t - lane;
        const bool active = (r >= 1) && (r <= rows) && (lane < nAct);
        bool deadNow = false;
        // values from the lane on my left
        int dMS = __shfl_up_sync(FULL, lastOldMS, 1);
        int dDEL = __shfl_up_sync(FULL, lastOldDEL, 1);
        int dINS = __shfl_up_sync(FULL, lastOldINS, 1);
        int lMS = __shfl_up_sync(FULL, lastNewMS, 1);
        int lDEL = __shfl_up_sync(FULL, lastNewDEL, 1);
        int mmInCur = __shfl_up_sync(FULL, mmNew, 1);
        int mmInPrev = __shfl_up_sync(FULL, mmOld, 1);
        if (lane == 0) {
            const int rr = imin(imax(r, 0), PEN_TAB - 1);
            const int up0 = rr >= 1 ? bs.insc[rr - 1] : 0;
            dMS = up0; dDEL = up0; dINS = up0;                 // column 0, row r-1
            lMS = bs.insc[rr]; lDEL = lMS;                     // column 0, row r
            mmInCur = MM_NONE; mmInPrev = MM_NONE;
        }
        if (active) {
            CellRow R;
            R.call1 = ws.read[r - 1];
            R.call0 = r < 2 ? '?' : ws.read[r - 2];
            R.callN = (R.call1 == 'N');
            R.vlimit = LIMITED ? ws.vl[r] : 0;
            R.delBar = (r < 3) || (r > rows - 3);
            const bool insTop = (r < 2), insBot = (r > rows - 2);
            const bool inP = LIMITED ? ((r == 1) || (mmInPrev != MM_NONE)) : true;
            const int dn0 = r - c0 - 1, in0 = (rows - r) - (cols - c0) - 1;
            unsigned gCur = 0;
            tbw word = 0;
            int ref0 = refLeft;
#pragma unroll
            for (int j = 0; j < W; ++j) {
                const int c = c0 + j;
                const int ref1 = refc[j];
                const bool inRange = (c <= cols);
                bool visit = inRange;
                if (LIMITED) {
                    const bool leftOK = inP || ((gPrev & ((2u << j) - 1u)) != 0u);
                    visit = visit && leftOK;
                    if (BAND) visit = visit && (c >= r - hb) && (c <= r + 2 * hb);
                }
                int delNeeded = 0, insNeeded = 0;
                if (LIMITED) {
                    delNeeded = imax(0, dn0 - j);            // max(0,row-col-1)
                    insNeeded = imax(0, in0 + j);            // max(0,(rows-row)-(cols-col)-1)  (<= rows)
                }
                const bool insBar = (insTop && c > 1) || (insBot && c < cols - 1);
                const CellOut o = msa_cell<LIMITED>(K, R, dMS, dDEL, dINS, lMS, lDEL, pMS[j], pINS[j], ref1, ref0,
                                                   insBar, LIMITED ? hl[j] : 0, delNeeded, insNeeded, bs);
                const int nMS = visit ? o.ms : subfloor, nDEL = visit ? o.del : subfloor, nINS = visit ? o.ins : subfloor;
                const unsigned code = o.code;
                const bool good = visit && o.good;
                word |= (tbw)code << (4 * j);
                gCur |= (good ? 1u : 0u) << j;
                if (inRange) {
                    if (DUMP && P.dump) {
                        // dense dump [3][rows+1][cols+2]; the host replays the reference's write pattern from it
                        const long long plane = (long long)(rows + 1) * (cols + 2);
                        const long long idx = (long long)r * (cols + 2) + c;
                        P.dump[idx] = nMS; P.dump[plane + idx] = nDEL; P.dump[2 * plane + idx] = nINS;
                    }
                    if (r == rows) {
                        // last row: candidates for the final scan (jni/...JNI.c:672-686): state-major, first max wins
                        const bool counted = LIMITED ? visit : true;
                        if (counted) {
                            if (c < firstVisited) firstVisited = c;
                            const int s0 = nMS & SMASK, s1 = nDEL & SMASK, s2 = nINS & SMASK;
                            // within one column the scan order between states is by state; across columns by (state, col)
                            if (s0 > bestScore || (s0 == bestScore && (0 < bestState))) { bestScore = s0; bestCol = c; bestState = 0; bestPacked = nMS; }
                            if (s1 > bestScore || (s1 == bestScore && (1 < bestState))) { bestScore = s1; bestCol = c; bestState = 1; bestPacked = nDEL; }
                            if (s2 > bestScore) { bestScore = s2; bestCol = c; bestState = 2; bestPacked = nINS; }
                        }
                    }
                }
                // shift the window: old row values of this column become the next column's diagonal
                dMS = pMS[j]; dDEL = pDEL[j]; dINS = pINS[j];
                pMS[j] = nMS; pDEL[j] = nDEL; pINS[j] = nINS;
                lMS = nMS; lDEL = nDEL;
                ref0 = ref1;
            }
            lastOldMS = dMS; lastOldDEL = dDEL; lastOldINS = dINS;
            lastNewMS = lMS; lastNewDEL = lDEL;
            tb[(size_t)t * 32 + lane] = word;
            if (LIMITED) {
                // running (min,max) good column of row r
                int mm = mmInCur;
                if (gCur) {
                    const int first = c0 + __ffs(gCur) - 1, lastc = c0 + 31 - __clz(gCur);
                    const int mn = (mm == MM_NONE) ? first : (mm >> 16);
                    mm = (mn << 16) | lastc;
                }
                mmOld = mmNew; mmNew = mm;
                gPrev = gCur;
                if (lane == nAct - 1) {
                    // row r is complete: replay the reference's per-row control flow (jni/...JNI.c:440-449, 660-668)
                    if (!broke) {
                        const int colStart = hb < 1 ? prevMin : imax(prevMin, r - hb);
                        const int colStop = hb < 1 ? prevMax : imin(prevMax, r + hb * 2 - 1);
                        if (colStart < 0 || colStop < colStart) { broke = true; }
                        else {
                            const int curMin = (mm == MM_NONE) ? -1 : (mm >> 16);
                            const int curMax = (mm == MM_NONE) ? -2 : (mm & 0xffff);
                            int lastc;
                            if (BAND) { lastc = imin(cols, colStop + 1); if (curMax > lastc) misspec = true; if (curMin >= 0 && curMin < colStart) misspec = true; }
                            else lastc = imin(cols, imax(colStop, curMax) + 1);
                            iters += lastc - colStart + 1;
                            prevMin = curMin; prevMax = curMax;
                            // an empty row ends the fill at the next row (colStart<0, jni/...JNI.c:449): nothing after it is visited
                            if (curMin < 0 && r < rows) { broke = true; deadNow = true; }
                        }
                    }
                }
            }
        }
        if (LIMITED) { if (__any_sync(FULL, deadNow)) break; }
    }
    __syncwarp();

    // ---- final scan over the last row: reduce (score desc, state asc, col asc) ----
    const int src = nAct - 1;
    int brokeAll = __shfl_sync(FULL, (int)broke, src);
    int missAll = __shfl_sync(FULL, (int)misspec, src);
    long long itAll = __shfl_sync(FULL, iters, src);
    int fv = firstVisited;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const int os = __shfl_xor_sync(FULL, bestScore, o), oc = __shfl_xor_sync(FULL, bestCol, o);
        const int ost = __shfl_xor_sync(FULL, bestState, o), op = __shfl_xor_sync(FULL, bestPacked, o);
        const int ofv = __shfl_xor_sync(FULL, fv, o);
        fv = imin(fv, ofv);
        bool take = false;
        if (oc >= 0) {
            if (bestCol < 0) take = true;
            else if (os > bestScore) take = true;
            else if (os == bestScore && (ost < bestState || (ost == bestState && oc < bestCol))) take = true;
        }
        if (take) { bestScore = os; bestCol = oc; bestState = ost; bestPacked = op; }
    }
    int maxCol = bestCol, maxState = bestState, maxScoreOff = bestScore, maxPacked = bestPacked;
    int fail = 0;
    long long iterations;
    if (LIMITED) {
        iterations = itAll;
        if (brokeAll || bestCol < 0) { maxCol = 1; maxState = 0; maxScoreOff = BADOFF; maxPacked = BADOFF; }
        else if (bestScore == subfloor && fv > 1) { maxCol = fv - 1; maxState = 0; maxPacked = subfloor; }  // (rows,colStart-1) was set to subfloor and is scanned first
        fail = (maxScoreOff < minScore_off) ? 1 : 0;
    } else {
        iterations = (long long)rows * cols;
    }
    const bool javaMode = (T.flags & (BBM_TF_RAW_LIMITED | BBM_TF_RAW_UNLIMITED)) == 0;

    if (BAND && missAll) {
        // the banded right-edge assumption was violated: hand the alignment to the generic kernel
        if (lane == 0) {
            const unsigned k = atomicAdd(P.overflow_count, 1u);
            P.overflow_list[k] = (int)taskId;
            out->status = 1;   // pending
        }
        return;
    }

    if (lane == 0) {
        out->path = LIMITED ? 0 : 1;
        out->iterations = iterations;
        out->status = 0;
        out->score_len = 0;
        out->match_len = -1;
        out->pad_ = 0;
        if (fail && javaMode) { out->result[0] = rows; out->result[1] = 0; out->result[2] = 0; out->result[3] = 0; out->result[4] = 1; }
        else {
            out->result[0] = rows; out->result[1] = maxCol; out->result[2] = maxState;
            out->result[3] = fail ? maxScoreOff : (maxScoreOff >> TBITS); out->result[4] = fail;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) out->score[k] = 0;
    }
    if (fail || (T.flags & (BBM_TF_SCORE | BBM_TF_TRACEBACK)) == 0) return;

    // ---- score2 + traceback2: one walk over the predecessor codes ----
    const bool wantTb = (T.flags & BBM_TF_TRACEBACK) != 0 && P.match_buf != nullptr;
    int8_t* mslot = nullptr; long long mcap = 0;
    if (wantTb) { mslot = P.match_buf + P.match_off[taskId]; mcap = P.match_off[taskId + 1] - P.match_off[taskId]; }
    int nOps = 0, gapsSeen = 0;
    if (lane == 0) {
        int row = rows, col = maxCol, state = maxState;
        const int bestRefStop = T.a + col - 1;
        int stateTime = 0;
        // ops are staged backwards at the end of the slot, then moved to the front by the whole warp
        while (row > 0 && col > 0) {
            const int L = (col - 1) / W, j = (col - 1) - L * W;
            const tbw w = tb[(size_t)(row + L) * 32 + L];
            const unsigned code = (unsigned)(w >> (4 * j)) & 15u;
            int prev;
            char op;
            if (state == ST_MS) {
                prev = code & 3u;
                if (wantTb) {
                    const int c = ws.read[row - 1], rf = ref[col - 1];
                    op = (c == rf) ? 'm' : ((!base_defined(c) || !base_defined(rf)) ? 'N' : 'S');
                }
                row--; col--;
            } else if (state == ST_DEL) {
                prev = ((code >> 2) & 1u) ? ST_DEL : ST_MS;
                if (wantTb) { const int rf = ref[col - 1]; if (rf == '-') { op = '-'; gapsSeen++; } else op = 'D'; }
                col--;
            } else {
                prev = ((code >> 3) & 1u) ? ST_INS : ST_MS;
                if (wantTb) op = (col == 0) ? 'X' : ((col >= cols) ? 'Y' : 'I');
                row--;
            }
            if (wantTb) { if (nOps < mcap) mslot[mcap - 1 - nOps] = op; }
            nOps++;
            stateTime = (state == prev) ? stateTime + 1 : 0;
            state = prev;
        }
        const int rowEnd = row, colEnd = col;
        if (wantTb && colEnd != rowEnd) {
            int rr = rowEnd;
            while (rr > 0) { if (nOps < mcap) mslot[mcap - 1 - nOps] = 'X'; nOps++; rr--; }
        }
        if (T.flags & BBM_TF_SCORE) {
            int colf = colEnd;
            if (rowEnd > colEnd) colf -= rowEnd;
            const int bestRefStart = T.a + colf;
            int padLeft = 0, padRight = 0;
            if (bestRefStart < T.a) padLeft = imax(0, T.a - bestRefStart);
            else if (bestRefStart == T.a && state == ST_INS) padLeft = stateTime;
            if (bestRefStop > score_ref_end(T)) padRight = imax(0, bestRefStop - score_ref_end(T));
            else if (bestRefStop == score_ref_end(T) && maxState == ST_INS) padRight = maxPacked & TMASK;
            out->score[0] = maxScoreOff >> TBITS; out->score[1] = bestRefStart; out->score[2] = bestRefStop;
            out->score[3] = rows; out->score[4] = maxCol; out->score[5] = maxState;
            out->score[6] = padLeft; out->score[7] = padRight;
            out->score_len = (padLeft > 0 || padRight > 0) ? 8 : 6;
        }
    }
    if (!wantTb) return;
    nOps = __shfl_sync(FULL, nOps, 0);
    gapsSeen = __shfl_sync(FULL, gapsSeen, 0);
    __syncwarp();
    if (nOps > mcap) { if (lane == 0) { out->status = BBM_E_CAPACITY; out->match_len = -1; } return; }
    if (gapsSeen == 0) {
        // the staged string is already in forward order at [mcap-nOps, mcap); slide it to the slot start
        const long long shift = mcap - nOps;
        if (shift > 0) {
            for (int base = 0; base < nOps; base += 32) {
                const int i = base + lane;
                int8_t v = 0;
                if (i < nOps) v = mslot[shift + i];
                __syncwarp();
                if (i < nOps) mslot[i] = v;
                __syncwarp();
            }
        }
        if (lane == 0) out->match_len = nOps;
    } else {
        // gapped reference: every '-' expands to GAPLEN 'D' (…JNI.java:478-493).  Rare; done by one lane.
        if (lane == 0) {
            const long long total = (long long)nOps + (long long)gapsSeen * 127;
            if (total > mcap) { out->status = BBM_E_CAPACITY; out->match_len = -1; }
            else {
                // expand from the back so the staged ops (at the end of the slot) are not overwritten before use:
                // first move ops to a compact front area in reverse? simpler: two passes through a forward copy.
                long long src = mcap - nOps;     // forward order begins here
                // pass 1: move forward ops to the very end (already there). pass 2: write expanded string from the front.
                // Since expansion only grows, writing position j never passes reading position src+i when total<=mcap
                // only if j <= src+i; guarantee by checking, otherwise report capacity.
                long long j = 0; bool okc = true;
                for (int i = 0; i < nOps && okc; ++i) {
                    const int8_t c = mslot[src + i];
                    if (c != '-') { if (j > src + i) okc = false; else mslot[j++] = c; }
                    else { if (j + 128 > src + i + 1) okc = false; else { for (int k = 0; k < 128; ++k) mslot[j++] = 'D'; } }
                }
                if (okc) out->match_len = (int)total; else { out->status = BBM_E_CAPACITY; out->match_len = -1; }
            }
        }
    }
}

}  // namespace bbm
