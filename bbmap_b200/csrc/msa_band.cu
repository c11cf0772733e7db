// msa_band.cu — banded fillLimitedX (halfband > 0), one THREAD per alignment, the band held by diagonal in shared memory.
//
// Reference: jni/MultiStateAligner11tsJNI.c:361-704 with `halfband` = max(min(bandwidth, 8 + rows*bandwidthRatio), columns-rows+8)/2 (:392-393):
// row `row` visits columns colStart = max(minGoodCol, row-halfband) .. min(colStop, columns-1)+1 with colStop = min(maxGoodCol, row+2*halfband-1)
// (:441-442, 660-661), i.e. never more than 3*halfband+1 cells around the diagonal — 25 for the default-padded 150-bp window, 61 with bandwidth 40 —
// while the register-tiled kernel evaluates the whole rectangle and must redo an alignment in the row-sequential kernel whenever its right-edge
// assumption fails (9-20 % of them on the configs[2] band sweep, which is why that sweep ran at 5-7 GCUPS).
//
// Here a thread walks its alignment in the reference's own row and column order.  The three states of the cells it can reach live in ONE rolling row
// indexed by diagonal, slot(row, col) = col - row + halfband + 1, laid out [state][slot][thread] in shared memory (bank = thread): moving to the next
// row moves every cell one slot down, so the diagonal predecessor of a cell is the previous content of its own slot, the upper predecessor is the
// content of the next slot, and the left predecessor stays in registers — the in-place scheme of msa_narrow.cuh for a band of any width up to 128
// slots.  What the reference writes explicitly around the visited interval — `subfloor` into (row, colStart-1) and into (row-1, col+1) when it walks
// past colStop (:451-456, 662-667) — is written into the corresponding slots, so every cell read is either computed in the row it is read from or
// `subfloor`, exactly as in the reference's matrix.  The cell itself is the branch-free msa_cell<true> of the tiled and strip kernels.  4-bit
// predecessor codes go to a per-thread scratch block (one byte per band cell) and the same thread runs score2 / traceback2 over them afterwards
// (MultiStateAligner11tsJNI.java:376-495, 537-658).
#include <climits>
#include "msa_kernels.cuh"

namespace bbm {

constexpr int BAND_THREADS = 64;
constexpr int BAND_MAX_ND = 128;              // slots: 3*halfband+3 <= 128  (halfband <= 41)

struct BandParams {
    MsaParams P;
    const int* list; const unsigned int* endPtr; unsigned int base; int nlist;
    int nd, maxRows, maxCols;                 // slots per row (>= 3*halfband+3 of every task), largest shape in the list
    int pen, delcN;                           // sizes of the penalty tables in shared memory (cell_tables_init_dyn)
    int* lim;                                 // [threads][maxRows + maxCols + 8]  vertLimit / horizLimit of the thread's current alignment
    unsigned char* tb;                        // [threads][(maxRows + 1) * nd]     predecessor codes by (row, slot)
    unsigned int* counter;
};

__global__ void __launch_bounds__(BAND_THREADS) msa_band_kernel(BandParams S) {
    extern __shared__ int bandBuf[];          // penalty tables sized for this launch's shapes, then [3][nd][BAND_THREADS]
    CellTablesDyn tab;
    cell_tables_init_dyn(bandBuf, S.pen, S.delcN, tab);
    __syncthreads();
    const MsaParams& P = S.P;
    const int tid = threadIdx.x, nd = S.nd;
    const long long gthread = (long long)blockIdx.x * BAND_THREADS + tid;
    int* bM = bandBuf + ((cell_tables_dyn_ints(S.pen, S.delcN) + 31) & ~31) + tid; int* bD = bM + nd * BAND_THREADS; int* bI = bD + nd * BAND_THREADS;
#define SLOT(p, s) (p)[(s) * BAND_THREADS]
    int* vl = S.lim + gthread * (S.maxRows + S.maxCols + 8);
    int* hl = vl + S.maxRows + 4;
    unsigned char* tb = S.tb + gthread * (long long)(S.maxRows + 1) * nd;
    const int n = S.endPtr ? min(S.nlist, (int)(*S.endPtr - S.base)) : S.nlist;
    const int lane = tid & 31;
    for (;;) {
        // the 32 lanes of a warp take 32 consecutive alignments of the list (ordered by slot class and read length) and start them together: same
        // number of rows, bands of similar width, so the row and column loops below stay converged
        __syncwarp();
        unsigned k0 = 0;
        if (lane == 0) k0 = atomicAdd(S.counter, 32u);
        k0 = __shfl_sync(0xffffffffu, k0, 0);
        if (k0 >= (unsigned)n) break;
        const unsigned k = k0 + lane;
        bool active = k < (unsigned)n;                         // lanes without an alignment still take part in the warp votes of the fill loop
        const int id = active ? S.list[k] : 0;
        bbm_msa_task task = {};
        if (active) task = P.tasks[id];
        bbm_msa_out* out = P.outs + id;
        TaskCtx T; T.rows = 0; T.cols = 0; T.a = 0; T.b = 0; T.minScore = 0; T.limited = 0; T.halfband = 0; T.flags = 0;
        if (active) resolve_task(task, P.bandwidth, P.ratio, T);
        const int rows = T.rows, cols = T.cols, hb = T.halfband;
        if (active && (!T.limited || hb < 1 || 3 * hb + 3 > nd || rows > S.maxRows || cols > S.maxCols)) {       // the classifier never sends such a task here
            bbm_msa_out o = {}; o.status = BBM_E_SHAPE; o.match_len = -1; *out = o; active = false;
        }
        const int8_t* __restrict__ read = P.reads + task.read_off;
        const int8_t* __restrict__ ref = P.refs + task.ref_off + T.a;
        const int maxGain = (rows - 1) * P_MATCH2 + P_MATCH;
        const int minScore_off = (int)((unsigned)T.minScore << TBITS);
        CellConst K; K.floor_ = minScore_off - maxGain; K.subfloor = K.floor_ - 5 * P_MATCH2;
        const int subfloor = K.subfloor;
        if (active) {
            // vertLimit / horizLimit (jni/...JNI.c:405-438)
            vl[rows] = minScore_off;
            bool pd = false;
            for (int i = rows - 1; i >= 0; --i) {
                const bool d = base_defined(read[i]);
                vl[i] = imax(vl[i + 1] - (d ? (pd ? P_MATCH2 : P_MATCH) : 0), K.floor_);
                pd = d;
            }
            hl[cols] = minScore_off;
            pd = false;
            for (int i = cols - 1; i >= 0; --i) {
                const int c = ref[i];
                const bool d = base_defined(c);
                hl[i] = imax(hl[i + 1] - (d ? (pd ? P_MATCH2 : P_MATCH) : ((pd && c == '-') ? P_DEL : 0)), K.floor_);
                pd = d;
            }
            for (int s = 0; s < nd; ++s) { SLOT(bM, s) = 0; SLOT(bD, s) = 0; SLOT(bI, s) = 0; }        // row 0 of the matrix is all zero
        }
        int minGood = 1, maxGood = cols, lastLo = 1, lastHi = 0;
        long long iters = 0;
        bool broke = false;
        // The fill as a per-lane state machine: every pass of the loop below first lets the lanes that stand between two rows open their next row
        // (cheap, divergent), then evaluates ONE cell for every lane that has one (the expensive part, converged): a lane whose band is 4 cells wide
        // does not wait at the end of each row for a neighbour whose band is 25 cells wide, it simply gets through its rows sooner.
        int row = 0, col = 0, colStop = 0, off = 0, lM = 0, lD = 0, ref0 = 0;
        int refCur = 0, hlCur = 0;                             // ref[col-1] and horizLimit[col] of the cell about to be evaluated: loaded one cell ahead
        bool inRow = false, fillDone = !active, insTop = false, insBot = false;
        CellRow R; R.call1 = 0; R.call0 = 0; R.callN = false; R.vlimit = 0; R.delBar = false;
        unsigned char* tbr = tb;
        for (;;) {
            if (!inRow && !fillDone) {
                ++row;
                if (row > rows) fillDone = true;
                else {
                    const int colStart = imax(minGood, row - hb);
                    colStop = imin(maxGood, row + hb * 2 - 1);
                    minGood = -1; maxGood = -2;
                    if (colStart < 0 || colStop < colStart) { broke = true; fillDone = true; }
                    else {
                        off = hb + 1 - row;                    // slot(row, col) = col + off;  slot(row-1, col) = col + off + 1
                        if (colStart > 1) {
                            lM = subfloor; lD = subfloor;
                            const int s0 = colStart - 1 + off; // (row, colStart-1) := subfloor  (:451-456)
                            SLOT(bM, s0) = subfloor; SLOT(bD, s0) = subfloor; SLOT(bI, s0) = subfloor;
                        } else { lM = tab.insc[row]; lD = lM; }   // column 0 (…JNI.java:105-111)
                        R.call1 = read[row - 1]; R.call0 = row < 2 ? '?' : read[row - 2];
                        R.callN = (R.call1 == 'N'); R.vlimit = vl[row];
                        R.delBar = (row < 3) || (row > rows - 3);
                        insTop = (row < 2); insBot = (row > rows - 2);
                        tbr = tb + (long long)row * nd;
                        col = colStart;
                        ref0 = col < 2 ? '!' : (ref[col - 2] == 'N' ? 0x100 : (int)ref[col - 2]);
                        refCur = ref[col - 1]; hlCur = hl[col];
                        if (row == rows) lastLo = colStart;
                        inRow = true;
                    }
                }
            }
            if (__all_sync(0xffffffffu, fillDone)) break;
            if (inRow) {
                const int slot = col + off;
                int ref1 = refCur; if (ref1 == 'N') ref1 = 0x100;
                const int hlim = hlCur;
                if (col < cols) { refCur = ref[col]; hlCur = hl[col + 1]; }      // the next cell of the row: its loads overlap this cell's arithmetic
                int dM, dD, dI;
                if (col == 1) { dM = row == 1 ? 0 : tab.insc[row - 1]; dD = dM; dI = dM; }        // (row-1, 0)
                else { dM = SLOT(bM, slot); dD = SLOT(bD, slot); dI = SLOT(bI, slot); }
                const int uM = SLOT(bM, slot + 1), uI = SLOT(bI, slot + 1);
                const int delNeeded = imax(0, row - col - 1), insNeeded = imax(0, (rows - row) - (cols - col) - 1);
                const bool insBar = (insTop && col > 1) || (insBot && col < cols - 1);
                const CellOut o = msa_cell<true>(K, R, dM, dD, dI, lM, lD, uM, uI, ref1, ref0, insBar, hlim, delNeeded, insNeeded, tab);
                SLOT(bM, slot) = o.ms; SLOT(bD, slot) = o.del; SLOT(bI, slot) = o.ins;
                tbr[slot] = (unsigned char)o.code;
                iters++;
                if (o.good) { maxGood = col; if (minGood < 0) minGood = col; }
                lM = o.ms; lD = o.del; ref0 = ref1;
                bool rowEnds = false;
                if (col >= colStop) {
                    if (col > colStop) rowEnds = true;         // halfband > 0: the cell after colStop is the last one (:660-661)
                    else if (row > 1) { SLOT(bM, slot + 2) = subfloor; SLOT(bD, slot + 2) = subfloor; SLOT(bI, slot + 2) = subfloor; }      // (row-1, col+1) := subfloor (:662-667)
                }
                if (!rowEnds && col + 1 > cols) { rowEnds = true; if (row == rows) lastHi = cols; }
                else if (rowEnds && row == rows) lastHi = col;
                if (rowEnds) inRow = false; else ++col;
            }
        }
        if (!active) continue;
        // final scan (:672-686): unvisited cells of the last row hold BADoff, (rows, colStart-1) holds subfloor; states 0,1,2, columns ascending, strict >
        int maxCol = 1, maxState = 0, maxScore = BADOFF & SMASK, maxPacked = BADOFF;
        if (!broke) {
            const int off = hb + 1 - rows;
            const int c0 = imax(1, lastLo - 1);
            if (c0 > 1) { /* column 1 is unvisited and comes first in scan order: it stays the answer unless something is strictly larger */ }
            else { maxScore = INT_MIN; }                        // column 1 is inside [lastLo-1, lastHi]: no BADoff cell precedes it
            for (int st = 0; st < 3; ++st) {
                const int* p = st == 0 ? bM : (st == 1 ? bD : bI);
                for (int c = c0; c <= lastHi; ++c) {
                    const int v = (c == lastLo - 1) ? subfloor : SLOT(p, c + off);
                    const int x = v & SMASK;
                    if (x > maxScore) { maxScore = x; maxCol = c; maxState = st; maxPacked = v; }
                }
            }
        }
        const int fail = (maxScore < minScore_off) ? 1 : 0;
        const bool javaMode = (T.flags & (BBM_TF_RAW_LIMITED | BBM_TF_RAW_UNLIMITED)) == 0;
        out->path = 0; out->iterations = iters; out->status = 0; out->score_len = 0; out->match_len = -1; out->pad_ = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) out->score[q] = 0;
        if (fail && javaMode) { out->result[0] = rows; out->result[1] = 0; out->result[2] = 0; out->result[3] = 0; out->result[4] = 1; }
        else {
            out->result[0] = rows; out->result[1] = maxCol; out->result[2] = maxState;
            out->result[3] = fail ? maxScore : (maxScore >> TBITS); out->result[4] = fail;
        }
        if (fail || (T.flags & (BBM_TF_SCORE | BBM_TF_TRACEBACK)) == 0) continue;

        // score2 + traceback2 over the predecessor codes
        const bool wantTb = (T.flags & BBM_TF_TRACEBACK) != 0 && P.match_buf != nullptr;
        int8_t* mslot = nullptr; long long mcap = 0;
        if (wantTb) { mslot = P.match_buf + P.match_off[id]; mcap = P.match_off[id + 1] - P.match_off[id]; }
        int wr = rows, wc = maxCol, state = maxState, stateTime = 0, nOps = 0, gapsSeen = 0;
        const int bestRefStop = T.a + wc - 1;
        while (wr > 0 && wc > 0) {
            const unsigned code = tb[(long long)wr * nd + (wc + hb + 1 - wr)];
            int prev; char op = 0;
            if (state == ST_MS) {
                prev = code & 3u;
                const int c = read[wr - 1], rf = ref[wc - 1];
                op = (c == rf) ? 'm' : ((!base_defined(c) || !base_defined(rf)) ? 'N' : 'S');
                wr--; wc--;
            } else if (state == ST_DEL) {
                prev = ((code >> 2) & 1u) ? ST_DEL : ST_MS;
                if (ref[wc - 1] == '-') { op = '-'; gapsSeen++; } else op = 'D';
                wc--;
            } else {
                prev = ((code >> 3) & 1u) ? ST_INS : ST_MS;
                op = (wc == 0) ? 'X' : ((wc >= cols) ? 'Y' : 'I');
                wr--;
            }
            if (wantTb && nOps < mcap) mslot[mcap - 1 - nOps] = op;
            nOps++;
            stateTime = (state == prev) ? stateTime + 1 : 0;
            state = prev;
        }
        const int rowEnd = wr, colEnd = wc;
        if (wantTb && colEnd != rowEnd) { int rr = rowEnd; while (rr > 0) { if (nOps < mcap) mslot[mcap - 1 - nOps] = 'X'; nOps++; rr--; } }
        if (T.flags & BBM_TF_SCORE) {
            int colf = colEnd; if (rowEnd > colEnd) colf -= rowEnd;
            const int bestRefStart = T.a + colf;
            int padLeft = 0, padRight = 0;
            if (bestRefStart < T.a) padLeft = imax(0, T.a - bestRefStart);
            else if (bestRefStart == T.a && state == ST_INS) padLeft = stateTime;
            if (bestRefStop > score_ref_end(T)) padRight = imax(0, bestRefStop - score_ref_end(T));
            else if (bestRefStop == score_ref_end(T) && maxState == ST_INS) padRight = maxPacked & TMASK;
            out->score[0] = maxScore >> TBITS; out->score[1] = bestRefStart; out->score[2] = bestRefStop;
            out->score[3] = rows; out->score[4] = maxCol; out->score[5] = maxState; out->score[6] = padLeft; out->score[7] = padRight;
            out->score_len = (padLeft > 0 || padRight > 0) ? 8 : 6;
        }
        if (!wantTb) continue;
        const long long total = (long long)nOps + (long long)gapsSeen * 127;
        if (nOps > mcap || total > mcap) { out->status = BBM_E_CAPACITY; out->match_len = -1; continue; }
        const long long src = mcap - nOps;
        long long j = 0;
        for (int i = 0; i < nOps; ++i) {
            const int8_t c = mslot[src + i];
            if (c != '-') mslot[j++] = c; else for (int q = 0; q < 128; ++q) mslot[j++] = 'D';
        }
        out->match_len = (int)total;
    }
#undef SLOT
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_msa_band_threads() { return BAND_THREADS; }
extern "C" int bbm_msa_band_max_nd() { return BAND_MAX_ND; }
static inline int band_pen(int maxRows, int maxCols) { return (maxRows > maxCols ? maxRows : maxCols) + 8; }
static inline int band_delc(int maxRows, int maxCols) { return maxRows + maxCols + 8; }
// dynamic shared memory of one block: the penalty tables for shapes up to maxRows x maxCols + three band rows of nd slots per thread
extern "C" size_t bbm_msa_band_smem_bytes(int maxRows, int maxCols, int nd) {
    const int tab = (cell_tables_dyn_ints(band_pen(maxRows, maxCols), band_delc(maxRows, maxCols)) + 31) & ~31;
    return ((size_t)tab + (size_t)3 * nd * BAND_THREADS) * sizeof(int);
}
extern "C" size_t bbm_msa_band_thread_bytes(int maxRows, int maxCols, int nd) {
    return (size_t)(maxRows + maxCols + 8) * 4 + (size_t)(maxRows + 1) * nd;
}
extern "C" int bbm_launch_msa_band(const MsaParams* P, const int* list, int nlist, const unsigned int* endPtr, unsigned int base, int nd, int maxRows, int maxCols,
                                   void* scratch, unsigned int* counter, int blocks, cudaStream_t st) {
    BandParams S;
    S.P = *P; S.list = list; S.nlist = nlist; S.endPtr = endPtr; S.base = base; S.nd = nd; S.maxRows = maxRows; S.maxCols = maxCols; S.counter = counter;
    const size_t threads = (size_t)blocks * BAND_THREADS;
    S.lim = (int*)scratch;
    S.tb = (unsigned char*)scratch + threads * (size_t)(maxRows + maxCols + 8) * 4;
    S.pen = band_pen(maxRows, maxCols); S.delcN = band_delc(maxRows, maxCols);
    const size_t smem = bbm_msa_band_smem_bytes(maxRows, maxCols, nd);
    // tables + band rows: above the 48 KB default from 64 slots on; the opt-in is per device, so it is simply made on every launch
    cudaError_t ea = cudaFuncSetAttribute(msa_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bbm_msa_band_smem_bytes(MAXR, STRIP_MAX_COLS, BAND_MAX_ND));
    if (ea != cudaSuccess) return (int)ea;
    msa_band_kernel<<<blocks, BAND_THREADS, smem, st>>>(S);
    return (int)cudaGetLastError();
}
