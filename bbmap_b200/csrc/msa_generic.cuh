// msa_generic.cuh — row-sequential (one thread per alignment) MultiStateAligner11ts for the shapes and cases the
// register-tiled kernel does not take: windows wider than 512 columns, reads longer than MAXR, and banded fills whose
// right-edge assumption failed.  It walks the rows and columns in the reference's own order
// (jni/MultiStateAligner11tsJNI.c:440-670 limited, :134-290 unlimited) on two rolling rows per state, so it is exact by
// construction, and records the same 4-bit predecessor codes as the tiled kernel for score2/traceback2
// (MultiStateAligner11tsJNI.java:376-495, 537-658).  Throughput is not the point of this kernel.
#pragma once
#include "msa_common.cuh"

namespace bbm {

struct TaskCtx;

__host__ __device__ inline long long msa_generic_scratch_ints(int rows, int cols) {
    const long long rowbuf = 2LL * 3 * (cols + 2);
    const long long lim = (rows + 2) + (cols + 2);
    const long long tbBytes = (long long)(rows + 1) * (cols + 1);
    return rowbuf + lim + (tbBytes + 3) / 4 + 8;
}

// `fast` (optional): a buffer of msa_generic_fast_ints(rows, cols) ints for the rolling rows and the two limit vectors — the
// shared-memory variant of the kernel passes dynamic shared memory here, which takes the ~10 dependent global round trips per cell
// (one thread, L2 latency) out of the loop; the predecessor codes always live in `scratch`.
__host__ __device__ inline long long msa_generic_fast_ints(int rows, int cols) { return 2LL * 3 * (cols + 2) + (rows + 2) + (cols + 2); }

__device__ void msa_generic_task(const MsaParams& P, const TaskCtx& T, const bbm_msa_task& task, long long taskId,
                                 int* scratch, long long scratchInts, bbm_msa_out* out, int* fast = nullptr, int lane = -1) {
    const int rows = T.rows, cols = T.cols;
    const bool limited = T.limited != 0;
    const int8_t* __restrict__ read = P.reads + task.read_off;
    const int8_t* __restrict__ ref = P.refs + task.ref_off + T.a;
    if (msa_generic_scratch_ints(rows, cols) > scratchInts) { out->status = BBM_E_SHAPE; return; }
    const int stride = cols + 2;
    int* rowbuf = fast ? fast : scratch;             // [2][3][stride]
    int* vl = rowbuf + 2 * 3 * stride;               // [rows+2]
    int* hl = vl + rows + 2;                         // [cols+2]
    unsigned char* tb = reinterpret_cast<unsigned char*>(scratch + 2 * 3 * stride + (rows + 2) + (cols + 2));   // [(rows+1)*(cols+1)]
    const int tbStride = cols + 1;

    const int maxGain = (rows - 1) * P_MATCH2 + P_MATCH;
    const int minScore_off = (int)((unsigned)T.minScore << TBITS);
    const int floor_ = limited ? minScore_off - maxGain : 0;
    const int subfloor = limited ? floor_ - 5 * P_MATCH2 : 0 - 2 * maxGain;
    const int hb = T.halfband;
    // lane >= 0: called by all 32 lanes of a warp that share `fast`.  Un-banded limited fills are then evaluated 32 columns at a time
    // (MS and INS of a row only depend on the previous row; DEL is a left-to-right chain resolved lane by lane with shuffles); every
    // other case runs on lane 0 exactly as in the single-thread form.
    const bool warpFill = lane >= 0 && fast != nullptr && (!limited || hb < 1) && P.dump == nullptr;
    if (lane > 0 && !warpFill) return;
    const bool lead = lane <= 0;

    if (limited && lead) {
        vl[rows] = minScore_off;
        bool pd = false;
        for (int i = rows - 1; i >= 0; --i) {
            const bool d = base_defined(read[i]);
            vl[i] = max(vl[i + 1] - (d ? (pd ? P_MATCH2 : P_MATCH) : 0), floor_);
            pd = d;
        }
        hl[cols] = minScore_off;
        pd = false;
        for (int i = cols - 1; i >= 0; --i) {
            const int c = ref[i];
            const bool d = base_defined(c);
            hl[i] = max(hl[i + 1] - (d ? (pd ? P_MATCH2 : P_MATCH) : ((pd && c == '-') ? P_DEL : 0)), floor_);
            pd = d;
        }
    }
    // row 0: all zero
    if (warpFill) { for (int i = lane; i < 3 * stride; i += 32) rowbuf[i] = 0; __syncwarp(); }
    else for (int s = 0; s < 3; ++s) for (int c = 0; c <= cols + 1; ++c) rowbuf[s * stride + c] = 0;

    int minGood = 1, maxGood = cols;
    long long iters = 0;
    bool broke = false;
    int lastRowLo = 1, lastRowHi = 0;       // visited interval of the final row (limited)
    if (warpFill) {
        constexpr unsigned WFULL = 0xffffffffu;
        for (int row = 1; row <= rows; ++row) {
            int* up = rowbuf + ((row - 1) & 1) * 3 * stride;
            int* cur = rowbuf + (row & 1) * 3 * stride;
            const int *uM = up, *uD = up + stride, *uI = up + 2 * stride;
            int *cM = cur, *cD = cur + stride, *cI = cur + 2 * stride;
            const int col0 = ins_score_offset(row);
            const int colStart = limited ? minGood : 1, colStop = limited ? maxGood : cols;      // an unlimited fill visits every cell (jni/...JNI.c:134-290)
            minGood = -1; maxGood = -2;
            if (colStart < 0 || colStop < colStart) { broke = true; break; }
            __syncwarp();
            if (lane == 0) {
                cM[0] = col0; cD[0] = col0; cI[0] = col0;
                if (colStart > 1) { cM[colStart - 1] = subfloor; cD[colStart - 1] = subfloor; cI[colStart - 1] = subfloor; }
            }
            __syncwarp();
            const int vlimit = vl[row];
            const int call1 = read[row - 1], call0 = row < 2 ? '?' : read[row - 2];
            const bool delBar = (row < 3) || (row > rows - 3);
            // the reference overwrites (row-1, col+1) with subfloor whenever it walks past colStop (jni/...JNI.c:662-667): beyond colStop the
            // previous row reads as subfloor (row 1 sits on the all-zero row 0 and colStop == cols there)
            const bool pastIsSub = limited && row > 1;
            int carryM = cM[colStart - 1], carryD = cD[colStart - 1];
            int rowHi = colStart - 1;
            bool stop = false;
            for (int c0 = colStart; c0 <= cols && !stop; c0 += 32) {
                const int col = c0 + lane;
                const bool valid = col <= cols;
                const int nIn = min(32, cols - c0 + 1);
                int msv = subfloor, insv = subfloor, delv = subfloor;
                unsigned code = 0; bool good = false;
                int limit = 0, delNeeded = 0, insNeeded = 0, insPen = 0, r1 = 0; bool gap = false;
                if (valid) {
                    r1 = ref[col - 1];
                    const int r0 = col < 2 ? '!' : ref[col - 2];
                    gap = (r1 == '-');
                    const bool match = (call1 == r1 && r1 != 'N'), prevMatch = (call0 == r0 && r0 != 'N');
                    int limit3 = 0, delPen = 0;
                    if (limited) {
                        limit = max(vlimit, hl[col]);
                        limit3 = max(floor_, match ? limit - P_MATCH2 : limit - P_SUB3);
                        delNeeded = max(0, row - col - 1);
                        insNeeded = max(0, (rows - row) - (cols - col) - 1);
                        delPen = del_score_offset(delNeeded); insPen = ins_score_offset(insNeeded);
                    }
                    const bool dSub = pastIsSub && (col - 1) > colStop, uSub = pastIsSub && col > colStop;
                    {   // MS
                        const int dm = dSub ? subfloor : uM[col - 1], dd = dSub ? subfloor : uD[col - 1], di = dSub ? subfloor : uI[col - 1];
                        const int sM = dm & SMASK, sD = dd & SMASK, sI = di & SMASK, streak = dm & TMASK;
                        if (gap || (limited && sM <= limit3 && sD <= limit3 && sI <= limit3)) msv = subfloor;
                        else {
                            int a_, o;
                            if (match) { a_ = sM + (prevMatch ? P_MATCH2 : P_MATCH); o = P_MATCH; }
                            else {
                                a_ = sM + ((r1 != 'N' && call1 != 'N') ? (prevMatch ? (streak <= 1 ? P_SUBR : P_SUB)
                                           : (streak == 0 ? P_SUB : (streak < 5 ? P_SUB2 : P_SUB3))) : 0);
                                o = P_SUB;
                            }
                            const int b_ = sD + o, c_ = sI + o;
                            int score, time;
                            if (a_ >= b_ && a_ >= c_) { score = a_; time = (match == prevMatch) ? streak + 1 : 1; }
                            else if (b_ >= c_) { score = b_; time = 1; }
                            else { score = c_; time = 1; }
                            if (limited) {
                                const int lim2 = delNeeded > 0 ? limit - delPen : (insNeeded > 0 ? limit - insPen : limit);
                                if (score >= lim2) good = true; else score = subfloor;
                            }
                            if (time > MAX_TIME) time = TIME_WRAP;
                            msv = score | time;
                            code |= (time > 1) ? 0u : ((sM >= sD && sM >= sI) ? 0u : (sD >= sI ? 1u : 2u));
                        }
                    }
                    {   // INS
                        const int um = uSub ? subfloor : uM[col], ui = uSub ? subfloor : uI[col];
                        const int sM = um & SMASK, sI = ui & SMASK, streak = ui & TMASK;
                        if (gap || (limited && sM <= limit && sI <= limit) || (row < 2 && col > 1) || (row > rows - 2 && col < cols - 1)) insv = subfloor;
                        else {
                            const int a_ = sM + P_INS;
                            const int b_ = sI + (streak == 0 ? P_INS : (streak < LIM3 ? P_INS2 : (streak < LIM4 ? P_INS3 : P_INS4)));
                            int score, time;
                            if (a_ >= b_) { score = a_; time = 1; } else { score = b_; time = streak + 1; }
                            if (limited) {
                                const int lim2 = delNeeded > 0 ? limit - delPen
                                               : (insNeeded > 0 ? limit - ins_score_offset(time + insNeeded) + ins_score_offset(time) : limit);
                                if (score >= lim2) good = true; else score = subfloor;
                            }
                            if (time > MAX_TIME) time = TIME_WRAP;
                            insv = score | time;
                            code |= ((time > 1) ? 1u : (sM >= sI ? 0u : 1u)) << 3;
                        }
                    }
                }
                // DEL: (row, col) from (row, col-1) — MS of the left neighbour is known, DEL is the chain
                int msLeft = __shfl_up_sync(WFULL, msv, 1);
                if (lane == 0) msLeft = carryM;
                // a DEL cell is subfloor outright when neither predecessor beats its limit; with a dead chain only the lanes whose left MS
                // could open a deletion need a step, and a skipped lane leaves exactly `subfloor` (time 0) behind, as the reference writes it
                if (!limited) {
                    // unlimited fill: no limit tests, so the chain step is the same few instructions for every column — run it uniformly
                    // (lane j's inputs broadcast, every lane follows the carry) instead of one divergent lane at a time
                    const int adjv = (r1 == 'N') ? P_DEL_REF_N : (gap ? P_GAP : 0);
                    const int aOpen = (msLeft & SMASK) + P_DEL + adjv;
                    auto step = [&](int j) {
                        const int aj = __shfl_sync(WFULL, aOpen, j), adjj = __shfl_sync(WFULL, adjv, j);
                        int nv = subfloor; unsigned cbit = 0;
                        if (!delBar) {
                            const int sD = carryD & SMASK, streak = carryD & TMASK;
                            const int b_ = sD + adjj + (streak == 0 ? P_DEL : (streak < LIM3 ? P_DEL2 : (streak < LIM4 ? P_DEL3 : (streak < LIM5 ? P_DEL4 :
                                                       (((streak & 3) == 0) ? P_DEL5 : 0)))));
                            const bool msw = aj >= b_;
                            int time = msw ? 1 : streak + 1;
                            if (time > MAX_TIME) time = TIME_WRAP;
                            nv = (msw ? aj : b_) | time;
                            cbit = ((time > 1) ? 1u : ((aj - P_DEL - adjj) >= sD ? 0u : 1u)) << 2;
                        }
                        if (lane == j) { delv = nv; code |= cbit; }
                        carryD = nv;
                                        };
                    if (nIn == 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) step(j);     // full chunk: unrolled, so the two broadcasts of a step issue ahead of the chain
                    } else for (int j = 0; j < nIn; ++j) step(j);
                }
                const unsigned openMask = limited ? __ballot_sync(WFULL, valid && !delBar && (msLeft & SMASK) > limit) : 0u;
                for (int j = 0; limited && j < nIn; ++j) {
                    if ((carryD & SMASK) == subfloor && !((openMask >> j) & 1u)) {
                        const unsigned m = openMask & (0xfffffffeu << j);
                        if (!m) { carryD = subfloor; break; }
                        j = __ffs(m) - 1;
                        carryD = subfloor;
                    }
                    if (lane == j) {
                        const int sM = msLeft & SMASK, sD = carryD & SMASK, streak = carryD & TMASK;
                        if ((limited && sM <= limit && sD <= limit) || delBar) delv = subfloor;
                        else {
                            int a_ = sM + P_DEL;
                            int b_ = sD + (streak == 0 ? P_DEL : (streak < LIM3 ? P_DEL2 : (streak < LIM4 ? P_DEL3 : (streak < LIM5 ? P_DEL4 :
                                           (((streak & 3) == 0) ? P_DEL5 : 0)))));
                            if (r1 == 'N') { a_ += P_DEL_REF_N; b_ += P_DEL_REF_N; } else if (gap) { a_ += P_GAP; b_ += P_GAP; }
                            int score, time;
                            if (a_ >= b_) { score = a_; time = 1; } else { score = b_; time = streak + 1; }
                            if (limited) {
                                const int lim2 = insNeeded > 0 ? limit - insPen
                                               : (delNeeded > 0 ? limit - del_score_offset(time + delNeeded) + del_score_offset(time) : limit);
                                if (score >= lim2) good = true; else score = subfloor;
                            }
                            if (time > MAX_TIME) time = TIME_WRAP;
                            delv = score | time;
                            code |= ((time > 1) ? 1u : (sM >= sD ? 0u : 1u)) << 2;
                        }
                    }
                    carryD = __shfl_sync(WFULL, delv, j);
                }
                carryM = __shfl_sync(WFULL, msv, 31);
                // the reference leaves the row at the first cell past colStop that is not good (the cell itself is evaluated and stored)
                const unsigned term = __ballot_sync(WFULL, limited && valid && col > colStop && !good);
                int lastLane = nIn - 1;
                if (term) { lastLane = __ffs(term) - 1; stop = true; }
                const bool visited = valid && lane <= lastLane;
                if (visited) {
                    cM[col] = msv; cD[col] = delv; cI[col] = insv;
                    tb[(long long)row * tbStride + col] = (unsigned char)code;
                }
                const unsigned gm = __ballot_sync(WFULL, visited && good);
                if (gm) { if (minGood < 0) minGood = c0 + __ffs(gm) - 1; maxGood = c0 + 31 - __clz(gm); }
                iters += lastLane + 1;
                rowHi = c0 + lastLane;
            }
            if (row == rows) { lastRowLo = colStart; lastRowHi = rowHi; }
        }
        __syncwarp();
        if (lane != 0) return;
    } else
    for (int row = 1; row <= rows; ++row) {
        int* up = rowbuf + ((row - 1) & 1) * 3 * stride;
        int* cur = rowbuf + (row & 1) * 3 * stride;
        int *uM = up, *uD = up + stride, *uI = up + 2 * stride;
        int *cM = cur, *cD = cur + stride, *cI = cur + 2 * stride;
        const int col0 = ins_score_offset(row);
        cM[0] = col0; cD[0] = col0; cI[0] = col0;
        int colStart = 1, colStop = cols;
        if (limited) {
            colStart = hb < 1 ? minGood : max(minGood, row - hb);
            colStop = hb < 1 ? maxGood : min(maxGood, row + hb * 2 - 1);
            minGood = -1; maxGood = -2;
            if (colStart < 0 || colStop < colStart) { broke = true; break; }
            if (colStart > 1) { cM[colStart - 1] = subfloor; cD[colStart - 1] = subfloor; cI[colStart - 1] = subfloor; }
        }
        const int vlimit = limited ? vl[row] : 0;
        const int call1 = read[row - 1], call0 = row < 2 ? '?' : read[row - 2];
        const bool delBar = (row < 3) || (row > rows - 3);
        int col = colStart;
        for (; col <= cols; ++col) {
            const int r1 = ref[col - 1];
            const int r0 = col < 2 ? '!' : ref[col - 2];
            const bool gap = (r1 == '-'), match = (call1 == r1 && r1 != 'N'), prevMatch = (call0 == r0 && r0 != 'N');
            iters++;
            int limit = 0, limit3 = 0, delNeeded = 0, insNeeded = 0, delPen = 0, insPen = 0;
            if (limited) {
                limit = max(vlimit, hl[col]);
                limit3 = max(floor_, match ? limit - P_MATCH2 : limit - P_SUB3);
                delNeeded = max(0, row - col - 1);
                insNeeded = max(0, (rows - row) - (cols - col) - 1);
                delPen = del_score_offset(delNeeded); insPen = ins_score_offset(insNeeded);
            }
            unsigned code = 0;
            bool good = false;
            {   // MS
                const int dm = uM[col - 1], sM = dm & SMASK, sD = uD[col - 1] & SMASK, sI = uI[col - 1] & SMASK, streak = dm & TMASK;
                if (gap || (limited && sM <= limit3 && sD <= limit3 && sI <= limit3)) cM[col] = subfloor;
                else {
                    int a_, o;
                    if (match) { a_ = sM + (prevMatch ? P_MATCH2 : P_MATCH); o = P_MATCH; }
                    else {
                        a_ = sM + ((r1 != 'N' && call1 != 'N') ? (prevMatch ? (streak <= 1 ? P_SUBR : P_SUB)
                                   : (streak == 0 ? P_SUB : (streak < 5 ? P_SUB2 : P_SUB3))) : 0);
                        o = P_SUB;
                    }
                    const int b_ = sD + o, c_ = sI + o;
                    int score, time;
                    if (a_ >= b_ && a_ >= c_) { score = a_; time = (match == prevMatch) ? streak + 1 : 1; }
                    else if (b_ >= c_) { score = b_; time = 1; }
                    else { score = c_; time = 1; }
                    if (limited) {
                        const int lim2 = delNeeded > 0 ? limit - delPen : (insNeeded > 0 ? limit - insPen : limit);
                        if (score >= lim2) good = true; else score = subfloor;
                    }
                    if (time > MAX_TIME) time = TIME_WRAP;
                    cM[col] = score | time;
                    code |= (time > 1) ? 0u : ((sM >= sD && sM >= sI) ? 0u : (sD >= sI ? 1u : 2u));
                }
            }
            {   // DEL
                const int lm = cM[col - 1], ld = cD[col - 1], sM = lm & SMASK, sD = ld & SMASK, streak = ld & TMASK;
                if ((limited && sM <= limit && sD <= limit) || delBar) cD[col] = subfloor;
                else {
                    int a_ = sM + P_DEL;
                    int b_ = sD + (streak == 0 ? P_DEL : (streak < LIM3 ? P_DEL2 : (streak < LIM4 ? P_DEL3 : (streak < LIM5 ? P_DEL4 :
                                   (((streak & 3) == 0) ? P_DEL5 : 0)))));
                    if (r1 == 'N') { a_ += P_DEL_REF_N; b_ += P_DEL_REF_N; } else if (gap) { a_ += P_GAP; b_ += P_GAP; }
                    int score, time;
                    if (a_ >= b_) { score = a_; time = 1; } else { score = b_; time = streak + 1; }
                    if (limited) {
                        const int lim2 = insNeeded > 0 ? limit - insPen
                                       : (delNeeded > 0 ? limit - del_score_offset(time + delNeeded) + del_score_offset(time) : limit);
                        if (score >= lim2) good = true; else score = subfloor;
                    }
                    if (time > MAX_TIME) time = TIME_WRAP;
                    cD[col] = score | time;
                    code |= ((time > 1) ? 1u : (sM >= sD ? 0u : 1u)) << 2;
                }
            }
            {   // INS
                const int um = uM[col], ui = uI[col], sM = um & SMASK, sI = ui & SMASK, streak = ui & TMASK;
                if (gap || (limited && sM <= limit && sI <= limit) || (row < 2 && col > 1) || (row > rows - 2 && col < cols - 1)) cI[col] = subfloor;
                else {
                    const int a_ = sM + P_INS;
                    const int b_ = sI + (streak == 0 ? P_INS : (streak < LIM3 ? P_INS2 : (streak < LIM4 ? P_INS3 : P_INS4)));
                    int score, time;
                    if (a_ >= b_) { score = a_; time = 1; } else { score = b_; time = streak + 1; }
                    if (limited) {
                        const int lim2 = delNeeded > 0 ? limit - delPen
                                       : (insNeeded > 0 ? limit - ins_score_offset(time + insNeeded) + ins_score_offset(time) : limit);
                        if (score >= lim2) good = true; else score = subfloor;
                    }
                    if (time > MAX_TIME) time = TIME_WRAP;
                    cI[col] = score | time;
                    code |= ((time > 1) ? 1u : (sM >= sI ? 0u : 1u)) << 3;
                }
            }
            tb[(long long)row * tbStride + col] = (unsigned char)code;
            if (P.dump) {   // dense dump [3][rows+1][cols+2] for the single-alignment twins (capi.cu replays the reference's writes)
                const long long plane = (long long)(rows + 1) * (cols + 2), idx = (long long)row * (cols + 2) + col;
                P.dump[idx] = cM[col]; P.dump[plane + idx] = cD[col]; P.dump[2 * plane + idx] = cI[col];
            }
            if (limited) {
                if (good) { maxGood = col; if (minGood < 0) minGood = col; }
                if (col >= colStop) {
                    if (col > colStop && (maxGood < col || hb > 0)) { break; }
                    if (row > 1) { uM[col + 1] = subfloor; uD[col + 1] = subfloor; uI[col + 1] = subfloor; }
                }
            }
        }
        if (row == rows) { lastRowLo = colStart; lastRowHi = min(col, cols); }
    }

    // final scan (jni/...JNI.c:672-686): unvisited last-row cells are BADoff, (rows,colStart-1) is subfloor
    int maxCol = -1, maxState = -1, maxScore = INT_MIN, maxPacked = 0;
    {
        const int* last = rowbuf + (rows & 1) * 3 * stride;
        for (int st = 0; st < 3; ++st)
            for (int c = 1; c <= cols; ++c) {
                int v;
                if (!limited) v = last[st * stride + c];
                else if (broke) v = BADOFF;
                else if (c >= lastRowLo && c <= lastRowHi) v = last[st * stride + c];
                else if (c == lastRowLo - 1) v = subfloor;
                else v = BADOFF;
                const int x = v & SMASK;
                if (x > maxScore) { maxScore = x; maxCol = c; maxState = st; maxPacked = v; }
            }
    }
    int fail = 0;
    if (limited && maxScore < minScore_off) fail = 1;
    const bool javaMode = (T.flags & (BBM_TF_RAW_LIMITED | BBM_TF_RAW_UNLIMITED)) == 0;
    out->path = limited ? 0 : 1;
    out->iterations = limited ? iters : (long long)rows * cols;
    out->status = 0; out->score_len = 0; out->match_len = -1; out->pad_ = 0;
    for (int k = 0; k < 8; ++k) out->score[k] = 0;
    if (fail && javaMode) { out->result[0] = rows; out->result[1] = 0; out->result[2] = 0; out->result[3] = 0; out->result[4] = 1; }
    else {
        out->result[0] = rows; out->result[1] = maxCol; out->result[2] = maxState;
        out->result[3] = fail ? maxScore : (maxScore >> TBITS); out->result[4] = fail;
    }
    if (fail || (T.flags & (BBM_TF_SCORE | BBM_TF_TRACEBACK)) == 0) return;

    const bool wantTb = (T.flags & BBM_TF_TRACEBACK) != 0 && P.match_buf != nullptr;
    int8_t* mslot = nullptr; long long mcap = 0;
    if (wantTb) { mslot = P.match_buf + P.match_off[taskId]; mcap = P.match_off[taskId + 1] - P.match_off[taskId]; }
    int row = rows, col = maxCol, state = maxState, stateTime = 0, nOps = 0, gapsSeen = 0;
    const int bestRefStop = T.a + col - 1;
    while (row > 0 && col > 0) {
        const unsigned code = tb[(long long)row * tbStride + col];
        int prev; char op = 0;
        if (state == ST_MS) {
            prev = code & 3u;
            const int c = read[row - 1], rf = ref[col - 1];
            op = (c == rf) ? 'm' : ((!base_defined(c) || !base_defined(rf)) ? 'N' : 'S');
            row--; col--;
        } else if (state == ST_DEL) {
            prev = ((code >> 2) & 1u) ? ST_DEL : ST_MS;
            if (ref[col - 1] == '-') { op = '-'; gapsSeen++; } else op = 'D';
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
        if (bestRefStart < T.a) padLeft = max(0, T.a - bestRefStart);
        else if (bestRefStart == T.a && state == ST_INS) padLeft = stateTime;
        if (bestRefStop > score_ref_end(T)) padRight = max(0, bestRefStop - score_ref_end(T));
        else if (bestRefStop == score_ref_end(T) && maxState == ST_INS) padRight = maxPacked & TMASK;
        out->score[0] = maxScore >> TBITS; out->score[1] = bestRefStart; out->score[2] = bestRefStop;
        out->score[3] = rows; out->score[4] = maxCol; out->score[5] = maxState; out->score[6] = padLeft; out->score[7] = padRight;
        out->score_len = (padLeft > 0 || padRight > 0) ? 8 : 6;
    }
    if (!wantTb) return;
    const long long total = (long long)nOps + (long long)gapsSeen * 127;
    if (nOps > mcap || total > mcap) { out->status = BBM_E_CAPACITY; out->match_len = -1; return; }
    const long long src = mcap - nOps;
    long long j = 0;
    for (int i = 0; i < nOps; ++i) {
        const int8_t c = mslot[src + i];
        if (c != '-') mslot[j++] = c; else for (int k = 0; k < 128; ++k) mslot[j++] = 'D';
    }
    out->match_len = (int)total;
}

}  // namespace bbm
