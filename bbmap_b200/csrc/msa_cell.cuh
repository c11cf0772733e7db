// msa_cell.cuh — one cell (row,col) of the MultiStateAligner11ts 3-state recurrence, branch-free.
//
// Reference: jni/MultiStateAligner11tsJNI.c:460-658 (limited) and :137-288 (unlimited); identical arithmetic in
// current/align2/MultiStateAligner11ts.java:177-560.  The reference's if/else ladders are restated as selects so that the
// 32 lanes of a warp — which sit on different rows/columns (tiled kernel) or different alignments (narrow kernel) — never
// diverge on data; the only branches left are the two rare "needs an indel to finish" limit adjustments.
//
// Also produces the 4-bit predecessor code that score2/traceback2 (MultiStateAligner11tsJNI.java:391-447, 573-611) would
// derive from the stored matrix: bits 0-1 = predecessor of the MS state (0 MS, 1 DEL, 2 INS), bit 2 = DEL came from DEL,
// bit 3 = INS came from INS.
#pragma once
#include "msa_common.cuh"

namespace bbm {

struct CellConst {          // per-alignment constants
    int floor_, subfloor;
};

struct CellRow {            // per-(lane,row) values
    int call1, call0;       // read[row-1], read[row-2] (or '?')
    bool callN;             // call1=='N'
    bool delBar;            // row<3 || row>rows-3
    int vlimit;             // vertLimit[row]
};

struct CellOut {
    int ms, del, ins;       // packed score|time
    unsigned code;          // 4-bit predecessor code
    bool good;
};

// Penalty tables live in shared memory: insc[i]=POINTSoff_INS_ARRAY_C[i], delc[i]=calcDelScoreOffset(i); both [0]=0.
template <bool LIMITED, bool CLAMP_TIME>
__device__ __forceinline__ CellOut msa_cell(const CellConst& K, const CellRow& R,
                                            int dMS, int dDEL, int dINS,        // (row-1,col-1)
                                            int lMS, int lDEL,                  // (row,  col-1)
                                            int uMS, int uINS,                  // (row-1,col)
                                            int ref1, int ref0,                 // mapped reference bytes ('N' -> 0x100)
                                            bool refN, bool gap, bool insBar,   // insBar: INS state is barred at this cell
                                            int hlimit,                         // horizLimit[col]
                                            int delNeeded, int insNeeded,
                                            const int* __restrict__ insc, const int* __restrict__ delc) {
    CellOut o;
    const bool match = (R.call1 == ref1);
    const bool prevMatch = (R.call0 == ref0);
    int limit = 0, limit3 = 0, lim2MS = 0, lim2DEL = 0, lim2INS = 0;
    if (LIMITED) {
        limit = imax(R.vlimit, hlimit);
        limit3 = imax(K.floor_, limit - (match ? P_MATCH2 : P_SUB3));
        const int delPen = delc[delNeeded];                 // 0 when delNeeded==0
        const int insPen = insc[insNeeded];                 // 0 when insNeeded==0
        lim2MS = limit - (delNeeded > 0 ? delPen : insPen);
        lim2DEL = limit - insPen;                           // delNeeded>0 && insNeeded==0 refined below (needs `time`)
        lim2INS = limit - delPen;                           // insNeeded>0 && delNeeded==0 refined below
    }
    bool good = false;
    unsigned code;
    // ---------------- MS (jni/...JNI.c:491-564) ----------------
    {
        const int sM = dMS & SMASK, sD = dDEL & SMASK, sI = dINS & SMASK, streak = dMS & TMASK;
        const int addMatch = prevMatch ? P_MATCH2 : P_MATCH;
        const int subNoPrev = streak == 0 ? P_SUB : (streak < 5 ? P_SUB2 : P_SUB3);      // POINTSoff_SUB_ARRAY[streak+1]
        const int subPrev = streak <= 1 ? P_SUBR : P_SUB;
        const int addSub = (refN || R.callN) ? 0 : (prevMatch ? subPrev : subNoPrev);
        const int a_ = sM + (match ? addMatch : addSub);
        const int mx = imax(sD, sI) + (match ? P_MATCH : P_SUB);
        const bool msWins = a_ >= mx;
        int score = imax(a_, mx);
        int time = (msWins && (match == prevMatch)) ? streak + 1 : 1;
        if (CLAMP_TIME) time = time > MAX_TIME ? TIME_WRAP : time;
        const unsigned raw = (sM >= sD && sM >= sI) ? 0u : (sD >= sI ? 1u : 2u);
        code = (time > 1) ? 0u : raw;
        bool skip = gap;
        if (LIMITED) {
            skip = skip || (imax3(sM, sD, sI) <= limit3);
            const bool ok = score >= lim2MS;
            good = ok && !skip;
            score = ok ? score : K.subfloor;
        }
        o.ms = skip ? K.subfloor : (score | time);
    }
    // ---------------- DEL (jni/...JNI.c:566-617) ----------------
    {
        const int sM = lMS & SMASK, sD = lDEL & SMASK, streak = lDEL & TMASK;
        const int ext = streak == 0 ? P_DEL : (streak < LIM3 ? P_DEL2 : (streak < LIM4 ? P_DEL3 : (streak < LIM5 ? P_DEL4 :
                        (((streak & 3) == 0) ? P_DEL5 : 0))));
        const int adj = refN ? P_DEL_REF_N : (gap ? P_GAP : 0);
        const int a_ = sM + P_DEL, b_ = sD + ext;
        const bool msWins = a_ >= b_;
        int score = imax(a_, b_) + adj;
        int time = msWins ? 1 : streak + 1;
        bool skip = R.delBar;
        if (LIMITED) {
            skip = skip || (imax(sM, sD) <= limit);
            int lim2 = lim2DEL;
            if (delNeeded > 0 && insNeeded == 0 && !skip)           // rare: below the diagonal with a live DEL state
                lim2 = limit - del_score_offset(time + delNeeded) + del_score_offset(time);
            const bool ok = score >= lim2;
            good = good || (ok && !skip);
            score = ok ? score : K.subfloor;
        }
        if (CLAMP_TIME) time = time > MAX_TIME ? TIME_WRAP : time;
        code |= ((time > 1) ? 1u : (sM >= sD ? 0u : 1u)) << 2;
        o.del = skip ? K.subfloor : (score | time);
    }
    // ---------------- INS (jni/...JNI.c:619-658) ----------------
    {
        const int sM = uMS & SMASK, sI = uINS & SMASK, streak = uINS & TMASK;
        const int ext = streak == 0 ? P_INS : (streak < LIM3 ? P_INS2 : (streak < LIM4 ? P_INS3 : P_INS4));   // POINTSoff_INS_ARRAY[streak+1]
        const int a_ = sM + P_INS, b_ = sI + ext;
        const bool msWins = a_ >= b_;
        int score = imax(a_, b_);
        int time = msWins ? 1 : streak + 1;
        bool skip = gap || insBar;
        if (LIMITED) {
            skip = skip || (imax(sM, sI) <= limit);
            int lim2 = lim2INS;
            if (insNeeded > 0 && delNeeded == 0 && !skip)           // rare: right of the end diagonal with a live INS state
                lim2 = limit - insc[imin(time + insNeeded, PEN_TAB - 1)] + insc[imin(time, PEN_TAB - 1)];
            const bool ok = score >= lim2;
            good = good || (ok && !skip);
            score = ok ? score : K.subfloor;
        }
        if (CLAMP_TIME) time = time > MAX_TIME ? TIME_WRAP : time;
        code |= ((time > 1) ? 1u : (sM >= sI ? 0u : 1u)) << 3;
        o.ins = skip ? K.subfloor : (score | time);
    }
    o.code = code;
    o.good = good;
    return o;
}

}  // namespace bbm
