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

// Look-up tables in shared memory (one copy per block).  The streak-dependent penalties are read from tables instead of being
// recomputed with compare/select ladders: the recurrence is bound by the integer ALU pipe, table reads go through the idle LSU.
struct CellTables {
    int insc[PEN_TAB];        // POINTSoff_INS_ARRAY_C[i]  (also column 0 of the matrix)
    int delc[DELC_TAB];       // calcDelScoreOffset(i), i up to columns+rows
    int delExt[PEN_TAB];      // extension cost of a DEL run of length `streak`   (jni/...JNI.c:572-576)
    int insExt[PEN_TAB];      // POINTSoff_INS_ARRAY[streak+1]                     (…JNI.java:1583-1603)
    int subExt[PEN_TAB];      // POINTSoff_SUB_ARRAY[streak+1]                     (…JNI.java:1610-1625)
};

__device__ __forceinline__ void cell_tables_init(CellTables& t) {
    for (int i = threadIdx.x; i < PEN_TAB; i += blockDim.x) {
        t.insc[i] = ins_score_offset(i);
        t.delExt[i] = i == 0 ? P_DEL : (i < LIM3 ? P_DEL2 : (i < LIM4 ? P_DEL3 : (i < LIM5 ? P_DEL4 : (((i & 3) == 0) ? P_DEL5 : 0))));
        t.insExt[i] = i == 0 ? P_INS : (i < LIM3 ? P_INS2 : (i < LIM4 ? P_INS3 : P_INS4));
        t.subExt[i] = i == 0 ? P_SUB : (i < 5 ? P_SUB2 : P_SUB3);
    }
    for (int i = threadIdx.x; i < DELC_TAB; i += blockDim.x) t.delc[i] = del_score_offset(i);
}

// The same tables behind pointers, sized for the shapes of one launch (msa_band.cu: the band kernel's occupancy is bound by shared memory, and the
// fixed-size struct above costs it 18 KB per block of 64 threads).  pen > the longest streak (max(rows, columns)), delcN > rows + columns.
struct CellTablesDyn {
    const int* insc; const int* delc; const int* delExt; const int* insExt; const int* subExt;
};
__host__ __device__ inline int cell_tables_dyn_ints(int pen, int delcN) { return 4 * pen + delcN; }
__device__ __forceinline__ void cell_tables_init_dyn(int* base, int pen, int delcN, CellTablesDyn& t) {
    int* insc = base; int* delExt = base + pen; int* insExt = base + 2 * pen; int* subExt = base + 3 * pen; int* delc = base + 4 * pen;
    for (int i = threadIdx.x; i < pen; i += blockDim.x) {
        insc[i] = ins_score_offset(i);
        delExt[i] = i == 0 ? P_DEL : (i < LIM3 ? P_DEL2 : (i < LIM4 ? P_DEL3 : (i < LIM5 ? P_DEL4 : (((i & 3) == 0) ? P_DEL5 : 0))));
        insExt[i] = i == 0 ? P_INS : (i < LIM3 ? P_INS2 : (i < LIM4 ? P_INS3 : P_INS4));
        subExt[i] = i == 0 ? P_SUB : (i < 5 ? P_SUB2 : P_SUB3);
    }
    for (int i = threadIdx.x; i < delcN; i += blockDim.x) delc[i] = del_score_offset(i);
    t.insc = insc; t.delc = delc; t.delExt = delExt; t.insExt = insExt; t.subExt = subExt;
}

// One cell.  Streaks never exceed PEN_TAB-1 here (rows <= 606, columns <= TAB_MAX_COLS = 768 in the kernels that use this function), so the
// tables need no index clamp and `time` never reaches MAX_TIME (the generic kernel handles the wrap for wider windows).
template <bool LIMITED, class TT = CellTables>
__device__ __forceinline__ CellOut msa_cell(const CellConst& K, const CellRow& R,
                                            int dMS, int dDEL, int dINS,        // (row-1,col-1)
                                            int lMS, int lDEL,                  // (row,  col-1)
                                            int uMS, int uINS,                  // (row-1,col)
                                            int ref1, int ref0,                 // mapped reference bytes ('N' -> 0x100)
                                            bool insBar,                        // INS state is barred at this cell
                                            int hlimit,                         // horizLimit[col]
                                            int delNeeded, int insNeeded,
                                            const TT& T) {
    CellOut o;
    const bool match = (R.call1 == ref1);
    const bool prevMatch = (R.call0 == ref0);
    const bool refN = (ref1 == 0x100);
    const bool gap = (ref1 == '-');
    int limit = 0, lim2MS = 0, lim2DEL = 0, lim2INS = 0;
    bool skipMS = gap, skipDEL = R.delBar, skipINS = gap || insBar;
    const int sMd = dMS & SMASK, sDd = dDEL & SMASK, sId = dINS & SMASK, streakM = dMS & TMASK;
    const int sMl = lMS & SMASK, sDl = lDEL & SMASK, streakD = lDEL & TMASK;
    const int sMu = uMS & SMASK, sIu = uINS & SMASK, streakI = uINS & TMASK;
    const int mxDI = imax(sDd, sId);
    if (LIMITED) {
        limit = imax(R.vlimit, hlimit);
        const int limit3 = __viaddmax_s32(limit, match ? -P_MATCH2 : -P_SUB3, K.floor_);
        const int delPen = T.delc[delNeeded];               // 0 when delNeeded==0
        const int insPen = T.insc[insNeeded];               // 0 when insNeeded==0
        lim2MS = limit - (delNeeded > 0 ? delPen : insPen);
        lim2DEL = limit - insPen;                           // delNeeded>0 && insNeeded==0 refined below (needs `time`)
        lim2INS = limit - delPen;                           // insNeeded>0 && delNeeded==0 refined below
        skipMS = skipMS || (imax(sMd, mxDI) <= limit3);
        skipDEL = skipDEL || (imax(sMl, sDl) <= limit);
        skipINS = skipINS || (imax(sMu, sIu) <= limit);
    }
    bool good = false;
    unsigned code;
    // ---------------- MS (jni/...JNI.c:491-564) ----------------
    {
        const int addMatch = prevMatch ? P_MATCH2 : P_MATCH;
        const int subPrev = streakM <= 1 ? P_SUBR : P_SUB;
        int addSub = prevMatch ? subPrev : T.subExt[streakM];
        addSub = (refN || R.callN) ? 0 : addSub;
        const int a_ = sMd + (match ? addMatch : addSub);
        const int mx = mxDI + (match ? P_MATCH : P_SUB);
        const bool msWins = a_ >= mx;
        int score = imax(a_, mx);
        const bool keep = msWins && (match == prevMatch);
        const int time = keep ? streakM + 1 : 1;
        // predecessor the traceback would pick: MS if time>1, else the raw arg-max of the three diagonal scores
        const bool preMS = (keep && streakM >= 1) || (sMd >= mxDI);
        code = preMS ? 0u : (sDd >= sId ? 1u : 2u);
        if (LIMITED) {
            const bool ok = score >= lim2MS;
            good = ok && !skipMS;
            score = ok ? score : K.subfloor;
        }
        o.ms = skipMS ? K.subfloor : (score | time);
    }
    // ---------------- DEL (jni/...JNI.c:566-617) ----------------
    {
        const int adj = refN ? P_DEL_REF_N : (gap ? P_GAP : 0);
        const int a_ = sMl + P_DEL, b_ = sDl + T.delExt[streakD];
        const bool msWins = a_ >= b_;
        int score = imax(a_, b_) + adj;
        const int time = msWins ? 1 : streakD + 1;
        if (LIMITED) {
            int lim2 = lim2DEL;
            if (delNeeded > 0 && insNeeded == 0 && !skipDEL)         // rare: below the diagonal with a live DEL state
                lim2 = limit - (T.delc[time + delNeeded] - T.delc[time]);
            const bool ok = score >= lim2;
            good = good || (ok && !skipDEL);
            score = ok ? score : K.subfloor;
        }
        const bool preDEL = (!msWins && streakD >= 1) || (sMl < sDl);
        code |= preDEL ? 4u : 0u;
        o.del = skipDEL ? K.subfloor : (score | time);
    }
    // ---------------- INS (jni/...JNI.c:619-658) ----------------
    {
        const int a_ = sMu + P_INS, b_ = sIu + T.insExt[streakI];
        const bool msWins = a_ >= b_;
        int score = imax(a_, b_);
        const int time = msWins ? 1 : streakI + 1;
        if (LIMITED) {
            int lim2 = lim2INS;
            if (insNeeded > 0 && delNeeded == 0 && !skipINS)         // rare: right of the end diagonal with a live INS state
                lim2 = limit - (T.insc[time + insNeeded] - T.insc[time]);
            const bool ok = score >= lim2;
            good = good || (ok && !skipINS);
            score = ok ? score : K.subfloor;
        }
        const bool preINS = (!msWins && streakI >= 1) || (sMu < sIu);
        code |= preINS ? 8u : 0u;
        o.ins = skipINS ? K.subfloor : (score | time);
    }
    o.code = code;
    o.good = good;
    return o;
}

}  // namespace bbm
