// sam.cu — SAM record fields of mapped reads on the device.
//
// Reference (Java only): SamLine(Read,int) current/stream/SamLine.java:82-330 (scaffold-relative coordinates, POS, PNEXT, TLEN, RNEXT),
// toCigar13/toCigar14 :600-750, makeFlag :2134-2151, toMapq :1709-1723, countLeadingClip/countTrailingClip/countLeadingIndels/
// countTrailingIndels :924-1020, Data.isSingleScaffold/scaffoldIndex/scaffoldRelativeLoc current/dna/Data.java:1089-1140,
// Read.containsNonM/containsNonNMS current/stream/Read.java:1815-1863.
//
// One thread per record: a run-length pass over the long-format match string (reference order) writes the CIGAR text; everything else is
// a few integer operations.  Byte work, HBM-bound: match_len bytes in, ~10-20 bytes out per read.  The (float) log2(length) of toMapq comes
// from a table computed on the host with the same libm call the oracle uses.
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

constexpr int SAM_LOG2_TAB = 4096;
__constant__ float c_sam_log2[SAM_LOG2_TAB];

struct SamParams {
    const bbm_sam_task* tasks; long long n; const int8_t* match_buf;
    const int* scaf_off; const int* scaf_loc; const int* scaf_len; int nchroms;
    bbm_sam_cfg cfg; bbm_sam_out* outs; int8_t* cigar_buf; const long long* cigar_off;
};
struct Side { int mapped, paired, has_match, idx, a, b, scaflen, pos0, pos1, gscaf; };

__device__ int bsearch_java(const int* a, int n, int key) {       // Arrays.binarySearch
    int lo = 0, hi = n - 1;
    while (lo <= hi) { const int mid = (int)(((unsigned)lo + (unsigned)hi) >> 1); const int v = a[mid]; if (v < key) lo = mid + 1; else if (v > key) hi = mid - 1; else return mid; }
    return -(lo + 1);
}
__device__ bool is_single_scaffold(const int* loc, int n, int pad, int loc1, int loc2) {
    if (n < 2) return true;
    const int idx = bsearch_java(loc, n, loc1 + pad);
    const int scaf = idx >= 0 ? idx : imax(0, (-1 - idx) - 1);
    if (scaf == n - 1) return true;
    const int lowerBound = loc[scaf] - pad, upperBound = loc[scaf + 1];
    if (loc2 < lowerBound || loc1 > upperBound) return false;
    return loc2 < upperBound;
}
__device__ int scaffold_index(const int* loc, int n, int pad, int l) {
    if (n < 2) return 0;
    const int idx = bsearch_java(loc, n, l + pad / 2);
    return idx >= 0 ? idx : imax(0, (-1 - idx) - 1);
}
__device__ void resolve_side(const SamParams& P, const bbm_sam_task& t, Side& s) {
    s.mapped = (t.flags & BBM_RF_MAPPED) != 0; s.paired = (t.flags & BBM_RF_PAIRED) != 0; s.has_match = t.match_len > 0;
    s.idx = -1; s.gscaf = -1; s.a = 0; s.b = 0; s.scaflen = 0; s.pos0 = 0; s.pos1 = 0;
    if (s.mapped) {
        const int base = P.scaf_off[t.chrom - 1], n = P.scaf_off[t.chrom] - base;
        const int* loc = P.scaf_loc + base;
        if (is_single_scaffold(loc, n, P.cfg.inter_scaffold_padding, t.start, t.stop)) {
            s.idx = scaffold_index(loc, n, P.cfg.inter_scaffold_padding, (t.start + t.stop) / 2);
            s.gscaf = base + s.idx; s.scaflen = P.scaf_len[s.gscaf];
            s.a = t.start - loc[s.idx]; s.b = s.a - t.start + t.stop;
        } else { s.mapped = 0; s.paired = 0; s.has_match = 0; }       // multi-scaffold alignment (SamLine.java:136-141)
    }
}
__device__ void positions(const SamParams& P, const bbm_sam_task& t, Side& s) {
    if (!s.mapped) { s.pos0 = 0; s.pos1 = 0; return; }
    const int8_t* m = P.match_buf + t.match_off; const int n = s.has_match ? t.match_len : 0;
    int clip = 0;                                                      // countLeadingClip (:924-945)
    if (n >= 1 && m[0] == 'C') {
        int current = 0;
        for (int i = 0; i < n; i++) {
            const int b = m[i];
            if (b >= '0' && b <= '9') current = current * 10 + (b - '0');
            else { if (current > 0) clip = clip + current - 1; current = 0; if (b != 'C') break; clip++; }
        }
        if (current > 0) clip = clip + current - 1;
    }
    int ci = 0;                                                        // countLeadingIndels (:975-996)
    if (n > 0 && s.a < 0) { int rloc = s.a, dels = 0, inss = 0; for (int i = 0; i < n && rloc < 0; i++) { const int b = m[i]; if (b == 'D') { dels++; rloc++; } else if (b == 'I') inss++; else rloc++; } ci = dels - inss; }
    int tclip = 0;                                                     // countTrailingClip (:959-973); countTrailingIndels is 0 for b >= 0 (:999)
    for (int i = n - 1; i >= 0; i--) { if (m[i] == 'C') tclip++; else break; }
    s.pos0 = (s.a + 1) + clip + ci;
    s.pos1 = (s.b + 1) - tclip;
}
__device__ int put_int(int8_t* out, int v) {
    char tmp[12]; int n = 0;
    if (v == 0) tmp[n++] = '0';
    while (v > 0) { tmp[n++] = (char)('0' + v % 10); v /= 10; }
    for (int i = 0; i < n; i++) out[i] = tmp[n - 1 - i];
    return n;
}
__device__ int java_round_f(float x) { return (int)floorf(__fadd_rn(x, 0.5f)); }
__device__ int to_mapq(int score, int length, bool mapped, bool ambig, int penalize) {     // SamLine.java:1709-1723
    if (!mapped || length < 1) return 0;
    if (ambig && penalize) {
        const float adjusted = __fdiv_rn(__fmul_rn((float)score, 3.f), __fmul_rn(100.f, (float)length));
        return imax(1, java_round_f(adjusted));
    }
    const float score2 = __fmul_rn((float)(score - length * 40), 1.6f);
    const float lg = length < SAM_LOG2_TAB ? c_sam_log2[length] : (float)(log((double)length) * (1 / log(2.0)));
    const float mx = __fadd_rn(__fmul_rn(1.5f, lg), 36.f);
    const float adjusted = __fdiv_rn(__fmul_rn(score2, mx), __fmul_rn(100.f, (float)length));
    return imax(4, java_round_f(adjusted));
}

__global__ void __launch_bounds__(128) sam_kernel(SamParams P) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    const bbm_sam_task t1 = P.tasks[i];
    const bool hasMate = t1.mate >= 0;
    bbm_sam_task t2 = t1;
    if (hasMate) t2 = P.tasks[t1.mate];
    Side s1, s2;
    resolve_side(P, t1, s1);
    s2.mapped = 0; s2.paired = 0; s2.has_match = 0; s2.idx = -1; s2.gscaf = -1; s2.pos0 = 0; s2.pos1 = 0; s2.scaflen = 0; s2.a = 0; s2.b = 0;
    if (hasMate) {
        resolve_side(P, t2, s2);
        if ((t1.flags & BBM_RF_MAPPED) && !s1.mapped) s2.paired = 0;
        if ((t2.flags & BBM_RF_MAPPED) && !s2.mapped) s1.paired = 0;
    }
    const bool sameScaf = hasMate && s1.idx > -1 && s1.idx == s2.idx && t1.chrom == t2.chrom;
    int flag = 0;
    if (hasMate) {
        flag |= 0x1;
        if (s1.mapped && s1.has_match && sameScaf && s1.paired && s2.mapped && s2.has_match) flag |= 0x2;
        flag |= (t1.flags & BBM_RF_PAIRNUM1) ? 0x80 : 0x40;
    }
    if (!s1.mapped) flag |= 0x4;
    if (hasMate && !s2.mapped) flag |= 0x8;
    if (t1.flags & BBM_RF_MINUS) flag |= 0x10;
    if (hasMate && (t2.flags & BBM_RF_MINUS)) flag |= 0x20;
    if (t1.flags & BBM_RF_SECONDARY) flag |= 0x100;
    if (t1.flags & BBM_RF_DISCARDED) flag |= 0x200;
    positions(P, t1, s1);
    if (s1.mapped) { if (s1.pos1 > s1.scaflen) s1.pos1 = s1.scaflen; if (s1.pos0 < 1) s1.pos0 = 1; }
    if (hasMate) {
        positions(P, t2, s2);
        if (s2.mapped && s2.pos0 < 1) s2.pos0 = 1;
        if (s2.mapped && s2.pos1 > s1.scaflen) s1.pos1 = s1.scaflen;     // `if(pos1_mate>scaflen){pos1=scaflen;}` (SamLine.java:205): touches pos1
    }
    int pos, pnext, tlen = 0;
    if (!hasMate) { pos = s1.pos0; pnext = 0; }
    else if (s1.mapped && s2.mapped) { pos = s1.pos0; pnext = s2.pos0; tlen = sameScaf ? 1 + (imax(s1.pos1, s2.pos1) - imin(s1.pos0, s2.pos0)) : 0; }
    else if (s1.mapped) { pos = s1.pos0; pnext = s1.pos0; }
    else if (s2.mapped) { pos = s2.pos0; pnext = s2.pos0; }
    else { pos = s1.pos0; pnext = s2.pos0; }
    if (!(!hasMate || t1.start < t2.start || (t1.start == t2.start && !(t1.flags & BBM_RF_PAIRNUM1)))) tlen = -tlen;
    bbm_sam_out O;
    O.flag = flag; O.pos = pos; O.pnext = pnext; O.tlen = tlen;
    O.mapq = to_mapq(t1.score, t1.read_len, s1.mapped, (t1.flags & BBM_RF_AMBIGUOUS) != 0, P.cfg.penalize_ambig);
    O.scaffold = s1.mapped ? s1.gscaf : ((hasMate && s2.mapped) ? s2.gscaf : -1);
    O.rnext = (!hasMate || (!s1.mapped && !s2.mapped)) ? -1 : ((s1.mapped && s2.mapped) ? (sameScaf ? -2 : s2.gscaf) : -2);
    O.cigar_len = -1;
    if (s1.mapped && s1.has_match && t1.read_len > 0) {
        const int8_t* m = P.match_buf + t1.match_off; const int ml = t1.match_len;
        int8_t* cg = P.cigar_buf + P.cigar_off[i];
        const bool inbounds = (s1.a >= 0 && s1.b < s1.scaflen), perfect = (t1.flags & BBM_RF_PERFECT) != 0, v14 = P.cfg.version14 != 0;
        bool nonM = false, nonNMS = false;
        for (int k = 0; k < ml; k++) { const int b = m[k]; nonM |= (b > '9' && b != 'm'); nonNMS |= (b > '9' && b != 'm' && b != 's' && b != 'N' && b != 'S'); }
        if (v14 ? (inbounds && perfect && !nonM) : (inbounds && (perfect || !nonNMS))) {
            int o = put_int(cg, t1.read_len); cg[o++] = v14 ? '=' : 'M'; O.cigar_len = o;
        } else if (s1.a == s1.b) O.cigar_len = -1;                          // toCigar: readStart==readStop -> null
        else {
            // toCigar13 / toCigar14 (:600-750)
            int count = 0, o = 0, refloc = s1.a; char mode = '=', lastMode = '=';
            const int reflen = s1.scaflen; bool bad = false;
            for (int mpos = 0; mpos < ml; mpos++) {
                const int mm = m[mpos];
                bool sfd = false;
                if (P.cfg.soft_clip && (refloc < 0 || refloc >= reflen)) { mode = 'S'; if (mm != 'I') refloc++; if (mm == 'D') sfd = true; }
                else if (mm == 'I' || mm == 'X' || mm == 'Y') mode = 'I';
                else if (mm == 'D') { mode = 'D'; refloc++; }
                else if (mm == 'C') { mode = 'S'; refloc++; }
                else if (v14) {
                    if (mm == 'm' || mm == 's') { mode = '='; refloc++; }
                    else if (mm == 'S') { mode = 'X'; refloc++; }
                    else if (mm == 'N' || mm == 'B') { mode = 'M'; refloc++; }
                    else { bad = true; break; }
                } else {
                    if (mm == 'm' || mm == 's' || mm == 'S' || mm == 'N' || mm == 'B') { mode = 'M'; refloc++; }
                    else { bad = true; break; }
                }
                if (mode != lastMode) {
                    if (count > 0) { o += put_int(cg + o, count); cg[o++] = (lastMode == 'D' && count > P.cfg.intron_limit) ? 'N' : lastMode; }
                    count = 0; lastMode = mode;
                }
                count++;
                if (sfd) count--;
            }
            if (bad) O.cigar_len = -2;                                      // the reference throws "Invalid match string character"
            else { o += put_int(cg + o, count); cg[o++] = (mode == 'D' && count > P.cfg.intron_limit) ? 'N' : mode; O.cigar_len = o; }
        }
    }
    P.outs[i] = O;
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_sam_upload_table(const float* log2tab) { return (int)cudaMemcpyToSymbol(c_sam_log2, log2tab, sizeof(float) * SAM_LOG2_TAB); }
extern "C" int bbm_sam_log2_tab() { return SAM_LOG2_TAB; }
extern "C" int bbm_launch_sam(const bbm_sam_task* tasks, long long n, const int8_t* match_buf, const int* scaf_off, const int* scaf_loc, const int* scaf_len,
                              int nchroms, const bbm_sam_cfg* cfg, bbm_sam_out* outs, int8_t* cigar_buf, const long long* cigar_off, cudaStream_t st) {
    SamParams P; P.tasks = tasks; P.n = n; P.match_buf = match_buf; P.scaf_off = scaf_off; P.scaf_loc = scaf_loc; P.scaf_len = scaf_len; P.nchroms = nchroms;
    P.cfg = *cfg; P.outs = outs; P.cigar_buf = cigar_buf; P.cigar_off = cigar_off;
    sam_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(P);
    return (int)cudaGetLastError();
}
