// samtext.cu — SAM lines of a mapped batch: SamLine.toBytes (current/stream/SamLine.java:1925-1960) with the optional tags BBMap writes by
// default, makeOptionalTags (:1481-1560: XT:A:R for ambiguous reads, NM:i, AM:i; MAKE_NM_TAG / MAKE_AM_TAG :2405-2406), the QNAME rule of
// SamLine(Read,int) (:100-112) and the reverse-complemented SEQ / reversed QUAL of minus-strand reads (:2061-2095).
//
// HBM-bound byte work: pass 1 (thread per read) computes every line's length and its NM value, an exclusive scan places the lines, pass 2
// (warp per read) writes them — lane 0 formats the numeric fields, all lanes copy QNAME / RNAME / CIGAR / SEQ / QUAL with coalesced accesses.
#include <cuda_runtime.h>
#include "mapper_kernels.cuh"

namespace bbm {


__device__ __forceinline__ int ndigits(int v) {          // length of Integer.toString(v)
    int n = v < 0 ? 1 : 0; unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
    do { n++; u /= 10; } while (u);
    return n;
}
__device__ int put_i(int8_t* out, int v) {
    int n = 0; unsigned u = v < 0 ? (unsigned)(-(long long)v) : (unsigned)v;
    if (v < 0) out[n++] = '-';
    char tmp[12]; int k = 0;
    do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
    while (k) out[n++] = tmp[--k];
    return n;
}
// QNAME: r.id with tabs replaced, "/1" "/2" " 1" " 2" cut for pairs (:100-112)
__device__ __forceinline__ int qname_len(const SamTextParams& P, long long r) {
    if (!P.names) return 1;
    const long long a = P.name_off[r]; int n = (int)(P.name_off[r + 1] - a);
    if (n == 0) return 0;
    if (P.paired && n > 2) {
        const int8_t c = P.names[a + n - 2]; const int num = P.names[a + n - 1] - '1';
        if ((num == 0 || num == 1) && (c == ' ' || c == '/')) n -= 2;
    }
    return n;
}
__device__ __forceinline__ int scaf_len(const SamTextParams& P, int idx) { return (idx < 0 || !P.scaf_names) ? 1 : (int)(P.scaf_name_off[idx + 1] - P.scaf_name_off[idx]); }

struct LineShape { int qn, rn, cg, rx, len, am; bool mapped, minus, ambig, tags, hasNM; };

__device__ LineShape line_shape(const SamTextParams& P, long long r, int nm) {
    const bbm_map_rec q = P.recs[r]; const bbm_sam_out o = P.sam[r];
    LineShape L;
    L.mapped = (o.flag & 0x4) == 0; L.minus = (o.flag & 0x10) != 0; L.ambig = (q.flags & 4) != 0;
    L.len = (int)(P.read_off[r + 1] - P.read_off[r]);
    L.qn = qname_len(P, r);
    L.rn = scaf_len(P, o.scaffold);
    L.cg = o.cigar_len > 0 ? o.cigar_len : 1;
    L.rx = (o.rnext == -1 || o.rnext == -2) ? 1 : scaf_len(P, o.rnext);
    L.tags = (q.flags & 1) != 0;                           // makeOptionalTags: r.mapped()
    L.hasNM = L.tags && ((q.flags & 2) || q.match_len > 0);
    int am = o.mapq;
    if (P.paired) { const bbm_map_rec m = P.recs[r ^ 1]; const int ml = (int)(P.read_off[(r ^ 1) + 1] - P.read_off[r ^ 1]); const int v = (m.flags & 1) ? max(1, ml > 0 ? m.map_score / ml : 0) : 0; am = min(am, v); }
    L.am = am;
    (void)nm;
    return L;
}
__device__ int line_length(const SamTextParams& P, const LineShape& L, const bbm_sam_out& o, int nm) {
    int n = L.qn + 1 + ndigits(o.flag) + 1 + L.rn + 1 + ndigits(o.pos) + 1 + ndigits(o.mapq) + 1 + L.cg + 1 + L.rx + 1 + ndigits(o.pnext) + 1 + ndigits(o.tlen) + 1;
    n += (L.len > 0 ? L.len : 1) + 1 + ((P.quality && L.len > 0) ? L.len : 1);
    if (L.tags) {
        if (L.ambig) n += 7;                               // \tXT:A:R
        if (L.hasNM) n += 6 + ndigits(nm);                 // \tNM:i:
        n += 6 + ndigits(L.am);                            // \tAM:i:
    }
    return n + 1;                                          // '\n'
}

__global__ void __launch_bounds__(128) samtext_len_kernel(SamTextParams P) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= P.nreads) return;
    const bbm_map_rec q = P.recs[r]; const bbm_sam_out o = P.sam[r];
    int nm = 0;
    if ((q.flags & 1) && !(q.flags & 2) && q.match_len > 0 && q.match_slot >= 0) {       // makeOptionalTags :1513-1541
        const int len = (int)(P.read_off[r + 1] - P.read_off[r]);
        int leftclip = 0, rightclip = 0;
        if (o.cigar_len > 0) {
            const int8_t* cg = P.cigar + P.cigar_off[r]; const int cl = o.cigar_len;
            int v = 0;
            for (int i = 0; i < cl; i++) { const int c = cg[i]; if (c >= '0' && c <= '9') v = v * 10 + (c - '0'); else { leftclip = (c == 'S') ? v : 0; break; } }
            if (cg[cl - 1] == 'S') { int p = cl - 2; while (p >= 0 && cg[p] >= '0' && cg[p] <= '9') p--; v = 0; for (int i = p + 1; i < cl - 1; i++) v = v * 10 + (cg[i] - '0'); rightclip = v; }
        }
        const int from = leftclip, to = len - rightclip;
        const int8_t* m = P.mslots + (r * GM_SLOTS + q.match_slot) * P.ms;
        int dels = 0;
        for (int i = 0, cpos = 0; i < q.match_len; i++) {
            const int8_t b = m[i];
            if (cpos >= from && cpos < to) {
                if (b == 'I' || b == 'S' || b == 'N' || b == 'X' || b == 'Y') nm++;
                if (b == 'D') dels++; else { if (dels <= P.intron_limit) nm += dels; dels = 0; }
            }
            if (b != 'D') cpos++;
        }
        if (dels <= P.intron_limit) nm += dels;
    }
    P.nm[r] = nm;
    const LineShape L = line_shape(P, r, nm);
    P.lens[r] = line_length(P, L, o, nm);
}

__device__ __forceinline__ void wcopy(int8_t* dst, const int8_t* src, int n, int lane) { for (int i = lane; i < n; i += 32) dst[i] = src[i]; }

__global__ void __launch_bounds__(256) samtext_write_kernel(SamTextParams P) {
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= P.nreads) return;
    const bbm_sam_out o = P.sam[r];
    const int nm = P.nm[r];
    const LineShape L = line_shape(P, r, nm);
    int8_t* out = P.text + P.line_off[r];
    if (lane == 0) P.text_off[r] = P.line_off[r];
    if (r == P.nreads - 1 && lane == 0) P.text_off[P.nreads] = P.line_off[P.nreads];
    int p = 0;
    // QNAME
    if (!P.names) { if (lane == 0) out[0] = '*'; }
    else { const int8_t* s = P.names + P.name_off[r]; for (int i = lane; i < L.qn; i += 32) { const int8_t c = s[i]; out[i] = c == '\t' ? '_' : c; } }
    p += L.qn;
    if (lane == 0) { int k = p; out[k++] = '\t'; k += put_i(out + k, o.flag); out[k++] = '\t'; }
    p += 2 + ndigits(o.flag);
    if (o.scaffold < 0 || !P.scaf_names) { if (lane == 0) out[p] = '*'; } else wcopy(out + p, P.scaf_names + P.scaf_name_off[o.scaffold], L.rn, lane);
    p += L.rn;
    if (lane == 0) { int k = p; out[k++] = '\t'; k += put_i(out + k, o.pos); out[k++] = '\t'; k += put_i(out + k, o.mapq); out[k++] = '\t'; }
    p += 3 + ndigits(o.pos) + ndigits(o.mapq);
    if (o.cigar_len > 0) wcopy(out + p, P.cigar + P.cigar_off[r], o.cigar_len, lane); else if (lane == 0) out[p] = '*';
    p += L.cg;
    if (lane == 0) out[p] = '\t';
    p += 1;
    if (o.rnext == -1) { if (lane == 0) out[p] = '*'; } else if (o.rnext == -2) { if (lane == 0) out[p] = '='; }
    else if (!P.scaf_names) { if (lane == 0) out[p] = '*'; } else wcopy(out + p, P.scaf_names + P.scaf_name_off[o.rnext], L.rx, lane);
    p += L.rx;
    if (lane == 0) { int k = p; out[k++] = '\t'; k += put_i(out + k, o.pnext); out[k++] = '\t'; k += put_i(out + k, o.tlen); out[k++] = '\t'; }
    p += 3 + ndigits(o.pnext) + ndigits(o.tlen);
    // SEQ: bases as sequenced, or their reverse complement when the line says mapped and minus (:1942-1948)
    const bool rc = L.mapped && L.minus;
    if (L.len > 0) wcopy(out + p, (rc ? P.basesM : P.bases) + P.read_off[r], L.len, lane); else if (lane == 0) out[p] = '*';
    p += (L.len > 0 ? L.len : 1);
    if (lane == 0) out[p] = '\t';
    p += 1;
    if (P.quality && L.len > 0) {
        const int8_t* qv = P.quality + P.read_off[r];
        for (int i = lane; i < L.len; i += 32) out[p + i] = (int8_t)(qv[rc ? L.len - 1 - i : i] + 33);
        p += L.len;
    } else { if (lane == 0) out[p] = '*'; p += 1; }
    if (lane == 0) {
        int k = p;
        if (L.tags) {
            if (L.ambig) { const char* t = "\tXT:A:R"; for (int i = 0; i < 7; i++) out[k++] = t[i]; }
            if (L.hasNM) { const char* t = "\tNM:i:"; for (int i = 0; i < 6; i++) out[k++] = t[i]; k += put_i(out + k, nm); }
            { const char* t = "\tAM:i:"; for (int i = 0; i < 6; i++) out[k++] = t[i]; k += put_i(out + k, L.am); }
        }
        out[k++] = '\n';
    }
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_launch_samtext_len(const SamTextParams* P, cudaStream_t st) {
    samtext_len_kernel<<<(unsigned)((P->nreads + 127) / 128), 128, 0, st>>>(*P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_samtext_write(const SamTextParams* P, cudaStream_t st) {
    samtext_write_kernel<<<(unsigned)((P->nreads * 32 + 255) / 256), 256, 0, st>>>(*P);
    return (int)cudaGetLastError();
}
extern "C" size_t bbm_samtext_params_size() { return sizeof(SamTextParams); }
