// ingest.cu — read normalisation for read batches: what every read goes through before mapping.
//
// Reference: Read.validate (current/stream/Read.java:81-215; switches :3406-3418) — junk detection through
// AminoAcid.baseToNumberExtended (dna/AminoAcid.java:110-114,586-595), quality clamped to [MIN_CALLED_QUALITY=2,
// MAX_CALLED_QUALITY=41] for fully defined bases and zeroed otherwise, '-' '.' 'X' 'n' -> 'N' — followed by the minus-strand
// copy AminoAcid.reverseComplementBases (dna/AminoAcid.java:203-211, table :633-647) that AbstractMapThread makes once per
// read (AbstractMapThread.java:492-503).
//
// HBM-bound byte work: 2 bytes read + 3 bytes written per base.  A block takes INGEST_READS consecutive reads, whose bytes
// are one contiguous span of the batch buffers: the span is staged with 16-byte loads, normalised in shared memory (thread per
// byte, the reverse complement is a shared-memory permutation), and written back with 16-byte stores.
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

constexpr int INGEST_THREADS = 256;

// baseToNumberExtended[b] >= 0: IUPAC letters ACMGRSVTWYHKDBNX in either case, plus U/u (dna/AminoAcid.java:110-114,586-595)
__device__ __forceinline__ bool iupac_known(int b) {
    if (b & 0x80) return false;
    const int u = b & 0xDF;
    if (u < 'A' || u > 'Y' || (b & 0x40) == 0) return false;
    // A B C D G H K M N R S T U V W X Y
    const unsigned mask = (1u << ('A' - 'A')) | (1u << ('B' - 'A')) | (1u << ('C' - 'A')) | (1u << ('D' - 'A')) | (1u << ('G' - 'A')) |
                          (1u << ('H' - 'A')) | (1u << ('K' - 'A')) | (1u << ('M' - 'A')) | (1u << ('N' - 'A')) | (1u << ('R' - 'A')) |
                          (1u << ('S' - 'A')) | (1u << ('T' - 'A')) | (1u << ('U' - 'A')) | (1u << ('V' - 'A')) | (1u << ('W' - 'A')) |
                          (1u << ('X' - 'A')) | (1u << ('Y' - 'A'));
    return (mask >> (u - 'A')) & 1u;
}

// baseToComplementExtended (dna/AminoAcid.java:129-133,633-647); entries the table leaves at -1 stay -1
__device__ __forceinline__ int complement_extended(int b) {
    if (b & 0x80) return -1;
    if (b == '?' || b == ' ' || b == '-' || b == '*' || b == '.') return b;
    const int u = b & 0xDF, lower = b & 0x20;
    if ((b & 0x40) == 0) return -1;
    int c;
    switch (u) {
        case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'M': c = 'K'; break; case 'G': c = 'C'; break;
        case 'R': c = 'Y'; break; case 'S': c = 'W'; break; case 'V': c = 'B'; break; case 'T': c = 'A'; break;
        case 'W': c = 'S'; break; case 'Y': c = 'R'; break; case 'H': c = 'D'; break; case 'K': c = 'M'; break;
        case 'D': c = 'H'; break; case 'B': c = 'V'; break; case 'N': c = 'N'; break; case 'X': c = 'X'; break;
        case 'U': c = 'A'; break;
        default: return -1;
    }
    return c | lower;
}

struct IngestParams {
    int8_t* bases; int8_t* quality; const long long* read_off; long long nreads;
    int8_t* basesM; int* readFlags; int flags; int readsPerBlock; int stageBytes;
};

__global__ void __launch_bounds__(INGEST_THREADS) ingest_kernel(IngestParams P) {
    extern __shared__ __align__(16) int8_t smem[];
    int8_t* sB = smem; int8_t* sQ = sB + P.stageBytes; int8_t* sM = sQ + P.stageBytes;
    __shared__ int sOff[66];                          // byte offsets (relative to the staging window) of this block's reads
    __shared__ int sJunk[64];                         // Read.junk() per read of this block
    const int tid = threadIdx.x;
    const bool fixJunk = P.flags & BBM_ING_FIX_JUNK, uToT = P.flags & BBM_ING_U_TO_T, toUpper = P.flags & BBM_ING_TO_UPPER_CASE,
               lowerToN = P.flags & BBM_ING_LOWER_CASE_TO_N;
    for (long long first = (long long)blockIdx.x * P.readsPerBlock; first < P.nreads; first += (long long)gridDim.x * P.readsPerBlock) {
        const long long lastp1 = first + P.readsPerBlock < P.nreads ? first + P.readsPerBlock : P.nreads;
        const int nr = (int)(lastp1 - first);
        const long long byte0 = P.read_off[first], byte1 = P.read_off[lastp1];
        const long long a0 = byte0 & ~15LL;
        const int span = (int)(byte1 - a0), lead = (int)(byte0 - a0);
        const int nvec = (span + 15) >> 4;
        __syncthreads();
        if (tid <= nr) sOff[tid] = (int)(P.read_off[first + tid] - a0);
        if (tid < nr) sJunk[tid] = 0;
        for (int v = tid; v < nvec; v += INGEST_THREADS) {
            reinterpret_cast<int4*>(sB)[v] = reinterpret_cast<const int4*>(P.bases + a0)[v];
            if (P.quality) reinterpret_cast<int4*>(sQ)[v] = reinterpret_cast<const int4*>(P.quality + a0)[v];
        }
        __syncthreads();
        // one thread per 16 staged bytes: find the read of its first byte once, then walk (Read.java:113-214 is elementwise apart
        // from the junk flag of the read; the reverse complement is a permutation inside the read)
        for (int v = tid; v < nvec; v += INGEST_THREADS) {
            const int lo = imax(v << 4, lead), hi = imin((v << 4) + 16, span);
            if (lo >= hi) continue;
            int a = 0, b = nr;                         // largest r with sOff[r] <= lo
            while (b - a > 1) { const int m = (a + b) >> 1; if (sOff[m] <= lo) a = m; else b = m; }
            int r = a;
            for (int p = lo; p < hi; ++p) {
                while (p >= sOff[r + 1]) ++r;          // empty reads are stepped over
                int bch = sB[p];
                if (uToT && (bch == 'U' || bch == 'u')) bch = (bch == 'U' ? 'T' : 't');
                if (!iupac_known(bch)) { if (fixJunk) bch = 'N'; else sJunk[r] = 1; }
                int nb = bch;
                if (P.quality) {
                    int q = sQ[p];
                    if (base_defined(bch)) { q = q < 2 ? 2 : (q > 41 ? 41 : q); }
                    else { q = 0; if (bch == '-' || bch == '.' || bch == 'X' || bch == 'n') nb = 'N'; }
                    if (toUpper && bch > 90) nb -= 32;
                    else if (lowerToN && bch > 90) nb = 'N';
                    sQ[p] = (int8_t)q;
                } else if (toUpper) {
                    if (bch > 90) nb -= 32;
                    if (bch == '-' || bch == '.' || bch == 'X') nb = 'N';
                } else if (lowerToN) {
                    if (bch > 90) nb = 'N'; else if (bch == '-' || bch == '.' || bch == 'X') nb = 'N';
                } else {
                    if (bch == '-' || bch == '.' || bch == 'X') nb = 'N';
                }
                sB[p] = (int8_t)nb;
                if (P.basesM) sM[sOff[r] + sOff[r + 1] - 1 - p] = (int8_t)complement_extended(nb);
            }
        }
        __syncthreads();
        if (tid < nr && P.readFlags) P.readFlags[first + tid] = sJunk[tid] ? BBM_READ_JUNK : 0;
        // write back: whole 16-byte vectors inside the span, single bytes at the two ragged edges (they belong to other blocks)
        for (int v = tid; v < nvec; v += INGEST_THREADS) {
            const int lo = v << 4, hi = lo + 16;
            if (lo >= lead && hi <= span) {
                reinterpret_cast<int4*>(P.bases + a0)[v] = reinterpret_cast<const int4*>(sB)[v];
                if (P.quality) reinterpret_cast<int4*>(P.quality + a0)[v] = reinterpret_cast<const int4*>(sQ)[v];
                if (P.basesM) reinterpret_cast<int4*>(P.basesM + a0)[v] = reinterpret_cast<const int4*>(sM)[v];
            } else {
                for (int j = (lo > lead ? lo : lead); j < (hi < span ? hi : span); ++j) {
                    P.bases[a0 + j] = sB[j];
                    if (P.quality) P.quality[a0 + j] = sQ[j];
                    if (P.basesM) P.basesM[a0 + j] = sM[j];
                }
            }
        }
    }
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_ingest_threads() { return INGEST_THREADS; }
extern "C" int bbm_launch_ingest(int8_t* bases, int8_t* quality, const long long* read_off, long long nreads, int8_t* basesM, int* readFlags,
                                 int flags, int readsPerBlock, int stageBytes, int blocks, cudaStream_t st) {
    IngestParams P; P.bases = bases; P.quality = quality; P.read_off = read_off; P.nreads = nreads; P.basesM = basesM; P.readFlags = readFlags;
    P.flags = flags; P.readsPerBlock = readsPerBlock; P.stageBytes = stageBytes;
    const size_t smem = (size_t)3 * stageBytes;
    cudaError_t e = cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    ingest_kernel<<<blocks, INGEST_THREADS, smem, st>>>(P);
    return (int)cudaGetLastError();
}
