// index_build.cu — device-side build and analysis of BBMap's k-mer index.
//
// What the reference does on the host:
//   IndexMaker4.BlockMaker/CountThread  current/align2/IndexMaker4.java:160-421  — counting sort of every k-mer start into
//       Block{starts[4^k+1], sites[]} per block of 2^chrombits chromosomes; site = (chrom&LOW)<<(31-chrombits) | pos
//       (BBIndex.java:3036-3057); lists ordered by (chrom,pos); keys with period <= 2 banned (IndexMaker4.java:335).
//   BBIndex.analyzeIndex                current/align2/BBIndex.java:101-191     — COUNTS[key]=fwd+rc list length, clumpy keys
//       zeroed, lengthHistogram (Tools.makeLengthHistogram3, Tools.java:1797-1850), MAX_USABLE_LENGTH.
// Here: one pass emits (key, site) for every position in (chrom,pos) order, a stable LSD radix sort by key (CUB — this is
// index set-up, not the measured path) turns that into the concatenated hit lists, an exclusive scan of the per-key counts
// gives `starts`.  Lists therefore come out sorted exactly like the reference's.
#include <cuda_runtime.h>
#include <cub/cub.cuh>
#include "msa_common.cuh"

namespace bbm {

__device__ __forceinline__ int b2n_dev(int c) {
    if (c & 0x80) return -1;
    const int u = c & 0xDF;
    return u == 'A' ? 0 : (u == 'C' ? 1 : (u == 'G' ? 2 : ((u == 'T' || u == 'U') ? 3 : -1)));
}

// AminoAcid.reverseComplementBinaryFast (dna/AminoAcid.java:258-271)
__device__ __forceinline__ int rcomp_fast_dev(int kmer, int k) {
    int out = 0;
    const int extra = k & 3;
    for (int i = 0; i < extra; ++i) { out = (out << 2) | ((~kmer) & 3); kmer >>= 2; }
    k -= extra;
    for (int i = 0; i < k; i += 4) {
        int b = kmer & 0xFF, r = 0;
        for (int j = 0; j < 4; ++j) { r = (r << 2) | ((~b) & 3); b >>= 2; }
        out = (out << 8) | (int)(short)r;
        kmer >>= 8;
    }
    return out;
}

// one thread per position of one chromosome: emit (key, site) or (INVALID, 0); count per key
__global__ void index_emit_kernel(const int8_t* __restrict__ chrom, int chromLen, int k, int siteHigh, unsigned* __restrict__ keys,
                                  int* __restrict__ vals, long long outBase, int* __restrict__ sizes, unsigned invalidKey) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    const int max = (chromLen - 1) - k + 1;              // a < max  (IndexMaker4.java:312,327)
    if (a >= chromLen) return;
    unsigned key = invalidKey; int val = 0;
    if (a < max) {
        const int first = chrom[a];
        if (first == 'A' || first == 'C' || first == 'G' || first == 'T') {
            int kk = 0; bool bad = false;
            for (int i = 0; i < k; ++i) { const int x = b2n_dev(chrom[a + i]); bad = bad || (x < 0); kk = (kk << 2) | (x & 3); }
            const int banmask = ~((-1) << (2 * k - 4));
            if (!bad && (kk >> 4) != (kk & banmask)) { key = (unsigned)kk; val = siteHigh | a; atomicAdd(&sizes[kk], 1); }
        }
    }
    keys[outBase + a] = key; vals[outBase + a] = val;
}

__global__ void count_defined_kernel(const int8_t* __restrict__ bytes, long long n, unsigned long long* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    for (; i < n; i += (long long)gridDim.x * blockDim.x) c += b2n_dev(bytes[i]) >= 0 ? 1 : 0;
    for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

// per block: COUNTS[key] += len (saturating), clumps[min(key,rkey)] += #adjacent pairs with 0 < dif <= 5
__global__ void index_analyze_block_kernel(const int* __restrict__ starts, const int* __restrict__ sites, int k, int* __restrict__ COUNTS,
                                           unsigned long long* __restrict__ clump) {
    const long long key = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long keyspace = 1LL << (2 * k);
    if (key >= keyspace) return;
    const int s = starts[key], e = starts[key + 1], len = e - s;
    const long long t = (long long)COUNTS[key] + len;
    COUNTS[key] = (int)(t > 2147483647LL ? 2147483647LL : t);
    unsigned long long clumps = 0;
    for (int i = s + 1; i < e; ++i) { const int dif = sites[i] - sites[i - 1]; clumps += (dif > 0 && dif <= 5) ? 1 : 0; }
    if (clumps) { const int r = rcomp_fast_dev((int)key, k); atomicAdd(&clump[key < r ? key : r], clumps); }
}

__global__ void index_merge_rc_kernel(int k, int* __restrict__ COUNTS) {
    const long long key = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= (1LL << (2 * k))) return;
    const int rkey = rcomp_fast_dev((int)key, k);
    if (key < rkey) {
        const long long x = (long long)COUNTS[key] + (long long)COUNTS[rkey];
        const int v = (int)(x > 2147483647LL ? 2147483647LL : x);
        COUNTS[key] = v; COUNTS[rkey] = v;
    }
}

__global__ void index_clumpy_kernel(int k, int* __restrict__ COUNTS, const unsigned long long* __restrict__ clump) {
    const long long key = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= (1LL << (2 * k))) return;
    const unsigned long long clumps = clump[key];
    if (clumps > 0) {
        const long long len = COUNTS[key];
        if (len > 2000 && (float)(long long)clumps > __fmul_rn(0.75f, (float)len)) { const int rkey = rcomp_fast_dev((int)key, k); COUNTS[key] = 0; COUNTS[rkey] = 0; }
    }
}

__global__ void index_max_kernel(const int* __restrict__ COUNTS, long long n, int* out) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int m = 0;
    for (; i < n; i += (long long)gridDim.x * blockDim.x) m = max(m, COUNTS[i]);
    for (int o = 16; o >= 1; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// counts[a]++ over COUNTS (Tools.makeLengthHistogram3, current/align2/Tools.java:1797-1815).  Nearly every k-mer of a genome-sized
// index occurs 0-3 times, so those four bins are counted in registers and everything below 1024 in shared memory; only the rare long
// lists touch global atomics (a plain global atomicAdd per element serialises on four addresses: 38 ms for k=13).
__global__ void __launch_bounds__(256) index_lenhist_kernel(const int* __restrict__ COUNTS, long long n, int* __restrict__ lenCounts) {
    __shared__ int sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        int v[4];
        if (i + 3 < n) { const int4 q = *reinterpret_cast<const int4*>(COUNTS + i); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
        else { for (int j = 0; j < 4; ++j) v[j] = (i + j < n) ? COUNTS[i + j] : -1; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int a = v[j];
            if (a < 0) continue;
            if (a == 0) ++c0; else if (a == 1) ++c1; else if (a == 2) ++c2; else if (a == 3) ++c3;
            else if (a < 1024) atomicAdd(&sh[a], 1);
            else atomicAdd(&lenCounts[a], 1);
        }
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&sh[0], c0); atomicAdd(&sh[1], c1); atomicAdd(&sh[2], c2); atomicAdd(&sh[3], c3); }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) if (sh[i]) atomicAdd(&lenCounts[i], sh[i]);
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_index_emit(const int8_t* chrom, int chromLen, int k, int siteHigh, unsigned* keys, int* vals, long long outBase, int* sizes,
                              unsigned invalidKey, cudaStream_t st) {
    if (chromLen <= 0) return 0;
    index_emit_kernel<<<(chromLen + 255) / 256, 256, 0, st>>>(chrom, chromLen, k, siteHigh, keys, vals, outBase, sizes, invalidKey);
    return (int)cudaGetLastError();
}
extern "C" int bbm_index_sort_pairs(void* temp, size_t* tempBytes, const unsigned* keysIn, unsigned* keysOut, const int* valsIn, int* valsOut,
                                    long long n, int endBit, cudaStream_t st) {
    return (int)cub::DeviceRadixSort::SortPairs(temp, *tempBytes, keysIn, keysOut, valsIn, valsOut, n, 0, endBit, st);
}
extern "C" int bbm_index_scan(void* temp, size_t* tempBytes, const int* in, int* out, long long n, cudaStream_t st) {
    return (int)cub::DeviceScan::ExclusiveSum(temp, *tempBytes, in, out, n, st);
}
extern "C" int bbm_index_count_defined(const int8_t* bytes, long long n, unsigned long long* out, cudaStream_t st) {
    count_defined_kernel<<<1024, 256, 0, st>>>(bytes, n, out);
    return (int)cudaGetLastError();
}
extern "C" int bbm_index_analyze_block(const int* starts, const int* sites, int k, int* COUNTS, unsigned long long* clump, cudaStream_t st) {
    const long long ks = 1LL << (2 * k);
    index_analyze_block_kernel<<<(unsigned)((ks + 255) / 256), 256, 0, st>>>(starts, sites, k, COUNTS, clump);
    return (int)cudaGetLastError();
}
extern "C" int bbm_index_finish_counts(int k, int* COUNTS, const unsigned long long* clump, int* maxOut, cudaStream_t st) {
    const long long ks = 1LL << (2 * k);
    index_merge_rc_kernel<<<(unsigned)((ks + 255) / 256), 256, 0, st>>>(k, COUNTS);
    index_clumpy_kernel<<<(unsigned)((ks + 255) / 256), 256, 0, st>>>(k, COUNTS, clump);
    index_max_kernel<<<1024, 256, 0, st>>>(COUNTS, ks, maxOut);
    return (int)cudaGetLastError();
}
extern "C" int bbm_index_lenhist(int k, const int* COUNTS, int* lenCounts, cudaStream_t st) {
    index_lenhist_kernel<<<1184, 256, 0, st>>>(COUNTS, 1LL << (2 * k), lenCounts);
    return (int)cudaGetLastError();
}
