// seed.cu — KeyRing seeding for read batches (north-star kernel 1).
//
// Per read, exactly what AbstractMapThread.quickMap does before the index search
// (current/align2/AbstractMapThread.java:643-733 with BBMap defaults, current/align2/BBMap.java:45-65):
//   key densities (:663-676) -> QualityTools.makeKeyProbs (QualityTools.java:188-247) -> KeyRing.makeOffsets3
//   (KeyRing.java:396-506) -> makeByteScoreArray (:145-162) -> makeKeyScores (:125-133) + probAllErrors (:712-725)
//   -> KeyRing.makeKeys / ChromosomeArray.toNumber (KeyRing.java:23-36, dna/ChromosomeArray.java:297-307).
// Java float semantics are pinned with __fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn (no FMA contraction);
// Math.round(float) = floorf(x+0.5f).
//
// Layout: one thread per read, 128 reads per block.  The block first stages the bases and qualities of its 128 reads into
// shared memory with coalesced 16-byte loads (the reads of a block are contiguous in the batch buffers), so the per-read
// sequential scans (the running product in makeKeyProbs is order-sensitive) run out of shared memory.  The per-read
// key-error-probability vector lives in a block-private global scratch, interleaved by thread so accesses coalesce.
#include <cuda_runtime.h>
#include "msa_common.cuh"

namespace bbm {

__constant__ float c_prob_correct[128];
__constant__ float c_prob_correct_inv[128];

constexpr int SEED_THREADS = 128;
constexpr int SEED_STAGE_BYTES = 22 * 1024;      // per array (bases, quality)

struct SeedParams {
    const int8_t* bases; const int8_t* quality; const long long* read_off; long long nreads;
    bbm_seed_cfg cfg; int maxKeys;
    int* nkeys; int* offsets; int* keys; int* keyScores; int8_t* baseScores;
    float* probScratch; long long probStride; int maxProbLen;
    unsigned int* counter;
};

__device__ __forceinline__ int base_to_number(int c) {      // AminoAcid.baseToNumber (dna/AminoAcid.java:615-624)
    if (c & 0x80) return -1;
    const int u = c & 0xDF;
    return u == 'A' ? 0 : (u == 'C' ? 1 : (u == 'G' ? 2 : ((u == 'T' || u == 'U') ? 3 : -1)));
}
__device__ __forceinline__ int java_round(float x) { return (int)floorf(__fadd_rn(x, 0.5f)); }

__device__ __forceinline__ int desired_keys(int readlen, int blocksize, float density, int minKeysDesired) {
    const int slots = readlen - blocksize + 1;
    const float t = __fdiv_rn(__fmul_rn((float)readlen, density), (float)blocksize);
    int desired = (int)ceil((double)t);
    desired = imax(minKeysDesired, desired);
    return imin(slots, desired);
}

// AminoAcid.reverseComplementBinaryFast (dna/AminoAcid.java:258-271); rcompBinaryTable entries are shorts
__device__ __forceinline__ int rcomp_key_fast(int kmer, int k) {
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

__global__ void __launch_bounds__(SEED_THREADS) seed_kernel(SeedParams P) {
    __shared__ __align__(16) int8_t sBases[SEED_STAGE_BYTES];
    __shared__ __align__(16) int8_t sQual[SEED_STAGE_BYTES];
    __shared__ unsigned sFirst;
    const int tid = threadIdx.x;
    float* kp = P.probScratch + ((long long)blockIdx.x * SEED_THREADS + tid);     // element i at kp[i*probStride]
    const long long stride = P.probStride;
    for (;;) {
        __syncthreads();
        if (tid == 0) sFirst = atomicAdd(P.counter, (unsigned)SEED_THREADS);
        __syncthreads();
        const long long first = sFirst;
        if (first >= P.nreads) break;
        const long long lastp1 = first + SEED_THREADS < P.nreads ? first + SEED_THREADS : P.nreads;
        const long long byte0 = P.read_off[first], byte1 = P.read_off[lastp1];
        const long long a0 = byte0 & ~15LL;                      // 16-byte aligned staging window
        const bool staged = (byte1 - a0) <= SEED_STAGE_BYTES;
        if (staged) {
            const int nvec = (int)((byte1 - a0 + 15) >> 4);
            for (int v = tid; v < nvec; v += SEED_THREADS) {
                reinterpret_cast<int4*>(sBases)[v] = reinterpret_cast<const int4*>(P.bases + a0)[v];
                if (P.quality) reinterpret_cast<int4*>(sQual)[v] = reinterpret_cast<const int4*>(P.quality + a0)[v];
            }
        }
        __syncthreads();
        const long long r = first + tid;
        if (r >= P.nreads) continue;
        const long long o = P.read_off[r];
        const int len = (int)(P.read_off[r + 1] - o);
        const int8_t* bases = staged ? sBases + (o - a0) : P.bases + o;
        const int8_t* qual = P.quality ? (staged ? sQual + (o - a0) : P.quality + o) : nullptr;
        int* of = P.offsets + r * P.maxKeys; int* ke = P.keys + r * P.maxKeys; int* ks = P.keyScores + r * P.maxKeys;
        int8_t* bs = P.baseScores + o;
        const int K = P.cfg.keylen;
        int n = 0;
        bool discard = false;
        if (len < K) { n = 0; discard = true; }
        else if (len - K + 1 > P.maxProbLen) { n = -2; discard = true; }
        else {
            int und = 0;
            for (int i = 0; i < len; ++i) und += base_to_number(bases[i]) < 0 ? 1 : 0;
            if (und > 25 && len - und < und) { n = -1; discard = true; }
        }
        if (!discard) {
            // key densities (AbstractMapThread.java:663-676)
            float keyDen2 = __fdiv_rn((float)(P.cfg.maxDesiredKeys * K), (float)len);
            keyDen2 = fmaxf(P.cfg.minKeyDensity, keyDen2);
            keyDen2 = fminf(fminf(P.cfg.keyDensity, keyDen2), (float)K);
            float keyDen3;
            if (len <= 50) keyDen3 = P.cfg.maxKeyDensity;
            else if (len >= 200) keyDen3 = __fsub_rn(P.cfg.maxKeyDensity, 0.5f);
            else keyDen3 = __fsub_rn(P.cfg.maxKeyDensity, __fmul_rn(0.003333333333f, (float)(len - 50)));
            keyDen3 = fmaxf(P.cfg.keyDensity, keyDen3);
            keyDen3 = fminf((float)K, keyDen3);
            // makeKeyProbs (QualityTools.java:188-247; USE_MODULO=false)
            const int nprob = len - K + 1;
            if (!qual) { for (int i = 0; i < nprob; ++i) kp[i * stride] = 0.f; }
            else {
                float key1 = 1.f; int tsz = 0;
                for (int i = 0; i < K; ++i) {
                    const int q = qual[i] & 127;
                    tsz = qual[i] > 0 ? tsz + 1 : 0;
                    key1 = __fmul_rn(key1, c_prob_correct[q]);
                }
                kp[0] = tsz < K ? 1.f : __fsub_rn(1.f, key1);
                for (int a = 0, b = K; b < len; ++a, ++b) {
                    const int qa = qual[a] & 127, qb = qual[b] & 127;
                    tsz = qual[b] > 0 ? tsz + 1 : 0;
                    key1 = __fmul_rn(__fmul_rn(key1, c_prob_correct_inv[qa]), c_prob_correct[qb]);
                    kp[(long long)(a + 1) * stride] = tsz < K ? 1.f : __fsub_rn(1.f, key1);
                }
            }
            // makeOffsets3 (KeyRing.java:396-506), semiperfectmode=false, minKeysDesired=2
            const int maxProbIndex = len - K;
            int left = 0, right = maxProbIndex;
            const float errorLimit2 = 0.9999f, errorLimit1 = 0.94f;
            while (left <= right && kp[left * stride] >= errorLimit1) left++;
            while (right >= left && kp[right * stride] >= errorLimit1) right--;
            int potentialKeys = 0;
            for (int i = left; i <= right; ++i) potentialKeys += kp[i * stride] < errorLimit2 ? 1 : 0;
            n = -1;
            if (potentialKeys > 0 && right >= left) {
                const int readlen2 = right - left + K;
                int desiredKeys = desired_keys(len, K, keyDen2, 2);
                if (readlen2 < len) desiredKeys = imin(desiredKeys, desired_keys(readlen2, K, keyDen3, 2));
                desiredKeys = imin(desiredKeys, potentialKeys);
                const float interval = __fdiv_rn((float)(right - left), (float)imax(desiredKeys - 1, 1));
                const int intervalInt = ((int)interval) + 1;
                float f = (float)left;
                int prev = -1; n = 0;
                for (int i = 0, j = left; i < desiredKeys; ++i) {
                    int x = -1;
                    if (prev < j) {
                        if (kp[j * stride] < errorLimit2 && (prev < 0 || j - prev > 0)) x = j;
                        else {
                            for (int k = j - 1, lim = prev + 2; k > lim; --k) if (kp[k * stride] < errorLimit2) { x = k; break; }
                            if (x < 0) {
                                const int lim = imin(j + intervalInt, right);
                                for (int k = j + 1; k < lim; ++k) if (kp[k * stride] < errorLimit2) { x = k; break; }
                            }
                        }
                    }
                    if (x > -1) { if (n < P.maxKeys) of[n] = x; n++; prev = x; }
                    else prev = imax(prev, j - 2);
                    f = __fadd_rn(f, interval);
                    j = imin(maxProbIndex, imax(j + 1, java_round(f)));
                }
                if (n > P.maxKeys) n = -3;                       // caller's maxKeys too small
                else if (n < P.cfg.minApproxHitsToKeep) n = -1;
            }
            if (n > 0) {
                // base scores, key scores, probAllErrors (AbstractMapThread.java:693-725)
                if (qual) for (int i = 0; i < len; ++i) bs[i] = (int8_t)(java_round(__fmul_rn(100.f, c_prob_correct[qual[i] & 127])) - 100);
                else for (int i = 0; i < len; ++i) bs[i] = 0;
                const int a = P.cfg.baseKeyHitScore, baseKeyScore = a / 8, range = a - baseKeyScore;
                float probAll = 1.f;
                for (int i = 0; i < n; ++i) {
                    const float p = kp[(long long)of[i] * stride];
                    ks[i] = baseKeyScore + java_round(__fmul_rn((float)range, __fsub_rn(1.f, p)));
                    probAll = __fmul_rn(probAll, p);
                }
                if (probAll > 0.50f) n = -1;
            }
            if (n > 0) {
                // makeKeys: 2-bit packing MSB first, -1 if any base is not ACGTU (ChromosomeArray.toNumber)
                for (int i = 0; i < n; ++i) {
                    int out = 0; bool bad = false;
                    for (int p = of[i]; p < of[i] + K; ++p) { const int x = base_to_number(bases[p]); bad = bad || (x < 0); out = (out << 2) | (x & 3); }
                    ke[i] = bad ? -1 : out;
                }
            }
        }
        P.nkeys[r] = n;
        const int keep = n > 0 ? n : 0;
        for (int i = keep; i < P.maxKeys; ++i) { of[i] = -1; ke[i] = -1; ks[i] = 0; }
        if (n <= 0) for (int i = 0; i < len; ++i) bs[i] = 0;
    }
}

// minus-strand keys/offsets for a key list (KeyRing.reverseComplementKeys :38-45, reverseOffsets :125-137)
__global__ void seed_reverse_kernel(const int* nkeys, const int* offsets, const int* keys, const long long* read_off, long long nreads,
                                    int maxKeys, int keylen, int* offsetsM, int* keysM) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nreads) return;
    const int n = nkeys[r] > 0 ? nkeys[r] : 0;
    const int len = (int)(read_off[r + 1] - read_off[r]);
    const int* of = offsets + r * maxKeys; const int* ke = keys + r * maxKeys;
    int* om = offsetsM + r * maxKeys; int* km = keysM + r * maxKeys;
    for (int i = 0; i < n; ++i) { om[i] = len - (of[n - 1 - i] + keylen); km[i] = rcomp_key_fast(ke[n - 1 - i], keylen); }
    for (int i = n; i < maxKeys; ++i) { om[i] = -1; km[i] = -1; }
}

}  // namespace bbm

using namespace bbm;

extern "C" int bbm_seed_upload_tables(const float* pc, const float* pci) {
    cudaError_t e = cudaMemcpyToSymbol(c_prob_correct, pc, 127 * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_prob_correct_inv, pci, 127 * sizeof(float));
    return (int)e;
}
extern "C" int bbm_launch_seed(const int8_t* bases, const int8_t* quality, const long long* read_off, long long nreads, const bbm_seed_cfg* cfg,
                               int maxKeys, int* nkeys, int* offsets, int* keys, int* keyScores, int8_t* baseScores,
                               float* probScratch, int blocks, int maxProbLen, unsigned int* counter, cudaStream_t st) {
    SeedParams P; P.bases = bases; P.quality = quality; P.read_off = read_off; P.nreads = nreads; P.cfg = *cfg; P.maxKeys = maxKeys;
    P.nkeys = nkeys; P.offsets = offsets; P.keys = keys; P.keyScores = keyScores; P.baseScores = baseScores;
    P.probScratch = probScratch; P.probStride = (long long)blocks * SEED_THREADS; P.maxProbLen = maxProbLen; P.counter = counter;
    seed_kernel<<<blocks, SEED_THREADS, 0, st>>>(P);
    return (int)cudaGetLastError();
}
extern "C" int bbm_launch_seed_reverse(const int* nkeys, const int* offsets, const int* keys, const long long* read_off, long long nreads,
                                       int maxKeys, int keylen, int* offsetsM, int* keysM, cudaStream_t st) {
    seed_reverse_kernel<<<(unsigned)((nreads + 127) / 128), 128, 0, st>>>(nkeys, offsets, keys, read_off, nreads, maxKeys, keylen, offsetsM, keysM);
    return (int)cudaGetLastError();
}
extern "C" int bbm_seed_threads() { return SEED_THREADS; }
