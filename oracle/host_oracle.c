/*
 * host_oracle.c — TEST INFRASTRUCTURE ONLY.
 * C restatement of the Java-only seeding stage of BBMap ("KeyRing" path) as executed by
 * AbstractMapThread.quickMap (current/align2/AbstractMapThread.java:643-733) with BBMap's defaults
 * (current/align2/BBMap.java:45-65: k=13, keyDensity 1.9, maxKeyDensity 3, minKeyDensity 1.5, maxDesiredKeys 15):
 *   QualityTools tables            current/align2/QualityTools.java:475-480, 519-539
 *   QualityTools.makeKeyProbs      :188-247 (quality) / :250-280 (no quality)
 *   KeyRing.desiredKeysFromDensity current/align2/KeyRing.java:269-282
 *   KeyRing.makeOffsets3           :396-506
 *   QualityTools.makeByteScoreArray:145-162, makeKeyScores :125-133
 *   KeyRing.makeKeys               :23-36 -> ChromosomeArray.toNumber dna/ChromosomeArray.java:297-307
 *   KeyRing.reverseComplementKeys  :38-45 -> AminoAcid.reverseComplementBinaryFast dna/AminoAcid.java:258-271
 *   KeyRing.reverseOffsets         :125-137
 * Java float semantics are kept: every operation is a separate IEEE single-precision operation (no FMA; gcc is told
 * -ffp-contract=off for this file), Math.round(float) = floor(x+0.5f), Math.ceil on the widened double.
 * PARITY UNPINNED: no JVM here, so these functions are pinned only by the invariants the reference asserts
 * (offsets strictly ascending and in range, BBIndex.checkOffsets :200-205) — not by Java outputs.
 */
#pragma GCC optimize ("fp-contract=off")
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "host_oracle.h"

static float PC[127], PCI[127];
static int tables_ready = 0;
static int8_t B2N[128];

void orc_quality_tables(float* prob_correct, float* prob_correct_inverse) {
    for (int i = 0; i < 127; i++) {
        float pe = (float)pow(10.0, 0 - .1 * i);
        if (i == 0) pe = .8f;
        const float pc = 1 - pe;
        PC[i] = pc; PCI[i] = 1 / pc;
    }
    memset(B2N, -1, 128);
    const char* b = "ACGT";
    for (int i = 0; i < 4; i++) { B2N[(int)b[i]] = (int8_t)i; B2N[(int)b[i] + 32] = (int8_t)i; }
    B2N['U'] = 3; B2N['u'] = 3;
    tables_ready = 1;
    if (prob_correct) memcpy(prob_correct, PC, sizeof(PC));
    if (prob_correct_inverse) memcpy(prob_correct_inverse, PCI, sizeof(PCI));
}
static void ready(void) { if (!tables_ready) orc_quality_tables(0, 0); }

static int java_round_f(float x) { return (int)floorf(x + 0.5f); }

void orc_make_key_probs(const int8_t* quality, const int8_t* bases, int len, int keylen, float* out) {
    ready();
    (void)bases;      /* USE_MODULO=false (IndexMaker4.java:522): bases only matter for the modulo filter */
    const int n = len - keylen + 1;
    if (!quality) { for (int i = 0; i < n; i++) out[i] = 0; return; }
    float key1 = 1;
    int timeSinceZero = 0;
    for (int i = 0; i < keylen; i++) {
        const int q = quality[i];
        if (q > 0) timeSinceZero++; else timeSinceZero = 0;
        key1 *= PC[q];
    }
    out[0] = 1 - key1;
    if (timeSinceZero < keylen) out[0] = 1;
    for (int a = 0, b = keylen; b < len; a++, b++) {
        const int qa = quality[a], qb = quality[b];
        if (qb > 0) timeSinceZero++; else timeSinceZero = 0;
        const float ipa = PCI[qa], pb = PC[qb];
        key1 = key1 * ipa * pb;
        out[a + 1] = 1 - key1;
        if (timeSinceZero < keylen) out[a + 1] = 1;
    }
}

static int desired_keys(int readlen, int blocksize, float density, int minKeysDesired) {
    const int slots = readlen - blocksize + 1;
    int desired = (int)ceil((double)((readlen * density) / blocksize));
    if (desired < minKeysDesired) desired = minKeysDesired;
    if (desired > slots) desired = slots;
    return desired;
}

int orc_make_offsets3(const float* kep, int readlenOriginal, int blocksize, float density, float maxDensity, int minKeysDesired,
                      int semiperfect, int* offsets /* cap >= desired */) {
    int readlen = readlenOriginal;
    const int maxProbIndex = readlen - blocksize;
    int left = 0, right = maxProbIndex;
    const float errorLimit2 = 0.9999f, errorLimit1 = semiperfect ? 0.99f : 0.94f;
    while (left <= right && kep[left] >= errorLimit1) left++;
    while (right >= left && kep[right] >= errorLimit1) right--;
    int potentialKeys = 0;
    for (int i = left; i <= right; i++) if (kep[i] < errorLimit2) potentialKeys++;
    if (potentialKeys == 0) return -1;
    if (right < left) return -1;
    readlen = right - left + blocksize;
    int desiredKeys = desired_keys(readlenOriginal, blocksize, density, minKeysDesired);
    if (readlen < readlenOriginal) {
        const int d2 = desired_keys(readlen, blocksize, maxDensity, minKeysDesired);
        if (d2 < desiredKeys) desiredKeys = d2;
    }
    if (potentialKeys < desiredKeys) desiredKeys = potentialKeys;
    const float interval = (right - left) / (float)(desiredKeys - 1 > 1 ? desiredKeys - 1 : 1);
    const int intervalInt = ((int)interval) + 1;
    float f = left;
    int prev = -1, n = 0;
    for (int i = 0, j = left; i < desiredKeys; i++) {
        int x = -1;
        if (prev < j) {
            if (kep[j] < errorLimit2 && (prev < 0 || j - prev > 0)) x = j;
            else {
                for (int k = j - 1, lim = prev + 2; k > lim; k--) if (kep[k] < errorLimit2) { x = k; break; }
                if (x < 0) {
                    const int lim = (j + intervalInt < right) ? j + intervalInt : right;
                    for (int k = j + 1; k < lim; k++) if (kep[k] < errorLimit2) { x = k; break; }
                }
            }
        }
        if (x > -1) { offsets[n++] = x; prev = x; }
        else { prev = prev > j - 2 ? prev : j - 2; }
        f += interval;
        const int rf = java_round_f(f);
        j = j + 1 > rf ? j + 1 : rf;
        if (j > maxProbIndex) j = maxProbIndex;
    }
    return n;
}

static int to_number(const int8_t* bases, int a, int b) {
    int out = 0;
    for (int i = a; i <= b; i++) {
        const int c = bases[i];
        const int x = (c >= 0) ? B2N[c] : -1;
        if (x < 0) return -1;
        out = (out << 2) | x;
    }
    return out;
}

static int rcomp_binary(int kmer, int k) {      /* AminoAcid.reverseComplementBinary: reverse the 2-bit letters, complement each */
    int out = 0;
    for (int i = 0; i < k; i++) { out = (out << 2) | ((~kmer) & 3); kmer >>= 2; }
    return out;
}
int orc_rcomp_key_fast(int kmer, int k) {
    /* dna/AminoAcid.java:258-271 with rcompBinaryTable[i]=(short)reverseComplementBinary(i,4) */
    int out = 0;
    const int extra = k & 3;
    for (int i = 0; i < extra; i++) { out = (out << 2) | ((~kmer) & 3); kmer >>= 2; }
    k -= extra;
    for (int i = 0; i < k; i += 4) { out = (out << 8) | (int)(int16_t)rcomp_binary(kmer & 0xFF, 4); kmer >>= 8; }
    return out;
}

/* quickMap's seeding for one read.  Returns n keys (>=1), 0 (read shorter than k) or -1 (discarded). */
int orc_quickmap_seed(const int8_t* bases, const int8_t* quality, int len, const orc_seed_cfg* cfg,
                      int32_t* offsets, int32_t* keys, int32_t* keyScores, int8_t* baseScores, float* keyProbsScratch) {
    ready();
    const int K = cfg->keylen;
    if (len < K) return 0;
    {   /* DISCARD_MOSTLY_UNDEFINED_READS (AbstractMapThread.java:651-654) */
        int n = 0;
        for (int i = 0; i < len; i++) { const int c = bases[i]; if (c < 0 || B2N[c] < 0) n++; }
        if (n > 25 && len - n < n) return -1;
    }
    const int keyProbLen = len - K + 1;
    float keyDen2 = ((cfg->maxDesiredKeys * K) / (float)len);
    if (keyDen2 < cfg->minKeyDensity) keyDen2 = cfg->minKeyDensity;
    { float m = cfg->keyDensity < keyDen2 ? cfg->keyDensity : keyDen2; if ((float)K < m) m = (float)K; keyDen2 = m; }
    float keyDen3;
    if (len <= 50) keyDen3 = cfg->maxKeyDensity;
    else if (len >= 200) keyDen3 = cfg->maxKeyDensity - 0.5f;
    else keyDen3 = cfg->maxKeyDensity - 0.003333333333f * (len - 50);
    if (keyDen3 < cfg->keyDensity) keyDen3 = cfg->keyDensity;
    if ((float)K < keyDen3) keyDen3 = (float)K;
    orc_make_key_probs(quality, bases, len, K, keyProbsScratch);
    const int n = orc_make_offsets3(keyProbsScratch, len, K, keyDen2, keyDen3, 2, 0, offsets);
    if (n < cfg->minApproxHitsToKeep) return -1;
    if (quality) for (int i = 0; i < len; i++) baseScores[i] = (int8_t)(java_round_f(100 * PC[(int)quality[i]]) - 100);
    else memset(baseScores, 0, (size_t)len);
    const int a = cfg->baseKeyHitScore, baseKeyScore = a / 8, range = a - baseKeyScore;
    float probAllErrors = 1.f;
    for (int i = 0; i < n; i++) {
        const float p = keyProbsScratch[offsets[i]];
        keyScores[i] = baseKeyScore + java_round_f(range * (1 - p));
        probAllErrors *= p;
    }
    (void)keyProbLen;
    if (probAllErrors > 0.50f) return -1;
    for (int i = 0; i < n; i++) keys[i] = to_number(bases, offsets[i], offsets[i] + K - 1);
    return n;
}

void orc_seed_batch(const int8_t* bases, const int8_t* quality, const int64_t* read_off, int64_t nreads, const orc_seed_cfg* cfg,
                    int32_t maxKeys, int32_t* nkeys, int32_t* offsets, int32_t* keys, int32_t* keyScores, int8_t* baseScores) {
    float kp[4096];
    for (int64_t r = 0; r < nreads; r++) {
        const int64_t o = read_off[r]; const int len = (int)(read_off[r + 1] - o);
        int32_t* of = offsets + r * maxKeys; int32_t* ke = keys + r * maxKeys; int32_t* ks = keyScores + r * maxKeys;
        for (int i = 0; i < maxKeys; i++) { of[i] = -1; ke[i] = -1; ks[i] = 0; }
        memset(baseScores + o, 0, (size_t)len);
        nkeys[r] = (len - cfg->keylen + 1 > 4096) ? -2 : orc_quickmap_seed(bases + o, quality ? quality + o : 0, len, cfg, of, ke, ks, baseScores + o, kp);
        if (nkeys[r] <= 0) { for (int i = 0; i < maxKeys; i++) { of[i] = -1; ke[i] = -1; ks[i] = 0; } memset(baseScores + o, 0, (size_t)len); }
    }
}

/* ---- MSA.scoreNoIndels / scoreNoIndelsAndMakeMatchString (MultiStateAligner11tsJNI.java:1033-1089, 1243-1318) ---- */
int orc_score_no_indels(const int8_t* read, int len, const int8_t* ref, int refLen, int refStart, int8_t* match /* NULL = plain scoreNoIndels */) {
    int score = 0, mode = -1, timeInMode = 0, readStart = 0, readStop = len;
    const long long refStop = (long long)refStart + len;
    if (match && (refStart < 0 || refStop > refLen)) return -99999;
    if (refStart < 0) readStart = -refStart;            /* POINTS_NOREF*readStart == 0 */
    if (refStop > refLen) readStop -= (int)(refStop - refLen);
    for (int i = readStart; i < readStop; i++) {
        const int8_t c = read[i], r = ref[refStart + i];
        if (c == r && c != 'N') {
            if (mode == 0) { timeInMode++; score += 100; } else { timeInMode = 0; score += 70; }
            if (match) match[i] = 'm';
            mode = 0;
        } else if (c < 0 || c == 'N') { if (match) match[i] = 'N'; }
        else if (r < 0 || r == 'N') { if (match) match[i] = 'N'; }
        else {
            if (match) match[i] = 'S';
            if (mode == 3) timeInMode++; else timeInMode = 0;
            const int t = timeInMode + 1;
            score += t > 5 ? -25 : (t > 1 ? -51 : -127);
            mode = 3;
        }
    }
    return score;
}
void orc_noindel_batch(const int8_t* reads, const int8_t* refs, const orc_noindel_task* tasks, int32_t* scores, int8_t* match_buf,
                       const int64_t* match_off, int64_t n) {
    for (int64_t i = 0; i < n; i++) {
        const orc_noindel_task* T = &tasks[i];
        int8_t* m = ((T->flags & 1) && match_buf) ? match_buf + match_off[i] : 0;
        scores[i] = orc_score_no_indels(reads + T->read_off, T->read_len, refs + T->ref_off, T->ref_len, T->ref_start, m);
    }
}

/* ---------------- Read.validate + AminoAcid.reverseComplementBases ----------------
 * current/stream/Read.java:81-215 with the switches of :3406-3418 passed as `flags` (1 FIX_JUNK, 2 U_TO_T, 4 TO_UPPER_CASE,
 * 8 LOWER_CASE_TO_N); nucleotide reads only.  Tables: current/dna/AminoAcid.java:110-133, 586-595, 615-647.
 * Returns 1 if the read was flagged junk. */
static int8_t g_b2nExt[128], g_compExt[128], g_ing_init = 0;
static void ingest_init(void) {
    static const char* ext = " ACMGRSVTWYHKDBNX"; static const char* cext = " TGKCYWBASRDMHVNX";
    memset(g_b2nExt, -1, 128); memset(g_compExt, -1, 128);
    for (int i = 0; ext[i]; i++) {
        const int x = ext[i], x2 = cext[i];
        if (x != ' ') { g_b2nExt[x] = (int8_t)i; g_b2nExt[tolower(x)] = (int8_t)i; }
        g_compExt[x] = (int8_t)x2; g_compExt[tolower(x)] = (int8_t)tolower(x2);
    }
    g_b2nExt['U'] = 8; g_b2nExt['u'] = 8;
    g_compExt['U'] = 'A'; g_compExt['u'] = 'a'; g_compExt['?'] = '?'; g_compExt[' '] = ' '; g_compExt['-'] = '-'; g_compExt['*'] = '*'; g_compExt['.'] = '.';
    g_ing_init = 1;
}
static int fully_defined(int8_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T' || b == 'U' || b == 'a' || b == 'c' || b == 'g' || b == 't' || b == 'u'; }

int orc_ingest_read(int8_t* bases, int8_t* quality, int len, int flags, int8_t* basesM) {
    if (!g_ing_init) ingest_init();
    const int fixJunk = flags & 1, uToT = flags & 2, toUpper = flags & 4, lowerToN = flags & 8;
    int junk = 0;
    if (uToT) for (int i = 0; i < len; i++) { if (bases[i] == 'U') bases[i] = 'T'; else if (bases[i] == 'u') bases[i] = 't'; }
    for (int i = 0; i < len; i++) {
        const int8_t b = bases[i];
        const int num = b < 0 ? -1 : g_b2nExt[(int)b];      /* Java would throw on a negative index; treated as junk */
        if (num < 0) { if (fixJunk) bases[i] = 'N'; else { junk = 1; break; } }
    }
    if (quality) {
        for (int i = 0; i < len; i++) {
            const int8_t b = bases[i], q = quality[i];
            if (fully_defined(b)) { if (q < 2) quality[i] = 2; else if (q > 41) quality[i] = 41; }
            else { quality[i] = 0; if (b == '-' || b == '.' || b == 'X' || b == 'n') bases[i] = 'N'; }
            if (toUpper && b > 90) bases[i] -= 32;
            else if (lowerToN && b > 90) bases[i] = 'N';
        }
    } else if (toUpper) {
        for (int i = 0; i < len; i++) { const int8_t b = bases[i]; if (b > 90) bases[i] -= 32; if (b == '-' || b == '.' || b == 'X') bases[i] = 'N'; }
    } else if (lowerToN) {
        for (int i = 0; i < len; i++) { const int8_t b = bases[i]; if (b > 90) bases[i] = 'N'; else if (b == '-' || b == '.' || b == 'X') bases[i] = 'N'; }
    } else {
        for (int i = 0; i < len; i++) { const int8_t b = bases[i]; if (b == '-' || b == '.' || b == 'X') bases[i] = 'N'; }
    }
    if (basesM) for (int i = 0; i < len; i++) { const int8_t b = bases[len - 1 - i]; basesM[i] = b < 0 ? -1 : g_compExt[(int)b]; }
    return junk;
}

void orc_ingest_batch(int8_t* bases, int8_t* quality, const int64_t* read_off, int64_t nreads, int flags, int8_t* basesM, int32_t* readFlags) {
    for (int64_t r = 0; r < nreads; r++) {
        const int64_t o = read_off[r]; const int len = (int)(read_off[r + 1] - o);
        const int j = orc_ingest_read(bases + o, quality ? quality + o : 0, len, flags, basesM ? basesM + o : 0);
        if (readFlags) readFlags[r] = j;
    }
}
