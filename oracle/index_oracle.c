/*
 * index_oracle.c — TEST INFRASTRUCTURE ONLY.
 * C restatement of BBMap's k-mer index build and analysis, Java-only in the reference:
 *   IndexMaker4.BlockMaker / CountThread   current/align2/IndexMaker4.java:160-421   (counting sort into Block{starts,sites})
 *   site codec, chrom bits                 current/align2/BBIndex.java:3036-3057, 3148-3164
 *   BBIndex.analyzeIndex                   current/align2/BBIndex.java:101-191     (COUNTS, clumpy keys, lengthHistogram, limits)
 *   Tools.makeLengthHistogram3/4           current/align2/Tools.java:1797-1850
 *   small-genome retune                    current/align2/BBMap.java:367-382, BBIndex.setFractionToExclude :3201-3209
 * PARITY UNPINNED against Java (no JVM); pinned only by structural invariants (lists sorted by (chrom,pos), banned
 * period-<=2 keys, COUNTS symmetric under reverse complement).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "index_oracle.h"

extern int orc_rcomp_key_fast(int kmer, int k);

static int b2n(int c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': case 'U': case 'u': return 3; default: return -1; } }

static int to_number(const int8_t* bases, int a, int b) {
    int out = 0;
    for (int i = a; i <= b; i++) { const int x = b2n(bases[i]); if (x < 0) return -1; out = (out << 2) | x; }
    return out;
}

void orc_index_cfg_init(orc_index_cfg* c, int k, int chrombits, int64_t numDefinedBases) {
    /* BBIndex statics + BBMap.loadIndex retune (BBMap.java:367-382) */
    memset(c, 0, sizeof(*c));
    c->keylen = k; c->chrombits = chrombits;
    c->max_hits_reduction2 = 2; c->maximum_max_hits_reduction = 3; c->hit_reduction_div = 5;
    float f = 0.03f;
    if (numDefinedBases < 300000000LL) {
        c->max_hits_reduction2 += 1; c->maximum_max_hits_reduction += 1;
        if (numDefinedBases < 30000000LL) { f = f * 0.5f; c->maximum_max_hits_reduction += 1; c->hit_reduction_div = c->hit_reduction_div - 1 > 3 ? c->hit_reduction_div - 1 : 3; }
        else if (numDefinedBases < 100000000LL) f = f * 0.6f;
        else f = f * 0.75f;
    }
    c->fraction_to_exclude = f;
    /* setFractionToExclude (BBIndex.java:3201-3209): double arithmetic, truncation */
    c->min_index_to_drop_long_hit_list = (int)(1000 * (1 - 3.5 * f));
    c->max_average_list_to_search = (int)(1000 * (1 - 2.3 * f));
    c->max_average_list_to_search2 = (int)(1000 * (1 - 1.4 * f));
    c->max_single_list_to_search = (int)(1000 * (1 - 1.0 * f));
    c->max_shortest_list_to_search = (int)(1000 * (1 - 2.8 * f));
    c->shift_length = 32 - 1 - chrombits;
    c->chroms_per_block = 1 << chrombits;
}

int orc_auto_chrombits(const int64_t* chrom_off, int nchroms) {
    /* BBMap.java:317-321: numberOfLeadingZeros(max chromLength)-1, capped at 16 */
    int64_t maxLen = 0;
    for (int i = 0; i < nchroms; i++) { const int64_t l = chrom_off[i + 1] - chrom_off[i]; if (l > maxLen) maxLen = l; }
    int nlz = 0; uint32_t v = (uint32_t)maxLen;
    if (v == 0) nlz = 32; else { while (!(v & 0x80000000u)) { v <<= 1; nlz++; } }
    int bits = nlz - 1;
    return bits < 16 ? bits : 16;
}

/* Build one block covering chromosomes [minChrom,maxChrom] (1-based).  starts has 4^k+1 entries, sites is malloc'd. */
int64_t orc_index_build_block(const int8_t* chroms, const int64_t* chrom_off, int minChrom, int maxChrom, const orc_index_cfg* c,
                              int32_t* starts, int32_t** sites_out) {
    const int k = c->keylen, skip = k - 1;
    const int64_t keyspace = 1LL << (2 * k);
    const int banshift = 4; const int banmask = ~((-1) << ((2 * k) - banshift));
    const int lowmask = c->chroms_per_block - 1;
    int32_t* sizes = (int32_t*)calloc((size_t)keyspace + 1, sizeof(int32_t));
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            int32_t sum = 0;
            for (int64_t i = 0; i <= keyspace; i++) { const int32_t t = sizes[i]; sizes[i] = sum; sum += t; }
            *sites_out = (int32_t*)calloc((size_t)(sum > 0 ? sum : 1), sizeof(int32_t));
        }
        for (int chrom = minChrom; chrom <= maxChrom; chrom++) {
            const int8_t* array = chroms + chrom_off[chrom - 1];
            const int maxIndex = (int)(chrom_off[chrom] - chrom_off[chrom - 1]) - 1;
            const int max = maxIndex - k + 1;
            for (int a = 0, b = skip; a < max; a++, b++) {
                const int first = array[a];
                if (first != 'A' && first != 'C' && first != 'G' && first != 'T') continue;      /* array[a]==idb of one of the 4 CountThreads */
                const int key = to_number(array, a, b);
                if (key >= 0 && (key >> banshift) != (key & banmask)) {
                    if (pass == 0) sizes[key]++;
                    else { (*sites_out)[sizes[key]++] = ((chrom & lowmask) << c->shift_length) | a; }
                }
            }
        }
    }
    /* after the fill pass sizes[key] = end of list key = start of key+1: shift back (IndexMaker4.java:242-245) */
    for (int64_t i = keyspace - 1; i >= 0; i--) sizes[i + 1] = sizes[i];
    sizes[0] = 0;
    memcpy(starts, sizes, sizeof(int32_t) * ((size_t)keyspace + 1));
    const int64_t n = sizes[keyspace];
    free(sizes);
    return n;
}

static int imax_(int a, int b) { return a > b ? a : b; }

/* analyzeIndex over all blocks: COUNTS[4^k], lengthHistogram[1001], MAX_USABLE_LENGTH(2), POINTS_PER_SITE */
void orc_index_analyze(int nblocks, int32_t* const* starts, int32_t* const* sites, orc_index_cfg* c, int32_t* COUNTS, int32_t* hist1001) {
    const int k = c->keylen; const int64_t keyspace = 1LL << (2 * k);
    memset(COUNTS, 0, sizeof(int32_t) * (size_t)keyspace);
    int64_t* clump = (int64_t*)calloc((size_t)keyspace, sizeof(int64_t));     /* cmap keyed by min(key,rkey) */
    for (int b = 0; b < nblocks; b++) {
        const int32_t* st = starts[b]; const int32_t* si = sites[b];
        for (int64_t key = 0; key < keyspace; key++) {
            const int start1 = st[key], stop1 = st[key + 1], len1 = stop1 - start1;
            const int64_t t = (int64_t)COUNTS[key] + len1;
            COUNTS[key] = (int32_t)(t > 2147483647LL ? 2147483647LL : t);
            int64_t clumps = 0;
            for (int i = start1 + 1; i < stop1; i++) { const int dif = si[i] - si[i - 1]; if (dif > 0 && dif <= 5) clumps++; }
            if (clumps > 0) { const int r = orc_rcomp_key_fast((int)key, k); clump[key < r ? key : r] += clumps; }
        }
    }
    for (int64_t key = 0; key < keyspace; key++) {
        const int rkey = orc_rcomp_key_fast((int)key, k);
        if (key < rkey) {
            const int64_t x = (int64_t)COUNTS[key] + (int64_t)COUNTS[rkey];
            COUNTS[key] = COUNTS[rkey] = (int32_t)(x > 2147483647LL ? 2147483647LL : x);
        }
    }
    for (int64_t key = 0; key < keyspace; key++) {
        const int64_t clumps = clump[key];
        if (clumps > 0) {
            const int64_t len = COUNTS[key];
            if (len > 2000 && (float)clumps > 0.75f * (float)len) { const int rkey = orc_rcomp_key_fast((int)key, k); COUNTS[key] = 0; COUNTS[rkey] = 0; }
        }
    }
    free(clump);
    /* Tools.makeLengthHistogram3 -> 4 with buckets=1000 */
    int max = 0;
    for (int64_t i = 0; i < keyspace; i++) if (COUNTS[i] > max) max = COUNTS[i];
    int32_t* counts = (int32_t*)calloc((size_t)max + 1, sizeof(int32_t));
    int64_t total = 0;
    for (int64_t i = 0; i < keyspace; i++) { const int a = COUNTS[i]; if (a >= 0) { counts[a]++; total += a; } }
    if (total <= 0) { total = 0; for (int i = 1; i <= max; i++) total += (int64_t)(i * counts[i]); }
    const int buckets = 1000;
    int64_t sum = 0; int ptr = 0;
    for (int i = 0; i < buckets; i++) {
        const int64_t nextLimit = ((total * i) + buckets / 2) / buckets;
        while (ptr < max + 1 && sum < nextLimit) { sum += (int32_t)(counts[ptr] * ptr); ptr++; }
        hist1001[i] = imax_(0, ptr - 1);
    }
    hist1001[buckets] = max;
    free(counts);
    /* limits (BBIndex.java:168-190) */
    const float f = c->fraction_to_exclude;
    const int idx1 = (int)((1 - f) * (1001 - 1));
    const int idx2 = (int)((1 - f * 0.25f) * (1001 - 1));
    c->max_usable_length = imax_(2 * 20, hist1001[idx1]);
    c->max_usable_length2 = imax_(6 * 20, hist1001[idx2]);
    int pps = (int)floor((double)((-50 * 4000.f) / imax_(2 * 20, hist1001[c->max_average_list_to_search])));     /* Solver.BASE_POINTS_PER_SITE=-50 */
    if (pps == 0) pps = -1;
    c->points_per_site = pps;
}
