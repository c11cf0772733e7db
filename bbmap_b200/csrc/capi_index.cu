// capi_index.cu — k-mer index build/analysis and the index search
// Part of the C ABI of libbbmapcuda.so (include/bbmap_cuda.h): host-side glue only (device buffers, streams, launches).
// No CPU implementation of any compute path lives here: without a device every call fails loudly.
#include "ctx.h"

// =====================  k-mer index build + analysis  =====================
void index_free(bbm_ctx* c) {
    c->map_sites_hint = 0;
    if (c->index_shared) {          // borrowed from another context (bbm_index_share): nothing to free here
        c->iblocks.clear(); c->d_counts = nullptr; c->has_index = false; c->d_icfg = c->d_iblocks = nullptr; c->d_ihist = nullptr; c->d_chrom_off = nullptr;
        c->index_shared = false; return;
    }
    for (auto& b : c->iblocks) { if (b.starts) cudaFree(b.starts); if (b.sites) cudaFree(b.sites); }
    c->iblocks.clear();
    if (c->d_counts) cudaFree(c->d_counts);
    c->d_counts = nullptr; c->has_index = false;
    if (c->d_icfg) cudaFree(c->d_icfg); if (c->d_iblocks) cudaFree(c->d_iblocks); if (c->d_ihist) cudaFree(c->d_ihist); if (c->d_chrom_off) cudaFree(c->d_chrom_off);
    c->d_icfg = c->d_iblocks = nullptr; c->d_ihist = nullptr; c->d_chrom_off = nullptr;
}

static void index_cfg_init(bbm_index_cfg* c, int k, int chrombits, long long numDefinedBases) {
    // BBIndex statics (current/align2/BBIndex.java:3169-3262) + the small-genome retune of BBMap.loadIndex (BBMap.java:367-382)
    memset(c, 0, sizeof(*c));
    c->keylen = k; c->chrombits = chrombits;
    c->max_hits_reduction2 = 2; c->maximum_max_hits_reduction = 3; c->hit_reduction_div = 5;
    float f = 0.03f;
    if (numDefinedBases < 300000000LL) {
        c->max_hits_reduction2 += 1; c->maximum_max_hits_reduction += 1;
        if (numDefinedBases < 30000000LL) { f = f * 0.5f; c->maximum_max_hits_reduction += 1; c->hit_reduction_div = std::max(c->hit_reduction_div - 1, 3); }
        else if (numDefinedBases < 100000000LL) f = f * 0.6f;
        else f = f * 0.75f;
    }
    c->fraction_to_exclude = f;
    c->min_index_to_drop_long_hit_list = (int)(1000 * (1 - 3.5 * f));      // setFractionToExclude: double arithmetic
    c->max_average_list_to_search = (int)(1000 * (1 - 2.3 * f));
    c->max_average_list_to_search2 = (int)(1000 * (1 - 1.4 * f));
    c->max_single_list_to_search = (int)(1000 * (1 - 1.0 * f));
    c->max_shortest_list_to_search = (int)(1000 * (1 - 2.8 * f));
    c->shift_length = 32 - 1 - chrombits;
    c->chroms_per_block = 1 << chrombits;
}

static int index_finalize(bbm_ctx* c, const int8_t* d_chroms, const int64_t* chrom_off, int32_t nchroms, cudaStream_t st);

// chrombits (automatic rule BBMap.java:317-321), the BBIndex statics for this genome; leaves the context without blocks
static int index_begin(bbm_ctx* c, const int8_t* d_chroms, const int64_t* chrom_off, int32_t nchroms, int32_t keylen, int32_t chrombits, cudaStream_t st) {
    index_free(c);
    long long maxLen = 0, total = chrom_off[nchroms] - chrom_off[0];
    for (int i = 0; i < nchroms; ++i) maxLen = std::max<long long>(maxLen, chrom_off[i + 1] - chrom_off[i]);
    if (chrombits < 0) { int nlz = maxLen == 0 ? 32 : __builtin_clz((unsigned)maxLen); chrombits = std::min(nlz - 1, 16); }
    if (maxLen - 1 > (long long)(~((-1) << (32 - 1 - chrombits)))) return fail(BBM_E_ARG, "bbm_index_build: chromosome longer than MAX_ALLOWED_CHROM_INDEX for these chrombits");
    unsigned long long* d_def = nullptr; CK(cudaMalloc(&d_def, 8)); CK(cudaMemsetAsync(d_def, 0, 8, st));
    int e = bbm_index_count_defined(d_chroms + chrom_off[0], total, d_def, st);
    if (e) return fail(BBM_E_CUDA, "count_defined", (cudaError_t)e);
    unsigned long long nDefined = 0; CK(cudaMemcpyAsync(&nDefined, d_def, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st)); cudaFree(d_def);
    c->launches++;
    index_cfg_init(&c->icfg, keylen, chrombits, (long long)nDefined);
    return BBM_OK;
}

extern "C" int bbm_index_build(bbm_ctx* c, const int8_t* d_chroms, const int64_t* chrom_off, int32_t nchroms, int32_t keylen, int32_t chrombits,
                               bbm_index_cfg* cfg_out, int32_t* nblocks_out) {
    if (!c || !d_chroms || !chrom_off || nchroms < 1 || keylen < 8 || keylen > 15) return fail(BBM_E_ARG, "bbm_index_build: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const auto t_build0 = std::chrono::steady_clock::now();
    if (int rc = index_begin(c, d_chroms, chrom_off, nchroms, keylen, chrombits, st)) return rc;
    const int k = keylen;
    const long long keyspace = 1LL << (2 * k);
    int e = 0;
    const int cpb = c->icfg.chroms_per_block, low = cpb - 1, shift = c->icfg.shift_length;
    // blocks: chromosomes sharing (chrom & ~low); chrom numbers are 1-based (IndexMaker4.makeIndex :44-62)
    for (int chrom = 1; chrom <= nchroms;) {
        const int a = std::max(1, chrom & ~low), b = std::min(nchroms, (chrom & ~low) + cpb - 1);
        bbm_ctx::IndexBlock B; B.minChrom = a; B.maxChrom = b;
        long long n = 0;
        for (int ch = a; ch <= b; ++ch) n += chrom_off[ch] - chrom_off[ch - 1];
        unsigned *k0 = nullptr, *k1 = nullptr; int *v0 = nullptr, *v1 = nullptr, *sizes = nullptr;
        CK(cudaMalloc(&k0, (size_t)n * 4 + 16)); CK(cudaMalloc(&k1, (size_t)n * 4 + 16)); CK(cudaMalloc(&v0, (size_t)n * 4 + 16)); CK(cudaMalloc(&v1, (size_t)n * 4 + 16));
        CK(cudaMalloc(&sizes, (size_t)(keyspace + 1) * 4)); CK(cudaMemsetAsync(sizes, 0, (size_t)(keyspace + 1) * 4, st));
        CK(cudaMalloc(&B.starts, (size_t)(keyspace + 1) * 4));
        const unsigned invalid = 1u << (2 * k);
        long long base = 0;
        for (int ch = a; ch <= b; ++ch) {
            const int len = (int)(chrom_off[ch] - chrom_off[ch - 1]);
            e = bbm_index_emit(d_chroms + chrom_off[ch - 1], len, k, (ch & low) << shift, k0, v0, base, sizes, invalid, st);
            if (e) return fail(BBM_E_CUDA, "index_emit_kernel", (cudaError_t)e);
            c->launches++;
            base += len;
        }
        size_t tb1 = 0, tb2 = 0;
        bbm_index_sort_pairs(nullptr, &tb1, k0, k1, v0, v1, n, 2 * k + 1, st);
        bbm_index_scan(nullptr, &tb2, sizes, B.starts, keyspace + 1, st);
        void* temp = nullptr; CK(cudaMalloc(&temp, std::max(tb1, tb2) + 16));
        e = bbm_index_sort_pairs(temp, &tb1, k0, k1, v0, v1, n, 2 * k + 1, st);
        if (e) return fail(BBM_E_CUDA, "radix sort", (cudaError_t)e);
        e = bbm_index_scan(temp, &tb2, sizes, B.starts, keyspace + 1, st);
        if (e) return fail(BBM_E_CUDA, "scan", (cudaError_t)e);
        c->launches += 2;
        int nsites = 0;
        CK(cudaMemcpyAsync(&nsites, B.starts + keyspace, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        B.nsites = nsites;
        CK(cudaMalloc(&B.sites, (size_t)std::max(nsites, 1) * 4));
        CK(cudaMemcpyAsync(B.sites, v1, (size_t)nsites * 4, cudaMemcpyDeviceToDevice, st));      // valid pairs sort before the invalid key
        CK(cudaStreamSynchronize(st));
        cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(sizes); cudaFree(temp);
        c->iblocks.push_back(B);
        chrom = b + 1;
    }
    if (int rc = index_finalize(c, d_chroms, chrom_off, nchroms, st)) return rc;
    c->index_build_us = (long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t_build0).count();
    if (cfg_out) *cfg_out = c->icfg;
    if (nblocks_out) *nblocks_out = (int)c->iblocks.size();
    return BBM_OK;
}

// BBIndex.analyzeIndex over the blocks now resident (COUNTS, clumpy keys, lengthHistogram, the derived limits) + the device-side descriptors
static int index_finalize(bbm_ctx* c, const int8_t* d_chroms, const int64_t* chrom_off, int32_t nchroms, cudaStream_t st) {
    const int k = c->icfg.keylen;
    const long long keyspace = 1LL << (2 * k);
    int e = 0;
    // analyzeIndex
    unsigned long long* d_clump = nullptr; int* d_max = nullptr;
    CK(cudaMalloc(&c->d_counts, (size_t)keyspace * 4)); CK(cudaMemsetAsync(c->d_counts, 0, (size_t)keyspace * 4, st));
    CK(cudaMalloc(&d_clump, (size_t)keyspace * 8)); CK(cudaMemsetAsync(d_clump, 0, (size_t)keyspace * 8, st));
    CK(cudaMalloc(&d_max, 4)); CK(cudaMemsetAsync(d_max, 0, 4, st));
    for (auto& B : c->iblocks) { e = bbm_index_analyze_block(B.starts, B.sites, k, c->d_counts, d_clump, st); if (e) return fail(BBM_E_CUDA, "analyze_block", (cudaError_t)e); c->launches++; }
    e = bbm_index_finish_counts(k, c->d_counts, d_clump, d_max, st);
    if (e) return fail(BBM_E_CUDA, "finish_counts", (cudaError_t)e);
    c->launches += 3;
    int maxv = 0; CK(cudaMemcpyAsync(&maxv, d_max, 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    int* d_len = nullptr; CK(cudaMalloc(&d_len, (size_t)(maxv + 1) * 4)); CK(cudaMemsetAsync(d_len, 0, (size_t)(maxv + 1) * 4, st));
    e = bbm_index_lenhist(k, c->d_counts, d_len, st);
    if (e) return fail(BBM_E_CUDA, "lenhist", (cudaError_t)e);
    c->launches++;
    std::vector<int> lenCounts((size_t)maxv + 1);
    CK(cudaMemcpyAsync(lenCounts.data(), d_len, (size_t)(maxv + 1) * 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    cudaFree(d_clump); cudaFree(d_max); cudaFree(d_len);
    {   // Tools.makeLengthHistogram4 (Tools.java:1817-1850), buckets=1000, on the (small) histogram of list lengths
        long long tot = 0;
        for (int i = 1; i <= maxv; ++i) tot += (long long)i * lenCounts[i];
        long long sum = 0; int ptr = 0; const int buckets = 1000;
        for (int i = 0; i < buckets; ++i) {
            const long long nextLimit = ((tot * i) + buckets / 2) / buckets;
            while (ptr < maxv + 1 && sum < nextLimit) { sum += (int)(lenCounts[ptr] * ptr); ptr++; }
            c->ihist[i] = std::max(0, ptr - 1);
        }
        c->ihist[buckets] = maxv;
        const float f = c->icfg.fraction_to_exclude;
        c->icfg.max_usable_length = std::max(2 * 20, c->ihist[(int)((1 - f) * (1001 - 1))]);
        c->icfg.max_usable_length2 = std::max(6 * 20, c->ihist[(int)((1 - f * 0.25f) * (1001 - 1))]);
        int pps = (int)floor((double)((-50 * 4000.f) / std::max(2 * 20, c->ihist[c->icfg.max_average_list_to_search])));
        if (pps == 0) pps = -1;
        c->icfg.points_per_site = pps;
    }
    c->d_chroms = d_chroms; c->chrom_off.assign(chrom_off, chrom_off + nchroms + 1);
    {   // device-side descriptors for the search kernel
        struct Blk { const int* starts; const int* sites; };
        std::vector<Blk> hb;
        for (auto& B : c->iblocks) hb.push_back(Blk{B.starts, B.sites});
        CK(cudaMalloc(&c->d_icfg, sizeof(bbm_index_cfg))); CK(cudaMemcpy(c->d_icfg, &c->icfg, sizeof(bbm_index_cfg), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&c->d_iblocks, hb.size() * sizeof(Blk))); CK(cudaMemcpy(c->d_iblocks, hb.data(), hb.size() * sizeof(Blk), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&c->d_ihist, sizeof(c->ihist))); CK(cudaMemcpy(c->d_ihist, c->ihist, sizeof(c->ihist), cudaMemcpyHostToDevice));
        std::vector<long long> rel(nchroms + 1);
        for (int i = 0; i <= nchroms; ++i) rel[i] = chrom_off[i];
        CK(cudaMalloc(&c->d_chrom_off, rel.size() * 8)); CK(cudaMemcpy(c->d_chrom_off, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice));
    }
    CK(cudaStreamSynchronize(st));
    c->has_index = true;
    return BBM_OK;
}

// ---- index persistence in the reference's own on-disk format (formats in wire.cpp) ----
// IndexMaker4.makeIndex writes every block it builds with Block.write (IndexMaker4.java:197-200); BBIndex/IndexMaker4 load it back with
// Block.read when both files exist (:135-139).  root_index is the reference's Data.ROOT_INDEX ("<path>/ref/index/"), `build` its genome build number.
extern "C" int bbm_index_save(bbm_ctx* c, const char* root_index, int32_t build) {
    if (!c || !c->has_index || !root_index) return fail(BBM_E_ARG, "bbm_index_save: no index in this context");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const long long nstarts = (1LL << (2 * c->icfg.keylen)) + 1;
    std::vector<int32_t> starts((size_t)nstarts), sites;
    for (const auto& B : c->iblocks) {
        char fname[4096];
        if (bbm_wire_block_fname(fname, sizeof fname, root_index, B.minChrom, B.maxChrom, c->icfg.keylen, c->icfg.chrombits, build)) return fail(BBM_E_ARG, bbm_wire_last_error());
        sites.resize((size_t)std::max<long long>(B.nsites, 1));
        CK(cudaMemcpy(starts.data(), B.starts, (size_t)nstarts * 4, cudaMemcpyDeviceToHost));
        if (B.nsites) CK(cudaMemcpy(sites.data(), B.sites, (size_t)B.nsites * 4, cudaMemcpyDeviceToHost));
        if (bbm_wire_write_block(fname, sites.data(), B.nsites, starts.data(), nstarts)) return fail(BBM_E_ARG, bbm_wire_last_error());
    }
    return BBM_OK;
}

extern "C" int bbm_index_load(bbm_ctx* c, const int8_t* d_chroms, const int64_t* chrom_off, int32_t nchroms, int32_t keylen, int32_t chrombits,
                              const char* root_index, int32_t build, bbm_index_cfg* cfg_out, int32_t* nblocks_out) {
    if (!c || !d_chroms || !chrom_off || nchroms < 1 || keylen < 8 || keylen > 15 || !root_index) return fail(BBM_E_ARG, "bbm_index_load: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    if (int rc = index_begin(c, d_chroms, chrom_off, nchroms, keylen, chrombits, st)) return rc;
    const int cpb = c->icfg.chroms_per_block, low = cpb - 1;
    const long long nstartsWant = (1LL << (2 * keylen)) + 1;
    for (int chrom = 1; chrom <= nchroms;) {
        const int a = std::max(1, chrom & ~low), b = std::min(nchroms, (chrom & ~low) + cpb - 1);
        char fname[4096];
        if (bbm_wire_block_fname(fname, sizeof fname, root_index, a, b, keylen, c->icfg.chrombits, build)) { index_free(c); return fail(BBM_E_ARG, bbm_wire_last_error()); }
        int32_t *sites = nullptr, *starts = nullptr; int64_t nsites = 0, nstarts = 0;
        if (bbm_wire_read_block(fname, &sites, &nsites, &starts, &nstarts)) { index_free(c); return fail(BBM_E_ARG, bbm_wire_last_error()); }
        bbm_ctx::IndexBlock B; B.minChrom = a; B.maxChrom = b; B.nsites = nsites;
        cudaError_t ce = cudaSuccess;
        if (nstarts != nstartsWant) ce = cudaErrorInvalidValue;
        if (ce == cudaSuccess) ce = cudaMalloc(&B.starts, (size_t)nstarts * 4);
        if (ce == cudaSuccess) ce = cudaMalloc(&B.sites, (size_t)std::max<int64_t>(nsites, 1) * 4);
        if (ce == cudaSuccess) ce = cudaMemcpy(B.starts, starts, (size_t)nstarts * 4, cudaMemcpyHostToDevice);
        if (ce == cudaSuccess && nsites) ce = cudaMemcpy(B.sites, sites, (size_t)nsites * 4, cudaMemcpyHostToDevice);
        bbm_wire_free(sites); bbm_wire_free(starts);
        if (ce != cudaSuccess) {
            if (B.starts) cudaFree(B.starts); if (B.sites) cudaFree(B.sites);
            index_free(c);
            return nstarts != nstartsWant ? fail(BBM_E_ARG, "bbm_index_load: the block on disk was built with another key length") : fail(BBM_E_CUDA, "bbm_index_load: upload", ce);
        }
        c->iblocks.push_back(B);
        chrom = b + 1;
    }
    if (int rc = index_finalize(c, d_chroms, chrom_off, nchroms, st)) return rc;
    if (cfg_out) *cfg_out = c->icfg;
    if (nblocks_out) *nblocks_out = (int)c->iblocks.size();
    return BBM_OK;
}

extern "C" int bbm_index_block_sites(bbm_ctx* c, int32_t block, int64_t* nsites_out) {
    if (!c || !c->has_index || block < 0 || block >= (int)c->iblocks.size() || !nsites_out) return fail(BBM_E_ARG, "bbm_index_block_sites: bad argument");
    *nsites_out = c->iblocks[block].nsites;
    return BBM_OK;
}

extern "C" int bbm_index_download(bbm_ctx* c, int32_t block, int32_t* starts, int32_t* sites, int32_t* counts, int32_t* hist1001) {
    if (!c || !c->has_index || block < 0 || block >= (int)c->iblocks.size()) return fail(BBM_E_ARG, "bbm_index_download: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const long long keyspace = 1LL << (2 * c->icfg.keylen);
    const auto& B = c->iblocks[block];
    if (starts) CK(cudaMemcpy(starts, B.starts, (size_t)(keyspace + 1) * 4, cudaMemcpyDeviceToHost));
    if (sites && B.nsites) CK(cudaMemcpy(sites, B.sites, (size_t)B.nsites * 4, cudaMemcpyDeviceToHost));
    if (counts) CK(cudaMemcpy(counts, c->d_counts, (size_t)keyspace * 4, cudaMemcpyDeviceToHost));
    if (hist1001) memcpy(hist1001, c->ihist, sizeof(c->ihist));
    return BBM_OK;
}


// =====================  index search (BBIndex.find)  =====================
int run_search(bbm_ctx* c, const int8_t* db, const int8_t* dbs, const int64_t* doff, int64_t nreads, const int* dn, const int* dof,
                      const int* dks, int maxKeys, int quit2, bbm_search_head* dh, bbm_site* ds, int maxSites, int maxReadLen, cudaStream_t st, float* ms_out) {
    if (!c->has_index) return fail(BBM_E_ARG, "bbm_search: no index in this context (call bbm_index_build first)");
    if (nreads <= 0) { if (ms_out) *ms_out = 0.f; return BBM_OK; }
    if ((int)c->iblocks.size() > 64) return fail(BBM_E_SHAPE, "bbm_search: more than 64 index blocks");
    const int T = bbm_search_threads();
    int blocks = c->sms * 8;
    const long long need = (nreads + T - 1) / T;
    if (need < blocks) blocks = (int)need;
    (void)maxReadLen;     // reserved: lets a later kernel size its per-read working set to the batch
    if (c->searchCtx.ensure((size_t)c->sms * 8 * T * bbm_search_pool_bytes())) return fail(BBM_E_CUDA, "cudaMalloc search scratch");
    unsigned int* cb = (unsigned int*)c->counters.p;
    if (c->search_prof) CK(cudaMemsetAsync(cb + 208, 0, 40, st));
    unsigned long long* prof = c->search_prof ? (unsigned long long*)(cb + 208) : nullptr;
    const int nblk = (int)c->iblocks.size(), nchr = (int)c->chrom_off.size() - 1;
    CK(cudaEventRecord(c->ev0, st));
    if (c->search_split) {
        // one phase of BBIndex.find per launch (key filtering, prescan, walk): all lanes of a warp run the same phase
        const int stride = bbm_search_mid_stride(maxKeys, nblk);
        if (c->searchRev.ensure((size_t)nreads * stride * 4)) return fail(BBM_E_CUDA, "cudaMalloc search phase state");
        for (int ph = 1; ph <= 4; ph <<= 1) {
            if (ph == 2 && c->search_split >= 2) {
                // prescan with one warp per read; reads with more than 32 keys are left to the thread-per-read launch that follows
                CK(cudaMemsetAsync(cb + 202, 0, 4, st));
                int wblocks = c->sms * 8; const long long wneed = (nreads + 3) / 4; if (wneed < wblocks) wblocks = (int)wneed;
                int e = bbm_launch_search_prescan_warp((const bbm_index_cfg*)c->d_icfg, c->d_iblocks, nblk, nchr, c->d_counts, (const long long*)doff, nreads, dn,
                                                       maxKeys, dh, cb + 202, wblocks, (int*)c->searchRev.p, stride, st);
                if (e) return fail(BBM_E_CUDA, "prescan_warp_kernel launch", (cudaError_t)e);
                c->launches++;
                if (maxKeys <= 32) continue;
            }
            if (ph == 4 && c->search_split >= 3) {
                // slowWalk3 / extendScore with one warp per read; reads with more than 32 keys are left to the thread-per-read launch that follows
                CK(cudaMemsetAsync(cb + 202, 0, 4, st));
                int wblocks = c->sms * 8; const long long wneed = (nreads + 3) / 4; if (wneed < wblocks) wblocks = (int)wneed;
                int e = bbm_launch_search_walk_warp((const bbm_index_cfg*)c->d_icfg, c->d_iblocks, nblk, nchr, c->d_counts, c->d_chroms, c->d_chrom_off, db, dbs,
                                                    (const long long*)doff, nreads, dn, maxKeys, quit2, dh, ds, maxSites, cb + 202, wblocks, (int*)c->searchRev.p, stride, st);
                if (e) return fail(BBM_E_CUDA, "walk_warp_kernel launch", (cudaError_t)e);
                c->launches++;
                if (maxKeys <= 32) continue;
            }
            CK(cudaMemsetAsync(cb + 202, 0, 4, st));
            int e = bbm_launch_search((const bbm_index_cfg*)c->d_icfg, c->d_iblocks, nblk, nchr, c->d_counts, c->d_ihist, c->d_chroms, c->d_chrom_off, db, dbs,
                                      (const long long*)doff, nreads, dn, dof, dks, maxKeys, quit2, dh, ds, maxSites, c->searchCtx.p, cb + 202, prof, blocks,
                                      c->search_shared ? 0 : 1, ph, (int*)c->searchRev.p, stride, st);
            if (e) return fail(BBM_E_CUDA, "search_kernel launch", (cudaError_t)e);
            c->launches++;
        }
    } else {
        CK(cudaMemsetAsync(cb + 202, 0, 4, st));
        int e = bbm_launch_search((const bbm_index_cfg*)c->d_icfg, c->d_iblocks, nblk, nchr, c->d_counts, c->d_ihist, c->d_chroms, c->d_chrom_off, db, dbs,
                                  (const long long*)doff, nreads, dn, dof, dks, maxKeys, quit2, dh, ds, maxSites, c->searchCtx.p, cb + 202, prof, blocks,
                                  c->search_shared ? 0 : 1, 7, nullptr, 0, st);
        if (e) return fail(BBM_E_CUDA, "search_kernel launch", (cudaError_t)e);
        c->launches++;
    }
    CK(cudaEventRecord(c->ev1, st));
    CK(cudaStreamSynchronize(st));
    if (ms_out) { float ms = 0.f; CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1)); *ms_out = ms; }
    if (c->search_prof) CK(cudaMemcpy(c->search_cycles, cb + 208, 40, cudaMemcpyDeviceToHost));
    return BBM_OK;
}

extern "C" int bbm_search_batch_dev(bbm_ctx* c, const int8_t* d_bases, const int8_t* d_baseScores, const int64_t* d_read_off, int64_t nreads,
                                    const int32_t* d_nkeys, const int32_t* d_offsets, const int32_t* d_keyScores, int32_t maxKeys,
                                    int32_t quit2, bbm_search_head* d_heads, bbm_site* d_sites, int32_t max_sites, int32_t max_read_len, void* stream, float* kernel_ms_out) {
    if (!c || !d_bases || !d_baseScores || !d_read_off || !d_nkeys || !d_offsets || !d_keyScores || !d_heads || !d_sites || max_sites < 1 || maxKeys < 1 || maxKeys > 96)
        return fail(BBM_E_ARG, "bbm_search_batch_dev: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    return run_search(c, d_bases, d_baseScores, d_read_off, nreads, d_nkeys, d_offsets, d_keyScores, maxKeys, quit2, d_heads, d_sites, max_sites,
                      max_read_len, stream ? (cudaStream_t)stream : c->stream, kernel_ms_out);
}

extern "C" int bbm_search_batch_host(bbm_ctx* c, const int8_t* bases, const int8_t* baseScores, const int64_t* read_off, int64_t nreads,
                                     const int32_t* nkeys, const int32_t* offsets, const int32_t* keyScores, int32_t maxKeys,
                                     int32_t quit2, bbm_search_head* heads, bbm_site* sites, int32_t max_sites) {
    if (!c || !bases || !baseScores || !read_off || !nkeys || !offsets || !keyScores || !heads || !sites || max_sites < 1 || maxKeys < 1 || maxKeys > 96)
        return fail(BBM_E_ARG, "bbm_search_batch_host: bad argument");
    if (nreads <= 0) return BBM_OK;
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t nb = (size_t)read_off[nreads], kb = (size_t)nreads * maxKeys * 4;
    int max_len = 1;
    for (int64_t i = 0; i < nreads; ++i) { const int64_t l = read_off[i + 1] - read_off[i]; if (l > max_len) max_len = (int)(l > 100000 ? 100000 : l); }
    const size_t hb = (size_t)nreads * sizeof(bbm_search_head), sb = (size_t)nreads * max_sites * sizeof(bbm_site);
    DevBuf* B = c->d_srch;   // 0 bases, 1 baseScores, 2 off, 3 nkeys, 4 offsets, 5 keyScores, 6 heads, 7 sites
    if (B[0].ensure(nb + 32) || B[1].ensure(nb + 32) || B[2].ensure((size_t)(nreads + 1) * 8) || B[3].ensure((size_t)nreads * 4) || B[4].ensure(kb) ||
        B[5].ensure(kb) || B[6].ensure(hb) || B[7].ensure(sb))
        return fail(BBM_E_CUDA, "cudaMalloc staging");
    CK(cudaMemcpyAsync(B[0].p, bases, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[1].p, baseScores, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[2].p, read_off, (size_t)(nreads + 1) * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[3].p, nkeys, (size_t)nreads * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[4].p, offsets, kb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(B[5].p, keyScores, kb, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(B[7].p, 0, sb, st));
    int rc = run_search(c, (const int8_t*)B[0].p, (const int8_t*)B[1].p, (const int64_t*)B[2].p, nreads, (const int*)B[3].p, (const int*)B[4].p,
                        (const int*)B[5].p, maxKeys, quit2, (bbm_search_head*)B[6].p, (bbm_site*)B[7].p, max_sites, max_len, st, nullptr);
    if (rc) return rc;
    CK(cudaMemcpyAsync(heads, B[6].p, hb, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(sites, B[7].p, sb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return BBM_OK;
}


// A second context on the same device that maps against the SAME resident index and reference (no copy): the reference keeps one index per
// process and one MSA per mapping thread (AbstractMapThread.java:133-136); here a context is the unit that owns scratch buffers and a stream, so
// several batches can be in flight against one index.  `src` must outlive `dst`.
extern "C" int bbm_index_share(bbm_ctx* dst, bbm_ctx* src) {
    if (!dst || !src || dst == src) return fail(BBM_E_ARG, "bbm_index_share: bad argument");
    if (!src->has_index) return fail(BBM_E_ARG, "bbm_index_share: the source context has no index");
    if (dst->device != src->device) return fail(BBM_E_ARG, "bbm_index_share: contexts live on different devices");
    std::lock_guard<std::mutex> lk(dst->mu);
    index_free(dst);
    dst->iblocks = src->iblocks; dst->d_counts = src->d_counts; memcpy(dst->ihist, src->ihist, sizeof(dst->ihist)); dst->icfg = src->icfg;
    dst->d_chroms = src->d_chroms; dst->chrom_off = src->chrom_off; dst->d_icfg = src->d_icfg; dst->d_iblocks = src->d_iblocks; dst->d_ihist = src->d_ihist;
    dst->d_chrom_off = src->d_chrom_off; dst->has_index = true; dst->index_shared = true;
    return BBM_OK;
}
