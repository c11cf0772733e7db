// capi_mapper.cu — bbm_map_batch_{dev,host}: the whole mapping chain behind one call (BBMapThread.processRead / processReadPair over a batch).
// Part of the C ABI of libbbmapcuda.so (include/bbmap_cuda.h): host-side glue only — it sequences the device stages, sizes their buffers and
// reads a handful of counters between rounds.  No CPU implementation of any compute path lives here.
#include "ctx.h"
#include "mapper_kernels.cuh"

using namespace bbm;

extern "C" int bbm_launch_sitelist(int op, bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const int8_t* basesP,
                                   const int8_t* basesM, const int8_t* refs, const long long* chrom_off, const bbm_policy_cfg* cfg, bbm_read_out* out,
                                   cudaStream_t st);
extern "C" int bbm_launch_sitelist_from_search(const bbm_search_head* heads, const bbm_site* sites, long long nreads, int maxSites, bbm_ss* lists,
                                               int* nss, int cap, cudaStream_t st);
extern "C" int bbm_launch_sitelist_tipdel(bbm_ss* lists, const int* nss, long long nreads, int cap, const long long* read_off, const int8_t* basesP,
                                          const int8_t* basesM, const int8_t* quality, const int8_t* refs, const long long* chrom_off,
                                          const int* chrom_min_index, const bbm_tipdel_cfg* tc, bbm_read_out* out, cudaStream_t st);
extern "C" int bbm_launch_sitelist_bounds(bbm_ss* lists, int* nss, long long nreads, int cap, const long long* read_off, const int* chrom_max_index,
                                          const int* scaf_off, const int* scaf_loc, int pad, int sam_out, int expected_len_limit, bbm_read_out* out, cudaStream_t st);
extern "C" int bbm_sitelist_max_cap();
extern "C" int bbm_launch_rescue(const int8_t* reads, const int8_t* refs, const bbm_rescue_task* tasks, long long n, const bbm_rescue_cfg* cfg,
                                 bbm_rescue_out* outs, cudaStream_t st);

namespace {
struct StageTimer {                 // host wall time per stage (every stage ends in a stream synchronisation or is followed by one)
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    float lap(cudaStream_t st) { cudaStreamSynchronize(st); auto t1 = std::chrono::steady_clock::now(); const float ms = std::chrono::duration<float, std::milli>(t1 - t0).count(); t0 = t1; return ms; }
};
enum { MB_BASESM = 0, MB_RFLAGS, MB_NKEYS, MB_OFFSETS, MB_KEYS, MB_KSCORES, MB_BSCORES, MB_OFFM, MB_KEYSM, MB_HEADS, MB_SITES, MB_LISTS, MB_NSS, MB_OUT, MB_OUT2,
       MB_RUN, MB_MASKED, MB_SLOWST, MB_GMSTATE, MB_MSLOTS, MB_MLEN, MB_TASKS, MB_OUTS, MB_RMATCH, MB_RMOFF, MB_GTASKS, MB_GAPS, MB_GOUTS, MB_GMATCH, MB_GMOFF,
       MB_COUNTERS, MB_STASKS, MB_CIGAR, MB_CIGOFF, MB_RECS, MB_SAM, MB_TLENS, MB_TNM, MB_TLINEOFF, MB_TSCAN, MB_TEXT, MB_TEXTOFF,
       MB_PSTATE, MB_RFLAGS2, MB_RTASKS, MB_ROUTS, MB_RAUX, MB_RSITES, MB_RTASKOF, MB_PSTATS, MB_COUNT };
}

static_assert(MB_COUNT <= 64, "bbm_ctx::mapBuf is too small");

#define LAUNCH(call, what) do { int e_ = (call); if (e_) return fail(BBM_E_CUDA, what, (cudaError_t)e_); c->launches++; } while (0)

extern "C" int bbm_map_set_scaffolds(bbm_ctx* c, const int32_t* scaf_off, const int32_t* scaf_loc, const int32_t* scaf_len, int32_t nchroms,
                                     const int8_t* names_buf, const int64_t* name_off) {
    if (!c || !scaf_off || !scaf_loc || !scaf_len || nchroms < 1) return fail(BBM_E_ARG, "bbm_map_set_scaffolds: bad argument");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    const int ns = scaf_off[nchroms];
    if (c->mapScaf[0].ensure((size_t)(nchroms + 1) * 4) || c->mapScaf[1].ensure((size_t)ns * 4 + 4) || c->mapScaf[2].ensure((size_t)ns * 4 + 4))
        return fail(BBM_E_CUDA, "cudaMalloc scaffold table");
    CK(cudaMemcpy(c->mapScaf[0].p, scaf_off, (size_t)(nchroms + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->mapScaf[1].p, scaf_loc, (size_t)ns * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c->mapScaf[2].p, scaf_len, (size_t)ns * 4, cudaMemcpyHostToDevice));
    c->map_nchroms = nchroms; c->map_nscaf = ns;
    c->map_has_names = false;
    if (names_buf && name_off) {
        const size_t nb = (size_t)name_off[ns];
        if (c->mapScaf[3].ensure(nb + 16) || c->mapScaf[4].ensure((size_t)(ns + 1) * 8)) return fail(BBM_E_CUDA, "cudaMalloc scaffold names");
        CK(cudaMemcpy(c->mapScaf[3].p, names_buf, nb, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->mapScaf[4].p, name_off, (size_t)(ns + 1) * 8, cudaMemcpyHostToDevice));
        c->map_has_names = true;
    }
    return BBM_OK;
}

static int map_args(bbm_ctx* c, const bbm_mapper_cfg* cfg, int64_t nreads) {
    if (!c || !cfg) return fail(BBM_E_ARG, "bbm_map_batch: null argument");
    if (!c->has_index) return fail(BBM_E_ARG, "bbm_map_batch: no index in this context (call bbm_index_build first)");
    if (c->map_nchroms != (int)c->chrom_off.size() - 1) return fail(BBM_E_ARG, "bbm_map_batch: no scaffold table for this index (call bbm_map_set_scaffolds)");
    if (cfg->max_sites < 1 || cfg->max_sites > bbm_sitelist_max_cap() || cfg->max_keys < 1) return fail(BBM_E_ARG, "bbm_map_batch: max_sites must be in 1..64");
    if (cfg->map.paired && (nreads & 1)) return fail(BBM_E_ARG, "bbm_map_batch: paired input needs an even number of reads");
    if (nreads > 0x3fffffffLL) return fail(BBM_E_ARG, "bbm_map_batch: too many reads in one batch");
    return BBM_OK;
}

// the chain; everything device-resident; d_bases/d_quality are modified in place (Read.validate)
static int map_locked(bbm_ctx* c, int8_t* d_bases, int8_t* d_quality, const int64_t* d_off, int64_t n, int maxLen, const bbm_mapper_cfg* cfg,
                      bbm_map_rec* d_recs, bbm_sam_out* d_sam, int8_t* d_match, int64_t match_stride, cudaStream_t st, bbm_map_stats* stats,
                      int64_t totalBases, const int8_t* d_names, const int64_t* d_name_off) {
    DevBuf* B = c->mapBuf;
    bbm_map_stats S; memset(&S, 0, sizeof S); S.reads = n;
    if (n <= 0) { if (stats) *stats = S; return BBM_OK; }
    StageTimer T; StageTimer Tall;
    // key slots per read: the caller's figure (32 covers reads up to ~200 bp), raised for longer reads to 2 keys per k bases + 3 (the key density never
    // exceeds 1.9, AbstractMapThread.java:663-676), at most 96 — a read must never lose its seeds to the size of a scratch array
    const int maxK = std::max(cfg->max_keys, std::min(96, 2 * maxLen / std::max(1, (int)cfg->seed.keylen) + 3)), paired = cfg->map.paired;
    const long long* off = (const long long*)d_off;
    const int8_t* refs = c->d_chroms; const long long* chrom_off = c->d_chrom_off;
    const int nchroms = c->map_nchroms;
    // ChromosomeArray.maxIndex per chromosome array (cached with the index)
    if (c->mapScaf[5].cap == 0 || c->map_maxidx_for != nchroms) {
        std::vector<int> mx(nchroms);
        for (int i = 0; i < nchroms; i++) mx[i] = (int)(c->chrom_off[i + 1] - c->chrom_off[i]) - 1;
        if (c->mapScaf[5].ensure((size_t)nchroms * 4)) return fail(BBM_E_CUDA, "cudaMalloc maxIndex");
        CK(cudaMemcpy(c->mapScaf[5].p, mx.data(), (size_t)nchroms * 4, cudaMemcpyHostToDevice));
        c->map_maxidx_for = nchroms;
    }
    const size_t nb = (size_t)totalBases + 64;
    if (B[MB_BASESM].ensure(nb) || B[MB_RFLAGS].ensure(n * 4) || B[MB_NKEYS].ensure(n * 4) || B[MB_OFFSETS].ensure((size_t)n * maxK * 4) || B[MB_KEYS].ensure((size_t)n * maxK * 4) ||
        B[MB_KSCORES].ensure((size_t)n * maxK * 4) || B[MB_BSCORES].ensure(nb) || B[MB_OFFM].ensure((size_t)n * maxK * 4) || B[MB_KEYSM].ensure((size_t)n * maxK * 4) ||
        B[MB_HEADS].ensure((size_t)n * sizeof(bbm_search_head)) || B[MB_NSS].ensure(n * 4) || B[MB_OUT].ensure((size_t)n * sizeof(bbm_read_out)) ||
        B[MB_OUT2].ensure((size_t)n * sizeof(bbm_read_out)) || B[MB_RUN].ensure(n * 4) || B[MB_MASKED].ensure(n * 4) || B[MB_SLOWST].ensure(n * 4) || B[MB_COUNTERS].ensure(256))
        return fail(BBM_E_CUDA, "cudaMalloc mapper buffers");
    int8_t* basesM = (int8_t*)B[MB_BASESM].p;
    int* nkeys = (int*)B[MB_NKEYS].p; int* nss = (int*)B[MB_NSS].p;
    bbm_search_head* heads = (bbm_search_head*)B[MB_HEADS].p;
    bbm_read_out* out = (bbm_read_out*)B[MB_OUT].p; bbm_read_out* out2 = (bbm_read_out*)B[MB_OUT2].p;
    int* counters = (int*)B[MB_COUNTERS].p;

    // ---- Read.validate, KeyRing, BBIndex.find ----
    if (int rc = run_ingest(c, d_bases, d_quality, d_off, n, maxLen, cfg->ingest_flags, basesM, (int*)B[MB_RFLAGS].p, st, nullptr)) return rc;
    if (int rc = run_seed(c, d_bases, d_quality, d_off, n, maxLen, &cfg->seed, maxK, nkeys, (int*)B[MB_OFFSETS].p, (int*)B[MB_KEYS].p, (int*)B[MB_KSCORES].p,
                          (int8_t*)B[MB_BSCORES].p, (int*)B[MB_OFFM].p, (int*)B[MB_KEYSM].p, st, nullptr)) return rc;
    // the reference keeps an unbounded ArrayList; here a read that overflows its slots sends the batch through the search again with more.  The size a
    // batch ended up needing is remembered in the context, so the next batch against the same index starts there (human-scale repeats: 16 -> 64 slots
    // meant three searches per batch, 465 of 746 ms).
    int maxSites = std::max(cfg->max_sites, std::min(c->map_sites_hint, bbm_sitelist_max_cap()));
    for (;;) {
        if (B[MB_SITES].ensure((size_t)n * maxSites * sizeof(bbm_site))) return fail(BBM_E_CUDA, "cudaMalloc site slots");
        if (int rc = run_search(c, d_bases, (const int8_t*)B[MB_BSCORES].p, d_off, n, nkeys, (const int*)B[MB_OFFSETS].p, (const int*)B[MB_KSCORES].p, maxK,
                                paired ? 0 : 1, heads, (bbm_site*)B[MB_SITES].p, maxSites, maxLen, st, nullptr)) return rc;
        CK(cudaMemsetAsync(counters, 0, 4, st));
        LAUNCH(bbm_launch_map_overflow(heads, n, maxSites, counters, st), "map_overflow_kernel launch");
        int over = 0;
        CK(cudaMemcpyAsync(&over, counters, 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        if (over == 0 || maxSites >= bbm_sitelist_max_cap()) break;
        maxSites = maxSites * 2 > bbm_sitelist_max_cap() ? bbm_sitelist_max_cap() : maxSites * 2;
    }
    const int cap = maxSites;
    S.max_sites_used = maxSites;
    c->map_sites_hint = maxSites > cfg->max_sites ? maxSites : 0;
    S.ms_seed_search = T.lap(st);

    // ---- site lists up to the final policy ----
    if (B[MB_LISTS].ensure((size_t)n * cap * sizeof(bbm_ss))) return fail(BBM_E_CUDA, "cudaMalloc site lists");
    bbm_ss* lists = (bbm_ss*)B[MB_LISTS].p;
    LAUNCH(bbm_launch_sitelist_from_search(heads, (const bbm_site*)B[MB_SITES].p, n, maxSites, lists, nss, cap, st), "sitelist_from_search_kernel launch");
    LAUNCH(bbm_launch_sitelist_bounds(lists, nss, n, cap, off, (const int*)c->mapScaf[5].p, (const int*)c->mapScaf[0].p, (const int*)c->mapScaf[1].p,
                                      cfg->sam.inter_scaffold_padding, 1, cfg->slow.expected_len_limit, out2, st), "sitelist_bounds_kernel launch");
    PairParams PP; memset(&PP, 0, sizeof PP);
    const long long npairs = n / 2;
    const int maxTasks = (int)std::min<long long>(0x3fffffffLL, 4 * n);
    if (paired) {
        if (B[MB_PSTATE].ensure((size_t)npairs * PAIR_STATE * 4) || B[MB_RFLAGS2].ensure(n * 4) || B[MB_RTASKS].ensure((size_t)maxTasks * sizeof(bbm_rescue_task)) ||
            B[MB_ROUTS].ensure((size_t)maxTasks * sizeof(bbm_rescue_out)) || B[MB_RAUX].ensure((size_t)maxTasks * sizeof(RescueAux)) || B[MB_RSITES].ensure((size_t)maxTasks * sizeof(bbm_ss)) ||
            B[MB_RTASKOF].ensure((size_t)n * cap * 4) || B[MB_PSTATS].ensure(64) || B[MB_TASKS].ensure((size_t)maxTasks * sizeof(bbm_msa_task)) ||
            B[MB_OUTS].ensure((size_t)maxTasks * sizeof(bbm_msa_out))) return fail(BBM_E_CUDA, "cudaMalloc pairing buffers");
        PP.lists = lists; PP.nss = nss; PP.npairs = npairs; PP.cap = cap; PP.read_off = off; PP.basesP = d_bases; PP.basesM = basesM; PP.quality = d_quality; PP.refs = refs;
        PP.chrom_off = chrom_off; PP.nkeys = nkeys; PP.cfg = cfg->map; PP.pc = cfg->policy; PP.tc = cfg->tip; PP.clearzone1e = cfg->slow.clearzone1e;
        PP.pstate = (int*)B[MB_PSTATE].p; PP.rflags = (int*)B[MB_RFLAGS2].p; PP.rtasks = (bbm_rescue_task*)B[MB_RTASKS].p; PP.routs = (const bbm_rescue_out*)B[MB_ROUTS].p;
        PP.raux = (RescueAux*)B[MB_RAUX].p; PP.rsites = (bbm_ss*)B[MB_RSITES].p; PP.rtask_of = (int*)B[MB_RTASKOF].p; PP.maxTasks = maxTasks;
        PP.mtasks = (bbm_msa_task*)B[MB_TASKS].p; PP.mouts = (const bbm_msa_out*)B[MB_OUTS].p; PP.counters = counters; PP.stats = (unsigned long long*)B[MB_PSTATS].p;
        CK(cudaMemsetAsync(B[MB_PSTATS].p, 0, 64, st));
        LAUNCH(bbm_launch_pair(&PP, PAIR_OP_INIT, st), "pair_kernel launch");           // pairSiteScoresInitial, paired trimList, score reset
    } else {
        LAUNCH(bbm_launch_sitelist(BBM_SL_TRIM, lists, nss, n, cap, off, d_bases, basesM, refs, chrom_off, &cfg->policy, out, st), "sitelist_kernel launch");
    }
    LAUNCH(bbm_launch_sitelist(BBM_SL_NOINDEL, lists, nss, n, cap, off, d_bases, basesM, refs, chrom_off, &cfg->policy, out, st), "sitelist_kernel launch");
    LAUNCH(bbm_launch_map_runmask(out, nss, n, paired, (int*)B[MB_RUN].p, (int*)B[MB_MASKED].p, st), "map_runmask_kernel launch");
    LAUNCH(bbm_launch_sitelist_tipdel(lists, (const int*)B[MB_MASKED].p, n, cap, off, d_bases, basesM, d_quality, refs, chrom_off, nullptr, &cfg->tip, out2, st),
           "sitelist_tipdel_kernel launch");
    S.ms_lists = T.lap(st);
    int64_t aligned = 0;
    if (int rc = scoreslow_locked(c, lists, nss, n, cap, d_off, d_bases, basesM, refs, (const int64_t*)chrom_off, (const int*)B[MB_RUN].p, &cfg->slow, (int*)B[MB_SLOWST].p,
                                  maxLen, st, &aligned, nullptr)) return rc;
    S.slow_alignments = aligned;
    S.ms_slow = T.lap(st);
    if (paired) {
        LAUNCH(bbm_launch_sitelist(BBM_SL_MERGE, lists, nss, n, cap, off, d_bases, basesM, refs, chrom_off, &cfg->policy, out, st), "sitelist_kernel launch");
        const bbm_rescue_cfg rcfg = {70, 100, 1, 100};           // POINTS_MATCH / POINTS_MATCH2, USE_AFFINE_SCORE, BASE_HIT_SCORE
        for (int dir = 0; dir < 2; dir++) {                     // rescue(r, r2, ...) then rescue(r2, r, ...): the second sees what the first added
            CK(cudaMemsetAsync(counters, 0, 32, st));
            LAUNCH(bbm_launch_pair(&PP, dir == 0 ? PAIR_OP_RESCUE_PREP0 : PAIR_OP_RESCUE_PREP1, st), "pair_kernel launch");
            int h[3];
            CK(cudaMemcpyAsync(h, counters, 12, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
            const int ntasks = std::min(h[1], maxTasks);
            if (ntasks > 0) {
                LAUNCH(bbm_launch_rescue(d_bases, refs, PP.rtasks, ntasks, &rcfg, (bbm_rescue_out*)B[MB_ROUTS].p, st), "rescue_kernel launch");
                LAUNCH(bbm_launch_rescue_mid(&PP, ntasks, st), "rescue_mid_kernel launch");
                CK(cudaMemcpyAsync(h, counters, 12, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
                S.rescue_scans += ntasks; S.rescue_fills += h[2];
                if (h[2] > 0) if (int rc = run_msa(c, d_bases, refs, PP.mtasks, (bbm_msa_out*)B[MB_OUTS].p, h[2], nullptr, nullptr, maxLen, 0, st, nullptr, nullptr)) return rc;
            }
            LAUNCH(bbm_launch_pair(&PP, dir == 0 ? PAIR_OP_RESCUE_APPLY0 : PAIR_OP_RESCUE_APPLY1, st), "pair_kernel launch");
        }
        LAUNCH(bbm_launch_pair(&PP, PAIR_OP_FINAL, st), "pair_kernel launch");
        S.ms_rescue = T.lap(st);
    } else {
        LAUNCH(bbm_launch_sitelist(BBM_SL_FINAL, lists, nss, n, cap, off, d_bases, basesM, refs, chrom_off, &cfg->policy, out, st), "sitelist_kernel launch");
    }

    // ---- genMatchString in rounds ----
    long long ms = ((std::max<long long>(2ll * maxLen + 128, cfg->map.match_slot) + 15) / 16) * 16;
    if (B[MB_GMSTATE].ensure((size_t)n * GM_STATE * 4) || B[MB_MLEN].ensure((size_t)n * GM_SLOTS * 4) || B[MB_MSLOTS].ensure((size_t)n * GM_SLOTS * ms + 64) ||
        B[MB_TASKS].ensure((size_t)n * sizeof(bbm_msa_task)) || B[MB_OUTS].ensure((size_t)n * sizeof(bbm_msa_out)) || B[MB_GTASKS].ensure((size_t)n * sizeof(bbm_gapped_task)) ||
        B[MB_GAPS].ensure((size_t)n * BBM_MAX_GAPS * 4) || B[MB_GOUTS].ensure((size_t)n * sizeof(bbm_msa_out)) || B[MB_RECS].ensure((size_t)n * sizeof(bbm_map_rec)))
        return fail(BBM_E_CUDA, "cudaMalloc genMatchString buffers");
    GmParams G; memset(&G, 0, sizeof G);
    G.lists = lists; G.nss = nss; G.nreads = n; G.cap = cap; G.read_off = off; G.basesP = d_bases; G.basesM = basesM; G.refs = refs; G.chrom_off = chrom_off;
    G.cfg = cfg->map; G.setSSScore = paired ? 0 : 1; G.rflags = nullptr;
    G.state = (int*)B[MB_GMSTATE].p; G.mslots = (int8_t*)B[MB_MSLOTS].p; G.ms = ms; G.mlen = (int*)B[MB_MLEN].p;
    G.tasks = (bbm_msa_task*)B[MB_TASKS].p; G.outs = (const bbm_msa_out*)B[MB_OUTS].p; G.gtasks = (bbm_gapped_task*)B[MB_GTASKS].p; G.gaps = (int*)B[MB_GAPS].p;
    G.gouts = (const bbm_msa_out*)B[MB_GOUTS].p; G.counters = counters;
    G.rmatch = nullptr; G.gmatch = nullptr; G.rstride = 0; G.gstride = 0;
    long long rmoffFor = -1, gmoffFor = -1;             // what the match_off arrays were generated for (count * 2^20 + stride)
    int rounds = 0;
    for (G.first = 1;; G.first = 0, rounds++) {
        if (rounds > 64) return fail(BBM_E_CUDA, "genMatchString did not converge in 64 rounds");
        CK(cudaMemsetAsync(counters, 0, 32, st));
        LAUNCH(bbm_launch_genmatch(&G, st), "genmatch_kernel launch");
        int h[5];
        CK(cudaMemcpyAsync(h, counters, 20, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        if (h[0] == 0) break;
        S.realign_fills += h[1] + h[2];
        if (h[1] > 0) {
            const long long stride = ((h[3] + 15) / 16) * 16;
            if (B[MB_RMATCH].ensure((size_t)h[1] * stride + 64) || B[MB_RMOFF].ensure((size_t)(h[1] + 1) * 8)) return fail(BBM_E_CUDA, "cudaMalloc realign match strings");
            if (rmoffFor != ((long long)h[1] << 20) + stride) { LAUNCH(bbm_launch_map_arange((long long*)B[MB_RMOFF].p, h[1], stride, st), "map_arange_kernel launch"); rmoffFor = ((long long)h[1] << 20) + stride; }
            if (int rc = run_msa(c, d_bases, refs, (const bbm_msa_task*)B[MB_TASKS].p, (bbm_msa_out*)B[MB_OUTS].p, h[1], (int8_t*)B[MB_RMATCH].p, (const int64_t*)B[MB_RMOFF].p,
                                 maxLen, 0, st, nullptr, nullptr)) return rc;
            G.rmatch = (const int8_t*)B[MB_RMATCH].p; G.rstride = stride;
        }
        if (h[2] > 0) {
            const long long stride = ((h[4] + 15) / 16) * 16;
            if (B[MB_GMATCH].ensure((size_t)h[2] * stride + 64) || B[MB_GMOFF].ensure((size_t)(h[2] + 1) * 8)) return fail(BBM_E_CUDA, "cudaMalloc realign match strings");
            if (gmoffFor != ((long long)h[2] << 20) + stride) { LAUNCH(bbm_launch_map_arange((long long*)B[MB_GMOFF].p, h[2], stride, st), "map_arange_kernel launch"); gmoffFor = ((long long)h[2] << 20) + stride; }
            if (int rc = run_msa_gapped(c, d_bases, refs, (const bbm_gapped_task*)B[MB_GTASKS].p, (const int32_t*)B[MB_GAPS].p, (bbm_msa_out*)B[MB_GOUTS].p, h[2],
                                        (int8_t*)B[MB_GMATCH].p, (const int64_t*)B[MB_GMOFF].p, st, nullptr)) return rc;
            G.gmatch = (const int8_t*)B[MB_GMATCH].p; G.gstride = stride;
        }
    }
    S.genmatch_rounds = rounds;

    // ---- the rest of processRead, Read -> SamLine ----
    bbm_map_rec* recs = d_recs ? d_recs : (bbm_map_rec*)B[MB_RECS].p;
    FinParams F; memset(&F, 0, sizeof F);
    F.lists = lists; F.nss = nss; F.nreads = n; F.cap = cap; F.read_off = off; F.basesP = d_bases; F.basesM = basesM; F.refs = refs; F.chrom_off = chrom_off;
    F.pc = cfg->policy; F.cfg = cfg->map; F.flags = out; F.state = G.state; F.mslots = G.mslots; F.ms = ms; F.mlen = G.mlen; F.recs = recs;
    if (paired) {
        LAUNCH(bbm_launch_pair_finish(&PP, &F, st), "pair_finish_kernel launch");
        unsigned long long ph[2];
        CK(cudaMemcpyAsync(ph, B[MB_PSTATS].p, 16, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        S.mated_pairs = (int64_t)ph[0]; S.inner_length_sum = (int64_t)ph[1];
    } else LAUNCH(bbm_launch_map_finish(&F, st), "map_finish_kernel launch");
    CK(cudaMemsetAsync(counters, 0, 32, st));
    LAUNCH(bbm_launch_map_status(heads, maxSites, (const int*)B[MB_SLOWST].p, nkeys, recs, n, (unsigned long long*)counters, st), "map_status_kernel launch");
    S.ms_genmatch = T.lap(st);
    if (B[MB_STASKS].ensure((size_t)n * sizeof(bbm_sam_task)) || B[MB_CIGOFF].ensure((size_t)(n + 1) * 8) || B[MB_SAM].ensure((size_t)n * sizeof(bbm_sam_out)))
        return fail(BBM_E_CUDA, "cudaMalloc SAM buffers");
    const long long cs = 2 * ms + 16;                 // toCigar: at most two bytes per match symbol + the count digits of the last run
    if (B[MB_CIGAR].ensure((size_t)n * cs + 64)) return fail(BBM_E_CUDA, "cudaMalloc CIGAR buffer");
    LAUNCH(bbm_launch_map_arange((long long*)B[MB_CIGOFF].p, n, cs, st), "map_arange_kernel launch");
    LAUNCH(bbm_launch_map_sam_tasks(recs, n, off, ms, paired, (bbm_sam_task*)B[MB_STASKS].p, st), "map_sam_tasks_kernel launch");
    bbm_sam_out* sam = d_sam ? d_sam : (bbm_sam_out*)B[MB_SAM].p;
    if (int rc = run_sam(c, (const bbm_sam_task*)B[MB_STASKS].p, n, G.mslots, (const int*)c->mapScaf[0].p, (const int*)c->mapScaf[1].p, (const int*)c->mapScaf[2].p, nchroms,
                         &cfg->sam, sam, (int8_t*)B[MB_CIGAR].p, (const int64_t*)B[MB_CIGOFF].p, st, nullptr)) return rc;
    if (d_match && match_stride > 0) LAUNCH(bbm_launch_map_copy_match(recs, G.mslots, ms, d_match, match_stride, n, st), "map_copy_match_kernel launch");
    if (cfg->sam_text) {            // SamLine.toBytes for every read, in input order
        size_t tb = 0;
        bbm_index_scan(nullptr, &tb, nullptr, nullptr, n + 1, st);
        if (B[MB_TLENS].ensure((size_t)(n + 1) * 4) || B[MB_TNM].ensure((size_t)n * 4) || B[MB_TLINEOFF].ensure((size_t)(n + 1) * 4) || B[MB_TSCAN].ensure(tb + 16) ||
            B[MB_TEXTOFF].ensure((size_t)(n + 1) * 8)) return fail(BBM_E_CUDA, "cudaMalloc SAM text offsets");
        SamTextParams X; memset(&X, 0, sizeof X);
        X.recs = recs; X.sam = sam; X.nreads = n; X.read_off = off; X.bases = d_bases; X.basesM = basesM; X.quality = d_quality;
        X.names = d_names; X.name_off = (const long long*)d_name_off;
        X.scaf_names = c->map_has_names ? (const int8_t*)c->mapScaf[3].p : nullptr; X.scaf_name_off = c->map_has_names ? (const long long*)c->mapScaf[4].p : nullptr;
        X.cigar = (const int8_t*)B[MB_CIGAR].p; X.cigar_off = (const long long*)B[MB_CIGOFF].p; X.mslots = G.mslots; X.ms = ms;
        X.paired = paired; X.intron_limit = cfg->sam.intron_limit;
        X.lens = (int*)B[MB_TLENS].p; X.nm = (int*)B[MB_TNM].p; X.line_off = (const int*)B[MB_TLINEOFF].p;
        CK(cudaMemsetAsync((int*)B[MB_TLENS].p + n, 0, 4, st));
        LAUNCH(bbm_launch_samtext_len(&X, st), "samtext_len_kernel launch");
        LAUNCH(bbm_index_scan(B[MB_TSCAN].p, &tb, (const int*)B[MB_TLENS].p, (int*)B[MB_TLINEOFF].p, n + 1, st), "SAM text scan");
        int total = 0;
        CK(cudaMemcpyAsync(&total, (const int*)B[MB_TLINEOFF].p + n, 4, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
        if (total < 0) return fail(BBM_E_CAPACITY, "bbm_map_batch: more than 2 GiB of SAM text in one batch (use smaller batches)");
        if (B[MB_TEXT].ensure((size_t)total + 64)) return fail(BBM_E_CUDA, "cudaMalloc SAM text");
        X.text = (int8_t*)B[MB_TEXT].p; X.text_off = (long long*)B[MB_TEXTOFF].p;
        LAUNCH(bbm_launch_samtext_write(&X, st), "samtext_write_kernel launch");
        S.sam_bytes = total;
    }
    unsigned long long hc[3];
    CK(cudaMemcpyAsync(hc, counters, 24, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    S.mapped = (int64_t)hc[0]; S.status_reads = (int64_t)hc[1]; S.site_overflow_reads = (int64_t)hc[2];
    S.ms_sam = T.lap(st);
    S.ms_total = Tall.lap(st);
    c->map_last_cs = cs; c->map_last_ms = ms;
    if (stats) *stats = S;
    return BBM_OK;
}

extern "C" int bbm_map_batch_dev(bbm_ctx* c, int8_t* d_bases, int8_t* d_quality, const int64_t* d_read_off, int64_t nreads, int32_t max_read_len,
                                 const bbm_mapper_cfg* cfg, bbm_map_rec* d_recs, bbm_sam_out* d_sam, int8_t* d_match, int64_t match_stride,
                                 void* stream, bbm_map_stats* stats) {
    if (int rc = map_args(c, cfg, nreads)) return rc;
    if (!d_bases || !d_read_off || max_read_len < 1) return fail(BBM_E_ARG, "bbm_map_batch_dev: null pointer");
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    long long total = 0;
    CK(cudaMemcpyAsync(&total, d_read_off + nreads, 8, cudaMemcpyDeviceToHost, st)); CK(cudaStreamSynchronize(st));
    return map_locked(c, d_bases, d_quality, d_read_off, nreads, max_read_len, cfg, d_recs, d_sam, d_match, match_stride, st, stats, total, nullptr, nullptr);
}

extern "C" int bbm_map_batch_host(bbm_ctx* c, const int8_t* bases, const int8_t* quality, const int64_t* read_off, int64_t nreads,
                                  const int8_t* names, const int64_t* name_off, const bbm_mapper_cfg* cfg, bbm_map_rec* recs, bbm_sam_out* sam,
                                  int8_t* match, int64_t match_stride, int8_t* sam_text, int64_t sam_cap, int64_t* sam_off, bbm_map_stats* stats) {
    if (int rc = map_args(c, cfg, nreads)) return rc;
    if (!bases || !read_off) return fail(BBM_E_ARG, "bbm_map_batch_host: null pointer");
    if (nreads <= 0) { if (stats) memset(stats, 0, sizeof *stats); return BBM_OK; }
    int maxLen = 1;
    for (int64_t r = 0; r < nreads; r++) {
        const int64_t l = read_off[r + 1] - read_off[r];
        if (l < 0 || l > 600) return fail(BBM_E_SHAPE, "bbm_map_batch_host: read length outside 0..600 (ALIGN_ROWS)");
        if (l > maxLen) maxLen = (int)l;
    }
    std::lock_guard<std::mutex> lk(c->mu);
    CK(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const size_t nb = (size_t)read_off[nreads];
    DevBuf* H = c->mapHost;
    if (H[0].ensure(nb + 64) || (quality && H[1].ensure(nb + 64)) || H[2].ensure((size_t)(nreads + 1) * 8) || H[3].ensure((size_t)nreads * sizeof(bbm_map_rec)) ||
        H[4].ensure((size_t)nreads * sizeof(bbm_sam_out)) || (match && match_stride > 0 && H[5].ensure((size_t)nreads * match_stride + 64)))
        return fail(BBM_E_CUDA, "cudaMalloc mapper staging");
    const bool haveNames = names && name_off;
    if (haveNames) {
        if (H[6].ensure((size_t)name_off[nreads] + 64) || H[7].ensure((size_t)(nreads + 1) * 8)) return fail(BBM_E_CUDA, "cudaMalloc mapper staging");
        CK(cudaMemcpyAsync(H[6].p, names, (size_t)name_off[nreads], cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(H[7].p, name_off, (size_t)(nreads + 1) * 8, cudaMemcpyHostToDevice, st));
    }
    CK(cudaMemcpyAsync(H[0].p, bases, nb, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync((char*)H[0].p + nb, 0, 64, st));
    if (quality) { CK(cudaMemcpyAsync(H[1].p, quality, nb, cudaMemcpyHostToDevice, st)); CK(cudaMemsetAsync((char*)H[1].p + nb, 0, 64, st)); }
    CK(cudaMemcpyAsync(H[2].p, read_off, (size_t)(nreads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (match && match_stride > 0) CK(cudaMemsetAsync(H[5].p, 0, (size_t)nreads * match_stride, st));
    bbm_map_stats S;
    int rc = map_locked(c, (int8_t*)H[0].p, quality ? (int8_t*)H[1].p : nullptr, (const int64_t*)H[2].p, nreads, maxLen, cfg, (bbm_map_rec*)H[3].p, (bbm_sam_out*)H[4].p,
                        (match && match_stride > 0) ? (int8_t*)H[5].p : nullptr, match_stride, st, &S, (int64_t)nb,
                        haveNames ? (const int8_t*)H[6].p : nullptr, haveNames ? (const int64_t*)H[7].p : nullptr);
    if (rc) return rc;
    if (cfg->sam_text) {
        if (stats) *stats = S;
        if (!sam_text || !sam_off || sam_cap < S.sam_bytes) return fail(BBM_E_CAPACITY, "bbm_map_batch_host: sam_text buffer too small (stats->sam_bytes holds the size needed)");
        CK(cudaMemcpyAsync(sam_text, c->mapBuf[MB_TEXT].p, (size_t)S.sam_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(sam_off, c->mapBuf[MB_TEXTOFF].p, (size_t)(nreads + 1) * 8, cudaMemcpyDeviceToHost, st));
    }
    if (recs) CK(cudaMemcpyAsync(recs, H[3].p, (size_t)nreads * sizeof(bbm_map_rec), cudaMemcpyDeviceToHost, st));
    if (sam) CK(cudaMemcpyAsync(sam, H[4].p, (size_t)nreads * sizeof(bbm_sam_out), cudaMemcpyDeviceToHost, st));
    if (match && match_stride > 0) CK(cudaMemcpyAsync(match, H[5].p, (size_t)nreads * match_stride, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (stats) *stats = S;
    return BBM_OK;
}
