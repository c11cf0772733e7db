/*
 * bbmap_cuda.h — C ABI of libbbmapcuda.so, the B200 (sm_100a) drop-in for BBMap's native plug-in.
 *
 * What it replaces in the reference (cavelandiah/BBMap, BBTools 36.19):
 *   - libbbtoolsjni's MultiStateAligner11ts fill kernels  jni/MultiStateAligner11tsJNI.c:100-314 (fillUnlimited),
 *     :361-704 (fillLimitedX) and their JNI shims :707-812 (header jni/align2_MultiStateAligner11tsJNI.h:165-174),
 *     loaded by current/align2/MultiStateAligner11tsJNI.java:11-42 when Shared.USE_JNI (current/align2/MSA.java:44-49);
 *   - the Java half that consumes the filled matrix: fillLimited dispatch (MultiStateAligner11tsJNI.java:116-164),
 *     score/score2 (:499-658), traceback/traceback2 (:362-495), MSA.fillAndScoreLimited (MSA.java:103-134) —
 *     moved onto the device so the `packed` matrix never has to exist in memory;
 *   - BandedAligner  jni/BandedAlignerJNI.c:123-585 and shims :588-757 (header jni/align2_BandedAlignerJNI.h:17-41).
 *
 * Conventions: plain pointers and sizes only.  Every entry point returns 0 on success or a negative BBM_E_* code;
 * nothing ever calls exit() (the reference does, jni/MultiStateAligner11tsJNI.c:130-132 — documented deviation).
 * There is NO CPU fallback: without a CUDA device every compute entry point returns BBM_E_NODEVICE.
 */
#ifndef BBMAP_CUDA_H
#define BBMAP_CUDA_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define BBM_OK            0
#define BBM_E_NODEVICE   -1   /* no CUDA device / driver */
#define BBM_E_CUDA       -2   /* a CUDA call failed; see bbm_last_error() */
#define BBM_E_ARG        -3   /* bad argument */
#define BBM_E_SHAPE      -4   /* rows/columns outside what the kernels support */
#define BBM_E_CAPACITY   -5   /* an output buffer was too small */

/* ---- task flags (bbm_msa_task.flags) ---- */
#define BBM_TF_RAW_LIMITED    1  /* exactly fillLimitedX(...minScore...)   (jni/...JNI.c:361): no dispatch rule, no -120 */
#define BBM_TF_RAW_UNLIMITED  2  /* exactly fillUnlimited(...)             (jni/...JNI.c:100) */
                                 /* neither bit: MultiStateAligner11tsJNI.fillLimited semantics (…JNI.java:132-164):
                                    limited-vs-unlimited rule, then minScore-=MIN_SCORE_ADJUST(120) */
#define BBM_TF_CLAMP          4  /* clamp the window to [0, ref_len-1] like MSA.fillAndScoreLimited (MSA.java:104-105) */
#define BBM_TF_SCORE          8  /* also run score2  (…JNI.java:537-658) */
#define BBM_TF_TRACEBACK     16  /* also run traceback2 (…JNI.java:376-495) and emit the match string */
#define BBM_TF_GAPPED        32  /* set by the library on tasks it rewrote onto a gapped reference: the fill runs on
                                    [0, greflimit] but score2 is called with refEndLoc = greflimit-1 (…JNI.java:127, 508-516) */

/* One (read, candidate window) alignment.  40 bytes. */
typedef struct {
    int64_t read_off;   /* byte offset of the read in the reads buffer */
    int64_t ref_off;    /* byte offset of the reference array (chromosome) in the reference buffer */
    int32_t read_len;   /* rows */
    int32_t ref_len;    /* length of that reference array (Java: ref.length) */
    int32_t ref_start;  /* refStartLoc, inclusive, relative to ref_off */
    int32_t ref_end;    /* refEndLoc, inclusive */
    int32_t min_score;  /* minScore (ignored by RAW_UNLIMITED) */
    int32_t flags;
} bbm_msa_task;

/* Result of one alignment.  80 bytes. */
typedef struct {
    int32_t result[5];   /* {rows, maxCol, maxState, maxScore, fail} exactly as the C writes them (unlimited: fail=0);
                            in Java-semantics mode a failed limited fill gives {rows,0,0,0,1} (Java returns null) */
    int32_t path;        /* 0 = fillLimitedX ran, 1 = fillUnlimited ran */
    int64_t iterations;  /* the reference's iterationsLimited/iterationsUnlimited increment for this call */
    int32_t score[8];    /* score2: {score,bestRefStart,bestRefStop,maxRow,maxCol,maxState,padLeft,padRight} */
    int32_t score_len;   /* 0 (not requested / fill failed), 6, or 8 (padding suggested) */
    int32_t match_len;   /* -1 (not requested / fill failed) or length of the match string */
    int32_t status;      /* 0 or BBM_E_* for this task */
    int32_t pad_;
} bbm_msa_out;

typedef struct bbm_ctx bbm_ctx;

/* Lifecycle.  `device` is a CUDA ordinal.  MSA.bandwidth / MSA.bandwidthRatio (MSA.java:864-865) are per-context. */
int  bbm_init(int device, bbm_ctx** out);
void bbm_destroy(bbm_ctx* ctx);
int  bbm_set_band(bbm_ctx* ctx, int32_t bandwidth, float bandwidthRatio);
const char* bbm_last_error(void);
int  bbm_device_count(void);

/* Reference / read residency: copy host bytes to device buffers owned by the context (returns device pointer). */
int  bbm_upload(bbm_ctx* ctx, const void* host, int64_t nbytes, void** dev_out);
int  bbm_free_dev(bbm_ctx* ctx, void* dev);

/* Batched MultiStateAligner11ts — everything already resident in device memory (pointers are device pointers).
 * match_off has ntasks+1 entries; task i's match string is written at match_buf+match_off[i] (capacity
 * match_off[i+1]-match_off[i], rows+columns is always enough without '-' symbols).  `stream` is a cudaStream_t (or 0).
 * kernel_ms_out (optional, host) receives the device time of the launches measured with CUDA events. */
int  bbm_msa_batch_dev(bbm_ctx* ctx, const int8_t* d_reads, const int8_t* d_refs, const bbm_msa_task* d_tasks,
                       bbm_msa_out* d_outs, int64_t ntasks, int8_t* d_match_buf, const int64_t* d_match_off,
                       int32_t max_rows, int32_t max_cols, void* stream, float* kernel_ms_out);

/* Same, from HOST buffers: copies tasks/reads in, runs, copies outs/match strings back (the reference-facing call;
 * the reference arrays `d_refs` stay resident, uploaded once with bbm_upload like the reference keeps chromosomes
 * in memory, dna/Data.java).  reads_bytes = size of the reads buffer. */
int  bbm_msa_batch_host(bbm_ctx* ctx, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs,
                        const bbm_msa_task* tasks, bbm_msa_out* outs, int64_t ntasks,
                        int8_t* match_buf, const int64_t* match_off);

/* ---- gapped references: MSA.fillAndScoreLimited(read, ref, start-thresh, stop+thresh, minScore, gaps) (MSA.java:103-134) ->
 * makeGref (…JNI.java:668-757) + translateFromGappedCoordinate (:759-779).  A task with ngaps>0 is aligned against the
 * gapped reference built on the device from its gap array {start0,stop0,start1,stop1,...} (SiteScore.gaps); bestRefStart /
 * bestRefStop (score[1], score[2]) come back in chromosome coordinates.  Tasks with ngaps==0 behave as in bbm_msa_batch_*.
 * A gapped reference longer than ALIGN_COLUMNS+2 = 3002 bytes gives status BBM_E_SHAPE (the reference asserts).
 * match_off capacity per gapped task: rows + columns + 127 * ('-' symbols). */
typedef struct {                /* 48 bytes */
    bbm_msa_task t;             /* ref_start/ref_end = the window before clamping; flags: BBM_TF_SCORE / BBM_TF_TRACEBACK */
    int32_t gaps_off;           /* index of the first int of this task's gap array in the gaps buffer */
    int32_t ngaps;              /* ints in the gap array (even); 0 = no gaps */
} bbm_gapped_task;
typedef struct { int32_t origin, greflimit, greflimit2, status; } bbm_gref_info;   /* grefRefOrigin, greflimit, greflimit2 */
int  bbm_msa_gapped_batch_dev(bbm_ctx* ctx, const int8_t* d_reads, const int8_t* d_refs, const bbm_gapped_task* d_tasks,
                              const int32_t* d_gaps, bbm_msa_out* d_outs, int64_t ntasks, int8_t* d_match_buf,
                              const int64_t* d_match_off, void* stream, float* kernel_ms_out);
int  bbm_msa_gapped_batch_host(bbm_ctx* ctx, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs,
                               const bbm_gapped_task* tasks, const int32_t* gaps, int64_t ngap_ints, bbm_msa_out* outs,
                               int64_t ntasks, int8_t* match_buf, const int64_t* match_off);

/* Tuning / introspection.  bbm_set_option keys:
 *   "narrow"  0 = off, 1 = every shape-eligible limited fill first tries the 16-diagonal thread-per-alignment kernel, n>1 = only those
 *             whose minScore lies within n points of the best possible score (default 1000; the others go straight to the strip kernel);
 *   "strip"   0 = limited un-banded fills use the warp-per-alignment tiled kernel, n>0 = work-estimate buckets (of 4096 cells) below n use
 *             the thread-per-alignment strip kernel (default 16 = all);  "strip_budget_mb" = device scratch the strip kernel may use;
 *   "search_shared" 1 = the index-search kernel keeps its per-read walk arrays in shared memory when a batch has <= 32 keys per read
 *             (A/B; measured slower than the per-thread global pool on B200, so off by default);
 *   "search_split" 0 = BBIndex.find in one thread-per-read launch, 1 = key filtering / prescan / walk as three thread-per-read launches,
 *             2 (default) = the prescan with one warp per read (reads with more than 32 keys fall back to the thread-per-read prescan);
 *   "search_profile", "strip_debug" (diagnostics).
 * Results are bit-identical for every setting.  bbm_get_stat keys: "launches", "band_misses" (banded alignments re-run by the
 * row-sequential kernel), "tasks_total", "narrow_tried", "narrow_handed_over", "strip_tasks", "index_build_us". */
int     bbm_set_option(bbm_ctx* ctx, const char* key, int value);
int64_t bbm_get_stat(const bbm_ctx* ctx, const char* key);

/* Measures the integer / DPX issue peak of this GPU (roofline denominator for the DP kernels): giga lane-ops per second of
 * instruction kind 0 IADD3, 1 LOP3, 2 VIMNMX3 (DPX), 3 VIADDMNMX (DPX), 4 IMAD, 5 half IMAD + half LOP3, 6 compare+select. */
int  bbm_int_peak(bbm_ctx* ctx, int kind, double* gops_out);

/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
int64_t bbm_launch_count(const bbm_ctx* ctx);

/* ---- read ingest: Read.validate (current/stream/Read.java:81-215; switches :3406-3418) applied in place to a read batch, plus
 * the minus-strand copy AminoAcid.reverseComplementBases (current/dna/AminoAcid.java:203-211) the mapper makes once per read
 * (AbstractMapThread.java:492-503).  quality holds phred values with the ASCII offset removed, or NULL (FASTA input).
 * flags mirror the reference's switches (all off = BBMap defaults).  basesM / read_flags may be NULL.  Device buffers must be
 * 16-byte aligned and readable up to 16 bytes past read_off[nreads] (the kernel moves whole 16-byte vectors). */
#define BBM_ING_FIX_JUNK         1   /* Read.FIX_JUNK (fixjunk): non-IUPAC bytes become 'N' instead of flagging the read */
#define BBM_ING_U_TO_T           2   /* Read.U_TO_T */
#define BBM_ING_TO_UPPER_CASE    4   /* Read.TO_UPPER_CASE (touppercase) */
#define BBM_ING_LOWER_CASE_TO_N  8   /* Read.LOWER_CASE_TO_N (lowercaseton) */
#define BBM_READ_JUNK            1   /* read_flags bit: Read.junk() */
int  bbm_ingest_batch_dev(bbm_ctx* ctx, int8_t* d_bases, int8_t* d_quality, const int64_t* d_read_off, int64_t nreads, int32_t max_len,
                          int32_t flags, int8_t* d_basesM, int32_t* d_read_flags, void* stream, float* kernel_ms_out);
int  bbm_ingest_batch_host(bbm_ctx* ctx, int8_t* bases, int8_t* quality, const int64_t* read_off, int64_t nreads, int32_t flags,
                           int8_t* basesM, int32_t* read_flags);

/* ---- k-mer index: IndexMaker4 (build) + BBIndex.analyzeIndex (current/align2/IndexMaker4.java:160-421, BBIndex.java:101-191) ---- */
typedef struct {        /* 80 bytes: the BBIndex statics as they stand after BBMap.loadIndex + analyzeIndex for this genome */
    int32_t keylen, chrombits, shift_length, chroms_per_block;
    int32_t max_hits_reduction2, maximum_max_hits_reduction, hit_reduction_div, points_per_site;
    int32_t min_index_to_drop_long_hit_list, max_average_list_to_search, max_average_list_to_search2, max_single_list_to_search;
    int32_t max_shortest_list_to_search, max_usable_length, max_usable_length2, pad_;
    float fraction_to_exclude, padf_[3];
} bbm_index_cfg;
/* d_chroms: the chromosome arrays (upper-case ACGTN bytes, N-padded exactly as FastaToChromArrays2 lays them out),
 * concatenated, device-resident; chrom_off (HOST) has nchroms+1 byte offsets; chromosome numbers are 1-based.
 * chrombits<0 = automatic (BBMap.java:317-321).  The index stays resident in the context. */
int  bbm_index_build(bbm_ctx* ctx, const int8_t* d_chroms, const int64_t* chrom_off, int32_t nchroms, int32_t keylen, int32_t chrombits,
                     bbm_index_cfg* cfg_out, int32_t* nblocks_out);
/* Lets `dst` (same device) search and map against the index and reference resident in `src`, without copying them: one index per process, one
 * context (scratch buffers + stream) per batch in flight, the way the reference keeps one index and one MSA per mapping thread
 * (AbstractMapThread.java:133-136).  `src` must outlive `dst`. */
int  bbm_index_share(bbm_ctx* dst, bbm_ctx* src);
/* Sizes / contents for inspection and parity tests: nsites of a block; copies of starts[4^k+1], sites[nsites] of one block;
 * COUNTS[4^k] and lengthHistogram[1001] (either may be NULL). */
int  bbm_index_block_sites(bbm_ctx* ctx, int32_t block, int64_t* nsites_out);
int  bbm_index_download(bbm_ctx* ctx, int32_t block, int32_t* starts, int32_t* sites, int32_t* counts, int32_t* hist1001);

/* ---- index persistence: the reference's own on-disk formats, host side, no device needed for the bbm_wire_* calls ----
 * Files are Java ObjectOutputStream streams (current/fileIO/ReadWrite.java:200-240, :739-757); a name ending in .gz is gzipped
 * (ReadWrite.getOutputStream :349-392).  Arrays returned by the readers are malloc'ed: release with bbm_wire_free.
 * All return 0 or a negative BBM_E_* code with the reason in bbm_wire_last_error() (thread-local). */
typedef struct {        /* ref/genome/<build>/summary.txt (dna/FastaToChromArrays2.java:229-250) */
    int64_t chroms, bases, defined, undefined, contigs, scaffolds, interpad;
    int32_t version, pad_;
    char name[256];
} bbm_genome_summary;
const char* bbm_wire_last_error(void);
void bbm_wire_free(void* p);
/* ReadWrite.write(int[] x, fname) / ReadWrite.read(int[].class, fname) */
int  bbm_wire_write_int_array(const char* path, const int32_t* data, int64_t n);
int  bbm_wire_read_int_array(const char* path, int32_t** data_out, int64_t* n_out);
/* IndexMaker4.fname (current/align2/IndexMaker4.java:477-488): <root_index><build>/chr<a>[-<b>]_index_k<k>_c<chrombits>_b<build>.block */
int  bbm_wire_block_fname(char* out, size_t cap, const char* root_index, int minChrom, int maxChrom, int k, int chrombits, int build);
/* Block.write / Block.read (current/align2/Block.java:74-160): `fname` holds int[] sites, `fname`+"2.gz" the delta-coded int[] starts
 * (nstarts = 4^k + 1 entries). */
int  bbm_wire_write_block(const char* fname, const int32_t* sites, int64_t nsites, const int32_t* starts, int64_t nstarts);
int  bbm_wire_read_block(const char* fname, int32_t** sites, int64_t* nsites, int32_t** starts, int64_t* nstarts);
/* dna.ChromosomeArray as chrN.chrom.gz (current/dna/ChromosomeArray.java:14-22,63-71,415-419; written FastaToChromArrays2.java:347-353) */
int  bbm_wire_write_chrom(const char* path, int32_t chromosome, const int8_t* array, int32_t len, int32_t minIndex, int32_t maxIndex, int8_t strand);
int  bbm_wire_read_chrom(const char* path, int32_t* chromosome, int8_t** array, int32_t* len, int32_t* minIndex, int32_t* maxIndex, int8_t* strand);
int  bbm_wire_write_summary(const char* path, const bbm_genome_summary* g);
int  bbm_wire_read_summary(const char* path, bbm_genome_summary* g);
/* Writes every block of the resident index the way IndexMaker4.makeIndex does (IndexMaker4.java:197-200) / loads them back instead of
 * building (:135-139; the analysis — COUNTS, lengthHistogram, derived limits — is recomputed on the device exactly as after a build). */
int  bbm_index_save(bbm_ctx* ctx, const char* root_index, int32_t build);
int  bbm_index_load(bbm_ctx* ctx, const int8_t* d_chroms, const int64_t* chrom_off, int32_t nchroms, int32_t keylen, int32_t chrombits,
                    const char* root_index, int32_t build, bbm_index_cfg* cfg_out, int32_t* nblocks_out);

/* ---- read batching in front of the mapper: ReformatReads.breakReads(list, max, min) as AbstractMapThread.run applies it to every list of reads when
 * `maxlen` / `minlen` are set (current/jgi/ReformatReads.java:1179-1219, current/align2/AbstractMapThread.java:441-443; how configs[4]'s 1-kbp reads reach a
 * 600-row aligner).  A read shorter than `min_len` is dropped; a read longer than `max_len` (> 0) is cut into pieces [0,max) [max,2max) ... while the piece
 * start is < length - min_len, piece number n (1-based) named "<name>_<n>"; other reads pass through unchanged.  Host code, no device involved.
 * Call once with out_* NULL to get the sizes (n_out, bases_out_len, names_out_len), then with buffers of at least those sizes.  src[i] = index of the input
 * read piece i came from, piece_start[i] = its offset in that read.  Paired input with a read longer than max_len is BBM_E_ARG (the reference asserts). ---- */
int bbm_break_reads(const int8_t* bases, const int8_t* quality /* may be NULL */, const int64_t* read_off, int64_t nreads,
                    const int8_t* names, const int64_t* name_off, int32_t paired, int32_t max_len, int32_t min_len,
                    int64_t* n_out, int64_t* bases_out_len, int64_t* names_out_len,
                    int8_t* out_bases, int8_t* out_quality /* may be NULL */, int64_t* out_read_off /* n_out+1 */,
                    int8_t* out_names, int64_t* out_name_off /* n_out+1 */, int64_t* src /* n_out */, int32_t* piece_start /* n_out */);

/* ---- index search: BBIndex.find (current/align2/BBIndex.java:403-639) — seeds -> candidate sites (SiteScore) ---- */
#define BBM_MAX_GAPS 10
#define BBM_ST_ANOMALY        1   /* extendScore located no base (the reference prints an anomaly and scores -99999) */
#define BBM_ST_SITE_OVERFLOW  2   /* more sites than max_sites for one read (extra sites dropped) */
#define BBM_ST_GAP_OVERFLOW   4   /* gap array longer than 9 ints */
#define BBM_ST_GAPFIX         8   /* subsumption into a site that carries gaps: GapTools.fixGaps is not implemented */
#define BBM_ST_BADARG        16
typedef struct {            /* 64 bytes: one SiteScore as BBIndex emits it (stream/SiteScore.java) */
    int32_t chrom, start, stop, hits, score, ngaps;
    int8_t strand, perfect, semiperfect, pad_;
    int32_t gaps[BBM_MAX_GAPS - 1];
} bbm_site;
typedef struct {            /* 48 bytes per read */
    int32_t nsites, status, num_hits, max_score, max_quick_score, pad_;
    int32_t best_scores[6];     /* bestScores[] at the end of find(): top score, max hits, qcutoff, best qscore, maxQuickScore, perfects */
} bbm_search_head;
/* Uses the index resident in the context (bbm_index_build).  Inputs are the seeding outputs (bbm_seed_batch_*): per read
 * nkeys, offsets[maxKeys], keyScores[maxKeys] and the base scores; sites is [nreads][max_sites].  quit_after_two_perfects
 * mirrors AbstractIndex.QUIT_AFTER_TWO_PERFECTS (true single-ended, false paired: BBMap.java:434). */
int  bbm_search_batch_dev(bbm_ctx* ctx, const int8_t* d_bases, const int8_t* d_baseScores, const int64_t* d_read_off, int64_t nreads,
                          const int32_t* d_nkeys, const int32_t* d_offsets, const int32_t* d_keyScores, int32_t maxKeys,
                          int32_t quit_after_two_perfects, bbm_search_head* d_heads, bbm_site* d_sites, int32_t max_sites,
                          int32_t max_read_len /* longest read of the batch, 0 = unknown */, void* stream, float* kernel_ms_out);
int  bbm_search_batch_host(bbm_ctx* ctx, const int8_t* bases, const int8_t* baseScores, const int64_t* read_off, int64_t nreads,
                           const int32_t* nkeys, const int32_t* offsets, const int32_t* keyScores, int32_t maxKeys,
                           int32_t quit_after_two_perfects, bbm_search_head* heads, bbm_site* sites, int32_t max_sites);

/* ---- ungapped site scoring: MSA.scoreNoIndels / scoreNoIndelsAndMakeMatchString (…JNI.java:1033-1089, 1243-1318) ---- */
typedef struct {                /* 32 bytes */
    int64_t read_off, ref_off;  /* byte offsets of the read / of the reference array */
    int32_t read_len, ref_len;  /* ref_len = length of the reference array (Java ref.length) */
    int32_t ref_start;          /* refStart (may be negative / run past the end: scored as POINTS_NOREF) */
    int32_t flags;              /* bit0: also write the match string ('m','S','N'); then out-of-bounds sites score -99999 */
} bbm_noindel_task;
int  bbm_noindel_batch_dev(bbm_ctx* ctx, const int8_t* d_reads, const int8_t* d_refs, const bbm_noindel_task* d_tasks, int32_t* d_scores,
                           int8_t* d_match_buf, const int64_t* d_match_off, int64_t ntasks, void* stream, float* kernel_ms_out);
int  bbm_noindel_batch_host(bbm_ctx* ctx, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs, const bbm_noindel_task* tasks,
                            int32_t* scores, int8_t* match_buf, const int64_t* match_off, int64_t ntasks);

/* ---- KeyRing seeding (AbstractMapThread.quickMap up to the index search; QualityTools / KeyRing) ---- */
typedef struct {                /* 32 bytes; defaults = BBMap.setDefaults (current/align2/BBMap.java:45-65) */
    int32_t keylen;             /* 13 */
    int32_t maxDesiredKeys;     /* 15 */
    int32_t baseKeyHitScore;    /* BASE_HIT_SCORE*keylen = 1300 (AbstractIndex.java:17) */
    int32_t minApproxHitsToKeep;/* 1 */
    float keyDensity, maxKeyDensity, minKeyDensity, pad_;   /* 1.9, 3.0, 1.5 */
} bbm_seed_cfg;
/* reads: concatenated bases (ASCII) and, optionally, qualities (phred, offset removed; NULL = FASTA path); read_off has
 * nreads+1 entries.  Per read r: nkeys[r] = number of seeds (>=1), 0 (shorter than k) or -1 (discarded by quickMap's rules);
 * offsets/keys/keyScores are [nreads][maxKeys] (unused slots -1/-1/0); baseScores is parallel to bases.  With
 * offsetsM/keysM non-NULL also emits the minus-strand lists (KeyRing.reverseOffsets / reverseComplementKeys).
 * Everything device-resident for *_dev; *_host copies in and out. */
int  bbm_seed_batch_dev(bbm_ctx* ctx, const int8_t* d_bases, const int8_t* d_quality, const int64_t* d_read_off, int64_t nreads,
                        int32_t max_read_len, const bbm_seed_cfg* cfg, int32_t maxKeys, int32_t* d_nkeys, int32_t* d_offsets,
                        int32_t* d_keys, int32_t* d_keyScores, int8_t* d_baseScores, int32_t* d_offsetsM, int32_t* d_keysM,
                        void* stream, float* kernel_ms_out);
int  bbm_seed_batch_host(bbm_ctx* ctx, const int8_t* bases, const int8_t* quality, const int64_t* read_off, int64_t nreads,
                         const bbm_seed_cfg* cfg, int32_t maxKeys, int32_t* nkeys, int32_t* offsets, int32_t* keys,
                         int32_t* keyScores, int8_t* baseScores, int32_t* offsetsM, int32_t* keysM);

/* ---- BandedAligner (jni/BandedAlignerJNI.c) ---- */
#define BBM_DIR_FORWARD     0   /* alignForward    jni/BandedAlignerJNI.c:123-239 */
#define BBM_DIR_FORWARD_RC  1   /* alignForwardRC  :241-355 */
#define BBM_DIR_REVERSE     2   /* alignReverse    :357-470 */
#define BBM_DIR_REVERSE_RC  3   /* alignReverseRC  :472-585 */
typedef struct {                /* 48 bytes; argument meaning as in the C signatures */
    int64_t query_off, ref_off; /* byte offsets into the query / reference buffers */
    int32_t query_len, ref_len, qstart, rstart, max_edits, max_width, exact, dir;
} bbm_band_task;
typedef struct {                /* 32 bytes */
    int32_t edits;              /* return value */
    int32_t rv[5];              /* {lastQueryLoc,lastRefLoc,lastRow,lastEdits,lastOffset} (BandedAlignerJNI.java returnVals) */
    int32_t status, pad_;
} bbm_band_out;
/* Batched, device-resident buffers / host buffers.  Band widths up to 127 cells (maxEdits<=63) are supported. */
int  bbm_banded_batch_dev(bbm_ctx* ctx, const int8_t* d_queries, const int8_t* d_refs, const bbm_band_task* d_tasks,
                          bbm_band_out* d_outs, int64_t ntasks, void* stream, float* kernel_ms_out);
int  bbm_banded_batch_host(bbm_ctx* ctx, const int8_t* queries, int64_t query_bytes, const int8_t* refs, int64_t ref_bytes,
                           const bbm_band_task* tasks, bbm_band_out* outs, int64_t ntasks);

/* ---- SAM record fields of mapped reads: SamLine(Read,int) (current/stream/SamLine.java:82-330) with toCigar13/toCigar14 (:600-750),
 * makeFlag (:2134-2151), toMapq (:1709-1723) and the scaffold lookups of dna/Data.java:1089-1140.  One record per read as the mapper
 * leaves it; mates point at each other through `mate`.  Match strings are long-format (m S N I D X Y C B s), in reference order. ---- */
#define BBM_RF_MAPPED     1
#define BBM_RF_MINUS      2    /* r.strand()==Gene.MINUS */
#define BBM_RF_PERFECT    4
#define BBM_RF_AMBIGUOUS  8
#define BBM_RF_SECONDARY  16
#define BBM_RF_DISCARDED  32
#define BBM_RF_PAIRED     64   /* r.paired(): the pairing logic accepted the pair */
#define BBM_RF_PAIRNUM1   128  /* r.pairnum()==1 (second fragment) */
typedef struct {                /* 48 bytes */
    int64_t match_off;          /* byte offset of the match string in the match buffer */
    int32_t match_len;          /* 0 = r.match==null */
    int32_t chrom, start, stop; /* chromosome-array coordinates, chrom 1-based (r.chrom, r.start, r.stop) */
    int32_t read_len;           /* r.length() */
    int32_t score;              /* r.mapScore */
    int32_t mate;               /* index of the mate's record in this batch, -1 = r.mate==null */
    int32_t flags;              /* BBM_RF_* */
    int32_t pad_;
} bbm_sam_task;
typedef struct {                /* 32 bytes */
    int32_t flag, pos, mapq;    /* FLAG, POS (1-based, scaffold-relative), MAPQ */
    int32_t scaffold;           /* RNAME as an index into the scaffold table, -1 = '*' */
    int32_t rnext;              /* -1 = '*', -2 = '=', else scaffold index */
    int32_t pnext, tlen;
    int32_t cigar_len;          /* -1 = '*' (null cigar), -2 = invalid match character (the reference throws) */
} bbm_sam_out;
typedef struct {                /* 32 bytes; defaults: SamLine.VERSION=1.4, SOFT_CLIP=true, INTRON_LIMIT=Integer.MAX_VALUE, PENALIZE_AMBIG=true (:2424-2434);
                                   inter_scaffold_padding = Data.interScaffoldPadding = FastaToChromArrays2.MID_PADDING = 300 */
    int32_t version14, soft_clip, intron_limit, penalize_ambig, inter_scaffold_padding, pad_[3];
} bbm_sam_cfg;
/* Scaffold table (dna/Data.scaffoldLocs / scaffoldLengths): scaffolds of chromosome c (1-based) are entries scaf_off[c-1] .. scaf_off[c]-1 of
 * scaf_loc (start inside the chromosome array, ascending) and scaf_len.  cigar i is written at cigar_buf + cigar_off[i] (capacity
 * 2*match_len+12 is always enough). */
int  bbm_sam_batch_dev(bbm_ctx* ctx, const bbm_sam_task* d_tasks, int64_t n, const int8_t* d_match_buf, const int32_t* d_scaf_off,
                       const int32_t* d_scaf_loc, const int32_t* d_scaf_len, int32_t nchroms, const bbm_sam_cfg* cfg, bbm_sam_out* d_outs,
                       int8_t* d_cigar_buf, const int64_t* d_cigar_off, void* stream, float* kernel_ms_out);
int  bbm_sam_batch_host(bbm_ctx* ctx, const bbm_sam_task* tasks, int64_t n, const int8_t* match_buf, int64_t match_bytes, const int32_t* scaf_off,
                        const int32_t* scaf_loc, const int32_t* scaf_len, int32_t nchroms, const bbm_sam_cfg* cfg, bbm_sam_out* outs,
                        int8_t* cigar_buf, const int64_t* cigar_off);

/* ---- ungapped scans around candidate sites: tip-deletion search and mate rescue ----
 * bbm_tipdel_*: AbstractMapThread.findTipDeletions(SiteScore ss, bases, maxImperfectScore, lookRight, lookLeft)
 *   (current/align2/AbstractMapThread.java:1107-1141) with findTipDeletionsRight/Left (:2178-2294): per site, how far stop may move
 *   right / start may move left so that slow alignment sees a deletion near the read tip.  The caller then rescoring the site with
 *   scoreNoIndels (bbm_noindel_*) is AbstractMapThread.findTipDeletions(Read,...) :1073-1104.
 * bbm_rescue_*: AbstractMapThread.quickRescue (:2303-2405): best ungapped placement of the loose mate within searchDist of loc,
 *   plus SiteScore.setPerfect / isInBounds of the site it returns (stream/SiteScore.java:239-291, 425-428). */
typedef struct {                /* 48 bytes */
    int64_t read_off, ref_off;  /* bases on the site's strand; chromosome array (ChromosomeArray.array) */
    int32_t read_len, ref_len;  /* ref_len = array.length */
    int32_t min_index;          /* ChromosomeArray.minIndex */
    int32_t start, stop;        /* SiteScore.start / stop */
    int32_t slow_score, max_imperfect;   /* nothing is searched when slowScore >= maxImperfectScore */
    int32_t flags;              /* bit0 lookRight, bit1 lookLeft */
} bbm_tipdel_task;
typedef struct { int32_t start, stop, right, left; } bbm_tipdel_out;      /* new start/stop; right = x, left = y (0 = unchanged) */
typedef struct {                /* defaults: TIP_SEARCH_DIST=100 (BBMap.java:59), TIP_DELETION_MAX_TIPLEN=8 (AbstractMapThread.java:2989),
                                   ALIGN_COLUMNS=3000, SLOW_RESCUE_PADDING=8 (BBMap.java:57-58) */
    int32_t search_range, max_tiplen, align_columns, slow_rescue_padding;
} bbm_tipdel_cfg;
typedef struct {                /* 56 bytes */
    int64_t read_off, ref_off;
    int32_t read_len, ref_len, min_index, max_index;   /* max_index = ChromosomeArray.maxIndex (isInBounds) */
    int32_t loc, search_dist, ideal_start, max_mismatches;
    int32_t flags, pad_;        /* bit0 searchRight */
} bbm_rescue_task;
typedef struct {                /* 32 bytes; start = -1: quickRescue returned null */
    int32_t start, stop, mismatches /* SiteScore.slowScore as left by quickRescue */, max_contig, score;
    int32_t perfect;            /* bit0 perfect, bit1 semiperfect */
    int32_t in_bounds, pad_;
} bbm_rescue_out;
typedef struct { int32_t points_match /*70*/, points_match2 /*100*/, use_affine /*1*/, base_hit_score /*100*/; } bbm_rescue_cfg;
int  bbm_tipdel_batch_dev(bbm_ctx* ctx, const int8_t* d_reads, const int8_t* d_refs, const bbm_tipdel_task* d_tasks, int64_t n,
                          const bbm_tipdel_cfg* cfg, bbm_tipdel_out* d_outs, void* stream, float* kernel_ms_out);
int  bbm_tipdel_batch_host(bbm_ctx* ctx, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs, const bbm_tipdel_task* tasks, int64_t n,
                           const bbm_tipdel_cfg* cfg, bbm_tipdel_out* outs);
int  bbm_rescue_batch_dev(bbm_ctx* ctx, const int8_t* d_reads, const int8_t* d_refs, const bbm_rescue_task* d_tasks, int64_t n,
                          const bbm_rescue_cfg* cfg, bbm_rescue_out* d_outs, void* stream, float* kernel_ms_out);
int  bbm_rescue_batch_host(bbm_ctx* ctx, const int8_t* reads, int64_t reads_bytes, const int8_t* d_refs, const bbm_rescue_task* tasks, int64_t n,
                           const bbm_rescue_cfg* cfg, bbm_rescue_out* outs);

/* ---- per-read site-list policies of the unpaired mapping loop (part of SURVEY 8f.1) ----
 * The list handling BBMapThread.processRead does around the alignment stages (current/align2/BBMapThread.java:420-553), on the device
 * so that search -> scoreNoIndels -> slow alignment can be chained without a host round trip:
 *   BBM_SL_TRIM     Collections.sort + trimList (:428-431; trimList :140-249; Tools.trimSiteList / trimSitesBelowCutoff, Tools.java:654-673,1106-1160)
 *   BBM_SL_NOINDEL  AbstractMapThread.scoreNoIndels(Read,...) (AbstractMapThread.java:762-855) + Collections.sort (:442)
 *   BBM_SL_FINAL    mergeDuplicateSites + sort, Read.setPerfectFlag, clearzone / ambiguity, removeLowQualitySitesUnpaired (:478-553)
 * One list of `cap` bbm_ss slots per read; nss[r] live entries. */
typedef struct {                /* 80 bytes: the SiteScore fields the policies read and write (stream/SiteScore.java) */
    int32_t chrom, start, stop, hits, score, quick_score, slow_score, paired_score;
    int8_t strand, perfect, semiperfect, rescued;
    int32_t ngaps;              /* 0 = gaps == null */
    int32_t gaps[BBM_MAX_GAPS - 1];
    int32_t has_match;          /* set where the reference attaches genMatchNoIndels (text: bbm_noindel_* with flag bit0) */
} bbm_ss;
typedef struct {                /* 80 bytes; defaults in bbmap_b200/sitelist.py (BBMapThread.java:38-62,114-118; AbstractMapThread.java:142,3004) */
    int32_t trim_list, min_trim_sites_to_retain, max_trim_sites_to_retain, quick_match_strings;
    int32_t clearzone1, clearzone1b, clearzone1c, clearzonep, clearzone3, clearzone1e, clearzone_limit1e, print_secondary;
    float min_align_ratio, cz1b_scale, cz1b_flat, cz1c_scale;
    float cz1c_flat; int32_t pad_[3];
} bbm_policy_cfg;
typedef struct { int32_t near_perfect /* scoreNoIndels' return value */, flags /* bit0 mapped, bit1 perfect, bit2 ambiguous */, clearzone, best_sites; } bbm_read_out;
#define BBM_SL_TRIM 1
#define BBM_SL_NOINDEL 2
#define BBM_SL_FINAL 3
#define BBM_SL_MERGE 4                /* Tools.mergeDuplicateSites(list, true, true) only (processReadPair, BBMapThread.java:1043,1059) */
/* sites as BBIndex.find emits them -> SiteScore(chrom, strand, start, stop, hits, quickScore): score = quickScore (SiteScore.java:40-52) */
int  bbm_sitelist_from_search_dev(bbm_ctx* ctx, const bbm_search_head* d_heads, const bbm_site* d_sites, int64_t nreads, int32_t max_sites,
                                  bbm_ss* d_lists, int32_t* d_nss, int32_t cap, void* stream);
/* d_read_off has nreads+1 entries (read r is bases[read_off[r] .. read_off[r+1])); d_basesP/d_basesM/d_refs/d_chrom_off are only read by
 * BBM_SL_NOINDEL (chrom_off as in bbm_index_build: chromosome c occupies refs[chrom_off[c-1] .. chrom_off[c])). */
int  bbm_sitelist_batch_dev(bbm_ctx* ctx, int32_t op, bbm_ss* d_lists, int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                            const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_refs, const int64_t* d_chrom_off,
                            const bbm_policy_cfg* cfg, bbm_read_out* d_out, void* stream, float* kernel_ms_out);
int  bbm_sitelist_batch_host(bbm_ctx* ctx, int32_t op, bbm_ss* lists, int32_t* nss, int64_t nreads, int32_t cap, const int64_t* read_off,
                             const int8_t* basesP, const int8_t* basesM, const int8_t* d_refs, const int64_t* chrom_off, int32_t nchroms,
                             const bbm_policy_cfg* cfg, bbm_read_out* out);

/* AbstractMapThread.removeOutOfBounds (current/align2/AbstractMapThread.java:2444-2479), quickMap's step right after the index search: sites hanging over
 * the chromosome array (start < 0 or stop > ChromosomeArray.maxIndex) are removed; with SAM output (sam_out != 0) so are sites that span two scaffolds
 * (Data.isSingleScaffold; d_scaf_off/d_scaf_loc as in bbm_sam_batch_*, NULL = no scaffold table); over-long ungapped sites are cut to read length + 40.
 * d_out[r].best_sites = sites removed; flags bit3 = an over-long gapped site was left alone (needs GapTools.fixGaps). */
int  bbm_sitelist_bounds_dev(bbm_ctx* ctx, bbm_ss* d_lists, int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                             const int32_t* d_chrom_max_index, const int32_t* d_scaf_off, const int32_t* d_scaf_loc, int32_t inter_scaffold_padding,
                             int32_t sam_out, int32_t expected_len_limit, bbm_read_out* d_out, void* stream);

/* AbstractMapThread.findTipDeletions(Read r, basesP, basesM, maxSwScore, maxImperfectScore) (current/align2/AbstractMapThread.java:1073-1104) on
 * every read's list: quality gate (d_quality NULL = FASTA), findTipDeletions per eligible site, and for changed sites the rescoring with
 * scoreNoIndels and the perfect/semiperfect update.  d_out[r].best_sites = sites changed; flags bit3 = a gapped site was skipped
 * (setStart/setStop on gapped sites need GapTools.fixGaps).  d_chrom_min_index may be NULL (minIndex 0 everywhere). */
int  bbm_sitelist_tipdel_dev(bbm_ctx* ctx, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                             const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_quality, const int8_t* d_refs, const int64_t* d_chrom_off,
                             const int32_t* d_chrom_min_index, const bbm_tipdel_cfg* cfg, bbm_read_out* d_out, void* stream, float* kernel_ms_out);

/* removeDuplicateBestSites (current/align2/AbstractMapThread.java:1328-1349; BBMapThread.java:624-628), then the clearzone-3 block and the final score gate of
 * processRead, after the primary site's match string exists (BBMapThread.java:667-684, 698-700; AbstractMapThread.applyClearzone3 :1820-1870, calcCZ3_fraction :1893-1911): the near-ties behind the top site lower every score of the list,
 * a read pushed under MINIMUM_ALIGNMENT_SCORE_RATIO becomes ambiguous, AMBIGUOUS_TOSS (ambiguous_toss != 0) and the ratio gate clear the mapping.
 * d_io is in/out: flags as BBM_SL_FINAL left them; afterwards flags updated, near_perfect = r.mapScore (= the top site's slowScore, Read.java:1178),
 * best_sites = the amount subtracted (0: applyClearzone3 returned false). */
int  bbm_sitelist_clearzone3_dev(bbm_ctx* ctx, bbm_ss* d_lists, int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                 const bbm_policy_cfg* cfg, int32_t ambiguous_toss, bbm_read_out* d_io, void* stream);
/* PENALIZE_AMBIG block of processRead (BBMapThread.java:706-709): AbstractMapThread.calcTipScorePenalty(r, maxSwScore, tiplen) (:2499-2567; tiplen 7)
 * on the long-format match string of the primary site, then applyScorePenalty (:2601-2609).  Read r: bases d_bases[read_off[r]..) as sequenced (r.bases),
 * match d_match[match_off[r] .. match_off[r+1]) (empty = r.match == null), mapped = d_flags[r].flags bit0.  d_penalty[r] = the penalty applied;
 * d_status[r] (may be NULL): bit0 the string ended inside a tip (the reference would throw), bit1 short-format digits (not supported). */
int  bbm_sitelist_tip_penalty_dev(bbm_ctx* ctx, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                  const int8_t* d_bases, const int8_t* d_match, const int64_t* d_match_off, const bbm_read_out* d_flags,
                                  int32_t tiplen, int32_t* d_penalty, int32_t* d_status, void* stream);

/* Read.setFromTopSite / setFromSite (current/stream/Read.java:1171-1190, 1213-1224; processRead :557) for unpaired reads: the top site of every list becomes the
 * record bbm_sam_batch_* reads (chrom, strand, start, stop, mapScore = slowScore, perfect = ss.perfect, ambiguous from d_flags); an empty list or a cleared
 * mapping gives Read.clearSite (:1278-1286).  d_match_off (nreads+1 entries, may be NULL = no match strings yet, CIGAR '*') places read r's match string. */
int  bbm_sam_tasks_from_lists_dev(bbm_ctx* ctx, const bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                                  const bbm_read_out* d_flags, const int64_t* d_match_off, bbm_sam_task* d_tasks, void* stream);

/* ---- BBMapThread.scoreSlow over a batch of reads, in rounds (current/align2/BBMapThread.java:252-386; part of SURVEY 8f.1) ----
 * Round k slow-aligns the k-th site of every read whose run[r] != 0 (processRead calls scoreSlow when scoreNoIndels found no
 * near-perfect site, :463-465): preamble, MSA.fillAndScoreLimited(bases, ss, SLOW_ALIGN_PADDING, max(slowScore, minMsaLimit)), the
 * "more padding" retry with EXTRA_PADDING, setSlowScore/setLimits, the minMsaLimit ratchet and the perfect/semiperfect flags.  Default
 * flags (QUICK_MATCH_STRINGS=false).  Sites with a gap array are aligned against the gapped reference (bbm_msa_gapped semantics) and keep their
 * gap array consistent through SiteScore.setLimits/setStop + GapTools.fixGaps.  status[r] (may be NULL): BBM_SLOW_* bits. */
typedef struct {                /* 32 bytes; defaults: MINIMUM_ALIGNMENT_SCORE_RATIO 0.56, ..._PRE_RESCUE (paired only), CLEARZONE1e 258, CLEARZONE3 800,
                                   SLOW_ALIGN_PADDING 4, EXTRA_PADDING 10, EXPECTED_LEN_LIMIT (3000*17)/20-2*(4+10) = 2522 */
    int32_t paired; float min_ratio, min_ratio_pre_rescue;
    int32_t clearzone1e, clearzone3, slow_align_padding, extra_padding, expected_len_limit;
} bbm_slow_cfg;
#define BBM_SLOW_GAPPED         1   /* reserved (round 1 left gapped sites unaligned; they now go through the gapped aligner and GapTools.fixGaps) */
#define BBM_SLOW_ALIGNER_ERROR  2   /* the aligner reported a per-task error (shape outside 601 x 3000) */
int  bbm_scoreslow_dev(bbm_ctx* ctx, bbm_ss* d_lists, const int32_t* d_nss, int64_t nreads, int32_t cap, const int64_t* d_read_off,
                       const int8_t* d_basesP, const int8_t* d_basesM, const int8_t* d_refs, const int64_t* d_chrom_off, const int32_t* d_run,
                       const bbm_slow_cfg* cfg, int32_t* d_status, int32_t max_read_len, void* stream, int64_t* alignments_out, float* ms_out);
int  bbm_scoreslow_host(bbm_ctx* ctx, bbm_ss* lists, const int32_t* nss, int64_t nreads, int32_t cap, const int64_t* read_off,
                        const int8_t* basesP, const int8_t* basesM, const int8_t* d_refs, const int64_t* chrom_off, int32_t nchroms,
                        const int32_t* run, const bbm_slow_cfg* cfg, int32_t* status, int64_t* alignments_out);


/* ---- the batched mapper: BBMapThread.processRead / processReadPair over a batch of reads, FASTQ-shaped bytes in, SAM records out ----
 * (current/align2/BBMapThread.java:389-733, 943-1362; AbstractMapThread.quickMap :643-755; genMatchString :860-1068;
 *  TranslateColorspaceRead.realign_new :229-660; SamLine(Read,int) stream/SamLine.java:82-410, toBytes :1925-1960).
 * One call chains every device stage with all intermediates resident: Read.validate -> KeyRing seeds -> BBIndex.find -> removeOutOfBounds ->
 * trimList -> scoreNoIndels -> findTipDeletions -> scoreSlow (rounds) -> final list policy -> genMatchString / realign_new (rounds) ->
 * removeDuplicateBestSites, clearzone 3, toLocalAlignment for X/Y/C tips, score gates, tip-score penalty -> SamLine fields, CIGAR, SAM text.
 * Needs the index (bbm_index_build) and the scaffold table (bbm_map_set_scaffolds) in the context. */
#define BBM_MAP_ST_MATCH_OVERFLOW  1   /* a match string did not fit its slot */
#define BBM_MAP_ST_TIP             2   /* calcTipScorePenalty ran off the match string (the reference would throw) */
#define BBM_MAP_ST_SLOTS           4   /* more than 3 sites of one read needed a match string at the same time */
#define BBM_MAP_ST_ALIGNER         8   /* the aligner reported a per-task error (shape outside 601 x 3000) */
#define BBM_MAP_ST_SITE_OVERFLOW  16   /* BBIndex.find emitted more sites than the mapper's site slots (the reference keeps an unbounded list) */
#define BBM_MAP_ST_SLOW           32   /* scoreSlow reported a status bit for this read */
#define BBM_MAP_ST_LIST_OVERFLOW   64   /* a rescued site did not fit the read's list (or the batch's rescue task slots) */
typedef struct {                /* 80 bytes; defaults in bbmap_b200/mapper.py */
    int32_t paired;             /* reads 2i / 2i+1 are mates (processReadPair) */
    float min_ratio, min_ratio_paired, min_ratio_pre_rescue, secondary_site_score_ratio;   /* MINIMUM_ALIGNMENT_SCORE_RATIO* (AbstractMapThread.java:104-107) */
    int32_t slow_align_padding, max_indel, ambiguous_toss, penalize_ambig;
    int32_t average_pair_dist, max_pair_dist, max_rescue_dist, max_rescue_mismatches;
    int32_t do_rescue, kill_bad_pairs, require_correct_strands, same_strand_pairs;
    int32_t match_slot;         /* bytes a site's match string may take inside the mapper; 0 = 2 * longest read + 128 (enough for indels up to ~1 kbp).  Spliced
                                   reads need read length + intron length: a longer string sets BBM_MAP_ST_MATCH_OVERFLOW and leaves the read without CIGAR */
    int32_t pad_[2];
} bbm_map_cfg;
typedef struct {                /* 48 bytes: the Read fields SamLine(Read,int) reads, as the mapper leaves them */
    int32_t chrom, start, stop, strand, map_score;
    int32_t flags;              /* bit0 mapped, bit1 perfect, bit2 ambiguous, bit3 paired, bit4 rescued, bit5 discarded */
    int32_t match_len, cz3_sub, tip_penalty;
    int32_t status;             /* BBM_MAP_ST_* */
    int32_t match_slot, pad_;
} bbm_map_rec;
typedef struct {                /* every switch of the chain in one record */
    bbm_map_cfg map; bbm_seed_cfg seed; bbm_policy_cfg policy; bbm_slow_cfg slow; bbm_tipdel_cfg tip; bbm_sam_cfg sam;
    int32_t ingest_flags;       /* BBM_ING_* */
    int32_t max_keys;           /* seed slots per read (32) */
    int32_t max_sites;          /* site slots per read for BBIndex.find = list capacity (16; raised automatically up to 64 when a read overflows) */
    int32_t sam_text;           /* 1 = also format SAM lines */
} bbm_mapper_cfg;
typedef struct {                /* what a call did (host) */
    int64_t reads, mapped, slow_alignments, realign_fills, site_overflow_reads, status_reads, sam_bytes;
    int32_t max_sites_used, genmatch_rounds;
    float ms_total, ms_seed_search, ms_lists, ms_slow, ms_genmatch, ms_sam;
    float ms_rescue; int32_t pad_;
    int64_t rescue_scans, rescue_fills;         /* quickRescue scans / slowRescue alignments (pairs) */
    int64_t mated_pairs, inner_length_sum;      /* numMated / innerLengthSum of this batch (AbstractMapThread.java:1543-1557): the host keeps the running
                                                   AVERAGE_PAIR_DIST = innerLengthSum / numMated once numMated > 1000 (BBMapThread.java:1307-1309) and passes it
                                                   with the next batch (bbm_map_cfg.average_pair_dist) */
} bbm_map_stats;
/* Scaffold table (bbm_sam_batch_* layout, host arrays) + scaffold names for RNAME (names_buf/name_off with nscaffolds+1 offsets; may be NULL: "*"). */
int  bbm_map_set_scaffolds(bbm_ctx* ctx, const int32_t* scaf_off, const int32_t* scaf_loc, const int32_t* scaf_len, int32_t nchroms,
                           const int8_t* names_buf, const int64_t* name_off);
/* Everything resident.  d_bases / d_quality are validated in place (Read.validate); d_quality may be NULL (FASTA).  Outputs: d_recs[nreads],
 * d_sam[nreads], the primary match strings at d_match + r * match_stride (may be NULL).  Buffers must be readable 16 bytes past their end. */
int  bbm_map_batch_dev(bbm_ctx* ctx, int8_t* d_bases, int8_t* d_quality, const int64_t* d_read_off, int64_t nreads, int32_t max_read_len,
                       const bbm_mapper_cfg* cfg, bbm_map_rec* d_recs, bbm_sam_out* d_sam, int8_t* d_match, int64_t match_stride,
                       void* stream, bbm_map_stats* stats);
/* The reference-facing call: host buffers in (FASTQ-shaped: bases, phred qualities or NULL, offsets, read names or NULL), host buffers out
 * (records, SamLine fields, primary match strings, and — with cfg->sam_text — the SAM lines: sam_text/sam_off[nreads+1], capacity sam_cap bytes;
 * BBM_E_CAPACITY with *sam_bytes_needed set when too small).  bases/quality are not modified. */
int  bbm_map_batch_host(bbm_ctx* ctx, const int8_t* bases, const int8_t* quality, const int64_t* read_off, int64_t nreads,
                        const int8_t* names, const int64_t* name_off, const bbm_mapper_cfg* cfg, bbm_map_rec* recs, bbm_sam_out* sam,
                        int8_t* match, int64_t match_stride, int8_t* sam_text, int64_t sam_cap, int64_t* sam_off, bbm_map_stats* stats);

/* ---- 1:1 twins of the reference's plain C entry points (single alignment; latency path) ----
 * Same argument meaning as jni/MultiStateAligner11tsJNI.c:100-114 / :361-382.  `packed` (host, 3*(maxRows+1)*(maxColumns+1)
 * ints) receives exactly the cells the reference would have written (values included), so Java's score2/traceback2
 * keep working on it; the penalty tables are fixed to the 11ts constants (the arrays are accepted and ignored). */
int  bbm_fillUnlimited(bbm_ctx* ctx, const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
                       int32_t refStartLoc, int32_t refEndLoc, int32_t* result4, int64_t* iterationsUnlimited,
                       int32_t* packed, int32_t maxRows, int32_t maxColumns);
int  bbm_fillLimitedX(bbm_ctx* ctx, const int8_t* read, const int8_t* ref, int32_t read_length, int32_t ref_length,
                      int32_t refStartLoc, int32_t refEndLoc, int32_t minScore, int32_t* result5, int64_t* iterationsLimited,
                      int32_t* packed, int32_t maxRows, int32_t maxColumns, int32_t bandwidth, float bandwidthRatio,
                      int32_t* vertLimit, int32_t* horizLimit);

#ifdef __cplusplus
}
#endif
#endif /* BBMAP_CUDA_H */
