/*
 * bwamem_b200.h -- C ABI of the B200-native BWA-MEM hot paths (libbwamem_b200.so).
 *
 * Two data-parallel paths, each a drop-in for one library boundary of the reference's
 * `gase_aln` driver (sflorescu/BWA-MEM_GPU):
 *
 *   1. SMEM seeding over the FMD index      replaces GPUSeed   (src/GPUSeed/seed_gen.h:92-106)
 *   2. ksw_extend2-equivalent seed extension replaces GASAL2 KSW (GASAL2/src/gasal_align.h:96-102,
 *                                            ctors.h:5-15, host_batch.h:13, interfaces.h:9)
 *
 * Everything here is plain C: pointers, sizes, opaque handles.  The source-compatible
 * wrappers that carry the reference's own names (seed_gpu(), gasal_aln_async(), ...) live in
 * include/compat/ and forward to these entry points.
 *
 * Results are bit-exact with the reference's CPU functions bwt_smem1 / bwt_sa
 * (src/bwt.c:483-566, bwa_index/bwt.c:151-172) and ksw_extend2 (src/ksw.c:864-986).
 * There is no CPU fallback: every compute entry point returns BWA_B200_ERR_CUDA when no
 * sm_100 device is usable.
 *
 * Conventions: all functions return 0 on success or a negative BWA_B200_ERR_* code;
 * bwa_b200_last_error() returns a thread-local message.  "dev" pointers are device
 * pointers on the handle's device; "host" pointers should be pinned for full PCIe speed
 * (bwa_b200_host_alloc) but pageable memory works.
 */
#ifndef BWAMEM_B200_H
#define BWAMEM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BWA_B200_OK             0
#define BWA_B200_ERR_ARG       -1
#define BWA_B200_ERR_IO        -2
#define BWA_B200_ERR_FORMAT    -3
#define BWA_B200_ERR_CUDA      -4
#define BWA_B200_ERR_NOMEM     -5
#define BWA_B200_ERR_CAPACITY  -6

const char *bwa_b200_last_error(void);
int  bwa_b200_version(void);

/* ------------------------------------------------------------------ runtime */
int  bwa_b200_device_count(void);
/* pinned host memory helpers (replace cudaMallocHost / cudaHostAlloc in the callers,
 * seed_gen.cu:1670-1672, GASAL2/src/host_batch.cpp:17) */
void *bwa_b200_host_alloc(size_t bytes);
void  bwa_b200_host_free(void *p);

/* -------------------------------------------------------------------- index */
/* FMD index in the reference's GPU file layout: .bwt = u64 primary, u64 L2[1..4], then one
 * 32-byte bucket per 64 BWT symbols {u32 cnt[A,C,G,T]; u32 sym[4]} (bwa_index/bwtindex.c:174-197);
 * .sa = sampled suffix array, u32 + packed high bits (bwa_index/bwt.c:472-487). */
typedef struct bwa_b200_index bwa_b200_index_t;

typedef struct {
    uint64_t primary, seq_len, L2[5];
    uint64_t n_buckets;        /* 32-byte occurrence buckets resident in HBM */
    uint64_t n_sa;
    int32_t  sa_intv, pack_size;
    int32_t  device;
    uint64_t hbm_bytes;        /* total device bytes held by this index */
} bwa_b200_index_info_t;

/* replaces bwt_restore_bwt_gpu + bwt_restore_sa_gpu + gpu_cpy_wrapper
 * (seed_gen.cu:1386-1468,1524-1556; called from src/fastmap.c:446-453) */
int  bwa_b200_index_load(const char *bwt_path, const char *sa_path, int device, bwa_b200_index_t **out);
/* same from host arrays (bwt_words = file payload after the 40-byte header; sa has n_sa entries) */
int  bwa_b200_index_from_host(uint64_t primary, const uint64_t L2[5], const uint32_t *bwt_words, uint64_t n_words,
                              const uint32_t *sa, const uint32_t *sa_hi, uint64_t n_sa, int sa_intv, int pack_size,
                              int device, bwa_b200_index_t **out);
/* replicate a resident index onto another device with a peer copy over NVLink */
int  bwa_b200_index_clone_to(const bwa_b200_index_t *src, int device, bwa_b200_index_t **out);
int  bwa_b200_index_info(const bwa_b200_index_t *idx, bwa_b200_index_info_t *info);
/* k-mer interval table beside the index (built on the device at load time for indexes beyond 256 MB of buckets, K = 12 or 13;
 * environment BWA_B200_KMER_K overrides): the suffix-array interval of every pattern of up to K bases, so that an extension step
 * whose result has at most K bases is one 8-byte table load instead of two dependent occurrence-bucket sectors.  Results do not change.  This call
 * rebuilds it with another K (0 = drop it, at most 14: 8 * (4^(K+1) - 4) / 3 bytes); the reference's unused hook for the same idea is
 * pre_calc_seed_intervals_wrapper (seed_gen.h:101, src/fastmap.c:455). */
int  bwa_b200_index_set_kmer_table(bwa_b200_index_t *idx, int K);
void bwa_b200_index_free(bwa_b200_index_t *idx);          /* replaces free_gpuseed_data */

/* Host-side index construction in the reference's on-disk format (bwa_index/bwtindex.c:287-358,
 * build_index.sh): writes <prefix>.bwt (32-bit occ, 64-symbol buckets) and <prefix>.sa; with
 * also_stock_layout != 0 additionally <prefix>.bwt128 (stock 64-bit occ, 128-symbol buckets) so
 * the unmodified CPU bwa can be timed on the same index.  fwd = forward strand, codes 0..3.
 * Output is byte-identical to the reference's `bwa index -s sa -r R` + `bwa index -s bwt`.
 * Texts of 2^32 - 1 rows and more (2*l_pac; human-sized genomes) are sorted with 64-bit suffix indexes and
 * their SA samples carry the high bits in the packed array of bwa_index/bwt.c:78-147 (about 10 bytes of
 * host memory per row); BWA_B200_ERR_CAPACITY if one base occurs 2^32 times or more (the 32-bit bucket counts
 * of the reference's GPU layout could not hold it). */
int  bwa_b200_build_index(const uint8_t *fwd, uint64_t l_pac, int sa_intv, const char *prefix,
                          int also_stock_layout, int n_threads);

/* ---------------------------------------------------------------- read batch */
/* 4-bit packing used on the wire and in HBM: 8 bases per u32, base 0 in bits 31..28, codes
 * A0 C1 G2 T3 N4 (as pack_4bit_fow, seed_gen.cu:1088-1108); every read starts on a word
 * boundary.  word_off has n_reads+1 entries. */
size_t bwa_b200_packed_words(const uint32_t *read_len, uint64_t n_reads);
int  bwa_b200_pack_ascii(const char *bases, const uint64_t *base_off, uint64_t n_reads,
                         uint32_t *packed, uint64_t *word_off, uint32_t *read_len, int n_threads);
int  bwa_b200_pack_codes(const uint8_t *codes, const uint64_t *base_off, uint64_t n_reads,
                         uint32_t *packed, uint64_t *word_off, uint32_t *read_len, int n_threads);

/* ------------------------------------------------------------------- seeding */
typedef struct bwa_b200_seeder bwa_b200_seeder_t;

typedef struct {
    int32_t min_seed_len;      /* opt->min_seed_len, default 19                              */
    int32_t max_occ;           /* > 0: locate only the rows mem_chain reads (bwa_index/bwamem.c:278-283):
                                  count = min(s, max_occ) rows k + t*step, step = s > max_occ ? s/max_occ : 1.
                                  <= 0: locate all s rows of every SMEM (reference GPU layout,
                                  seed_gen.cu:520-545)                                       */
    int32_t reseed;            /* 0: SMEM pass 1 only -- what the reference GPU path does (README.md:93, the
                                  "no re-seeding" limitation).  != 0: also passes 2 and 3 of mem_collect_intv
                                  (bwa_index/bwamem.c:132-161), i.e. the seed set of stock `bwa mem`; the four
                                  fields below are then read                                 */
    float   split_factor;      /* opt->split_factor  (1.5): pass 2 splits SMEMs of length >= min_seed_len * this */
    int32_t split_width;       /* opt->split_width   (10):  ... that have at most this many occurrences          */
    int32_t max_mem_intv;      /* opt->max_mem_intv  (20):  pass 3 (bwt_seed_strategy1) threshold; <= 0 skips it */
} bwa_b200_seed_params_t;
/* min_seed_len 19, max_occ 500, reseed off; split_factor 1.5, split_width 10, max_mem_intv 20 (bwa_index/bwamem.c:56-62) */
void bwa_b200_seed_params_default(bwa_b200_seed_params_t *p);

/* flat result, layout of mem_seed_v_gpu (seed_gen.h:68-75): seeds of read r are
 * [seed_off[r], seed_off[r] + n_seeds[r]), ordered by SMEM (ascending query start) then SA row;
 * score = occurrence count s of the SMEM on the first seed of each SMEM group, 0 on the others. */
typedef struct {
    uint64_t  n_reads, n_seeds;
    uint64_t *rbeg;            /* reference position in [0, 2*l_pac)          */
    int32_t  *qbeg_qend;       /* pairs {qbeg, qend}  (int2 in the reference) */
    uint32_t *score;
    uint32_t *n_seeds_per_read;
    uint64_t *seed_off;        /* exclusive prefix sum                        */
} bwa_b200_seeds_t;

int  bwa_b200_seeder_create(const bwa_b200_index_t *idx, uint64_t max_reads, uint64_t max_words,
                            bwa_b200_seeder_t **out);
void bwa_b200_seeder_destroy(bwa_b200_seeder_t *s);

/* host in, host out: H2D of the packed batch, kernels, D2H of the seeds.  The arrays of `out`
 * are allocated with malloc() (the reference's caller frees them with free(), src/fastmap.c:537-542). */
int  bwa_b200_seed_host(bwa_b200_seeder_t *s, const uint32_t *packed, const uint64_t *word_off,
                        const uint32_t *read_len, uint64_t n_reads, const bwa_b200_seed_params_t *p,
                        bwa_b200_seeds_t *out);
void bwa_b200_seeds_free(bwa_b200_seeds_t *r);

/* device in, device out (inputs already resident): enqueues on the seeder's stream.  The result
 * stays in the seeder's workspace until the next call; bwa_b200_seed_device_result synchronises
 * and reports the device pointers. */
int  bwa_b200_seed_device(bwa_b200_seeder_t *s, const uint32_t *dev_packed, const uint64_t *dev_word_off,
                          const uint32_t *dev_read_len, uint64_t n_reads, const bwa_b200_seed_params_t *p);
int  bwa_b200_seed_device_result(bwa_b200_seeder_t *s, bwa_b200_seeds_t *dev_view);
void *bwa_b200_seeder_stream(bwa_b200_seeder_t *s);       /* cudaStream_t */
/* per-phase launch counters and SMEM-only view for tests: qbeg,qend,k,s per SMEM in read order */
int  bwa_b200_seed_device_smems(bwa_b200_seeder_t *s, uint64_t n_reads, uint32_t *host_n_smems,
                                int32_t *host_qbeg, int32_t *host_qend, uint64_t *host_k, uint64_t *host_s,
                                uint64_t cap, uint64_t *total);
uint64_t bwa_b200_seeder_launches(const bwa_b200_seeder_t *s);
/* what the pass-1 kernels actually ask the memory system for (bench.py's roofline): enable != 0 switches the counting on for the
 * following batches; out (may be NULL) receives {fwd_kernel bucket sectors, fwd_kernel k-mer table entries, back_kernel bucket
 * sectors, back_kernel k-mer table entries} of the last counted batch */
int  bwa_b200_seeder_request_counts(bwa_b200_seeder_t *s, int enable, uint64_t out[4]);
/* random 32-byte-sector gather throughput (GB/s) over a scratch buffer of `bytes`: the measured
 * denominator for the seeding roofline; > L2-sized buffers measure HBM, small ones measure L2 */
double bwa_b200_measure_random_sector_gbs(int device, uint64_t bytes, int iters, int reps);
/* measured denominator for the extension roofline: issue rate, in warp-instructions per clock per SM, of the integer instructions
 * the extension kernels are made of (op i is named by bwa_b200_int_alu_op_name(i): IADD3, LOP3, PRMT, VIMNMX.S32, VIADDMNMX.S32,
 * VIMNMX3.S32, VIADDMNMX.S16x2, VIMNMX3.S16x2, IMAD), at full occupancy with independent chains; *sm_mhz = the SM clock measured
 * under that load, *n_sm = the SM count.  Returns the number of rates written (9) or a negative error code. */
int  bwa_b200_measure_int_alu(int device, double *rates, int cap, double *sm_mhz, int *n_sm);
const char *bwa_b200_int_alu_op_name(int i);

/* ----------------------------------------------------------------- extension */
typedef struct bwa_b200_extender bwa_b200_extender_t;

/* the arguments of ksw_extend2 that are constant over a batch (src/ksw.c:864-866) plus the
 * clip penalty of the local-vs-to-end rule (src/bwamem.c:1892-1901).  The reference boundary
 * cannot express w/zdrop/end_bonus/o_ins/e_ins; this struct is how they arrive. */
typedef struct {
    int8_t  mat[25];           /* 5x5 substitution matrix, row = target base (bwa_fill_scmat) */
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t w;                 /* band width                                                 */
    int32_t end_bonus;
    int32_t zdrop;
    int32_t use_band;          /* opt_ext of the fork's ksw_extend2 (src/ksw.c:902-907)       */
    int32_t pen_clip;          /* for the (score,qend,tend) triple                            */
} bwa_b200_ext_params_t;

void bwa_b200_ext_params_default(bwa_b200_ext_params_t *p);   /* a=1 b=4 o=6 e=1 w=100 zdrop=100 clip=5 */
void bwa_b200_fill_scmat(int a, int b, int8_t mat[25]);

typedef struct {               /* all six outputs of ksw_extend2 */
    int32_t score, qle, tle, gtle, gscore, max_off;
} bwa_b200_ext_result_t;

int  bwa_b200_extender_create(int device, uint64_t max_jobs, uint64_t max_query_bytes, uint64_t max_target_bytes,
                              bwa_b200_extender_t **out);
void bwa_b200_extender_destroy(bwa_b200_extender_t *e);

/* Job batch in the GASAL host layout (GASAL2/src/host_batch.cpp:79-153, src/bwamem.c:1102-1167):
 * one byte per base, codes 0..4, job a = qseq[qoff[a] .. qoff[a]+qlen[a]) vs tseq[toff[a] ..);
 * h0[a] = host_seed_scores[a].  Asynchronous: returns after enqueueing H2D, kernels and D2H on the
 * extender's stream (gasal_aln_async).  res6 / triple may be NULL.  triple receives
 * {aln_score, query_batch_end, target_batch_end} per job after the local-vs-to-end rule. */
int  bwa_b200_extend_async(bwa_b200_extender_t *e, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                           const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                           const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen,
                           const uint32_t *h0, bwa_b200_ext_result_t *res6,
                           int32_t *aln_score, int32_t *query_end, int32_t *target_end);
/* same with each sequence buffer given as a list of host pages (the linked pinned pages of
 * GASAL2's host_batch_t, GASAL2/src/gasal_align.cu:146-169): page i covers batch bytes
 * [offset, offset + bytes) */
typedef struct { const uint8_t *data; uint64_t offset, bytes; } bwa_b200_host_page_t;
int  bwa_b200_extend_async_paged(bwa_b200_extender_t *e, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                                 const bwa_b200_host_page_t *qpages, int n_qpages, uint64_t q_bytes,
                                 const uint32_t *qoff, const uint32_t *qlen,
                                 const bwa_b200_host_page_t *tpages, int n_tpages, uint64_t t_bytes,
                                 const uint32_t *toff, const uint32_t *tlen,
                                 const uint32_t *h0, bwa_b200_ext_result_t *res6,
                                 int32_t *aln_score, int32_t *query_end, int32_t *target_end);
/* 0 = done, 1 = still running (gasal_is_aln_async_done returns 0 / -1) */
int  bwa_b200_extend_query(bwa_b200_extender_t *e);
int  bwa_b200_extend_wait(bwa_b200_extender_t *e);

/* device-resident variant: sequences already packed 4-bit in HBM (8 bases per u32, first base in
 * the high nibble, each sequence starting on a word boundary; offsets in bases, multiples of 8). */
int  bwa_b200_extend_device(bwa_b200_extender_t *e, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                            const uint32_t *dev_qpacked, const uint32_t *dev_qoff, const uint32_t *dev_qlen,
                            const uint32_t *dev_tpacked, const uint32_t *dev_toff, const uint32_t *dev_tlen,
                            const uint32_t *dev_h0, bwa_b200_ext_result_t *dev_res6);
/* pack a byte-per-base device buffer into the 4-bit layout (gasal_pack_kernel equivalent) */
int  bwa_b200_pack_device(bwa_b200_extender_t *e, const uint8_t *dev_bytes, uint64_t n_bytes, uint32_t *dev_packed);
void *bwa_b200_extender_stream(bwa_b200_extender_t *e);
uint64_t bwa_b200_extender_launches(const bwa_b200_extender_t *e);
/* cells evaluated by the last batch (sum over rows of end-beg), counted on device */
uint64_t bwa_b200_extender_last_cells(bwa_b200_extender_t *e);
/* Jobs of the last batch answered without a matrix: a query that equals the head of its target except for a few substituted
 * bases (up to six at the default penalties; the usual flank of a maximal exact match: the base that ended the match and little else) has a closed-form ksw_extend2 result
 * when the scoring is bwa_fill_scmat's, min(o_del + e_del, o_ins + e_ins) > a + b and -- with two or more substitutions -- no diagonal an
 * affordable gap reaches matches all the way between two neighbouring ones.  Proof and conditions: bwa-mem_gpu_b200/csrc/ext_pair_core.cuh
 * (closed_form_job).  Such jobs are not in last_cells.  set_closed_form(e, 0) sends every job through the kernels (also: environment
 * BWA_B200_EXT_NO_CLOSED at creation); the results are identical either way. */
uint64_t bwa_b200_extender_last_closed_form(bwa_b200_extender_t *e);
int  bwa_b200_extender_set_closed_form(bwa_b200_extender_t *e, int enable);

/* ------------------------------------------------- fused seed -> extend pipeline */
/* B200-first entry point for the headline metric (reads/s, seed + extend): reads go in once,
 * SMEM seeds stay in HBM, the extension jobs of each read are cut on the device from a resident
 * 2-bit copy of the reference (the .pac content) and extended, and one fixed-size record per read
 * comes back.  Job shapes follow mem_chain2aln for a one-seed chain (cal_max_gap / rmax window,
 * src/bwamem.c:996-1002,1180-1201; left and right both start from h0 = seed_len * a as in the fork,
 * src/bwamem.c:1360,1384).  The seed extended is the longest located seed of the read (first on
 * ties) that does not bridge the forward/reverse boundary; chaining proper stays with the caller
 * (it is not on the hot path).  The two-boundary flow (bwa_b200_seed_host, then
 * bwa_b200_extend_async) remains available and returns the same numbers. */
int  bwa_b200_index_attach_ref(bwa_b200_index_t *idx, const uint8_t *fwd_codes, uint64_t l_pac);

typedef struct {
    int64_t seed_rbeg;              /* -1 when the read has no usable seed */
    int32_t seed_qbeg, seed_qend;
    int32_t n_seeds;                /* located seeds of the read           */
    int32_t h0;
    bwa_b200_ext_result_t left, right;   /* score = h0 and gscore = -1 where no extension was needed */
} bwa_b200_read_result_t;

typedef struct bwa_b200_pipeline bwa_b200_pipeline_t;
int  bwa_b200_pipeline_create(const bwa_b200_index_t *idx, uint64_t max_reads, uint64_t max_words,
                              uint32_t max_read_len, bwa_b200_pipeline_t **out);
void bwa_b200_pipeline_destroy(bwa_b200_pipeline_t *p);
/* host buffers in and out: H2D, kernels, D2H, synchronised on return */
int  bwa_b200_seed_extend_host(bwa_b200_pipeline_t *p, const uint32_t *packed, const uint64_t *word_off,
                               const uint32_t *read_len, uint64_t n_reads, const bwa_b200_seed_params_t *sp,
                               const bwa_b200_ext_params_t *ep, bwa_b200_read_result_t *host_out);
/* device buffers in and out, asynchronous on the pipeline stream; call _sync before reading */
int  bwa_b200_seed_extend_device(bwa_b200_pipeline_t *p, const uint32_t *dev_packed, const uint64_t *dev_word_off,
                                 const uint32_t *dev_read_len, uint64_t n_reads, uint32_t max_read_len,
                                 const bwa_b200_seed_params_t *sp, const bwa_b200_ext_params_t *ep,
                                 bwa_b200_read_result_t *dev_out);
int  bwa_b200_pipeline_sync(bwa_b200_pipeline_t *p);
void *bwa_b200_pipeline_stream(bwa_b200_pipeline_t *p);
uint64_t bwa_b200_pipeline_launches(const bwa_b200_pipeline_t *p);
/* totals of the last batch: [0] seeds, [1] extension jobs with qlen > 0, [2] DP cells evaluated */
int  bwa_b200_pipeline_totals(bwa_b200_pipeline_t *p, uint64_t out[3]);
/* per-kernel device time of the last batch, measured with CUDA events on the pipeline stream:
 * enable, run a batch, sync, then read (name, milliseconds) pairs; returns the number of records */
int  bwa_b200_pipeline_profile(bwa_b200_pipeline_t *p, int enable);
int  bwa_b200_pipeline_kernel_times(bwa_b200_pipeline_t *p, const char **names, float *ms, int cap);

/* ------------------------------------------- global alignment with backtrack: CIGAR, score, NM */
/* The dynamic programming of the reference's output stage: mem_reg2aln -> bwa_gen_cigar2 -> ksw_global2
 * (src/bwamem.c:2344-2438, src/bwa.c:111-216, src/ksw.c:1120-1241).  A job is (query, target, band w): both ends fixed, the
 * arguments of ksw_global2; mat / o_del / e_del / o_ins / e_ins are the fields of the same name in the extension parameter block; its other fields are ignored.
 * Per job: the score ksw_global2 returns, the CIGAR (len << 4 | op, op 0 M / 1 I / 2 D, BAM encoding as in the reference) and
 * the NM of bwa_gen_cigar2 (mismatches + inserted + deleted bases, a deletion at either end of the CIGAR not counted,
 * src/bwa.c:178-205).  The caller fetches the reference window and reverses both sequences for a reverse-strand hit, as
 * bwa_gen_cigar2 does (src/bwa.c:139-150), and takes the band from bwa_b200_cigar_band.  The MD string is text formatting and
 * stays with the caller.  Bands wider than 127 are refused (BWA_B200_ERR_ARG). */
typedef struct bwa_b200_cigar bwa_b200_cigar_t;
typedef struct {
    uint64_t  n_jobs, n_ops;
    int32_t  *score, *nm;
    uint32_t *n_cigar;
    uint64_t *cigar_off;       /* exclusive prefix sum of n_cigar */
    uint32_t *cigar;           /* n_ops operations, job after job */
} bwa_b200_cigars_t;
int  bwa_b200_cigar_create(int device, bwa_b200_cigar_t **out);
void bwa_b200_cigar_destroy(bwa_b200_cigar_t *c);
/* the band bwa_gen_cigar2 passes to ksw_global2 for a query of l_query bases against rlen reference bases when it was
 * called with band w_ (src/bwa.c:161-169) */
int  bwa_b200_cigar_band(const bwa_b200_ext_params_t *p, int w_, int l_query, int64_t rlen);
/* host in (byte per base, codes 0..4, as bwa_b200_extend_async), host out: arrays of `out` are malloc'ed, release them with
 * bwa_b200_cigars_free */
int  bwa_b200_global_host(bwa_b200_cigar_t *c, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                          const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                          const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen,
                          const uint32_t *w, bwa_b200_cigars_t *out);
void bwa_b200_cigars_free(bwa_b200_cigars_t *r);
/* the same with the results in pinned host buffers owned by the handle (valid until its next call or its destruction; not to be
 * freed): the per-batch path of a driver -- no allocation, no pageable staging; with the inputs in pinned memory too, every copy
 * is one asynchronous DMA (the reference keeps its result arrays pinned for the same reason, GASAL2/src/res.cpp:10-60) */
int  bwa_b200_global_host_view(bwa_b200_cigar_t *c, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                               const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                               const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen,
                               const uint32_t *w, bwa_b200_cigars_t *view);
/* sequences and job tables already in HBM; host copies of qlen, tlen and w drive the band-width binning and the sizing.
 * aligned8 != 0: the caller guarantees the GASAL layout -- every offset a multiple of 8 and every sequence padded to a multiple of 8
 * bytes inside its buffer -- so the kernel loads 8 bases at a time (bwa_b200_global_host checks this itself).  Results stay on the
 * device until the next call: bwa_b200_global_device_view synchronises and reports the device pointers. */
int  bwa_b200_global_device(bwa_b200_cigar_t *c, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                            const uint8_t *dev_qseq, const uint32_t *dev_qoff, const uint32_t *dev_qlen,
                            const uint8_t *dev_tseq, const uint32_t *dev_toff, const uint32_t *dev_tlen,
                            const uint32_t *host_qlen, const uint32_t *host_tlen, const uint32_t *host_w, int aligned8);
int  bwa_b200_global_device_view(bwa_b200_cigar_t *c, bwa_b200_cigars_t *dev_view);
void *bwa_b200_cigar_stream(bwa_b200_cigar_t *c);
uint64_t bwa_b200_cigar_launches(const bwa_b200_cigar_t *c);
uint64_t bwa_b200_cigar_last_cells(const bwa_b200_cigar_t *c);     /* DP cells of the last batch (sum of end - beg over rows) */
int  bwa_b200_cigar_profile(bwa_b200_cigar_t *c, int enable);
int  bwa_b200_cigar_kernel_times(bwa_b200_cigar_t *c, const char **names, float *ms, int cap);

/* mem_reg2aln over a batch (src/bwamem.c:2344-2438): per alignment region the band inference (infer_bw), bwa_gen_cigar2 with its
 * band-doubling retries -- each retry is one more wave of the kernel over the regions still improving --, the squeeze of a leading or
 * trailing deletion, the soft clips, the forward position and contig.  Query and reference windows are cut on the device from the
 * packed reads and the reference attached to the index (bwa_b200_index_attach_ref), in the orientation bwa_gen_cigar2 aligns them.
 * What stays with the caller: mapq, flags, the MD text.  ctg_off = bntann1_t offsets; match_score = opt->a; p->w = opt->w. */
typedef struct { uint32_t read; int32_t qb, qe; int64_t rb, re; int32_t truesc, w; } bwa_b200_aln_in_t;   /* the mem_alnreg_t fields it reads */
typedef struct {                   /* the mem_aln_t fields it sets */
    int64_t  pos;                  /* 0-based position on contig rid, forward strand; -1 for an unmapped record (rb or re < 0) */
    int32_t  rid, is_rev;
    int32_t  score;                /* the last global-alignment score (the value mem_reg2aln compares with truesc)             */
    int32_t  nm, n_cigar;          /* CIGAR in BAM encoding, soft clips (op 3) included                                        */
    int32_t  band, n_waves;        /* w2 after the retry loop; calls of bwa_gen_cigar2                                          */
    uint64_t cigar_off;            /* first operation in the flat CIGAR array                                                  */
} bwa_b200_aln_out_t;
/* `cigar` is malloc'ed (n_ops operations): release it with free() */
int  bwa_b200_reg2aln_host(bwa_b200_cigar_t *c, const bwa_b200_index_t *idx, int32_t n_ctg, const int64_t *ctg_off,
                           const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len, uint64_t n_reads,
                           const bwa_b200_aln_in_t *alns, uint64_t n_alns, const bwa_b200_ext_params_t *p, int32_t match_score,
                           bwa_b200_aln_out_t *out, uint32_t **cigar, uint64_t *n_ops);

/* -------------------------------------------------- local alignment of mate rescue: ksw_align2 */
/* The Smith-Waterman call of mate rescue (mem_matesw, src/bwamem_pair.c:119-175, call at :159) and of mem_seed_sw (src/bwamem.c:774-830):
 * ksw_align2 (src/ksw.c:698-736) over the fork's striped kernels ksw_u8 / ksw_i16 (src/ksw.c:440-696, the SSE2 build: opt->use_avx2 = 0).
 * One job = (query, target, xtra): xtra carries the reference's flags -- KSW_XBYTE 0x10000 (byte kernel), KSW_XSTOP 0x20000, KSW_XSUBO
 * 0x40000 (second-best above xtra & 0xffff), KSW_XSTART 0x80000 (second pass over the reversed prefixes for the start positions).
 * The result is kswr_t field for field.  Bit-exact with the reference's SIMD code, whose values differ from the textbook recurrence
 * (E is not corrected after the lazy-F pass, F restarts at every stripe): the kernel replays the striped order, see csrc/sw_core.cuh.
 * A byte-kernel overflow reports score 255 and te, as the reference's first pass does (its second pass is undefined there).
 * mat / o_del / e_del / o_ins / e_ins are read from the extension parameter block; sequences are byte-per-base codes 0..4. */
typedef struct bwa_b200_sw bwa_b200_sw_t;
typedef struct { int32_t score, te, qe, score2, te2, tb, qb; } bwa_b200_sw_result_t;
int  bwa_b200_sw_create(int device, bwa_b200_sw_t **out);
void bwa_b200_sw_destroy(bwa_b200_sw_t *s);
int  bwa_b200_sw_align2_host(bwa_b200_sw_t *s, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                             const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                             const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen,
                             const uint32_t *xtra, bwa_b200_sw_result_t *out);
uint64_t bwa_b200_sw_launches(const bwa_b200_sw_t *s);
float bwa_b200_sw_last_kernel_ms(const bwa_b200_sw_t *s);      /* device time of the last batch's kernel (CUDA events on its stream) */

/* ------------------------------------- seeds -> chains -> extension jobs -> alignment regions */
/* The step between the two hot paths in the reference worker (src/bwamem.c:2055-2093 and :2286-2306), on the
 * device, so that a read batch goes seeds -> chains -> jobs -> extension -> regions without leaving HBM
 * (SURVEY.md 8f row 1).  Bit-exact with the fork's host code:
 *   mem_chain (:404-476, test_and_merge :337-359, the chain kbtree of src/kbtree.h), mem_chain_flt (:488-560),
 *   mem_chain2aln (:1170-1479: rmax window per chain, seeds by descending score, the estimated-extent test that
 *   skips seeds inside an earlier region, left / right jobs with h0 = seed length, SHORT / LONG batch choice),
 *   the local-vs-to-end rule (:1892-1901) and the region arithmetic (:2286-2306).
 * mem_flt_chained_seeds (:970-990) acts on reads of 757 bases and more at the default options (5.5 ln L <= 0.05 L): every seed of a kept
 * chain is scored by mem_seed_sw (:774-808, ksw_align2 around the seed) and the low ones dropped.  That runs on the device as well
 * (seedsw_kernel / chain_long_kernel, csrc/chain.cu), so reads of any length go through the same call.
 * mem_sort_dedup_patch and everything after it stay with the caller. */
typedef struct {               /* the mem_opt_t fields this stage reads (src/bwamem.h:34-73) */
    int32_t a, b, o_del, e_del, o_ins, e_ins, w;
    int32_t min_seed_len, max_occ, max_chain_gap, min_chain_weight, max_chain_extend;
    float   mask_level, drop_ratio;
} bwa_b200_chain_params_t;
void bwa_b200_chain_params_default(bwa_b200_chain_params_t *p);      /* mem_opt_init, src/bwamem.c:107-150 */

typedef struct {               /* mem_chain_t after mem_chain_flt (src/bwamem.c:318-324) */
    int64_t pos;               /* rbeg of the first seed (the kbtree key)              */
    int32_t rid, n, w, kept, first, is_alt;
    float   frac_rep;
    int32_t seed_off;          /* first seed of the chain within the read's chain seeds */
} bwa_b200_chain_t;
typedef struct { int64_t rbeg; int32_t qbeg, len, score, pad; } bwa_b200_chain_seed_t;      /* mem_seed_t */

typedef struct {               /* one mem_alnreg_t (src/bwamem.h:82-112), the fields this stage sets */
    int64_t rb, re;            /* [rb, re) on the reference, after extension            */
    int64_t rb_est, re_est, target_seed_begin;
    int32_t qb, qe, score, truesc;
    int32_t qb_est, qe_est, rid, align_sides, where_is_long, query_seed_begin, seedlen0, seedcov, w;
    float   frac_rep;
    int32_t left_tlen, right_tlen;   /* target window lengths of the left / right extension job    */
    int32_t job_short, job_long;     /* index of the region's job in the SHORT / LONG batch, or -1 */
} bwa_b200_region_t;

typedef struct { uint32_t qoff, qlen, toff, tlen, h0; } bwa_b200_job_t;   /* offsets in bases, multiples of 8 */

/* flat result of a batch; arrays are malloc'ed by the call and released with bwa_b200_alignments_free.
 * Regions of read r are regions[region_off[r] .. + n_regions_per_read[r]), in the order mem_chain2aln creates
 * them.  With want_detail != 0 the chains and the extension jobs are returned as well: jobs[0 .. n_jobs_short)
 * are the SHORT batch in the reference's order, the LONG batch follows; job sequences are 4-bit packed. */
typedef struct {
    uint64_t n_reads, n_regions, n_chains, n_chain_seeds, n_jobs_short, n_jobs_long, q_words, t_words;
    uint32_t *n_regions_per_read; uint64_t *region_off; bwa_b200_region_t *regions;
    uint32_t *n_chains_per_read;  uint64_t *chain_off;  bwa_b200_chain_t *chains;
    uint64_t *chain_seed_off;     bwa_b200_chain_seed_t *chain_seeds;
    bwa_b200_job_t *jobs; uint32_t *qpacked, *tpacked; bwa_b200_ext_result_t *job_res;
} bwa_b200_alignments_t;
void bwa_b200_alignments_free(bwa_b200_alignments_t *a);

typedef struct bwa_b200_aligner bwa_b200_aligner_t;
/* idx must have the forward reference attached (bwa_b200_index_attach_ref).  Contigs default to one sequence
 * [0, l_pac); set the reference's bntann1_t offsets / lengths / is_alt flags for a multi-sequence index. */
int  bwa_b200_aligner_create(const bwa_b200_index_t *idx, uint64_t max_reads, uint64_t max_words, bwa_b200_aligner_t **out);
int  bwa_b200_aligner_set_contigs(bwa_b200_aligner_t *a, int32_t n, const int64_t *offset, const int32_t *len, const int32_t *is_alt);
void bwa_b200_aligner_destroy(bwa_b200_aligner_t *a);
/* reads in, regions out: seeding (as bwa_b200_seed_host) then chains, jobs, extension, regions */
int  bwa_b200_align_host(bwa_b200_aligner_t *a, const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len,
                         uint64_t n_reads, const bwa_b200_seed_params_t *sp, const bwa_b200_chain_params_t *cp,
                         const bwa_b200_ext_params_t *ep, int want_detail, bwa_b200_alignments_t *out);
/* the same, regions only, into pinned host buffers owned by the aligner (valid until its next call or its
 * destruction; not to be freed): the per-batch path of a driver, no allocation and no pageable staging.  The
 * reference keeps its result arrays pinned for the same reason (GASAL2/src/res.cpp:10-60). */
int  bwa_b200_align_host_view(bwa_b200_aligner_t *a, const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len,
                              uint64_t n_reads, const bwa_b200_seed_params_t *sp, const bwa_b200_chain_params_t *cp,
                              const bwa_b200_ext_params_t *ep, uint64_t *n_regions, const uint32_t **n_regions_per_read,
                              const uint64_t **region_off, const bwa_b200_region_t **regions);
/* the same from given seeds (host arrays in the layout of bwa_b200_seeds_t).  layout_all != 0: every SMEM group
 * holds all `score` rows and is sampled with step score / max_occ, the reference's mem_seed_v_gpu
 * (src/bwamem.c:419-431); 0: a group holds only the sampled rows (bwa_b200_seed_* with max_occ > 0). */
int  bwa_b200_align_seeds_host(bwa_b200_aligner_t *a, const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len,
                               uint64_t n_reads, const bwa_b200_seeds_t *seeds, int layout_all, const bwa_b200_chain_params_t *cp,
                               const bwa_b200_ext_params_t *ep, int want_detail, bwa_b200_alignments_t *out);
/* device-resident batch: reads in HBM, regions left in HBM (bwa_b200_align_device_view); returns when done */
int  bwa_b200_align_device(bwa_b200_aligner_t *a, const uint32_t *dev_packed, const uint64_t *dev_word_off, const uint32_t *dev_read_len,
                           uint64_t n_reads, uint32_t max_read_len, const bwa_b200_seed_params_t *sp,
                           const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep);
typedef struct {
    uint64_t n_reads, n_regions, n_jobs_short, n_jobs_long, n_seeds, cells;
    const uint32_t *n_regions_per_read; const uint64_t *region_off; const bwa_b200_region_t *regions;   /* device pointers */
    uint64_t closed_form_jobs;  /* extension jobs answered in closed form (bwa_b200_extender_last_closed_form); not in cells */
} bwa_b200_align_view_t;
int  bwa_b200_align_device_view(bwa_b200_aligner_t *a, bwa_b200_align_view_t *v);
/* Kept for ABI compatibility: earlier versions left the reads mem_flt_chained_seeds acts on to the caller and listed them here; they are
 * now aligned on the device, so *n is always 0 and *read_idx NULL. */
int  bwa_b200_aligner_skipped_reads(bwa_b200_aligner_t *a, uint64_t *n, const uint32_t **read_idx);
/* ---- the compact host boundary of the aligner: the smallest transfers that carry the same information (SURVEY 8e names the host leg
 * -- PCIe and host memory traffic of 8 GPUs behind one socket -- as the scaling risk; the reference ships one byte per base,
 * GASAL2/src/host_batch.cpp:79-153).
 * Reads: 2 bits per base (A0 C1 G2 T3), 16 bases per u32, base 0 in bits 31..30, every read on a word boundary; bases that are not
 * A/C/G/T are listed apart, one entry (read << 32 | position) each, and patched in on the device.  read_len == NULL: every read
 * has uniform_len bases, and no per-read array crosses the bus at all.  Regions: 40-byte records holding the mem_alnreg_t fields
 * the stages after extension read (src/bwamem.h:82-112); regions of read r follow those of read r - 1, n_regions_per_read[r] of
 * them.  Reads beyond 65535 bases do not fit the record (use bwa_b200_align_host_view).  Results live in pinned buffers owned by
 * the aligner, valid until its next call. */
typedef struct {
    int64_t  rb;                 /* [rb, rb + rlen) on the reference */
    int32_t  rlen;
    uint16_t qb, qe;
    int32_t  score, truesc, seedcov, rid;
    uint16_t w, seedlen0;
    float    frac_rep;
} bwa_b200_region_compact_t;
size_t bwa_b200_packed2_words(const uint32_t *read_len, uint64_t n_reads, uint32_t uniform_len);
/* codes 0..3 or anything else (= N) / ASCII -> the 2-bit layout; n_list receives up to n_cap entries, *n_n their number (sorted by
 * read, then position); BWA_B200_ERR_CAPACITY when there are more */
int  bwa_b200_pack2_codes(const uint8_t *codes, const uint64_t *base_off, uint64_t n_reads, uint32_t *packed2, uint32_t *read_len,
                          uint64_t *n_list, uint64_t n_cap, uint64_t *n_n, int n_threads);
int  bwa_b200_pack2_ascii(const char *bases, const uint64_t *base_off, uint64_t n_reads, uint32_t *packed2, uint32_t *read_len,
                          uint64_t *n_list, uint64_t n_cap, uint64_t *n_n, int n_threads);
int  bwa_b200_align_host_compact(bwa_b200_aligner_t *a, const uint32_t *packed2, const uint32_t *read_len, uint32_t uniform_len,
                                 uint64_t n_reads, const uint64_t *n_list, uint64_t n_n, const bwa_b200_seed_params_t *sp,
                                 const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep, uint64_t *n_regions,
                                 const uint32_t **n_regions_per_read, const bwa_b200_region_compact_t **regions);

/* ---- every GPU of the box behind one call: a dispatcher deals chunks of chunk_reads reads to worker threads,
 * workers_per_device per device (two keep one chunk's copies under the other's kernels), each with its own aligner; the index
 * -- with its attached reference and k-mer table -- is replicated to devices[1..] by peer copies (bwa_b200_index_clone_to).  The
 * reference has no counterpart: it never selects a device (src/fastmap.c:143).  Input as bwa_b200_align_host_compact (pinned
 * host memory makes the copies asynchronous; the N list sorted by read).  Output, in pinned memory owned by the handle and valid
 * until its next call: n_regions_per_read in read order; the records of chunk k (reads [k * chunk_reads, ...)) are
 * regions[chunk_region_off[k] ...], in read order within the chunk. */
typedef struct bwa_b200_multi bwa_b200_multi_t;
typedef struct {
    uint64_t n_reads, n_regions, n_chunks, chunk_reads;
    const uint32_t *n_regions_per_read;
    const uint64_t *chunk_region_off;
    const bwa_b200_region_compact_t *regions;
} bwa_b200_multi_result_t;
int  bwa_b200_multi_create(const bwa_b200_index_t *idx, const int *devices, int n_devices, int workers_per_device,
                           uint64_t chunk_reads, uint32_t max_read_len, bwa_b200_multi_t **out);
int  bwa_b200_multi_set_contigs(bwa_b200_multi_t *m, int32_t n, const int64_t *offset, const int32_t *len, const int32_t *is_alt);
int  bwa_b200_multi_align_compact(bwa_b200_multi_t *m, const uint32_t *packed2, const uint32_t *read_len, uint32_t uniform_len,
                                  uint64_t n_reads, const uint64_t *n_list, uint64_t n_n, const bwa_b200_seed_params_t *sp,
                                  const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep, bwa_b200_multi_result_t *out);
/* The same with two batches in flight: submit returns at once with a ticket (0 or 1), wait blocks until that batch is done and hands
 * out its results, which stay valid until the second submit after it.  The workers drain the older batch first and go straight on to
 * the next, so the tail of one batch (its last chunk's results on their way to the host) and the head of the next (its first chunk's
 * reads on their way to the device) run under kernels instead of between calls -- how a driver that streams batches uses the library
 * (the reference keeps NB_STREAMS = 2 batches per thread in flight the same way, src/fastmap.c:31, src/bwamem.c:2040-2181).  The read
 * buffers of a batch must stay untouched until its wait returns; the parameter blocks are copied.  Submit and wait from one thread.
 * bwa_b200_multi_align_compact is submit + wait. */
int  bwa_b200_multi_submit_compact(bwa_b200_multi_t *m, const uint32_t *packed2, const uint32_t *read_len, uint32_t uniform_len,
                                   uint64_t n_reads, const uint64_t *n_list, uint64_t n_n, const bwa_b200_seed_params_t *sp,
                                   const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep, int *ticket);
int  bwa_b200_multi_wait(bwa_b200_multi_t *m, int ticket, bwa_b200_multi_result_t *out);
int  bwa_b200_multi_n_workers(const bwa_b200_multi_t *m);
uint64_t bwa_b200_multi_worker_chunks(const bwa_b200_multi_t *m, int worker);   /* chunks a worker has processed since creation */
uint64_t bwa_b200_multi_launches(const bwa_b200_multi_t *m);
void bwa_b200_multi_destroy(bwa_b200_multi_t *m);
void *bwa_b200_aligner_stream(bwa_b200_aligner_t *a);
uint64_t bwa_b200_aligner_launches(const bwa_b200_aligner_t *a);
int  bwa_b200_aligner_profile(bwa_b200_aligner_t *a, int enable);
int  bwa_b200_aligner_kernel_times(bwa_b200_aligner_t *a, const char **names, float *ms, int cap);

/* ---------------------------------------------------------------- regions -> the records SAM is written from
 * What the reference does to a read's alignment regions once the extension results are in: mem_sort_dedup_patch (redundant hits
 * dropped, colinear hits joined when one global alignment of the joint span scores within 10 % of the prediction; src/bwamem.c:
 * 580-681), the is_alt marking of its caller (:2321-2325), mem_mark_primary_se (:715-760; id = index of the read in the run, it
 * seeds the tie-breaking hash) and mem_approx_mapq_se as mem_reg2aln applies it (:1690-1716, :2363: 0 for secondary hits).
 * One GPU lane per read runs the reference's own sequence of comparisons (its introsort is not stable, equal keys are common).
 * The records carry the mem_alnreg_t fields the stage reads or writes (src/bwamem.h:83-112). */
typedef struct {
    int32_t a, b, o_del, e_del, o_ins, e_ins, w, min_seed_len, max_chain_gap, mapQ_coef_fac;   /* mapQ_coef_fac is an int in the reference */
    float   mask_level, mask_level_redun, mapQ_coef_len;
} bwa_b200_region_opt_t;
typedef struct {
    int64_t  rb, re;
    uint64_t hash;
    int32_t  qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0, n_comp, is_alt;
    float    frac_rep;
    int32_t  mapq;
} bwa_b200_alnreg_t;
void bwa_b200_region_opt_default(bwa_b200_region_opt_t *o);   /* mem_opt_init's values (src/bwamem.c:100-140) */
/* regs[region_off[r] .. region_off[r+1]) are read r's regions on entry; on return the first n_regs_out[r] of that slice are its
 * finished regions in the reference's final order, n_pri[r] = mem_mark_primary_se's return value.  The reads are 4-bit packed
 * (bwa_b200_pack_*), the reference is the one attached to idx; ctg_alt (n_ctg flags, may be NULL) = bntann1_t.is_alt.
 * BWA_B200_ERR_ARG when a region lies outside its read ([qb, qe) within read_len), outside the text ([rb, re) within 2*l_pac) or names
 * a contig beyond n_ctg: checked on the host before any device work. */
int  bwa_b200_finish_regions_host(const bwa_b200_index_t *idx, int32_t n_ctg, const int32_t *ctg_alt,
                                  const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len, uint64_t n_reads,
                                  const uint64_t *region_off, bwa_b200_alnreg_t *regs, uint32_t *n_regs_out, int32_t *n_pri,
                                  int64_t first_read_id, const bwa_b200_region_opt_t *opt);

#ifdef __cplusplus
}
#endif
#endif /* BWAMEM_B200_H */
