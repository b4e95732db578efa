/*
 * oracle/fmd_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's
 * FMD-index seeding path).  Nothing in the product library links or calls this;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may use it.
 *
 * Parity status: PINNED against the reference's own CPU functions compiled from
 * /root/reference/bwa_index (oracle/_ref, see oracle/Makefile + tests/test_oracle_vs_ref.py)
 * and against golden vectors generated from them (tests/golden/).  The reference ships
 * no golden vectors of its own for this path.
 *
 * Index layout consumed here is the reference's *GPU* file layout
 * (bwa_index/bwtindex.c:174-197, seed_gen.cu:28,42-48): one 32-byte bucket per
 * 64 BWT symbols = u32 cnt[A,C,G,T] followed by 4 x u32 of 2-bit symbols.
 */
#ifndef FMD_ORACLE_H
#define FMD_ORACLE_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint64_t primary;      /* row of '$' (bwa_index/bwt.h:48)                    */
    uint64_t L2[5];        /* cumulative base counts                             */
    uint64_t seq_len;      /* 2 * l_pac                                          */
    uint64_t n_words;      /* u32 words in bwt[] (file payload)                  */
    uint32_t *bwt;         /* buckets: 8 words each (+ trailing 4 count words)   */
    int      sa_intv;
    uint64_t n_sa;
    uint32_t *sa;          /* low 32 bits of sampled SA                          */
    uint32_t *sa_hi;       /* packed high bits, pack_size bits per sample        */
    int      pack_size;
    int      owns;         /* arrays malloc'ed by the loader                     */
} fmd_index_t;

typedef struct { uint64_t k, l, s; int32_t beg, end; } fmd_intv_t; /* x[0],x[1],x[2],info */

typedef struct {           /* instrumentation: roofline numerators */
    uint64_t n_extend;     /* bwt_extend calls                                  */
    uint64_t n_bucket;     /* distinct 32B buckets touched by occ lookups       */
    uint64_t n_lf;         /* LF steps inside SA lookups                        */
    uint64_t n_located;    /* SA lookups                                        */
    uint64_t n_smem;       /* SMEMs kept (len >= min_seed_len)                  */
    uint64_t n_extend_fwd; /* subset of n_extend done by forward phases         */
    uint64_t n_bucket_fwd; /* subset of n_bucket touched by forward phases      */
} fmd_counters_t;

/* loaders for the reference file formats (bwa_index/bwt.c:461-487 writers,
 * seed_gen.cu:1386-1468 readers) */
int  fmd_load(fmd_index_t *idx, const char *bwt_path, const char *sa_path);
void fmd_free(fmd_index_t *idx);

/* src/bwt.c:340-403 */
void fmd_occ4(const fmd_index_t *idx, uint64_t k, uint64_t cnt[4], fmd_counters_t *c);
uint64_t fmd_occ(const fmd_index_t *idx, uint64_t k, int base, fmd_counters_t *c);
/* src/bwt.c:455-470 */
void fmd_extend(const fmd_index_t *idx, const fmd_intv_t *ik, fmd_intv_t ok[4], int is_back,
                fmd_counters_t *c);
/* src/bwt.c:483-566 (bwt_smem1a with max_intv = 0); returns next x. out must hold len+1. */
int  fmd_smem1(const fmd_index_t *idx, int len, const uint8_t *q, int x, int min_intv,
               fmd_intv_t *out, int *n_out, fmd_counters_t *c);
/* bwa_index/bwt.c:151-172 (packed-SA variant of src/bwt.c:105-115) */
uint64_t fmd_sa(const fmd_index_t *idx, uint64_t k, fmd_counters_t *c);

/* bwa_index/bwamem.c:121-131 pass 1 only: SMEMs with length >= min_seed_len in the
 * order bwt_smem1 emits them (ascending start).  out must hold len+1 entries. */
int  fmd_collect_pass1(const fmd_index_t *idx, int len, const uint8_t *q, int min_seed_len,
                       fmd_intv_t *out, fmd_counters_t *c);

/* re-seeding parameters of mem_collect_intv (mem_opt_t split_factor 1.5, split_width 10, max_mem_intv 20,
 * bwa_index/bwamem.c:60-62); enable == 0 (or a NULL pointer) = pass 1 only, which is all the reference GPU path does */
typedef struct { int32_t enable; float split_factor; int32_t split_width, max_mem_intv; } fmd_reseed_t;
/* bwa_index/bwt.c:434-455 */
int  fmd_seed_strategy1(const fmd_index_t *idx, int len, const uint8_t *q, int x, int min_len, int max_intv,
                        fmd_intv_t *m, fmd_counters_t *c);
/* bwa_index/bwamem.c:114-162: pass 1, then (rs->enable) passes 2 and 3 and the sort by (start, end).
 * *out is a growable malloc'ed array (capacity *out_cap entries). */
int  fmd_collect_intv(const fmd_index_t *idx, int len, const uint8_t *q, int min_seed_len, const fmd_reseed_t *rs,
                      fmd_intv_t **out, size_t *out_cap, fmd_counters_t *c);
int64_t fmd_seed_batch_rs(const fmd_index_t *idx, const uint8_t *reads, const uint64_t *read_off,
                          int64_t n_reads, int min_seed_len, int max_occ, const fmd_reseed_t *rs,
                          uint32_t *n_seeds, uint64_t *seed_off,
                          uint64_t *rbeg, int32_t *qbeg, int32_t *qend, uint32_t *score,
                          int64_t cap, int n_threads, fmd_counters_t *c);
int64_t fmd_smem_batch_rs(const fmd_index_t *idx, const uint8_t *reads, const uint64_t *read_off,
                          int64_t n_reads, int min_seed_len, const fmd_reseed_t *rs,
                          uint32_t *n_smems, int32_t *qbeg, int32_t *qend, uint64_t *k, uint64_t *s,
                          int64_t cap, int n_threads, fmd_counters_t *c);

/* Whole-batch driver mirroring the result layout of seed_gpu() (seed_gen.h:68-75) but
 * locating only the rows mem_chain consumes (bwa_index/bwamem.c:278-283 sampling rule):
 *   per read r:  n_seeds[r]; seeds [off[r], off[r]+n_seeds[r]) : rbeg, qbeg, qend, and
 *   score = SMEM occurrence count on the first seed of each SMEM group, 0 elsewhere.
 * reads: codes 0..4, concatenated; read_off[n_reads+1].
 * Returns total seeds, or -1 if cap exceeded. */
int64_t fmd_seed_batch(const fmd_index_t *idx, const uint8_t *reads, const uint64_t *read_off,
                       int64_t n_reads, int min_seed_len, int max_occ,
                       uint32_t *n_seeds, uint64_t *seed_off,
                       uint64_t *rbeg, int32_t *qbeg, int32_t *qend, uint32_t *score,
                       int64_t cap, int n_threads, fmd_counters_t *c);

/* SMEM-only batch (no locate): per read list of (qbeg,qend,k,s). */
int64_t fmd_smem_batch(const fmd_index_t *idx, const uint8_t *reads, const uint64_t *read_off,
                       int64_t n_reads, int min_seed_len,
                       uint32_t *n_smems, int32_t *qbeg, int32_t *qend, uint64_t *k, uint64_t *s,
                       int64_t cap, int n_threads, fmd_counters_t *c);

#ifdef __cplusplus
}
#endif
#endif
