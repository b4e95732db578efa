/*
 * include/compat/seed_gen.h -- source-compatible replacement for the GPUSeed boundary of the
 * reference (src/GPUSeed/seed_gen.h:21-106).  The `gase_aln` driver (src/fastmap.c:432-465,
 * src/bwamem.c:404-431) compiles against this header unchanged and links libbwamem_b200.so
 * instead of GPUSeed: same type names, same field names, same seven extern "C" entry points,
 * same ownership rules (results are malloc()ed and freed by the caller with free()).
 *
 * Behind these names sit the B200 kernels of bwa-mem_gpu_b200/csrc/seed.cu; the results follow
 * the CPU definition (bwt_smem1 + bwt_sa), which is what mem_chain expects.
 *
 * Differences a maintainer should know about:
 *  - is_smem == 0 (the -g "MEM" mode) is not provided; seed_gpu() reports it and exits.
 *  - bwt_t_gpu returned by gpu_cpy_wrapper() carries device pointers exactly as before; the
 *    library keeps its own handle keyed on bwt_gpu.bwt.
 */
#ifndef B200_COMPAT_SEED_GEN_H
#define B200_COMPAT_SEED_GEN_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#if defined(__CUDACC__) || defined(__VECTOR_TYPES_H__)
#include <vector_types.h>
#elif defined(__has_include)
#  if __has_include(<vector_types.h>)
#    include <vector_types.h>
#  else
typedef struct { int x, y; } int2;
typedef struct { unsigned int x, y; } uint2;
#  endif
#else
typedef struct { int x, y; } int2;
typedef struct { unsigned int x, y; } uint2;
#endif

typedef uint64_t bwtint_t_gpu;

/* FM index as the driver sees it: host copy after bwt_restore_*_gpu, device copy after
 * gpu_cpy_wrapper (seed_gen.h:21-33) */
typedef struct {
    bwtint_t_gpu  primary;
    bwtint_t_gpu *L2;
    bwtint_t_gpu  seq_len;
    bwtint_t_gpu  bwt_size;
    uint32_t     *bwt;
    int           sa_intv;
    bwtint_t_gpu  n_sa;
    uint32_t     *sa;
    uint32_t     *sa_upper_bits;
    uint8_t       pack_size;
} bwt_t_gpu;

/* Host-side types the fork's driver takes from this header although GPUSeed itself never touches them
 * (seed_gen.h:35-66): the per-read seed record / vector used by mem_chain (src/bwamem.c:323-431) and the
 * bntseq mirror structs. */
typedef struct {
    int64_t  offset;
    int32_t  len;
    int32_t  n_ambs;
    uint32_t gi;
    int32_t  is_alt;
    char    *name, *anno;
} bntann2_t;

typedef struct {
    int64_t offset;
    int32_t len;
    char    amb;
} bntamb2_t;

typedef struct {
    int64_t    l_pac;
    int32_t    n_seqs;
    uint32_t   seed;
    bntann2_t *anns;
    int32_t    n_holes;
    bntamb2_t *ambs;
    FILE      *fp_pac;
} bntseq2_t;

typedef struct {
    int64_t rbeg;
    int32_t qbeg, len;
    int     score;
} mem_seed_t;

typedef struct { size_t n, m; mem_seed_t *a; int seed_counter; } mem_seed_v;

/* flat per-file seed table consumed by mem_chain (seed_gen.h:68-75, src/bwamem.c:415-431) */
typedef struct {
    bwtint_t_gpu *rbeg;
    int2         *qbeg;                           /* x = qbeg, y = qend */
    uint32_t     *score;                          /* SMEM occurrence count on the first seed of a group */
    uint32_t     *n_ref_pos_fow_rev_results;      /* seeds per read */
    uint32_t     *n_ref_pos_fow_rev_prefix_sums;  /* exclusive prefix sums over the file */
    uint64_t      file_bytes_skip;
} mem_seed_v_gpu;

typedef struct {
    char      *read_file;
    char      *query_file;
    bwt_t_gpu *bwt;
    bwt_t_gpu  bwt_gpu;
    uint2     *pre_calc_seed_intervals;
    int        pre_calc_seed_intervals_flag;
    int        pre_calc_seed_len;
    int        min_seed_size;
    int        is_smem;
    uint64_t   file_bytes_skip;
} gpuseed_storage_vector;

#ifdef __cplusplus
extern "C" {
#endif

void            bwt_destroy_gpu(bwt_t_gpu *bwt);
void            bwt_restore_sa_gpu(const char *fn, bwt_t_gpu *bwt);
bwt_t_gpu      *bwt_restore_bwt_gpu(const char *fn);
bwt_t_gpu       gpu_cpy_wrapper(bwt_t_gpu *bwt);
void            pre_calc_seed_intervals_wrapper(uint2 *pre_calc_seed_intervals, int pre_calc_seed_len, bwt_t_gpu bwt_gpu);
void            free_gpuseed_data(gpuseed_storage_vector *gpuseed_data);
mem_seed_v_gpu *seed_gpu(gpuseed_storage_vector *gpuseed_data);

/* B200 additions (optional): device ordinal for the next gpu_cpy_wrapper/seed_gpu on this thread,
 * and the occurrence cap; max_occ <= 0 (default) keeps the reference layout where every SMEM
 * contributes all of its occurrences. */
void            gpuseed_b200_set_device(int device);
void            gpuseed_b200_set_max_occ(int max_occ);
/* re-seeding for the next seed_gpu on this thread: enable != 0 adds passes 2 and 3 of mem_collect_intv (bwa_index/bwamem.c:132-161), the
 * seed set of stock bwa mem, lifting the reference's documented limitation (README.md:93); split_factor 1.5, split_width 10,
 * max_mem_intv 20 are mem_opt_init's.  Default: off, i.e. exactly what GPUSeed returns.  The layout of mem_seed_v_gpu is unchanged. */
void            gpuseed_b200_set_reseed(int enable, float split_factor, int split_width, int max_mem_intv);

#ifdef __cplusplus
}
#endif
#endif
