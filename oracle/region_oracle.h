/*
 * oracle/region_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of what the reference does to a read's alignment regions
 * between the extension results and SAM: mem_sort_dedup_patch -> is_alt -> mem_mark_primary_se -> mem_approx_mapq_se).
 * Nothing in the product library links or calls this.
 *
 * Parity status: PINNED against the reference fork's own functions compiled from /root/reference/src into
 * oracle/_ref/libforkmem.so (fork_finish_regs in oracle/fork_mem_shim.cpp; tests/test_region_oracle.py) and against golden
 * vectors generated from them (tests/golden/region_golden.npz).  The reference ships no vectors of its own.
 */
#ifndef REGION_ORACLE_H
#define REGION_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {                 /* the mem_opt_t fields these functions read (src/bwamem.h:34-73, defaults src/bwamem.c:100-140) */
    int32_t a, b, o_del, e_del, o_ins, e_ins, w, min_seed_len, max_chain_gap, mapQ_coef_fac;   /* mapQ_coef_fac is an int there */
    float mask_level, mask_level_redun, mapQ_coef_len;
} region_opt_t;

typedef struct {                 /* the mem_alnreg_t fields they read or write (src/bwamem.h:83-112) */
    int64_t rb, re;
    uint64_t hash;
    int32_t qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0, n_comp, is_alt;
    float frac_rep;
    int32_t mapq;                /* mem_reg2aln's: mem_approx_mapq_se when secondary < 0, else 0 (src/bwamem.c:2363) */
} region_t;

void region_opt_default(region_opt_t *o);
/* src/bwamem.c:620-681; fwd = forward reference, one code per base; query = codes 0..4 of the whole read.  Returns the new count. */
int region_sort_dedup_patch(const region_opt_t *o, int64_t l_pac, const uint8_t *fwd, const uint8_t *query, int n, region_t *a);
/* src/bwamem.c:715-760; returns n_pri */
int region_mark_primary_se(const region_opt_t *o, int n, region_t *a, int64_t id);
/* src/bwamem.c:1690-1716 */
int region_approx_mapq_se(const region_opt_t *o, const region_t *a);
/* the stage for one read, in the reference's order (src/bwamem.c:2313-2326, 2459, 2363): dedup + patch, is_alt from the contig table
 * (ctg_alt may be NULL), primary marking with id = index of the read in the run, mapq.  Returns the new count; *n_pri as above. */
int region_finish_read(const region_opt_t *o, int64_t l_pac, const int32_t *ctg_alt, const uint8_t *fwd, const uint8_t *query,
                       int n, region_t *a, int64_t id, int *n_pri);
/* the sorts alone (src/ksort.h:146-226 with the comparators of src/bwamem.c:565-575): which = 0 end, 1 score, 2 hash, 3 hash2 */
void region_combsort(int which, int n, region_t *a);
void region_introsort(int which, int n, region_t *a);
#ifdef __cplusplus
}
#endif
#endif
