/*
 * oracle/global_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of the reference's CIGAR path:
 * bwa_gen_cigar2 -> ksw_global2).  Nothing in the product library links or calls this.
 *
 * Parity status: PINNED against the reference's own ksw_global2 / bwa_gen_cigar2 compiled from
 * /root/reference/bwa_index into oracle/_ref/libbwaref.so (tests/test_global_oracle.py) and against golden vectors
 * generated from them (tests/golden/global_golden.npz).  The reference ships no vectors of its own.
 */
#ifndef GLOBAL_ORACLE_H
#define GLOBAL_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define GLB_MINUS_INF (-0x40000000)

/* src/ksw.c:1120-1241 (= bwa_index/ksw.c:504-606): banded global alignment with backtrack.  cigar (len << 4 | op,
 * op 0 M / 1 I / 2 D) receives at most cap entries; *n_cigar is the full count.  Returns the score. */
int glb_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t *cigar, int cap, uint64_t *cells);
/* the band bwa_gen_cigar2 hands to ksw_global2 (src/bwa.c:161-169) */
int glb_band(const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w_, int l_query, int64_t rlen);
/* NM of bwa_gen_cigar2 (src/bwa.c:178-205): mismatches + inserted + deleted bases, a deletion that is the first or the
 * last CIGAR operation not counted */
int glb_nm(int n_cigar, const uint32_t *cigar, const uint8_t *query, const uint8_t *rseq);
/* batch of independent jobs (byte per base, codes 0..4); per job score, nm, n_cigar and up to cig_stride operations at
 * cigar[a * cig_stride ..].  w[a] is the band given to ksw_global2. */
void glb_batch(int64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen, const uint8_t *tseq, const uint32_t *toff,
               const uint32_t *tlen, const uint32_t *w, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
               int32_t *score, int32_t *nm, uint32_t *n_cigar, uint32_t *cigar, int cig_stride, int n_threads, uint64_t *cells);
/* bwa_gen_cigar2 over forward reference codes (one byte per base) instead of the 2-bit pac: fetch [rb, re) as bns_get_seq
 * does (bwa_index/bntseq.c:404-425), reverse query and reference when rb >= l_pac, band rule, ksw_global2 (or the ungapped
 * shortcut when l_query == re - rb and w_ == 0), NM.  Returns n_cigar, or -1 when the reference rejects the job. */
int glb_gen_cigar2(const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *fwd,
                   int l_query, const uint8_t *query, int64_t rb, int64_t re, int *score, int *nm, uint32_t *cigar, int cap);
/* mem_reg2aln (src/bwamem.c:2344-2438) without the text parts: the fields of mem_aln_t it computes from the DP */
typedef struct { int64_t pos; int32_t rid, is_rev, score, nm, n_cigar, band, n_waves; } glb_aln_t;
/* score = the last global score, band = the last w2 handed to bwa_gen_cigar2, n_waves = calls of bwa_gen_cigar2; the CIGAR has
 * the leading / trailing deletion squeezed out and the soft clips added.  Returns n_cigar, -1 when cap is too small. */
int glb_reg2aln(const int8_t *mat, int a, int o_del, int e_del, int o_ins, int e_ins, int opt_w, int64_t l_pac, const uint8_t *fwd,
                int n_ctg, const int64_t *ctg_off, int l_query, const uint8_t *query, int qb, int qe, int64_t rb, int64_t re,
                int truesc, int ar_w, glb_aln_t *out, uint32_t *cigar, int cap);
#ifdef __cplusplus
}
#endif
#endif
