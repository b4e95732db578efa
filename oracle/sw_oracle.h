/*
 * oracle/sw_oracle.h -- TEST INFRASTRUCTURE ONLY.  CPU restatement of ksw_align2 (src/ksw.c:698-736) and the two striped
 * Smith-Waterman kernels behind it, ksw_u8 (src/ksw.c:440-572) and ksw_i16 (src/ksw.c:574-696), as the fork compiles them without
 * AVX2: mate rescue (mem_matesw, src/bwamem_pair.c:119-175) and mem_seed_sw (src/bwamem.c:774-830) call it.
 * Pinned against the reference's own SSE2 functions (oracle/_ref/libforkksw.so) by tests/test_sw_oracle.py and
 * tests/golden/sw_golden.npz.
 */
#ifndef SW_ORACLE_H
#define SW_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SW_XBYTE  0x10000
#define SW_XSTOP  0x20000
#define SW_XSUBO  0x40000
#define SW_XSTART 0x80000

typedef struct { int32_t score, te, qe, score2, te2, tb, qb; } sw_result_t;    /* kswr_t, src/ksw.h:14-19 */

/* one striped kernel run (size 1 = ksw_u8, 2 = ksw_i16) */
sw_result_t sw_striped_oracle(int size, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                              int o_del, int e_del, int o_ins, int e_ins, int xtra);
/* ksw_align2 with qry == NULL and avx2 == 0 */
sw_result_t sw_align2_oracle(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                             int o_del, int e_del, int o_ins, int e_ins, int xtra);
void sw_align2_batch_oracle(int64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen,
                            const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen, const uint32_t *xtra,
                            int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, sw_result_t *res, int n_threads);
#ifdef __cplusplus
}
#endif
#endif
