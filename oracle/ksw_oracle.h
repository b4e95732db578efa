/*
 * oracle/ksw_oracle.h -- TEST INFRASTRUCTURE ONLY (CPU restatement of ksw_extend2,
 * src/ksw.c:864-986 == bwa_index/ksw.c:380-479 with opt_ext = 1).
 * Parity pinned against the reference function itself (oracle/_ref) by
 * tests/test_oracle_vs_ref.py and tests/golden/ksw_*.npz.
 */
#ifndef KSW_ORACLE_H
#define KSW_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t score, qle, tle, gtle, gscore, max_off;
} ksw_ext_result_t;

typedef struct {
    uint64_t cells;   /* sum over rows of (end - beg) actually evaluated */
    uint64_t rows;
    uint64_t rect;    /* qlen * tlen */
} ksw_counters_t;

typedef struct {
    int8_t  mat[25];
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t w, end_bonus, zdrop;
    int32_t use_band;          /* the fork's opt_ext argument (src/ksw.c:902-907) */
    int32_t pen_clip;          /* local-vs-to-end rule, src/bwamem.c:1892-1901    */
} ksw_params_t;

void ksw_fill_mat(int a, int b, int8_t mat[25]);   /* bwa_fill_scmat, src/bwa.c */

int ksw_extend2_oracle(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                       const ksw_params_t *p, int h0, ksw_ext_result_t *res, ksw_counters_t *cnt);

/* batch over jobs laid out like the GASAL host batch (byte codes 0..4, per-job offset+len) */
void ksw_extend_batch_oracle(int64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen,
                             const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen,
                             const uint32_t *h0, const ksw_params_t *p,
                             ksw_ext_result_t *res, int n_threads, ksw_counters_t *cnt);
#ifdef __cplusplus
}
#endif
#endif
