/*
 * oracle/fork_ksw_shim.c -- TEST INFRASTRUCTURE ONLY.
 * Entry point around the fork's ksw_extend2 (src/ksw.c:864, the variant with the extra
 * opt_ext argument that switches the band off).  Compiled by oracle/build_ref.sh together
 * with /root/reference/src/ksw.c into oracle/_ref/libforkksw.so.  No reference code here.
 */
#include <stdint.h>
#include "ksw.h"

int fork_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                     int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                     int opt_ext, int32_t out[6])
{
    int qle, tle, gtle, gscore, max_off;
    int sc = ksw_extend2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                         &qle, &tle, &gtle, &gscore, &max_off, opt_ext);
    out[0] = sc; out[1] = qle; out[2] = tle; out[3] = gtle; out[4] = gscore; out[5] = max_off;
    return sc;
}

/* the fork's ksw_align2 (src/ksw.c:698-736: ksw_u8 / ksw_i16 as compiled here, i.e. the SSE2 kernels) with qry = NULL and avx2 = 0;
 * it reverses its inputs in place and restores them, so they are copied first.  out7 = score te qe score2 te2 tb qb */
#include <stdlib.h>
#include <string.h>
void fork_ksw_align2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                     int o_del, int e_del, int o_ins, int e_ins, int xtra, int32_t out7[7])
{
    uint8_t *q = (uint8_t *)malloc((size_t)qlen + 16), *t = (uint8_t *)malloc((size_t)tlen + 16);
    memcpy(q, query, (size_t)qlen); memcpy(t, target, (size_t)tlen);
    kswr_t r = ksw_align2(qlen, q, tlen, t, m, mat, o_del, e_del, o_ins, e_ins, xtra, 0, 0);
    out7[0] = r.score; out7[1] = r.te; out7[2] = r.qe; out7[3] = r.score2; out7[4] = r.te2; out7[5] = r.tb; out7[6] = r.qb;
    free(q); free(t);
}
void fork_ksw_align2_batch(int64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen, const uint8_t *tseq, const uint32_t *toff,
                           const uint32_t *tlen, const uint32_t *xtra, int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int32_t *out7)
{
    for (int64_t a = 0; a < n; ++a)
        fork_ksw_align2((int)qlen[a], qseq + qoff[a], (int)tlen[a], tseq + toff[a], m, mat, o_del, e_del, o_ins, e_ins, (int)xtra[a], out7 + 7 * a);
}
