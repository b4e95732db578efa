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
