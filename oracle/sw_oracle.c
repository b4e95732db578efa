/*
 * oracle/sw_oracle.c -- TEST INFRASTRUCTURE ONLY (see sw_oracle.h).
 *
 * The reference's kernels are Farrar-striped SSE2 code: a vector holds p = 16 (u8) or 8 (i16) query positions slen apart,
 * position of (vector j, lane l) = l * slen + j, slen = ceil(qlen / p) (ksw_qinit, src/ksw.c:387-436).  Their results are NOT those
 * of the textbook recurrence, so the restatement keeps what makes them differ, in scalar code over positions:
 *   - main pass (src/ksw.c:484-507 / :615-631): positions of a lane are visited in order with F restarted at 0 at the first position
 *     of every lane; E(i+1, .) is computed from this pass's H, before any F carried across lanes has been applied;
 *   - lazy-F loop (src/ksw.c:509-522 / :632-644): the F left at the end of every lane moves to the next lane and decays through it,
 *     raising H only; the loop is vector-wide -- up to 16 rounds (in this fork for both element sizes), each up to slen vector steps,
 *     left at the first step after which no lane's F exceeds H - oe_ins -- and is replayed here step for step;
 *   - the row maximum is taken in the main pass only; saturation: unsigned bytes with the matrix biased by `shift` (u8), signed
 *     16-bit adds and unsigned saturating subtractions (i16);
 *   - the bookkeeping of ksw_u8 / ksw_i16: rows at or above the XSUBO threshold collected in runs, Hmax, the early stop of XSTOP and
 *     of a byte overflow, qe by the smallest position among equal maxima, score2 / te2 outside the window around te.
 */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "sw_oracle.h"

static inline int subs(int a, int b) { return a > b ? a - b : 0; }                      /* _mm_subs_epu8 / _mm_subs_epu16 on values >= 0 */
static inline int adds16(int a, int b) { int v = a + b; return v > 32767 ? 32767 : (v < -32768 ? -32768 : v); }
static inline int imax(int a, int b) { return a > b ? a : b; }

sw_result_t sw_striped_oracle(int size, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                              int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
    sw_result_t r = {0, -1, -1, -1, -1, -1, -1};                                        /* g_defr, src/ksw.c:50 */
    const int p = size == 1 ? 16 : 8, slen = (qlen + p - 1) / p, n = slen * p;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int shift = 127, mdiff = 0;                                                          /* src/ksw.c:404-412 */
    for (int a = 0; a < m * m; ++a) { if (mat[a] < shift) shift = mat[a]; if (mat[a] > mdiff) mdiff = mat[a]; }
    const int qmax = mdiff;
    shift = (256 - shift) & 255;
    const int minsc = (xtra & SW_XSUBO) ? xtra & 0xffff : 0x10000, endsc = (xtra & SW_XSTOP) ? xtra & 0xffff : 0x10000;
    if (slen == 0) return r;            /* the reference would read H0[-1]; not a defined case */
    int *const base = (int *)calloc((size_t)n * 4, sizeof(int));
    int *H0 = base, *H1 = H0 + n, *E = H1 + n, *Hmax = E + n;
    uint64_t *b = NULL;
    int n_b = 0, m_b = 0, gmax = 0, te = -1;
    int fv[16];
    for (int i = 0; i < tlen; ++i) {
        const int8_t *ma = mat + (int)target[i] * m;
        int rowmax = 0;
        /* main pass: lane l = positions l*slen .. l*slen + slen - 1, F from 0 */
        for (int l = 0; l < p; ++l) {
            int f = 0;
            for (int j = 0; j < slen; ++j) {
                const int pos = l * slen + j;
                int h = pos > 0 ? H0[pos - 1] : 0;                                       /* H(i-1, pos-1): the shifted last vector for j = 0 */
                const int sc = pos >= qlen ? 0 : ma[query[pos]];
                if (size == 1) { h = h + sc + shift; h = h > 255 ? 255 : h; h = subs(h, shift); }
                else h = adds16(h, sc);
                int e = E[pos];
                h = imax(h, e); h = imax(h, f);
                rowmax = imax(rowmax, h);
                H1[pos] = h;
                e = imax(subs(e, e_del), subs(h, oe_del));
                E[pos] = e;
                f = imax(subs(f, e_ins), subs(h, oe_ins));
            }
            fv[l] = f;
        }
        /* lazy-F loop, vector-wide */
        for (int k = 0, done = 0; k < 16 && !done; ++k) {
            for (int l = p - 1; l > 0; --l) fv[l] = fv[l - 1];
            fv[0] = 0;
            for (int j = 0; j < slen; ++j) {
                int any = 0;
                for (int l = 0; l < p; ++l) {
                    const int pos = l * slen + j;
                    int h = imax(H1[pos], fv[l]);
                    H1[pos] = h;
                    h = subs(h, oe_ins);
                    fv[l] = subs(fv[l], e_ins);
                    if (fv[l] > h) any = 1;
                }
                if (!any) { done = 1; break; }
            }
        }
        if (rowmax >= minsc) {                                                           /* src/ksw.c:526-537 */
            if (n_b == 0 || (int32_t)b[n_b - 1] + 1 != i) {
                if (n_b == m_b) { m_b = m_b ? m_b << 1 : 8; b = (uint64_t *)realloc(b, 8 * (size_t)m_b); }
                b[n_b++] = (uint64_t)rowmax << 32 | (uint32_t)i;
            } else if ((int)(b[n_b - 1] >> 32) < rowmax) b[n_b - 1] = (uint64_t)rowmax << 32 | (uint32_t)i;
        }
        if (rowmax > gmax) {
            gmax = rowmax; te = i;
            memcpy(Hmax, H1, sizeof(int) * (size_t)n);
            if ((size == 1 && gmax + shift >= 255) || gmax >= endsc) break;
        }
        int *t = H1; H1 = H0; H0 = t;
    }
    r.score = size == 1 ? (gmax + shift < 255 ? gmax : 255) : gmax;
    r.te = te;
    if (size != 1 || r.score != 255) {
        int mx = -1;
        if (size != 1) r.qe = -1;
        /* the reference walks Hmax in memory order (vector j, lane l); qe = smallest position among the maxima */
        for (int i2 = 0; i2 < n; ++i2) {
            const int pos = i2 / p + i2 % p * slen, v = Hmax[pos];
            if (v > mx) { mx = v; r.qe = pos; }
            else if (v == mx && pos < r.qe) r.qe = pos;
        }
        if (b) {
            int w = (r.score + qmax - 1) / qmax, low = te - w, high = te + w;
            for (int i2 = 0; i2 < n_b; ++i2) {
                const int e = (int32_t)b[i2];
                if ((e < low || e > high) && (int)(b[i2] >> 32) > r.score2) { r.score2 = (int)(b[i2] >> 32); r.te2 = e; }
            }
        }
    }
    free(b);
    free(base);
    return r;
}

sw_result_t sw_align2_oracle(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat,
                             int o_del, int e_del, int o_ins, int e_ins, int xtra)
{
    const int size = (xtra & SW_XBYTE) ? 1 : 2;
    sw_result_t r = sw_striped_oracle(size, qlen, query, tlen, target, m, mat, o_del, e_del, o_ins, e_ins, xtra);
    if ((xtra & SW_XSTART) == 0 || ((xtra & SW_XSUBO) && r.score < (xtra & 0xffff))) return r;
    if (r.qe < 0 || r.te < 0) return r;          /* byte overflow (score 255, qe unset): the reference's second pass is undefined */
    /* second pass over the reversed prefixes; the target keeps its tail behind the reversed part (src/ksw.c:722-733) */
    uint8_t *rq = (uint8_t *)malloc((size_t)r.qe + 1), *rt = (uint8_t *)malloc((size_t)tlen + 1);
    for (int i = 0; i <= r.qe; ++i) rq[i] = query[r.qe - i];
    memcpy(rt, target, (size_t)tlen);
    for (int i = 0; i <= r.te; ++i) rt[i] = target[r.te - i];
    sw_result_t rr = sw_striped_oracle(size, r.qe + 1, rq, tlen, rt, m, mat, o_del, e_del, o_ins, e_ins, SW_XSTOP | r.score);
    free(rq); free(rt);
    if (r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
    return r;
}

void sw_align2_batch_oracle(int64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen,
                            const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen, const uint32_t *xtra,
                            int m, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, sw_result_t *res, int n_threads)
{
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 16)
    for (int64_t a = 0; a < n; ++a)
        res[a] = sw_align2_oracle((int)qlen[a], qseq + qoff[a], (int)tlen[a], tseq + toff[a], m, mat, o_del, e_del, o_ins, e_ins, (int)xtra[a]);
}
