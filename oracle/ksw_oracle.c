/*
 * oracle/ksw_oracle.c -- TEST INFRASTRUCTURE ONLY.
 * Row-by-row restatement of ksw_extend2 (src/ksw.c:864-986; stock copy
 * bwa_index/ksw.c:380-479).  The reference keeps one array eh[0..qlen] whose entry j
 * holds { H(i-1, j-1), E(i, j) }; entries outside the evaluated window [beg, end) of a
 * row keep whatever an earlier row (or the initial row) left there.  That staleness is
 * observable (the window can grow by two columns per row), so it is reproduced here
 * with the same array semantics rather than re-derived.
 */
#include "ksw_oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void ksw_fill_mat(int a, int b, int8_t mat[25])
{
    int i, j, k = 0;
    for (i = 0; i < 4; ++i) {
        for (j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? a : -b);
        mat[k++] = -1; /* ambiguous base */
    }
    for (j = 0; j < 5; ++j) mat[k++] = -1;
}

int ksw_extend2_oracle(int qlen, const uint8_t *query, int tlen, const uint8_t *target,
                       const ksw_params_t *p, int h0, ksw_ext_result_t *res, ksw_counters_t *cnt)
{
    const int o_del = p->o_del, e_del = p->e_del, o_ins = p->o_ins, e_ins = p->e_ins;
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int *Hd = (int *)calloc((size_t)qlen + 1, sizeof(int)); /* eh[j].h */
    int *E  = (int *)calloc((size_t)qlen + 1, sizeof(int)); /* eh[j].e */
    int i, j, w = p->w;

    /* first row (src/ksw.c:880-883) */
    Hd[0] = h0;
    if (qlen >= 1) Hd[1] = h0 > oe_ins ? h0 - oe_ins : 0;
    for (j = 2; j <= qlen && Hd[j - 1] > e_ins; ++j) Hd[j] = Hd[j - 1] - e_ins;

    /* band clamp (src/ksw.c:885-893) */
    int mx = 0;
    for (i = 0; i < 25; ++i) mx = mx > p->mat[i] ? mx : p->mat[i];
    int max_ins = (int)((double)(qlen * mx + p->end_bonus - o_ins) / e_ins + 1.);
    if (max_ins < 1) max_ins = 1;
    if (w > max_ins) w = max_ins;
    int max_del = (int)((double)(qlen * mx + p->end_bonus - o_del) / e_del + 1.);
    if (max_del < 1) max_del = 1;
    if (w > max_del) w = max_del;

    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    for (i = 0; i < tlen; ++i) {
        int f = 0, h1, rowmax = 0, rowmax_j = -1;
        const int8_t *srow = p->mat + target[i] * 5;
        if (p->use_band) {
            if (beg < i - w) beg = i - w;
            if (end > i + w + 1) end = i + w + 1;
            if (end > qlen) end = qlen;
        }
        if (beg == 0) {
            h1 = h0 - (o_del + e_del * (i + 1));
            if (h1 < 0) h1 = 0;
        } else h1 = 0;
        for (j = beg; j < end; ++j) {
            int M = Hd[j], e = E[j], h, t;
            Hd[j] = h1;
            M = M ? M + srow[query[j]] : 0;
            h = M > e ? M : e;
            h = h > f ? h : f;
            h1 = h;
            rowmax_j = rowmax > h ? rowmax_j : j;
            rowmax = rowmax > h ? rowmax : h;
            t = M - oe_del; t = t > 0 ? t : 0;
            e -= e_del; e = e > t ? e : t;
            E[j] = e;
            t = M - oe_ins; t = t > 0 ? t : 0;
            f -= e_ins; f = f > t ? f : t;
        }
        if (cnt) { cnt->cells += (uint64_t)(end > beg ? end - beg : 0); cnt->rows++; }
        Hd[end] = h1; E[end] = 0;
        if (j == qlen) {
            best_ie = gscore > h1 ? best_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (rowmax == 0) break;
        if (rowmax > best) {
            int d = rowmax_j - i; if (d < 0) d = -d;
            best = rowmax; best_i = i; best_j = rowmax_j;
            if (d > max_off) max_off = d;
        } else if (p->zdrop > 0) {
            int di = i - best_i, dj = rowmax_j - best_j;
            if (di > dj) { if (best - rowmax - (di - dj) * e_del > p->zdrop) break; }
            else         { if (best - rowmax - (dj - di) * e_ins > p->zdrop) break; }
        }
        for (j = beg; j < end && Hd[j] == 0 && E[j] == 0; ++j) ;
        beg = j;
        for (j = end; j >= beg && Hd[j] == 0 && E[j] == 0; --j) ;
        end = j + 2 < qlen ? j + 2 : qlen;
    }
    free(Hd); free(E);
    if (cnt) cnt->rect += (uint64_t)qlen * (uint64_t)tlen;
    res->score = best; res->qle = best_j + 1; res->tle = best_i + 1;
    res->gtle = best_ie + 1; res->gscore = gscore; res->max_off = max_off;
    return best;
}

void ksw_extend_batch_oracle(int64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen,
                             const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen,
                             const uint32_t *h0, const ksw_params_t *p,
                             ksw_ext_result_t *res, int n_threads, ksw_counters_t *cnt)
{
    if (n_threads < 1) n_threads = 1;
    uint64_t cells = 0, rows = 0, rect = 0;
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 256) reduction(+:cells,rows,rect)
    for (int64_t a = 0; a < n; ++a) {
        ksw_counters_t c = {0, 0, 0};
        ksw_extend2_oracle((int)qlen[a], qseq + qoff[a], (int)tlen[a], tseq + toff[a], p, (int)h0[a], &res[a], &c);
        cells += c.cells; rows += c.rows; rect += c.rect;
    }
    if (cnt) { cnt->cells += cells; cnt->rows += rows; cnt->rect += rect; }
}
