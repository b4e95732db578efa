/* oracle/global_oracle.c -- TEST INFRASTRUCTURE ONLY; see global_oracle.h. */
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "global_oracle.h"

typedef struct { int32_t h, e; } eh_t;

static int push_op(int n, uint32_t *cigar, int cap, uint32_t *last, int op, int len)
{ /* push_cigar, src/ksw.c: merge with the previous operation when it has the same code.  *last mirrors cigar[n-1] so that
     operations beyond cap are still merged and counted correctly */
    if (n == 0 || op != (int)(*last & 0xf)) {
        *last = (uint32_t)len << 4 | (uint32_t)op;
        if (n < cap) cigar[n] = *last;
        return n + 1;
    }
    *last += (uint32_t)len << 4;
    if (n - 1 < cap) cigar[n - 1] = *last;
    return n;
}

int glb_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar_, uint32_t *cigar, int cap, uint64_t *cells)
{
    const int oe_del = o_del + e_del, oe_ins = o_ins + e_ins;
    int i, j, score;
    const int n_col = qlen < 2 * w + 1 ? qlen : 2 * w + 1;
    uint8_t *z = (uint8_t *)malloc((size_t)(n_col > 0 ? n_col : 1) * (size_t)(tlen > 0 ? tlen : 1));
    eh_t *eh = (eh_t *)calloc((size_t)qlen + 1, sizeof(eh_t));
    eh[0].h = 0; eh[0].e = GLB_MINUS_INF;
    for (j = 1; j <= qlen && j <= w; ++j) { eh[j].h = -(o_ins + e_ins * j); eh[j].e = GLB_MINUS_INF; }
    for (; j <= qlen; ++j) eh[j].h = eh[j].e = GLB_MINUS_INF;
    for (i = 0; i < tlen; ++i) {
        int32_t f = GLB_MINUS_INF, h1, t;
        const int8_t *srow = mat + target[i] * 5;
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
        uint8_t *zi = z + (size_t)i * n_col;
        h1 = beg == 0 ? -(o_del + e_del * (i + 1)) : GLB_MINUS_INF;
        if (cells && end > beg) *cells += (uint64_t)(end - beg);
        for (j = beg; j < end; ++j) {
            eh_t *p = &eh[j];
            int32_t h, m = p->h, e = p->e;
            uint8_t d;
            p->h = h1;
            m += srow[query[j]];
            d = m >= e ? 0 : 1;
            h = m >= e ? m : e;
            d = h >= f ? d : 2;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del;
            e -= e_del;
            d |= e > t ? 1 << 2 : 0;
            e = e > t ? e : t;
            p->e = e;
            t = m - oe_ins;
            f -= e_ins;
            d |= f > t ? 2 << 4 : 0;
            f = f > t ? f : t;
            zi[j - beg] = d;
        }
        eh[end].h = h1; eh[end].e = GLB_MINUS_INF;
    }
    score = eh[qlen].h;
    {
        int n = 0, which = 0, k;
        uint32_t last = 0;
        i = tlen - 1;
        k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        while (i >= 0 && k >= 0) {
            which = z[(size_t)i * n_col + (k - (i > w ? i - w : 0))] >> (which << 1) & 3;
            if (which == 0) { n = push_op(n, cigar, cap, &last, 0, 1); --i; --k; }
            else if (which == 1) { n = push_op(n, cigar, cap, &last, 2, 1); --i; }
            else { n = push_op(n, cigar, cap, &last, 1, 1); --k; }
        }
        if (i >= 0) n = push_op(n, cigar, cap, &last, 2, i + 1);
        if (k >= 0) n = push_op(n, cigar, cap, &last, 1, k + 1);
        if (n <= cap)
            for (i = 0; i < n >> 1; ++i) { uint32_t tmp = cigar[i]; cigar[i] = cigar[n - 1 - i]; cigar[n - 1 - i] = tmp; }
        *n_cigar_ = n;
    }
    free(eh); free(z);
    return score;
}

int glb_band(const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w_, int l_query, int64_t rlen)
{
    int w, max_gap, max_ins, max_del, min_w;
    max_ins = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_ins) / e_ins + 1.);
    max_del = (int)((double)(((l_query + 1) >> 1) * mat[0] - o_del) / e_del + 1.);
    max_gap = max_ins > max_del ? max_ins : max_del;
    max_gap = max_gap > 1 ? max_gap : 1;
    w = (max_gap + abs((int)(rlen - l_query)) + 1) >> 1;
    w = w < w_ ? w : w_;
    min_w = abs((int)(rlen - l_query)) + 3;
    w = w > min_w ? w : min_w;
    return w;
}

int glb_nm(int n_cigar, const uint32_t *cigar, const uint8_t *query, const uint8_t *rseq)
{
    int k, x = 0, y = 0, n_mm = 0, n_gap = 0, i;
    for (k = 0; k < n_cigar; ++k) {
        const int op = cigar[k] & 0xf, len = (int)(cigar[k] >> 4);
        if (op == 0) {
            for (i = 0; i < len; ++i) if (query[x + i] != rseq[y + i]) ++n_mm;
            x += len; y += len;
        } else if (op == 2) {
            if (k > 0 && k < n_cigar - 1) n_gap += len;
            y += len;
        } else if (op == 1) { x += len; n_gap += len; }
    }
    return n_mm + n_gap;
}

void glb_batch(int64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen, const uint8_t *tseq, const uint32_t *toff,
               const uint32_t *tlen, const uint32_t *w, const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins,
               int32_t *score, int32_t *nm, uint32_t *n_cigar, uint32_t *cigar, int cig_stride, int n_threads, uint64_t *cells)
{
    uint64_t tot = 0;
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 256) reduction(+:tot)
    for (int64_t a = 0; a < n; ++a) {
        int nc = 0;
        uint64_t c = 0;
        uint32_t *cg = cigar + (size_t)a * cig_stride;
        score[a] = glb_global2((int)qlen[a], qseq + qoff[a], (int)tlen[a], tseq + toff[a], mat, o_del, e_del, o_ins, e_ins, (int)w[a], &nc, cg, cig_stride, &c);
        n_cigar[a] = (uint32_t)nc;
        nm[a] = nc <= cig_stride ? glb_nm(nc, cg, qseq + qoff[a], tseq + toff[a]) : -1;
        tot += c;
    }
    if (cells) *cells = tot;
}

int glb_gen_cigar2(const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *fwd,
                   int l_query, const uint8_t *query_in, int64_t rb, int64_t re, int *score, int *nm, uint32_t *cigar, int cap)
{
    int n_cigar = 0, i;
    *nm = -1; *score = 0;
    if (l_query <= 0 || rb >= re || (rb < l_pac && re > l_pac)) return -1;
    /* bns_get_seq */
    int64_t beg = rb, end = re, rlen;
    if (end > l_pac << 1) end = l_pac << 1;
    if (beg < 0) beg = 0;
    rlen = end - beg;
    if (re - rb != rlen) return -1;                          /* out of range */
    uint8_t *rseq = (uint8_t *)malloc((size_t)rlen + 1), *query = (uint8_t *)malloc((size_t)l_query);
    memcpy(query, query_in, (size_t)l_query);
    if (beg >= l_pac) {
        int64_t beg_f = (l_pac << 1) - 1 - end, end_f = (l_pac << 1) - 1 - beg, k, l = 0;
        for (k = end_f; k > beg_f; --k) rseq[l++] = (uint8_t)(3 - fwd[k]);
    } else memcpy(rseq, fwd + beg, (size_t)rlen);
    if (rb >= l_pac) {                                       /* reverse both so that indels are placed leftmost */
        for (i = 0; i < l_query >> 1; ++i) { uint8_t t = query[i]; query[i] = query[l_query - 1 - i]; query[l_query - 1 - i] = t; }
        for (i = 0; i < rlen >> 1; ++i) { uint8_t t = rseq[i]; rseq[i] = rseq[rlen - 1 - i]; rseq[rlen - 1 - i] = t; }
    }
    if (l_query == re - rb && w_ == 0) {
        if (cap > 0) cigar[0] = (uint32_t)l_query << 4;
        n_cigar = 1;
        for (i = 0; i < l_query; ++i) *score += mat[rseq[i] * 5 + query[i]];
    } else {
        const int w = glb_band(mat, o_del, e_del, o_ins, e_ins, w_, l_query, rlen);
        *score = glb_global2(l_query, query, (int)rlen, rseq, mat, o_del, e_del, o_ins, e_ins, w, &n_cigar, cigar, cap, 0);
    }
    if (n_cigar <= cap) *nm = glb_nm(n_cigar, cigar, query, rseq);
    free(rseq); free(query);
    return n_cigar;
}

/* ---- mem_reg2aln (src/bwamem.c:2344-2438): band inference, bwa_gen_cigar2 with the band-doubling retry, squeeze of a
 * leading / trailing deletion, soft clips, position.  The MD string, mapq and the flags stay out (text / other inputs). */
static int infer_bw(int l1, int l2, int score, int a, int q, int r)
{ /* src/bwamem.c:1486-1494 */
    int w;
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
    if (w < abs(l1 - l2)) w = abs(l1 - l2);
    return w;
}

static int pos2rid(int n_ctg, const int64_t *ctg_off, int64_t l_pac, int64_t pos_f)
{ /* bns_pos2rid, src/bntseq.c:349-363 */
    int left = 0, mid = 0, right = n_ctg;
    if (pos_f >= l_pac) return -1;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= ctg_off[mid]) {
            if (mid == n_ctg - 1) break;
            if (pos_f < ctg_off[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}

int glb_reg2aln(const int8_t *mat, int a, int o_del, int e_del, int o_ins, int e_ins, int opt_w, int64_t l_pac, const uint8_t *fwd,
                int n_ctg, const int64_t *ctg_off, int l_query, const uint8_t *query, int qb, int qe, int64_t rb, int64_t re,
                int truesc, int ar_w, glb_aln_t *out, uint32_t *cigar, int cap)
{
    int i, w2, tmp, NM = -1, score = 0, last_sc = -(1 << 30), n_cigar = 0, is_rev, n_waves = 0;
    int64_t pos;
    memset(out, 0, sizeof(*out));
    if (rb < 0 || re < 0) { out->rid = -1; out->pos = -1; return 0; }
    tmp = infer_bw(qe - qb, (int)(re - rb), truesc, a, o_del, e_del);
    w2 = infer_bw(qe - qb, (int)(re - rb), truesc, a, o_ins, e_ins);
    w2 = w2 > tmp ? w2 : tmp;
    if (w2 > opt_w) w2 = w2 < ar_w ? w2 : ar_w;
    i = 0;
    do {
        w2 = w2 < opt_w << 2 ? w2 : opt_w << 2;
        int sc = 0, nm = -1;
        int n = glb_gen_cigar2(mat, o_del, e_del, o_ins, e_ins, w2, l_pac, fwd, qe - qb, query + qb, rb, re, &sc, &nm, cigar, cap);
        ++n_waves;
        if (n >= 0) { n_cigar = n; score = sc; NM = nm; } else { n_cigar = 0; NM = -1; }   /* a rejected job leaves score as it was */
        if (score == last_sc || w2 == opt_w << 2) break;
        last_sc = score;
        w2 <<= 1;
    } while (++i < 3 && score < truesc - a);
    if (n_cigar > cap) return -1;
    pos = rb < l_pac ? rb : re - 1;
    is_rev = pos >= l_pac;
    if (is_rev) pos = (l_pac << 1) - 1 - pos;
    if (n_cigar > 0) {
        if ((cigar[0] & 0xf) == 2) { pos += cigar[0] >> 4; --n_cigar; memmove(cigar, cigar + 1, (size_t)n_cigar * 4); }
        else if ((cigar[n_cigar - 1] & 0xf) == 2) --n_cigar;
    }
    if (qb != 0 || qe != l_query) {
        const int clip5 = is_rev ? l_query - qe : qb, clip3 = is_rev ? qb : l_query - qe;
        if (n_cigar + 2 > cap) return -1;
        if (clip5) { memmove(cigar + 1, cigar, (size_t)n_cigar * 4); cigar[0] = (uint32_t)clip5 << 4 | 3; ++n_cigar; }
        if (clip3) cigar[n_cigar++] = (uint32_t)clip3 << 4 | 3;
    }
    out->rid = pos2rid(n_ctg, ctg_off, l_pac, pos);
    out->pos = pos - (out->rid >= 0 ? ctg_off[out->rid] : 0);
    out->is_rev = is_rev; out->score = score; out->nm = NM; out->n_cigar = n_cigar; out->band = w2; out->n_waves = n_waves;
    return n_cigar;
}
