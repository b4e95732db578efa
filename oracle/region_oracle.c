/*
 * oracle/region_oracle.c -- TEST INFRASTRUCTURE ONLY.  See region_oracle.h for what this restates and how it is pinned.
 * Floating-point comparisons keep the operand types of the reference (float options against 64-bit or 32-bit integers, float
 * constants against doubles): the outcomes at the thresholds depend on them.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "region_oracle.h"
#include "global_oracle.h"

void region_opt_default(region_opt_t *o)
{ /* src/bwamem.c:100-140 */
    o->a = 1; o->b = 4; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1; o->w = 100; o->min_seed_len = 19; o->max_chain_gap = 10000;
    o->mask_level = 0.50f; o->mask_level_redun = 0.95f; o->mapQ_coef_len = 50; o->mapQ_coef_fac = (int)log(o->mapQ_coef_len);
}

static void fill_mat(const region_opt_t *o, int8_t mat[25])
{ /* bwa_fill_scmat, src/bwa.c:38-50 */
    int i, j, k;
    for (i = k = 0; i < 4; ++i) {
        for (j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? o->a : -o->b);
        mat[k++] = -1;
    }
    for (j = 0; j < 5; ++j) mat[k++] = -1;
}

/* ---- ks_introsort over region_t (src/ksort.h:146-226): an unstable sort, so equal keys come out in the order this exact
 * sequence of comparisons and swaps leaves them in ---- */
typedef int (*lt_fn)(const region_t *, const region_t *);
static int lt_end(const region_t *a, const region_t *b) { return a->re < b->re; }                                   /* src/bwamem.c:565 */
static int lt_score(const region_t *a, const region_t *b)                                                             /* :568 */
{ return a->score > b->score || (a->score == b->score && (a->rb < b->rb || (a->rb == b->rb && a->qb < b->qb))); }
static int lt_hash(const region_t *a, const region_t *b)                                                              /* :571 */
{ return a->score > b->score || (a->score == b->score && (a->is_alt < b->is_alt || (a->is_alt == b->is_alt && a->hash < b->hash))); }
static int lt_hash2(const region_t *a, const region_t *b)                                                             /* :574 */
{ return a->is_alt < b->is_alt || (a->is_alt == b->is_alt && (a->score > b->score || (a->score == b->score && a->hash < b->hash))); }

#define SWAP(x, y) do { region_t tmp_ = (x); (x) = (y); (y) = tmp_; } while (0)
static void insertsort(lt_fn lt, region_t *s, region_t *t)
{
    for (region_t *i = s + 1; i < t; ++i)
        for (region_t *j = i; j > s && lt(j, j - 1); --j) SWAP(*j, *(j - 1));
}
static void combsort(lt_fn lt, size_t n, region_t *a)
{
    const double shrink = 1.2473309501039786540366528676643;
    int swapped;
    size_t gap = n;
    do {
        if (gap > 2) { gap = (size_t)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
        swapped = 0;
        for (region_t *i = a; i < a + n - gap; ++i)
            if (lt(i + gap, i)) { SWAP(*i, *(i + gap)); swapped = 1; }
    } while (swapped || gap > 2);
    if (gap != 1) insertsort(lt, a, a + n);
}
static void introsort(lt_fn lt, size_t n, region_t *a)
{
    typedef struct { region_t *left, *right; int depth; } frame_t;
    int d;
    if (n < 1) return;
    if (n == 2) { if (lt(&a[1], &a[0])) SWAP(a[0], a[1]); return; }
    for (d = 2; 1ul << d < n; ++d) {}
    frame_t *stack = (frame_t *)malloc(sizeof(frame_t) * (sizeof(size_t) * d + 2)), *top = stack;
    region_t *s = a, *t = a + (n - 1), *i, *j, *k, pivot;
    d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { combsort(lt, (size_t)(t - s + 1), s); t = s; continue; }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (lt(k, i)) { if (lt(k, j)) k = j; }
            else k = lt(j, i) ? i : j;
            pivot = *k;
            if (k != t) SWAP(*k, *t);
            for (;;) {
                do ++i; while (lt(i, &pivot));
                do --j; while (i <= j && lt(&pivot, j));
                if (j <= i) break;
                SWAP(*i, *j);
            }
            SWAP(*i, *t);
            if (i - s > t - i) {
                if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) { free(stack); insertsort(lt, a, a + n); return; }
            --top; s = top->left; t = top->right; d = top->depth;
        }
    }
}

/* the comb-sort fallback of the introsort runs only when quicksort degenerates (depth 2 log2 n), which random inputs never reach:
 * exported so that the tests can drive it directly.  which: 0 end, 1 score, 2 hash, 3 hash2 */
void region_combsort(int which, int n, region_t *a)
{
    static const lt_fn f[4] = {lt_end, lt_score, lt_hash, lt_hash2};
    if (n > 0) combsort(f[which & 3], (size_t)n, a);
}
void region_introsort(int which, int n, region_t *a)
{
    static const lt_fn f[4] = {lt_end, lt_score, lt_hash, lt_hash2};
    introsort(f[which & 3], (size_t)n, a);
}

/* ---- mem_patch_reg, src/bwamem.c:580-618: can hit a (upstream) be joined with hit b by one global alignment? ---- */
#define PATCH_MAX_R_BW 0.05f
#define PATCH_MIN_SC_RATIO 0.90f
static int patch_reg(const region_opt_t *o, const int8_t *mat, int64_t l_pac, const uint8_t *fwd, const uint8_t *query,
                     const region_t *a, const region_t *b, int *w_out)
{
    int w, score = 0, nm, q_s, r_s;
    double r;
    if (fwd == 0 || query == 0) return 0;
    if (a->rb < l_pac && b->rb >= l_pac) return 0;                       /* on different strands */
    if (a->qb >= b->qb || a->qe >= b->qe || a->re >= b->re) return 0;    /* not colinear */
    w = (int)((a->re - b->rb) - (a->qe - b->qb));                        /* required bandwidth */
    w = w > 0 ? w : -w;
    r = (double)(a->re - b->rb) / (b->re - a->rb) - (double)(a->qe - b->qb) / (b->qe - a->qb);   /* relative bandwidth */
    r = r > 0. ? r : -r;
    if (a->re < b->rb || a->qe < b->qb) {                                /* no overlap on query or on ref */
        if (w > o->w << 1 || r >= PATCH_MAX_R_BW) return 0;
    } else if (w > o->w << 2 || r >= PATCH_MAX_R_BW * 2) return 0;       /* more permissive if overlapping on both */
    w += a->w + b->w;
    w = w < o->w << 2 ? w : o->w << 2;
    {   /* bwa_gen_cigar2 for its score only (n_cigar = NM = NULL there, src/bwa.c:111-216).  When it rejects the job it leaves
         * the caller's `score` uninitialised in the reference; here that reads as 0 = no merge. */
        uint32_t cig[8];
        glb_gen_cigar2(mat, o->o_del, o->e_del, o->o_ins, o->e_ins, w, l_pac, fwd, b->qe - a->qb, query + a->qb, a->rb, b->re, &score, &nm, cig, 8);
    }
    q_s = (int)((double)(b->qe - a->qb) / ((b->qe - b->qb) + (a->qe - a->qb)) * (b->score + a->score) + .499);   /* predicted from the query */
    r_s = (int)((double)(b->re - a->rb) / ((b->re - b->rb) + (a->re - a->rb)) * (b->score + a->score) + .499);   /* predicted from the ref */
    if ((double)score / (q_s > r_s ? q_s : r_s) < PATCH_MIN_SC_RATIO) return 0;
    *w_out = w;
    return score;
}

int region_sort_dedup_patch(const region_opt_t *o, int64_t l_pac, const uint8_t *fwd, const uint8_t *query, int n, region_t *a)
{
    int m, i, j;
    int8_t mat[25];
    if (n <= 1) return n;
    fill_mat(o, mat);
    introsort(lt_end, (size_t)n, a);                                     /* by END position */
    for (i = 0; i < n; ++i) a[i].n_comp = 1;
    for (i = 1; i < n; ++i) {
        region_t *p = &a[i];
        if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + o->max_chain_gap) continue;
        for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + o->max_chain_gap; --j) {
            region_t *q = &a[j];
            int64_t pr, pq, mr, mq;
            int score, w;
            if (q->qe == q->qb) continue;                                /* excluded earlier */
            pr = q->re - p->rb;                                          /* overlap on the reference */
            pq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;          /* overlap on the query */
            mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
            mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
            if (pr > o->mask_level_redun * mr && pq > o->mask_level_redun * mq) {   /* one of the two is redundant */
                if (p->score < q->score) { p->qe = p->qb; break; }
                else q->qe = q->qb;
            } else if (q->rb < p->rb && (score = patch_reg(o, mat, l_pac, fwd, query, q, p, &w)) > 0) {   /* merge q into p */
                p->n_comp += q->n_comp + 1;
                p->seedcov = p->seedcov > q->seedcov ? p->seedcov : q->seedcov;
                p->sub = p->sub > q->sub ? p->sub : q->sub;
                p->csub = p->csub > q->csub ? p->csub : q->csub;
                p->qb = q->qb; p->rb = q->rb;
                p->truesc = p->score = score;
                p->w = w;
                q->qb = q->qe;
            }
        }
    }
    for (i = 0, m = 0; i < n; ++i)                                       /* drop the excluded */
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    n = m;
    introsort(lt_score, (size_t)n, a);
    for (i = 1; i < n; ++i)                                              /* identical hits */
        if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
    for (i = 1, m = 1; i < n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    return m;
}

static uint64_t hash_64(uint64_t key)
{ /* src/utils.h:126-137 */
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}

/* src/bwamem.c:685-713; z = indexes of the hits that are primary so far */
static void mark_primary_core(const region_opt_t *o, int n, region_t *a, int *z)
{
    int i, k, nz = 0, tmp;
    tmp = o->a + o->b;
    tmp = o->o_del + o->e_del > tmp ? o->o_del + o->e_del : tmp;
    tmp = o->o_ins + o->e_ins > tmp ? o->o_ins + o->e_ins : tmp;
    z[nz++] = 0;
    for (i = 1; i < n; ++i) {
        for (k = 0; k < nz; ++k) {
            const int j = z[k];
            const int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            const int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {                                         /* overlap on the query */
                const int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
                if (e_min - b_max >= min_l * o->mask_level) {            /* significant */
                    if (a[j].sub == 0) a[j].sub = a[i].score;
                    if (a[j].score - a[i].score <= tmp && (a[j].is_alt || !a[i].is_alt)) ++a[j].sub_n;
                    break;
                }
            }
        }
        if (k == nz) z[nz++] = i;
        else a[i].secondary = z[k];
    }
}

int region_mark_primary_se(const region_opt_t *o, int n, region_t *a, int64_t id)
{
    int i, n_pri;
    if (n == 0) return 0;
    int *z = (int *)malloc(sizeof(int) * (size_t)n);
    for (i = n_pri = 0; i < n; ++i) {
        a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = hash_64((uint64_t)(id + i));
        if (!a[i].is_alt) ++n_pri;
    }
    introsort(lt_hash, (size_t)n, a);
    mark_primary_core(o, n, a, z);
    for (i = 0; i < n; ++i) {
        region_t *p = &a[i];
        p->secondary_all = i;                                            /* rank of the first round */
        if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt) p->alt_sc = a[p->secondary].score;
    }
    if (n_pri >= 0 && n_pri < n) {
        if (n_pri > 0) introsort(lt_hash2, (size_t)n, a);
        for (i = 0; i < n; ++i) z[a[i].secondary_all] = i;
        for (i = 0; i < n; ++i) {
            if (a[i].secondary >= 0) {
                a[i].secondary_all = z[a[i].secondary];
                if (a[i].is_alt) a[i].secondary = 0x7fffffff;
            } else a[i].secondary_all = -1;
        }
        if (n_pri > 0) {                                                 /* primary marking among the primary-assembly hits only */
            for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
            mark_primary_core(o, n_pri, a, z);
        }
    } else {
        for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
    }
    free(z);
    return n_pri;
}

#define MEM_MAPQ_COEF 30.0
int region_approx_mapq_se(const region_opt_t *o, const region_t *a)
{
    int mapq, l, sub = a->sub ? a->sub : o->min_seed_len * o->a;
    double identity;
    sub = a->csub > sub ? a->csub : sub;
    if (sub >= a->score) return 0;
    l = a->qe - a->qb > a->re - a->rb ? a->qe - a->qb : (int)(a->re - a->rb);
    identity = 1. - (double)(l * o->a - a->score) / (o->a + o->b) / l;
    if (a->score == 0) mapq = 0;
    else if (o->mapQ_coef_len > 0) {
        double tmp = l < o->mapQ_coef_len ? 1. : o->mapQ_coef_fac / log(l);
        tmp *= identity * identity;
        mapq = (int)(6.02 * (a->score - sub) / o->a * tmp * tmp + .499);
    } else {
        mapq = (int)(MEM_MAPQ_COEF * (1. - (double)sub / a->score) * log(a->seedcov) + .499);
        mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
    }
    if (a->sub_n > 0) mapq -= (int)(4.343 * log(a->sub_n + 1) + .499);
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int)(mapq * (1. - a->frac_rep) + .499);
    return mapq;
}

int region_finish_read(const region_opt_t *o, int64_t l_pac, const int32_t *ctg_alt, const uint8_t *fwd, const uint8_t *query,
                       int n, region_t *a, int64_t id, int *n_pri)
{
    int i;
    n = region_sort_dedup_patch(o, l_pac, fwd, query, n, a);
    for (i = 0; i < n; ++i)
        if (a[i].rid >= 0 && ctg_alt && ctg_alt[a[i].rid]) a[i].is_alt = 1;
    *n_pri = region_mark_primary_se(o, n, a, id);
    for (i = 0; i < n; ++i) a[i].mapq = a[i].secondary < 0 ? region_approx_mapq_se(o, &a[i]) : 0;
    return n;
}
