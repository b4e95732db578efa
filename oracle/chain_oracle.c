/*
 * oracle/chain_oracle.c -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement of the reference fork's seed -> chain -> extension-job stage (SURVEY.md section 8f row 1
 * and section 8a "job construction"), one read at a time, in plain C:
 *
 *   chain_oracle_read      mem_chain            src/bwamem.c:404-476   (seeds come from mem_seed_v_gpu)
 *                          test_and_merge       src/bwamem.c:337-359
 *                          kbtree put/interval  src/kbtree.h:117-131,151-168,174-224 (KB_DEFAULT_SIZE 512 and
 *                                               sizeof(mem_chain_t) = 40 give t = 5: at most 9 chains per node)
 *                          bns_intv2rid         src/bntseq.c:349-373
 *                          mem_chain_weight     src/bwamem.c:361-384
 *                          mem_chain_flt        src/bwamem.c:488-560, ks_introsort src/ksort.h:146-226
 *                          mem_flt_chained_seeds src/bwamem.c:970-990: only its "short read" early return is
 *                                               restated; a read for which it would run mem_seed_sw gives -2
 *   chain2aln_oracle_read  mem_chain2aln        src/bwamem.c:1170-1479 with cal_max_gap :996-1002,
 *                                               bns_fetch_seq src/bntseq.c:531-560 and fill_extension :1102-1167
 *   chain_regs_finish      result gathering     src/bwamem.c:2286-2306
 *
 * The B-tree is restated as a B-tree, not as a sorted list: when two chains of a read start at the same
 * reference position the reference's answer depends on where the equal keys sit in the tree.
 * Pinned against the unmodified reference code (oracle/_ref/libforkmem.so) by tests/test_chain_oracle.py.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "chain_oracle.h"
#include "sw_oracle.h"

void chain_opt_default(chain_opt_t *o)
{ /* mem_opt_init, src/bwamem.c:107-150 (the fork's defaults: w = 300) */
    o->a = 1; o->b = 4; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1; o->w = 300;
    o->min_seed_len = 19; o->max_occ = 500; o->max_chain_gap = 10000; o->min_chain_weight = 0; o->max_chain_extend = 1 << 30;
    o->mask_level = 0.50f; o->drop_ratio = 0.50f;
}

/* ------------------------------------------------------------------ reference coordinates */
static int pos2rid(int64_t l_pac, int n_ctg, const int64_t *off, int64_t pos_f)
{ /* src/bntseq.c:349-363 */
    int left = 0, mid = 0, right = n_ctg;
    if (pos_f >= l_pac) return -1;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= off[mid]) {
            if (mid == n_ctg - 1) break;
            if (pos_f < off[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}
static int64_t depos(int64_t l_pac, int64_t pos, int *is_rev) { return (*is_rev = (pos >= l_pac)) ? (l_pac << 1) - 1 - pos : pos; }
static int intv2rid(int64_t l_pac, int n_ctg, const int64_t *off, int64_t rb, int64_t re)
{ /* src/bntseq.c:365-373 */
    int is_rev, rid_b, rid_e;
    if (rb < l_pac && re > l_pac) return -2;
    rid_b = pos2rid(l_pac, n_ctg, off, depos(l_pac, rb, &is_rev));
    rid_e = rb < re ? pos2rid(l_pac, n_ctg, off, depos(l_pac, re - 1, &is_rev)) : rid_b;
    return rid_b == rid_e ? rid_b : -1;
}

/* ------------------------------------------------------------------ chains under construction */
typedef struct { int64_t rbeg; int32_t qbeg, len, score; } seed_t;
typedef struct { int n, m, first, rid, w, kept, is_alt; int64_t pos; seed_t *seeds; } chn_t;

/* ------------------------------------------------------------------ kbtree with t = 5 */
#define KB_T 5
#define KB_N (2 * KB_T - 1)
typedef struct { int n, is_internal; int key[KB_N]; int ptr[KB_N + 1]; } kbn_t;
typedef struct { kbn_t *nd; int n_nodes, cap, root; const chn_t *ch; } kbt_t;

static int kb_new(kbt_t *b, int internal)
{
    if (b->n_nodes == b->cap) { b->cap = b->cap * 2 + 8; b->nd = (kbn_t *)realloc(b->nd, (size_t)b->cap * sizeof(kbn_t)); }
    kbn_t *x = &b->nd[b->n_nodes];
    memset(x, 0, sizeof(*x));
    x->is_internal = internal;
    return b->n_nodes++;
}
static int kb_cmp(int64_t a, int64_t b) { return (b < a) - (a < b); }
static int kb_getp_aux(const kbt_t *b, const kbn_t *x, int64_t k, int *r)
{ /* src/kbtree.h:117-131: first key >= k; *r = 0 on an exact hit */
    int tr, *rr = r ? r : &tr, begin = 0, end = x->n;
    if (x->n == 0) return -1;
    while (begin < end) {
        int mid = (begin + end) >> 1;
        if (kb_cmp(b->ch[x->key[mid]].pos, k) < 0) begin = mid + 1; else end = mid;
    }
    if (begin == x->n) { *rr = 1; return x->n - 1; }
    if ((*rr = kb_cmp(k, b->ch[x->key[begin]].pos)) < 0) --begin;
    return begin;
}
static int kb_lower(const kbt_t *b, int64_t k)
{ /* kb_intervalp, src/kbtree.h:151-168: the chain `lower` points at, or -1 */
    int x = b->root, lower = -1, r = 0;
    while (x >= 0) {
        const kbn_t *nd = &b->nd[x];
        int i = kb_getp_aux(b, nd, k, &r);
        if (i >= 0 && r == 0) return nd->key[i];
        if (i >= 0) lower = nd->key[i];
        if (!nd->is_internal) return lower;
        x = nd->ptr[i + 1];
    }
    return lower;
}
static void kb_split(kbt_t *b, int xi, int i, int yi)
{ /* src/kbtree.h:176-191 */
    int zi = kb_new(b, b->nd[yi].is_internal);
    kbn_t *x = &b->nd[xi], *y = &b->nd[yi], *z = &b->nd[zi];
    z->n = KB_T - 1;
    memcpy(z->key, y->key + KB_T, sizeof(int) * (KB_T - 1));
    if (y->is_internal) memcpy(z->ptr, y->ptr + KB_T, sizeof(int) * KB_T);
    y->n = KB_T - 1;
    memmove(x->ptr + i + 2, x->ptr + i + 1, sizeof(int) * (size_t)(x->n - i));
    x->ptr[i + 1] = zi;
    memmove(x->key + i + 1, x->key + i, sizeof(int) * (size_t)(x->n - i));
    x->key[i] = y->key[KB_T - 1];
    ++x->n;
}
static void kb_putp_aux(kbt_t *b, int xi, int c)
{ /* src/kbtree.h:192-209 */
    const int64_t k = b->ch[c].pos;
    kbn_t *x = &b->nd[xi];
    if (!x->is_internal) {
        int i = kb_getp_aux(b, x, k, 0);
        if (i != x->n - 1) memmove(x->key + i + 2, x->key + i + 1, (size_t)(x->n - i - 1) * sizeof(int));
        x->key[i + 1] = c;
        ++x->n;
    } else {
        int i = kb_getp_aux(b, x, k, 0) + 1;
        if (b->nd[x->ptr[i]].n == KB_N) {
            kb_split(b, xi, i, x->ptr[i]);
            x = &b->nd[xi];
            if (kb_cmp(k, b->ch[x->key[i]].pos) > 0) ++i;
        }
        kb_putp_aux(b, b->nd[xi].ptr[i], c);
    }
}
static void kb_put(kbt_t *b, int c)
{ /* src/kbtree.h:210-224 */
    if (b->nd[b->root].n == KB_N) {
        int s = kb_new(b, 1), r = b->root;
        b->root = s;
        b->nd[s].ptr[0] = r;
        kb_split(b, s, 0, r);
    }
    kb_putp_aux(b, b->root, c);
}
static void kb_traverse(const kbt_t *b, int xi, int *out, int *n)
{ /* in-order, src/kbtree.h:336-358 */
    const kbn_t *x = &b->nd[xi];
    for (int i = 0; i <= x->n; ++i) {
        if (x->is_internal) kb_traverse(b, x->ptr[i], out, n);
        if (i < x->n) out[(*n)++] = x->key[i];
    }
}

/* ------------------------------------------------------------------ mem_chain */
static int test_and_merge(const chain_opt_t *opt, int64_t l_pac, chn_t *c, const seed_t *p, int seed_rid)
{ /* src/bwamem.c:337-359 */
    const seed_t *last = &c->seeds[c->n - 1];
    int64_t qend = last->qbeg + last->len, rend = last->rbeg + last->len, x, y;
    if (seed_rid != c->rid) return 0;
    if (p->qbeg >= c->seeds[0].qbeg && p->qbeg + p->len <= qend && p->rbeg >= c->seeds[0].rbeg && p->rbeg + p->len <= rend) return 1;
    if ((last->rbeg < l_pac || c->seeds[0].rbeg < l_pac) && p->rbeg >= l_pac) return 0;
    x = p->qbeg - last->qbeg;
    y = p->rbeg - last->rbeg;
    if (y >= 0 && x - y <= opt->w && y - x <= opt->w && x - last->len < opt->max_chain_gap && y - last->len < opt->max_chain_gap) {
        if (c->n == c->m) { c->m <<= 1; c->seeds = (seed_t *)realloc(c->seeds, (size_t)c->m * sizeof(seed_t)); }
        c->seeds[c->n++] = *p;
        return 1;
    }
    return 0;
}

static int chain_weight(const chn_t *c)
{ /* src/bwamem.c:361-384 */
    int64_t end;
    int j, w = 0, tmp;
    for (j = 0, end = 0; j < c->n; ++j) {
        const seed_t *s = &c->seeds[j];
        if (s->qbeg >= end) w += s->len;
        else if (s->qbeg + s->len > end) w += s->qbeg + s->len - end;
        end = end > s->qbeg + s->len ? end : s->qbeg + s->len;
    }
    tmp = w; w = 0;
    for (j = 0, end = 0; j < c->n; ++j) {
        const seed_t *s = &c->seeds[j];
        if (s->rbeg >= end) w += s->len;
        else if (s->rbeg + s->len > end) w += s->rbeg + s->len - end;
        end = end > s->rbeg + s->len ? end : s->rbeg + s->len;
    }
    w = w < tmp ? w : tmp;
    return w < 1 << 30 ? w : (1 << 30) - 1;
}

/* ks_introsort(mem_flt) on chain indices; __sort_lt(a, b) = a.w > b.w  (src/ksort.h:146-226, src/bwamem.c:485-486) */
#define FLT_LT(a, b) (ch[a].w > ch[b].w)
static void flt_insertsort(const chn_t *ch, int *s, int *t)
{
    for (int *i = s + 1; i < t; ++i)
        for (int *j = i; j > s && FLT_LT(*j, *(j - 1)); --j) { int tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}
static void flt_combsort(const chn_t *ch, size_t n, int *a)
{
    const double shrink_factor = 1.2473309501039786540366528676643;
    int do_swap;
    size_t gap = n;
    do {
        if (gap > 2) { gap = (size_t)(gap / shrink_factor); if (gap == 9 || gap == 10) gap = 11; }
        do_swap = 0;
        for (int *i = a; i < a + n - gap; ++i) {
            int *j = i + gap;
            if (FLT_LT(*j, *i)) { int tmp = *i; *i = *j; *j = tmp; do_swap = 1; }
        }
    } while (do_swap || gap > 2);
    if (gap != 1) flt_insertsort(ch, a, a + n);
}
static void flt_introsort(const chn_t *ch, size_t n, int *a)
{
    typedef struct { int *left, *right; int depth; } stk_t;
    int d;
    stk_t *top, *stack;
    int rp, tmp, *s, *t, *i, *j, *k;
    if (n < 1) return;
    else if (n == 2) { if (FLT_LT(a[1], a[0])) { tmp = a[0]; a[0] = a[1]; a[1] = tmp; } return; }
    for (d = 2; 1ul << d < n; ++d);
    stack = (stk_t *)malloc(sizeof(stk_t) * ((sizeof(size_t) * d) + 2));
    top = stack; s = a; t = a + (n - 1); d <<= 1;
    while (1) {
        if (s < t) {
            if (--d == 0) { flt_combsort(ch, (size_t)(t - s + 1), s); t = s; continue; }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (FLT_LT(*k, *i)) { if (FLT_LT(*k, *j)) k = j; }
            else k = FLT_LT(*j, *i) ? i : j;
            rp = *k;
            if (k != t) { tmp = *k; *k = *t; *t = tmp; }
            for (;;) {
                do ++i; while (FLT_LT(*i, rp));
                do --j; while (i <= j && FLT_LT(rp, *j));
                if (j <= i) break;
                tmp = *i; *i = *j; *j = tmp;
            }
            tmp = *i; *i = *t; *t = tmp;
            if (i - s > t - i) {
                if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) { free(stack); flt_insertsort(ch, a, a + n); return; }
            else { --top; s = top->left; t = top->right; d = top->depth; }
        }
    }
}

#define CHN_BEG(c) ((c).seeds[0].qbeg)
#define CHN_END(c) ((c).seeds[(c).n - 1].qbeg + (c).seeds[(c).n - 1].len)

/* seeds of the read: rbeg[i], {qbeg, qend} = qbeg_qend[2i..], score[i] = occurrence count s on the first seed of an
 * SMEM group.  layout_all != 0: a group holds all s rows and is sampled with step s / max_occ (the reference's
 * mem_seed_v_gpu, src/bwamem.c:419-431); layout_all == 0: a group holds only the sampled rows, consecutively
 * (bwa_b200_seeds_t with max_occ > 0).  chains / cseeds need room for n_seeds entries.
 * Returns 0; -2 when mem_flt_chained_seeds would not return early for this read length. */
static int chain_read_impl(const chain_opt_t *opt, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                           int l_query, uint32_t n_seeds, const uint64_t *rbeg, const int32_t *qq, const uint32_t *score, int layout_all,
                           int32_t *n_chains, chain_rec_t *chains, chain_seed_t *cseeds, int stop_at_long);

int chain_oracle_read(const chain_opt_t *opt, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                      int l_query, uint32_t n_seeds, const uint64_t *rbeg, const int32_t *qq, const uint32_t *score, int layout_all,
                      int32_t *n_chains, chain_rec_t *chains, chain_seed_t *cseeds)
{
    return chain_read_impl(opt, l_pac, n_ctg, ctg_off, ctg_len, ctg_alt, l_query, n_seeds, rbeg, qq, score, layout_all, n_chains, chains, cseeds, 1);
}
/* the same for a read of any length: the caller runs chain_oracle_flt_seeds on the result (mem_flt_chained_seeds) */
int chain_oracle_read_any(const chain_opt_t *opt, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                          int l_query, uint32_t n_seeds, const uint64_t *rbeg, const int32_t *qq, const uint32_t *score, int layout_all,
                          int32_t *n_chains, chain_rec_t *chains, chain_seed_t *cseeds)
{
    return chain_read_impl(opt, l_pac, n_ctg, ctg_off, ctg_len, ctg_alt, l_query, n_seeds, rbeg, qq, score, layout_all, n_chains, chains, cseeds, 0);
}

static int chain_read_impl(const chain_opt_t *opt, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                           int l_query, uint32_t n_seeds, const uint64_t *rbeg, const int32_t *qq, const uint32_t *score, int layout_all,
                           int32_t *n_chains, chain_rec_t *chains, chain_seed_t *cseeds, int stop_at_long)
{
    (void)ctg_len;
    *n_chains = 0;
    if (l_query < opt->min_seed_len) return 0;
    if (stop_at_long) {   /* mem_flt_chained_seeds, src/bwamem.c:972-977 (MEM_HSP_COEF 1.1, MEM_MINSC_COEF 5.5, MEM_SEEDSW_COEF 0.05) */
        double min_l = opt->min_chain_weight ? 1.1f * opt->min_chain_weight : 5.5f * log(l_query);
        if (!(min_l > 0.05f * l_query)) return -2;
    }
    chn_t *ch = (chn_t *)calloc(n_seeds ? n_seeds : 1, sizeof(chn_t));
    int n_ch = 0;
    kbt_t bt;
    memset(&bt, 0, sizeof(bt));
    bt.ch = ch;
    bt.root = kb_new(&bt, 0);
    int tree_size = 0;

    /* fraction of the read covered by repetitive seeds, src/bwamem.c:415-422 */
    int b = 0, e = 0, l_rep = 0;
    uint32_t i;
#define GROUP(i_, s_, step_, cnt_, grp_)                                                                     \
    uint32_t s_ = score[i_], step_ = s_ > (uint32_t)opt->max_occ ? s_ / (uint32_t)opt->max_occ : 1u;          \
    uint32_t cnt_ = (s_ + step_ - 1) / step_;                                                                \
    if (cnt_ > (uint32_t)opt->max_occ) cnt_ = (uint32_t)opt->max_occ;                                        \
    uint32_t grp_ = layout_all ? s_ : cnt_;                                                                  \
    if (grp_ == 0) grp_ = 1;
    for (i = 0; i < n_seeds;) {
        GROUP(i, s, step, cnt, grp)
        int sb = qq[2 * i], se = qq[2 * i + 1];
        (void)step; (void)cnt;
        if (s > (uint32_t)opt->max_occ) {
            if (sb > e) { l_rep += e - b; b = sb; e = se; }
            else e = e > se ? e : se;
        }
        i += grp;
    }
    l_rep += e - b;

    for (i = 0; i < n_seeds;) {
        GROUP(i, s, step, cnt, grp)
        int slen = qq[2 * i + 1] - qq[2 * i];
        uint32_t k, count;
        (void)cnt;
        for (k = count = 0; k < s && count < (uint32_t)opt->max_occ; k += step, ++count) {
            seed_t sd;
            sd.rbeg = (int64_t)rbeg[layout_all ? i + k : i + count];
            sd.qbeg = qq[2 * i];
            sd.score = sd.len = slen;
            int rid = intv2rid(l_pac, n_ctg, ctg_off, sd.rbeg, sd.rbeg + sd.len), to_add = 0;
            if (rid < 0) continue;
            if (tree_size) {
                int lower = kb_lower(&bt, sd.rbeg);
                if (lower < 0 || !test_and_merge(opt, l_pac, &ch[lower], &sd, rid)) to_add = 1;
            } else to_add = 1;
            if (to_add) {
                chn_t *c = &ch[n_ch];
                c->n = 1; c->m = 4;
                c->seeds = (seed_t *)calloc((size_t)c->m, sizeof(seed_t));
                c->seeds[0] = sd;
                c->rid = rid; c->is_alt = ctg_alt ? !!ctg_alt[rid] : 0; c->pos = sd.rbeg;
                kb_put(&bt, n_ch);
                ++n_ch; ++tree_size;
            }
        }
        i += grp;
    }
    int *a = (int *)malloc(sizeof(int) * (size_t)(n_ch ? n_ch : 1)), n_chn = 0;
    kb_traverse(&bt, bt.root, a, &n_chn);
    const float frac_rep = (float)l_rep / l_query;

    /* ---- mem_chain_flt, src/bwamem.c:488-560 */
    int k, n_kept = 0;
    if (n_chn > 0) {
        int kk = 0;
        for (int ii = 0; ii < n_chn; ++ii) {
            chn_t *c = &ch[a[ii]];
            c->first = -1; c->kept = 0;
            c->w = chain_weight(c);
            if (c->w < opt->min_chain_weight) continue;
            a[kk++] = a[ii];
        }
        n_chn = kk;
        flt_introsort(ch, (size_t)n_chn, a);
        if (n_chn > 0) {
            int *kept_idx = (int *)malloc(sizeof(int) * (size_t)n_chn), n_k = 0;
            ch[a[0]].kept = 3;
            kept_idx[n_k++] = 0;
            for (int ii = 1; ii < n_chn; ++ii) {
                int large_ovlp = 0;
                chn_t *ci = &ch[a[ii]];
                for (k = 0; k < n_k; ++k) {
                    int j = kept_idx[k];
                    chn_t *cj = &ch[a[j]];
                    int b_max = CHN_BEG(*cj) > CHN_BEG(*ci) ? CHN_BEG(*cj) : CHN_BEG(*ci);
                    int e_min = CHN_END(*cj) < CHN_END(*ci) ? CHN_END(*cj) : CHN_END(*ci);
                    if (e_min > b_max && (!cj->is_alt || ci->is_alt)) {
                        int li = CHN_END(*ci) - CHN_BEG(*ci);
                        int lj = CHN_END(*cj) - CHN_BEG(*cj);
                        int min_l = li < lj ? li : lj;
                        if (e_min - b_max >= min_l * opt->mask_level && min_l < opt->max_chain_gap) {
                            large_ovlp = 1;
                            if (cj->first < 0) cj->first = ii;
                            if (ci->w < cj->w * opt->drop_ratio && cj->w - ci->w >= opt->min_seed_len << 1) break;
                        }
                    }
                }
                if (k == n_k) { kept_idx[n_k++] = ii; ci->kept = large_ovlp ? 2 : 3; }
            }
            for (int ii = 0; ii < n_k; ++ii) {
                chn_t *c = &ch[a[kept_idx[ii]]];
                if (c->first >= 0) ch[a[c->first]].kept = 1;
            }
            free(kept_idx);
            int ii;
            for (ii = k = 0; ii < n_chn; ++ii) {
                if (ch[a[ii]].kept == 0 || ch[a[ii]].kept == 3) continue;
                if (++k >= opt->max_chain_extend) break;
            }
            for (; ii < n_chn; ++ii) if (ch[a[ii]].kept < 3) ch[a[ii]].kept = 0;
            for (ii = 0; ii < n_chn; ++ii) if (ch[a[ii]].kept != 0) a[n_kept++] = a[ii];
        }
    }
    int so = 0;
    for (int ii = 0; ii < n_kept; ++ii) {
        const chn_t *c = &ch[a[ii]];
        chain_rec_t *o = &chains[ii];
        o->pos = c->pos; o->rid = c->rid; o->n = c->n; o->w = c->w; o->kept = c->kept; o->first = c->first; o->is_alt = c->is_alt;
        o->frac_rep = frac_rep; o->seed_off = so;
        for (int j = 0; j < c->n; ++j, ++so) {
            cseeds[so].rbeg = c->seeds[j].rbeg; cseeds[so].qbeg = c->seeds[j].qbeg; cseeds[so].len = c->seeds[j].len;
            cseeds[so].score = c->seeds[j].score; cseeds[so].pad = 0;
        }
    }
    *n_chains = n_kept;
    for (int ii = 0; ii < n_ch; ++ii) free(ch[ii].seeds);
    free(ch); free(a); free(bt.nd);
    return 0;
}

/* ------------------------------------------------------------------ mem_chain2aln */
static int cal_max_gap(const chain_opt_t *opt, int qlen)
{ /* src/bwamem.c:996-1002 */
    int l_del = (int)((double)(qlen * opt->a - opt->o_del) / opt->e_del + 1.);
    int l_ins = (int)((double)(qlen * opt->a - opt->o_ins) / opt->e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < opt->w << 1 ? l : opt->w << 1;
}
static uint8_t text_base(const uint8_t *fwd, int64_t l_pac, int64_t p)
{ /* base p of fwd + revcomp(fwd): bns_get_seq, src/bntseq.c:506-529 */
    return p < l_pac ? fwd[p] : (uint8_t)(3 - fwd[(l_pac << 1) - 1 - p]);
}
/* mem_seed_sw, src/bwamem.c:774-808: a local alignment of the read around the seed against the reference around it, 50 bases either
 * side; -1 when the seed or its windows reach MEM_SHORT_LEN (200) */
static int seed_sw(const chain_opt_t *opt, const int8_t *mat, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len,
                   const uint8_t *fwd, int l_query, const uint8_t *query, const chain_seed_t *s)
{
    int qb, qe, is_rev, rid;
    int64_t rb, re, mid, far_beg, far_end;
    uint8_t rseq[256];
    if (s->len >= 200) return -1;
    qb = s->qbeg; qe = s->qbeg + s->len;
    rb = s->rbeg; re = s->rbeg + s->len;
    mid = (rb + re) >> 1;
    qb -= 50; qb = qb > 0 ? qb : 0;
    qe += 50; qe = qe < l_query ? qe : l_query;
    rb -= 50; rb = rb > 0 ? rb : 0;
    re += 50; re = re < l_pac << 1 ? re : l_pac << 1;
    if (rb < l_pac && l_pac < re) { if (mid < l_pac) re = l_pac; else rb = l_pac; }
    if (qe - qb >= 200 || re - rb >= 200) return -1;
    /* bns_fetch_seq, src/bntseq.c:531-552: clamp to the contig that holds mid */
    rid = pos2rid(l_pac, n_ctg, ctg_off, depos(l_pac, mid, &is_rev));
    far_beg = ctg_off[rid]; far_end = far_beg + ctg_len[rid];
    if (is_rev) { int64_t tmp = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - tmp; }
    rb = rb > far_beg ? rb : far_beg;
    re = re < far_end ? re : far_end;
    for (int64_t k = rb; k < re; ++k) rseq[k - rb] = text_base(fwd, l_pac, k);
    sw_result_t x = sw_align2_oracle(qe - qb, query + qb, (int)(re - rb), rseq, 5, mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, SW_XSTART);
    return x.score;
}

/* mem_flt_chained_seeds, src/bwamem.c:970-990, on the chains chain_oracle_read_any left: seeds whose local alignment scores below
 * min_HSP_score leave their chain (chains[i].n shrinks; the kept seeds of all chains are packed and seed_off re-assigned), the others take the
 * alignment's score (seed length * a when mem_seed_sw declined).  Returns 1 when the read was long enough for the filter to run. */
int chain_oracle_flt_seeds(const chain_opt_t *opt, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len,
                           const uint8_t *fwd, int l_query, const uint8_t *query, int n_chains, chain_rec_t *chains, chain_seed_t *cseeds)
{
    double min_l = opt->min_chain_weight ? 1.1f * opt->min_chain_weight : 5.5f * log(l_query);
    int min_HSP_score = (int)(opt->a * min_l + .499);
    int8_t mat[25];
    int i, j, k;
    if (min_l > 0.05f * l_query) return 0;
    for (i = k = 0; i < 4; ++i) { for (j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? opt->a : -opt->b); mat[k++] = -1; }   /* bwa_fill_scmat */
    for (j = 0; j < 5; ++j) mat[k++] = -1;
    int so = 0;                                   /* the kept seeds are packed: chain after chain, seed_off re-assigned */
    for (i = 0; i < n_chains; ++i) {
        const chain_seed_t *sd = cseeds + chains[i].seed_off;
        const int first = so;
        for (j = 0; j < chains[i].n; ++j) {
            chain_seed_t s = sd[j];
            s.score = seed_sw(opt, mat, l_pac, n_ctg, ctg_off, ctg_len, fwd, l_query, query, &s);
            if (s.score < 0 || s.score >= min_HSP_score) {
                s.score = s.score < 0 ? s.len * opt->a : s.score;
                cseeds[so++] = s;
            }
        }
        chains[i].n = so - first;
        chains[i].seed_off = first;
    }
    (void)k;
    return 1;
}

static int cmp_u64(const void *x, const void *y) { uint64_t a = *(const uint64_t *)x, b = *(const uint64_t *)y; return a < b ? -1 : a > b; }

typedef struct { chain_job_t *jobs; uint8_t *q, *t; uint32_t nq, nt; int n; } side_t;

static int fill_ext(side_t *sd, int cap_jobs, uint64_t cap_bytes, const uint8_t *ref, const uint8_t *read, int ref_len, int read_len, int h0)
{ /* fill_extension, src/bwamem.c:1102-1167: sequences are appended and padded to a multiple of 8 with N (code 4) */
    if (sd->n >= cap_jobs) return -1;
    uint32_t tpad = (8 - (uint32_t)ref_len % 8) % 8, qpad = (8 - (uint32_t)read_len % 8) % 8;
    if ((uint64_t)sd->nt + ref_len + tpad > cap_bytes || (uint64_t)sd->nq + read_len + qpad > cap_bytes) return -1;
    chain_job_t *j = &sd->jobs[sd->n++];
    j->toff = sd->nt; j->qoff = sd->nq; j->tlen = (uint32_t)ref_len; j->qlen = (uint32_t)read_len; j->h0 = (uint32_t)h0;
    memcpy(sd->t + sd->nt, ref, (size_t)ref_len); memset(sd->t + sd->nt + ref_len, 4, tpad); sd->nt += (uint32_t)ref_len + tpad;
    memcpy(sd->q + sd->nq, read, (size_t)read_len); memset(sd->q + sd->nq + read_len, 4, qpad); sd->nq += (uint32_t)read_len + qpad;
    return 0;
}

/* chains / cseeds as produced by chain_oracle_read.  fwd: forward reference, one code 0..3 per base.
 * Jobs go to side 0 (SHORT) or 1 (LONG); qbuf[s] / tbuf[s] receive their sequences (cap_bytes each).
 * Returns 0, -1 when a capacity is too small. */
int chain2aln_oracle_read(const chain_opt_t *opt, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len,
                          const uint8_t *fwd, int l_query, const uint8_t *query,
                          int n_chains, const chain_rec_t *chains, const chain_seed_t *cseeds,
                          int32_t *n_regs_out, chain_reg_t *regs, int cap_regs,
                          int32_t n_jobs[2], chain_job_t *jobs_short, chain_job_t *jobs_long, int cap_jobs,
                          uint8_t *qbuf[2], uint8_t *tbuf[2], uint64_t cap_bytes)
{
    side_t side[2] = {{jobs_short, qbuf[0], tbuf[0], 0, 0, 0}, {jobs_long, qbuf[1], tbuf[1], 0, 0, 0}};
    int n_regs = 0, rc = 0;
    for (int ci = 0; ci < n_chains && rc == 0; ++ci) {
        const chain_rec_t *c = &chains[ci];
        const chain_seed_t *sd = cseeds + c->seed_off;
        if (c->n == 0) continue;
        int64_t rmax[2];
        int i, k;
        rmax[0] = l_pac << 1; rmax[1] = 0;
        for (i = 0; i < c->n; ++i) {   /* src/bwamem.c:1180-1201 */
            const chain_seed_t *t = &sd[i];
            int64_t b = t->rbeg - (t->qbeg + cal_max_gap(opt, t->qbeg));
            int64_t e = t->rbeg + t->len + ((l_query - t->qbeg - t->len) + cal_max_gap(opt, l_query - t->qbeg - t->len));
            rmax[0] = rmax[0] < b ? rmax[0] : b;
            rmax[1] = rmax[1] > e ? rmax[1] : e;
        }
        rmax[0] = rmax[0] > 0 ? rmax[0] : 0;
        rmax[1] = rmax[1] < l_pac << 1 ? rmax[1] : l_pac << 1;
        if (rmax[0] < l_pac && l_pac < rmax[1]) { if (sd[0].rbeg < l_pac) rmax[1] = l_pac; else rmax[0] = l_pac; }
        {   /* bns_fetch_seq clamps the window to the contig of the first seed, src/bntseq.c:531-552 */
            int is_rev;
            int rid = pos2rid(l_pac, n_ctg, ctg_off, depos(l_pac, sd[0].rbeg, &is_rev));
            int64_t far_beg = ctg_off[rid], far_end = far_beg + ctg_len[rid];
            if (is_rev) { int64_t tmp = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - tmp; }
            rmax[0] = rmax[0] > far_beg ? rmax[0] : far_beg;
            rmax[1] = rmax[1] < far_end ? rmax[1] : far_end;
        }
        const int64_t l_refer = rmax[1] - rmax[0];
        uint8_t *rseq = (uint8_t *)malloc((size_t)(l_refer > 0 ? l_refer : 1));
        for (int64_t p = 0; p < l_refer; ++p) rseq[p] = text_base(fwd, l_pac, rmax[0] + p);
        uint64_t *srt = (uint64_t *)malloc((size_t)c->n * 8);
        for (i = 0; i < c->n; ++i) srt[i] = (uint64_t)sd[i].score << 32 | (uint32_t)i;
        qsort(srt, (size_t)c->n, 8, cmp_u64);      /* ks_introsort_64: the keys are distinct, any sort gives this order */

        for (k = c->n - 1; k >= 0 && rc == 0; --k) {
            const chain_seed_t *s = &sd[(uint32_t)srt[k]];
            for (i = 0; i < n_regs; ++i) {          /* src/bwamem.c:1225-1244: estimated extents of earlier regions */
                const chain_reg_t *p = &regs[i];
                int64_t rd;
                int qd, w, max_gap;
                if (s->rbeg < p->rb_est || s->rbeg + s->len > p->re_est || s->qbeg < p->qb_est || s->qbeg + s->len > p->qe_est) continue;
                if (s->len - p->seedlen0 > .1 * l_query) continue;
                qd = s->qbeg - p->qb_est; rd = s->rbeg - p->rb_est;
                max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
                w = max_gap < p->w ? max_gap : p->w;
                if (qd - rd < w && rd - qd < w) break;
                qd = p->qe_est - (s->qbeg + s->len); rd = p->re_est - (s->rbeg + s->len);
                max_gap = cal_max_gap(opt, qd < rd ? qd : (int)rd);
                w = max_gap < p->w ? max_gap : p->w;
                if (qd - rd < w && rd - qd < w) break;
            }
            if (i < n_regs) {                       /* src/bwamem.c:1246-1262 */
                for (i = k + 1; i < c->n; ++i) {
                    const chain_seed_t *t;
                    if (srt[i] == 0) continue;
                    t = &sd[(uint32_t)srt[i]];
                    if (t->len < s->len * .95) continue;
                    if (s->qbeg <= t->qbeg && s->qbeg + s->len - t->qbeg >= s->len >> 2 && t->qbeg - s->qbeg != t->rbeg - s->rbeg) break;
                    if (t->qbeg <= s->qbeg && t->qbeg + t->len - s->qbeg >= s->len >> 2 && s->qbeg - t->qbeg != s->rbeg - t->rbeg) break;
                }
                if (i == c->n) { srt[k] = 0; continue; }
            }
            if (n_regs >= cap_regs) { rc = -1; break; }
            chain_reg_t *a = &regs[n_regs++];
            memset(a, 0, sizeof(*a));
            a->w = opt->w;
            a->score = a->truesc = -1;
            a->rid = c->rid;
            {   /* src/bwamem.c:1284-1298 (FILTER_COEF 0.85) */
                int fwd_ = (int)(0.85 * (l_query - (s->qbeg + s->len)));
                a->qe_est = (s->qbeg + s->len) + fwd_ < l_query ? (s->qbeg + s->len) + fwd_ : l_query;
                a->re_est = (s->rbeg + s->len) + fwd_ < l_pac << 1 ? (s->rbeg + s->len) + fwd_ : l_pac << 1;
                int back = (int)(0.85 * (s->qbeg + 1));
                a->qb_est = (s->qbeg - back) > 0 ? (s->qbeg - back) : 0;
                a->rb_est = (s->rbeg - back) > 0 ? (s->rbeg - back) : 0;
                if (a->rb_est < l_pac && l_pac < a->qe_est) {     /* sic: the reference compares l_pac with qe_est */
                    if (s->rbeg < l_pac) a->re_est = l_pac; else a->rb_est = l_pac;
                }
            }
            const int lq = s->qbeg, lt = (int)(s->rbeg - rmax[0]);
            const int rq = l_query - (lq + s->len), rt = (int)(l_refer - (lt + s->len));
            uint8_t *left_query = NULL, *left_refer = NULL;
            const uint8_t *right_query = query + lq + s->len, *right_refer = rseq + lt + s->len;
            if (lq > 0) {
                left_query = (uint8_t *)malloc((size_t)lq);
                for (i = 0; i < lq; ++i) left_query[i] = query[lq - 1 - i];
                left_refer = (uint8_t *)malloc((size_t)(lt > 0 ? lt : 1));
                for (i = 0; i < lt; ++i) left_refer[i] = rseq[lt - 1 - i];
            }
            a->score = s->len; a->truesc = a->score;
            a->query_seed_begin = s->qbeg; a->target_seed_begin = s->rbeg;
            if (lq == 0 && rq > 0) {
                a->align_sides = 1;
                rc |= fill_ext(&side[1], cap_jobs, cap_bytes, right_refer, right_query, rt, rq, s->len);
                a->where_is_long = 1;
            } else if (lq > 0 && rq == 0) {
                a->align_sides = 1;
                rc |= fill_ext(&side[1], cap_jobs, cap_bytes, left_refer, left_query, lt, lq, s->len);
                a->where_is_long = 0;
            } else if (lq > 0 && rq > 0) {
                a->align_sides = 2;
                if (s->qbeg + (s->len / 2) < l_query / 2) {
                    rc |= fill_ext(&side[0], cap_jobs, cap_bytes, left_refer, left_query, lt, lq, s->len);
                    rc |= fill_ext(&side[1], cap_jobs, cap_bytes, right_refer, right_query, rt, rq, s->len);
                    a->where_is_long = 1;
                } else {
                    rc |= fill_ext(&side[1], cap_jobs, cap_bytes, left_refer, left_query, lt, lq, s->len);
                    rc |= fill_ext(&side[0], cap_jobs, cap_bytes, right_refer, right_query, rt, rq, s->len);
                    a->where_is_long = 0;
                }
            } else {
                a->align_sides = 0;
                a->score = a->truesc = s->score;
            }
            free(left_query); free(left_refer);
            {   /* seedcov, src/bwamem.c:1459-1466: evaluated before any extension result exists, i.e. with qb = qe = rb = re = 0
                   unless the seed spans the whole read */
                int64_t qb = 0, qe = 0, rb = 0, re = 0;
                if (a->align_sides == 0) { qb = 0; qe = l_query; rb = s->rbeg; re = s->rbeg + s->len; }
                for (i = 0, a->seedcov = 0; i < c->n; ++i) {
                    const chain_seed_t *t = &sd[i];
                    if (t->qbeg >= qb && t->qbeg + t->len <= qe && t->rbeg >= rb && t->rbeg + t->len <= re) a->seedcov += t->len;
                }
            }
            a->seedlen0 = s->len;
            a->frac_rep = c->frac_rep;
        }
        free(rseq); free(srt);
    }
    *n_regs_out = n_regs;
    n_jobs[0] = side[0].n; n_jobs[1] = side[1].n;
    return rc ? -1 : 0;
}

/* triples = {aln_score, query_batch_end, target_batch_end} per job of each side, in job order */
void chain_regs_finish(int l_query, int n_regs, const chain_reg_t *regs, const int32_t *short_triples, const int32_t *long_triples,
                       chain_aln_t *out)
{ /* src/bwamem.c:2217-2306 */
    int is = 0, il = 0;
    for (int i = 0; i < n_regs; ++i) {
        const chain_reg_t *a = &regs[i];
        chain_aln_t *o = &out[i];
        int32_t part[2][3] = {{0, 0, 0}, {0, 0, 0}};     /* [LEFT 0 / RIGHT 1] = {score, query_end, ref_end} */
        if (a->seedlen0 != l_query && a->align_sides > 0) {
            const int32_t *t = long_triples + 3 * il++;
            memcpy(part[a->where_is_long ? 1 : 0], t, sizeof(int32_t) * 3);
            if (a->align_sides == 2) {
                const int32_t *u = short_triples + 3 * is++;
                memcpy(part[a->where_is_long ? 0 : 1], u, sizeof(int32_t) * 3);
            }
            o->score = part[0][0] + part[1][0] - (a->align_sides == 2 ? a->seedlen0 : 0);
            o->qb = a->query_seed_begin - part[0][1];
            o->qe = a->query_seed_begin + a->seedlen0 + part[1][1];
            o->rb = a->target_seed_begin - part[0][2];
            o->re = a->target_seed_begin + a->seedlen0 + part[1][2];
            o->truesc = o->score;
        } else {    /* src/bwamem.c:1437: the seed covers the read */
            o->score = o->truesc = a->score; o->qb = 0; o->qe = l_query; o->rb = a->target_seed_begin; o->re = a->target_seed_begin + a->seedlen0;
        }
    }
}
