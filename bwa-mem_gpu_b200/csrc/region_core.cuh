// region_core.cuh -- a read's alignment regions between the extension results and its SAM records, for ONE read, as the reference
// fork's host code does it:  mem_sort_dedup_patch (src/bwamem.c:620-681) with mem_patch_reg (:580-618, its global alignment score =
// bwa_gen_cigar2 src/bwa.c:111-216 -> ksw_global2 src/ksw.c:1120-1241 without the backtrack), the is_alt marking of its caller
// (:2321-2325), mem_mark_primary_se (:715-760) and mem_approx_mapq_se as mem_reg2aln applies it (:1690-1716, :2363).
//
// Shared source, like chain_core.cuh: the same functions compile for the device (one lane per read: the reference's logic is
// sequential per read -- sort, then every region against the regions before it) and for the host
// (tests/host_emul/region_host.cpp, run against oracle/region_oracle.c and the fork's golden vectors on the CPU box).
// The sorts are the reference's ks_introsort (src/ksort.h:146-226) step by step: it is not stable and equal keys are common
// (equal scores, equal ends), so any other sort gives other answers.  Float comparisons keep the reference's operand types.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define RG_FN __host__ __device__ inline
#define RG_MFN __host__ __device__
#else
#define RG_FN static inline
#define RG_MFN
#endif

namespace b200region {

constexpr int32_t MINUS_INF = -0x40000000;

struct Opt {                    // the mem_opt_t fields read here (src/bwamem.h:34-73)
    int32_t a, b, o_del, e_del, o_ins, e_ins, w, min_seed_len, max_chain_gap, mapQ_coef_fac;
    float mask_level, mask_level_redun, mapQ_coef_len;
};

struct Reg {                    // the mem_alnreg_t fields read or written here (src/bwamem.h:83-112); layout of oracle/region_oracle.h
    int64_t rb, re;
    uint64_t hash;
    int32_t qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0, n_comp, is_alt;
    float frac_rep;
    int32_t mapq;
};

struct EH { int32_t h, e; };

// bases of the read (codes 0..4) and of the forward reference (codes 0..3); the kernels pass accessors over the packed words
struct ByteQuery { const uint8_t *q; RG_MFN int operator()(int i) const { return q[i]; } };
struct ByteRef { const uint8_t *fwd; RG_MFN int operator()(int64_t p) const { return fwd[p]; } };

RG_FN int8_t sub_score(const Opt &o, int t, int q) { return (t > 3 || q > 3) ? (int8_t)-1 : (int8_t)(t == q ? o.a : -o.b); }   // bwa_fill_scmat

// ------------------------------------------------------------------------------------------ ks_introsort
struct LtEnd { RG_MFN bool operator()(const Reg &a, const Reg &b) const { return a.re < b.re; } };
struct LtScore { RG_MFN bool operator()(const Reg &a, const Reg &b) const
    { return a.score > b.score || (a.score == b.score && (a.rb < b.rb || (a.rb == b.rb && a.qb < b.qb))); } };
struct LtHash { RG_MFN bool operator()(const Reg &a, const Reg &b) const
    { return a.score > b.score || (a.score == b.score && (a.is_alt < b.is_alt || (a.is_alt == b.is_alt && a.hash < b.hash))); } };
struct LtHash2 { RG_MFN bool operator()(const Reg &a, const Reg &b) const
    { return a.is_alt < b.is_alt || (a.is_alt == b.is_alt && (a.score > b.score || (a.score == b.score && a.hash < b.hash))); } };

RG_FN void swap_reg(Reg &x, Reg &y) { Reg t = x; x = y; y = t; }

template <class Lt> RG_FN void insertsort(Lt lt, Reg *s, Reg *t)
{
    for (Reg *i = s + 1; i < t; ++i)
        for (Reg *j = i; j > s && lt(*j, *(j - 1)); --j) swap_reg(*j, *(j - 1));
}
template <class Lt> RG_FN void combsort(Lt lt, int64_t n, Reg *a)
{
    const double shrink = 1.2473309501039786540366528676643;
    bool swapped;
    int64_t gap = n;
    do {
        if (gap > 2) { gap = (int64_t)((double)gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
        swapped = false;
        for (Reg *i = a; i < a + n - gap; ++i)
            if (lt(*(i + gap), *i)) { swap_reg(*i, *(i + gap)); swapped = true; }
    } while (swapped || gap > 2);
    if (gap != 1) insertsort(lt, a, a + n);
}
template <class Lt> RG_FN void introsort(Lt lt, int64_t n, Reg *a)
{
    struct Frame { Reg *left, *right; int depth; };
    Frame stack[64], *top = stack;          // the larger side is pushed, the smaller one continued: depth <= log2 n
    int d;
    if (n < 1) return;
    if (n == 2) { if (lt(a[1], a[0])) swap_reg(a[0], a[1]); return; }
    for (d = 2; ((int64_t)1 << d) < n; ++d) {}
    Reg *s = a, *t = a + (n - 1), *i, *j, *k;
    d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { combsort(lt, (int64_t)(t - s + 1), s); t = s; continue; }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (lt(*k, *i)) { if (lt(*k, *j)) k = j; }
            else k = lt(*j, *i) ? i : j;
            const Reg pivot = *k;
            if (k != t) swap_reg(*k, *t);
            for (;;) {
                do ++i; while (lt(*i, pivot));
                do --j; while (i <= j && lt(pivot, *j));
                if (j <= i) break;
                swap_reg(*i, *j);
            }
            swap_reg(*i, *t);
            if (i - s > t - i) {
                if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == stack) { insertsort(lt, a, a + n); return; }
            --top; s = top->left; t = top->right; d = top->depth;
        }
    }
}

// ------------------------------------------------------------------------------------------ global score of query[q0, q0+lq) vs [rb, re)
// bwa_gen_cigar2 for its score: reference bases as bns_get_seq returns them, both sequences reversed on the reverse strand so
// that the band and the tie-breaking are the reference's; eh = lq + 1 cells of scratch.  A job the reference rejects scores 0
// (there the caller's variable stays uninitialised).
template <class Q, class R>
RG_FN int global_score(const Opt &o, int w_, int64_t l_pac, R ref, Q query, int q0, int lq, int64_t rb, int64_t re, EH *eh)
{
    if (lq <= 0 || rb >= re || (rb < l_pac && re > l_pac) || rb < 0 || re > (l_pac << 1)) return 0;
    const int64_t rlen = re - rb;
    const bool rev = rb >= l_pac;
    // base i of the (possibly reversed) sequences
    auto qat = [&](int j) { return query(q0 + (rev ? lq - 1 - j : j)); };
    auto tat = [&](int64_t i) {
        if (!rev) return ref(rb + i);
        const int64_t p = rb + (rlen - 1 - i);                      // position on the doubled text, reverse strand
        return 3 - ref((l_pac << 1) - 1 - p);
    };
    if (lq == rlen && w_ == 0) {                                        // no gap: no DP
        int score = 0;
        for (int i = 0; i < lq; ++i) score += sub_score(o, tat(i), qat(i));
        return score;
    }
    int max_ins = (int)((double)(((lq + 1) >> 1) * o.a - o.o_ins) / o.e_ins + 1.);
    int max_del = (int)((double)(((lq + 1) >> 1) * o.a - o.o_del) / o.e_del + 1.);
    int max_gap = max_ins > max_del ? max_ins : max_del;
    max_gap = max_gap > 1 ? max_gap : 1;
    const int diff = (int)(rlen > lq ? rlen - lq : lq - rlen);
    int w = (max_gap + diff + 1) >> 1;
    w = w < w_ ? w : w_;
    w = w > diff + 3 ? w : diff + 3;
    // ksw_global2 without the backtrack
    const int oe_del = o.o_del + o.e_del, oe_ins = o.o_ins + o.e_ins, tlen = (int)rlen;
    int j;
    eh[0].h = 0; eh[0].e = MINUS_INF;
    for (j = 1; j <= lq && j <= w; ++j) { eh[j].h = -(o.o_ins + o.e_ins * j); eh[j].e = MINUS_INF; }
    for (; j <= lq; ++j) eh[j].h = eh[j].e = MINUS_INF;
    for (int i = 0; i < tlen; ++i) {
        int32_t f = MINUS_INF, h1, t;
        const int tb = tat(i);
        const int beg = i > w ? i - w : 0;
        const int end = i + w + 1 < lq ? i + w + 1 : lq;
        h1 = beg == 0 ? -(o.o_del + o.e_del * (i + 1)) : MINUS_INF;
        for (j = beg; j < end; ++j) {
            EH *p = &eh[j];
            int32_t h, m = p->h, e = p->e;
            p->h = h1;
            m += sub_score(o, tb, qat(j));
            h = m >= e ? m : e;
            h = h >= f ? h : f;
            h1 = h;
            t = m - oe_del; e -= o.e_del; e = e > t ? e : t; p->e = e;
            t = m - oe_ins; f -= o.e_ins; f = f > t ? f : t;
        }
        eh[end].h = h1; eh[end].e = MINUS_INF;
    }
    return eh[lq].h;
}

// mem_patch_reg: can hit a (upstream) be joined with hit b by one global alignment?  Returns its score or 0.
template <class Q, class R>
RG_FN int patch_reg(const Opt &o, int64_t l_pac, R ref, Q query, const Reg &a, const Reg &b, int *w_out, EH *eh)
{
    if (a.rb < l_pac && b.rb >= l_pac) return 0;                         // on different strands
    if (a.qb >= b.qb || a.qe >= b.qe || a.re >= b.re) return 0;          // not colinear
    int w = (int)((a.re - b.rb) - (a.qe - b.qb));                        // required bandwidth
    w = w > 0 ? w : -w;
    double r = (double)(a.re - b.rb) / (double)(b.re - a.rb) - (double)(a.qe - b.qb) / (double)(b.qe - a.qb);   // relative bandwidth
    r = r > 0. ? r : -r;
    if (a.re < b.rb || a.qe < b.qb) {                                    // no overlap on query or on ref
        if (w > o.w << 1 || r >= 0.05f) return 0;
    } else if (w > o.w << 2 || r >= 0.05f * 2) return 0;
    w += a.w + b.w;
    w = w < o.w << 2 ? w : o.w << 2;
    const int score = global_score(o, w, l_pac, ref, query, a.qb, b.qe - a.qb, a.rb, b.re, eh);
    const int q_s = (int)((double)(b.qe - a.qb) / ((b.qe - b.qb) + (a.qe - a.qb)) * (b.score + a.score) + .499);
    const int r_s = (int)((double)(b.re - a.rb) / (double)((b.re - b.rb) + (a.re - a.rb)) * (b.score + a.score) + .499);
    if ((double)score / (q_s > r_s ? q_s : r_s) < 0.90f) return 0;
    *w_out = w;
    return score;
}

template <class Q, class R>
RG_FN int sort_dedup_patch(const Opt &o, int64_t l_pac, R ref, Q query, int n, Reg *a, EH *eh)
{
    int m, i, j;
    if (n <= 1) return n;
    introsort(LtEnd(), n, a);                                            // by END position
    for (i = 0; i < n; ++i) a[i].n_comp = 1;
    for (i = 1; i < n; ++i) {
        Reg *p = &a[i];
        if (p->rid != a[i - 1].rid || p->rb >= a[i - 1].re + o.max_chain_gap) continue;
        for (j = i - 1; j >= 0 && p->rid == a[j].rid && p->rb < a[j].re + o.max_chain_gap; --j) {
            Reg *q = &a[j];
            int score, w;
            if (q->qe == q->qb) continue;                                // excluded earlier
            const int64_t pr = q->re - p->rb;                            // overlap on the reference
            const int64_t pq = q->qb < p->qb ? q->qe - p->qb : p->qe - q->qb;
            const int64_t mr = q->re - q->rb < p->re - p->rb ? q->re - q->rb : p->re - p->rb;
            const int64_t mq = q->qe - q->qb < p->qe - p->qb ? q->qe - q->qb : p->qe - p->qb;
            if ((float)pr > o.mask_level_redun * (float)mr && (float)pq > o.mask_level_redun * (float)mq) {   // one of the two is redundant
                if (p->score < q->score) { p->qe = p->qb; break; }
                else q->qe = q->qb;
            } else if (q->rb < p->rb && (score = patch_reg(o, l_pac, ref, query, *q, *p, &w, eh)) > 0) {      // merge q into p
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
    for (i = 0, m = 0; i < n; ++i)                                       // drop the excluded
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    n = m;
    introsort(LtScore(), n, a);
    for (i = 1; i < n; ++i)                                              // identical hits
        if (a[i].score == a[i - 1].score && a[i].rb == a[i - 1].rb && a[i].qb == a[i - 1].qb) a[i].qe = a[i].qb;
    for (i = 1, m = 1; i < n; ++i)
        if (a[i].qe > a[i].qb) { if (m != i) a[m++] = a[i]; else ++m; }
    return m;
}

RG_FN uint64_t hash_64(uint64_t key)
{ // src/utils.h:126-137
    key += ~(key << 32); key ^= (key >> 22); key += ~(key << 13); key ^= (key >> 8);
    key += (key << 3); key ^= (key >> 15); key += ~(key << 27); key ^= (key >> 31);
    return key;
}

// z = n ints of scratch: indexes of the hits that are primary so far
RG_FN void mark_primary_core(const Opt &o, int n, Reg *a, int32_t *z)
{
    int i, k, nz = 0, tmp;
    tmp = o.a + o.b;
    tmp = o.o_del + o.e_del > tmp ? o.o_del + o.e_del : tmp;
    tmp = o.o_ins + o.e_ins > tmp ? o.o_ins + o.e_ins : tmp;
    z[nz++] = 0;
    for (i = 1; i < n; ++i) {
        for (k = 0; k < nz; ++k) {
            const int j = z[k];
            const int b_max = a[j].qb > a[i].qb ? a[j].qb : a[i].qb;
            const int e_min = a[j].qe < a[i].qe ? a[j].qe : a[i].qe;
            if (e_min > b_max) {                                         // overlap on the query
                const int min_l = a[i].qe - a[i].qb < a[j].qe - a[j].qb ? a[i].qe - a[i].qb : a[j].qe - a[j].qb;
                if ((float)(e_min - b_max) >= (float)min_l * o.mask_level) {   // significant
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

RG_FN int mark_primary_se(const Opt &o, int n, Reg *a, int64_t id, int32_t *z)
{
    int i, n_pri;
    if (n == 0) return 0;
    for (i = n_pri = 0; i < n; ++i) {
        a[i].sub = a[i].alt_sc = 0; a[i].secondary = a[i].secondary_all = -1; a[i].hash = hash_64((uint64_t)(id + i));
        if (!a[i].is_alt) ++n_pri;
    }
    introsort(LtHash(), n, a);
    mark_primary_core(o, n, a, z);
    for (i = 0; i < n; ++i) {
        Reg *p = &a[i];
        p->secondary_all = i;                                            // rank of the first round
        if (!p->is_alt && p->secondary >= 0 && a[p->secondary].is_alt) p->alt_sc = a[p->secondary].score;
    }
    if (n_pri >= 0 && n_pri < n) {
        if (n_pri > 0) introsort(LtHash2(), n, a);
        for (i = 0; i < n; ++i) z[a[i].secondary_all] = i;
        for (i = 0; i < n; ++i) {
            if (a[i].secondary >= 0) {
                a[i].secondary_all = z[a[i].secondary];
                if (a[i].is_alt) a[i].secondary = 0x7fffffff;
            } else a[i].secondary_all = -1;
        }
        if (n_pri > 0) {                                                 // among the primary-assembly hits only
            for (i = 0; i < n_pri; ++i) { a[i].sub = 0; a[i].secondary = -1; }
            mark_primary_core(o, n_pri, a, z);
        }
    } else {
        for (i = 0; i < n; ++i) a[i].secondary_all = a[i].secondary;
    }
    return n_pri;
}

RG_FN int approx_mapq_se(const Opt &o, const Reg &a)
{
    int mapq, l, sub = a.sub ? a.sub : o.min_seed_len * o.a;
    sub = a.csub > sub ? a.csub : sub;
    if (sub >= a.score) return 0;
    l = a.qe - a.qb > a.re - a.rb ? a.qe - a.qb : (int)(a.re - a.rb);
    const double identity = 1. - (double)(l * o.a - a.score) / (o.a + o.b) / l;
    if (a.score == 0) mapq = 0;
    else if (o.mapQ_coef_len > 0) {
        double tmp = (float)l < o.mapQ_coef_len ? 1. : o.mapQ_coef_fac / log((double)l);
        tmp *= identity * identity;
        mapq = (int)(6.02 * (a.score - sub) / o.a * tmp * tmp + .499);
    } else {
        mapq = (int)(30.0 * (1. - (double)sub / a.score) * log((double)a.seedcov) + .499);
        mapq = identity < 0.95 ? (int)(mapq * identity * identity + .499) : mapq;
    }
    if (a.sub_n > 0) mapq -= (int)(4.343 * log((double)(a.sub_n + 1)) + .499);
    if (mapq > 60) mapq = 60;
    if (mapq < 0) mapq = 0;
    mapq = (int)(mapq * (1. - a.frac_rep) + .499);
    return mapq;
}

// the stage for one read; eh = l_query + 1 cells, z = n ints.  Returns the new count.
template <class Q, class R>
RG_FN int finish_read(const Opt &o, int64_t l_pac, const int32_t *ctg_alt, R ref, Q query, int n, Reg *a, int64_t id, int *n_pri, EH *eh, int32_t *z)
{
    n = sort_dedup_patch(o, l_pac, ref, query, n, a, eh);
    for (int i = 0; i < n; ++i)
        if (a[i].rid >= 0 && ctg_alt && ctg_alt[a[i].rid]) a[i].is_alt = 1;
    *n_pri = mark_primary_se(o, n, a, id, z);
    for (int i = 0; i < n; ++i) a[i].mapq = a[i].secondary < 0 ? approx_mapq_se(o, a[i]) : 0;
    return n;
}

} // namespace b200region
