// chain_core.cuh -- seeds -> chains -> extension jobs for ONE read, as the reference fork's host code does it.
//
// Shared between the CUDA kernels (chain.cu: one lane per read) and a host build of the same source
// (tests/host_emul/chain_host.cpp, run against the oracle on the CPU box).  Parity target, all in the fork's
// src/bwamem.c:  mem_chain :404-476 with test_and_merge :337-359 and the kbtree it keeps chains in
// (src/kbtree.h:117-224; KB_DEFAULT_SIZE 512 and a 40-byte key give t = 5, so a node holds at most 9 chains),
// mem_chain_weight :361-384, mem_chain_flt :488-560 (ks_introsort, src/ksort.h:146-226: not stable, so it is
// followed step by step), mem_chain2aln :1170-1479 with cal_max_gap :996-1002, bns_intv2rid / bns_fetch_seq's
// contig clamp (src/bntseq.c:349-373,531-552), and the result gathering :2286-2306.
//
// The chain tree is kept as a B-tree rather than a sorted list because the reference's answer depends on the
// tree's shape when two chains of a read start at the same reference position.  Work per read is sequential in
// the reference (every seed is tested against the chains built so far; every seed of a chain against the
// regions accepted so far), so the decomposition is one lane per read; a read's working set lives in its own
// slice of flat scratch arrays indexed by the read's seed offset.
#pragma once
#include <stdint.h>
#include <math.h>
#include "bwamem_b200.h"
#include "sw_core.cuh"

#ifdef __CUDACC__
#define CH_FN __host__ __device__ inline
#define CH_DEV __device__ inline
#else
#define CH_FN static inline
#define CH_DEV static inline
#endif

namespace b200chain {

constexpr int KB_T = 5;
constexpr int KB_N = 2 * KB_T - 1;

struct Contigs { const int64_t *off; const int32_t *len; const int32_t *alt; int32_t n; int64_t l_pac; };

struct ChainW {                 // a chain under construction: what test_and_merge reads of it
    int64_t pos, last_rbeg;     // rbeg of the first / last seed
    int32_t first_qbeg, last_qbeg, last_len, rid;
    int32_t head, tail, n;      // seeds as a linked list over the read's seed slots
    int32_t w, kept, first;
};

struct KbNode { int32_t n, is_internal; int32_t key[KB_N]; int32_t ptr[KB_N + 1]; };

struct ReadIO {
    // the read's seeds (bwa_b200_seeds_t slice)
    const uint64_t *rbeg; const int32_t *qq; const uint32_t *score; uint32_t ns; int32_t l_query; int32_t layout_all;
    // scratch, ns entries each (nodes: n_node_cap entries)
    ChainW *ch; int32_t *nxt; int32_t *sq; int32_t *ord; int32_t *kidx; KbNode *nodes; int32_t n_node_cap;
    // outputs, ns entries each
    bwa_b200_chain_t *chains; bwa_b200_chain_seed_t *cseeds;
};

CH_FN int nodes_needed(uint32_t ns) { return ns <= (uint32_t)KB_N ? 1 : (int)(ns / 3 + 4); }

// ------------------------------------------------------------------------------------ reference coordinates
CH_FN int pos2rid(const Contigs &c, int64_t pos_f)
{ // src/bntseq.c:349-363
    int left = 0, mid = 0, right = c.n;
    if (pos_f >= c.l_pac) return -1;
    while (left < right) {
        mid = (left + right) >> 1;
        if (pos_f >= c.off[mid]) {
            if (mid == c.n - 1) break;
            if (pos_f < c.off[mid + 1]) break;
            left = mid + 1;
        } else right = mid;
    }
    return mid;
}
CH_FN int64_t depos(const Contigs &c, int64_t pos, int *is_rev) { return (*is_rev = (pos >= c.l_pac)) ? (c.l_pac << 1) - 1 - pos : pos; }
CH_FN int intv2rid(const Contigs &c, int64_t rb, int64_t re)
{ // src/bntseq.c:365-373
    int is_rev;
    if (rb < c.l_pac && re > c.l_pac) return -2;
    const int rid_b = pos2rid(c, depos(c, rb, &is_rev));
    const int rid_e = rb < re ? pos2rid(c, depos(c, re - 1, &is_rev)) : rid_b;
    return rid_b == rid_e ? rid_b : -1;
}

// ------------------------------------------------------------------------------------ kbtree, t = 5
struct Tree { KbNode *nd; const ChainW *ch; int32_t n_nodes, cap, root; };

CH_FN int kb_cmp(int64_t a, int64_t b) { return (int)(b < a) - (int)(a < b); }
CH_FN int kb_new(Tree &b, int internal)
{
    if (b.n_nodes >= b.cap) return -1;
    KbNode &x = b.nd[b.n_nodes];
    x.n = 0; x.is_internal = internal;
    return b.n_nodes++;
}
CH_FN int kb_getp_aux(const Tree &b, const KbNode &x, int64_t k, int *r)
{ // src/kbtree.h:117-131
    int begin = 0, end = x.n;
    if (x.n == 0) return -1;
    while (begin < end) {
        const int mid = (begin + end) >> 1;
        if (kb_cmp(b.ch[x.key[mid]].pos, k) < 0) begin = mid + 1; else end = mid;
    }
    if (begin == x.n) { *r = 1; return x.n - 1; }
    if ((*r = kb_cmp(k, b.ch[x.key[begin]].pos)) < 0) --begin;
    return begin;
}
CH_FN int kb_lower(const Tree &b, int64_t k)
{ // kb_intervalp, src/kbtree.h:151-168
    int x = b.root, lower = -1, r = 0;
    for (;;) {
        const KbNode &nd = b.nd[x];
        const int i = kb_getp_aux(b, nd, k, &r);
        if (i >= 0 && r == 0) return nd.key[i];
        if (i >= 0) lower = nd.key[i];
        if (!nd.is_internal) return lower;
        x = nd.ptr[i + 1];
    }
}
CH_FN int kb_split(Tree &b, int xi, int i, int yi)
{ // src/kbtree.h:176-191
    const int zi = kb_new(b, b.nd[yi].is_internal);
    if (zi < 0) return -1;
    KbNode &x = b.nd[xi], &y = b.nd[yi], &z = b.nd[zi];
    z.n = KB_T - 1;
    for (int j = 0; j < KB_T - 1; ++j) z.key[j] = y.key[KB_T + j];
    if (y.is_internal) for (int j = 0; j < KB_T; ++j) z.ptr[j] = y.ptr[KB_T + j];
    y.n = KB_T - 1;
    for (int j = x.n; j > i; --j) x.ptr[j + 1] = x.ptr[j];
    x.ptr[i + 1] = zi;
    for (int j = x.n; j > i; --j) x.key[j] = x.key[j - 1];
    x.key[i] = y.key[KB_T - 1];
    ++x.n;
    return 0;
}
CH_FN int kb_put(Tree &b, int c)
{ // kb_putp + __kb_putp_aux, src/kbtree.h:192-224 (the recursion is a descent: written as a loop)
    const int64_t k = b.ch[c].pos;
    int r = 0;
    if (b.nd[b.root].n == KB_N) {
        const int s = kb_new(b, 1), old = b.root;
        if (s < 0) return -1;
        b.root = s;
        b.nd[s].ptr[0] = old;
        if (kb_split(b, s, 0, old)) return -1;
    }
    int xi = b.root;
    for (;;) {
        KbNode &x = b.nd[xi];
        if (!x.is_internal) {
            const int i = kb_getp_aux(b, x, k, &r);
            for (int j = x.n - 1; j > i; --j) x.key[j + 1] = x.key[j];
            x.key[i + 1] = c;
            ++x.n;
            return 0;
        }
        int i = kb_getp_aux(b, x, k, &r) + 1;
        if (b.nd[x.ptr[i]].n == KB_N) {
            if (kb_split(b, xi, i, x.ptr[i])) return -1;
            if (kb_cmp(k, b.ch[b.nd[xi].key[i]].pos) > 0) ++i;
        }
        xi = b.nd[xi].ptr[i];
    }
}
// in-order traversal, src/kbtree.h:336-358
CH_FN int kb_traverse(const Tree &b, int32_t *out)
{
    int sx[16], si[16], sp = 0, n = 0;
    sx[0] = b.root; si[0] = 0;
    for (;;) {
        while (sx[sp] >= 0 && si[sp] <= b.nd[sx[sp]].n) {
            const KbNode &x = b.nd[sx[sp]];
            if (sp >= 14) return -1;
            sx[sp + 1] = x.is_internal ? x.ptr[si[sp]] : -1; si[sp + 1] = 0;
            ++sp;
        }
        --sp;
        if (sp < 0) break;
        if (sx[sp] >= 0 && si[sp] < b.nd[sx[sp]].n) out[n++] = b.nd[sx[sp]].key[si[sp]];
        ++si[sp];
    }
    return n;
}

// ------------------------------------------------------------------------------------ mem_chain_flt pieces
CH_FN int chain_weight(const ReadIO &io, const ChainW &c)
{ // src/bwamem.c:361-384
    int64_t end = 0;
    int w = 0, tmp;
    for (int j = c.head; j >= 0; j = io.nxt[j]) {
        const int qb = io.sq[2 * j], ln = io.sq[2 * j + 1];
        if (qb >= end) w += ln;
        else if (qb + ln > end) w += (int)(qb + ln - end);
        end = end > qb + ln ? end : qb + ln;
    }
    tmp = w; w = 0; end = 0;
    for (int j = c.head; j >= 0; j = io.nxt[j]) {
        const int64_t rb = (int64_t)io.rbeg[j];
        const int ln = io.sq[2 * j + 1];
        if (rb >= end) w += ln;
        else if (rb + ln > end) w += (int)(rb + ln - end);
        end = end > rb + ln ? end : rb + ln;
    }
    w = w < tmp ? w : tmp;
    return w < 1 << 30 ? w : (1 << 30) - 1;
}

#define CH_LT(x, y) (io.ch[x].w > io.ch[y].w)      // flt_lt, src/bwamem.c:485
CH_FN void flt_insertsort(const ReadIO &io, int32_t *a, int s, int t)
{ // [s, t)
    for (int i = s + 1; i < t; ++i)
        for (int j = i; j > s && CH_LT(a[j], a[j - 1]); --j) { const int32_t tmp = a[j]; a[j] = a[j - 1]; a[j - 1] = tmp; }
}
CH_FN void flt_combsort(const ReadIO &io, int32_t *a, int s, int n)
{ // src/ksort.h:154-175 on a[s .. s+n)
    const double shrink_factor = 1.2473309501039786540366528676643;
    int do_swap;
    int gap = n;
    do {
        if (gap > 2) { gap = (int)(gap / shrink_factor); if (gap == 9 || gap == 10) gap = 11; }
        do_swap = 0;
        for (int i = s; i < s + n - gap; ++i) {
            const int j = i + gap;
            if (CH_LT(a[j], a[i])) { const int32_t tmp = a[i]; a[i] = a[j]; a[j] = tmp; do_swap = 1; }
        }
    } while (do_swap || gap > 2);
    if (gap != 1) flt_insertsort(io, a, s, s + n);
}
CH_FN void flt_introsort(const ReadIO &io, int32_t *a, int n)
{ // src/ksort.h:176-226
    int sl[72], sr[72], sd[72], top = 0;
    int d, s, t, i, j, k;
    int32_t rp, tmp;
    if (n < 1) return;
    if (n == 2) { if (CH_LT(a[1], a[0])) { tmp = a[0]; a[0] = a[1]; a[1] = tmp; } return; }
    for (d = 2; (1u << d) < (unsigned)n; ++d) {}
    s = 0; t = n - 1; d <<= 1;
    for (;;) {
        if (s < t) {
            if (--d == 0) { flt_combsort(io, a, s, t - s + 1); t = s; continue; }
            i = s; j = t; k = i + ((j - i) >> 1) + 1;
            if (CH_LT(a[k], a[i])) { if (CH_LT(a[k], a[j])) k = j; }
            else k = CH_LT(a[j], a[i]) ? i : j;
            rp = a[k];
            if (k != t) { tmp = a[k]; a[k] = a[t]; a[t] = tmp; }
            for (;;) {
                do ++i; while (CH_LT(a[i], rp));
                do --j; while (i <= j && CH_LT(rp, a[j]));
                if (j <= i) break;
                tmp = a[i]; a[i] = a[j]; a[j] = tmp;
            }
            tmp = a[i]; a[i] = a[t]; a[t] = tmp;
            if (i - s > t - i) {
                if (i - s > 16) { sl[top] = s; sr[top] = i - 1; sd[top] = d; ++top; }
                s = t - i > 16 ? i + 1 : t;
            } else {
                if (t - i > 16) { sl[top] = i + 1; sr[top] = t; sd[top] = d; ++top; }
                t = i - s > 16 ? i - 1 : s;
            }
        } else {
            if (top == 0) { flt_insertsort(io, a, 0, n); return; }
            --top; s = sl[top]; t = sr[top]; d = sd[top];
        }
    }
}
#undef CH_LT

// ------------------------------------------------------------------------------------ mem_chain + mem_chain_flt
// Returns the number of chains kept (written to io.chains / io.cseeds in the reference's final order), or
//   -3  internal capacity (node pool / traversal depth)
// A read for which flt_seeds_applies() holds goes through seed_sw / flt_seeds_apply before chain2aln_read.
CH_FN int chain_read(const bwa_b200_chain_params_t &P, const Contigs &ctg, const ReadIO &io)
{
    if (io.l_query < P.min_seed_len || io.ns == 0) return 0;
    Tree bt;
    bt.nd = io.nodes; bt.ch = io.ch; bt.n_nodes = 0; bt.cap = io.n_node_cap;
    bt.root = kb_new(bt, 0);
    int n_ch = 0;

#define CH_GROUP(i_)                                                                                          \
    const uint32_t s = io.score[i_], step = s > (uint32_t)P.max_occ ? s / (uint32_t)P.max_occ : 1u;            \
    uint32_t cnt = (s + step - 1) / step;                                                                     \
    if (cnt > (uint32_t)P.max_occ) cnt = (uint32_t)P.max_occ;                                                 \
    uint32_t grp = io.layout_all ? s : cnt;                                                                   \
    if (grp == 0) grp = 1;

    int b = 0, e = 0, l_rep = 0;          // src/bwamem.c:415-422
    for (uint32_t i = 0; i < io.ns;) {
        CH_GROUP(i)
        const int sb = io.qq[2 * i], se = io.qq[2 * i + 1];
        if (s > (uint32_t)P.max_occ) {
            if (sb > e) { l_rep += e - b; b = sb; e = se; }
            else e = e > se ? e : se;
        }
        i += grp;
    }
    l_rep += e - b;

    for (uint32_t i = 0; i < io.ns;) {   // src/bwamem.c:423-452
        CH_GROUP(i)
        const int qbeg = io.qq[2 * i], slen = io.qq[2 * i + 1] - io.qq[2 * i];
        for (uint32_t k = 0, count = 0; k < s && count < (uint32_t)P.max_occ; k += step, ++count) {
            const int idx = (int)(io.layout_all ? i + k : i + count);
            const int64_t rb = (int64_t)io.rbeg[idx];
            const int rid = intv2rid(ctg, rb, rb + slen);
            if (rid < 0) continue;
            int res = 0;                 // 0: new chain, 1: contained, 2: appended
            if (n_ch) {
                const int lower = kb_lower(bt, rb);
                if (lower >= 0) {        // test_and_merge, src/bwamem.c:337-359
                    ChainW &c = io.ch[lower];
                    const int64_t qend = c.last_qbeg + c.last_len, rend = c.last_rbeg + c.last_len;
                    if (rid != c.rid) res = 0;
                    else if (qbeg >= c.first_qbeg && qbeg + slen <= qend && rb >= c.pos && rb + slen <= rend) res = 1;
                    else if ((c.last_rbeg < ctg.l_pac || c.pos < ctg.l_pac) && rb >= ctg.l_pac) res = 0;
                    else {
                        const int64_t x = qbeg - c.last_qbeg, y = rb - c.last_rbeg;
                        if (y >= 0 && x - y <= P.w && y - x <= P.w && x - c.last_len < P.max_chain_gap && y - c.last_len < P.max_chain_gap) {
                            io.nxt[c.tail] = idx; io.nxt[idx] = -1;
                            io.sq[2 * idx] = qbeg; io.sq[2 * idx + 1] = slen;
                            c.tail = idx; c.last_rbeg = rb; c.last_qbeg = qbeg; c.last_len = slen; ++c.n;
                            res = 2;
                        }
                    }
                }
            }
            if (res == 0) {
                ChainW &c = io.ch[n_ch];
                c.pos = rb; c.last_rbeg = rb; c.first_qbeg = qbeg; c.last_qbeg = qbeg; c.last_len = slen; c.rid = rid;
                c.head = c.tail = idx; c.n = 1; c.w = 0; c.kept = 0; c.first = -1;
                io.nxt[idx] = -1; io.sq[2 * idx] = qbeg; io.sq[2 * idx + 1] = slen;
                if (kb_put(bt, n_ch)) return -3;
                ++n_ch;
            }
        }
        i += grp;
    }
#undef CH_GROUP
    int32_t *a = io.ord;
    int n_chn = kb_traverse(bt, a);
    if (n_chn < 0) return -3;
    const float frac_rep = (float)l_rep / io.l_query;

    // ---- mem_chain_flt, src/bwamem.c:488-560
    int n_kept = 0;
    if (n_chn > 0) {
        int kk = 0;
        for (int ii = 0; ii < n_chn; ++ii) {
            ChainW &c = io.ch[a[ii]];
            c.first = -1; c.kept = 0;
            c.w = chain_weight(io, c);
            if (c.w < P.min_chain_weight) continue;
            a[kk++] = a[ii];
        }
        n_chn = kk;
        flt_introsort(io, a, n_chn);
        if (n_chn > 0) {
            int n_k = 0, k;
            io.ch[a[0]].kept = 3;
            io.kidx[n_k++] = 0;
            for (int ii = 1; ii < n_chn; ++ii) {
                int large_ovlp = 0;
                ChainW &ci = io.ch[a[ii]];
                const int beg_i = ci.first_qbeg, end_i = ci.last_qbeg + ci.last_len;
                for (k = 0; k < n_k; ++k) {
                    ChainW &cj = io.ch[a[io.kidx[k]]];
                    const int beg_j = cj.first_qbeg, end_j = cj.last_qbeg + cj.last_len;
                    const int b_max = beg_j > beg_i ? beg_j : beg_i, e_min = end_j < end_i ? end_j : end_i;
                    if (e_min > b_max && (!ctg.alt[cj.rid] || ctg.alt[ci.rid])) {
                        const int li = end_i - beg_i, lj = end_j - beg_j, min_l = li < lj ? li : lj;
                        if ((float)(e_min - b_max) >= (float)min_l * P.mask_level && min_l < P.max_chain_gap) {
                            large_ovlp = 1;
                            if (cj.first < 0) cj.first = ii;
                            if ((float)ci.w < (float)cj.w * P.drop_ratio && cj.w - ci.w >= P.min_seed_len << 1) break;
                        }
                    }
                }
                if (k == n_k) { io.kidx[n_k++] = ii; ci.kept = large_ovlp ? 2 : 3; }
            }
            for (int ii = 0; ii < n_k; ++ii) {
                const ChainW &c = io.ch[a[io.kidx[ii]]];
                if (c.first >= 0) io.ch[a[c.first]].kept = 1;
            }
            int ii;
            for (ii = k = 0; ii < n_chn; ++ii) {
                const int kept = io.ch[a[ii]].kept;
                if (kept == 0 || kept == 3) continue;
                if (++k >= P.max_chain_extend) break;
            }
            for (; ii < n_chn; ++ii) if (io.ch[a[ii]].kept < 3) io.ch[a[ii]].kept = 0;
            for (ii = 0; ii < n_chn; ++ii) if (io.ch[a[ii]].kept != 0) a[n_kept++] = a[ii];
        }
    }
    int so = 0;
    for (int ii = 0; ii < n_kept; ++ii) {
        const ChainW &c = io.ch[a[ii]];
        bwa_b200_chain_t o;
        o.pos = c.pos; o.rid = c.rid; o.n = c.n; o.w = c.w; o.kept = c.kept; o.first = c.first; o.is_alt = ctg.alt[c.rid] ? 1 : 0;
        o.frac_rep = frac_rep; o.seed_off = so;
        io.chains[ii] = o;
        for (int j = c.head; j >= 0; j = io.nxt[j], ++so) {
            bwa_b200_chain_seed_t sd;
            sd.rbeg = (int64_t)io.rbeg[j]; sd.qbeg = io.sq[2 * j]; sd.len = io.sq[2 * j + 1]; sd.score = sd.len; sd.pad = 0;
            io.cseeds[so] = sd;
        }
    }
    return n_kept;
}

// ------------------------------------------------------------------------------------ mem_flt_chained_seeds
// src/bwamem.c:970-990: reads long enough that 5.5 ln L <= 0.05 L (757 bases at the defaults) have every seed of every kept chain
// scored by a local alignment of the read around the seed against the reference around it (mem_seed_sw, :774-808); seeds scoring below
// min_HSP_score leave their chain.  The three steps are separate so that the alignments of a read run on as many lanes as it has seeds:
// flt_seeds_applies (per read), seed_sw (per seed, independent), flt_seeds_apply (per read, sequential).
CH_FN double flt_min_l(const bwa_b200_chain_params_t &P, int l_query)
{ return P.min_chain_weight ? 1.1f * P.min_chain_weight : 5.5f * log((double)l_query); }
CH_FN bool flt_seeds_applies(const bwa_b200_chain_params_t &P, int l_query) { return !(flt_min_l(P, l_query) > 0.05f * l_query); }

constexpr int SEEDSW_MAX = 200;          // MEM_SHORT_LEN: windows of 200 bases and more are not aligned
constexpr int SEEDSW_EXT = 50;           // MEM_SHORT_EXT

CH_FN int pac_base(const uint32_t *pac, int64_t f) { return (int)((pac[f >> 4] >> (30 - ((f & 15) << 1))) & 3u); }
CH_FN int text_base(const uint32_t *pac, int64_t l_pac, int64_t p) { return p < l_pac ? pac_base(pac, p) : 3 - pac_base(pac, (l_pac << 1) - 1 - p); }
CH_FN int read_base(const uint32_t *rd, int p) { const int c = (int)((rd[p >> 3] >> (28 - ((p & 7) << 2))) & 15u); return c > 4 ? 4 : c; }

struct SeedSwQ { const uint8_t *q; size_t NS; SW_MEM int operator()(int i) const { return q[(size_t)i * NS]; } };
struct SeedSwT { const uint32_t *pac; int64_t l_pac, rb; SW_MEM int operator()(int i) const { return text_base(pac, l_pac, rb + i); } };

// mem_seed_sw: the score of ksw_align2(query window, reference window, KSW_XSTART), or -1 when the seed or a window reaches 200 bases.
// H, E: 200 int16 each, qs: 200 bytes, all at element stride NS (per-lane state of the caller).
// the two windows of mem_seed_sw (src/bwamem.c:780-800): false when the seed or a window reaches 200 bases (no alignment: -1)
CH_FN bool seed_sw_window(const Contigs &ctg, int l_query, const bwa_b200_chain_seed_t &s, int &qb, int &qe, int64_t &rb, int64_t &re)
{
    if (s.len >= SEEDSW_MAX) return false;
    qb = s.qbeg - SEEDSW_EXT; qe = s.qbeg + s.len + SEEDSW_EXT;
    rb = s.rbeg - SEEDSW_EXT; re = s.rbeg + s.len + SEEDSW_EXT;
    const int64_t mid = (s.rbeg + s.rbeg + s.len) >> 1, l2 = ctg.l_pac << 1;
    qb = qb > 0 ? qb : 0; qe = qe < l_query ? qe : l_query;
    rb = rb > 0 ? rb : 0; re = re < l2 ? re : l2;
    if (rb < ctg.l_pac && ctg.l_pac < re) { if (mid < ctg.l_pac) re = ctg.l_pac; else rb = ctg.l_pac; }
    if (qe - qb >= SEEDSW_MAX || re - rb >= SEEDSW_MAX) return false;
    {   // bns_fetch_seq, src/bntseq.c:531-552: the window is cut to the contig that holds mid
        int is_rev;
        const int rid = pos2rid(ctg, depos(ctg, mid, &is_rev));
        int64_t far_beg = ctg.off[rid], far_end = far_beg + ctg.len[rid];
        if (is_rev) { const int64_t t = far_beg; far_beg = l2 - far_end; far_end = l2 - t; }
        rb = rb > far_beg ? rb : far_beg; re = re < far_end ? re : far_end;
    }
    return true;
}
CH_DEV int seed_sw(const bwa_b200_chain_params_t &P, const Contigs &ctg, const uint32_t *pac, const uint32_t *rd, int l_query,
                   const bwa_b200_chain_seed_t &s, int16_t *H, int16_t *E, uint8_t *qs, size_t NS)
{
    int qb, qe;
    int64_t rb, re;
    if (!seed_sw_window(ctg, l_query, s, qb, qe, rb, re)) return -1;
    SwParams S;
    {   // bwa_fill_scmat, src/bwa.c:93-105
        int k = 0;
        for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) S.mat[k++] = (int8_t)(i == j ? P.a : -P.b); S.mat[k++] = -1; }
        for (int j = 0; j < 5; ++j) S.mat[k++] = -1;
    }
    S.m = 5; S.o_del = P.o_del; S.e_del = P.e_del; S.o_ins = P.o_ins; S.e_ins = P.e_ins;
    for (int k = 0; k < qe - qb; ++k) qs[(size_t)k * NS] = (uint8_t)read_base(rd, qb + k);
    return sw_i16_score(qe - qb, SeedSwQ{qs, NS}, (int)(re - rb), SeedSwT{pac, ctg.l_pac, rb}, S, H, E, NS);
}

// the filter itself, on chains whose seeds carry seed_sw's result in .score: kept seeds packed chain after chain
CH_FN void flt_seeds_apply(const bwa_b200_chain_params_t &P, int l_query, int n_chains, bwa_b200_chain_t *chains, bwa_b200_chain_seed_t *cseeds)
{
    const int min_HSP_score = (int)(P.a * flt_min_l(P, l_query) + .499);
    int so = 0;
    for (int i = 0; i < n_chains; ++i) {
        const bwa_b200_chain_seed_t *sd = cseeds + chains[i].seed_off;
        const int first = so, n = chains[i].n;
        for (int j = 0; j < n; ++j) {
            bwa_b200_chain_seed_t s = sd[j];
            if (s.score < 0 || s.score >= min_HSP_score) {
                s.score = s.score < 0 ? s.len * P.a : s.score;
                cseeds[so++] = s;
            }
        }
        chains[i].n = so - first; chains[i].seed_off = first;
    }
}

// ------------------------------------------------------------------------------------ mem_chain2aln
CH_FN int cal_max_gap(const bwa_b200_chain_params_t &P, int qlen)
{ // src/bwamem.c:996-1002
    const int l_del = (int)((double)(qlen * P.a - P.o_del) / P.e_del + 1.);
    const int l_ins = (int)((double)(qlen * P.a - P.o_ins) / P.e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < P.w << 1 ? l : P.w << 1;
}

struct AlnIO {
    int32_t l_query;
    int32_t n_chains; const bwa_b200_chain_t *chains; const bwa_b200_chain_seed_t *cseeds;
    uint64_t *srt;                 // scratch: as many entries as the read has chain seeds
    bwa_b200_region_t *regs;       // output: as many entries as the read has chain seeds
};

// Regions of the read in the order the reference creates them; n_short / n_long = jobs handed to the SHORT / LONG
// batch (fill_extension, src/bwamem.c:1102-1167).  region.job_short / job_long are indices within the read.
CH_FN int chain2aln_read(const bwa_b200_chain_params_t &P, const Contigs &ctg, const AlnIO &io, int *n_short, int *n_long)
{
    int n_regs = 0, ns = 0, nl = 0;
    const int l_query = io.l_query;
    const int64_t l_pac = ctg.l_pac;
    for (int ci = 0; ci < io.n_chains; ++ci) {
        const bwa_b200_chain_t &c = io.chains[ci];
        const bwa_b200_chain_seed_t *sd = io.cseeds + c.seed_off;
        if (c.n == 0) continue;
        int64_t rmax0 = l_pac << 1, rmax1 = 0;
        for (int i = 0; i < c.n; ++i) {            // src/bwamem.c:1180-1201
            const bwa_b200_chain_seed_t &t = sd[i];
            const int64_t b = t.rbeg - (t.qbeg + cal_max_gap(P, t.qbeg));
            const int64_t e = t.rbeg + t.len + ((l_query - t.qbeg - t.len) + cal_max_gap(P, l_query - t.qbeg - t.len));
            rmax0 = rmax0 < b ? rmax0 : b;
            rmax1 = rmax1 > e ? rmax1 : e;
        }
        rmax0 = rmax0 > 0 ? rmax0 : 0;
        rmax1 = rmax1 < l_pac << 1 ? rmax1 : l_pac << 1;
        if (rmax0 < l_pac && l_pac < rmax1) { if (sd[0].rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
        {   // bns_fetch_seq clamps the window to the contig of the chain's first seed, src/bntseq.c:531-552
            int is_rev;
            const int rid = pos2rid(ctg, depos(ctg, sd[0].rbeg, &is_rev));
            int64_t far_beg = ctg.off[rid], far_end = far_beg + ctg.len[rid];
            if (is_rev) { const int64_t tmp = far_beg; far_beg = (l_pac << 1) - far_end; far_end = (l_pac << 1) - tmp; }
            rmax0 = rmax0 > far_beg ? rmax0 : far_beg;
            rmax1 = rmax1 < far_end ? rmax1 : far_end;
        }
        const int64_t l_refer = rmax1 - rmax0;
        // seeds by (score, index) ascending; the keys are distinct, so any sort gives ks_introsort_64's order
        uint64_t *srt = io.srt + c.seed_off;
        for (int i = 0; i < c.n; ++i) srt[i] = (uint64_t)(uint32_t)sd[i].score << 32 | (uint32_t)i;
        for (int gap = c.n >> 1; gap > 0; gap >>= 1)
            for (int i = gap; i < c.n; ++i) {
                const uint64_t v = srt[i];
                int j = i;
                for (; j >= gap && srt[j - gap] > v; j -= gap) srt[j] = srt[j - gap];
                srt[j] = v;
            }
        for (int k = c.n - 1; k >= 0; --k) {
            const bwa_b200_chain_seed_t &s = sd[(uint32_t)srt[k]];
            int i;
            for (i = 0; i < n_regs; ++i) {         // src/bwamem.c:1225-1244
                const bwa_b200_region_t &p = io.regs[i];
                int64_t rd;
                int qd, w, max_gap;
                if (s.rbeg < p.rb_est || s.rbeg + s.len > p.re_est || s.qbeg < p.qb_est || s.qbeg + s.len > p.qe_est) continue;
                if (s.len - p.seedlen0 > .1 * l_query) continue;
                qd = s.qbeg - p.qb_est; rd = s.rbeg - p.rb_est;
                max_gap = cal_max_gap(P, qd < rd ? qd : (int)rd);
                w = max_gap < p.w ? max_gap : p.w;
                if (qd - rd < w && rd - qd < w) break;
                qd = p.qe_est - (s.qbeg + s.len); rd = p.re_est - (s.rbeg + s.len);
                max_gap = cal_max_gap(P, qd < rd ? qd : (int)rd);
                w = max_gap < p.w ? max_gap : p.w;
                if (qd - rd < w && rd - qd < w) break;
            }
            if (i < n_regs) {                      // src/bwamem.c:1246-1262
                for (i = k + 1; i < c.n; ++i) {
                    if (srt[i] == 0) continue;
                    const bwa_b200_chain_seed_t &t = sd[(uint32_t)srt[i]];
                    if (t.len < s.len * .95) continue;
                    if (s.qbeg <= t.qbeg && s.qbeg + s.len - t.qbeg >= s.len >> 2 && t.qbeg - s.qbeg != t.rbeg - s.rbeg) break;
                    if (t.qbeg <= s.qbeg && t.qbeg + t.len - s.qbeg >= s.len >> 2 && s.qbeg - t.qbeg != s.rbeg - t.rbeg) break;
                }
                if (i == c.n) { srt[k] = 0; continue; }
            }
            bwa_b200_region_t a;
            a.rb = a.re = 0; a.qb = a.qe = 0;
            a.w = P.w; a.score = a.truesc = -1; a.rid = c.rid;
            {   // src/bwamem.c:1284-1298 (FILTER_COEF 0.85)
                const int fwd = (int)(0.85 * (l_query - (s.qbeg + s.len)));
                a.qe_est = (s.qbeg + s.len) + fwd < l_query ? (s.qbeg + s.len) + fwd : l_query;
                a.re_est = (s.rbeg + s.len) + fwd < l_pac << 1 ? (s.rbeg + s.len) + fwd : l_pac << 1;
                const int back = (int)(0.85 * (s.qbeg + 1));
                a.qb_est = (s.qbeg - back) > 0 ? (s.qbeg - back) : 0;
                a.rb_est = (s.rbeg - back) > 0 ? (s.rbeg - back) : 0;
                if (a.rb_est < l_pac && l_pac < a.qe_est) { if (s.rbeg < l_pac) a.re_est = l_pac; else a.rb_est = l_pac; }   // sic (qe_est)
            }
            const int lq = s.qbeg, lt = (int)(s.rbeg - rmax0);
            const int rq = l_query - (lq + s.len), rt = (int)(l_refer - (lt + s.len));
            a.left_tlen = lq > 0 ? lt : 0; a.right_tlen = rq > 0 ? rt : 0;
            a.score = s.len; a.truesc = a.score;
            a.query_seed_begin = s.qbeg; a.target_seed_begin = s.rbeg;
            a.job_short = -1; a.job_long = -1;
            if (lq == 0 && rq > 0) { a.align_sides = 1; a.where_is_long = 1; a.job_long = nl++; }
            else if (lq > 0 && rq == 0) { a.align_sides = 1; a.where_is_long = 0; a.job_long = nl++; }
            else if (lq > 0 && rq > 0) {
                a.align_sides = 2;
                a.where_is_long = (s.qbeg + (s.len / 2) < l_query / 2) ? 1 : 0;
                a.job_short = ns++; a.job_long = nl++;
            } else { a.align_sides = 0; a.where_is_long = 0; a.score = a.truesc = s.score; }
            {   // seedcov, src/bwamem.c:1459-1466: taken before any extension result exists
                int64_t qb = 0, qe = 0, rb = 0, re = 0;
                if (a.align_sides == 0) { qe = l_query; rb = s.rbeg; re = s.rbeg + s.len; }
                a.seedcov = 0;
                for (i = 0; i < c.n; ++i) {
                    const bwa_b200_chain_seed_t &t = sd[i];
                    if (t.qbeg >= qb && t.qbeg + t.len <= qe && t.rbeg >= rb && t.rbeg + t.len <= re) a.seedcov += t.len;
                }
            }
            a.seedlen0 = s.len;
            a.frac_rep = c.frac_rep;
            io.regs[n_regs++] = a;
        }
    }
    *n_short = ns; *n_long = nl;
    return n_regs;
}

// result gathering, src/bwamem.c:2286-2306; sc/qe/te = (aln_score, query_batch_end, target_batch_end) of the region's jobs
CH_FN void region_finish(bwa_b200_region_t &a, int l_query, const int32_t *long3, const int32_t *short3)
{
    if (a.seedlen0 != l_query && a.align_sides > 0) {
        int32_t part[2][3] = {{0, 0, 0}, {0, 0, 0}};          // [LEFT 0 / RIGHT 1] = {score, query_end, ref_end}
        const int L = a.where_is_long ? 1 : 0;
        part[L][0] = long3[0]; part[L][1] = long3[1]; part[L][2] = long3[2];
        if (a.align_sides == 2) { part[1 - L][0] = short3[0]; part[1 - L][1] = short3[1]; part[1 - L][2] = short3[2]; }
        a.score = part[0][0] + part[1][0] - (a.align_sides == 2 ? a.seedlen0 : 0);
        a.qb = a.query_seed_begin - part[0][1];
        a.qe = a.query_seed_begin + a.seedlen0 + part[1][1];
        a.rb = a.target_seed_begin - part[0][2];
        a.re = a.target_seed_begin + a.seedlen0 + part[1][2];
        a.truesc = a.score;
    } else {
        a.qb = 0; a.qe = l_query; a.rb = a.target_seed_begin; a.re = a.target_seed_begin + a.seedlen0;
    }
}


// ------------------------------------------------------------------------------------ job sequences
// Sequences of an extension job, cut eight bases (one 4-bit packed word, first base in the high nibble) at a time:
// fill_extension's copies (src/bwamem.c:1102-1167) of the read slice and of bns_fetch_seq's window, reversed for a left
// job (src/bwamem.c:1356-1372), padded with N (code 4) to the word boundary as gasal_host_batch_fill pads to 8
// (GASAL2/src/host_batch.cpp:100-102).  The reference text is T = fwd + revcomp(fwd) over the 2-bit forward strand
// (16 bases per word, base 0 in the top two bits); a job's window never crosses l_pac (the rmax clamp above).
struct JobAux { int64_t start; int32_t qfrom; uint32_t read_flags; };   // flags: bit 31 left job, bit 30 reverse strand
constexpr uint32_t AUX_LEFT = 1u << 31, AUX_REV = 1u << 30, AUX_READ = AUX_REV - 1;

CH_FN uint32_t ld_word(const uint32_t *a, int64_t i, int64_t n) { return (i >= 0 && i < n) ? a[i] : 0u; }
CH_FN uint32_t pac8(const uint32_t *pac, int64_t n_words, int64_t f)
{   // fwd[f .. f+8) as 16 bits, fwd[f] in the top two
    const int64_t wi = f >> 4;
    const uint64_t w = (uint64_t)ld_word(pac, wi, n_words) << 32 | ld_word(pac, wi + 1, n_words);
    return (uint32_t)((w << ((f & 15) << 1)) >> 48);
}
CH_FN uint32_t read8(const uint32_t *rd, int64_t n_words, int64_t p)
{   // read[p .. p+8) as eight nibbles
    const int64_t wi = p >> 3;
    const uint64_t w = (uint64_t)ld_word(rd, wi, n_words) << 32 | ld_word(rd, wi + 1, n_words);
    return (uint32_t)((w << ((p & 7) << 2)) >> 32);
}
CH_FN uint32_t rev8x2(uint32_t x)
{   // eight 2-bit fields of the low 16 bits in reverse order
    x = ((x & 0x00ffu) << 8) | ((x >> 8) & 0x00ffu);
    x = ((x & 0x0f0fu) << 4) | ((x >> 4) & 0x0f0fu);
    return ((x & 0x3333u) << 2) | ((x >> 2) & 0x3333u);
}
CH_FN uint32_t rev8x4(uint32_t x)
{   // eight nibbles in reverse order
    x = (x << 16) | (x >> 16);
    x = ((x & 0x00ff00ffu) << 8) | ((x >> 8) & 0x00ff00ffu);
    return ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);
}
CH_FN uint32_t spread2to4(uint32_t x)
{   // 8 x 2 bits -> 8 nibbles, order kept
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    return (x | (x << 2)) & 0x33333333u;
}
CH_FN uint32_t pad_n(uint32_t w, int64_t valid)
{   // the first `valid` nibbles of w, N in the others
    if (valid >= 8) return w;
    const uint32_t m = valid <= 0 ? 0u : 0xffffffffu << ((8 - (int)valid) << 2);
    return (w & m) | (0x44444444u & ~m);
}
CH_FN uint32_t cut_target_word(const uint32_t *pac, int64_t pac_words, int64_t l_pac, const JobAux &a, uint32_t wi, uint32_t tlen)
{
    const int64_t i0 = (int64_t)wi << 3;
    const bool left = a.read_flags & AUX_LEFT, rev = a.read_flags & AUX_REV;
    uint32_t x;
    if (!left) x = !rev ? pac8(pac, pac_words, a.start + i0) : rev8x2(pac8(pac, pac_words, (l_pac << 1) - 8 - (a.start + i0))) ^ 0xffffu;
    else x = !rev ? rev8x2(pac8(pac, pac_words, a.start - 8 - i0)) : pac8(pac, pac_words, (l_pac << 1) - a.start + i0) ^ 0xffffu;
    return pad_n(spread2to4(x), (int64_t)tlen - i0);
}
CH_FN uint32_t cut_query_word(const uint32_t *rd, int64_t rd_words, const JobAux &a, uint32_t wi, uint32_t qlen)
{
    const int64_t i0 = (int64_t)wi << 3;
    const uint32_t x = !(a.read_flags & AUX_LEFT) ? read8(rd, rd_words, a.qfrom + i0) : rev8x4(read8(rd, rd_words, a.qfrom - 8 - i0));
    return pad_n(x, (int64_t)qlen - i0);
}

// the two jobs of a region: side 0 = left, 1 = right -> {qlen, tlen}; which one is the LONG batch's: where_is_long
CH_FN void region_job(const bwa_b200_region_t &a, int l_query, int side, uint32_t *qlen, uint32_t *tlen, JobAux *aux, uint32_t read, int64_t l_pac)
{
    const uint32_t rev = a.target_seed_begin >= l_pac ? AUX_REV : 0u;
    if (side == 0) { *qlen = (uint32_t)a.query_seed_begin; *tlen = (uint32_t)a.left_tlen; aux->start = a.target_seed_begin; aux->qfrom = a.query_seed_begin; aux->read_flags = read | AUX_LEFT | rev; }
    else { *qlen = (uint32_t)(l_query - a.query_seed_begin - a.seedlen0); *tlen = (uint32_t)a.right_tlen; aux->start = a.target_seed_begin + a.seedlen0; aux->qfrom = a.query_seed_begin + a.seedlen0; aux->read_flags = read | rev; }
}

// every job of a read in the order fill_extension receives them: emit(region index, is_long, qlen, tlen, h0, aux)
template <class F>
CH_FN void read_jobs(const bwa_b200_region_t *regs, int n_regs, int l_query, uint32_t read, int64_t l_pac, F &&emit)
{
    for (int i = 0; i < n_regs; ++i) {
        const bwa_b200_region_t &a = regs[i];
        if (a.align_sides <= 0) continue;
        uint32_t ql, tl;
        JobAux aux;
        const int long_side = a.where_is_long ? 1 : 0;
        if (a.align_sides == 2) { region_job(a, l_query, 1 - long_side, &ql, &tl, &aux, read, l_pac); emit(i, 0, ql, tl, (uint32_t)a.seedlen0, aux); }
        region_job(a, l_query, long_side, &ql, &tl, &aux, read, l_pac);
        emit(i, 1, ql, tl, (uint32_t)a.seedlen0, aux);
    }
}

// (aln_score, query_batch_end, target_batch_end) of a job: the local-vs-to-end rule, src/bwamem.c:1892-1901
CH_FN void ext_triple(const bwa_b200_ext_result_t &r, int qlen, int pen_clip, int32_t out[3])
{
    if (r.gscore <= 0 || r.gscore <= r.score - pen_clip) { out[0] = r.score; out[1] = r.qle; out[2] = r.tle; }
    else { out[0] = r.gscore; out[1] = qlen; out[2] = r.gtle; }
}

} // namespace b200chain
