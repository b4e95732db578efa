/*
 * oracle/fmd_oracle.c -- TEST INFRASTRUCTURE ONLY.
 * CPU restatement of the reference's SMEM seeding path over the FMD index:
 *   bwt_occ / bwt_occ4 / bwt_2occ4   src/bwt.c:235-403
 *   bwt_extend                       src/bwt.c:455-470
 *   bwt_smem1a (max_intv == 0)       src/bwt.c:483-566
 *   bwt_sa / bwt_invPsi              bwa_index/bwt.c:54-60,151-172 (u32 + packed hi-bit SA)
 *   mem_collect_intv pass 1          bwa_index/bwamem.c:121-131
 *   occurrence sampling rule         bwa_index/bwamem.c:278-283
 * written against the reference's GPU index layout (32-byte buckets of 64 symbols,
 * bwa_index/bwtindex.c:174-197).  Parity is pinned by tests/test_oracle_vs_ref.py
 * (against oracle/_ref built from the reference sources) and tests/golden/.
 */
#include "fmd_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define NONE64 ((uint64_t)-1)

/* ---------------------------------------------------------------- loaders */

int fmd_load(fmd_index_t *idx, const char *bwt_path, const char *sa_path)
{
    memset(idx, 0, sizeof(*idx));
    FILE *fp = fopen(bwt_path, "rb");
    if (!fp) return -1;
    fseek(fp, 0, SEEK_END);
    long fsz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    /* header: u64 primary, u64 L2[1..4]  (bwa_index/bwt.c:461-470) */
    idx->n_words = (uint64_t)(fsz - 5 * 8) >> 2;
    idx->bwt = (uint32_t *)malloc(idx->n_words * 4 + 64);
    int ok = fread(&idx->primary, 8, 1, fp) == 1 && fread(idx->L2 + 1, 8, 4, fp) == 4 &&
             fread(idx->bwt, 4, idx->n_words, fp) == idx->n_words;
    fclose(fp);
    if (!ok) return -2;
    idx->L2[0] = 0;
    idx->seq_len = idx->L2[4];
    idx->owns = 1;
    if (!sa_path) return 0;
    fp = fopen(sa_path, "rb");
    if (!fp) return -3;
    /* u64 primary, 4 x u64 (skipped), 8-byte sa_intv, u64 seq_len, (n_sa-1) x u32,
     * u8 pack_size, (pack_size*n_sa/32+1) x u32   (bwa_index/bwt.c:472-487) */
    uint64_t primary, skip[4], intv, seq_len;
    ok = fread(&primary, 8, 1, fp) == 1 && fread(skip, 8, 4, fp) == 4 &&
         fread(&intv, 8, 1, fp) == 1 && fread(&seq_len, 8, 1, fp) == 1;
    if (!ok || primary != idx->primary || seq_len != idx->seq_len) { fclose(fp); return -4; }
    idx->sa_intv = (int)intv;
    idx->n_sa = (idx->seq_len + intv) / intv;
    idx->sa = (uint32_t *)malloc(idx->n_sa * 4);
    idx->sa[0] = (uint32_t)-1;
    uint8_t ps = 0;
    ok = fread(idx->sa + 1, 4, idx->n_sa - 1, fp) == idx->n_sa - 1 && fread(&ps, 1, 1, fp) == 1;
    if (!ok) { fclose(fp); return -5; }
    idx->pack_size = ps;
    uint64_t nhi = (uint64_t)ps * idx->n_sa / 32 + 1;
    idx->sa_hi = (uint32_t *)calloc(nhi, 4);
    size_t got = fread(idx->sa_hi, 4, nhi, fp);
    (void)got; /* bwt_dump_sa writes pack_size*(n_sa/32)+1 words, which can be fewer */
    fclose(fp);
    return 0;
}

void fmd_free(fmd_index_t *idx)
{
    if (idx->owns) { free(idx->bwt); free(idx->sa); free(idx->sa_hi); }
    memset(idx, 0, sizeof(*idx));
}

/* ------------------------------------------------------------ occ lookups */

/* count occurrences of base c among the first n (0..16) symbols of a packed word;
 * symbol i sits at bits (15-i)*2 (bwa_index/bwtindex.c:149 bwt_B00). */
static inline uint32_t word_count(uint32_t w, int n, int c)
{
    if (n <= 0) return 0;
    uint32_t x = n >= 16 ? w : w >> (32 - 2 * n);
    uint32_t m = n >= 16 ? 0x55555555u : (((1u << (2 * n)) - 1) & 0x55555555u);
    uint32_t t = ~(x ^ ((uint32_t)c * 0x55555555u));
    return (uint32_t)__builtin_popcount(t & (t >> 1) & m);
}

static inline const uint32_t *bucket_of(const fmd_index_t *idx, uint64_t kk) { return idx->bwt + (kk >> 6) * 8; }

/* Occ for all four bases over BWT rows [0,k]; k == -1 gives zeros (src/bwt.c:340-360). */
void fmd_occ4(const fmd_index_t *idx, uint64_t k, uint64_t cnt[4], fmd_counters_t *c)
{
    if (k == NONE64) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
    k -= (k >= idx->primary);               /* '$' is not stored */
    const uint32_t *b = bucket_of(idx, k);
    int n = (int)(k & 63) + 1;               /* symbols of this bucket to include */
    if (c) c->n_bucket++;
    for (int a = 0; a < 4; ++a) {
        uint32_t x = b[a];
        for (int w = 0; w < 4; ++w) x += word_count(b[4 + w], n - 16 * w, a);
        cnt[a] = x;
    }
}

/* Occ of one base; handles the k == seq_len and k == -1 cases of src/bwt.c:235-262. */
uint64_t fmd_occ(const fmd_index_t *idx, uint64_t k, int base, fmd_counters_t *c)
{
    if (k == idx->seq_len) return idx->L2[base + 1] - idx->L2[base];
    if (k == NONE64) return 0;
    k -= (k >= idx->primary);
    const uint32_t *b = bucket_of(idx, k);
    int n = (int)(k & 63) + 1;
    if (c) c->n_bucket++;
    uint32_t x = b[base];
    for (int w = 0; w < 4; ++w) x += word_count(b[4 + w], n - 16 * w, base);
    return x;
}

/* same answers as two fmd_occ4 calls; only the bucket-touch counter knows about the
 * shared-bucket fast path of src/bwt.c:363-403. */
static void occ4_pair(const fmd_index_t *idx, uint64_t k, uint64_t l, uint64_t ck[4], uint64_t cl[4],
                      fmd_counters_t *c)
{
    fmd_occ4(idx, k, ck, c);
    fmd_occ4(idx, l, cl, c);
    if (c && k != NONE64 && l != NONE64) {
        uint64_t kk = k - (k >= idx->primary), ll = l - (l >= idx->primary);
        if ((kk >> 6) == (ll >> 6)) c->n_bucket--;
    }
}

/* ---------------------------------------------------------------- extend */

void fmd_extend(const fmd_index_t *idx, const fmd_intv_t *ik, fmd_intv_t ok[4], int is_back,
                fmd_counters_t *c)
{
    /* the side being extended is x[!is_back] in the reference: backward extension walks
     * x[0] (k), forward extension walks x[1] (l) with the complemented base. */
    uint64_t side = is_back ? ik->k : ik->l, other = is_back ? ik->l : ik->k;
    uint64_t tk[4], tl[4], nside[4], nother[4];
    occ4_pair(idx, side - 1, side - 1 + ik->s, tk, tl, c);
    if (c) c->n_extend++;
    for (int a = 0; a < 4; ++a) {
        nside[a] = idx->L2[a] + 1 + tk[a];
        ok[a].s = tl[a] - tk[a];
    }
    nother[3] = other + (side <= idx->primary && side + ik->s - 1 >= idx->primary);
    nother[2] = nother[3] + ok[3].s;
    nother[1] = nother[2] + ok[2].s;
    nother[0] = nother[1] + ok[1].s;
    for (int a = 0; a < 4; ++a) {
        if (is_back) { ok[a].k = nside[a]; ok[a].l = nother[a]; }
        else         { ok[a].l = nside[a]; ok[a].k = nother[a]; }
        ok[a].beg = ik->beg; ok[a].end = ik->end;
    }
}

static inline void set_intv(const fmd_index_t *idx, int c, fmd_intv_t *ik)
{
    ik->k = idx->L2[c] + 1;
    ik->s = idx->L2[c + 1] - idx->L2[c];
    ik->l = idx->L2[3 - c] + 1;
    ik->beg = 0; ik->end = 0;
}

/* ----------------------------------------------------------------- smem1 */

int fmd_smem1(const fmd_index_t *idx, int len, const uint8_t *q, int x, int min_intv,
              fmd_intv_t *out, int *n_out, fmd_counters_t *c)
{
    *n_out = 0;
    if (q[x] > 3) return x + 1;
    if (min_intv < 1) min_intv = 1;
    fmd_intv_t *lst0 = (fmd_intv_t *)malloc(sizeof(fmd_intv_t) * (size_t)(len + 2) * 2);
    fmd_intv_t *prev = lst0, *curr = lst0 + (len + 2), *swp;
    int n_prev = 0, n_curr = 0, i, j;
    fmd_intv_t ik, ok[4];

    /* forward phase: remember the interval each time its size changes */
    set_intv(idx, q[x], &ik);
    ik.end = x + 1;
    for (i = x + 1; i < len; ++i) {
        if (q[i] < 4) {
            int cb = 3 - q[i];
            uint64_t b_before = c ? c->n_bucket : 0;
            fmd_extend(idx, &ik, ok, 0, c);
            if (c) { c->n_extend_fwd++; c->n_bucket_fwd += c->n_bucket - b_before; }
            if (ok[cb].s != ik.s) {
                curr[n_curr++] = ik;
                if (ok[cb].s < (uint64_t)min_intv) break;
            }
            ik = ok[cb];
            ik.end = i + 1;
        } else {
            curr[n_curr++] = ik;
            break;
        }
    }
    if (i == len) curr[n_curr++] = ik;
    /* longest match first */
    for (j = 0; j < n_curr / 2; ++j) { fmd_intv_t t = curr[j]; curr[j] = curr[n_curr - 1 - j]; curr[n_curr - 1 - j] = t; }
    int ret = curr[0].end;
    swp = curr; curr = prev; prev = swp; n_prev = n_curr;

    /* backward phase */
    int n_mem = 0;
    for (i = x - 1; i >= -1; --i) {
        int cb = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
        n_curr = 0;
        for (j = 0; j < n_prev; ++j) {
            fmd_intv_t *p = &prev[j];
            if (cb >= 0) fmd_extend(idx, p, ok, 1, c);
            if (cb < 0 || ok[cb].s < (uint64_t)min_intv) {
                if (n_curr == 0) {
                    if (n_mem == 0 || i + 1 < out[n_mem - 1].beg) {
                        out[n_mem] = *p;
                        out[n_mem].beg = i + 1;
                        ++n_mem;
                    }
                }
            } else if (n_curr == 0 || ok[cb].s != curr[n_curr - 1].s) {
                ok[cb].end = p->end;
                curr[n_curr++] = ok[cb];
            }
        }
        if (n_curr == 0) break;
        swp = curr; curr = prev; prev = swp; n_prev = n_curr;
    }
    /* ascending start */
    for (j = 0; j < n_mem / 2; ++j) { fmd_intv_t t = out[j]; out[j] = out[n_mem - 1 - j]; out[n_mem - 1 - j] = t; }
    *n_out = n_mem;
    free(lst0);
    return ret;
}

int fmd_collect_pass1(const fmd_index_t *idx, int len, const uint8_t *q, int min_seed_len,
                      fmd_intv_t *out, fmd_counters_t *c)
{
    int n = 0, x = 0, m, i;
    fmd_intv_t *tmp = (fmd_intv_t *)malloc(sizeof(fmd_intv_t) * (size_t)(len + 2));
    while (x < len) {
        if (q[x] < 4) {
            x = fmd_smem1(idx, len, q, x, 1, tmp, &m, c);
            for (i = 0; i < m; ++i)
                if (tmp[i].end - tmp[i].beg >= min_seed_len) out[n++] = tmp[i];
        } else ++x;
    }
    free(tmp);
    if (c) c->n_smem += (uint64_t)n;
    return n;
}

/* bwa_index/bwt.c:434-455 bwt_seed_strategy1: forward extension from x until the interval is smaller than max_intv
 * and the match longer than min_len; *m gets s == 0 when nothing is found.  Returns the next x. */
int fmd_seed_strategy1(const fmd_index_t *idx, int len, const uint8_t *q, int x, int min_len, int max_intv,
                       fmd_intv_t *m, fmd_counters_t *c)
{
    fmd_intv_t ik, ok[4];
    int i;
    memset(m, 0, sizeof(*m));
    if (q[x] > 3) return x + 1;
    set_intv(idx, q[x], &ik);
    for (i = x + 1; i < len; ++i) {
        if (q[i] < 4) {
            int cb = 3 - q[i];
            uint64_t b_before = c ? c->n_bucket : 0;
            fmd_extend(idx, &ik, ok, 0, c);
            if (c) { c->n_extend_fwd++; c->n_bucket_fwd += c->n_bucket - b_before; }
            if (ok[cb].s < (uint64_t)max_intv && i - x >= min_len) {
                *m = ok[cb];
                m->beg = x; m->end = i + 1;
                return i + 1;
            }
            ik = ok[cb];
        } else return i + 1;
    }
    return len;
}

static int intv_cmp(const void *pa, const void *pb)
{ /* intv_lt: by info = start << 32 | end (bwa_index/bwamem.c:82-83); equal info means equal interval */
    const fmd_intv_t *a = (const fmd_intv_t *)pa, *b = (const fmd_intv_t *)pb;
    if (a->beg != b->beg) return a->beg < b->beg ? -1 : 1;
    if (a->end != b->end) return a->end < b->end ? -1 : 1;
    return 0;
}

/* bwa_index/bwamem.c:114-162 mem_collect_intv, all three passes and the final sort. */
int fmd_collect_intv(const fmd_index_t *idx, int len, const uint8_t *q, int min_seed_len, const fmd_reseed_t *rs,
                     fmd_intv_t **out, size_t *out_cap, fmd_counters_t *c)
{
    size_t n = 0;
    int x = 0, m, i;
    fmd_intv_t *tmp = (fmd_intv_t *)malloc(sizeof(fmd_intv_t) * (size_t)(len + 2));
#define PUSH(v) do { if (n == *out_cap) { *out_cap = *out_cap ? *out_cap * 2 : 64; *out = (fmd_intv_t *)realloc(*out, *out_cap * sizeof(fmd_intv_t)); } (*out)[n++] = (v); } while (0)
    while (x < len) {                                        /* pass 1: all SMEMs */
        if (q[x] < 4) {
            x = fmd_smem1(idx, len, q, x, 1, tmp, &m, c);
            for (i = 0; i < m; ++i)
                if (tmp[i].end - tmp[i].beg >= min_seed_len) PUSH(tmp[i]);
        } else ++x;
    }
    if (rs && rs->enable) {
        const int split_len = (int)(min_seed_len * rs->split_factor + .499);
        const size_t old_n = n;
        for (size_t k = 0; k < old_n; ++k) {                 /* pass 2: MEMs inside a long, rare SMEM */
            const fmd_intv_t p = (*out)[k];
            if (p.end - p.beg < split_len || p.s > (uint64_t)rs->split_width) continue;
            fmd_smem1(idx, len, q, (p.beg + p.end) >> 1, (int)p.s + 1, tmp, &m, c);
            for (i = 0; i < m; ++i)
                if (tmp[i].end - tmp[i].beg >= min_seed_len) PUSH(tmp[i]);
        }
        if (rs->max_mem_intv > 0) {                          /* pass 3: LAST-like */
            x = 0;
            while (x < len) {
                if (q[x] < 4) {
                    fmd_intv_t mm;
                    x = fmd_seed_strategy1(idx, len, q, x, min_seed_len, rs->max_mem_intv, &mm, c);
                    if (mm.s > 0) PUSH(mm);
                } else ++x;
            }
        }
        qsort(*out, n, sizeof(fmd_intv_t), intv_cmp);
    }
#undef PUSH
    free(tmp);
    if (c) c->n_smem += (uint64_t)n;
    return (int)n;
}

/* -------------------------------------------------------------------- SA */

static inline int bwt_sym(const fmd_index_t *idx, uint64_t j) /* j: index in the '$'-less string */
{
    const uint32_t *b = bucket_of(idx, j);
    return (int)(b[4 + ((j & 63) >> 4)] >> ((~j & 15) << 1) & 3);
}

static inline uint64_t inv_psi(const fmd_index_t *idx, uint64_t k, fmd_counters_t *c)
{
    if (k == idx->primary) return 0;
    uint64_t j = k - (k > idx->primary);
    int a = bwt_sym(idx, j);
    if (c) c->n_lf++;
    fmd_counters_t *cc = NULL; /* the symbol and its occ live in the same bucket: one touch */
    uint64_t r = idx->L2[a] + fmd_occ(idx, k, a, cc);
    if (c) c->n_bucket++;
    return r;
}

uint64_t fmd_sa(const fmd_index_t *idx, uint64_t k, fmd_counters_t *c)
{
    uint64_t steps = 0, mask = (uint64_t)idx->sa_intv - 1;
    while (k & mask) { ++steps; k = inv_psi(idx, k, c); }
    if (c) c->n_located++;
    uint64_t j = k / (uint64_t)idx->sa_intv;
    if (j == 0) return steps - 1;                         /* sa[0] == -1 */
    uint64_t hi = 0;
    if (idx->sa_hi && idx->pack_size > 0) {
        uint32_t per = 32u / (uint32_t)idx->pack_size;
        uint32_t msk = idx->pack_size >= 32 ? 0xffffffffu : ((1u << idx->pack_size) - 1);
        hi = (idx->sa_hi[j / per] >> ((j % per) * (uint32_t)idx->pack_size)) & msk;
    }
    return steps + ((uint64_t)idx->sa[j] | hi << 32);
}

/* --------------------------------------------------------------- batches */

typedef struct { uint64_t rbeg; int32_t qbeg, qend; uint32_t score; } seed_rec_t;
typedef struct { seed_rec_t *a; size_t n, m; } seed_vec_t;
static inline void sv_push(seed_vec_t *v, seed_rec_t r)
{
    if (v->n == v->m) { v->m = v->m ? v->m * 2 : 1024; v->a = (seed_rec_t *)realloc(v->a, v->m * sizeof(seed_rec_t)); }
    v->a[v->n++] = r;
}

int64_t fmd_seed_batch(const fmd_index_t *idx, const uint8_t *reads, const uint64_t *read_off,
                       int64_t n_reads, int min_seed_len, int max_occ,
                       uint32_t *n_seeds, uint64_t *seed_off,
                       uint64_t *rbeg, int32_t *qbeg, int32_t *qend, uint32_t *score,
                       int64_t cap, int n_threads, fmd_counters_t *cnt)
{
    return fmd_seed_batch_rs(idx, reads, read_off, n_reads, min_seed_len, max_occ, NULL, n_seeds, seed_off, rbeg, qbeg, qend, score,
                             cap, n_threads, cnt);
}

int64_t fmd_seed_batch_rs(const fmd_index_t *idx, const uint8_t *reads, const uint64_t *read_off,
                          int64_t n_reads, int min_seed_len, int max_occ, const fmd_reseed_t *rs,
                          uint32_t *n_seeds, uint64_t *seed_off,
                          uint64_t *rbeg, int32_t *qbeg, int32_t *qend, uint32_t *score,
                          int64_t cap, int n_threads, fmd_counters_t *cnt)
{
    if (n_threads < 1) n_threads = 1;
    seed_vec_t *tv = (seed_vec_t *)calloc((size_t)n_threads, sizeof(seed_vec_t));
    fmd_counters_t *tc = (fmd_counters_t *)calloc((size_t)n_threads, sizeof(fmd_counters_t));
    uint64_t *where = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(n_reads ? n_reads : 1));
    uint8_t *who = (uint8_t *)malloc((size_t)(n_reads ? n_reads : 1) * sizeof(uint16_t));
    uint16_t *who16 = (uint16_t *)who;
#pragma omp parallel num_threads(n_threads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        fmd_intv_t *mem = NULL; size_t mem_cap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_reads; ++r) {
            int len = (int)(read_off[r + 1] - read_off[r]);
            int n = len >= min_seed_len ? fmd_collect_intv(idx, len, reads + read_off[r], min_seed_len, rs, &mem, &mem_cap, &tc[tid]) : 0;
            where[r] = tv[tid].n; who16[r] = (uint16_t)tid;
            uint32_t ns = 0;
            for (int i = 0; i < n; ++i) {
                uint64_t s = mem[i].s, step = 1, count = s;
                if (max_occ > 0) { step = s > (uint64_t)max_occ ? s / (uint64_t)max_occ : 1; count = (s + step - 1) / step; if (count > (uint64_t)max_occ) count = (uint64_t)max_occ; }
                for (uint64_t t = 0; t < count; ++t) {
                    seed_rec_t rec;
                    rec.rbeg = fmd_sa(idx, mem[i].k + t * step, &tc[tid]);
                    rec.qbeg = mem[i].beg; rec.qend = mem[i].end;
                    rec.score = t == 0 ? (uint32_t)s : 0;
                    sv_push(&tv[tid], rec);
                    ++ns;
                }
            }
            n_seeds[r] = ns;
        }
        free(mem);
    }
    uint64_t tot = 0;
    for (int64_t r = 0; r < n_reads; ++r) { seed_off[r] = tot; tot += n_seeds[r]; }
    int64_t ret = (int64_t)tot;
    if ((int64_t)tot > cap) ret = -1;
    else {
#pragma omp parallel for num_threads(n_threads) schedule(static)
        for (int64_t r = 0; r < n_reads; ++r) {
            const seed_rec_t *src = tv[who16[r]].a + where[r];
            uint64_t o = seed_off[r];
            for (uint32_t i = 0; i < n_seeds[r]; ++i) {
                rbeg[o + i] = src[i].rbeg; qbeg[o + i] = src[i].qbeg; qend[o + i] = src[i].qend; score[o + i] = src[i].score;
            }
        }
    }
    for (int t = 0; t < n_threads; ++t) {
        if (cnt) { cnt->n_extend += tc[t].n_extend; cnt->n_bucket += tc[t].n_bucket; cnt->n_lf += tc[t].n_lf; cnt->n_located += tc[t].n_located; cnt->n_smem += tc[t].n_smem; cnt->n_extend_fwd += tc[t].n_extend_fwd; cnt->n_bucket_fwd += tc[t].n_bucket_fwd; }
        free(tv[t].a);
    }
    free(tv); free(tc); free(where); free(who);
    return ret;
}

int64_t fmd_smem_batch(const fmd_index_t *idx, const uint8_t *reads, const uint64_t *read_off,
                       int64_t n_reads, int min_seed_len,
                       uint32_t *n_smems, int32_t *qbeg, int32_t *qend, uint64_t *k, uint64_t *s,
                       int64_t cap, int n_threads, fmd_counters_t *cnt)
{
    return fmd_smem_batch_rs(idx, reads, read_off, n_reads, min_seed_len, NULL, n_smems, qbeg, qend, k, s, cap, n_threads, cnt);
}

int64_t fmd_smem_batch_rs(const fmd_index_t *idx, const uint8_t *reads, const uint64_t *read_off,
                          int64_t n_reads, int min_seed_len, const fmd_reseed_t *rs,
                          uint32_t *n_smems, int32_t *qbeg, int32_t *qend, uint64_t *k, uint64_t *s,
                          int64_t cap, int n_threads, fmd_counters_t *cnt)
{
    /* serial per-read fill; offsets are the running sum (small inputs only) */
    (void)n_threads;
    int64_t tot = 0;
    fmd_intv_t *mem = NULL; size_t mem_cap = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        int len = (int)(read_off[r + 1] - read_off[r]);
        int n = len >= min_seed_len ? fmd_collect_intv(idx, len, reads + read_off[r], min_seed_len, rs, &mem, &mem_cap, cnt) : 0;
        n_smems[r] = (uint32_t)n;
        if (tot + n > cap) { free(mem); return -1; }
        for (int i = 0; i < n; ++i) {
            qbeg[tot] = mem[i].beg; qend[tot] = mem[i].end; k[tot] = mem[i].k; s[tot] = mem[i].s; ++tot;
        }
    }
    free(mem);
    return tot;
}
