/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE ONLY.
 * Thin C entry points around the UNMODIFIED reference CPU functions, compiled by
 * oracle/build_ref.sh together with the reference's own bwa_index/ *.c files (taken from
 * /root/reference where they lie; OCC_INTV_SHIFT set to 7 exactly as the reference's
 * build_index.sh:46 does) into oracle/_ref/libbwaref.so.  This file contains no reference
 * code: it only calls bwt_restore_bwt / bwt_smem1 / bwt_sa / ksw_extend2 through the
 * reference headers.  The one thing it replaces is the .sa reader: the reference's
 * bwt_restore_sa (bwa_index/bwt.c:501-527) cannot read the u32 + packed-hi-bit file that
 * bwt_dump_sa (bwa_index/bwt.c:472-487) writes, so the samples are
 * loaded here into the same bwt_t fields.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "bwt.h"
#include "ksw.h"
#include "bwa.h"
#include "jobs_common.h"

void *ref_load(const char *bwt128_path, const char *sa_path)
{
    bwt_t *bwt = bwt_restore_bwt(bwt128_path);
    if (!sa_path) return bwt;
    FILE *fp = fopen(sa_path, "rb");
    if (!fp) return NULL;
    uint64_t hdr[7];
    if (fread(hdr, 8, 7, fp) != 7 || hdr[0] != bwt->primary || hdr[6] != bwt->seq_len) { fclose(fp); return NULL; }
    bwt->sa_intv = (int)hdr[5];
    bwt->n_sa = (bwt->seq_len + bwt->sa_intv) / bwt->sa_intv;
    bwt->sa = (uint32_t *)calloc(bwt->n_sa, 4);
    bwt->sa[0] = (uint32_t)-1;
    if (fread(bwt->sa + 1, 4, bwt->n_sa - 1, fp) != bwt->n_sa - 1) { fclose(fp); return NULL; }
    uint8_t ps;
    if (fread(&ps, 1, 1, fp) != 1) { fclose(fp); return NULL; }
    bwt->pack_size = ps;
    bwt->pack_mask = ps >= 32 ? 0xffffffffu : ((1u << ps) - 1);
    if ((bwt->seq_len >> 32) == 0) bwt->pack_mask = 0; /* bwa_index/bwt.c:88-91: msb == 0 */
    size_t nhi = (size_t)ps * bwt->n_sa / 32 + 1;
    bwt->sa_bits = (uint32_t *)calloc(nhi + 1, 4);
    size_t got = fread(bwt->sa_bits, 4, nhi, fp); (void)got;
    fclose(fp);
    return bwt;
}

void ref_free(void *h) { if (h) bwt_destroy((bwt_t *)h); }

uint64_t ref_seq_len(void *h) { return ((bwt_t *)h)->seq_len; }
uint64_t ref_primary(void *h) { return ((bwt_t *)h)->primary; }

/* out: 5 x u64 per interval = x0, x1, x2, start, end */
int ref_smem1(void *h, int len, const uint8_t *q, int x, int min_intv, uint64_t *out, int *n_out)
{
    bwtintv_v mem = {0, 0, 0};
    int ret = bwt_smem1((bwt_t *)h, len, q, x, min_intv, &mem, 0);
    for (size_t i = 0; i < mem.n; ++i) {
        out[5 * i + 0] = mem.a[i].x[0]; out[5 * i + 1] = mem.a[i].x[1]; out[5 * i + 2] = mem.a[i].x[2];
        out[5 * i + 3] = mem.a[i].info >> 32; out[5 * i + 4] = (uint32_t)mem.a[i].info;
    }
    *n_out = (int)mem.n;
    free(mem.a);
    return ret;
}

uint64_t ref_sa(void *h, uint64_t k) { return bwt_sa((bwt_t *)h, k); }

void ref_occ4(void *h, uint64_t k, uint64_t cnt[4]) { bwt_occ4((bwt_t *)h, k, cnt); }

void ref_extend(void *h, const uint64_t ik[3], uint64_t ok[12], int is_back)
{
    bwtintv_t i, o[4];
    i.x[0] = ik[0]; i.x[1] = ik[1]; i.x[2] = ik[2]; i.info = 0;
    bwt_extend((bwt_t *)h, &i, o, is_back);
    for (int a = 0; a < 4; ++a) { ok[3 * a] = o[a].x[0]; ok[3 * a + 1] = o[a].x[1]; ok[3 * a + 2] = o[a].x[2]; }
}

/* stock (always banded) ksw_extend2, bwa_index/ksw.c:380 */
int ref_ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0,
                    int32_t out[6])
{
    int qle, tle, gtle, gscore, max_off;
    int sc = ksw_extend2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, h0,
                         &qle, &tle, &gtle, &gscore, &max_off);
    out[0] = sc; out[1] = qle; out[2] = tle; out[3] = gtle; out[4] = gscore; out[5] = max_off;
    return sc;
}

/* ---- whole-batch drivers used as the CPU "reference" arm of bench.py ---- */

/* pass-1 SMEMs (len >= min_seed_len) + sampled SA lookups per read, like
 * mem_collect_intv pass 1 + the mem_chain loop (bwa_index/bwamem.c:121-131,278-283).
 * Returns the number of seeds produced; a checksum keeps the work observable. */
int64_t ref_seed_batch(void *h, const uint8_t *reads, const uint64_t *read_off, int64_t n_reads,
                       int min_seed_len, int max_occ, int n_threads, uint64_t *checksum)
{
    bwt_t *bwt = (bwt_t *)h;
    int64_t total = 0; uint64_t sum = 0;
#pragma omp parallel num_threads(n_threads) reduction(+:total,sum)
    {
        bwtintv_v mem = {0, 0, 0}, t0 = {0, 0, 0}, t1 = {0, 0, 0};
        bwtintv_v *tmpv[2] = {&t0, &t1};
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_reads; ++r) {
            const uint8_t *q = reads + read_off[r];
            int len = (int)(read_off[r + 1] - read_off[r]), x = 0;
            if (len < min_seed_len) continue;
            while (x < len) {
                if (q[x] < 4) {
                    x = bwt_smem1(bwt, len, q, x, 1, &mem, tmpv);
                    for (size_t i = 0; i < mem.n; ++i) {
                        bwtintv_t *p = &mem.a[i];
                        int slen = (int)((uint32_t)p->info - (p->info >> 32));
                        if (slen < min_seed_len) continue;
                        int64_t step = p->x[2] > (uint64_t)max_occ ? p->x[2] / max_occ : 1, k, count;
                        for (k = count = 0; k < (int64_t)p->x[2] && count < max_occ; k += step, ++count) {
                            uint64_t pos = bwt_sa(bwt, p->x[0] + k);
                            sum += pos ^ ((uint64_t)slen << 40);
                            ++total;
                        }
                    }
                } else ++x;
            }
        }
        free(mem.a); free(t0.a); free(t1.a);
    }
    if (checksum) *checksum = sum;
    return total;
}

/* the seeds themselves, in the layout of mem_seed_v_gpu (seed_gen.h:68-75): pass-1 SMEMs of length >= min_seed_len in query
 * order, rows k + t*step for t < count (max_occ <= 0: all rows), `score` = occurrence count on the first row of a group.
 * n_seeds[r] / seed_off[r] per read; returns the total, or -(total) when `cap` was too small (nothing past cap written). */
typedef struct { uint64_t rbeg; int32_t qb, qe; uint32_t score; } ref_seed_rec_t;
int64_t ref_seed_arrays(void *h, const uint8_t *reads, const uint64_t *read_off, int64_t n_reads, int min_seed_len, int max_occ, int n_threads,
                        uint32_t *n_seeds, uint64_t *seed_off, uint64_t *rbeg, int32_t *qbeg_qend, uint32_t *score, int64_t cap)
{
    bwt_t *bwt = (bwt_t *)h;
    if (n_threads < 1) n_threads = 1;
    ref_seed_rec_t **tv = (ref_seed_rec_t **)calloc((size_t)n_threads, sizeof(*tv));
    size_t *tn = (size_t *)calloc((size_t)n_threads, sizeof(size_t)), *tm = (size_t *)calloc((size_t)n_threads, sizeof(size_t));
    uint64_t *where = (uint64_t *)malloc(8 * (size_t)(n_reads ? n_reads : 1));
    uint16_t *who = (uint16_t *)malloc(2 * (size_t)(n_reads ? n_reads : 1));
#pragma omp parallel num_threads(n_threads)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        bwtintv_v mem = {0, 0, 0}, t0 = {0, 0, 0}, t1 = {0, 0, 0};
        bwtintv_v *tmpv[2] = {&t0, &t1};
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_reads; ++r) {
            const uint8_t *q = reads + read_off[r];
            int len = (int)(read_off[r + 1] - read_off[r]), x = 0;
            where[r] = tn[tid]; who[r] = (uint16_t)tid;
            uint32_t ns = 0;
            while (len >= min_seed_len && x < len) {
                if (q[x] < 4) {
                    x = bwt_smem1(bwt, len, q, x, 1, &mem, tmpv);
                    for (size_t i = 0; i < mem.n; ++i) {
                        bwtintv_t *p = &mem.a[i];
                        int qb = (int)(p->info >> 32), qe = (int)(uint32_t)p->info;
                        if (qe - qb < min_seed_len) continue;
                        uint64_t s = p->x[2], step = 1, count = s;
                        if (max_occ > 0) { step = s > (uint64_t)max_occ ? s / (uint64_t)max_occ : 1; count = (s + step - 1) / step; if (count > (uint64_t)max_occ) count = (uint64_t)max_occ; }
                        for (uint64_t t = 0; t < count; ++t) {
                            if (tn[tid] == tm[tid]) { tm[tid] = tm[tid] ? tm[tid] * 2 : 4096; tv[tid] = (ref_seed_rec_t *)realloc(tv[tid], tm[tid] * sizeof(ref_seed_rec_t)); }
                            ref_seed_rec_t *o = &tv[tid][tn[tid]++];
                            o->rbeg = bwt_sa(bwt, p->x[0] + t * step); o->qb = qb; o->qe = qe; o->score = t == 0 ? (uint32_t)s : 0u;
                            ++ns;
                        }
                    }
                } else ++x;
            }
            n_seeds[r] = ns;
        }
        free(mem.a); free(t0.a); free(t1.a);
    }
    uint64_t tot = 0;
    for (int64_t r = 0; r < n_reads; ++r) { seed_off[r] = tot; tot += n_seeds[r]; }
    int over = (int64_t)tot > cap;
    if (!over) {
#pragma omp parallel for num_threads(n_threads) schedule(static)
        for (int64_t r = 0; r < n_reads; ++r) {
            const ref_seed_rec_t *src = tv[who[r]] + where[r];
            for (uint32_t i = 0; i < n_seeds[r]; ++i) {
                uint64_t o = seed_off[r] + i;
                rbeg[o] = src[i].rbeg; qbeg_qend[2 * o] = src[i].qb; qbeg_qend[2 * o + 1] = src[i].qe; score[o] = src[i].score;
            }
        }
    }
    for (int t = 0; t < n_threads; ++t) free(tv[t]);
    free(tv); free(tn); free(tm); free(where); free(who);
    return over ? -(int64_t)tot : (int64_t)tot;
}

void ref_ksw_batch(int64_t n, const uint8_t *qseq, const uint32_t *qoff, const uint32_t *qlen,
                   const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen, const uint32_t *h0,
                   const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop,
                   int32_t *out6, int n_threads)
{
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 256)
    for (int64_t a = 0; a < n; ++a)
        ref_ksw_extend2((int)qlen[a], qseq + qoff[a], (int)tlen[a], tseq + toff[a], mat, o_del, e_del, o_ins, e_ins,
                        w, end_bonus, zdrop, (int)h0[a], out6 + 6 * a);
}

/* fused seed -> extend pass (same record layout as bwa_b200_read_result_t) computed with the
 * reference's bwt_smem1 / bwt_sa / ksw_extend2; job shapes from oracle/jobs_common.h */
typedef struct { int64_t seed_rbeg; int32_t seed_qbeg, seed_qend, n_seeds, h0; int32_t left[6], right[6]; } ref_read_result_t;

void ref_pipeline_batch(void *h, const uint8_t *fwd, int64_t l_pac, const uint8_t *reads, const uint64_t *read_off,
                        int64_t n_reads, int min_seed_len, int max_occ, const int8_t *mat,
                        int o_del, int e_del, int o_ins, int e_ins, int w, int end_bonus, int zdrop, int a,
                        ref_read_result_t *out, int n_threads)
{
    bwt_t *bwt = (bwt_t *)h;
    job_rules_t rules = {a, o_del, e_del, o_ins, e_ins, w};
#pragma omp parallel num_threads(n_threads)
    {
        bwtintv_v mem = {0, 0, 0}, t0 = {0, 0, 0}, t1 = {0, 0, 0};
        bwtintv_v *tmpv[2] = {&t0, &t1};
        uint64_t *rb = NULL; int32_t *qb = NULL, *qe = NULL; size_t sc = 0;
        uint8_t *qbuf = NULL, *tbuf = NULL; size_t qcap = 0, tcap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_reads; ++r) {
            const uint8_t *q = reads + read_off[r];
            int len = (int)(read_off[r + 1] - read_off[r]), x = 0;
            ref_read_result_t *o = &out[r];
            memset(o, 0, sizeof(*o));
            o->seed_qbeg = -1; o->seed_rbeg = -1; o->seed_qend = -1;
            o->left[4] = -1; o->right[4] = -1;
            size_t ns = 0;
            while (len >= min_seed_len && x < len) {
                if (q[x] < 4) {
                    x = bwt_smem1(bwt, len, q, x, 1, &mem, tmpv);
                    for (size_t i = 0; i < mem.n; ++i) {
                        bwtintv_t *p = &mem.a[i];
                        int beg = (int)(p->info >> 32), end = (int)(uint32_t)p->info;
                        if (end - beg < min_seed_len) continue;
                        int64_t step = p->x[2] > (uint64_t)max_occ ? p->x[2] / max_occ : 1, k, count;
                        for (k = count = 0; k < (int64_t)p->x[2] && count < max_occ; k += step, ++count) {
                            if (ns == sc) { sc = sc * 2 + 64; rb = (uint64_t *)realloc(rb, sc * 8); qb = (int32_t *)realloc(qb, sc * 4); qe = (int32_t *)realloc(qe, sc * 4); }
                            rb[ns] = bwt_sa(bwt, p->x[0] + k); qb[ns] = beg; qe[ns] = end; ++ns;
                        }
                    }
                } else ++x;
            }
            o->n_seeds = (int32_t)ns;
            int64_t best = jc_choose(rb, qb, qe, (int64_t)ns, l_pac);
            if (best < 0) continue;
            job_pair_t j;
            jc_shape(&rules, l_pac, len, (int64_t)rb[best], qb[best], qe[best], &j);
            o->seed_rbeg = j.rbeg; o->seed_qbeg = j.qbeg; o->seed_qend = j.qend; o->h0 = j.h0;
            if ((size_t)len + 8 > qcap) { qcap = (size_t)len + 8; qbuf = (uint8_t *)realloc(qbuf, qcap); }
            size_t tneed = (size_t)(j.lt > j.rt ? j.lt : j.rt) + 8;
            if (tneed > tcap) { tcap = tneed; tbuf = (uint8_t *)realloc(tbuf, tcap); }
            int32_t none[6] = {j.h0, 0, 0, 0, -1, 0};
            memcpy(o->left, none, sizeof(none)); memcpy(o->right, none, sizeof(none));
            if (j.lq > 0) {
                for (int i = 0; i < j.lq; ++i) qbuf[i] = q[j.qbeg - 1 - i];
                for (int i = 0; i < j.lt; ++i) tbuf[i] = jc_text(fwd, l_pac, j.lt_start - 1 - i);
                ref_ksw_extend2(j.lq, qbuf, j.lt, tbuf, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, j.h0, o->left);
            }
            if (j.rq > 0) {
                for (int i = 0; i < j.rt; ++i) tbuf[i] = jc_text(fwd, l_pac, j.rt_start + i);
                ref_ksw_extend2(j.rq, q + j.qend, j.rt, tbuf, mat, o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, j.h0, o->right);
            }
        }
        free(mem.a); free(t0.a); free(t1.a); free(rb); free(qb); free(qe); free(qbuf); free(tbuf);
    }
}

/* ---- CIGAR path: the reference's ksw_global2 (bwa_index/ksw.c:504) and bwa_gen_cigar2
 * (bwa_index/bwa.c:121).  The CIGAR is copied out (first `cap` operations) and the reference's buffer freed. */
int ref_ksw_global2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, const int8_t *mat,
                    int o_del, int e_del, int o_ins, int e_ins, int w, int *n_cigar, uint32_t *cigar_out, int cap)
{
    uint32_t *cigar = NULL;
    int sc = ksw_global2(qlen, query, tlen, target, 5, mat, o_del, e_del, o_ins, e_ins, w, n_cigar, &cigar);
    for (int i = 0; i < *n_cigar && i < cap; ++i) cigar_out[i] = cigar[i];
    free(cigar);
    return sc;
}

/* pac: 2-bit packed forward reference (4 bases per byte, bntseq.c _get_pac).  Returns n_cigar (0 when the reference
 * rejects the job and returns NULL). */
int ref_gen_cigar2(const int8_t *mat, int o_del, int e_del, int o_ins, int e_ins, int w_, int64_t l_pac, const uint8_t *pac,
                   int l_query, const uint8_t *query, int64_t rb, int64_t re, int *score, int *nm, uint32_t *cigar_out, int cap)
{
    int n_cigar = 0;
    uint8_t *q = (uint8_t *)malloc(l_query > 0 ? l_query : 1);
    memcpy(q, query, l_query > 0 ? l_query : 0);
    *score = 0; *nm = -1;
    uint32_t *cigar = bwa_gen_cigar2(mat, o_del, e_del, o_ins, e_ins, w_, l_pac, pac, l_query, q, rb, re, score, &n_cigar, nm);
    for (int i = 0; cigar && i < n_cigar && i < cap; ++i) cigar_out[i] = cigar[i];
    free(cigar); free(q);
    return n_cigar;
}
