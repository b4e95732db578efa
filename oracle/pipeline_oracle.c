/*
 * oracle/pipeline_oracle.c -- TEST INFRASTRUCTURE ONLY.
 * CPU statement of the benchmark pipeline  seed (pass-1 SMEMs + sampled SA rows) ->
 * pick the longest seed -> left/right ksw_extend2, built from fmd_oracle.c, jobs_common.h and
 * ksw_oracle.c.  Output layout = bwa_b200_read_result_t (include/bwamem_b200.h).
 */
#include <stdlib.h>
#include <string.h>
#include "fmd_oracle.h"
#include "ksw_oracle.h"
#include "jobs_common.h"
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int64_t seed_rbeg; int32_t seed_qbeg, seed_qend; int32_t n_seeds; int32_t h0;
    ksw_ext_result_t left, right;
} read_result_t;

void pipeline_oracle(const fmd_index_t *idx, const uint8_t *fwd, int64_t l_pac,
                     const uint8_t *reads, const uint64_t *read_off, int64_t n_reads,
                     int min_seed_len, int max_occ, const ksw_params_t *kp, int a,
                     read_result_t *out, int n_threads, fmd_counters_t *fc, ksw_counters_t *kc)
{
    if (n_threads < 1) n_threads = 1;
    job_rules_t rules = {a, kp->o_del, kp->e_del, kp->o_ins, kp->e_ins, kp->w};
    uint64_t fe = 0, fb = 0, fl = 0, fn = 0, fs = 0, fef = 0, fbf = 0, kcells = 0, krows = 0, krect = 0;
#pragma omp parallel num_threads(n_threads) reduction(+:fe,fb,fl,fn,fs,fef,fbf,kcells,krows,krect)
    {
        fmd_intv_t *mem = NULL; size_t mem_cap = 0;
        uint64_t *rb = NULL; int32_t *qb = NULL, *qe = NULL; size_t sc = 0;
        uint8_t *qbuf = NULL, *tbuf = NULL; size_t qcap = 0, tcap = 0;
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_reads; ++r) {
            fmd_counters_t c; memset(&c, 0, sizeof(c));
            ksw_counters_t k = {0, 0, 0};
            const uint8_t *q = reads + read_off[r];
            int len = (int)(read_off[r + 1] - read_off[r]);
            read_result_t *o = &out[r];
            memset(o, 0, sizeof(*o));
            o->seed_qbeg = -1; o->seed_rbeg = -1; o->seed_qend = -1;
            o->left.gscore = -1; o->right.gscore = -1;      /* no extension ran: score = h0 = 0, gscore = -1 */
            if ((size_t)len + 2 > mem_cap) { mem_cap = (size_t)len + 2; mem = (fmd_intv_t *)realloc(mem, mem_cap * sizeof(*mem)); }
            int n = len >= min_seed_len ? fmd_collect_pass1(idx, len, q, min_seed_len, mem, &c) : 0;
            size_t ns = 0;
            for (int i = 0; i < n; ++i) {
                uint64_t s = mem[i].s, step = s > (uint64_t)max_occ ? s / (uint64_t)max_occ : 1, cnt = (s + step - 1) / step;
                if (cnt > (uint64_t)max_occ) cnt = (uint64_t)max_occ;
                if (ns + cnt > sc) { sc = (ns + cnt) * 2 + 16; rb = (uint64_t *)realloc(rb, sc * 8); qb = (int32_t *)realloc(qb, sc * 4); qe = (int32_t *)realloc(qe, sc * 4); }
                for (uint64_t t = 0; t < cnt; ++t, ++ns) { rb[ns] = fmd_sa(idx, mem[i].k + t * step, &c); qb[ns] = mem[i].beg; qe[ns] = mem[i].end; }
            }
            o->n_seeds = (int32_t)ns;
            int64_t best = jc_choose(rb, qb, qe, (int64_t)ns, l_pac);
            if (best >= 0) {
                job_pair_t j;
                jc_shape(&rules, l_pac, len, (int64_t)rb[best], qb[best], qe[best], &j);
                o->seed_rbeg = j.rbeg; o->seed_qbeg = j.qbeg; o->seed_qend = j.qend; o->h0 = j.h0;
                if ((size_t)len + 8 > qcap) { qcap = (size_t)len + 8; qbuf = (uint8_t *)realloc(qbuf, qcap); }
                size_t tneed = (size_t)(j.lt > j.rt ? j.lt : j.rt) + 8;
                if (tneed > tcap) { tcap = tneed; tbuf = (uint8_t *)realloc(tbuf, tcap); }
                ksw_ext_result_t none = {j.h0, 0, 0, 0, -1, 0};
                o->left = none; o->right = none;
                if (j.lq > 0) {
                    for (int i = 0; i < j.lq; ++i) qbuf[i] = q[j.qbeg - 1 - i];
                    for (int i = 0; i < j.lt; ++i) tbuf[i] = jc_text(fwd, l_pac, j.lt_start - 1 - i);
                    ksw_extend2_oracle(j.lq, qbuf, j.lt, tbuf, kp, j.h0, &o->left, &k);
                }
                if (j.rq > 0) {
                    for (int i = 0; i < j.rt; ++i) tbuf[i] = jc_text(fwd, l_pac, j.rt_start + i);
                    ksw_extend2_oracle(j.rq, q + j.qend, j.rt, tbuf, kp, j.h0, &o->right, &k);
                }
            }
            fe += c.n_extend; fb += c.n_bucket; fl += c.n_lf; fn += c.n_located; fs += c.n_smem; fef += c.n_extend_fwd; fbf += c.n_bucket_fwd;
            kcells += k.cells; krows += k.rows; krect += k.rect;
        }
        free(mem); free(rb); free(qb); free(qe); free(qbuf); free(tbuf);
    }
    if (fc) { fc->n_extend += fe; fc->n_bucket += fb; fc->n_lf += fl; fc->n_located += fn; fc->n_smem += fs; fc->n_extend_fwd += fef; fc->n_bucket_fwd += fbf; }
    if (kc) { kc->cells += kcells; kc->rows += krows; kc->rect += krect; }
}
