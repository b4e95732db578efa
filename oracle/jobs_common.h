/*
 * oracle/jobs_common.h -- TEST INFRASTRUCTURE ONLY.
 * Extension-job construction for the seed -> extend benchmark pipeline, restated from the
 * reference's mem_chain2aln for a chain that holds one seed:
 *   cal_max_gap                    src/bwamem.c:996-1002  (bwa_index/bwamem.c:620-627)
 *   rmax[] window + strand clamp   src/bwamem.c:1180-1201 (bwa_index/bwamem.c:641-659)
 *   left / right job shapes        src/bwamem.c:1356-1424; both sides use h0 = seed_len * a as the
 *                                  fork does (src/bwamem.c:1360,1384)
 *   bridging seeds are dropped     bns_intv2rid < 0, src/bwamem.c:433-434
 * Chaining itself (mem_chain / mem_chain_flt) is outside the hot path; the
 * pipeline extends the longest located seed of each read (first one on ties).
 */
#ifndef JOBS_COMMON_H
#define JOBS_COMMON_H
#include <stdint.h>

typedef struct {
    int32_t a, o_del, e_del, o_ins, e_ins, w;
} job_rules_t;

typedef struct {                 /* one read's pair of jobs */
    int64_t  rbeg;               /* chosen seed; qbeg = -1 when the read has no usable seed */
    int32_t  qbeg, qend;
    int32_t  lq, lt, rq, rt;     /* query / target lengths of the left and right job       */
    int64_t  lt_start, rt_start; /* left target = T[lt_start-1 .. lt_start-lt] (reversed); right = T[rt_start .. +rt) */
    int32_t  h0;
} job_pair_t;

static inline int jc_max_gap(const job_rules_t *r, int qlen)
{
    int l_del = (int)((double)(qlen * r->a - r->o_del) / r->e_del + 1.);
    int l_ins = (int)((double)(qlen * r->a - r->o_ins) / r->e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < (r->w << 1) ? l : (r->w << 1);
}

/* base p of T = fwd + revcomp(fwd) */
static inline uint8_t jc_text(const uint8_t *fwd, int64_t l_pac, int64_t p)
{
    return p < l_pac ? fwd[p] : (uint8_t)(3 - fwd[2 * l_pac - 1 - p]);
}

/* longest seed that does not bridge the forward/reverse boundary, first on ties */
static inline int64_t jc_choose(const uint64_t *rbeg, const int32_t *qbeg, const int32_t *qend, int64_t n, int64_t l_pac)
{
    int64_t best = -1; int best_len = -1;
    for (int64_t i = 0; i < n; ++i) {
        int len = qend[i] - qbeg[i];
        if ((int64_t)rbeg[i] < l_pac && (int64_t)rbeg[i] + len > l_pac) continue;
        if (len > best_len) { best_len = len; best = i; }
    }
    return best;
}

static inline void jc_shape(const job_rules_t *r, int64_t l_pac, int read_len, int64_t rbeg, int qbeg, int qend, job_pair_t *j)
{
    int slen = qend - qbeg;
    int64_t rmax0 = rbeg - (qbeg + jc_max_gap(r, qbeg));
    int64_t rmax1 = rbeg + slen + ((read_len - qend) + jc_max_gap(r, read_len - qend));
    if (rmax0 < 0) rmax0 = 0;
    if (rmax1 > (l_pac << 1)) rmax1 = l_pac << 1;
    if (rmax0 < l_pac && l_pac < rmax1) { if (rbeg < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
    j->rbeg = rbeg; j->qbeg = qbeg; j->qend = qend; j->h0 = slen * r->a;
    j->lq = qbeg; j->lt = qbeg > 0 ? (int32_t)(rbeg - rmax0) : 0; j->lt_start = rbeg;
    j->rq = read_len - qend; j->rt = qend < read_len ? (int32_t)(rmax1 - (rbeg + slen)) : 0; j->rt_start = rbeg + slen;
}
#endif
