/*
 * oracle/fork_mem_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 * Entry point around the UNMODIFIED host code of the reference fork (src/bwamem.c compiled as
 * C++ exactly like the reference Makefile:14 does), built by oracle/build_ref.sh from the sources
 * where they lie into oracle/_ref/libforkmem.so.  It calls, per read and in the order of the
 * reference worker (src/bwamem.c:2055-2093):
 *     mem_chain -> mem_chain_flt -> mem_flt_chained_seeds -> mem_chain2aln for every chain
 * and returns the chains, the alignment-region records and the extension jobs that
 * fill_extension handed to the SHORT / LONG batches.  This file contains no reference code.
 *
 * What it has to provide because the fork's host code cannot run without a GPU:
 *   - gasal_host_batch_fill / gasal_host_alns_resize / Parameters: GASAL2's versions sit on
 *     cudaHostAlloc'ed pages (GASAL2/src/host_batch.cpp:79-153, interfaces.cpp:26-78); here the
 *     same contract (append `size` bytes at `idx`, pad to a multiple of 8 with N_CODE = 4, return
 *     the new running offset) over one malloc'ed buffer per side.
 *   - the timing / statistics globals that src/bwamem.c declares extern (src/fastmap.c:138-160).
 * mem_chain_t and mem_chain_v are private to src/bwamem.c (:318-329); their layout is declared
 * again here so the chains can be read back.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include "bwamem.h"
#include "bntseq.h"
#include "ksw.h"
#include "utils.h"
#include "GPUSeed/seed_gen.h"

/* ---- globals the fork's bwamem.c / kthread.c expect from fastmap.c ---- */
time_struct *extension_time = NULL;
uint64_t *no_of_extensions = NULL;
double *load_balance_waste_time = NULL;
double total_load_balance_waste_time = 0;
gasal_gpu_storage_v *gpu_storage_vec_arr = NULL;

/* ---- private types of src/bwamem.c:318-329 ---- */
typedef struct {
    int n, m, first, rid;
    uint32_t w : 29, kept : 2, is_alt : 1;
    float frac_rep;
    int64_t pos;
    mem_seed_t *seeds;
} mem_chain_t;
typedef struct { size_t n, m; mem_chain_t *a; } mem_chain_v;

mem_chain_v mem_chain(const mem_opt_t *opt, const bwt_t *bwt, const bntseq_t *bns, int len, const uint8_t *seq, mem_seed_v_gpu *gpu_results, int j);
int mem_chain_flt(const mem_opt_t *opt, int n_chn, mem_chain_t *a);
int mem_sort_dedup_patch(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, uint8_t *query, int n, mem_alnreg_t *a);
int mem_mark_primary_se(const mem_opt_t *opt, int n, mem_alnreg_t *a, int64_t id);
int mem_approx_mapq_se(const mem_opt_t *opt, const mem_alnreg_t *a);
void ks_combsort_mem_ars2(size_t n, mem_alnreg_t a[]);      /* KSORT_INIT instances of src/bwamem.c:565-575 */
void ks_combsort_mem_ars(size_t n, mem_alnreg_t a[]);
void ks_combsort_mem_ars_hash(size_t n, mem_alnreg_t a[]);
void ks_combsort_mem_ars_hash2(size_t n, mem_alnreg_t a[]);
void ks_introsort_mem_ars2(size_t n, mem_alnreg_t a[]);
void ks_introsort_mem_ars(size_t n, mem_alnreg_t a[]);
void ks_introsort_mem_ars_hash(size_t n, mem_alnreg_t a[]);
void ks_introsort_mem_ars_hash2(size_t n, mem_alnreg_t a[]);
void mem_flt_chained_seeds(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query, const uint8_t *query, int n_chn, mem_chain_t *a);
void mem_chain2aln(const mem_opt_t *opt, const bntseq_t *bns, const uint8_t *pac, int l_query, const uint8_t *query, const mem_chain_t *c,
                   mem_alnreg_v *regs, int *curr_read_offset, int *curr_ref_offset, gpu_batch *curr_gpu_batch_short, gpu_batch *curr_gpu_batch_long);

/* ---- GASAL2 host-side contract over plain memory ---- */
struct SideBuf { std::vector<uint8_t> q, t; };
static thread_local SideBuf g_buf[2];                       /* per thread: fork_align_batch runs reads on several threads */
static thread_local gasal_gpu_storage_t *g_sto[2] = {NULL, NULL};

Parameters::Parameters(int argc_, char **argv_) : sa(1), sb(4), gapo(6), gape(1), print_out(0), n_threads(1), k_band(0), isPacked(false), isReverseComplement(false), argc(argc_), argv(argv_) {}
Parameters::~Parameters() {}

uint32_t gasal_host_batch_fill(gasal_gpu_storage_t *sto, uint32_t idx, const char *data, uint32_t size, data_source SRC)
{
    int side = sto == g_sto[0] ? 0 : 1;
    std::vector<uint8_t> &v = SRC == QUERY ? g_buf[side].q : g_buf[side].t;
    uint32_t pad = (8 - size % 8) % 8;
    if (v.size() < (size_t)idx + size + pad) v.resize((size_t)idx + size + pad);
    memcpy(v.data() + idx, data, size);
    memset(v.data() + idx + size, 4 /* N_CODE */, pad);
    return idx + size + pad;
}
uint32_t gasal_host_batch_addbase(gasal_gpu_storage_t *sto, uint32_t idx, const char base, data_source SRC)
{ /* GASAL2/src/host_batch.cpp:156-159: one byte, no padding (only bns_get_seq_gpu uses it; not on this path) */
    int side = sto == g_sto[0] ? 0 : 1;
    std::vector<uint8_t> &v = SRC == QUERY ? g_buf[side].q : g_buf[side].t;
    if (v.size() < (size_t)idx + 1) v.resize((size_t)idx + 1);
    v[idx] = (uint8_t)base;
    return idx + 1;
}
void gasal_host_alns_resize(gasal_gpu_storage_t *sto, int new_max, Parameters *)
{
    sto->host_query_batch_offsets = (uint32_t *)realloc(sto->host_query_batch_offsets, (size_t)new_max * 4);
    sto->host_target_batch_offsets = (uint32_t *)realloc(sto->host_target_batch_offsets, (size_t)new_max * 4);
    sto->host_query_batch_lens = (uint32_t *)realloc(sto->host_query_batch_lens, (size_t)new_max * 4);
    sto->host_target_batch_lens = (uint32_t *)realloc(sto->host_target_batch_lens, (size_t)new_max * 4);
    sto->host_seed_scores = (uint32_t *)realloc(sto->host_seed_scores, (size_t)new_max * 4);
    sto->host_max_n_alns = (uint32_t)new_max;
}
void gasal_aln_async(gasal_gpu_storage_t *, uint32_t, uint32_t, uint32_t, Parameters *) { fprintf(stderr, "fork_mem_shim: gasal_aln_async is not available\n"); abort(); }
int gasal_is_aln_async_done(gasal_gpu_storage_t *) { return -2; }
extern "C" int bit_vec_filter_sse1(char *, char *, int, int) { fprintf(stderr, "fork_mem_shim: the SHD filter is not built\n"); abort(); return 0; }

static gasal_gpu_storage_t *new_storage(void)
{
    gasal_gpu_storage_t *s = (gasal_gpu_storage_t *)calloc(1, sizeof(*s));
    gasal_host_alns_resize(s, 1024, NULL);
    s->host_max_query_batch_bytes = s->host_max_target_batch_bytes = 0x7fffffff;
    return s;
}

/* ---- flat records handed back to Python ---- */
extern "C" {

typedef struct {
    int32_t a, b, o_del, e_del, o_ins, e_ins, w, min_seed_len, max_occ, max_chain_gap, min_chain_weight, max_chain_extend;
    float mask_level, drop_ratio;
} fork_opt_t;

typedef struct { int64_t pos; int32_t rid, n, w, kept, first, is_alt; float frac_rep; int32_t seed_off; } fork_chain_t;
typedef struct { int64_t rbeg; int32_t qbeg, len, score, pad; } fork_seed_t;
typedef struct {
    int64_t rb_est, re_est, target_seed_begin;
    int32_t qb_est, qe_est, rid, score, truesc, align_sides, where_is_long, query_seed_begin, seedlen0, seedcov, w;
    float frac_rep;
} fork_reg_t;
typedef struct { uint32_t qoff, qlen, toff, tlen, h0; } fork_job_t;

/* seeds in the reference layout of mem_seed_v_gpu (seed_gen.h:68-75): every SMEM group holds all `score` rows.
 * Returns 0, or -1 when an output capacity is too small.  Job sequences of side s (0 SHORT, 1 LONG) can be read
 * with fork_mem_seq() until the next call. */
int fork_mem_read(const fork_opt_t *fo, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                  const uint8_t *pac, int l_query, const uint8_t *query,
                  uint32_t n_seeds, const uint64_t *rbeg, const int32_t *qbeg_qend, const uint32_t *score,
                  int32_t *n_chains, fork_chain_t *chains, int cap_chains, fork_seed_t *cseeds, int cap_cseeds,
                  int32_t *n_regs, fork_reg_t *regs_out, int cap_regs,
                  int32_t n_jobs[2], fork_job_t *jobs_short, fork_job_t *jobs_long, int cap_jobs)
{
    mem_opt_t *opt = mem_opt_init();
    opt->a = fo->a; opt->b = fo->b; opt->o_del = fo->o_del; opt->e_del = fo->e_del; opt->o_ins = fo->o_ins; opt->e_ins = fo->e_ins;
    opt->w = fo->w; opt->min_seed_len = fo->min_seed_len; opt->max_occ = fo->max_occ; opt->max_chain_gap = fo->max_chain_gap;
    opt->min_chain_weight = fo->min_chain_weight; opt->max_chain_extend = fo->max_chain_extend;
    opt->mask_level = fo->mask_level; opt->drop_ratio = fo->drop_ratio;
    bwa_fill_scmat(opt->a, opt->b, opt->mat);

    bntseq_t bns;
    memset(&bns, 0, sizeof(bns));
    bns.l_pac = l_pac; bns.n_seqs = n_ctg;
    std::vector<bntann1_t> anns(n_ctg);
    char nm[] = "ctg";
    for (int i = 0; i < n_ctg; ++i) { memset(&anns[i], 0, sizeof(bntann1_t)); anns[i].offset = ctg_off[i]; anns[i].len = ctg_len[i]; anns[i].is_alt = ctg_alt ? ctg_alt[i] : 0; anns[i].name = nm; anns[i].anno = nm; }
    bns.anns = anns.data();

    mem_seed_v_gpu gr;
    memset(&gr, 0, sizeof(gr));
    uint32_t zero = 0;
    gr.rbeg = (bwtint_t_gpu *)rbeg; gr.qbeg = (int2 *)qbeg_qend; gr.score = (uint32_t *)score;
    gr.n_ref_pos_fow_rev_results = &n_seeds; gr.n_ref_pos_fow_rev_prefix_sums = &zero;

    if (!g_sto[0]) { g_sto[0] = new_storage(); g_sto[1] = new_storage(); }
    gpu_batch gb[2];
    for (int s = 0; s < 2; ++s) {
        memset(&gb[s], 0, sizeof(gpu_batch));
        gb[s].gpu_storage = g_sto[s];
        g_sto[s]->current_n_alns = 0;
        g_buf[s].q.clear(); g_buf[s].t.clear();
    }
    int cur_read_off[2] = {0, 0}, cur_ref_off[2] = {0, 0};

    std::vector<uint8_t> q(query, query + l_query);
    mem_chain_v chn = mem_chain(opt, NULL, &bns, l_query, q.data(), &gr, 0);
    chn.n = mem_chain_flt(opt, (int)chn.n, chn.a);
    mem_flt_chained_seeds(opt, &bns, pac, l_query, q.data(), (int)chn.n, chn.a);

    int rc = 0, so = 0;
    *n_chains = (int32_t)chn.n;
    for (size_t i = 0; i < chn.n; ++i) {
        const mem_chain_t *c = &chn.a[i];
        if ((int)i < cap_chains) {
            fork_chain_t *o = &chains[i];
            o->pos = c->pos; o->rid = c->rid; o->n = c->n; o->w = (int32_t)c->w; o->kept = (int32_t)c->kept; o->first = c->first;
            o->is_alt = (int32_t)c->is_alt; o->frac_rep = c->frac_rep; o->seed_off = so;
        } else rc = -1;
        for (int k = 0; k < c->n; ++k, ++so) {
            if (so < cap_cseeds) { cseeds[so].rbeg = c->seeds[k].rbeg; cseeds[so].qbeg = c->seeds[k].qbeg; cseeds[so].len = c->seeds[k].len; cseeds[so].score = c->seeds[k].score; cseeds[so].pad = 0; }
            else rc = -1;
        }
    }

    mem_alnreg_v regs;
    regs.n = regs.m = 0; regs.a = NULL;
    for (size_t i = 0; i < chn.n; ++i) {
        mem_chain2aln(opt, &bns, pac, l_query, q.data(), &chn.a[i], &regs, cur_read_off, cur_ref_off, &gb[SHORT], &gb[LONG]);
        free(chn.a[i].seeds);
    }
    free(chn.a);
    *n_regs = (int32_t)regs.n;
    for (size_t i = 0; i < regs.n; ++i) {
        if ((int)i >= cap_regs) { rc = -1; break; }
        const mem_alnreg_t *a = &regs.a[i];
        fork_reg_t *o = &regs_out[i];
        o->rb_est = a->rb_est; o->re_est = a->re_est; o->target_seed_begin = a->target_seed_begin;
        o->qb_est = a->qb_est; o->qe_est = a->qe_est; o->rid = a->rid; o->score = a->score; o->truesc = a->truesc;
        o->align_sides = a->align_sides; o->where_is_long = a->where_is_long; o->query_seed_begin = a->query_seed_begin;
        o->seedlen0 = a->seedlen0; o->seedcov = a->seedcov; o->w = a->w; o->frac_rep = a->frac_rep;
    }
    free(regs.a);
    for (int s = 0; s < 2; ++s) {
        n_jobs[s] = gb[s].n_seqs;
        fork_job_t *jo = s == 0 ? jobs_short : jobs_long;
        for (int k = 0; k < gb[s].n_seqs; ++k) {
            if (k >= cap_jobs) { rc = -1; break; }
            jo[k].qoff = g_sto[s]->host_query_batch_offsets[k]; jo[k].qlen = g_sto[s]->host_query_batch_lens[k];
            jo[k].toff = g_sto[s]->host_target_batch_offsets[k]; jo[k].tlen = g_sto[s]->host_target_batch_lens[k];
            jo[k].h0 = g_sto[s]->host_seed_scores[k];
        }
    }
    free(opt);
    return rc;
}

/* byte buffers of the last call: side 0 SHORT / 1 LONG, which 0 query / 1 target */
const uint8_t *fork_mem_seq(int side, int which, uint64_t *n_bytes)
{
    std::vector<uint8_t> &v = which == 0 ? g_buf[side].q : g_buf[side].t;
    *n_bytes = v.size();
    return v.data();
}

/* the fork's own mem_reg2aln (src/bwamem.c:2344-2438) on one alignment region.  pac: 2-bit packed forward reference; query: codes
 * 0..4 of the whole read.  out8 = pos, rid, is_rev, score (the mem_aln_t one = ar->score), NM, n_cigar, flag, mapq; the CIGAR
 * (soft clips included) is copied to cigar_out.  Returns n_cigar. */
int fork_reg2aln(const fork_opt_t *fo, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const uint8_t *pac,
                 int l_query, const uint8_t *query, int qb, int qe, int64_t rb, int64_t re, int truesc, int ar_w, int ar_score,
                 int64_t out8[8], uint32_t *cigar_out, int cap)
{
    mem_opt_t *opt = mem_opt_init();
    opt->a = fo->a; opt->b = fo->b; opt->o_del = fo->o_del; opt->e_del = fo->e_del; opt->o_ins = fo->o_ins; opt->e_ins = fo->e_ins; opt->w = fo->w;
    bwa_fill_scmat(opt->a, opt->b, opt->mat);
    bntseq_t bns;
    memset(&bns, 0, sizeof(bns));
    bns.l_pac = l_pac; bns.n_seqs = n_ctg;
    std::vector<bntann1_t> anns(n_ctg);
    char nm[] = "ctg";
    for (int i = 0; i < n_ctg; ++i) { memset(&anns[i], 0, sizeof(bntann1_t)); anns[i].offset = ctg_off[i]; anns[i].len = ctg_len[i]; anns[i].name = nm; anns[i].anno = nm; }
    bns.anns = anns.data();
    mem_alnreg_t ar;
    memset(&ar, 0, sizeof(ar));
    ar.rb = rb; ar.re = re; ar.qb = qb; ar.qe = qe; ar.truesc = truesc; ar.w = ar_w; ar.score = ar_score; ar.secondary = -1;
    if (rb >= 0 && re >= 0) {
        int is_rev;
        ar.rid = bns_pos2rid(&bns, bns_depos(&bns, rb < l_pac ? rb : re - 1, &is_rev));     /* mem_reg2aln asserts a.rid == ar->rid */
    }
    mem_aln_t a = mem_reg2aln(opt, &bns, pac, l_query, (const char *)query, &ar);
    out8[0] = a.pos; out8[1] = a.rid; out8[2] = a.is_rev; out8[3] = a.score; out8[4] = a.NM; out8[5] = a.n_cigar; out8[6] = a.flag; out8[7] = a.mapq;
    for (int i = 0; i < a.n_cigar && i < cap; ++i) cigar_out[i] = a.cigar[i];
    int n = a.n_cigar;
    free(a.cigar);
    free(opt);
    return n;
}

/* the fork's own mem_sort_dedup_patch (src/bwamem.c:620-681), the is_alt marking of its caller (:2321-2325), mem_mark_primary_se
 * (:715-760) and mem_approx_mapq_se as mem_reg2aln applies it (:1690-1716, :2363) on one read's alignment regions, in place.
 * The record is oracle/region_oracle.h's region_t.  opt13 = a b o_del e_del o_ins e_ins w min_seed_len max_chain_gap mapQ_coef_fac
 * then mask_level, mask_level_redun, mapQ_coef_len as floats.  Returns the new count. */
typedef struct {
    int64_t rb, re;
    uint64_t hash;
    int32_t qb, qe, rid, score, truesc, sub, alt_sc, csub, sub_n, w, seedcov, secondary, secondary_all, seedlen0, n_comp, is_alt;
    float frac_rep;
    int32_t mapq;
} fork_region_t;
typedef struct {
    int32_t a, b, o_del, e_del, o_ins, e_ins, w, min_seed_len, max_chain_gap, mapQ_coef_fac;
    float mask_level, mask_level_redun, mapQ_coef_len;
} fork_region_opt_t;

int fork_finish_regs(const fork_region_opt_t *fo, int64_t l_pac, int n_ctg, const int32_t *ctg_alt, const uint8_t *pac,
                     int l_query, const uint8_t *query_in, int n, fork_region_t *r, int64_t id, int *n_pri)
{
    mem_opt_t *opt = mem_opt_init();
    opt->a = fo->a; opt->b = fo->b; opt->o_del = fo->o_del; opt->e_del = fo->e_del; opt->o_ins = fo->o_ins; opt->e_ins = fo->e_ins; opt->w = fo->w;
    opt->min_seed_len = fo->min_seed_len; opt->max_chain_gap = fo->max_chain_gap; opt->mapQ_coef_fac = fo->mapQ_coef_fac;
    opt->mask_level = fo->mask_level; opt->mask_level_redun = fo->mask_level_redun; opt->mapQ_coef_len = fo->mapQ_coef_len;
    bwa_fill_scmat(opt->a, opt->b, opt->mat);
    bntseq_t bns;
    memset(&bns, 0, sizeof(bns));
    bns.l_pac = l_pac; bns.n_seqs = n_ctg;
    std::vector<bntann1_t> anns(n_ctg > 0 ? n_ctg : 1);
    char nm[] = "ctg";
    for (int i = 0; i < n_ctg; ++i) { memset(&anns[i], 0, sizeof(bntann1_t)); anns[i].name = nm; anns[i].anno = nm; anns[i].is_alt = ctg_alt ? ctg_alt[i] : 0; }
    bns.anns = anns.data();
    std::vector<uint8_t> query(query_in, query_in + l_query);
    std::vector<mem_alnreg_t> a(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) {
        mem_alnreg_t &x = a[i];
        memset(&x, 0, sizeof(x));
        x.rb = r[i].rb; x.re = r[i].re; x.qb = r[i].qb; x.qe = r[i].qe; x.rid = r[i].rid; x.score = r[i].score; x.truesc = r[i].truesc;
        x.sub = r[i].sub; x.alt_sc = r[i].alt_sc; x.csub = r[i].csub; x.sub_n = r[i].sub_n; x.w = r[i].w; x.seedcov = r[i].seedcov;
        x.secondary = r[i].secondary; x.secondary_all = r[i].secondary_all; x.seedlen0 = r[i].seedlen0; x.n_comp = r[i].n_comp;
        x.is_alt = r[i].is_alt; x.frac_rep = r[i].frac_rep; x.hash = r[i].hash;
    }
    n = mem_sort_dedup_patch(opt, &bns, pac, query.data(), n, a.data());
    for (int i = 0; i < n; ++i)
        if (a[i].rid >= 0 && bns.anns[a[i].rid].is_alt) a[i].is_alt = 1;
    *n_pri = mem_mark_primary_se(opt, n, a.data(), id);
    for (int i = 0; i < n; ++i) {
        const mem_alnreg_t &x = a[i];
        r[i].rb = x.rb; r[i].re = x.re; r[i].qb = x.qb; r[i].qe = x.qe; r[i].rid = x.rid; r[i].score = x.score; r[i].truesc = x.truesc;
        r[i].sub = x.sub; r[i].alt_sc = x.alt_sc; r[i].csub = x.csub; r[i].sub_n = x.sub_n; r[i].w = x.w; r[i].seedcov = x.seedcov;
        r[i].secondary = x.secondary; r[i].secondary_all = x.secondary_all; r[i].seedlen0 = x.seedlen0; r[i].n_comp = x.n_comp;
        r[i].is_alt = x.is_alt; r[i].frac_rep = x.frac_rep; r[i].hash = x.hash;
        r[i].mapq = x.secondary < 0 ? mem_approx_mapq_se(opt, &x) : 0;
    }
    free(opt);
    return n;
}

/* the fork's own sort instances on the same records: comb != 0 -> ks_combsort, else ks_introsort; which = 0 mem_ars2 (end), 1 mem_ars
 * (score), 2 mem_ars_hash, 3 mem_ars_hash2.  Only the fields the comparators read and an identity (seedlen0) travel. */
void fork_sort_regs(int comb, int which, int n, fork_region_t *r)
{
    std::vector<mem_alnreg_t> a(n > 0 ? n : 1);
    for (int i = 0; i < n; ++i) {
        memset(&a[i], 0, sizeof(mem_alnreg_t));
        a[i].rb = r[i].rb; a[i].re = r[i].re; a[i].qb = r[i].qb; a[i].qe = r[i].qe; a[i].score = r[i].score; a[i].is_alt = r[i].is_alt;
        a[i].hash = r[i].hash; a[i].seedlen0 = r[i].seedlen0;
    }
    if (n > 0) {
        if (comb) {
            if (which == 0) ks_combsort_mem_ars2(n, a.data()); else if (which == 1) ks_combsort_mem_ars(n, a.data());
            else if (which == 2) ks_combsort_mem_ars_hash(n, a.data()); else ks_combsort_mem_ars_hash2(n, a.data());
        } else {
            if (which == 0) ks_introsort_mem_ars2(n, a.data()); else if (which == 1) ks_introsort_mem_ars(n, a.data());
            else if (which == 2) ks_introsort_mem_ars_hash(n, a.data()); else ks_introsort_mem_ars_hash2(n, a.data());
        }
    }
    for (int i = 0; i < n; ++i) {
        r[i].rb = a[i].rb; r[i].re = a[i].re; r[i].qb = a[i].qb; r[i].qe = a[i].qe; r[i].score = a[i].score; r[i].is_alt = a[i].is_alt;
        r[i].hash = a[i].hash; r[i].seedlen0 = a[i].seedlen0;
    }
}

/* ---- the CPU arm of bench.py's chained step: a whole read batch through the fork's own host code, on n_threads threads ----
 * Per read, in the order of the reference worker (src/bwamem.c:2055-2093, :2286-2306):
 *   mem_chain -> mem_chain_flt -> mem_flt_chained_seeds -> mem_chain2aln per kept chain   (the fork's functions, unmodified)
 *   every extension job of the read's SHORT then LONG batch through the fork's ksw_extend2 (src/ksw.c:864) with the given band /
 *   z-drop / end bonus, the local-vs-to-end rule of decoy_cpu_align (src/bwamem.c:1892-1901),
 *   the result gathering of the worker (src/bwamem.c:2286-2306; restated here, it is inline code of the worker).
 * Seeds arrive in the reference layout of mem_seed_v_gpu for the whole batch (seed_off = exclusive prefix sums, every SMEM group
 * holding all its rows).  Output: n_regs[r] and, at reg_off[r] (exclusive prefix sums, computed here), one fork_aln_t per region in
 * the order mem_chain2aln created them.  Returns the total number of regions, or -1 when cap_regs is too small (n_regs is still
 * filled).  cells_out (may be NULL) receives nothing: the fork's ksw_extend2 has no cell counter. */
typedef struct { int64_t rb, re; int32_t qb, qe, score, truesc, rid, seedcov, seedlen0, w; float frac_rep; int32_t pad; } fork_aln_t;

int64_t fork_align_batch(const fork_opt_t *fo, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                         const uint8_t *pac, int64_t n_reads, const uint8_t *reads, const uint64_t *read_off,
                         const uint64_t *rbeg, const int32_t *qbeg_qend, const uint32_t *score, const uint32_t *n_seeds, const uint64_t *seed_off,
                         int ext_w, int ext_zdrop, int ext_end_bonus, int ext_use_band, int pen_clip,
                         uint32_t *n_regs, uint64_t *reg_off, fork_aln_t *regs_out, int64_t cap_regs, int n_threads, uint64_t *n_jobs_out)
{
    if (n_threads < 1) n_threads = 1;
    std::vector<std::vector<fork_aln_t> > per_read((size_t)n_reads);
    uint64_t n_jobs_total = 0;
#pragma omp parallel num_threads(n_threads) reduction(+:n_jobs_total)
    {
        mem_opt_t *opt = mem_opt_init();
        opt->a = fo->a; opt->b = fo->b; opt->o_del = fo->o_del; opt->e_del = fo->e_del; opt->o_ins = fo->o_ins; opt->e_ins = fo->e_ins;
        opt->w = fo->w; opt->min_seed_len = fo->min_seed_len; opt->max_occ = fo->max_occ; opt->max_chain_gap = fo->max_chain_gap;
        opt->min_chain_weight = fo->min_chain_weight; opt->max_chain_extend = fo->max_chain_extend;
        opt->mask_level = fo->mask_level; opt->drop_ratio = fo->drop_ratio;
        bwa_fill_scmat(opt->a, opt->b, opt->mat);
        bntseq_t bns;
        memset(&bns, 0, sizeof(bns));
        bns.l_pac = l_pac; bns.n_seqs = n_ctg;
        std::vector<bntann1_t> anns(n_ctg);
        char nm[] = "ctg";
        for (int i = 0; i < n_ctg; ++i) { memset(&anns[i], 0, sizeof(bntann1_t)); anns[i].offset = ctg_off[i]; anns[i].len = ctg_len[i]; anns[i].is_alt = ctg_alt ? ctg_alt[i] : 0; anns[i].name = nm; anns[i].anno = nm; }
        bns.anns = anns.data();
        if (!g_sto[0]) { g_sto[0] = new_storage(); g_sto[1] = new_storage(); }
        std::vector<uint8_t> q;
        std::vector<int32_t> tri[2];
#pragma omp for schedule(dynamic, 64)
        for (int64_t r = 0; r < n_reads; ++r) {
            const int l_query = (int)(read_off[r + 1] - read_off[r]);
            q.assign(reads + read_off[r], reads + read_off[r + 1]);
            mem_seed_v_gpu gr;
            memset(&gr, 0, sizeof(gr));
            uint32_t ns = n_seeds[r], zero = 0;
            gr.rbeg = (bwtint_t_gpu *)(rbeg + seed_off[r]); gr.qbeg = (int2 *)(qbeg_qend + 2 * seed_off[r]); gr.score = (uint32_t *)(score + seed_off[r]);
            gr.n_ref_pos_fow_rev_results = &ns; gr.n_ref_pos_fow_rev_prefix_sums = &zero;
            gpu_batch gb[2];
            for (int s = 0; s < 2; ++s) {
                memset(&gb[s], 0, sizeof(gpu_batch));
                gb[s].gpu_storage = g_sto[s];
                g_sto[s]->current_n_alns = 0;
                g_buf[s].q.clear(); g_buf[s].t.clear();
            }
            int cur_read_off[2] = {0, 0}, cur_ref_off[2] = {0, 0};
            mem_chain_v chn = mem_chain(opt, NULL, &bns, l_query, q.data(), &gr, 0);
            chn.n = mem_chain_flt(opt, (int)chn.n, chn.a);
            mem_flt_chained_seeds(opt, &bns, pac, l_query, q.data(), (int)chn.n, chn.a);
            mem_alnreg_v regs;
            regs.n = regs.m = 0; regs.a = NULL;
            for (size_t i = 0; i < chn.n; ++i) {
                mem_chain2aln(opt, &bns, pac, l_query, q.data(), &chn.a[i], &regs, cur_read_off, cur_ref_off, &gb[SHORT], &gb[LONG]);
                free(chn.a[i].seeds);
            }
            free(chn.a);
            for (int s = 0; s < 2; ++s) {                      /* the extension jobs of this read */
                tri[s].resize((size_t)gb[s].n_seqs * 3 + 3);
                n_jobs_total += (uint64_t)gb[s].n_seqs;
                for (int k = 0; k < gb[s].n_seqs; ++k) {
                    const uint32_t ql = g_sto[s]->host_query_batch_lens[k], tl = g_sto[s]->host_target_batch_lens[k];
                    const uint8_t *qs = g_buf[s].q.data() + g_sto[s]->host_query_batch_offsets[k], *ts = g_buf[s].t.data() + g_sto[s]->host_target_batch_offsets[k];
                    int qle, tle, gtle, gscore, max_off;
                    const int sc = ksw_extend2((int)ql, qs, (int)tl, ts, 5, opt->mat, opt->o_del, opt->e_del, opt->o_ins, opt->e_ins, ext_w, ext_end_bonus, ext_zdrop,
                                               (int)g_sto[s]->host_seed_scores[k], &qle, &tle, &gtle, &gscore, &max_off, ext_use_band);
                    int32_t *t3 = &tri[s][(size_t)k * 3];
                    if (gscore <= 0 || gscore <= sc - pen_clip) { t3[0] = sc; t3[1] = qle; t3[2] = tle; }
                    else { t3[0] = gscore; t3[1] = (int)ql; t3[2] = gtle; }
                }
            }
            std::vector<fork_aln_t> &out = per_read[(size_t)r];
            out.resize(regs.n);
            int is = 0, il = 0;
            for (size_t i = 0; i < regs.n; ++i) {              /* src/bwamem.c:2217-2306 */
                const mem_alnreg_t *a = &regs.a[i];
                fork_aln_t *o = &out[i];
                int32_t part[2][3] = {{0, 0, 0}, {0, 0, 0}};
                if (a->seedlen0 != l_query && a->align_sides > 0) {
                    memcpy(part[a->where_is_long ? 1 : 0], &tri[LONG][(size_t)3 * il++], sizeof(int32_t) * 3);
                    if (a->align_sides == 2) memcpy(part[a->where_is_long ? 0 : 1], &tri[SHORT][(size_t)3 * is++], sizeof(int32_t) * 3);
                    o->score = part[0][0] + part[1][0] - (a->align_sides == 2 ? a->seedlen0 : 0);
                    o->qb = a->query_seed_begin - part[0][1];
                    o->qe = a->query_seed_begin + a->seedlen0 + part[1][1];
                    o->rb = a->target_seed_begin - part[0][2];
                    o->re = a->target_seed_begin + a->seedlen0 + part[1][2];
                    o->truesc = o->score;
                } else {
                    o->score = o->truesc = a->score; o->qb = 0; o->qe = l_query; o->rb = a->target_seed_begin; o->re = a->target_seed_begin + a->seedlen0;
                }
                o->rid = a->rid; o->seedcov = a->seedcov; o->seedlen0 = a->seedlen0; o->w = a->w; o->frac_rep = a->frac_rep; o->pad = 0;
            }
            free(regs.a);
        }
        free(opt);
    }
    uint64_t tot = 0;
    for (int64_t r = 0; r < n_reads; ++r) { n_regs[r] = (uint32_t)per_read[(size_t)r].size(); reg_off[r] = tot; tot += n_regs[r]; }
    if (n_jobs_out) *n_jobs_out = n_jobs_total;
    if ((int64_t)tot > cap_regs) return -1;
    for (int64_t r = 0; r < n_reads; ++r)
        if (n_regs[r]) memcpy(regs_out + reg_off[r], per_read[(size_t)r].data(), sizeof(fork_aln_t) * n_regs[r]);
    return (int64_t)tot;
}

} /* extern "C" */