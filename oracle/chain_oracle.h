/*
 * oracle/chain_oracle.h -- TEST INFRASTRUCTURE ONLY (see chain_oracle.c).
 */
#ifndef CHAIN_ORACLE_H
#define CHAIN_ORACLE_H
#include <stdint.h>

typedef struct {                 /* the mem_opt_t fields the path reads (src/bwamem.h:34-73) */
    int32_t a, b, o_del, e_del, o_ins, e_ins, w, min_seed_len, max_occ, max_chain_gap, min_chain_weight, max_chain_extend;
    float mask_level, drop_ratio;
} chain_opt_t;

typedef struct { int64_t pos; int32_t rid, n, w, kept, first, is_alt; float frac_rep; int32_t seed_off; } chain_rec_t;
typedef struct { int64_t rbeg; int32_t qbeg, len, score, pad; } chain_seed_t;
typedef struct {                 /* one mem_alnreg_t as mem_chain2aln leaves it (src/bwamem.c:1263-1472) */
    int64_t rb_est, re_est, target_seed_begin;
    int32_t qb_est, qe_est, rid, score, truesc, align_sides, where_is_long, query_seed_begin, seedlen0, seedcov, w;
    float frac_rep;
} chain_reg_t;
typedef struct { uint32_t qoff, qlen, toff, tlen, h0; } chain_job_t;
typedef struct { int64_t rb, re; int32_t qb, qe, score, truesc; } chain_aln_t;   /* after the extension results are in */

void chain_opt_default(chain_opt_t *o);
int chain_oracle_read(const chain_opt_t *o, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                      int l_query, uint32_t n_seeds, const uint64_t *rbeg, const int32_t *qbeg_qend, const uint32_t *score, int layout_all,
                      int32_t *n_chains, chain_rec_t *chains, chain_seed_t *cseeds);
int chain_oracle_read_any(const chain_opt_t *o, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                          int l_query, uint32_t n_seeds, const uint64_t *rbeg, const int32_t *qbeg_qend, const uint32_t *score, int layout_all,
                          int32_t *n_chains, chain_rec_t *chains, chain_seed_t *cseeds);
int chain_oracle_flt_seeds(const chain_opt_t *o, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len,
                           const uint8_t *fwd, int l_query, const uint8_t *query, int n_chains, chain_rec_t *chains, chain_seed_t *cseeds);
int chain2aln_oracle_read(const chain_opt_t *o, int64_t l_pac, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len,
                          const uint8_t *fwd, int l_query, const uint8_t *query,
                          int n_chains, const chain_rec_t *chains, const chain_seed_t *cseeds,
                          int32_t *n_regs, chain_reg_t *regs, int cap_regs,
                          int32_t n_jobs[2], chain_job_t *jobs_short, chain_job_t *jobs_long, int cap_jobs,
                          uint8_t *qbuf[2], uint8_t *tbuf[2], uint64_t cap_bytes);
void chain_regs_finish(int l_query, int n_regs, const chain_reg_t *regs, const int32_t *short_triples, const int32_t *long_triples,
                       chain_aln_t *out);
#endif
