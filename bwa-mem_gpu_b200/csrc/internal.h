// internal.h -- structures shared between the seeding, extension and pipeline translation units.
#pragma once
#include "common.h"
#include <vector>

namespace b200 {

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline
// numbers come from here, not from a profiler).
struct Prof {
    struct Rec { const char *name; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    size_t used = 0;
    void reset() { used = 0; }
    void begin(const char *name, cudaStream_t st)
    {
        if (used == recs.size()) { Rec r; r.name = name; cudaEventCreate(&r.a); cudaEventCreate(&r.b); recs.push_back(r); }
        recs[used].name = name;
        cudaEventRecord(recs[used].a, st);
    }
    void end(cudaStream_t st) { cudaEventRecord(recs[used].b, st); ++used; }
    ~Prof() { for (auto &r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } }
};

#define B200_LAUNCH(prof, name, stream, ...)          \
    do {                                              \
        if (prof) (prof)->begin(name, stream);        \
        __VA_ARGS__;                                  \
        if (prof) (prof)->end(stream);                \
    } while (0)

struct __align__(16) Cand {           // one forward candidate / (after back_kernel) one SMEM
    uint64_t k;                       // x[0]: first SA row
    uint32_t s;                       // x[2]: interval size; 0 after back_kernel = not an SMEM
    uint16_t x;                       // segment start (candidate) / SMEM begin (after back_kernel)
    uint16_t end;                     // exclusive end on the read
};


} // namespace b200

struct bwa_b200_seeder {
    const bwa_b200_index *idx = nullptr;
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    uint64_t max_reads = 0, max_words = 0;
    uint32_t max_read_len = 0, cand_stride = 0, env_stride = 0;
    // inputs (host API)
    uint32_t *d_packed = nullptr, *d_len = nullptr;
    uint64_t *d_woff = nullptr;
    // workspace
    b200::Cand *d_cand = nullptr;
    uint64_t cand_cap = 0;
    uint32_t *d_ncand = nullptr, *d_nsmems = nullptr, *d_nseeds = nullptr, *d_env = nullptr;
    uint64_t *d_seed_off = nullptr, *d_smem_off = nullptr;
    unsigned long long *d_stats = nullptr;        // optional request counters (bwa_b200_seeder_request_counts)
    unsigned long long *d_counters = nullptr;     // [0] next_read, [1] next_seed, [2] total seeds, [3] total smems, [4] widest re-seeding row needed
    // re-seeding (passes 2 and 3 of mem_collect_intv): merged interval list per read (cand2, xstride per read) and the
    // pass-2 candidates (cand3, stride3 per read); [4] / [5] of d_counters = widest row a read needed in either
    b200::Cand *d_cand2 = nullptr, *d_cand3 = nullptr;
    uint64_t cand2_cap = 0, cand3_cap = 0;
    uint32_t xstride = 0, stride3 = 0, *d_ncand2 = nullptr, *d_ncand3 = nullptr, *d_rs_dummy = nullptr;
    uint32_t *d_nsmems1 = nullptr;       // pass-1 SMEM counts of the batch, kept aside: merge_kernel overwrites d_nsmems and the passes may be re-run
    bool rs_saved = false;
    // inputs of the current batch (re-seeding is re-run from them when a read overflows its row)
    const uint32_t *cur_packed = nullptr, *cur_len = nullptr; const uint64_t *cur_woff = nullptr; uint32_t cur_max_len = 0;
    void *d_cub = nullptr;
    size_t cub_bytes = 0;
    // outputs
    uint64_t *d_rbeg = nullptr;
    int2 *d_qq = nullptr;
    uint32_t *d_score = nullptr;
    uint64_t seed_cap = 0;
    uint64_t last_n_reads = 0, last_total = 0;
    int back_grid = 0, loc_grid = 0, fwd_minb = 10, back_minb = 10;
    uint64_t launches = 0;
    bwa_b200_seed_params_t last_p{19, 500};
    b200::Prof *prof = nullptr;
    bool filled = false;          // fill + locate already enqueued for the current batch
    bool redone = false;          // b200_seeder_finish had to redo part of the batch (re-seeding rows widened, or seed arrays grown):
                                  // whatever a caller enqueued behind b200_seeder_run consumed incomplete seeds and must be re-enqueued
    bool narrow_rows = true;      // every BWT row fits 32 bits (seq_len < 2^32) and BWA_B200_WIDE_ROWS is not set
    // pinned staging for the scalar read-backs
    unsigned long long *h_counters = nullptr;
};


struct bwa_b200_extender {
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[8] = {};            // bins of one batch run concurrently on these (forked from / joined to `stream`)
    cudaEvent_t ev_fork = nullptr, ev_join[8] = {};
    int n_side = 0;
    uint64_t max_jobs = 0, max_q = 0, max_t = 0;
    uint8_t *d_q = nullptr, *d_t = nullptr;
    uint32_t *d_qoff = nullptr, *d_qlen = nullptr, *d_toff = nullptr, *d_tlen = nullptr, *d_h0 = nullptr;
    uint32_t *d_keys = nullptr, *d_keys2 = nullptr, *d_vals = nullptr, *d_order = nullptr, *d_range = nullptr;
    bwa_b200_ext_result_t *d_res = nullptr;
    int32_t *d_tri = nullptr;
    unsigned long long *d_cells = nullptr;
    int *d_err = nullptr;
    void *d_cub = nullptr;
    size_t cub_bytes = 0;
    unsigned long long *h_cells = nullptr;
    int *h_err = nullptr;
    uint64_t launches = 0;
    bool pending = false;
    int smem_optin = 0;
    b200::Prof *prof = nullptr;          // an event pair around every launch (bins then run one after another)
    b200::Prof *phase_prof = nullptr;    // one event pair around the whole launch set (bins overlap on the side streams)
    bool own_stream = true;
    int2 *d_intra = nullptr;             // {H, E} column slabs of ext_intra_kernel, one per resident warp
    int intra_grid = 0, intra_max_q = 0;
    bool no_closed_form = false;         // BWA_B200_EXT_NO_CLOSED / bwa_b200_extender_set_closed_form: every job through the kernels
    int pair_unroll = 8; bool pair_no_ring = false;      // BWA_B200_PAIR_NO_RING: per-query state even with a band (A/B switch)
};


// seeding: enqueue the whole batch on s->stream (no host synchronisation while the output arrays
// are large enough); finish reads the totals back and repairs an overflow
int b200_seeder_run(bwa_b200_seeder *s, const uint32_t *d_packed, const uint64_t *d_woff, const uint32_t *d_len,
                    uint64_t n_reads, uint32_t max_len, const bwa_b200_seed_params_t *p);
int b200_seeder_finish(bwa_b200_seeder *s);
// extension over 4-bit packed device sequences
int b200_ext_run_packed(bwa_b200_extender *e, const bwa_b200_ext_params_t *p, uint32_t n,
                        const uint32_t *d_qp, const uint32_t *d_qoff, const uint32_t *d_qlen,
                        const uint32_t *d_tp, const uint32_t *d_toff, const uint32_t *d_tlen,
                        const uint32_t *d_h0, bwa_b200_ext_result_t *d_res, int64_t max_read_len = -1);
