// chain.cu -- seeds -> chains -> extension jobs -> extension -> alignment regions on the device.
//
// The step that sits between the two hot paths in the reference worker: mem_chain / mem_chain_flt / mem_chain2aln
// and the result gathering (src/bwamem.c:404-560,1170-1479,2286-2306).  The per-read logic is chain_core.cuh, the
// same source the CPU tests run against the oracle; this file is the batch machinery around it:
//
//   chain_kernel    one lane per read: chains, kept chains, regions and the job counts of the read.  A read's working
//                   set is its slice [seed_off, seed_off + n_seeds) of flat scratch arrays (the reference keeps a
//                   kbtree and kvecs per read on the heap).
//   scan            one exclusive scan over a 9-field counter record per read gives every read its slot in the region,
//                   chain, job and sequence arrays; the only host round trip of a batch reads the totals back
//                   (the job count is a launch parameter of the extension kernels).
//   jobs_kernel     one lane per read: job descriptors in the reference's batch order (SHORT batch, then LONG batch;
//                   fill_extension, src/bwamem.c:1102-1167), regions compacted to their final place.
//   cut_kernel      one warp per job: query words from the resident packed reads, target words from the resident
//                   2-bit reference, eight bases per lane step (funnel shift + bit spreading, no per-base loop).
//   extension       extend.cu, unchanged: every job of the batch in one sorted launch set.
//   finish_kernel   one lane per read: local-vs-to-end rule and the region arithmetic.
#include "internal.h"
#include "chain_core.cuh"
#include "sw_stripe.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <functional>
#include <vector>

using namespace b200chain;

namespace {

struct Cnt { uint64_t regs, chains, cseeds, n_short, n_long, qw_short, tw_short, qw_long, tw_long; };
struct CntAdd {
    __host__ __device__ Cnt operator()(const Cnt &a, const Cnt &b) const
    {
        return Cnt{a.regs + b.regs, a.chains + b.chains, a.cseeds + b.cseeds, a.n_short + b.n_short, a.n_long + b.n_long,
                   a.qw_short + b.qw_short, a.tw_short + b.tw_short, a.qw_long + b.qw_long, a.tw_long + b.tw_long};
    }
};

struct Scratch {       // per-seed-slot working arrays of chain_core.cuh
    ChainW *ch; int32_t *nxt, *sq, *ord, *kidx; KbNode *nodes; uint64_t *srt;
    bwa_b200_chain_t *chains; bwa_b200_chain_seed_t *cseeds; bwa_b200_region_t *regs;
};

struct SeedView { const uint64_t *rbeg; const int32_t *qq; const uint32_t *score; const uint32_t *n_seeds; const uint64_t *seed_off; uint64_t cap; int layout_all; };

// regions, job counts and sequence words of a read whose kept chains stand in W: the tail of chain_kernel / chain_long_kernel
__device__ __forceinline__ void read_regions(const bwa_b200_chain_params_t &P, const Contigs &ctg, const Scratch &W, uint64_t so, int nc, int l_query, uint32_t r, Cnt &c)
{
    AlnIO ao{l_query, nc, W.chains + so, W.cseeds + so, W.srt + so, W.regs + so};
    int n_short = 0, n_long = 0;
    const int nr = chain2aln_read(P, ctg, ao, &n_short, &n_long);
    c.regs = (uint64_t)nr; c.chains = (uint64_t)nc; c.n_short = (uint64_t)n_short; c.n_long = (uint64_t)n_long;
    // after mem_flt_chained_seeds a chain's seeds are seed_off .. seed_off + n; the detail copy takes the packed prefix
    c.cseeds = nc ? (uint64_t)(W.chains[so + nc - 1].seed_off + W.chains[so + nc - 1].n) : 0;
    read_jobs(W.regs + so, nr, l_query, r, ctg.l_pac, [&](int, int is_long, uint32_t ql, uint32_t tl, uint32_t, const JobAux &) {
        const uint64_t qw = (ql + 7) >> 3, tw = (tl + 7) >> 3;
        if (is_long) { c.qw_long += qw; c.tw_long += tw; } else { c.qw_short += qw; c.tw_short += tw; }
    });
}

__global__ void __launch_bounds__(128)
chain_kernel(uint32_t n_reads, bwa_b200_chain_params_t P, Contigs ctg, SeedView S, const uint32_t *__restrict__ read_len,
             Scratch W, Cnt *__restrict__ cnt, int *__restrict__ err, uint32_t *__restrict__ long_reads)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    Cnt c{0, 0, 0, 0, 0, 0, 0, 0, 0};
    const uint64_t so = S.seed_off[r];
    const uint32_t ns = S.n_seeds[r];
    const int l_query = (int)read_len[r];
    if (ns && so + ns <= S.cap) {
        ReadIO io{S.rbeg + so, S.qq + 2 * so, S.score + so, ns, l_query, S.layout_all,
                  W.ch + so, W.nxt + so, W.sq + 2 * so, W.ord + so, W.kidx + so, W.nodes + (so / 3 + 4ull * r), nodes_needed(ns),
                  W.chains + so, W.cseeds + so};
        int nc = chain_read(P, ctg, io);
        if (nc < 0) { atomicMax(err, 2); nc = 0; }
        if (nc > 0 && flt_seeds_applies(P, l_query)) {          // mem_flt_chained_seeds acts on this read: seedsw_kernel + chain_long_kernel finish it
            long_reads[atomicAdd(err + 1, 1)] = r;
            c.chains = (uint64_t)nc;
        } else read_regions(P, ctg, W, so, nc, l_query, r, c);
    }
    cnt[r] = c;
}

// mem_seed_sw for every chain seed of the long reads chain_kernel listed: one block per read, one seed per group of FOUR lanes -- the
// eight 16-bit lanes of the reference's striped ksw_i16, its vectors in registers (pass_i16_score, sw_stripe.cuh).  The first version
// replayed the striping position by position, one seed per lane with the state in shared memory: 231 ms for the 0.8 M seeds of
// 20 000 reads of 2 kb, 84 % of that batch's step.
constexpr int SEEDSW_NT = 64;
constexpr int SEEDSW_VEC = (SEEDSW_MAX + 7) / 8;             // vectors of the longest window (199 bases): 25
struct SeedSwQ4 { const uint32_t *rd; int qb; __device__ __forceinline__ int operator()(int i) const { return read_base(rd, qb + i); } };
struct SeedSwT2 { const uint32_t *pac; int64_t l_pac, rb; __device__ __forceinline__ int operator()(int i) const { return text_base(pac, l_pac, rb + i); } };
__global__ void __launch_bounds__(SEEDSW_NT)
seedsw_kernel(const int *__restrict__ err, const uint32_t *__restrict__ long_reads, bwa_b200_chain_params_t P, Contigs ctg,
              const uint64_t *__restrict__ seed_off, const uint32_t *__restrict__ read_len, const uint32_t *__restrict__ pac,
              const uint32_t *__restrict__ packed_reads, const uint64_t *__restrict__ word_off, Scratch W, const Cnt *__restrict__ cnt)
{
    __shared__ uint2 tabs[4];
    if (threadIdx.x < 4) {       // bwa_fill_scmat's row of target base t against query A/C/G/T; query N scores -1, a padded position 0
        uint32_t tab = 0;
        for (int q = 0; q < 4; ++q) tab |= (uint32_t)(uint8_t)(int8_t)(q == (int)threadIdx.x ? P.a : -P.b) << (8 * q);
        tabs[threadIdx.x] = make_uint2(tab, 0x008000ffu);
    }
    __syncthreads();
    const int grp = threadIdx.x >> 2, gt = threadIdx.x & 3;
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const uint32_t noe_del2 = (uint32_t)(uint16_t)(int16_t)(-oe_del) * 0x00010001u, ne_del2 = (uint32_t)(uint16_t)(int16_t)(-P.e_del) * 0x00010001u;
    const uint32_t noe_ins2 = (uint32_t)(uint16_t)(int16_t)(-oe_ins) * 0x00010001u, ne_ins2 = (uint32_t)(uint16_t)(int16_t)(-P.e_ins) * 0x00010001u;
    const uint32_t n_long = (uint32_t)err[1];
    for (uint32_t k = blockIdx.x; k < n_long; k += gridDim.x) {
        const uint32_t r = long_reads[k];
        const uint64_t so = seed_off[r];
        const int nc = (int)cnt[r].chains, l_query = (int)read_len[r];
        const int n_cs = W.chains[so + nc - 1].seed_off + W.chains[so + nc - 1].n;
        const uint32_t *rd = packed_reads + word_off[r];
        for (int base = 0; base < n_cs; base += SEEDSW_NT / 4) {
            const int i = base + grp;
            bwa_b200_chain_seed_t s;
            s.rbeg = 0; s.qbeg = 0; s.len = SEEDSW_MAX; s.score = 0; s.pad = 0;
            if (i < n_cs) s = W.cseeds[so + i];
            int qb = 0, qe = 0;
            int64_t rb = 0, re = 0;
            const bool job = i < n_cs && seed_sw_window(ctg, l_query, s, qb, qe, rb, re);
            const int sc = b200sw::pass_i16_score<SEEDSW_VEC>(job, SeedSwQ4{rd, qb}, qe - qb, SeedSwT2{pac, ctg.l_pac, rb}, (int)(re - rb), tabs,
                                                              noe_del2, ne_del2, noe_ins2, ne_ins2, gt);
            if (i < n_cs && gt == 0) W.cseeds[so + i].score = job ? sc : -1;
        }
    }
}

// one lane per long read: the filter on the scored seeds, then what chain_kernel does for the other reads
__global__ void __launch_bounds__(64)
chain_long_kernel(const int *__restrict__ err, const uint32_t *__restrict__ long_reads, bwa_b200_chain_params_t P, Contigs ctg,
                  const uint64_t *__restrict__ seed_off, const uint32_t *__restrict__ read_len, Scratch W, Cnt *__restrict__ cnt)
{
    const uint32_t n_long = (uint32_t)err[1];
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n_long; k += gridDim.x * blockDim.x) {
        const uint32_t r = long_reads[k];
        const uint64_t so = seed_off[r];
        const int nc = (int)cnt[r].chains, l_query = (int)read_len[r];
        flt_seeds_apply(P, l_query, nc, W.chains + so, W.cseeds + so);
        Cnt c{0, 0, 0, 0, 0, 0, 0, 0, 0};
        read_regions(P, ctg, W, so, nc, l_query, r, c);
        cnt[r] = c;
    }
}

__global__ void total_kernel(uint32_t n_reads, const Cnt *cnt, const Cnt *off, Cnt *tot) { *tot = CntAdd()(off[n_reads - 1], cnt[n_reads - 1]); }

struct JobArrays { uint32_t *qoff, *qlen, *toff, *tlen, *h0; JobAux *aux; };

__global__ void __launch_bounds__(128)
jobs_kernel(uint32_t n_reads, int64_t l_pac, const uint32_t *__restrict__ read_len, const uint64_t *__restrict__ seed_off,
            const Cnt *__restrict__ cnt, const Cnt *__restrict__ off, const Cnt *__restrict__ tot, Scratch W,
            bwa_b200_region_t *__restrict__ regions, uint32_t *__restrict__ n_regs, uint64_t *__restrict__ region_off, JobArrays J,
            bwa_b200_chain_t *__restrict__ chains_out, bwa_b200_chain_seed_t *__restrict__ cseeds_out,
            uint32_t *__restrict__ n_chains, uint64_t *__restrict__ chain_off, uint64_t *__restrict__ cseed_off)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const Cnt c = cnt[r], o = off[r];
    const uint64_t so = seed_off[r];
    n_regs[r] = (uint32_t)c.regs; region_off[r] = o.regs;
    if (chains_out) {
        n_chains[r] = (uint32_t)c.chains; chain_off[r] = o.chains; cseed_off[r] = o.cseeds;
        for (uint64_t i = 0; i < c.chains; ++i) chains_out[o.chains + i] = W.chains[so + i];
        for (uint64_t i = 0; i < c.cseeds; ++i) cseeds_out[o.cseeds + i] = W.cseeds[so + i];
    }
    if (c.regs == 0) return;
    const uint64_t n_short_all = tot->n_short;
    uint64_t jb[2] = {o.n_short, n_short_all + o.n_long};
    uint64_t qw[2] = {o.qw_short, tot->qw_short + o.qw_long}, tw[2] = {o.tw_short, tot->tw_short + o.tw_long};
    bwa_b200_region_t *dst = regions + o.regs;
    for (uint64_t i = 0; i < c.regs; ++i) dst[i] = W.regs[so + i];
    read_jobs(dst, (int)c.regs, (int)read_len[r], r, l_pac, [&](int i, int is_long, uint32_t ql, uint32_t tl, uint32_t h0, const JobAux &aux) {
        const uint64_t j = jb[is_long]++;
        J.qoff[j] = (uint32_t)(qw[is_long] << 3); J.qlen[j] = ql; J.toff[j] = (uint32_t)(tw[is_long] << 3); J.tlen[j] = tl; J.h0[j] = h0;
        J.aux[j] = aux;
        qw[is_long] += (ql + 7) >> 3; tw[is_long] += (tl + 7) >> 3;
        if (is_long) dst[i].job_long = (int32_t)(j - n_short_all); else dst[i].job_short = (int32_t)j;
    });
}

// LPJ lanes per job (32 / LPJ jobs per warp): a job of a 150 bp read has at most 17 query and 30 target words, so a whole warp per
// job leaves most lanes idle and pays the job's descriptor loads once per warp
template <int LPJ>
__global__ void __launch_bounds__(256)
cut_jobs_kernel(uint32_t n_jobs, int64_t l_pac, const uint32_t *__restrict__ pac, int64_t pac_words,
                const uint32_t *__restrict__ packed_reads, const uint64_t *__restrict__ word_off, JobArrays J,
                uint32_t *__restrict__ qp, uint32_t *__restrict__ tp)
{
    const uint32_t sub = threadIdx.x & (LPJ - 1);
    for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) / LPJ; j < n_jobs; j += (gridDim.x * blockDim.x) / LPJ) {
        const JobAux a = J.aux[j];
        const uint32_t ql = J.qlen[j], tl = J.tlen[j], read = a.read_flags & AUX_READ;
        const uint64_t w0 = word_off[read];
        const int64_t rd_words = (int64_t)(word_off[read + 1] - w0);
        uint32_t *q = qp + (J.qoff[j] >> 3), *t = tp + (J.toff[j] >> 3);
        for (uint32_t w = sub; w < (ql + 7) >> 3; w += LPJ) q[w] = cut_query_word(packed_reads + w0, rd_words, a, w, ql);
        for (uint32_t w = sub; w < (tl + 7) >> 3; w += LPJ) t[w] = cut_target_word(pac, pac_words, l_pac, a, w, tl);
    }
}

__global__ void __launch_bounds__(128)
finish_kernel(uint32_t n_reads, int pen_clip, const uint32_t *__restrict__ read_len, const uint32_t *__restrict__ n_regs,
              const uint64_t *__restrict__ region_off, const Cnt *__restrict__ tot, const uint32_t *__restrict__ job_qlen,
              const bwa_b200_ext_result_t *__restrict__ res, bwa_b200_region_t *__restrict__ regions)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t nr = n_regs[r];
    const uint64_t n_short_all = tot->n_short;
    bwa_b200_region_t *a = regions + region_off[r];
    for (uint32_t i = 0; i < nr; ++i) {
        int32_t l3[3] = {0, 0, 0}, s3[3] = {0, 0, 0};
        if (a[i].job_long >= 0) { const uint64_t j = n_short_all + (uint64_t)a[i].job_long; ext_triple(res[j], (int)job_qlen[j], pen_clip, l3); }
        if (a[i].job_short >= 0) { const uint64_t j = (uint64_t)a[i].job_short; ext_triple(res[j], (int)job_qlen[j], pen_clip, s3); }
        region_finish(a[i], (int)read_len[r], l3, s3);
    }
}

// ---- compact boundary: 2-bit reads -> the 4-bit layout of bwa_b200_pack_codes (padding 4), bases that are not A/C/G/T patched in
// from a sparse list; regions -> 40-byte records
__global__ void layout2_kernel(uint32_t n_reads, const uint32_t *len_in, uint32_t uniform_len, uint32_t *len,
                               uint64_t *__restrict__ w4, uint64_t *__restrict__ w2)
{ // per-read word counts (slot n_reads = 0, so that the exclusive scans end with the totals)
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    const uint32_t l = r == n_reads ? 0u : (len_in ? len_in[r] : uniform_len);
    if (r < n_reads) len[r] = l;
    w4[r] = ((uint64_t)l + 7) >> 3; w2[r] = ((uint64_t)l + 15) >> 4;
}
__global__ void layout2_uniform_kernel(uint32_t n_reads, uint32_t uniform_len, uint32_t *__restrict__ len, uint64_t *__restrict__ w4, uint64_t *__restrict__ w2)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    if (r < n_reads) len[r] = uniform_len;
    w4[r] = (uint64_t)r * ((uint64_t)(uniform_len + 7) >> 3); w2[r] = (uint64_t)r * ((uint64_t)(uniform_len + 15) >> 4);
}
// one warp per read, one lane per output word: 8 bases = 16 bits of the 2-bit word, each pair of bits spread into a nibble
__global__ void __launch_bounds__(256)
expand2_kernel(uint32_t n_reads, const uint32_t *__restrict__ p2, const uint64_t *__restrict__ w2, const uint32_t *__restrict__ len,
               const uint64_t *__restrict__ w4, uint32_t *__restrict__ p4)
{
    const uint32_t lane = threadIdx.x & 31;
    for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t l = len[r], nw = (l + 7) >> 3;
        const uint32_t *src = p2 + w2[r];
        uint32_t *dst = p4 + w4[r];
        for (uint32_t k = lane; k < nw; k += 32) {
            uint32_t x = (src[k >> 1] >> (16u * (1u - (k & 1u)))) & 0xffffu;
            x = (x | (x << 8)) & 0x00ff00ffu;
            x = (x | (x << 4)) & 0x0f0f0f0fu;
            x = (x | (x << 2)) & 0x33333333u;
            const uint32_t valid = l - 8u * k;                     // bases of this word inside the read (>= 1)
            if (valid < 8u) { const uint32_t pad = 0xffffffffu >> (4u * valid); x = (x & ~pad) | (0x44444444u & pad); }
            dst[k] = x;
        }
    }
}
__global__ void npatch_kernel(uint64_t n_n, const uint64_t *__restrict__ list, uint32_t n_reads, const uint32_t *__restrict__ len,
                              const uint64_t *__restrict__ w4, uint32_t *__restrict__ p4, int *__restrict__ err)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_n) return;
    const uint32_t r = (uint32_t)(list[k] >> 32), pos = (uint32_t)list[k];
    if (r >= n_reads || pos >= len[r]) { atomicMax(err, 3); return; }
    const uint32_t sh = 28u - 4u * (pos & 7u);
    uint32_t *wd = p4 + w4[r] + (pos >> 3);
    atomicAnd(wd, ~(0xfu << sh));
    atomicOr(wd, 4u << sh);
}
__global__ void compact_regions_kernel(uint64_t n, const bwa_b200_region_t *__restrict__ in, bwa_b200_region_compact_t *__restrict__ out, int *__restrict__ err)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const bwa_b200_region_t a = in[k];
    bwa_b200_region_compact_t c;
    c.rb = a.rb; c.rlen = (int32_t)(a.re - a.rb); c.qb = (uint16_t)a.qb; c.qe = (uint16_t)a.qe;
    c.score = a.score; c.truesc = a.truesc; c.seedcov = a.seedcov; c.rid = a.rid;
    c.w = (uint16_t)a.w; c.seedlen0 = (uint16_t)a.seedlen0; c.frac_rep = a.frac_rep;
    if (a.w < 0 || a.w > 0xffff || a.re - a.rb > 0x7fffffffll || a.re < a.rb) atomicMax(err, 4);
    out[k] = c;
}

template <typename T> int grow(T *&p, uint64_t &cap, uint64_t need, uint64_t slack = 0)
{
    if (need <= cap && p) return BWA_B200_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    const uint64_t n = need + slack + 16;
    if (cudaMalloc(&p, n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); b200::set_error("aligner: out of device memory (%llu bytes)", (unsigned long long)(n * sizeof(T))); return BWA_B200_ERR_NOMEM; }
    cap = n;
    return BWA_B200_OK;
}

} // namespace

struct bwa_b200_aligner {
    const bwa_b200_index *idx = nullptr;
    bwa_b200_seeder *seeder = nullptr;
    bwa_b200_extender *ext = nullptr;
    cudaStream_t stream = nullptr;
    int device = 0;
    uint64_t max_reads = 0;
    // contigs (bntann1_t offset / len / is_alt)
    int32_t n_ctg = 0;
    int64_t *d_ctg_off = nullptr; int32_t *d_ctg_len = nullptr, *d_ctg_alt = nullptr;
    // per-seed-slot scratch
    Scratch W{}; uint64_t slot_cap = 0, node_cap = 0;
    // per-read
    Cnt *d_cnt = nullptr, *d_off = nullptr, *d_tot = nullptr, *h_tot = nullptr;
    uint32_t *d_nregs = nullptr, *d_nchains = nullptr; uint64_t *d_region_off = nullptr, *d_chain_off = nullptr, *d_cseed_off = nullptr;
    void *d_cub = nullptr; size_t cub_bytes = 0;
    int *d_err = nullptr, *h_err = nullptr;          // [0] internal error, [1] reads listed in d_long (mem_flt_chained_seeds acts on them), [2] compact boundary
    uint32_t *d_long = nullptr;
    std::vector<uint32_t> skipped;                   // always empty since mem_seed_sw runs on the device (kept for the ABI)
    uint64_t n_long_reads = 0;
    // outputs
    bwa_b200_region_t *d_regions = nullptr; uint64_t region_cap = 0;
    bwa_b200_chain_t *d_chains = nullptr; uint64_t chain_cap = 0;
    bwa_b200_chain_seed_t *d_cseeds = nullptr; uint64_t cseed_cap = 0;
    JobArrays J{}; uint64_t job_cap = 0;
    bwa_b200_ext_result_t *d_res = nullptr;
    uint32_t *d_qp = nullptr, *d_tp = nullptr; uint64_t qp_cap = 0, tp_cap = 0;
    // given seeds (bwa_b200_align_seeds_host)
    uint64_t *g_rbeg = nullptr, *g_seed_off = nullptr; int32_t *g_qq = nullptr; uint32_t *g_score = nullptr, *g_nseeds = nullptr;
    uint64_t g_cap = 0, g_reads = 0;
    // last batch
    uint64_t b_n = 0, b_seeds = 0, b_cells = 0;
    int64_t b_max_len = -1;          // longest read of the batch when the caller knows it (bounds the extension jobs)
    uint32_t b_read_max = 0;         // longest read of the batch, 0 = unknown (decides whether the mem_seed_sw kernels are launched)
    Cnt b_tot{};
    bool b_detail = false;
    uint64_t launches = 0;
    b200::Prof prof; int profiling = 0;
    // pinned host result buffers of bwa_b200_align_host_view (grown geometrically, reused batch after batch)
    uint32_t *p_nregs = nullptr; uint64_t *p_region_off = nullptr; uint64_t p_reads = 0;
    bwa_b200_region_t *p_regions = nullptr; uint64_t p_region_cap = 0;
    // compact boundary (bwa_b200_align_host_compact): 2-bit reads in, 40-byte records out
    uint32_t *d_p2 = nullptr; uint64_t p2_cap = 0; uint64_t *d_woff2 = nullptr; uint64_t *d_nlist = nullptr; uint64_t nlist_cap = 0;
    bwa_b200_region_compact_t *d_cregions = nullptr; uint64_t creg_cap = 0;
    bwa_b200_region_compact_t *p_cregions = nullptr; uint64_t p_creg_cap = 0;
};

extern "C" void bwa_b200_chain_params_default(bwa_b200_chain_params_t *p)
{ // mem_opt_init, src/bwamem.c:107-150
    if (!p) return;
    p->a = 1; p->b = 4; p->o_del = p->o_ins = 6; p->e_del = p->e_ins = 1; p->w = 300;
    p->min_seed_len = 19; p->max_occ = 500; p->max_chain_gap = 10000; p->min_chain_weight = 0; p->max_chain_extend = 1 << 30;
    p->mask_level = 0.50f; p->drop_ratio = 0.50f;
}

extern "C" void bwa_b200_alignments_free(bwa_b200_alignments_t *a)
{
    if (!a) return;
    free(a->n_regions_per_read); free(a->region_off); free(a->regions); free(a->n_chains_per_read); free(a->chain_off); free(a->chains);
    free(a->chain_seed_off); free(a->chain_seeds); free(a->jobs); free(a->qpacked); free(a->tpacked); free(a->job_res);
    memset(a, 0, sizeof(*a));
}

static int aligner_upload_contigs(bwa_b200_aligner *a, int32_t n, const int64_t *off, const int32_t *len, const int32_t *alt)
{
    cudaFree(a->d_ctg_off); cudaFree(a->d_ctg_len); cudaFree(a->d_ctg_alt);
    a->d_ctg_off = nullptr; a->d_ctg_len = nullptr; a->d_ctg_alt = nullptr; a->n_ctg = 0;
    B200_CUDA(cudaMalloc(&a->d_ctg_off, (size_t)n * 8)); B200_CUDA(cudaMalloc(&a->d_ctg_len, (size_t)n * 4)); B200_CUDA(cudaMalloc(&a->d_ctg_alt, (size_t)n * 4));
    B200_CUDA(cudaMemcpy(a->d_ctg_off, off, (size_t)n * 8, cudaMemcpyHostToDevice));
    B200_CUDA(cudaMemcpy(a->d_ctg_len, len, (size_t)n * 4, cudaMemcpyHostToDevice));
    if (alt) B200_CUDA(cudaMemcpy(a->d_ctg_alt, alt, (size_t)n * 4, cudaMemcpyHostToDevice));
    else B200_CUDA(cudaMemset(a->d_ctg_alt, 0, (size_t)n * 4));
    a->n_ctg = n;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_aligner_set_contigs(bwa_b200_aligner_t *a, int32_t n, const int64_t *offset, const int32_t *len, const int32_t *is_alt)
{
    if (!a || n < 1 || !offset || !len) { b200::set_error("aligner_set_contigs: bad argument"); return BWA_B200_ERR_ARG; }
    int64_t end = 0;
    for (int32_t i = 0; i < n; ++i) {
        if (offset[i] != end || len[i] < 0) { b200::set_error("aligner_set_contigs: contigs must tile [0, l_pac) in order"); return BWA_B200_ERR_ARG; }
        end += len[i];
    }
    if ((uint64_t)end != a->idx->l_pac) { b200::set_error("aligner_set_contigs: contig lengths sum to %lld, l_pac is %llu", (long long)end, (unsigned long long)a->idx->l_pac); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(a->device));
    B200_CUDA(cudaStreamSynchronize(a->stream));
    return aligner_upload_contigs(a, n, offset, len, is_alt);
}

extern "C" int bwa_b200_aligner_create(const bwa_b200_index_t *idx, uint64_t max_reads, uint64_t max_words, bwa_b200_aligner_t **out)
{
    if (!idx || !out || !max_reads) { b200::set_error("aligner_create: bad argument"); return BWA_B200_ERR_ARG; }
    if (!idx->d_pac) { b200::set_error("aligner_create: no reference attached (bwa_b200_index_attach_ref)"); return BWA_B200_ERR_ARG; }
    if (max_reads > AUX_READ) { b200::set_error("aligner_create: at most 2^30 - 1 reads per batch"); return BWA_B200_ERR_ARG; }
    bwa_b200_aligner *a = new bwa_b200_aligner();
    a->idx = idx; a->device = idx->device; a->max_reads = max_reads;
    int rc = bwa_b200_seeder_create(idx, max_reads, max_words, &a->seeder);
    if (rc) { delete a; return rc; }
    rc = bwa_b200_extender_create(idx->device, 2 * max_reads, 1024, 1024, &a->ext);
    if (rc) { bwa_b200_seeder_destroy(a->seeder); delete a; return rc; }
    a->stream = a->seeder->stream;
    cudaStreamDestroy(a->ext->stream);          // one stream, in order (the extender forks its bins to side streams itself)
    a->ext->stream = a->stream; a->ext->own_stream = false;
    B200_CUDA(cudaMalloc(&a->d_cnt, max_reads * sizeof(Cnt))); B200_CUDA(cudaMalloc(&a->d_off, max_reads * sizeof(Cnt)));
    B200_CUDA(cudaMalloc(&a->d_tot, sizeof(Cnt))); B200_CUDA(cudaHostAlloc(&a->h_tot, sizeof(Cnt), cudaHostAllocDefault));
    B200_CUDA(cudaMalloc(&a->d_nregs, max_reads * 4)); B200_CUDA(cudaMalloc(&a->d_nchains, max_reads * 4));
    B200_CUDA(cudaMalloc(&a->d_region_off, max_reads * 8)); B200_CUDA(cudaMalloc(&a->d_chain_off, max_reads * 8)); B200_CUDA(cudaMalloc(&a->d_cseed_off, max_reads * 8));
    B200_CUDA(cudaMalloc(&a->d_err, 16)); B200_CUDA(cudaMemset(a->d_err, 0, 16)); B200_CUDA(cudaHostAlloc(&a->h_err, 16, cudaHostAllocDefault));
    a->h_err[0] = a->h_err[1] = a->h_err[2] = a->h_err[3] = 0;
    B200_CUDA(cudaMalloc(&a->d_long, (max_reads ? max_reads : 1) * 4));
    B200_CUDA(cub::DeviceScan::ExclusiveScan(nullptr, a->cub_bytes, a->d_cnt, a->d_off, CntAdd(), Cnt{}, (int)max_reads, a->stream));
    {   // the compact boundary scans word counts with the same scratch
        size_t b2 = 0;
        B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, b2, (uint64_t *)nullptr, (uint64_t *)nullptr, (int)(max_reads + 1), a->stream));
        a->cub_bytes = std::max(a->cub_bytes, b2);
    }
    B200_CUDA(cudaMalloc(&a->d_cub, a->cub_bytes + 16));
    if (idx->l_pac < 0x7fffffffull) {           // default: one sequence [0, l_pac)
        const int64_t off0 = 0; const int32_t len0 = (int32_t)idx->l_pac;
        rc = aligner_upload_contigs(a, 1, &off0, &len0, nullptr);
        if (rc) return rc;
    }
    *out = a;
    return BWA_B200_OK;
}

extern "C" void bwa_b200_aligner_destroy(bwa_b200_aligner_t *a)
{
    if (!a) return;
    cudaSetDevice(a->device);
    cudaStreamSynchronize(a->stream);
    a->seeder->prof = nullptr; a->ext->prof = nullptr; a->ext->phase_prof = nullptr;
    bwa_b200_extender_destroy(a->ext);
    bwa_b200_seeder_destroy(a->seeder);
    cudaFree(a->d_ctg_off); cudaFree(a->d_ctg_len); cudaFree(a->d_ctg_alt);
    cudaFree(a->W.ch); cudaFree(a->W.nxt); cudaFree(a->W.sq); cudaFree(a->W.ord); cudaFree(a->W.kidx); cudaFree(a->W.nodes); cudaFree(a->W.srt);
    cudaFree(a->W.chains); cudaFree(a->W.cseeds); cudaFree(a->W.regs);
    cudaFree(a->d_cnt); cudaFree(a->d_off); cudaFree(a->d_tot); cudaFreeHost(a->h_tot);
    cudaFree(a->d_nregs); cudaFree(a->d_nchains); cudaFree(a->d_region_off); cudaFree(a->d_chain_off); cudaFree(a->d_cseed_off);
    cudaFree(a->d_cub); cudaFree(a->d_err); cudaFree(a->d_long); cudaFreeHost(a->h_err);
    cudaFree(a->d_regions); cudaFree(a->d_chains); cudaFree(a->d_cseeds);
    cudaFree(a->J.qoff); cudaFree(a->J.qlen); cudaFree(a->J.toff); cudaFree(a->J.tlen); cudaFree(a->J.h0); cudaFree(a->J.aux);
    cudaFree(a->d_res); cudaFree(a->d_qp); cudaFree(a->d_tp);
    cudaFree(a->g_rbeg); cudaFree(a->g_seed_off); cudaFree(a->g_qq); cudaFree(a->g_score); cudaFree(a->g_nseeds);
    cudaFreeHost(a->p_nregs); cudaFreeHost(a->p_region_off); cudaFreeHost(a->p_regions);
    cudaFree(a->d_p2); cudaFree(a->d_woff2); cudaFree(a->d_nlist); cudaFree(a->d_cregions); cudaFreeHost(a->p_cregions);
    delete a;
}

static int aligner_ensure_slots(bwa_b200_aligner *a, uint64_t cap)
{
    if (cap <= a->slot_cap) return BWA_B200_OK;
    uint64_t c;
    int rc = 0;
#define G(ptr, mult) (c = 0, cudaFree(ptr), ptr = nullptr, rc |= grow(ptr, c, cap * (mult)))
    G(a->W.ch, 1); G(a->W.nxt, 1); G(a->W.sq, 2); G(a->W.ord, 1); G(a->W.kidx, 1); G(a->W.srt, 1); G(a->W.chains, 1); G(a->W.cseeds, 1); G(a->W.regs, 1);
#undef G
    c = 0; cudaFree(a->W.nodes); a->W.nodes = nullptr;
    rc |= grow(a->W.nodes, c, cap / 3 + 4 * a->max_reads + 8);
    if (rc) { a->slot_cap = 0; return BWA_B200_ERR_NOMEM; }
    a->slot_cap = cap;
    return BWA_B200_OK;
}

// chains -> regions -> jobs -> extension -> regions finished, for the seeds in S; leaves everything on the device
static int aligner_run(bwa_b200_aligner *a, const SeedView &S0, bool seeds_from_seeder, const uint32_t *d_packed, const uint64_t *d_woff,
                       const uint32_t *d_len, uint64_t n_reads, const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep, bool detail)
{
    cudaStream_t st = a->stream;
    b200::Prof *prof = a->profiling ? &a->prof : nullptr;
    a->ext->prof = a->profiling == 1 ? prof : nullptr; a->ext->phase_prof = a->profiling == 2 ? prof : nullptr;
    const uint32_t n = (uint32_t)n_reads;
    if (a->n_ctg < 1) { b200::set_error("align: the reference is longer than 2^31; set the contigs first (bwa_b200_aligner_set_contigs)"); return BWA_B200_ERR_ARG; }
    if (cp->e_del <= 0 || cp->e_ins <= 0 || cp->max_occ < 1) { b200::set_error("align: bad chain parameters"); return BWA_B200_ERR_ARG; }
    Contigs ctg{a->d_ctg_off, a->d_ctg_len, a->d_ctg_alt, a->n_ctg, (int64_t)a->idx->l_pac};
    SeedView S = S0;
    for (int attempt = 0;; ++attempt) {
        if (seeds_from_seeder) {               // the seeder may have re-allocated its output
            bwa_b200_seeder *s = a->seeder;
            S = SeedView{s->d_rbeg, (const int32_t *)s->d_qq, s->d_score, s->d_nseeds, s->d_seed_off, s->seed_cap, 0};
        }
        int rc = aligner_ensure_slots(a, S.cap);
        if (rc) return rc;
        B200_CUDA(cudaMemsetAsync(a->d_err, 0, 8, st));      // also before the retry: attempt 0 may have chained incomplete seed arrays
        B200_LAUNCH(prof, "chain_kernel", st, (chain_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, *cp, ctg, S, d_len, a->W, a->d_cnt, a->d_err, a->d_long)));
        if (a->b_read_max == 0 || flt_seeds_applies(*cp, (int)a->b_read_max)) {     // a batch that may hold reads mem_flt_chained_seeds acts on (length unknown: 0)
            const unsigned g = (unsigned)std::min<uint64_t>(n, (uint64_t)a->seeder->n_sm * 8);
            B200_LAUNCH(prof, "seedsw_kernel", st, (seedsw_kernel<<<g, SEEDSW_NT, 0, st>>>(a->d_err, a->d_long, *cp, ctg, S.seed_off, d_len, a->idx->d_pac,
                                                                                                 d_packed, d_woff, a->W, a->d_cnt)));
            B200_LAUNCH(prof, "chain_long_kernel", st, (chain_long_kernel<<<(unsigned)std::min<uint64_t>((n + 63) / 64, (uint64_t)a->seeder->n_sm * 8), 64, 0, st>>>(
                                                            a->d_err, a->d_long, *cp, ctg, S.seed_off, d_len, a->W, a->d_cnt)));
            a->launches += 2;
        }
        size_t tmp = a->cub_bytes;
        if (prof) prof->begin("chain_scan", st);
        B200_CUDA(cub::DeviceScan::ExclusiveScan(a->d_cub, tmp, a->d_cnt, a->d_off, CntAdd(), Cnt{}, (int)n, st));
        total_kernel<<<1, 1, 0, st>>>(n, a->d_cnt, a->d_off, a->d_tot);
        if (prof) prof->end(st);
        a->launches += 3;
        B200_CUDA(cudaMemcpyAsync(a->h_tot, a->d_tot, sizeof(Cnt), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaMemcpyAsync(a->h_err, a->d_err, 8, cudaMemcpyDeviceToHost, st));
        if (seeds_from_seeder) {
            const uint64_t cap_before = a->seeder->seed_cap;
            rc = b200_seeder_finish(a->seeder);            // synchronises; grows and refills the seed arrays on overflow
            if (rc) return rc;
            a->b_seeds = a->seeder->last_total;
            if ((a->seeder->last_total > cap_before || a->seeder->redone) && attempt == 0) continue;
        }
        B200_CUDA(cudaStreamSynchronize(st));
        break;
    }
    a->skipped.clear();
    if (a->h_err[0]) {
        a->h_err[0] = 0;
        b200::set_error("align: internal capacity exceeded while chaining a read");
        return BWA_B200_ERR_CAPACITY;
    }
    a->n_long_reads = (uint64_t)a->h_err[1];      // reads that went through mem_flt_chained_seeds on the device
    const Cnt T = *a->h_tot;
    a->b_tot = T; a->b_n = n_reads; a->b_detail = detail; a->b_cells = 0;
    const uint64_t n_jobs = T.n_short + T.n_long, qw = T.qw_short + T.qw_long, tw = T.tw_short + T.tw_long;
    if (n_jobs >= 0x7fffffffull || (qw << 3) >= 0xffffffffull || (tw << 3) >= 0xffffffffull) { b200::set_error("align: batch too large for 32-bit job offsets; use smaller batches"); return BWA_B200_ERR_CAPACITY; }
    int rc = grow(a->d_regions, a->region_cap, T.regs, T.regs / 8);
    if (detail) { rc |= grow(a->d_chains, a->chain_cap, T.chains, T.chains / 8); rc |= grow(a->d_cseeds, a->cseed_cap, T.cseeds, T.cseeds / 8); }
    if (n_jobs > a->job_cap || !a->J.qoff) {
        uint64_t c;
        const uint64_t need = n_jobs + n_jobs / 8;
#define G(ptr) (c = 0, cudaFree(ptr), ptr = nullptr, rc |= grow(ptr, c, need))
        G(a->J.qoff); G(a->J.qlen); G(a->J.toff); G(a->J.tlen); G(a->J.h0); G(a->J.aux); G(a->d_res);
#undef G
        a->job_cap = rc ? 0 : need;
    }
    rc |= grow(a->d_qp, a->qp_cap, qw, qw / 8);
    rc |= grow(a->d_tp, a->tp_cap, tw, tw / 8);
    if (rc) return BWA_B200_ERR_NOMEM;
    B200_LAUNCH(prof, "jobs_kernel", st,
        (jobs_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, (int64_t)a->idx->l_pac, d_len, S.seed_off, a->d_cnt, a->d_off, a->d_tot, a->W, a->d_regions,
                                                      a->d_nregs, a->d_region_off, a->J, detail ? a->d_chains : nullptr, a->d_cseeds, a->d_nchains,
                                                      a->d_chain_off, a->d_cseed_off)));
    a->launches += 1;
    if (n_jobs) {
        static const int lpj = getenv("BWA_B200_CUT_LPJ") ? atoi(getenv("BWA_B200_CUT_LPJ")) : 16;     // C2: 32 -> 0.53 ms, 16 -> 0.46, 8 -> 0.46
        auto kern = lpj <= 8 ? cut_jobs_kernel<8> : (lpj <= 16 ? cut_jobs_kernel<16> : cut_jobs_kernel<32>);
        const uint64_t groups = (n_jobs * (uint64_t)(lpj <= 8 ? 8 : (lpj <= 16 ? 16 : 32)) + 31) / 32;      // warps' worth of work
        const uint64_t warps = groups < (uint64_t)a->seeder->n_sm * 64 ? groups : (uint64_t)a->seeder->n_sm * 64;
        B200_LAUNCH(prof, "cut_jobs_kernel", st,
            (kern<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>((uint32_t)n_jobs, (int64_t)a->idx->l_pac, a->idx->d_pac, (int64_t)((a->idx->l_pac + 15) / 16 + 1),
                                                              d_packed, d_woff, a->J, a->d_qp, a->d_tp)));
        a->launches += 1;
        B200_CUDA(cudaGetLastError());
        rc = b200_ext_run_packed(a->ext, ep, (uint32_t)n_jobs, a->d_qp, a->J.qoff, a->J.qlen, a->d_tp, a->J.toff, a->J.tlen, a->J.h0, a->d_res, a->b_max_len);
        if (rc) return rc;
    }
    B200_LAUNCH(prof, "finish_kernel", st,
        (finish_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, ep->pen_clip, d_len, a->d_nregs, a->d_region_off, a->d_tot, a->J.qlen, a->d_res, a->d_regions)));
    a->launches += 1;
    B200_CUDA(cudaGetLastError());
    B200_CUDA(cudaStreamSynchronize(st));
    if (n_jobs) {
        rc = bwa_b200_extend_wait(a->ext);
        if (rc) return rc;
    }
    return BWA_B200_OK;
}

extern "C" int bwa_b200_align_device(bwa_b200_aligner_t *a, const uint32_t *dev_packed, const uint64_t *dev_word_off, const uint32_t *dev_read_len,
                                     uint64_t n_reads, uint32_t max_read_len, const bwa_b200_seed_params_t *sp,
                                     const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep)
{
    if (!a || !sp || !cp || !ep || (n_reads && (!dev_packed || !dev_word_off || !dev_read_len))) { b200::set_error("align_device: bad argument"); return BWA_B200_ERR_ARG; }
    if (n_reads > a->max_reads) { b200::set_error("align: %llu reads > capacity", (unsigned long long)n_reads); return BWA_B200_ERR_CAPACITY; }
    if (sp->max_occ <= 0 || sp->max_occ != cp->max_occ || sp->min_seed_len != cp->min_seed_len) { b200::set_error("align: seed and chain parameters must agree on min_seed_len and a positive max_occ"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(a->device));
    a->b_n = 0; a->b_tot = Cnt{}; a->b_seeds = 0;
    if (n_reads == 0) return BWA_B200_OK;
    b200::Prof *prof = a->profiling ? &a->prof : nullptr;
    a->seeder->prof = prof;
    if (prof) prof->reset();
    a->b_max_len = (int64_t)max_read_len; a->b_read_max = max_read_len;
    int rc = b200_seeder_run(a->seeder, dev_packed, dev_word_off, dev_read_len, n_reads, max_read_len, sp);
    if (rc) return rc;
    return aligner_run(a, SeedView{}, true, dev_packed, dev_word_off, dev_read_len, n_reads, cp, ep, a->b_detail);
}

extern "C" int bwa_b200_aligner_skipped_reads(bwa_b200_aligner_t *a, uint64_t *n, const uint32_t **read_idx)
{
    if (!a || !n) return BWA_B200_ERR_ARG;
    *n = a->skipped.size();
    if (read_idx) *read_idx = a->skipped.empty() ? nullptr : a->skipped.data();
    return BWA_B200_OK;
}

extern "C" int bwa_b200_align_device_view(bwa_b200_aligner_t *a, bwa_b200_align_view_t *v)
{
    if (!a || !v) return BWA_B200_ERR_ARG;
    memset(v, 0, sizeof(*v));
    v->n_reads = a->b_n; v->n_regions = a->b_tot.regs; v->n_jobs_short = a->b_tot.n_short; v->n_jobs_long = a->b_tot.n_long; v->n_seeds = a->b_seeds;
    v->cells = a->b_n && (a->b_tot.n_short + a->b_tot.n_long) ? bwa_b200_extender_last_cells(a->ext) : 0;
    v->closed_form_jobs = v->cells || (a->b_n && (a->b_tot.n_short + a->b_tot.n_long)) ? a->ext->h_cells[1] : 0;
    v->n_regions_per_read = a->d_nregs; v->region_off = a->d_region_off; v->regions = a->d_regions;
    return BWA_B200_OK;
}

template <typename T> static T *host_copy(const T *d, uint64_t n, cudaStream_t st, bool *ok)
{
    T *h = (T *)malloc((n ? n : 1) * sizeof(T));
    if (!h) { *ok = false; return nullptr; }
    if (n && cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, st) != cudaSuccess) *ok = false;
    return h;
}

static int aligner_download(bwa_b200_aligner *a, int want_detail, bwa_b200_alignments_t *out)
{
    const Cnt &T = a->b_tot;
    const uint64_t n = a->b_n;
    cudaStream_t st = a->stream;
    bool ok = true;
    out->n_reads = n; out->n_regions = T.regs;
    out->n_regions_per_read = host_copy(a->d_nregs, n, st, &ok);
    out->region_off = host_copy(a->d_region_off, n, st, &ok);
    out->regions = host_copy(a->d_regions, T.regs, st, &ok);
    uint32_t *qoff = nullptr, *qlen = nullptr, *toff = nullptr, *tlen = nullptr, *h0 = nullptr;
    const uint64_t nj = T.n_short + T.n_long;
    if (want_detail) {
        out->n_chains = T.chains; out->n_chain_seeds = T.cseeds; out->n_jobs_short = T.n_short; out->n_jobs_long = T.n_long;
        out->q_words = T.qw_short + T.qw_long; out->t_words = T.tw_short + T.tw_long;
        out->n_chains_per_read = host_copy(a->d_nchains, n, st, &ok);
        out->chain_off = host_copy(a->d_chain_off, n, st, &ok);
        out->chain_seed_off = host_copy(a->d_cseed_off, n, st, &ok);
        out->chains = host_copy(a->d_chains, T.chains, st, &ok);
        out->chain_seeds = host_copy(a->d_cseeds, T.cseeds, st, &ok);
        out->qpacked = host_copy(a->d_qp, out->q_words, st, &ok);
        out->tpacked = host_copy(a->d_tp, out->t_words, st, &ok);
        out->job_res = host_copy(a->d_res, nj, st, &ok);
        qoff = host_copy(a->J.qoff, nj, st, &ok); qlen = host_copy(a->J.qlen, nj, st, &ok); toff = host_copy(a->J.toff, nj, st, &ok);
        tlen = host_copy(a->J.tlen, nj, st, &ok); h0 = host_copy(a->J.h0, nj, st, &ok);
        out->jobs = (bwa_b200_job_t *)malloc((nj ? nj : 1) * sizeof(bwa_b200_job_t));
        if (!out->jobs) ok = false;
    }
    if (cudaStreamSynchronize(st) != cudaSuccess) ok = false;
    if (ok && want_detail)
        for (uint64_t j = 0; j < nj; ++j) out->jobs[j] = bwa_b200_job_t{qoff[j], qlen[j], toff[j], tlen[j], h0[j]};
    free(qoff); free(qlen); free(toff); free(tlen); free(h0);
    if (!ok) { bwa_b200_alignments_free(out); b200::set_error("align: copying the results back failed (host memory or CUDA error)"); return BWA_B200_ERR_NOMEM; }
    return BWA_B200_OK;
}

static int aligner_upload_reads(bwa_b200_aligner *a, const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len, uint64_t n_reads, uint32_t *max_len)
{
    bwa_b200_seeder *s = a->seeder;
    if (n_reads > a->max_reads || word_off[n_reads] > s->max_words) { b200::set_error("align: batch exceeds the aligner's capacity"); return BWA_B200_ERR_CAPACITY; }
    uint32_t m = 0;
    for (uint64_t r = 0; r < n_reads; ++r) m = read_len[r] > m ? read_len[r] : m;
    *max_len = m;
    B200_CUDA(cudaMemcpyAsync(s->d_packed, packed, word_off[n_reads] * 4, cudaMemcpyHostToDevice, a->stream));
    B200_CUDA(cudaMemcpyAsync(s->d_woff, word_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, a->stream));
    B200_CUDA(cudaMemcpyAsync(s->d_len, read_len, n_reads * 4, cudaMemcpyHostToDevice, a->stream));
    return BWA_B200_OK;
}

extern "C" int bwa_b200_align_host(bwa_b200_aligner_t *a, const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len,
                                   uint64_t n_reads, const bwa_b200_seed_params_t *sp, const bwa_b200_chain_params_t *cp,
                                   const bwa_b200_ext_params_t *ep, int want_detail, bwa_b200_alignments_t *out)
{
    if (!a || !sp || !cp || !ep || !out || (n_reads && (!packed || !word_off || !read_len))) { b200::set_error("align_host: bad argument"); return BWA_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (n_reads == 0) return BWA_B200_OK;
    B200_CUDA(cudaSetDevice(a->device));
    uint32_t max_len = 0;
    int rc = aligner_upload_reads(a, packed, word_off, read_len, n_reads, &max_len);
    if (rc) return rc;
    bwa_b200_seeder *s = a->seeder;
    a->b_detail = want_detail != 0;
    rc = bwa_b200_align_device(a, s->d_packed, s->d_woff, s->d_len, n_reads, max_len, sp, cp, ep);
    a->b_detail = false;
    if (rc) return rc;
    return aligner_download(a, want_detail, out);
}

// regions into pinned buffers owned by the aligner: no allocation and no pageable staging per batch (the reference's
// gasal_res_t arrays are pinned for the same reason, GASAL2/src/res.cpp)
extern "C" int bwa_b200_align_host_view(bwa_b200_aligner_t *a, const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len,
                                        uint64_t n_reads, const bwa_b200_seed_params_t *sp, const bwa_b200_chain_params_t *cp,
                                        const bwa_b200_ext_params_t *ep, uint64_t *n_regions, const uint32_t **n_regions_per_read,
                                        const uint64_t **region_off, const bwa_b200_region_t **regions)
{
    if (!a || !sp || !cp || !ep || !n_regions || !n_regions_per_read || !region_off || !regions || (n_reads && (!packed || !word_off || !read_len))) {
        b200::set_error("align_host_view: bad argument"); return BWA_B200_ERR_ARG;
    }
    *n_regions = 0; *n_regions_per_read = nullptr; *region_off = nullptr; *regions = nullptr;
    if (n_reads == 0) return BWA_B200_OK;
    B200_CUDA(cudaSetDevice(a->device));
    uint32_t max_len = 0;
    int rc = aligner_upload_reads(a, packed, word_off, read_len, n_reads, &max_len);
    if (rc) return rc;
    bwa_b200_seeder *s = a->seeder;
    rc = bwa_b200_align_device(a, s->d_packed, s->d_woff, s->d_len, n_reads, max_len, sp, cp, ep);
    if (rc) return rc;
    const uint64_t nr = a->b_tot.regs;
    if (n_reads > a->p_reads) {
        cudaFreeHost(a->p_nregs); cudaFreeHost(a->p_region_off); a->p_nregs = nullptr; a->p_region_off = nullptr; a->p_reads = 0;
        const uint64_t c = std::max<uint64_t>(n_reads, a->max_reads);
        B200_CUDA(cudaHostAlloc(&a->p_nregs, c * 4, cudaHostAllocDefault));
        B200_CUDA(cudaHostAlloc(&a->p_region_off, c * 8, cudaHostAllocDefault));
        a->p_reads = c;
    }
    if (nr > a->p_region_cap) {
        cudaFreeHost(a->p_regions); a->p_regions = nullptr; a->p_region_cap = 0;
        const uint64_t c = nr + nr / 4 + 1024;
        B200_CUDA(cudaHostAlloc(&a->p_regions, c * sizeof(bwa_b200_region_t), cudaHostAllocDefault));
        a->p_region_cap = c;
    }
    cudaStream_t st = a->stream;
    B200_CUDA(cudaMemcpyAsync(a->p_nregs, a->d_nregs, n_reads * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(a->p_region_off, a->d_region_off, n_reads * 8, cudaMemcpyDeviceToHost, st));
    if (nr) B200_CUDA(cudaMemcpyAsync(a->p_regions, a->d_regions, nr * sizeof(bwa_b200_region_t), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    *n_regions = nr; *n_regions_per_read = a->p_nregs; *region_off = a->p_region_off; *regions = a->p_regions;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_align_seeds_host(bwa_b200_aligner_t *a, const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len,
                                         uint64_t n_reads, const bwa_b200_seeds_t *seeds, int layout_all, const bwa_b200_chain_params_t *cp,
                                         const bwa_b200_ext_params_t *ep, int want_detail, bwa_b200_alignments_t *out)
{
    if (!a || !seeds || !cp || !ep || !out || (n_reads && (!packed || !word_off || !read_len))) { b200::set_error("align_seeds_host: bad argument"); return BWA_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (n_reads == 0) return BWA_B200_OK;
    if (seeds->n_reads != n_reads || !seeds->n_seeds_per_read || !seeds->seed_off || (seeds->n_seeds && (!seeds->rbeg || !seeds->qbeg_qend || !seeds->score))) {
        b200::set_error("align_seeds_host: the seeds do not describe this batch"); return BWA_B200_ERR_ARG;
    }
    B200_CUDA(cudaSetDevice(a->device));
    uint32_t max_len = 0;
    int rc = aligner_upload_reads(a, packed, word_off, read_len, n_reads, &max_len);
    if (rc) return rc;
    a->b_read_max = max_len;
    a->b_max_len = -1;            // given seeds: a seed's score (h0) is the caller's, not bounded by the read length
    const uint64_t ns = seeds->n_seeds;
    if (ns > a->g_cap || !a->g_rbeg) {
        uint64_t c;
        c = 0; cudaFree(a->g_rbeg); a->g_rbeg = nullptr; rc |= grow(a->g_rbeg, c, ns, ns / 8);
        c = 0; cudaFree(a->g_qq); a->g_qq = nullptr; rc |= grow(a->g_qq, c, 2 * ns, ns / 4);
        c = 0; cudaFree(a->g_score); a->g_score = nullptr; rc |= grow(a->g_score, c, ns, ns / 8);
        if (rc) { a->g_cap = 0; return BWA_B200_ERR_NOMEM; }
        a->g_cap = ns + ns / 8;
    }
    if (n_reads > a->g_reads || !a->g_nseeds) {
        uint64_t c;
        c = 0; cudaFree(a->g_nseeds); a->g_nseeds = nullptr; rc |= grow(a->g_nseeds, c, n_reads);
        c = 0; cudaFree(a->g_seed_off); a->g_seed_off = nullptr; rc |= grow(a->g_seed_off, c, n_reads);
        if (rc) { a->g_reads = 0; return BWA_B200_ERR_NOMEM; }
        a->g_reads = n_reads;
    }
    cudaStream_t st = a->stream;
    if (ns) {
        B200_CUDA(cudaMemcpyAsync(a->g_rbeg, seeds->rbeg, ns * 8, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(a->g_qq, seeds->qbeg_qend, ns * 8, cudaMemcpyHostToDevice, st));
        B200_CUDA(cudaMemcpyAsync(a->g_score, seeds->score, ns * 4, cudaMemcpyHostToDevice, st));
    }
    B200_CUDA(cudaMemcpyAsync(a->g_nseeds, seeds->n_seeds_per_read, n_reads * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(a->g_seed_off, seeds->seed_off, n_reads * 8, cudaMemcpyHostToDevice, st));
    b200::Prof *prof = a->profiling ? &a->prof : nullptr;
    if (prof) prof->reset();
    a->b_seeds = ns;
    bwa_b200_seeder *s = a->seeder;
    SeedView S{a->g_rbeg, a->g_qq, a->g_score, a->g_nseeds, a->g_seed_off, ns, layout_all != 0};
    rc = aligner_run(a, S, false, s->d_packed, s->d_woff, s->d_len, n_reads, cp, ep, want_detail != 0);
    if (rc) return rc;
    return aligner_download(a, want_detail, out);
}

// The compact boundary.  reserve(n_regions) returns where the batch's records go (pinned host memory), or nullptr to refuse.
int b200_align_compact(bwa_b200_aligner *a, const uint32_t *packed2, const uint32_t *read_len, uint32_t uniform_len, uint64_t n_reads,
                       const uint64_t *n_list, uint64_t n_n, const bwa_b200_seed_params_t *sp, const bwa_b200_chain_params_t *cp,
                       const bwa_b200_ext_params_t *ep, uint32_t *dst_nregs, uint64_t *n_regions,
                       const std::function<bwa_b200_region_compact_t *(uint64_t)> &reserve)
{
    bwa_b200_seeder *s = a->seeder;
    if (n_reads > a->max_reads) { b200::set_error("align_compact: %llu reads > capacity", (unsigned long long)n_reads); return BWA_B200_ERR_CAPACITY; }
    uint64_t tot2 = 0, tot4 = 0;
    uint32_t max_len = uniform_len;
    if (read_len) {
        max_len = 0;
        for (uint64_t r = 0; r < n_reads; ++r) { const uint32_t l = read_len[r]; tot2 += ((uint64_t)l + 15) >> 4; tot4 += ((uint64_t)l + 7) >> 3; max_len = l > max_len ? l : max_len; }
    } else { tot2 = n_reads * (((uint64_t)uniform_len + 15) >> 4); tot4 = n_reads * (((uint64_t)uniform_len + 7) >> 3); }
    if (max_len > 0xffffu) { b200::set_error("align_compact: reads beyond 65535 bases do not fit the compact record; use bwa_b200_align_host_view"); return BWA_B200_ERR_ARG; }
    if (tot4 > s->max_words) { b200::set_error("align_compact: batch exceeds the aligner's capacity"); return BWA_B200_ERR_CAPACITY; }
    B200_CUDA(cudaSetDevice(a->device));
    cudaStream_t st = a->stream;
    int rc = grow(a->d_p2, a->p2_cap, tot2, tot2 / 8);
    if (!a->d_woff2) { uint64_t c = 0; rc |= grow(a->d_woff2, c, a->max_reads + 1); }
    if (n_n) rc |= grow(a->d_nlist, a->nlist_cap, n_n, n_n / 4);
    if (rc) return BWA_B200_ERR_NOMEM;
    const uint32_t n = (uint32_t)n_reads;
    B200_CUDA(cudaMemcpyAsync(a->d_p2, packed2, tot2 * 4, cudaMemcpyHostToDevice, st));
    if (n_n) B200_CUDA(cudaMemcpyAsync(a->d_nlist, n_list, n_n * 8, cudaMemcpyHostToDevice, st));
    if (read_len) {
        // lengths ride in the seeder's own length array; the word counts are scanned in place into both offset arrays
        B200_CUDA(cudaMemcpyAsync(s->d_len, read_len, n_reads * 4, cudaMemcpyHostToDevice, st));
        layout2_kernel<<<(n + 256) / 256, 256, 0, st>>>(n, s->d_len, 0, s->d_len, s->d_woff, a->d_woff2);
        size_t tmp = a->cub_bytes;
        B200_CUDA(cub::DeviceScan::ExclusiveSum(a->d_cub, tmp, s->d_woff, s->d_woff, (int)(n + 1), st));
        tmp = a->cub_bytes;
        B200_CUDA(cub::DeviceScan::ExclusiveSum(a->d_cub, tmp, a->d_woff2, a->d_woff2, (int)(n + 1), st));
        a->launches += 3;
    } else {
        layout2_uniform_kernel<<<(n + 256) / 256, 256, 0, st>>>(n, uniform_len, s->d_len, s->d_woff, a->d_woff2);
        a->launches += 1;
    }
    {
        const uint64_t warps = n_reads < (uint64_t)s->n_sm * 64 ? n_reads : (uint64_t)s->n_sm * 64;
        expand2_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(n, a->d_p2, a->d_woff2, s->d_len, s->d_woff, s->d_packed);
        a->launches += 1;
    }
    int *d_err2 = a->d_err + 2;
    B200_CUDA(cudaMemsetAsync(d_err2, 0, 4, st));
    if (n_n) { npatch_kernel<<<(unsigned)((n_n + 255) / 256), 256, 0, st>>>(n_n, a->d_nlist, n, s->d_len, s->d_woff, s->d_packed, d_err2); a->launches += 1; }
    B200_CUDA(cudaGetLastError());
    rc = bwa_b200_align_device(a, s->d_packed, s->d_woff, s->d_len, n_reads, max_len, sp, cp, ep);
    if (rc) return rc;
    const uint64_t nr = a->b_tot.regs;
    rc = grow(a->d_cregions, a->creg_cap, nr, nr / 8);
    if (rc) return rc;
    if (nr) { compact_regions_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, st>>>(nr, a->d_regions, a->d_cregions, d_err2); a->launches += 1; }
    bwa_b200_region_compact_t *dst = reserve(nr);
    if (!dst && nr) { b200::set_error("align_compact: no room for %llu region records", (unsigned long long)nr); return BWA_B200_ERR_CAPACITY; }
    B200_CUDA(cudaMemcpyAsync(dst_nregs, a->d_nregs, n_reads * 4, cudaMemcpyDeviceToHost, st));
    if (nr) B200_CUDA(cudaMemcpyAsync(dst, a->d_cregions, nr * sizeof(bwa_b200_region_compact_t), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(a->h_err + 2, d_err2, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    if (a->h_err[2]) {
        const int e = a->h_err[2]; a->h_err[2] = 0;
        b200::set_error(e == 3 ? "align_compact: an entry of the N list names a read or a position outside the batch" : "align_compact: a region does not fit the compact record (band beyond 65535)");
        return BWA_B200_ERR_ARG;
    }
    *n_regions = nr;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_align_host_compact(bwa_b200_aligner_t *a, const uint32_t *packed2, const uint32_t *read_len, uint32_t uniform_len,
                                           uint64_t n_reads, const uint64_t *n_list, uint64_t n_n, const bwa_b200_seed_params_t *sp,
                                           const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep, uint64_t *n_regions,
                                           const uint32_t **n_regions_per_read, const bwa_b200_region_compact_t **regions)
{
    if (!a || !sp || !cp || !ep || !n_regions || !n_regions_per_read || !regions || (n_reads && !packed2) || (n_n && !n_list)) {
        b200::set_error("align_host_compact: bad argument"); return BWA_B200_ERR_ARG;
    }
    *n_regions = 0; *n_regions_per_read = nullptr; *regions = nullptr;
    if (n_reads == 0) return BWA_B200_OK;
    B200_CUDA(cudaSetDevice(a->device));
    if (n_reads > a->p_reads) {
        cudaFreeHost(a->p_nregs); cudaFreeHost(a->p_region_off); a->p_nregs = nullptr; a->p_region_off = nullptr; a->p_reads = 0;
        const uint64_t c = std::max<uint64_t>(n_reads, a->max_reads);
        B200_CUDA(cudaHostAlloc(&a->p_nregs, c * 4, cudaHostAllocDefault));
        B200_CUDA(cudaHostAlloc(&a->p_region_off, c * 8, cudaHostAllocDefault));
        a->p_reads = c;
    }
    bool alloc_failed = false;
    auto reserve = [&](uint64_t nr) -> bwa_b200_region_compact_t * {
        if (nr > a->p_creg_cap) {
            cudaFreeHost(a->p_cregions); a->p_cregions = nullptr; a->p_creg_cap = 0;
            const uint64_t c = nr + nr / 4 + 1024;
            if (cudaHostAlloc(&a->p_cregions, c * sizeof(bwa_b200_region_compact_t), cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); alloc_failed = true; return nullptr; }
            a->p_creg_cap = c;
        }
        return a->p_cregions;
    };
    const int rc = b200_align_compact(a, packed2, read_len, uniform_len, n_reads, n_list, n_n, sp, cp, ep, a->p_nregs, n_regions, reserve);
    if (rc) return alloc_failed ? BWA_B200_ERR_NOMEM : rc;
    *n_regions_per_read = a->p_nregs; *regions = a->p_cregions;
    return BWA_B200_OK;
}

extern "C" void *bwa_b200_aligner_stream(bwa_b200_aligner_t *a) { return a ? (void *)a->stream : nullptr; }
extern "C" uint64_t bwa_b200_aligner_launches(const bwa_b200_aligner_t *a) { return a ? a->launches + a->seeder->launches + a->ext->launches : 0; }
extern "C" int bwa_b200_aligner_profile(bwa_b200_aligner_t *a, int enable)
{
    if (!a) return BWA_B200_ERR_ARG;
    a->profiling = enable == 2 ? 2 : (enable != 0);
    return BWA_B200_OK;
}
extern "C" int bwa_b200_aligner_kernel_times(bwa_b200_aligner_t *a, const char **names, float *ms, int cap)
{
    if (!a) return BWA_B200_ERR_ARG;
    cudaSetDevice(a->device);
    cudaStreamSynchronize(a->stream);
    int n = 0;
    for (size_t i = 0; i < a->prof.used && n < cap; ++i, ++n) {
        float t = 0;
        cudaEventElapsedTime(&t, a->prof.recs[i].a, a->prof.recs[i].b);
        names[n] = a->prof.recs[i].name; ms[n] = t;
    }
    return n;
}
