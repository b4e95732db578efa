// region.cu -- alignment regions of a read batch -> the records SAM is written from: mem_sort_dedup_patch, is_alt, mem_mark_primary_se,
// mem_approx_mapq_se (reference: src/bwamem.c:620-760, 1690-1716, 2313-2326, 2363, 2459).  The per-read logic is region_core.cuh,
// shared with its host build (tests/host_emul/region_host.cpp); here one lane takes one read at a time (the reference's logic is
// sequential within a read: a sort, then each region against the ones before it), lanes stride over the batch so that the scratch
// (one DP row of the patch alignment per lane) is sized by the grid, not by the batch.
#include "common.h"
#include "region_core.cuh"
#include <algorithm>

using namespace b200region;

static_assert(sizeof(Reg) == sizeof(bwa_b200_alnreg_t) && sizeof(Reg) == 96, "bwa_b200_alnreg_t is region_core's Reg");
static_assert(sizeof(Opt) == sizeof(bwa_b200_region_opt_t), "bwa_b200_region_opt_t is region_core's Opt");

namespace {

struct PackedQuery {            // 4-bit packed read (pack.cu): base i at bits 28 - 4 * (i & 7) of word i >> 3; anything above 3 is N
    const uint32_t *w;
    __host__ __device__ int operator()(int i) const { const int c = (int)((w[i >> 3] >> (28 - 4 * (i & 7))) & 15u); return c > 4 ? 4 : c; }
};
struct PackedRef {              // 2-bit forward reference (bwa_b200_index_attach_ref): base p at bits (~p & 15) * 2 of word p >> 4
    const uint32_t *pac;
    __host__ __device__ int operator()(int64_t p) const { return (int)((pac[p >> 4] >> ((~p & 15) << 1)) & 3u); }
};

constexpr int FIN_THREADS = 128;

__global__ void __launch_bounds__(FIN_THREADS)
finish_kernel(Opt o, int64_t l_pac, const int32_t *__restrict__ ctg_alt, const uint32_t *__restrict__ pac,
              const uint32_t *__restrict__ packed, const uint64_t *__restrict__ woff, uint64_t n_reads,
              const uint64_t *__restrict__ reg_off, Reg *__restrict__ regs, uint32_t *__restrict__ n_out, int32_t *__restrict__ n_pri,
              int64_t first_id, EH *__restrict__ eh_all, uint32_t eh_stride, int32_t *__restrict__ z_all)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, n_lanes = (uint64_t)gridDim.x * blockDim.x;
    EH *eh = eh_all + tid * eh_stride;
    for (uint64_t r = tid; r < n_reads; r += n_lanes) {
        const uint64_t a0 = reg_off[r];
        const int n = (int)(reg_off[r + 1] - a0);
        int np = 0;
        n_out[r] = (uint32_t)finish_read(o, l_pac, ctg_alt, PackedRef{pac}, PackedQuery{packed + woff[r]}, n, regs + a0, first_id + (int64_t)r, &np, eh, z_all + a0);
        n_pri[r] = np;
    }
}

// The call owns a non-blocking stream and takes its buffers from the stream-ordered allocator: nothing here touches the legacy
// default stream or synchronises the device (cudaMalloc / cudaFree / cudaDeviceSynchronize would stall every other handle's streams --
// two batches in flight is the normal way to drive the library), and the pool keeps the memory between calls.
struct CallStream {
    cudaStream_t st = nullptr;
    ~CallStream() { if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); } }
    int open() { return cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess ? 0 : -1; }
};
struct DevBuf {
    void *p = nullptr;
    cudaStream_t st = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, st); }
    int alloc(size_t bytes, cudaStream_t s) { st = s; return cudaMallocAsync(&p, bytes ? bytes : 1, s) == cudaSuccess ? 0 : -1; }
};

} // namespace

extern "C" void bwa_b200_region_opt_default(bwa_b200_region_opt_t *o)
{ // src/bwamem.c:100-140
    if (!o) return;
    o->a = 1; o->b = 4; o->o_del = o->o_ins = 6; o->e_del = o->e_ins = 1; o->w = 100; o->min_seed_len = 19; o->max_chain_gap = 10000;
    o->mask_level = 0.50f; o->mask_level_redun = 0.95f; o->mapQ_coef_len = 50; o->mapQ_coef_fac = (int32_t)log(50.0);
}

extern "C" int bwa_b200_finish_regions_host(const bwa_b200_index_t *idx, int32_t n_ctg, const int32_t *ctg_alt,
                                            const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len, uint64_t n_reads,
                                            const uint64_t *region_off, bwa_b200_alnreg_t *regs, uint32_t *n_regs_out, int32_t *n_pri,
                                            int64_t first_read_id, const bwa_b200_region_opt_t *opt)
{
    if (!idx || !opt || !word_off || !read_len || !region_off || !n_regs_out || !n_pri || (n_reads && (!packed || !regs && region_off[n_reads])))
        { b200::set_error("finish_regions: bad argument"); return BWA_B200_ERR_ARG; }
    if (opt->e_del <= 0 || opt->e_ins <= 0 || opt->a <= 0) { b200::set_error("finish_regions: a, e_del and e_ins must be positive"); return BWA_B200_ERR_ARG; }
    // everything the lanes index with is checked here, before any device work: a region outside its read or outside the text would
    // send the patch alignment out of bounds
    const uint64_t n_regs = n_reads ? region_off[n_reads] : 0, n_words = n_reads ? word_off[n_reads] : 0;
    uint32_t max_len = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        max_len = std::max(max_len, read_len[r]);
        if (region_off[r + 1] < region_off[r] || word_off[r + 1] < word_off[r] + (read_len[r] + 7) / 8)
            { b200::set_error("finish_regions: offsets of read %llu are not increasing / its words do not hold %u bases", (unsigned long long)r, read_len[r]); return BWA_B200_ERR_ARG; }
        for (uint64_t a = region_off[r]; a < region_off[r + 1]; ++a) {
            const bwa_b200_alnreg_t &x = regs[a];
            if (x.qb < 0 || x.qe < x.qb || (uint32_t)x.qe > read_len[r] || x.rb < 0 || x.re < x.rb || (idx->l_pac && (uint64_t)x.re > 2 * idx->l_pac) ||
                (n_ctg > 0 && x.rid >= n_ctg))
                { b200::set_error("finish_regions: region %llu of read %llu lies outside its read, the reference or the contig table", (unsigned long long)(a - region_off[r]), (unsigned long long)r); return BWA_B200_ERR_ARG; }
        }
    }
    B200_CUDA(cudaSetDevice(idx->device));
    if (!idx->d_pac) { b200::set_error("finish_regions: the index has no reference attached (bwa_b200_index_attach_ref)"); return BWA_B200_ERR_ARG; }
    if (n_reads == 0) return BWA_B200_OK;
    int sm = 0;
    B200_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, idx->device));
    const unsigned grid = (unsigned)std::min<uint64_t>((n_reads + FIN_THREADS - 1) / FIN_THREADS, (uint64_t)sm * 8);
    const uint32_t eh_stride = max_len + 2;
    CallStream cs;                         // declared before the buffers: they are released on it, then it is drained and destroyed
    if (cs.open()) { cudaGetLastError(); b200::set_error("finish_regions: cudaStreamCreate failed"); return BWA_B200_ERR_CUDA; }
    cudaStream_t st = cs.st;
    {
        static bool pool_kept[64] = {};    // keep freed blocks in the device's pool instead of returning them to the driver at every sync
        if (idx->device < 64 && !pool_kept[idx->device]) {
            cudaMemPool_t pool;
            uint64_t keep = ~0ull;
            if (cudaDeviceGetDefaultMemPool(&pool, idx->device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            cudaGetLastError();
            pool_kept[idx->device] = true;
        }
    }
    DevBuf d_packed, d_woff, d_off, d_regs, d_n, d_pri, d_alt, d_eh, d_z;
    if (d_packed.alloc(n_words * 4, st) || d_woff.alloc((n_reads + 1) * 8, st) || d_off.alloc((n_reads + 1) * 8, st) || d_regs.alloc(n_regs * sizeof(Reg), st) ||
        d_n.alloc(n_reads * 4, st) || d_pri.alloc(n_reads * 4, st) || d_alt.alloc((size_t)(n_ctg > 0 ? n_ctg : 1) * 4, st) ||
        d_eh.alloc((size_t)grid * FIN_THREADS * eh_stride * sizeof(EH), st) || d_z.alloc(n_regs * 4, st))
        { cudaGetLastError(); b200::set_error("finish_regions: device allocation failed"); return BWA_B200_ERR_CUDA; }
    B200_CUDA(cudaMemcpyAsync(d_packed.p, packed, n_words * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_woff.p, word_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(d_off.p, region_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_regs) B200_CUDA(cudaMemcpyAsync(d_regs.p, regs, n_regs * sizeof(Reg), cudaMemcpyHostToDevice, st));
    const bool have_alt = n_ctg > 0 && ctg_alt;
    if (have_alt) B200_CUDA(cudaMemcpyAsync(d_alt.p, ctg_alt, (size_t)n_ctg * 4, cudaMemcpyHostToDevice, st));
    Opt o;
    memcpy(&o, opt, sizeof(o));
    finish_kernel<<<grid, FIN_THREADS, 0, st>>>(o, (int64_t)idx->l_pac, have_alt ? (const int32_t *)d_alt.p : nullptr, idx->d_pac,
                                                (const uint32_t *)d_packed.p, (const uint64_t *)d_woff.p, n_reads, (const uint64_t *)d_off.p,
                                                (Reg *)d_regs.p, (uint32_t *)d_n.p, (int32_t *)d_pri.p, first_read_id, (EH *)d_eh.p, eh_stride, (int32_t *)d_z.p);
    B200_CUDA(cudaGetLastError());
    if (n_regs) B200_CUDA(cudaMemcpyAsync(regs, d_regs.p, n_regs * sizeof(Reg), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(n_regs_out, d_n.p, n_reads * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(n_pri, d_pri.p, n_reads * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return BWA_B200_OK;
}
