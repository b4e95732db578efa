// pipeline.cu -- fused seed -> extend pass: seeds stay in HBM, extension jobs are cut on the
// device from a resident 2-bit reference, one record per read is produced.
//
// Job construction follows the reference's mem_chain2aln for a chain of one seed
// (cal_max_gap src/bwamem.c:996-1002; rmax window and strand clamp :1180-1201; left/right job
// shapes and h0 = seed_len * a :1356-1424); see include/bwamem_b200.h for what is and is not
// claimed.  Kernels here are plumbing around the two hot kernels (seed.cu, extend.cu): a
// per-read selection, a word-parallel sequence cutter and a gather.
#include "internal.h"
#include "chain_core.cuh"

namespace {

struct Rules { int a, o_del, e_del, o_ins, e_ins, w; };

__device__ __forceinline__ int max_gap(const Rules &r, int qlen)
{
    int l_del = (int)((double)(qlen * r.a - r.o_del) / r.e_del + 1.);
    int l_ins = (int)((double)(qlen * r.a - r.o_ins) / r.e_ins + 1.);
    int l = l_del > l_ins ? l_del : l_ins;
    l = l > 1 ? l : 1;
    return l < (r.w << 1) ? l : (r.w << 1);
}

using b200chain::JobAux;          // {start, qfrom, read | AUX_LEFT | AUX_REV}: shared with the chaining stage (chain_core.cuh)
using b200chain::AUX_LEFT;
using b200chain::AUX_REV;

// one lane per read: choose the seed, shape both jobs
__global__ void choose_kernel(uint32_t n_reads, int64_t l_pac, Rules R, const uint32_t *__restrict__ read_len,
                              const uint32_t *__restrict__ n_seeds, const uint64_t *__restrict__ seed_off,
                              const uint64_t *__restrict__ rbeg, const int2 *__restrict__ qq, uint64_t seed_cap,
                              uint32_t qstride, uint32_t tstride,
                              bwa_b200_read_result_t *__restrict__ out,
                              uint32_t *__restrict__ jq_len, uint32_t *__restrict__ jt_len, uint32_t *__restrict__ j_h0,
                              uint32_t *__restrict__ jq_off, uint32_t *__restrict__ jt_off, JobAux *__restrict__ aux)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int len = (int)read_len[r];
    const uint32_t ns = n_seeds[r];
    const uint64_t so = seed_off[r];
    int64_t best = -1; int best_len = -1;
    if (so + ns <= seed_cap)
        for (uint32_t i = 0; i < ns; ++i) {
            int2 q = qq[so + i];
            int sl = q.y - q.x;
            int64_t rb = (int64_t)rbeg[so + i];
            if (rb < l_pac && rb + sl > l_pac) continue;           // bridges forward/reverse (bns_intv2rid < 0)
            if (sl > best_len) { best_len = sl; best = (int64_t)i; }
        }
    bwa_b200_read_result_t o;
    o.seed_rbeg = -1; o.seed_qbeg = -1; o.seed_qend = -1; o.n_seeds = (int32_t)ns; o.h0 = 0;
    o.left = bwa_b200_ext_result_t{0, 0, 0, 0, 0, 0}; o.right = o.left;
    uint32_t lq = 0, lt = 0, rq = 0, rt = 0, h0 = 0;
    JobAux la{0, 0, r | AUX_LEFT}, ra{0, 0, r};
    if (best >= 0) {
        int2 q = qq[so + best];
        int64_t rb = (int64_t)rbeg[so + best];
        int slen = q.y - q.x;
        int64_t rmax0 = rb - (q.x + max_gap(R, q.x));
        int64_t rmax1 = rb + slen + ((len - q.y) + max_gap(R, len - q.y));
        if (rmax0 < 0) rmax0 = 0;
        if (rmax1 > (l_pac << 1)) rmax1 = l_pac << 1;
        if (rmax0 < l_pac && l_pac < rmax1) { if (rb < l_pac) rmax1 = l_pac; else rmax0 = l_pac; }
        h0 = (uint32_t)(slen * R.a);
        o.seed_rbeg = rb; o.seed_qbeg = q.x; o.seed_qend = q.y; o.h0 = (int32_t)h0;
        lq = (uint32_t)q.x; lt = q.x > 0 ? (uint32_t)(rb - rmax0) : 0u;
        rq = (uint32_t)(len - q.y); rt = q.y < len ? (uint32_t)(rmax1 - (rb + slen)) : 0u;
        const uint32_t rev = rb >= l_pac ? AUX_REV : 0u;
        la.start = rb; la.qfrom = q.x; la.read_flags |= rev; ra.start = rb + slen; ra.qfrom = q.y; ra.read_flags |= rev;
    }
    out[r] = o;
    const uint32_t jl = 2 * r, jr = 2 * r + 1;
    jq_len[jl] = lq; jt_len[jl] = lt; j_h0[jl] = h0; jq_off[jl] = jl * qstride * 8; jt_off[jl] = jl * tstride * 8; aux[jl] = la;
    jq_len[jr] = rq; jt_len[jr] = rt; j_h0[jr] = h0; jq_off[jr] = jr * qstride * 8; jt_off[jr] = jr * tstride * 8; aux[jr] = ra;
}

// one lane per output word (8 bases) of every job's query and target slot; a word is cut with one funnel shift of the
// packed read / 2-bit reference plus bit spreading (chain_core.cuh), not base by base
__global__ void __launch_bounds__(256)
cut_kernel(uint32_t n_jobs, int64_t l_pac, const uint32_t *__restrict__ pac, int64_t pac_words,
           const uint32_t *__restrict__ packed_reads, const uint64_t *__restrict__ word_off,
           const uint32_t *__restrict__ jq_len, const uint32_t *__restrict__ jt_len,
           const JobAux *__restrict__ aux, uint32_t qstride, uint32_t tstride,
           uint32_t *__restrict__ qp, uint32_t *__restrict__ tp)
{
    // one warp per job, lanes over its words: no division per word, the job's descriptors are warp-uniform loads, and the words a
    // short job does not have are never visited
    const uint32_t lane = threadIdx.x & 31u, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < n_jobs; j += warps) {
        const uint32_t ql = jq_len[j], tl = jt_len[j];
        if (ql == 0 && tl == 0) continue;
        const JobAux ax = aux[j];
        const uint64_t wo = word_off[j >> 1];
        const int64_t rw = (int64_t)(word_off[(j >> 1) + 1] - wo);
        const uint32_t qw = (ql + 7) >> 3, tw = (tl + 7) >> 3;
        for (uint32_t wi = lane; wi < qw; wi += 32) qp[(uint64_t)j * qstride + wi] = b200chain::cut_query_word(packed_reads + wo, rw, ax, wi, ql);
        for (uint32_t ti = lane; ti < tw; ti += 32) tp[(uint64_t)j * tstride + ti] = b200chain::cut_target_word(pac, pac_words, l_pac, ax, ti, tl);
    }
}

__global__ void gather_kernel(uint32_t n_reads, const bwa_b200_ext_result_t *__restrict__ res, const uint32_t *__restrict__ jq_len,
                              bwa_b200_read_result_t *__restrict__ out, unsigned long long *__restrict__ n_jobs_live)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t live = 0;
    if (r < n_reads) {
        out[r].left = res[2 * r];
        out[r].right = res[2 * r + 1];
        live = (jq_len[2 * r] > 0) + (jq_len[2 * r + 1] > 0);
    }
    for (int o = 16; o > 0; o >>= 1) live += __shfl_down_sync(0xffffffffu, live, o);
    if ((threadIdx.x & 31) == 0 && live) atomicAdd(n_jobs_live, (unsigned long long)live);
}

} // namespace

struct bwa_b200_pipeline {
    const bwa_b200_index *idx = nullptr;
    bwa_b200_seeder *seeder = nullptr;
    bwa_b200_extender *ext = nullptr;
    cudaStream_t stream = nullptr;
    int device = 0, n_sm = 0;
    uint64_t max_reads = 0;
    uint32_t max_read_len = 0, qstride = 0, tstride = 0;
    uint32_t *d_jq_len = nullptr, *d_jt_len = nullptr, *d_j_h0 = nullptr, *d_jq_off = nullptr, *d_jt_off = nullptr;
    JobAux *d_aux = nullptr;
    uint32_t *d_qp = nullptr, *d_tp = nullptr;
    uint64_t qp_words = 0, tp_words = 0;
    bwa_b200_ext_result_t *d_res = nullptr;
    bwa_b200_read_result_t *d_out = nullptr;
    unsigned long long *d_live = nullptr, *h_tot = nullptr;
    uint64_t launches = 0;
    b200::Prof prof;
    int profiling = 0;            // 0 off, 1 per kernel, 2 per phase (extension bins overlap as in an unprofiled step)
    // state of the batch in flight (for overflow repair)
    const uint32_t *b_packed = nullptr; const uint64_t *b_woff = nullptr; const uint32_t *b_len = nullptr;
    uint64_t b_n = 0; uint32_t b_maxlen = 0;
    bwa_b200_seed_params_t b_sp{19, 500}; bwa_b200_ext_params_t b_ep{}; bwa_b200_read_result_t *b_out = nullptr;
    // host-buffer path: copy streams so that the H2D of the next slice and the D2H of the previous one overlap the kernels
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_in[16] = {};
};

extern "C" int bwa_b200_index_attach_ref(bwa_b200_index_t *idx, const uint8_t *fwd, uint64_t l_pac)
{
    if (!idx || !fwd || 2 * l_pac != idx->v.seq_len) { b200::set_error("attach_ref: l_pac does not match the index (seq_len = 2*l_pac)"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(idx->device));
    uint64_t n_words = (l_pac + 15) / 16 + 1;
    std::vector<uint32_t> w(n_words, 0);
    for (uint64_t i = 0; i < l_pac; ++i) w[i >> 4] |= (uint32_t)(fwd[i] & 3) << ((~i & 15) << 1);
    if (idx->d_pac) cudaFree(idx->d_pac);
    idx->d_pac = nullptr;
    B200_CUDA(cudaMalloc(&idx->d_pac, n_words * 4));
    B200_CUDA(cudaMemcpy(idx->d_pac, w.data(), n_words * 4, cudaMemcpyHostToDevice));
    idx->l_pac = l_pac;
    return BWA_B200_OK;
}

static int pipe_alloc_jobs(bwa_b200_pipeline *p, uint32_t max_len, const bwa_b200_ext_params_t *ep)
{
    // slot sizes: query <= read, target <= query + min(cal_max_gap, 2w)
    uint32_t qstride = (max_len + 7) / 8;
    int a = ep->mat[0] > 0 ? ep->mat[0] : 1;
    double gd = (double)((int)max_len * a - ep->o_del) / ep->e_del + 1., gi = (double)((int)max_len * a - ep->o_ins) / ep->e_ins + 1.;
    int64_t gap = (int64_t)(gd > gi ? gd : gi);
    if (gap < 1) gap = 1;
    if (gap > 2 * (int64_t)ep->w) gap = 2 * (int64_t)ep->w;
    uint32_t tstride = (uint32_t)((max_len + gap + 7) / 8) + 1;
    uint64_t nj = 2 * p->max_reads;
    if (qstride > p->qstride || !p->d_qp) {
        cudaFree(p->d_qp); p->d_qp = nullptr;
        B200_CUDA(cudaMalloc(&p->d_qp, nj * qstride * 4));
        p->qstride = qstride;
    }
    if (tstride > p->tstride || !p->d_tp) {
        cudaFree(p->d_tp); p->d_tp = nullptr;
        B200_CUDA(cudaMalloc(&p->d_tp, nj * tstride * 4));
        p->tstride = tstride;
    }
    if ((uint64_t)nj * p->tstride * 8 >= 0xffffffffull) { b200::set_error("pipeline: batch too large for 32-bit job offsets; use smaller batches"); return BWA_B200_ERR_CAPACITY; }
    return BWA_B200_OK;
}

extern "C" int bwa_b200_pipeline_create(const bwa_b200_index_t *idx, uint64_t max_reads, uint64_t max_words,
                                        uint32_t max_read_len, bwa_b200_pipeline_t **out)
{
    if (!idx || !out || !max_reads) { b200::set_error("pipeline_create: bad argument"); return BWA_B200_ERR_ARG; }
    if (!idx->d_pac) { b200::set_error("pipeline_create: no reference attached (bwa_b200_index_attach_ref)"); return BWA_B200_ERR_ARG; }
    bwa_b200_pipeline *p = new bwa_b200_pipeline();
    p->idx = idx; p->device = idx->device; p->max_reads = max_reads; p->max_read_len = max_read_len;
    int rc = bwa_b200_seeder_create(idx, max_reads, max_words, &p->seeder);
    if (rc) return rc;
    rc = bwa_b200_extender_create(idx->device, 2 * max_reads, 1024, 1024, &p->ext);
    if (rc) return rc;
    p->stream = p->seeder->stream;
    cudaStreamDestroy(p->ext->stream);          // everything runs on one stream, in order
    p->ext->stream = p->stream; p->ext->own_stream = false;
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, idx->device));
    p->n_sm = prop.multiProcessorCount;
    uint64_t nj = 2 * max_reads;
    B200_CUDA(cudaMalloc(&p->d_jq_len, nj * 4)); B200_CUDA(cudaMalloc(&p->d_jt_len, nj * 4)); B200_CUDA(cudaMalloc(&p->d_j_h0, nj * 4));
    B200_CUDA(cudaMalloc(&p->d_jq_off, nj * 4)); B200_CUDA(cudaMalloc(&p->d_jt_off, nj * 4)); B200_CUDA(cudaMalloc(&p->d_aux, nj * sizeof(JobAux)));
    B200_CUDA(cudaMalloc(&p->d_res, nj * sizeof(bwa_b200_ext_result_t)));
    B200_CUDA(cudaMalloc(&p->d_out, max_reads * sizeof(bwa_b200_read_result_t)));
    B200_CUDA(cudaMalloc(&p->d_live, 8));
    B200_CUDA(cudaHostAlloc(&p->h_tot, 32, cudaHostAllocDefault));
    *out = p;
    return BWA_B200_OK;
}

extern "C" void bwa_b200_pipeline_destroy(bwa_b200_pipeline_t *p)
{
    if (!p) return;
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    p->seeder->prof = nullptr; p->ext->prof = nullptr; p->ext->phase_prof = nullptr;
    bwa_b200_extender_destroy(p->ext);
    bwa_b200_seeder_destroy(p->seeder);
    cudaFree(p->d_jq_len); cudaFree(p->d_jt_len); cudaFree(p->d_j_h0); cudaFree(p->d_jq_off); cudaFree(p->d_jt_off); cudaFree(p->d_aux);
    cudaFree(p->d_qp); cudaFree(p->d_tp); cudaFree(p->d_res); cudaFree(p->d_out); cudaFree(p->d_live);
    cudaFreeHost(p->h_tot);
    if (p->s_h2d) cudaStreamDestroy(p->s_h2d);
    if (p->s_d2h) cudaStreamDestroy(p->s_d2h);
    for (int k = 0; k < 16; ++k) if (p->ev_in[k]) cudaEventDestroy(p->ev_in[k]);
    delete p;
}

static int pipe_enqueue(bwa_b200_pipeline *p)
{
    const uint32_t n = (uint32_t)p->b_n;
    cudaStream_t st = p->stream;
    b200::Prof *prof = p->profiling ? &p->prof : nullptr;
    p->seeder->prof = prof; p->ext->prof = p->profiling == 1 ? prof : nullptr; p->ext->phase_prof = p->profiling == 2 ? prof : nullptr;
    if (prof) prof->reset();
    int rc = pipe_alloc_jobs(p, p->b_maxlen, &p->b_ep);
    if (rc) return rc;
    rc = b200_seeder_run(p->seeder, p->b_packed, p->b_woff, p->b_len, p->b_n, p->b_maxlen, &p->b_sp);
    if (rc) return rc;
    Rules R{p->b_ep.mat[0], p->b_ep.o_del, p->b_ep.e_del, p->b_ep.o_ins, p->b_ep.e_ins, p->b_ep.w};
    bwa_b200_seeder *s = p->seeder;
    B200_CUDA(cudaMemsetAsync(p->d_live, 0, 8, st));
    B200_LAUNCH(prof, "choose_kernel", st,
        (choose_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, (int64_t)p->idx->l_pac, R, p->b_len, s->d_nseeds, s->d_seed_off, s->d_rbeg,
                                                       s->d_qq, s->seed_cap, p->qstride, p->tstride, p->b_out, p->d_jq_len, p->d_jt_len,
                                                       p->d_j_h0, p->d_jq_off, p->d_jt_off, p->d_aux)));
    B200_LAUNCH(prof, "cut_kernel", st,
        (cut_kernel<<<p->n_sm * 16, 256, 0, st>>>(2 * n, (int64_t)p->idx->l_pac, p->idx->d_pac, (int64_t)((p->idx->l_pac + 15) / 16 + 1), p->b_packed, p->b_woff, p->d_jq_len,
                                                 p->d_jt_len, p->d_aux, p->qstride, p->tstride, p->d_qp, p->d_tp)));
    rc = b200_ext_run_packed(p->ext, &p->b_ep, 2 * n, p->d_qp, p->d_jq_off, p->d_jq_len, p->d_tp, p->d_jt_off, p->d_jt_len, p->d_j_h0, p->d_res, (int64_t)p->b_maxlen);
    if (rc) return rc;
    B200_LAUNCH(prof, "gather_kernel", st,
        (gather_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, p->d_res, p->d_jq_len, p->b_out, p->d_live)));
    p->launches += 3;
    B200_CUDA(cudaGetLastError());
    return BWA_B200_OK;
}

extern "C" int bwa_b200_seed_extend_device(bwa_b200_pipeline_t *p, const uint32_t *dev_packed, const uint64_t *dev_word_off,
                                           const uint32_t *dev_read_len, uint64_t n_reads, uint32_t max_read_len,
                                           const bwa_b200_seed_params_t *sp, const bwa_b200_ext_params_t *ep,
                                           bwa_b200_read_result_t *dev_out)
{
    if (!p || !sp || !ep || !dev_out || (n_reads && (!dev_packed || !dev_word_off || !dev_read_len))) { b200::set_error("seed_extend_device: bad argument"); return BWA_B200_ERR_ARG; }
    if (n_reads > p->max_reads) { b200::set_error("seed_extend: %llu reads > capacity", (unsigned long long)n_reads); return BWA_B200_ERR_CAPACITY; }
    if (sp->max_occ <= 0) { b200::set_error("seed_extend: max_occ must be positive"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(p->device));
    p->b_packed = dev_packed; p->b_woff = dev_word_off; p->b_len = dev_read_len; p->b_n = n_reads; p->b_maxlen = max_read_len;
    p->b_sp = *sp; p->b_ep = *ep; p->b_out = dev_out;
    if (n_reads == 0) return BWA_B200_OK;
    return pipe_enqueue(p);
}

// wait for the batch; if the seed arrays overflowed, they have been grown: run the batch again
extern "C" int bwa_b200_pipeline_sync(bwa_b200_pipeline_t *p)
{
    if (!p) return BWA_B200_ERR_ARG;
    B200_CUDA(cudaSetDevice(p->device));
    if (p->b_n == 0) return BWA_B200_OK;
    uint64_t cap_before = p->seeder->seed_cap;
    int rc = b200_seeder_finish(p->seeder);
    if (rc) return rc;
    if (p->seeder->last_total > cap_before || p->seeder->redone) {
        rc = pipe_enqueue(p);
        if (rc) return rc;
        rc = b200_seeder_finish(p->seeder);
        if (rc) return rc;
    }
    B200_CUDA(cudaStreamSynchronize(p->stream));
    return bwa_b200_extend_wait(p->ext);
}

extern "C" int bwa_b200_seed_extend_host(bwa_b200_pipeline_t *p, const uint32_t *packed, const uint64_t *word_off,
                                         const uint32_t *read_len, uint64_t n_reads, const bwa_b200_seed_params_t *sp,
                                         const bwa_b200_ext_params_t *ep, bwa_b200_read_result_t *host_out)
{
    if (!p || !sp || !ep || (n_reads && (!packed || !word_off || !read_len || !host_out))) { b200::set_error("seed_extend_host: bad argument"); return BWA_B200_ERR_ARG; }
    if (n_reads == 0) return BWA_B200_OK;
    bwa_b200_seeder *s = p->seeder;
    if (n_reads > p->max_reads || word_off[n_reads] > s->max_words) { b200::set_error("seed_extend_host: batch exceeds the pipeline capacity"); return BWA_B200_ERR_CAPACITY; }
    B200_CUDA(cudaSetDevice(p->device));
    uint32_t max_len = 0;
    for (uint64_t r = 0; r < n_reads; ++r) max_len = read_len[r] > max_len ? read_len[r] : max_len;
    // The batch is cut into slices of whole reads: every H2D copy is queued up front on a copy stream, the kernels of slice k
    // wait for its copy only, and its results go back on a third stream while slice k + 1 computes.  Offsets stay absolute
    // (one packed buffer, one output array), so a slice is just a sub-range of the batch.
    // Measured on B200 / C2 (tools/gpu_slices.sh): 1 slice 76.9, 2 slices 77.5, 4 slices 65.4, 8 slices 46.9 M reads/s -- at
    // PCIe 5 rates the copies are 2.9 ms of a 13 ms call and the fixed cost of a slice (about 30 launches, the host round
    // trip for the seed total, the tails of the extension bins) eats the overlap, so the default is one slice.
    int n_slices = getenv("BWA_B200_HOST_SLICES") ? atoi(getenv("BWA_B200_HOST_SLICES")) : 1;
    if (n_slices > 16) n_slices = 16;
    if (n_slices < 1 || n_reads < (uint64_t)n_slices * 32768) n_slices = 1;
    if (!p->s_h2d) {
        B200_CUDA(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
        B200_CUDA(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
        for (int k = 0; k < 16; ++k) B200_CUDA(cudaEventCreateWithFlags(&p->ev_in[k], cudaEventDisableTiming));
    }
    const uint64_t per = (n_reads + n_slices - 1) / n_slices;
    for (int k = 0; k < n_slices; ++k) {
        const uint64_t r0 = std::min<uint64_t>(n_reads, k * per), r1 = std::min<uint64_t>(n_reads, r0 + per);
        if (r1 > r0) {
            B200_CUDA(cudaMemcpyAsync(s->d_packed + word_off[r0], packed + word_off[r0], (word_off[r1] - word_off[r0]) * 4, cudaMemcpyHostToDevice, p->s_h2d));
            B200_CUDA(cudaMemcpyAsync(s->d_woff + r0, word_off + r0, (r1 - r0 + 1) * 8, cudaMemcpyHostToDevice, p->s_h2d));
            B200_CUDA(cudaMemcpyAsync(s->d_len + r0, read_len + r0, (r1 - r0) * 4, cudaMemcpyHostToDevice, p->s_h2d));
        }
        B200_CUDA(cudaEventRecord(p->ev_in[k], p->s_h2d));
    }
    for (int k = 0; k < n_slices; ++k) {
        const uint64_t r0 = std::min<uint64_t>(n_reads, k * per), r1 = std::min<uint64_t>(n_reads, r0 + per);
        if (r1 == r0) continue;
        B200_CUDA(cudaStreamWaitEvent(p->stream, p->ev_in[k], 0));
        int rc = bwa_b200_seed_extend_device(p, s->d_packed, s->d_woff + r0, s->d_len + r0, r1 - r0, max_len, sp, ep, p->d_out + r0);
        if (rc) return rc;
        rc = bwa_b200_pipeline_sync(p);          // slice k is complete (and its seed arrays were large enough, or it was re-run)
        if (rc) return rc;
        B200_CUDA(cudaMemcpyAsync(host_out + r0, p->d_out + r0, (r1 - r0) * sizeof(bwa_b200_read_result_t), cudaMemcpyDeviceToHost, p->s_d2h));
    }
    B200_CUDA(cudaStreamSynchronize(p->s_d2h));
    return BWA_B200_OK;
}

extern "C" void *bwa_b200_pipeline_stream(bwa_b200_pipeline_t *p) { return p ? (void *)p->stream : nullptr; }
extern "C" uint64_t bwa_b200_pipeline_launches(const bwa_b200_pipeline_t *p)
{
    return p ? p->launches + p->seeder->launches + p->ext->launches : 0;
}

extern "C" int bwa_b200_pipeline_totals(bwa_b200_pipeline_t *p, uint64_t out[3])
{
    if (!p || !out) return BWA_B200_ERR_ARG;
    B200_CUDA(cudaSetDevice(p->device));
    B200_CUDA(cudaMemcpyAsync(p->h_tot, p->d_live, 8, cudaMemcpyDeviceToHost, p->stream));
    B200_CUDA(cudaMemcpyAsync(p->h_tot + 1, p->ext->d_cells, 8, cudaMemcpyDeviceToHost, p->stream));
    B200_CUDA(cudaStreamSynchronize(p->stream));
    out[0] = p->seeder->last_total; out[1] = p->h_tot[0]; out[2] = p->h_tot[1];
    return BWA_B200_OK;
}

extern "C" int bwa_b200_pipeline_profile(bwa_b200_pipeline_t *p, int enable)
{
    if (!p) return BWA_B200_ERR_ARG;
    p->profiling = enable == 2 ? 2 : (enable != 0);
    return BWA_B200_OK;
}

extern "C" int bwa_b200_pipeline_kernel_times(bwa_b200_pipeline_t *p, const char **names, float *ms, int cap)
{
    if (!p) return BWA_B200_ERR_ARG;
    cudaSetDevice(p->device);
    cudaStreamSynchronize(p->stream);
    int n = 0;
    for (size_t i = 0; i < p->prof.used && n < cap; ++i, ++n) {
        float t = 0;
        cudaEventElapsedTime(&t, p->prof.recs[i].a, p->prof.recs[i].b);
        names[n] = p->prof.recs[i].name; ms[n] = t;
    }
    return n;
}
