// sw.cu -- ksw_align2 on the device: the local alignment of mate rescue (mem_matesw, src/bwamem_pair.c:159) and of mem_seed_sw
// (src/bwamem.c:803).  One job per lane; sw_core.cuh replays the reference's striped kernels so that score, end / start positions
// and the second-best hit are those of the SSE2 code, not of the textbook recurrence.  State (H0, H1, E, Hmax, row maxima) lives in a
// per-lane slice of a global workspace laid out [element][lane]: the 32 lanes of a warp touch one 64-byte row per access.
// Mate rescue is a minor share of a paired-end run (a few jobs per unpaired read); this kernel is about identity, not about the roofline.
#include "common.h"
#include "sw_core.cuh"
#include "sw_stripe.cuh"
#include <algorithm>
#include <vector>

struct bwa_b200_sw {
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    uint8_t *d_q = nullptr, *d_t = nullptr;
    uint64_t q_cap = 0, t_cap = 0, job_cap = 0;
    uint32_t *d_qoff = nullptr, *d_qlen = nullptr, *d_toff = nullptr, *d_tlen = nullptr, *d_xtra = nullptr;
    bwa_b200_sw_result_t *d_res = nullptr;
    int16_t *d_ws = nullptr;
    uint32_t *d_jobs = nullptr;                      // job indexes: the byte-kernel jobs of sw_stripe_kernel (longest target first), then the others
    uint64_t jobs_cap = 0;
    int smem_optin = 0;
    uint64_t ws_elems = 0;
    uint64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;        // around the kernel of the last batch
    float last_kernel_ms = 0.f;
};

namespace {

constexpr int SW_THREADS = 64;

__global__ void __launch_bounds__(SW_THREADS)
sw_align2_kernel(SwParams S, uint32_t n, const uint8_t *__restrict__ qseq, const uint32_t *__restrict__ qoff, const uint32_t *__restrict__ qlen,
                 const uint8_t *__restrict__ tseq, const uint32_t *__restrict__ toff, const uint32_t *__restrict__ tlen,
                 const uint32_t *__restrict__ xtra, int16_t *__restrict__ ws, uint32_t n_cap, uint32_t t_cap, bwa_b200_sw_result_t *__restrict__ res,
                 const uint32_t *__restrict__ jobs)
{
    const size_t NS = (size_t)gridDim.x * blockDim.x;
    const size_t lane = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    int16_t *my = ws + lane;                                   // H0 | H1 | E | Hmax, n_cap elements each, then t_cap row maxima
    int16_t *rowmax = ws + 4 * (size_t)n_cap * NS + lane;
    for (size_t k = lane; k < n; k += NS) {
        const size_t a = jobs[k];
        bwa_b200_sw_result_t r;
        sw_align2((int)qlen[a], qseq + qoff[a], (int)tlen[a], tseq + toff[a], S, (int)xtra[a], my, NS, n_cap, rowmax, r);
        res[a] = r;
    }
}

template <class T> int grow(T *&p, uint64_t &cap, uint64_t need)
{
    if (need <= cap) return BWA_B200_OK;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    B200_CUDA(cudaMalloc(&p, need * sizeof(T)));
    cap = need;
    return BWA_B200_OK;
}

} // namespace

extern "C" int bwa_b200_sw_create(int device, bwa_b200_sw_t **out)
{
    if (!out) { b200::set_error("sw_create: bad argument"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(device));
    bwa_b200_sw *s = new bwa_b200_sw();
    s->device = device;
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    s->n_sm = prop.multiProcessorCount;
    s->smem_optin = (int)prop.sharedMemPerBlockOptin;
    B200_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    B200_CUDA(cudaEventCreate(&s->ev0)); B200_CUDA(cudaEventCreate(&s->ev1));
    *out = s;
    return BWA_B200_OK;
}

extern "C" void bwa_b200_sw_destroy(bwa_b200_sw_t *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    cudaFree(s->d_q); cudaFree(s->d_t); cudaFree(s->d_qoff); cudaFree(s->d_qlen); cudaFree(s->d_toff); cudaFree(s->d_tlen); cudaFree(s->d_xtra);
    cudaFree(s->d_res); cudaFree(s->d_ws); cudaFree(s->d_jobs);
    cudaEventDestroy(s->ev0); cudaEventDestroy(s->ev1);
    cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" uint64_t bwa_b200_sw_launches(const bwa_b200_sw_t *s) { return s ? s->launches : 0; }
extern "C" float bwa_b200_sw_last_kernel_ms(const bwa_b200_sw_t *s) { return s ? s->last_kernel_ms : 0.f; }

extern "C" int bwa_b200_sw_align2_host(bwa_b200_sw_t *s, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                                       const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                                       const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen,
                                       const uint32_t *xtra, bwa_b200_sw_result_t *out)
{
    if (!s || !p || (n_jobs && (!qseq || !qoff || !qlen || !tseq || !toff || !tlen || !xtra || !out))) { b200::set_error("sw_align2_host: bad argument"); return BWA_B200_ERR_ARG; }
    if (n_jobs == 0) return BWA_B200_OK;
    if (n_jobs > 0x7fffffffull) { b200::set_error("sw_align2_host: at most 2^31 jobs per call"); return BWA_B200_ERR_ARG; }
    // Two classes.  Byte-kernel jobs (KSW_XBYTE: mem_matesw sets it when the read's best score fits a byte, src/bwamem_pair.c:150)
    // with at most 256 query bases and a target that fits the row buffer run in sw_stripe_kernel, longest target first so that the
    // four jobs of a warp finish together; everything else in the replay kernel.
    static const bool no_stripe = getenv("BWA_B200_SW_NO_STRIPE") != nullptr;
    constexpr uint32_t STRIPE_MAX_Q = 256, STRIPE_MAX_T = 6000;         // 32 jobs x 6000 row bytes = 192 KB of shared memory at most
    uint32_t max_q = 0, max_t = 0, fast_q = 0, fast_t = 0;
    std::vector<uint32_t> order(n_jobs);
    uint64_t n_fast = 0;
    {
        std::vector<uint32_t> slow;
        std::vector<uint32_t> cnt(STRIPE_MAX_T / 16 + 2, 0);
        std::vector<uint8_t> fast(n_jobs);
        for (uint64_t a = 0; a < n_jobs; ++a) {
            if ((uint64_t)qoff[a] + qlen[a] > q_bytes || (uint64_t)toff[a] + tlen[a] > t_bytes) { b200::set_error("sw_align2_host: job %llu lies outside its buffer", (unsigned long long)a); return BWA_B200_ERR_ARG; }
            if (qlen[a] == 0 || qlen[a] > 32000u || tlen[a] > 0x3fffffffu) { b200::set_error("sw_align2_host: job %llu: query of %u bases (1 .. 32000 supported)", (unsigned long long)a, qlen[a]); return BWA_B200_ERR_ARG; }
            fast[a] = !no_stripe && (xtra[a] & 0x10000u) && qlen[a] <= STRIPE_MAX_Q && tlen[a] <= STRIPE_MAX_T;
            if (fast[a]) { ++cnt[tlen[a] / 16]; ++n_fast; fast_q = std::max(fast_q, qlen[a]); fast_t = std::max(fast_t, tlen[a]); }
            else { slow.push_back((uint32_t)a); max_q = std::max(max_q, qlen[a]); max_t = std::max(max_t, tlen[a]); }
        }
        uint32_t at = 0;                                     // counting sort by target length / 16, descending
        for (size_t b = cnt.size(); b-- > 0;) { const uint32_t c = cnt[b]; cnt[b] = at; at += c; }
        for (uint64_t a = 0; a < n_jobs; ++a) if (fast[a]) order[cnt[tlen[a] / 16]++] = (uint32_t)a;
        std::copy(slow.begin(), slow.end(), order.begin() + n_fast);
    }
    const uint64_t n_slow = n_jobs - n_fast;
    B200_CUDA(cudaSetDevice(s->device));
    int rc;
    if ((rc = grow(s->d_q, s->q_cap, q_bytes + 16)) || (rc = grow(s->d_t, s->t_cap, t_bytes + 16))) return rc;
    if (n_jobs > s->job_cap) {
        uint64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0;
        const uint64_t need = n_jobs + n_jobs / 4 + 64;
        if ((rc = grow(s->d_qoff, c0, need)) || (rc = grow(s->d_qlen, c1, need)) || (rc = grow(s->d_toff, c2, need)) || (rc = grow(s->d_tlen, c3, need)) ||
            (rc = grow(s->d_xtra, c4, need)) || (rc = grow(s->d_res, c5, need))) { s->job_cap = 0; return rc; }
        s->job_cap = need;
    }
    // replay kernel: as many lanes as it has jobs, up to 8 blocks per SM, fewer when the per-lane state of this batch would not fit 2 GB
    const uint32_t n_cap = (max_q + 15) / 16 * 16 + 16, t_cap = max_t + 1;
    const uint64_t per_lane = 4ull * n_cap + t_cap;
    uint32_t grid = 0;
    if (n_slow) {
        uint64_t lanes = (uint64_t)s->n_sm * 8 * SW_THREADS;
        const uint64_t budget = (2ull << 30) / 2;                      // int16 elements
        if (lanes * per_lane > budget) lanes = budget / per_lane;
        if (lanes > n_slow) lanes = n_slow;
        grid = (uint32_t)((lanes + SW_THREADS - 1) / SW_THREADS);
        if (grid < 1) grid = 1;
        if ((rc = grow(s->d_ws, s->ws_elems, (uint64_t)grid * SW_THREADS * per_lane))) return rc;
    }
    if ((rc = grow(s->d_jobs, s->jobs_cap, n_jobs + n_jobs / 4 + 64))) return rc;
    cudaStream_t st = s->stream;
    B200_CUDA(cudaMemcpyAsync(s->d_jobs, order.data(), n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(s->d_q, qseq, q_bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(s->d_t, tseq, t_bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(s->d_qoff, qoff, n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(s->d_qlen, qlen, n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(s->d_toff, toff, n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(s->d_tlen, tlen, n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(s->d_xtra, xtra, n_jobs * 4, cudaMemcpyHostToDevice, st));
    SwParams S;
    memset(&S, 0, sizeof(S));
    memcpy(S.mat, p->mat, 25);
    S.m = 5; S.o_del = p->o_del; S.e_del = p->e_del; S.o_ins = p->o_ins; S.e_ins = p->e_ins;
    B200_CUDA(cudaEventRecord(s->ev0, st));
    if (n_fast) {
        b200sw::StripeParams SP;
        memset(&SP, 0, sizeof(SP));
        int mn = 127, mx = 0;
        for (int a = 0; a < 25; ++a) { mn = std::min(mn, (int)p->mat[a]); mx = std::max(mx, (int)p->mat[a]); }
        SP.shift = (256 - mn) & 255; SP.qmax = mx;
        for (int t = 0; t < 5; ++t) {
            for (int q = 0; q < 4; ++q) SP.tab[t] |= (uint32_t)(uint8_t)p->mat[t * 5 + q] << (8 * q);
            SP.tabn[t] = (uint32_t)(uint8_t)p->mat[t * 5 + 4] | 0x00800000u;     // byte 1 = 0 (padding), byte 2 = 0x80 (unused registers)
        }
        const int oe_del = p->o_del + p->e_del, oe_ins = p->o_ins + p->e_ins;
        SP.noe_del2 = (uint32_t)(uint16_t)(int16_t)(-oe_del) * 0x00010001u; SP.ne_del2 = (uint32_t)(uint16_t)(int16_t)(-p->e_del) * 0x00010001u;
        SP.noe_ins2 = (uint32_t)(uint16_t)(int16_t)(-oe_ins) * 0x00010001u; SP.ne_ins2 = (uint32_t)(uint16_t)(int16_t)(-p->e_ins) * 0x00010001u;
        const uint32_t tc = (fast_t + 15) / 16 * 16;
        const size_t smem = (size_t)b200sw::JOBS_PER_BLOCK * tc;
        const int slen_max = (int)(fast_q + 15) / 16;
        const bool sg = oe_del == oe_ins;
        // instantiated for the vector counts of the usual read lengths (7: up to 112 bases, 10: up to 160, 16: up to 256) besides 4, 8
        // and 12: a batch of one read length then runs the variant of the row loop that has no bound check (pass_u8<.., UNI>)
#define SW_PICK(N) (sg ? b200sw::sw_stripe_kernel<N, true> : b200sw::sw_stripe_kernel<N, false>)
        auto kern = slen_max <= 4 ? SW_PICK(4) : slen_max <= 7 ? SW_PICK(7) : slen_max <= 8 ? SW_PICK(8) : slen_max <= 10 ? SW_PICK(10)
                  : slen_max <= 12 ? SW_PICK(12) : SW_PICK(16);
#undef SW_PICK
        if (smem > 48 * 1024) B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, b200sw::BLOCK, smem));
        if (occ < 1) occ = 1;
        uint32_t g = (uint32_t)std::min<uint64_t>((n_fast + b200sw::JOBS_PER_BLOCK - 1) / b200sw::JOBS_PER_BLOCK, (uint64_t)s->n_sm * occ);
        kern<<<g, b200sw::BLOCK, smem, st>>>(SP, (uint32_t)n_fast, s->d_jobs, s->d_q, s->d_qoff, s->d_qlen, s->d_t, s->d_toff, s->d_tlen, s->d_xtra, tc, s->d_res);
        B200_CUDA(cudaGetLastError());
        s->launches += 1;
    }
    if (n_slow) {
        sw_align2_kernel<<<grid, SW_THREADS, 0, st>>>(S, (uint32_t)n_slow, s->d_q, s->d_qoff, s->d_qlen, s->d_t, s->d_toff, s->d_tlen, s->d_xtra,
                                                      s->d_ws, n_cap, t_cap, s->d_res, s->d_jobs + n_fast);
        B200_CUDA(cudaGetLastError());
        s->launches += 1;
    }
    B200_CUDA(cudaEventRecord(s->ev1, st));
    B200_CUDA(cudaMemcpyAsync(out, s->d_res, n_jobs * sizeof(bwa_b200_sw_result_t), cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    cudaEventElapsedTime(&s->last_kernel_ms, s->ev0, s->ev1);
    return BWA_B200_OK;
}
