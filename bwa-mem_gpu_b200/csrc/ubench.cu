// ubench.cu -- measured denominators for the extension roofline: issue rate of the integer instructions the extension
// kernels are made of, in warp-instructions per clock per SM, at full occupancy with 8 independent chains per lane.
// bench.py calls bwa_b200_measure_int_alu live and derives the INT-ALU peak from it instead of assuming a lane count.
#include "common.h"

namespace {

constexpr int CHAINS = 8;

template <int OP>
__device__ __forceinline__ uint32_t ub_op(uint32_t a, uint32_t b, uint32_t c)
{
    if (OP == 0) return a + b + c;                               // IADD3
    if (OP == 1) return (a & b) ^ c;                             // LOP3
    if (OP == 2) return __byte_perm(a, b, c);                    // PRMT
    if (OP == 3) return (uint32_t)max((int)a, (int)c);           // VIMNMX.S32
    if (OP == 4) return __viaddmax_s32(a, b, c);                 // VIADDMNMX.S32
    if (OP == 5) return __vimax3_s32(a, b, c);                   // VIMNMX3.S32
    if (OP == 6) return __viaddmax_s16x2(a, b, c);               // VIADDMNMX.S16x2
    if (OP == 7) return __vimax3_s16x2(a, b, c);                 // VIMNMX3.S16x2
    if (OP == 8) return a * b + c;                               // IMAD (fma pipe)
    if (OP == 9) return __umulhi(a, b) + c;                      // IMAD.HI (fma pipe)
    return a;
}

template <int OP>
__global__ void __launch_bounds__(256) ub_kernel(uint32_t *out, int iters, uint32_t s0, uint32_t s1)
{
    uint32_t v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = threadIdx.x * 7 + i + s0;
    const uint32_t c = s0 ^ 0x00030003;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) v[i] = ub_op<OP>(v[i], v[(i + 3) & 7], c);
    }
    uint32_t r = s1;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= v[i];
    if (r == 0x12345678u) out[0] = r;
}

// half of the chains on the ALU pipe (VIADDMNMX.S16x2), half on the FMA pipe (IMAD): do the two pipes issue side by side?
template <>
__global__ void __launch_bounds__(256) ub_kernel<10>(uint32_t *out, int iters, uint32_t s0, uint32_t s1)
{
    uint32_t v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = threadIdx.x * 7 + i + s0;
    const uint32_t c = s0 ^ 0x00030003;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) v[i] = (i & 1) ? v[i] * v[(i + 2) & 7] + c : __viaddmax_s16x2(v[i], v[(i + 2) & 7], c);
    }
    uint32_t r = s1;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= v[i];
    if (r == 0x12345678u) out[0] = r;
}

// SM clocks elapsed in a fixed spin, against the event time of the same launch: the clock the rates are divided by
__global__ void ub_clock_kernel(unsigned long long *out, long long spin)
{
    const long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(clock64() - t0);
}

template <int OP> double ub_run(uint32_t *out, int sms, double mhz)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 2048, blocks = sms * 8;
    ub_kernel<OP><<<blocks, 256>>>(out, 16, 1, 2);
    double best = 0;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a);
        ub_kernel<OP><<<blocks, 256>>>(out, iters, 1, 2);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double inst = (double)blocks * 8 /*warps*/ * iters * 4.0 * CHAINS;
        const double rate = inst / (ms * 1e-3) / (mhz * 1e6) / sms;
        if (rate > best) best = rate;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return best;
}

} // namespace

extern "C" const char *bwa_b200_int_alu_op_name(int i)
{
    static const char *names[] = {"IADD3", "LOP3", "PRMT", "VIMNMX.S32", "VIADDMNMX.S32", "VIMNMX3.S32", "VIADDMNMX.S16x2", "VIMNMX3.S16x2", "IMAD", "IMAD.HI", "VIADDMNMX.S16x2+IMAD"};
    return i >= 0 && i < 11 ? names[i] : nullptr;
}

// rates[i] = warp-instructions per clock per SM of op i (names above), at the SM clock measured under load (*sm_mhz);
// returns the number of ops written, or a negative error code
extern "C" int bwa_b200_measure_int_alu(int device, double *rates, int cap, double *sm_mhz, int *n_sm)
{
    if (!rates || cap < 9) { b200::set_error("measure_int_alu: need room for 9 rates"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    const int sms = prop.multiProcessorCount;
    uint32_t *out = nullptr;
    unsigned long long *clk = nullptr, h_clk = 0;
    B200_CUDA(cudaMalloc(&out, 4));
    B200_CUDA(cudaMalloc(&clk, 8));
    // warm the clocks up with real work, then measure the SM clock
    ub_kernel<0><<<sms * 8, 256>>>(out, 4096, 1, 2);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    ub_clock_kernel<<<1, 32>>>(clk, 20000000ll);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    B200_CUDA(cudaMemcpy(&h_clk, clk, 8, cudaMemcpyDeviceToHost));
    const double mhz = (double)h_clk / (ms * 1e-3) / 1e6;
    cudaEventDestroy(a); cudaEventDestroy(b);
    rates[0] = ub_run<0>(out, sms, mhz); rates[1] = ub_run<1>(out, sms, mhz); rates[2] = ub_run<2>(out, sms, mhz);
    rates[3] = ub_run<3>(out, sms, mhz); rates[4] = ub_run<4>(out, sms, mhz); rates[5] = ub_run<5>(out, sms, mhz);
    rates[6] = ub_run<6>(out, sms, mhz); rates[7] = ub_run<7>(out, sms, mhz); rates[8] = ub_run<8>(out, sms, mhz);
    int n_ops = 9;
    if (cap >= 11) { rates[9] = ub_run<9>(out, sms, mhz); rates[10] = ub_run<10>(out, sms, mhz); n_ops = 11; }
    cudaFree(out); cudaFree(clk);
    B200_CUDA(cudaGetLastError());
    if (sm_mhz) *sm_mhz = mhz;
    if (n_sm) *n_sm = sms;
    return n_ops;
}
