// ubench_int.cu -- issue-rate microbenchmark of the integer instructions the extension kernel uses
// (measured INT-ALU roofline denominator).  Prints warp-instructions per clock per SM
// for each op with 8 independent chains per lane at full occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_int.bin tools/ubench_int.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define CHAINS 8
template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{
    if (OP == 0) return __viaddmax_s16x2(a, b, c);
    if (OP == 1) return __vimax3_s16x2(a, b, c);
    if (OP == 2) return __viaddmax_s32(a, b, c);
    if (OP == 3) return __vimax3_s32(a, b, c);
    if (OP == 4) return a + b + c;                               // IADD3
    if (OP == 5) return (a & b) ^ c;                             // LOP3
    if (OP == 6) return __byte_perm(a, b, c);                    // PRMT
    if (OP == 7) return a * b + c;                               // IMAD
    if (OP == 8) return __funnelshift_l(a, b, c);                // SHF
    if (OP == 9) { bool ph, pl; uint32_t m = __vibmax_s16x2(a, c, &ph, &pl); return m + (pl ? b : 0u); } // VIMNMX.P + SEL-ish
    if (OP == 10) return __viaddmax_s16x2_relu(a, b, c);
    if (OP == 11) return max((int)a, (int)c);                    // VIMNMX s32
    if (OP == 12) return __popc(a) + c;                          // POPC
    return a;
}
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t *out, int iters, uint32_t s0, uint32_t s1)
{
    uint32_t v[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) v[i] = threadIdx.x * 7 + i + s0;
    uint32_t b = s1 | 1, c = s0 ^ 0x00030003;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) v[i] = op<OP>(v[i], v[(i + 3) & 7], c);
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= v[i];
    if (r == 0x12345678u) out[0] = r;
}
// mixed: the DP cell's DPX ops next to an IMAD (fma pipe) and LDS/STS, to see co-issue
__global__ void __launch_bounds__(256) kmix(uint32_t *out, int iters, uint32_t s0, uint32_t s1)
{
    __shared__ uint2 sm[16][256];
    uint32_t f = s0, h1 = s1, m = 0, acc = 0;
    for (int j = 0; j < 16; ++j) sm[j][threadIdx.x] = make_uint2(s0 + j, s1 ^ j);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint2 p = sm[j][threadIdx.x];
            uint32_t hk = p.x * 32u;
            uint32_t sc = __byte_perm(s0, s1, p.y ^ acc);
            uint32_t M = __viaddmin_s16x2(p.x, sc, hk);
            uint32_t h = __vimax3_s16x2(M, p.y, f);
            m = __vmaxs2(m, h);
            uint32_t t1 = __viaddmax_s16x2(M, s1, 0);
            uint32_t e = __viaddmax_s16x2(p.y, s0, t1);
            uint32_t t2 = __viaddmax_s16x2(M, s0, 0);
            f = __viaddmax_s16x2(f, s1, t2);
            sm[j][threadIdx.x] = make_uint2(h1, e);
            h1 = h;
        }
        acc += m;
    }
    if (acc == 0x12345678u) out[0] = acc + f;
}

template <int OP> void run(const char *name, int sms, double mhz)
{
    uint32_t *out; cudaMalloc(&out, 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 2048, blocks = sms * 8;
    k<OP><<<blocks, 256>>>(out, 16, 1, 2);
    cudaEventRecord(a);
    k<OP><<<blocks, 256>>>(out, iters, 1, 2);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double inst = (double)blocks * 8 /*warps*/ * iters * 4.0 * CHAINS;
    printf("%-22s %8.3f ms  %6.2f warp-inst/clk/SM (at %.0f MHz)  %8.1f Gop/s lanes\n", name, ms, inst / (ms * 1e-3) / (mhz * 1e6) / sms, mhz,
           inst * 32 / (ms * 1e-3) / 1e9);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double mhz = clk / 1e3; int sms = p.multiProcessorCount;
    printf("%s, %d SMs, %.0f MHz (attr)\n", p.name, sms, mhz);
    run<0>("VIADDMNMX.S16x2", sms, mhz); run<10>("VIADDMNMX.S16x2.RELU", sms, mhz); run<1>("VIMNMX3.S16x2", sms, mhz);
    run<2>("VIADDMNMX.S32", sms, mhz); run<3>("VIMNMX3.S32", sms, mhz); run<11>("VIMNMX.S32", sms, mhz);
    run<4>("IADD3", sms, mhz); run<5>("LOP3", sms, mhz); run<6>("PRMT", sms, mhz); run<7>("IMAD", sms, mhz); run<8>("SHF", sms, mhz);
    run<9>("VIMNMX.S16x2+P,SEL,IADD", sms, mhz); run<12>("POPC+IADD", sms, mhz);
    {
        uint32_t *out; cudaMalloc(&out, 4);
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        const int iters = 4096, blocks = sms * 8;
        kmix<<<blocks, 256>>>(out, 16, 1, 2);
        cudaEventRecord(a);
        kmix<<<blocks, 256>>>(out, iters, 0x00010001, 0xfffefffe);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        double cells2 = (double)blocks * 256 * iters * 16;
        printf("mixed s16x2 cell (8 ALU + IMAD + LDS.64 + STS.64): %.3f ms, %.1f G cell-pairs/s = %.1f GCUPS, %.2f clk/warp-cellpair/SMSP\n", ms,
               cells2 / (ms * 1e-3) / 1e9, 2 * cells2 / (ms * 1e-3) / 1e9, (ms * 1e-3) * mhz * 1e6 * sms * 4 / (cells2 / 32));
    }
    return 0;
}
