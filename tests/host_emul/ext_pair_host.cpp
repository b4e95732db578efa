// Host build of the column-pair s16x2 extension (bwa-mem_gpu_b200/csrc/ext_pair_core.cuh) with the
// integer intrinsics it uses emulated in plain C++.  TEST INFRASTRUCTURE: lets the exact kernel
// source be fuzzed against the oracle on the CPU box (tests/test_ext_pair_host.py); never shipped
// and never used as a compute path.
//   g++ -O2 -shared -fPIC -I include -I bwa-mem_gpu_b200/csrc tests/host_emul/ext_pair_host.cpp -o tests/host_emul/libextpair_host.so
#include <stdint.h>
#include <string.h>
#include <vector>

// prmt.b32, generic mode (PTX ISA): selector bit 3 replicates the sign of the selected byte; only c[15:0] is used
static inline uint32_t b200_prmt(uint32_t x, uint32_t y, uint32_t s)
{
    const uint64_t v = (uint64_t)y << 32 | x;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        const uint32_t sel = (s >> (4 * i)) & 0xfu;
        uint32_t b = (uint32_t)(v >> (8 * (sel & 7u))) & 0xffu;
        if (sel & 8u) b = (b & 0x80u) ? 0xffu : 0u;
        r |= b << (8 * i);
    }
    return r;
}
// __byte_perm (CUDA math API): only bits 2:0 of each selector nibble are used
static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) { return b200_prmt(x, y, s & 0x7777u); }
static inline int16_t lo16(uint32_t v) { return (int16_t)(v & 0xffffu); }
static inline int16_t hi16(uint32_t v) { return (int16_t)(v >> 16); }
static inline uint32_t pk(int lo, int hi) { return (uint32_t)(uint16_t)(int16_t)lo | (uint32_t)(uint16_t)(int16_t)hi << 16; }
static inline int mn(int a, int b) { return a < b ? a : b; }
static inline int mx(int a, int b) { return a > b ? a : b; }
static inline uint32_t __viaddmin_s16x2(uint32_t a, uint32_t b, uint32_t c)
{ return pk(mn((int16_t)(lo16(a) + lo16(b)), lo16(c)), mn((int16_t)(hi16(a) + hi16(b)), hi16(c))); }
static inline uint32_t __viaddmax_s16x2(uint32_t a, uint32_t b, uint32_t c)
{ return pk(mx((int16_t)(lo16(a) + lo16(b)), lo16(c)), mx((int16_t)(hi16(a) + hi16(b)), hi16(c))); }
static inline uint32_t __viaddmax_s16x2_relu(uint32_t a, uint32_t b, uint32_t c)
{ return pk(mx(mx((int16_t)(lo16(a) + lo16(b)), lo16(c)), 0), mx(mx((int16_t)(hi16(a) + hi16(b)), hi16(c)), 0)); }
static inline uint32_t __vimax3_s16x2(uint32_t a, uint32_t b, uint32_t c)
{ return pk(mx(mx(lo16(a), lo16(b)), lo16(c)), mx(mx(hi16(a), hi16(b)), hi16(c))); }
static inline uint32_t __vibmax_s16x2(uint32_t a, uint32_t b, bool *ph, bool *pl)
{ *pl = lo16(a) >= lo16(b); *ph = hi16(a) >= hi16(b); return pk(mx(lo16(a), lo16(b)), mx(hi16(a), hi16(b))); }
static inline uint32_t umx(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b)
{ return umx(a & 0xffffu, b & 0xffffu) | umx(a >> 16, b >> 16) << 16; }
static inline uint32_t __viaddmax_u16x2(uint32_t a, uint32_t b, uint32_t c)
{ return umx((a + b) & 0xffffu, c & 0xffffu) | umx(((a >> 16) + (b >> 16)) & 0xffffu, c >> 16) << 16; }
static inline uint32_t __vimax3_u16x2(uint32_t a, uint32_t b, uint32_t c) { return __vmaxu2(__vmaxu2(a, b), c); }
static inline uint32_t b200_mad(uint32_t a, uint32_t b, uint32_t c) { return a * b + c; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t __vimax3_u32(uint32_t a, uint32_t b, uint32_t c) { return max(max(a, b), c); }
static inline uint32_t __vmins2(uint32_t a, uint32_t b) { return pk(mn(lo16(a), lo16(b)), mn(hi16(a), hi16(b))); }

#include "ext_pair_core.cuh"

// the WIDE instantiation (scores up to 32700, queries up to 65535, ring state): every job whose score bound fits runs, whatever its
// length; skipped[a] = 1 otherwise.  Returns the evaluated cells, -1 for ineligible parameters, -2 when the band allows no ring.
extern "C" long long ext_pair_host_run_wide(const bwa_b200_ext_params_t *p, uint64_t n, const uint8_t *qseq, const uint32_t *qoff,
                                            const uint32_t *qlen, const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen,
                                            const uint32_t *h0, int32_t *res6, uint8_t *skipped)
{
    PairParams S;
    if (!pair_params_from(p, &S)) return -1;
    if (!p->use_band) return -2;
    ExtParams P;
    memset(&P, 0, sizeof(P));
    memcpy(P.mat, p->mat, 25);
    P.o_del = p->o_del; P.e_del = p->e_del; P.o_ins = p->o_ins; P.e_ins = p->e_ins;
    P.w = p->w; P.end_bonus = p->end_bonus; P.zdrop = p->zdrop; P.use_band = p->use_band; P.pen_clip = p->pen_clip;
    int mxs = 0;
    for (int i = 0; i < 25; ++i) mxs = mxs > p->mat[i] ? mxs : p->mat[i];
    P.max_score = mxs;
    JobView J{qseq, tseq, nullptr, nullptr, qoff, qlen, toff, tlen, h0};
    unsigned long long cells = 0;
    S.ring = p->w + 2; S.ring_magic = pair_ring_magic(S.ring);
    std::vector<uint2> HE(S.ring, uint2{0xdeadbeefu, 0xdeadbeefu});
    std::vector<uint16_t> QS(S.ring, 0xdeadu);
    const bool same = pair_same_gap(p);
    for (uint64_t a = 0; a < n; ++a) {
        const int ql = (int)qlen[a], tl = (int)tlen[a], h = (int)h0[a];
        const uint64_t bound = (uint64_t)h + (uint64_t)ql * (uint64_t)mxs;
        skipped[a] = (bound > (uint64_t)PAIR_WIDE_MAX_SCORE || ql > PAIR_WIDE_MAX_Q || ql < 1 || h < 1) ? 1 : 0;
        if (skipped[a]) continue;
        bwa_b200_ext_result_t r;
        if (same) pair_job<true, 1, true, true, false, 8, true>(P, S, S.tab, J, (uint32_t)a, ql, tl, h, HE.data(), QS.data(), r, cells);
        else pair_job<true, 1, false, true, false, 8, true>(P, S, S.tab, J, (uint32_t)a, ql, tl, h, HE.data(), QS.data(), r, cells);
        memcpy(res6 + a * 6, &r, 24);
    }
    return (long long)cells;
}

// jobs in the GASAL byte layout; res6 = n x 6 int32; returns the number of evaluated cells, or -1 if the parameters
// are not eligible for the pair kernel; skipped[a] = 1 for jobs outside its class (score bound > 1023, query > 512).
// variant: bit 0 = four pairs per trip of the longest unrolled loop instead of eight, bit 1 = band-sized ring state (when the band allows one)
extern "C" long long ext_pair_host_run(const bwa_b200_ext_params_t *p, int variant, uint64_t n, const uint8_t *qseq, const uint32_t *qoff,
                                       const uint32_t *qlen, const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen,
                                       const uint32_t *h0, int32_t *res6, uint8_t *skipped)
{
    PairParams S;
    if (!pair_params_from(p, &S)) return -1;
    ExtParams P;
    memset(&P, 0, sizeof(P));
    memcpy(P.mat, p->mat, 25);
    P.o_del = p->o_del; P.e_del = p->e_del; P.o_ins = p->o_ins; P.e_ins = p->e_ins;
    P.w = p->w; P.end_bonus = p->end_bonus; P.zdrop = p->zdrop; P.use_band = p->use_band; P.pen_clip = p->pen_clip;
    int mxs = 0;
    for (int i = 0; i < 25; ++i) mxs = mxs > p->mat[i] ? mxs : p->mat[i];
    P.max_score = mxs;
    JobView J{qseq, tseq, nullptr, nullptr, qoff, qlen, toff, tlen, h0};
    unsigned long long cells = 0;
    std::vector<uint2> HE;
    std::vector<uint16_t> QS;
    const bool same = pair_same_gap(p);
    for (uint64_t a = 0; a < n; ++a) {
        const int ql = (int)qlen[a], tl = (int)tlen[a], h = (int)h0[a];
        const uint64_t bound = (uint64_t)h + (uint64_t)ql * (uint64_t)mxs;
        skipped[a] = (bound > (uint64_t)PAIR_MAX_SCORE || ql > PAIR_MAX_Q || ql < 1 || h < 1) ? 1 : 0;
        if (skipped[a]) continue;
        // the kernel sizes the state by the longest query of the job's length bin; here: the next multiple of 16
        const int bin_q = (ql + 15) / 16 * 16;
        int slots = bin_q / 2 + 1;
        S.ring = 0; S.ring_magic = 0;
        if (variant & 2) slots = pair_slots(p, bin_q, &S);
        HE.assign(slots, uint2{0xdeadbeefu, 0xdeadbeefu}); QS.assign(slots, 0xdeadu);
        bwa_b200_ext_result_t r;
        const bool ring = S.ring != 0, chunked = slots > PAIR_CHUNK + 1;
#define RUN(SG, RG, CH, UU) pair_job<true, 1, SG, RG, CH, UU>(P, S, S.tab, J, (uint32_t)a, ql, tl, h, HE.data(), QS.data(), r, cells)
#define RUN_U(SG, RG, CH) do { if (variant & 1) RUN(SG, RG, CH, 4); else RUN(SG, RG, CH, 8); } while (0)
#define RUN_C(SG, RG) do { if (chunked) RUN_U(SG, RG, true); else RUN_U(SG, RG, false); } while (0)
#define RUN_R(SG) do { if (ring) RUN_C(SG, true); else RUN_C(SG, false); } while (0)
        if (same) RUN_R(true); else RUN_R(false);
        memcpy(res6 + a * 6, &r, 24);
    }
    return (long long)cells;
}

// closed_form_job over a batch, both sequence forms (bytes as given; 4-bit words packed here, 8 bases per word, base 0 in the top
// nibble, each sequence on a word boundary, padding 4).  flags[a] bit 0 / bit 1: the byte / packed form took the job; res6 holds the
// result where a flag is set (both forms must agree: -3 otherwise).  Returns the number of jobs taken, -1 when the parameters rule
// the shortcut out.
extern "C" long long ext_closed_form_host(const bwa_b200_ext_params_t *p, uint64_t n, const uint8_t *qseq, const uint32_t *qoff,
                                          const uint32_t *qlen, const uint8_t *tseq, const uint32_t *toff, const uint32_t *tlen,
                                          const uint32_t *h0, int32_t *res6, uint8_t *flags)
{
    const ClosedParams C = closed_params_from(p);
    if (!C.ok) return -1;
    std::vector<uint32_t> qp, tp, qo(n), to(n);
    auto pack = [](std::vector<uint32_t> &dst, const uint8_t *s, uint32_t len) -> uint32_t {
        const uint32_t at = (uint32_t)dst.size() * 8;
        for (uint32_t w = 0; w < (len + 7) / 8; ++w) {
            uint32_t x = 0;
            for (uint32_t k = 0; k < 8; ++k) { const uint32_t i = 8 * w + k; x = (x << 4) | (i < len ? (s[i] > 4 ? 4u : s[i]) : 4u); }
            dst.push_back(x);
        }
        return at;
    };
    for (uint64_t a = 0; a < n; ++a) { qo[a] = pack(qp, qseq + qoff[a], qlen[a]); to[a] = pack(tp, tseq + toff[a], tlen[a]); }
    qp.push_back(0); tp.push_back(0);
    JobView JB{qseq, tseq, nullptr, nullptr, qoff, qlen, toff, tlen, h0};
    JobView JP{nullptr, nullptr, qp.data(), tp.data(), qo.data(), qlen, to.data(), tlen, h0};
    long long taken = 0;
    for (uint64_t a = 0; a < n; ++a) {
        bwa_b200_ext_result_t rb, rp;
        const bool fb = closed_form_job<true>(C, JB, (uint32_t)a, &rb), fp = closed_form_job<false>(C, JP, (uint32_t)a, &rp);
        flags[a] = (uint8_t)((fb ? 1 : 0) | (fp ? 2 : 0));
        if (fb != fp || (fb && memcmp(&rb, &rp, sizeof(rb)) != 0)) return -3;
        if (fb) { memcpy(res6 + 6 * a, &rb, sizeof(rb)); ++taken; }
    }
    return taken;
}
