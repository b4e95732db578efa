// sw_stripe.cuh -- ksw_align2's byte kernel (ksw_u8, src/ksw.c:440-572) with the DP in registers: one job per group of EIGHT lanes.
//
// The reference's kernel is a striped SSE2 Smith-Waterman: query position l * slen + j sits in byte lane l of vector j (16 lanes,
// slen = ceil(qlen / 16) vectors per row).  Its results are those of that vector program, not of the textbook recurrence (F restarts
// at the head of every lane, the lazy-F loop raises H without correcting E and leaves at its first quiet step, bytes saturate), so
// sw_core.cuh replays it position by position, one job per lane, with the state in global memory: 38 GCUPS on a B200, slower than the
// CPU.  Here the vector program itself runs on a lane group: thread t of a group holds byte lanes 2t and 2t + 1 of every vector as
// the two halves of an s16x2 register (the byte values 0 .. 255 are exact in 16 bits; saturation becomes a min / a relu), so
//   * a vector operation is ONE packed DPX instruction per thread (VIADDMNMX / VIMNMX3 on s16x2) for two cells,
//   * H, E and Hmax of the whole query (slen vectors) stay in registers -- 3 * SMAX registers per thread, SMAX = 8 / 12 / 16 --,
//   * `_mm_slli_si128(v, 1)` is one SHFL from the neighbouring thread, the row maximum three SHFL.XOR steps, the lazy-F exit test one
//     ballot masked to the group's eight bits; four jobs share a warp and advance row by row in lockstep.
// Vectors are kept in reverse order (register k = vector slen - 1 - k) so that the vector the next row's first step reads, slen - 1,
// is register 0 whatever the job's slen, and every register index is a compile-time constant.
// The row maxima (the reference's list b of rows at or above minsc, src/ksw.c:526-537,559-568) are bytes in shared memory; the list is
// rebuilt from them at the end exactly as sw_core.cuh does.  Jobs outside the class (16-bit kernel, more than 16 * SMAX query bases,
// targets beyond the shared-memory row buffer) run in the replay kernel of sw.cu.
#pragma once
#include <stdint.h>
#include "bwamem_b200.h"

namespace b200sw {

constexpr int GRP = 8;                       // lanes per job
constexpr int JOBS_PER_BLOCK = 32;           // 256 threads
constexpr int BLOCK = GRP * JOBS_PER_BLOCK;

struct StripeParams {
    uint32_t tab[5];        // tab[t]: scores of target code t against query codes 0..3, one signed byte each
    uint32_t tabn[5];       // byte 0: score of target code t against a query N; byte 1: 0 (a padded query position); byte 2: 0x80 (see pass_u8)
    int32_t  shift, qmax;   // ksw_qinit's bias (-min of the matrix) and the largest matrix entry
    uint32_t noe_del2, ne_del2, noe_ins2, ne_ins2;   // negative penalties in both halves
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
// _mm_slli_si128(v, 1) over the group: byte lane l takes lane l - 1, lane 0 takes 0
__device__ __forceinline__ uint32_t lane_shift_up(uint32_t v, int gt)
{
    const uint32_t up = __shfl_up_sync(0xffffffffu, v, 1);
    return (gt ? (up >> 16) : 0u) | (v << 16);
}
__device__ __forceinline__ int group_max(uint32_t v)
{
    int m = max((int)(int16_t)(v & 0xffffu), (int)(int16_t)(v >> 16));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
    return max(m, __shfl_xor_sync(0xffffffffu, m, 4));
}
__device__ __forceinline__ uint32_t umax32(uint32_t a, uint32_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint32_t in_reg(uint32_t v)
{
    uint32_t r;
    asm volatile("mov.b32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

struct Seq {                // as SwSeq of sw_core.cuh: the first `rev` bases reversed, the rest as given
    const uint8_t *p;
    int rev;
    __device__ __forceinline__ int operator()(int i) const { const int c = p[i < rev ? rev - 1 - i : i]; return c > 4 ? 4 : c; }
};

// One pass of ksw_u8 for the group's job (job == false: the group only takes part in the warp's shuffles and votes).
// rowmax: the job's row-maximum bytes in shared memory, or null when the list is not wanted (no KSW_XSUBO).
// UNI: every job of the warp has exactly SMAX vectors (one read length and a kernel instantiated for it: the usual batch), so the row
// loop has no predicate and no bound check at all; otherwise the registers a job does not use (they come first in the loop) carry a
// selector that scores -32768, which leaves their H, E, the F chain and the row maximum at zero, and only the diagonal carried from
// vector to vector takes a SEL.
// SG: o_del + e_del == o_ins + e_ins (one gap-open term per vector instead of two).
template <int SMAX, bool UNI, bool SG>
__device__ __forceinline__ void pass_u8(bool job, const Seq &Q, int qlen, const Seq &T, int tlen, const StripeParams &S, const uint2 *tabs, int xtra,
                                        int gt, uint8_t *rowmax, bwa_b200_sw_result_t &r)
{
    const unsigned FULL = 0xffffffffu;
    const int gbase = (threadIdx.x & 31) & ~(GRP - 1);
    const int slen = job ? (qlen + 15) >> 4 : 0;
    const int minsc = (xtra & 0x40000) ? xtra & 0xffff : 0x10000, endsc = (xtra & 0x20000) ? xtra & 0xffff : 0x10000;
    // the constants of the row loop are pinned to registers (left alone the compiler re-reads each from the parameter bank at every use:
    // 63 LDC per row in the first version of this kernel)
    const uint32_t cap2 = in_reg((uint32_t)(255 - S.shift) * 0x00010001u);
    const uint32_t noe_del2 = in_reg(S.noe_del2), ne_del2 = in_reg(S.ne_del2), noe_ins2 = in_reg(S.noe_ins2), ne_ins2 = in_reg(S.ne_ins2);
    const int shift = S.shift;
    const int kmax = UNI ? SMAX : __reduce_max_sync(FULL, slen);          // registers at or beyond it hold no vector of any of the warp's jobs
    uint32_t H[SMAX], E[SMAX], Hm[SMAX], sel[SMAX];
#pragma unroll
    for (int k = 0; k < SMAX; ++k) {
        H[k] = 0; E[k] = 0; Hm[k] = 0;
        uint32_t s = 0x65656565u;                            // both halves 0x8000: tabn byte 1 (0x00) below tabn byte 2 (0x80)
        if (k < slen) {      // vector j = slen - 1 - k: positions (2 gt) * slen + j and (2 gt + 1) * slen + j
            const int j = slen - 1 - k, p0 = 2 * gt * slen + j, p1 = p0 + slen;
            const uint32_t c0 = p0 < qlen ? (uint32_t)Q(p0) : 5u, c1 = p1 < qlen ? (uint32_t)Q(p1) : 5u;
            s = (c0 * 17u + 0x80u) | (c1 * 17u + 0x80u) << 8;     // low nibble picks the score byte, high nibble replicates its sign
        }
        sel[k] = s;
    }
    int gmax = 0, te = -1, n_rows = 0;
    bool fin = false;
    // the target base of row i + 2 is loaded during row i and its score tables (shared memory) during row i + 1: nothing the row's
    // first vector needs is a load issued in the same row
    int t_b = (job && tlen > 1) ? T(1) : 0;
    uint2 tb = tabs[(job && tlen > 0) ? T(0) : 0];
    for (int i = 0;; ++i) {
        const bool act = job && !fin && i < tlen;
        if (!__any_sync(FULL, act)) break;
        const uint32_t tab = tb.x, tabn = tb.y;
        tb = tabs[t_b];
        t_b = (job && i + 2 < tlen) ? T(i + 2) : 0;
        // ---- the row's vectors in order (src/ksw.c:490-512)
        uint32_t hin = lane_shift_up(H[0], gt), f = 0, rm = 0;
#pragma unroll
        for (int k = SMAX - 1; k >= 0; --k) {
            if (k >= kmax) continue;
            const uint32_t sc = prmt(tab, tabn, sel[k]);
            uint32_t h = __viaddmin_s16x2(hin, sc, cap2);                   // adds_epu8(h, profile), subs_epu8(h, shift): the relu comes with the max below
            if (UNI) hin = H[k]; else hin = k < slen ? H[k] : hin;
            const uint32_t e = E[k];
            h = __vimax3_s16x2(h, e, f);
            rm = __vimax3_s16x2(rm, h, h);
            H[k] = h;
            const uint32_t t1 = __viaddmax_s16x2(h, noe_del2, 0u);
            E[k] = __viaddmax_s16x2(e, ne_del2, t1);
            const uint32_t t2 = SG ? t1 : __viaddmax_s16x2(h, noe_ins2, 0u);
            f = __viaddmax_s16x2(f, ne_ins2, t2);
        }
        // ---- lazy F (src/ksw.c:513-524): up to 16 rounds over the vectors, left at the first vector where no lane's F exceeds H - oe_ins.
        // In a row that crosses a good hit F reaches far to the right of the diagonal (it decays by e_ins per column), so most rows walk
        // several vectors here: the loop is branch-free per lane.  A group that has left (or a finished job) carries F = 0, which makes
        // every later step a no-op on its H (H >= 0); the warp leaves when every lane's F is zero.
        if (!act) f = 0;
        for (int round = 0; round < 16; ++round) {
            if (!__any_sync(FULL, f != 0u)) break;
            f = lane_shift_up(f, gt);
#pragma unroll
            for (int k = SMAX - 1; k >= 0; --k) {
                if (k >= kmax) continue;
                const bool on = UNI || k < slen;
                const uint32_t fe = on ? f : 0u;
                const uint32_t h = __vimax3_s16x2(H[k], fe, fe);
                H[k] = h;
                const uint32_t h2 = __viaddmax_s16x2(h, noe_ins2, 0u);
                const uint32_t f1 = __viaddmax_s16x2(fe, ne_ins2, 0u);
                const uint32_t b = __ballot_sync(FULL, __vimax3_s16x2(f1, h2, h2) != h2);
                if (on) f = ((b >> gbase) & 0xffu) ? f1 : 0u;
                if ((k & 1) == 0 && !__any_sync(FULL, f != 0u)) break;
            }
        }
        // ---- row bookkeeping (src/ksw.c:525-548)
        const int rowm = group_max(rm);
        if (act) {
            if (rowmax && gt == 0) rowmax[i] = (uint8_t)rowm;
            n_rows = i + 1;
            if (rowm > gmax) {
                gmax = rowm; te = i;
#pragma unroll
                for (int k = 0; k < SMAX; ++k) Hm[k] = H[k];
                if (gmax + shift >= 255 || gmax >= endsc) fin = true;
            }
        }
    }
    r.score = gmax + shift < 255 ? gmax : 255; r.te = te; r.qe = -1; r.score2 = -1; r.te2 = -1; r.tb = -1; r.qb = -1;
    // ---- qe: the smallest position among the maxima of the saved row (src/ksw.c:551-558); every lane of the group gets it
    uint32_t best = 0;            // value << 16 | 0xffff - position
#pragma unroll
    for (int k = 0; k < SMAX; ++k) {
        if (k < slen) {
            const int j = slen - 1 - k, p0 = 2 * gt * slen + j, p1 = p0 + slen;
            best = umax32(best, (Hm[k] & 0xffffu) << 16 | (uint32_t)(0xffff - p0));
            best = umax32(best, (Hm[k] & 0xffff0000u) | (uint32_t)(0xffff - p1));
        }
    }
    best = umax32(best, __shfl_xor_sync(FULL, best, 1));
    best = umax32(best, __shfl_xor_sync(FULL, best, 2));
    best = umax32(best, __shfl_xor_sync(FULL, best, 4));
    if (!job || r.score == 255) return;
    r.qe = 0xffff - (int)(best & 0xffffu);
    // ---- the second-best hit from the rows at or above minsc (src/ksw.c:526-537,559-568), rebuilt as in sw_core.cuh
    if (minsc <= 0xffff && rowmax && gt == 0) {
        const int w = (r.score + S.qmax - 1) / S.qmax, low = te - w, high = te + w;
        int last_sc = -1, last_e = -1;
        for (int i = 0; i < n_rows; ++i) {
            const int v = rowmax[i];
            if (v < minsc) continue;
            if (last_e < 0 || last_e + 1 != i) {
                if (last_e >= 0 && (last_e < low || last_e > high) && last_sc > r.score2) { r.score2 = last_sc; r.te2 = last_e; }
                last_sc = v; last_e = i;
            } else if (last_sc < v) { last_sc = v; last_e = i; }
        }
        if (last_e >= 0 && (last_e < low || last_e > high) && last_sc > r.score2) { r.score2 = last_sc; r.te2 = last_e; }
    }
}

template <int SMAX, bool SG>
__device__ __forceinline__ void pass_any(bool job, const Seq &Q, int qlen, const Seq &T, int tlen, const StripeParams &S, const uint2 *tabs, int xtra,
                                         int gt, uint8_t *rowmax, bwa_b200_sw_result_t &r)
{
    const int slen = job ? (qlen + 15) >> 4 : 0;
    if (__all_sync(0xffffffffu, !job || slen == SMAX)) pass_u8<SMAX, true, SG>(job, Q, qlen, T, tlen, S, tabs, xtra, gt, rowmax, r);
    else pass_u8<SMAX, false, SG>(job, Q, qlen, T, tlen, S, tabs, xtra, gt, rowmax, r);
}

// ---- the 16-bit kernel's score alone (ksw_i16, src/ksw.c:574-696, as mem_seed_sw reads it: xtra = KSW_XSTART, only the score used,
// src/bwamem.c:774-808): eight SSE lanes = a group of FOUR threads, up to SMAX vectors (8 SMAX query bases) in registers, eight jobs
// per warp.  No saved row, no row list, no early stop: H, E and the selectors are 3 SMAX registers.  Same conventions as pass_u8
// (reverse register order, unused registers scoring -32768, branch-free lazy-F loop); no cap on H: the 16-bit adds cannot saturate
// for windows of under 200 bases.  Q(pos) / T(row) return codes 0..4.
__device__ __forceinline__ uint32_t lane_shift_up4(uint32_t v, int gt) { return lane_shift_up(v, gt); }      // thread gt - 1 is in the same group for gt > 0
__device__ __forceinline__ int group_max4(uint32_t v)
{
    int m = max((int)(int16_t)(v & 0xffffu), (int)(int16_t)(v >> 16));
    m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
    return max(m, __shfl_xor_sync(0xffffffffu, m, 2));
}
template <int SMAX, class QF, class TF>
__device__ __forceinline__ int pass_i16_score(bool job, const QF &Q, int qlen, const TF &T, int tlen, const uint2 *tabs,
                                              uint32_t noe_del2_, uint32_t ne_del2_, uint32_t noe_ins2_, uint32_t ne_ins2_, int gt)
{
    const unsigned FULL = 0xffffffffu;
    const int gbase = (threadIdx.x & 31) & ~3;
    const int slen = job ? (qlen + 7) >> 3 : 0;
    const uint32_t noe_del2 = in_reg(noe_del2_), ne_del2 = in_reg(ne_del2_), noe_ins2 = in_reg(noe_ins2_), ne_ins2 = in_reg(ne_ins2_);
    const int kmax = __reduce_max_sync(FULL, slen);
    uint32_t H[SMAX], E[SMAX], sel[SMAX];
#pragma unroll
    for (int k = 0; k < SMAX; ++k) {
        H[k] = 0; E[k] = 0;
        uint32_t s = 0x65656565u;
        if (k < slen) {      // vector j = slen - 1 - k: positions (2 gt) * slen + j and (2 gt + 1) * slen + j
            const int j = slen - 1 - k, p0 = 2 * gt * slen + j, p1 = p0 + slen;
            const uint32_t c0 = p0 < qlen ? (uint32_t)Q(p0) : 5u, c1 = p1 < qlen ? (uint32_t)Q(p1) : 5u;
            s = (c0 * 17u + 0x80u) | (c1 * 17u + 0x80u) << 8;
        }
        sel[k] = s;
    }
    int gmax = 0;
    int t_b = (job && tlen > 1) ? T(1) : 0;
    uint2 tb = tabs[(job && tlen > 0) ? T(0) : 0];
    for (int i = 0;; ++i) {
        const bool act = job && i < tlen;
        if (!__any_sync(FULL, act)) break;
        const uint32_t tab = tb.x, tabn = tb.y;
        tb = tabs[t_b];
        t_b = (job && i + 2 < tlen) ? T(i + 2) : 0;
        uint32_t hin = lane_shift_up4(H[0], gt), f = 0, rm = 0;
#pragma unroll
        for (int k = SMAX - 1; k >= 0; --k) {
            if (k >= kmax) continue;
            const uint32_t sc = prmt(tab, tabn, sel[k]);
            const uint32_t e = E[k];
            uint32_t h = __viaddmax_s16x2(hin, sc, e);                      // adds_epi16(h, profile), max with E
            hin = k < slen ? H[k] : hin;
            h = __vimax3_s16x2(h, f, f);
            rm = __vimax3_s16x2(rm, h, h);
            H[k] = h;
            const uint32_t t1 = __viaddmax_s16x2(h, noe_del2, 0u);
            E[k] = __viaddmax_s16x2(e, ne_del2, t1);
            const uint32_t t2 = __viaddmax_s16x2(h, noe_ins2, 0u);
            f = __viaddmax_s16x2(f, ne_ins2, t2);
        }
        if (!act) f = 0;
        for (int round = 0; round < 16; ++round) {
            if (!__any_sync(FULL, f != 0u)) break;
            f = lane_shift_up4(f, gt);
#pragma unroll
            for (int k = SMAX - 1; k >= 0; --k) {
                if (k >= kmax) continue;
                const bool on = k < slen;
                const uint32_t fe = on ? f : 0u;
                const uint32_t h = __vimax3_s16x2(H[k], fe, fe);
                H[k] = h;
                const uint32_t h2 = __viaddmax_s16x2(h, noe_ins2, 0u);
                const uint32_t f1 = __viaddmax_s16x2(fe, ne_ins2, 0u);
                const uint32_t b = __ballot_sync(FULL, __vimax3_s16x2(f1, h2, h2) != h2);
                if (on) f = ((b >> gbase) & 0xfu) ? f1 : 0u;
                if ((k & 1) == 0 && !__any_sync(FULL, f != 0u)) break;
            }
        }
        const int rowm = group_max4(rm);
        if (act) gmax = max(gmax, rowm);
    }
    return gmax;
}

// ksw_align2 (src/ksw.c:698-736) for the byte kernel: the pass, then -- with KSW_XSTART and a score that reaches the threshold -- the
// pass on the reversed prefixes that yields the start of the hit.
template <int SMAX, bool SG>
__global__ void __launch_bounds__(BLOCK)
sw_stripe_kernel(StripeParams S, uint32_t n_fast, const uint32_t *__restrict__ jobs, const uint8_t *__restrict__ qseq, const uint32_t *__restrict__ qoff,
                 const uint32_t *__restrict__ qlen, const uint8_t *__restrict__ tseq, const uint32_t *__restrict__ toff, const uint32_t *__restrict__ tlen,
                 const uint32_t *__restrict__ xtra, uint32_t t_cap, bwa_b200_sw_result_t *__restrict__ res)
{
    extern __shared__ uint8_t rowmax_all[];
    __shared__ uint2 tabs[8];
    if (threadIdx.x < 5) tabs[threadIdx.x] = make_uint2(S.tab[threadIdx.x], S.tabn[threadIdx.x]);
    __syncthreads();
    const int grp = threadIdx.x / GRP, gt = threadIdx.x % GRP;
    uint8_t *rowmax = rowmax_all + (size_t)grp * t_cap;
    for (uint32_t base = blockIdx.x * JOBS_PER_BLOCK; base < n_fast; base += gridDim.x * JOBS_PER_BLOCK) {
        const uint32_t k = base + (uint32_t)grp;
        const bool job = k < n_fast;
        const uint32_t a = job ? jobs[k] : 0u;
        const int ql = job ? (int)qlen[a] : 0, tl = job ? (int)tlen[a] : 0, xt = job ? (int)xtra[a] : 0;
        const uint8_t *q = qseq + (job ? qoff[a] : 0u), *t = tseq + (job ? toff[a] : 0u);
        bwa_b200_sw_result_t r;
        pass_any<SMAX, SG>(job, Seq{q, 0}, ql, Seq{t, 0}, tl, S, tabs, xt, gt, rowmax, r);
        const bool second = job && (xt & 0x80000) != 0 && !((xt & 0x40000) && r.score < (xt & 0xffff)) && r.qe >= 0 && r.te >= 0;
        if (__any_sync(0xffffffffu, second)) {
            bwa_b200_sw_result_t rr;
            pass_any<SMAX, SG>(second, Seq{q, r.qe + 1}, r.qe + 1, Seq{t, r.te + 1}, tl, S, tabs, 0x20000 | r.score, gt, nullptr, rr);
            if (second && r.score == rr.score) { r.tb = r.te - rr.te; r.qb = r.qe - rr.qe; }
        }
        if (job && gt == 0) res[a] = r;
        __syncwarp();
    }
}

} // namespace b200sw
