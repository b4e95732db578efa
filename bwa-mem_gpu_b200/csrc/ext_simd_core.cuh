// ext_simd_core.cuh -- the two-row s16x2 extension of one job, shared between the CUDA kernel
// (extend.cu) and a host build that emulates the five integer intrinsics it uses
// (tests/host_emul/ext_simd_host.cpp: the same source, run against the oracle on the CPU test box).
#pragma once
#include <stdint.h>
#include "bwamem_b200.h"

#ifdef __CUDACC__
#define B200_DEV __device__ __forceinline__
#else
#define B200_DEV static inline
#endif

// prmt.b32 in its generic mode: selector bit 3 replicates the sign of the selected byte.  (__byte_perm masks
// the selector with 0x7777, so the score lookup needs the PTX instruction itself.)
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t b200_prmt(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#endif

struct ExtParams {
    int8_t  mat[32];
    int32_t o_del, e_del, o_ins, e_ins, w, end_bonus, zdrop, use_band, pen_clip;
    int32_t max_score;           // max entry of mat (band clamp, src/ksw.c:886-887)
    int32_t bias;                // -min(mat, 0): scores are kept as unsigned bytes score + bias
};

struct JobView {
    const uint8_t  *qb, *tb;     // byte-per-base sequences (BYTES) ...
    const uint32_t *qp, *tp;     // ... or 4-bit packed (offsets in bases, multiples of 8)
    const uint32_t *qoff, *qlen, *toff, *tlen, *h0;
};

struct SimdParams {
    uint32_t tab_lo, tab_hi;                 // PRMT score table
    uint32_t noe_del2, ne_del2, noe_ins2, ne_ins2;   // negative penalties, both halves
};

#define SIMD_STEP(C)                                                                             \
    {                                                                                            \
        const uint32_t pB_ = Bp[(C) * NT];                                                       \
        const uint32_t qb_ = Qp[(C) * NT];                                                       \
        const uint32_t Hin_ = __byte_perm(pB_, aw, 0x5410);                                      \
        const uint32_t Ein_ = __byte_perm(pB_, aw, 0x7632);                                      \
        const uint32_t q16_ = qprev * 256u + qb_;                                                \
        qprev = qb_;                                                                             \
        const uint32_t S_ = b200_prmt(S.tab_lo, S.tab_hi, (q16_ ^ T16) | R16);                   \
        const uint32_t M_ = __viaddmin_s16x2(Hin_, S_, Hin_ * 32u);                              \
        const uint32_t h_ = __vimax3_s16x2(M_, Ein_, F);                                         \
        bool pH_, pL_;                                                                           \
        mm = __vibmax_s16x2(h_, mm, &pH_, &pL_);                                                 \
        mjL = pL_ ? (C) : mjL;                                                                   \
        mjH = pH_ ? (C) : mjH;                                                                   \
        const uint32_t t1_ = __viaddmax_s16x2(M_, S.noe_del2, 0u);                               \
        const uint32_t En_ = __viaddmax_s16x2(Ein_, S.ne_del2, t1_);                             \
        const uint32_t t2_ = __viaddmax_s16x2(M_, S.noe_ins2, 0u);                               \
        F = __viaddmax_s16x2(F, S.ne_ins2, t2_);                                                 \
        aw = __byte_perm(h1, En_, 0x5410);                                                       \
        Ap[(C) * NT] = aw;                                                                       \
        Bp[((C) - 1) * NT] = __byte_perm(h1, En_, 0x7632);                                       \
        h1 = h_;                                                                                 \
    }

// one cell of one row, plain integers (the few columns around the two-row sweep)
#define SCALAR_CELL(IN, OUTP, TB, COL, H1, FF, MX, MJ, MJV)                                      \
    {                                                                                            \
        const uint32_t p_ = (IN);                                                                \
        const int hd_ = (int)(p_ & 0xffffu), e_ = (int)(p_ >> 16);                               \
        const int qc_ = (int)(Qp[(COL) * NT] & 7u);                                              \
        const int M_ = hd_ ? hd_ + (int)P.mat[(TB) * 5 + qc_] : 0;                               \
        int h_ = M_ > e_ ? M_ : e_;                                                              \
        h_ = h_ > FF ? h_ : FF;                                                                  \
        MJ = MX > h_ ? MJ : (MJV);                                                               \
        MX = MX > h_ ? MX : h_;                                                                  \
        int t_ = M_ - oe_del; t_ = t_ > 0 ? t_ : 0;                                              \
        int en_ = e_ - P.e_del; en_ = en_ > t_ ? en_ : t_;                                       \
        t_ = M_ - oe_ins; t_ = t_ > 0 ? t_ : 0;                                                  \
        FF -= P.e_ins; FF = FF > t_ ? FF : t_;                                                   \
        *(OUTP) = (uint32_t)H1 | ((uint32_t)en_ << 16);                                          \
        H1 = h_;                                                                                 \
    }


// Fill the per-batch constants of the two-row kernel; returns 0 when the matrix / penalties are not eligible.
static inline int simd_params_from(const bwa_b200_ext_params_t *p, SimdParams *S)
{
    int ok = 1;
    const int a = p->mat[0], mb = p->mat[1], nn = p->mat[4];
    for (int i = 0; i < 5 && ok; ++i)
        for (int j = 0; j < 5; ++j) {
            const int want = (i == 4 || j == 4) ? nn : (i == j ? a : mb);
            if (p->mat[i * 5 + j] != want) { ok = 0; break; }
        }
    if (a < 1 || a > 31 || mb > a || nn > a) ok = 0;
    const int oe_del = p->o_del + p->e_del, oe_ins = p->o_ins + p->e_ins;
    if (oe_del < 0 || oe_ins < 0 || oe_del > 16000 || oe_ins > 16000 || p->e_del > 16000 || p->e_ins > 16000) ok = 0;
    S->tab_lo = (uint32_t)(uint8_t)(int8_t)a | (uint32_t)(uint8_t)(int8_t)mb * 0x01010100u;
    S->tab_hi = (uint32_t)(uint8_t)(int8_t)nn * 0x01010101u;
    S->noe_del2 = (uint32_t)(uint16_t)(int16_t)(-oe_del) * 0x00010001u; S->ne_del2 = (uint32_t)(uint16_t)(int16_t)(-p->e_del) * 0x00010001u;
    S->noe_ins2 = (uint32_t)(uint16_t)(int16_t)(-oe_ins) * 0x00010001u; S->ne_ins2 = (uint32_t)(uint16_t)(int16_t)(-p->e_ins) * 0x00010001u;
    return ok;
}

// One job.  Bp / Ap / Qp are this lane's columns of the [column][lane] arrays (lane stride NT words / bytes).
template <bool BYTES, int NT>
B200_DEV void simd_job(const ExtParams &P, const SimdParams &S, const JobView &J, uint32_t a, int qlen, int tlen, int h0,
                       uint32_t *const Bp, uint32_t *const Ap, uint8_t *const Qp, bwa_b200_ext_result_t &r, unsigned long long &my_cells)
{
    const uint32_t qo = J.qoff[a], to = J.toff[a];
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    // stage the query: byte c = code | code << 4 | 0x80 (two PRMT selector nibbles, the upper with sign replication)
    for (int j8 = 0; j8 < qlen; j8 += 8) {
        uint32_t wv = 0;
        if (!BYTES) wv = J.qp[(qo + j8) >> 3];
        for (int u = 0; u < 8 && j8 + u < qlen; ++u) {
            uint32_t c = BYTES ? (uint32_t)J.qb[qo + j8 + u] : (wv >> (28 - 4 * u)) & 15u;
            c = c > 4u ? 4u : c;
            Qp[(j8 + u) * NT] = (uint8_t)(c * 17u + 0x80u);
        }
    }
    // first row: H(-1,-1) = h0, then one gap open, then extensions (src/ksw.c:880-883)
    {
        int v = h0;
        Bp[0] = (uint32_t)v;
        v = h0 > oe_ins ? h0 - oe_ins : 0;
        for (int j = 1; j <= qlen; ++j) {
            Bp[j * NT] = (uint32_t)v;
            v = v > P.e_ins ? v - P.e_ins : 0;
        }
    }
    // band clamp (src/ksw.c:885-893)
    int w = P.w;
    {
        int max_ins = (int)((double)(qlen * P.max_score + P.end_bonus - P.o_ins) / P.e_ins + 1.);
        max_ins = max_ins > 1 ? max_ins : 1;
        w = w < max_ins ? w : max_ins;
        int max_del = (int)((double)(qlen * P.max_score + P.end_bonus - P.o_del) / P.e_del + 1.);
        max_del = max_del > 1 ? max_del : 1;
        w = w < max_del ? w : max_del;
    }
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    uint32_t tword = 0;
    auto target_at = [&](int i) -> int {
        int tbv;
        if (BYTES) tbv = J.tb[to + i];
        else {
            if ((i & 7) == 0) tword = J.tp[(to + i) >> 3];
            tbv = (int)((tword >> (28 - 4 * (i & 7))) & 15u);
        }
        return tbv > 4 ? 4 : tbv;
    };
    // the reference's per-row epilogue (src/ksw.c:942-959); returns true when the job ends at this row
    auto row_done = [&](int i, int jfin, int h1fin, int m, int mj) -> bool {
        if (jfin == qlen) {
            best_ie = gscore > h1fin ? best_ie : i;
            gscore = gscore > h1fin ? gscore : h1fin;
        }
        if (m == 0) return true;
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            const int d = mj > i ? mj - i : i - mj;
            max_off = max_off > d ? max_off : d;
        } else if (P.zdrop > 0) {
            const int di = i - best_i, dj = mj - best_j;
            if (di > dj) { if (best - m - (di - dj) * P.e_del > P.zdrop) return true; }
            else         { if (best - m - (dj - di) * P.e_ins > P.zdrop) return true; }
        }
        return false;
    };

    for (int i = 0; i < tlen; i += 2) {
        const bool two = i + 1 < tlen;
        const int tb0 = target_at(i);
        const int tb1 = two ? target_at(i + 1) : 4;
        if (P.use_band) {
            if (beg < i - w) beg = i - w;
            if (end > i + w + 1) end = i + w + 1;
            if (end > qlen) end = qlen;
        }
        int hstart = beg;                                  // first column of row i+1 in the sweep
        if (P.use_band && hstart < i + 1 - w) hstart = i + 1 - w;
        int h1L = 0, fL = 0, mL = 0, mjL = -1, h1H = 0, fH = 0, mH = 0, mjH = 0;   // mjH holds column + 1
        if (beg == 0) { h1L = h0 - (P.o_del + P.e_del * (i + 1)); h1L = h1L < 0 ? 0 : h1L; }
        if (hstart == 0) { h1H = h0 - (P.o_del + P.e_del * (i + 2)); h1H = h1H < 0 ? 0 : h1H; }
        int t = beg;
        // low row alone until the high row can start behind it
        for (; t < end && t <= hstart; ++t) SCALAR_CELL(Bp[t * NT], Ap + t * NT, tb0, t, h1L, fL, mL, mjL, t)
        int hc = hstart;                                   // next column of row i+1 to evaluate
        if (t < end) {
            // per row pair: target halves of the PRMT selector
            const uint32_t T16 = (tb0 < 4 ? (uint32_t)tb0 * 17u : 0u) | (tb1 < 4 ? (uint32_t)tb1 * 17u : 0u) << 8;
            const uint32_t R16 = (tb0 < 4 ? 0u : 0x44u) | (tb1 < 4 ? 0u : 0x4400u);
            uint32_t h1 = (uint32_t)h1L | (uint32_t)h1H << 16, F = (uint32_t)fL | (uint32_t)fH << 16;
            uint32_t mm = (uint32_t)mL | (uint32_t)mH << 16;
            uint32_t aw = Ap[(t - 1) * NT], qprev = Qp[(t - 1) * NT];
            for (; t + 8 <= end; t += 8) {
                SIMD_STEP(t) SIMD_STEP(t + 1) SIMD_STEP(t + 2) SIMD_STEP(t + 3)
                SIMD_STEP(t + 4) SIMD_STEP(t + 5) SIMD_STEP(t + 6) SIMD_STEP(t + 7)
            }
            for (; t < end; ++t) SIMD_STEP(t)
            h1L = (int)(h1 & 0xffffu); h1H = (int)(h1 >> 16);
            fL = (int)(F & 0xffffu); fH = (int)(F >> 16);
            mL = (int)(mm & 0xffffu); mH = (int)(mm >> 16);
            hc = end - 1;
        }
        // ---- row i is complete
        Ap[end * NT] = (uint32_t)h1L;                       // eh[end] = {h1, 0}
        my_cells += (unsigned long long)(end > beg ? end - beg : 0);
        if (row_done(i, beg < end ? end : beg, h1L, mL, mjL)) break;
        if (!two) break;
        // the reference's next window from row i's output (src/ksw.c:965-970), then the band of row i+1
        int beg1, end1;
        {
            int j = beg;
            while (j < end && Ap[j * NT] == 0u) ++j;
            beg1 = j;
            j = end;
            while (j >= beg1 && Ap[j * NT] == 0u) --j;
            end1 = j + 2 < qlen ? j + 2 : qlen;
            if (P.use_band) {
                if (beg1 < i + 1 - w) beg1 = i + 1 - w;
                if (end1 > i + 1 + w + 1) end1 = i + 1 + w + 1;
                if (end1 > qlen) end1 = qlen;
            }
        }
        // ---- complete or roll back row i+1
        int h1fin;
        if (beg1 >= end1) {                                // empty window: nothing evaluated by the reference
            h1fin = 0;
            if (beg1 == 0) { h1fin = h0 - (P.o_del + P.e_del * (i + 2)); h1fin = h1fin < 0 ? 0 : h1fin; }
            mH = 0;
        } else if (end1 >= hc) {
            int j = hc > beg1 ? hc : beg1;
            for (; j < end1; ++j) {
                const uint32_t in = j <= end ? Ap[j * NT] : Bp[j * NT];
                SCALAR_CELL(in, Bp + j * NT, tb1, j, h1H, fH, mH, mjH, j + 1)
            }
            h1fin = h1H;
        } else {
            h1fin = (int)(Bp[end1 * NT] & 0xffffu);         // H(i+1, end1-1), stored with column end1
        }
        for (int j = end1 + 1; j <= end; ++j) Bp[j * NT] = Ap[j * NT];   // columns row i+1 does not own keep row i's output
        Bp[end1 * NT] = (uint32_t)h1fin;                    // eh[end] = {h1, 0}
        my_cells += (unsigned long long)(end1 > beg1 ? end1 - beg1 : 0);
        if (row_done(i + 1, beg1 < end1 ? end1 : beg1, h1fin, mH, mjH - 1)) break;
        // window of the next row pair
        {
            int j = beg1;
            while (j < end1 && Bp[j * NT] == 0u) ++j;
            beg = j;
            j = end1;
            while (j >= beg && Bp[j * NT] == 0u) --j;
            end = j + 2 < qlen ? j + 2 : qlen;
        }
    }
    r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = best_ie + 1; r.gscore = gscore; r.max_off = max_off;
}
