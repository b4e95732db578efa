// ext_pair_core.cuh -- one ksw_extend2 job per lane, two query COLUMNS per s16x2 register.
//
// Shared between the CUDA kernel (extend.cu) and a host build that emulates the integer intrinsics it
// uses (tests/host_emul/ext_pair_host.cpp: the same source, fuzzed against the oracle on the CPU box).
//
// Why columns pair up exactly: in ksw_extend2 (src/ksw.c:921-938) M(i,j), E(i+1,j) and the gap-open
// term of F depend only on the previous row, so they are data-parallel along a row; only
// F(i,j+1) = max(F(i,j) - e_ins, max(M(i,j) - oe_ins, 0)) carries from column to column.  For the
// column pair (j, j+1) that carry is two dependent VIADDMNMX on a chain register whose HIGH half is
// F(i,j); everything else is one packed instruction for both columns.  Rows stay in the reference's
// order with the reference's window [beg, end): an odd first / even last column of the window is a
// scalar cell, so no cell outside the window is ever read or written (src/ksw.c:909-970 semantics
// are kept exactly; nothing is speculated).
//
// Per-lane state (lane stride NT elements, conflict-free):
//   HE[p] = uint2 { H(i-1, 2p-1) | H(i-1, 2p) << 16 ,  E(i, 2p) | E(i, 2p+1) << 16 }     p = 0 .. qlen/2
//           i.e. the reference's eh[j] = {H(i-1,j-1), E(i,j)} for j = 2p (low halves) and 2p+1 (high halves)
//   QS[g] = PRMT selectors of query columns 4g .. 4g+3, one byte each: code * 17 + 0x80
//           (low nibble picks the score byte, high nibble replicates its sign into the upper byte)
#pragma once
#include <stdint.h>
#include "bwamem_b200.h"

#ifdef __CUDACC__
#define B200_DEV __device__ __forceinline__
// prmt.b32 in its generic mode: selector bit 3 replicates the sign of the selected byte.  (__byte_perm masks
// the selector with 0x7777, so the score lookup needs the PTX instruction itself.)
__device__ __forceinline__ uint32_t b200_prmt(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
#else
#define B200_DEV static inline
struct uint2 { uint32_t x, y; };
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

struct PairParams {
    uint32_t tab[5];             // tab[t]: score bytes against query codes 0..3 for target base t (t = 4: N)
    uint32_t tab_n;              // byte 0: score against query code 4 (N)
    uint32_t noe_del2, ne_del2, noe_ins2, ne_ins2;   // negative penalties in both halves
};

constexpr int PAIR_MAX_SCORE = 1023;   // H * 32 must stay below 2^15 (zero test of the diagonal, see PAIR_STEP)
constexpr int PAIR_KEYED_MAX_Q = 128;  // (score, column pair) fits one 16-bit key: score < 2^10, pair index < 2^6

// Fill the per-batch constants; returns 0 when the matrix / penalties are not eligible (any matrix whose
// entries fit a signed byte with 0 < max <= 31 is; penalties must fit 16 bits with room to spare).
static inline int pair_params_from(const bwa_b200_ext_params_t *p, PairParams *S)
{
    int ok = 1, mx = 0;
    for (int i = 0; i < 25; ++i) mx = mx > p->mat[i] ? mx : p->mat[i];
    if (mx < 1 || mx > 31) ok = 0;
    const int oe_del = p->o_del + p->e_del, oe_ins = p->o_ins + p->e_ins;
    if (p->o_del < 0 || p->o_ins < 0 || p->e_del < 1 || p->e_ins < 1 || oe_del > 16000 || oe_ins > 16000) ok = 0;
    for (int t = 0; t < 5; ++t) {
        uint32_t w = 0;
        for (int q = 0; q < 4; ++q) w |= (uint32_t)(uint8_t)p->mat[t * 5 + q] << (8 * q);
        S->tab[t] = w;
        if (p->mat[t * 5 + 4] != p->mat[4]) ok = 0;          // one score for "query is N" (bwa_fill_scmat: -1 everywhere)
    }
    S->tab_n = (uint32_t)(uint8_t)p->mat[4];
    S->noe_del2 = (uint32_t)(uint16_t)(int16_t)(-oe_del) * 0x00010001u; S->ne_del2 = (uint32_t)(uint16_t)(int16_t)(-p->e_del) * 0x00010001u;
    S->noe_ins2 = (uint32_t)(uint16_t)(int16_t)(-oe_ins) * 0x00010001u; S->ne_ins2 = (uint32_t)(uint16_t)(int16_t)(-p->e_ins) * 0x00010001u;
    return ok;
}

// One column pair (2p, 2p+1) of row i.  SEL: low 16 bits = the two PRMT selector bytes of the pair.
//   Hd   = {H(i-1,2p-1), H(i-1,2p)}                     the diagonal of both columns
//   M    = min(Hd + score, Hd * 32)                      `M = M ? M + s : 0` (src/ksw.c:924): Hd == 0 gives M <= 0,
//                                                        which every later use treats like 0; Hd > 0 gives Hd + score
//   c    : chain register, HIGH half = F(i,2p); Fh.hi = F(i,2p+1); new c.hi = F(i,2p+2)
//   max(x - pen, 0) is the RELU form of VIADDMNMX with the (negative) penalty as its own third operand
//   key  = h * 64 + p (KEYED): unsigned max keeps the LAST column among equal maxima (src/ksw.c:928);
//          pp = {p0, p0} for the loop iteration's first pair, KOFF the pair's offset from it
#define PAIR_STEP(HP, PIDX, SEL, KOFF)                                                           \
    {                                                                                            \
        const uint2 he_ = *(HP);                                                                 \
        const uint32_t S_ = b200_prmt(tlo, tab_n, (SEL));                                        \
        const uint32_t M_ = __viaddmin_s16x2(he_.x, S_, he_.x * 32u);                            \
        const uint32_t t2_ = __viaddmax_s16x2_relu(M_, noe_ins2, noe_ins2);                      \
        const uint32_t Fh_ = __viaddmax_s16x2(c, ne_ins2, t2_ << 16);                            \
        const uint32_t F_ = __byte_perm(c, Fh_, 0x7632);                                         \
        c = __viaddmax_s16x2(Fh_, ne_ins2, t2_);                                                 \
        const uint32_t h_ = __vimax3_s16x2(M_, he_.y, F_);                                       \
        if (KEYED) m2 = (KOFF) ? __viaddmax_u16x2(h_ * 64u + pp, (KOFF), m2) : __vmaxu2(m2, h_ * 64u + pp); \
        else {                                                                                   \
            bool pH_, pL_;                                                                       \
            m2 = __vibmax_s16x2(h_, m2, &pH_, &pL_);                                             \
            pjL = pL_ ? (PIDX) : pjL;                                                            \
            pjH = pH_ ? (PIDX) : pjH;                                                            \
        }                                                                                        \
        const uint32_t t1_ = __viaddmax_s16x2_relu(M_, noe_del2, noe_del2);                      \
        uint2 o_;                                                                                \
        o_.y = __viaddmax_s16x2(he_.y, ne_del2, t1_);                                            \
        o_.x = __byte_perm(hprev, h_, 0x5432);                                                   \
        *(HP) = o_;                                                                              \
        hprev = h_;                                                                              \
    }

// One job.  HEp / QSp are this lane's element 0 of the [index][lane] arrays.
template <bool BYTES, int NT, bool KEYED>
B200_DEV void pair_job(const ExtParams &P, const PairParams &S, const JobView &J, uint32_t a, int qlen, int tlen, int h0,
                       uint2 *const HEp, uint32_t *const QSp, bwa_b200_ext_result_t &r, unsigned long long &my_cells)
{
    const uint32_t qo = J.qoff[a], to = J.toff[a];
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    uint16_t *const hw = reinterpret_cast<uint16_t *>(HEp);
#define H16(j) hw[((j) >> 1) * (NT * 4) + ((j) & 1)]
#define E16(j) hw[((j) >> 1) * (NT * 4) + 2 + ((j) & 1)]
    // stage the query as PRMT selector bytes, four columns per word (columns >= qlen: N, never evaluated)
    for (int j8 = 0; j8 < qlen; j8 += 8) {
        uint32_t wv = 0;
        if (!BYTES) wv = J.qp[(qo + j8) >> 3];
        uint32_t s0 = 0, s1 = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t cde = 4u;
            if (j8 + u < qlen) { cde = BYTES ? (uint32_t)J.qb[qo + j8 + u] : (wv >> (28 - 4 * u)) & 15u; cde = cde > 4u ? 4u : cde; }
            const uint32_t sb = cde * 17u + 0x80u;
            if (u < 4) s0 |= sb << (8 * u); else s1 |= sb << (8 * (u - 4));
        }
        QSp[(j8 >> 2) * NT] = s0;
        if (j8 + 4 < qlen) QSp[((j8 >> 2) + 1) * NT] = s1;
    }
    // first row: H(-1,-1) = h0, then one gap open, then extensions (src/ksw.c:880-883); E = 0
    {
        int v = h0 > oe_ins ? h0 - oe_ins : 0;       // eh[1].h
        uint32_t lo = (uint32_t)h0;
        for (int p = 0; 2 * p <= qlen; ++p) {
            uint2 o;
            o.x = lo | (uint32_t)v << 16;            // eh[2p].h, eh[2p+1].h
            o.y = 0u;
            HEp[p * NT] = o;
            v = v > P.e_ins ? v - P.e_ins : 0;
            lo = (uint32_t)v;                        // eh[2p+2].h
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
    // per-batch constants, made to depend on a run-time zero so that they stay in registers (ptxas otherwise
    // re-reads the kernel parameter bank inside the column loop)
    const uint32_t rz = (uint32_t)tlen >> 31;
    const uint32_t tab_n = S.tab_n + rz, noe_del2 = S.noe_del2 + rz, ne_del2 = S.ne_del2 + rz, noe_ins2 = S.noe_ins2 + rz, ne_ins2 = S.ne_ins2 + rz;
    for (int i = 0; i < tlen; ++i) {
        int tbv;
        if (BYTES) tbv = J.tb[to + i];
        else {
            if ((i & 7) == 0) tword = J.tp[(to + i) >> 3];
            tbv = (int)((tword >> (28 - 4 * (i & 7))) & 15u);
        }
        tbv = tbv > 4 ? 4 : tbv;
        if (P.use_band) {
            if (beg < i - w) beg = i - w;
            if (end > i + w + 1) end = i + w + 1;
            if (end > qlen) end = qlen;
        }
        const uint32_t tlo = tbv == 0 ? S.tab[0] : (tbv == 1 ? S.tab[1] : (tbv == 2 ? S.tab[2] : (tbv == 3 ? S.tab[3] : S.tab[4])));
        int h1 = 0, f = 0, m = 0, mj = -1;
        if (beg == 0) { h1 = h0 - (P.o_del + P.e_del * (i + 1)); h1 = h1 < 0 ? 0 : h1; }
        // one cell, plain integers (src/ksw.c:921-938): the odd first / even last column of the window
#define PAIR_SCALAR_CELL(COL)                                                                    \
        {                                                                                        \
            const int j_ = (COL);                                                                \
            const int hd_ = (int)H16(j_), e0_ = (int)E16(j_);                                    \
            const uint32_t qs_ = (QSp[(j_ >> 2) * NT] >> (8 * (j_ & 3))) & 0xffu;                \
            const int sc_ = (int)(int16_t)(uint16_t)b200_prmt(tlo, tab_n, qs_);                \
            H16(j_) = (uint16_t)h1;                                                              \
            const int M_ = hd_ ? hd_ + sc_ : 0;                                                  \
            int h_ = M_ > e0_ ? M_ : e0_;                                                        \
            h_ = h_ > f ? h_ : f;                                                                \
            h1 = h_;                                                                             \
            mj = m > h_ ? mj : j_;                                                               \
            m = m > h_ ? m : h_;                                                                 \
            int t_ = M_ - oe_del; t_ = t_ > 0 ? t_ : 0;                                          \
            int en_ = e0_ - P.e_del; en_ = en_ > t_ ? en_ : t_;                                  \
            E16(j_) = (uint16_t)en_;                                                             \
            t_ = M_ - oe_ins; t_ = t_ > 0 ? t_ : 0;                                              \
            f -= P.e_ins; f = f > t_ ? f : t_;                                                   \
        }
        int j = beg;
        if ((j & 1) && j < end) { PAIR_SCALAR_CELL(j) ++j; }
        const int p_end = end >> 1;                      // pairs [j/2, end/2) lie wholly inside the window
        if ((j >> 1) < p_end) {
            uint32_t hprev = (uint32_t)h1 << 16, c = (uint32_t)f << 16, m2 = 0u;
            int p = j >> 1;
            uint32_t pp = (uint32_t)p * 0x00010001u;
            int pjL = p, pjH = p;
            uint2 *hp = HEp + p * NT;
            const uint32_t *qp = QSp + (p >> 1) * NT;
            if (p & 1) { const uint32_t qw = *qp; PAIR_STEP(hp, p, qw >> 16, 0u) ++p; pp += 0x00010001u; hp += NT; qp += NT; }
#pragma unroll 2
            for (; p + 2 <= p_end; p += 2, pp += 0x00020002u, hp += 2 * NT, qp += NT) {
                const uint32_t qw = *qp;
                PAIR_STEP(hp, p, qw, 0u)
                PAIR_STEP(hp + NT, p + 1, qw >> 16, 0x00010001u)
            }
            if (p < p_end) { const uint32_t qw = *qp; PAIR_STEP(hp, p, qw, 0u) ++p; }
            h1 = (int)(hprev >> 16); f = (int)(c >> 16);
            // row maximum of the paired columns; among equal maxima the last column wins (src/ksw.c:928)
            int sL, cL, sH, cH;
            if (KEYED) {
                const uint32_t kl = m2 & 0xffffu, kh = m2 >> 16;
                sL = (int)(kl >> 6); cL = (int)(kl & 63u) * 2; sH = (int)(kh >> 6); cH = (int)(kh & 63u) * 2 + 1;
            } else {
                sL = (int)(m2 & 0xffffu); cL = pjL * 2; sH = (int)(m2 >> 16); cH = pjH * 2 + 1;
            }
            const bool hi_wins = sH > sL || (sH == sL && cH > cL);
            const int sP = hi_wins ? sH : sL, cP = hi_wins ? cH : cL;
            if (sP >= m) { m = sP; mj = cP; }            // every paired column lies right of the scalar head cell
            j = p << 1;
        }
        if (j < end) { PAIR_SCALAR_CELL(j) ++j; }
        // ---- row i is complete (src/ksw.c:940-959)
        H16(end) = (uint16_t)h1; E16(end) = 0;           // eh[end] = {h1, 0}
        my_cells += (unsigned long long)(end > beg ? end - beg : 0);
        if ((beg < end ? end : beg) == qlen) {
            best_ie = gscore > h1 ? best_ie : i;
            gscore = gscore > h1 ? gscore : h1;
        }
        if (m == 0) break;
        if (m > best) {
            best = m; best_i = i; best_j = mj;
            const int d = mj > i ? mj - i : i - mj;
            max_off = max_off > d ? max_off : d;
        } else if (P.zdrop > 0) {
            const int di = i - best_i, dj = mj - best_j;
            if (di > dj) { if (best - m - (di - dj) * P.e_del > P.zdrop) break; }
            else         { if (best - m - (dj - di) * P.e_ins > P.zdrop) break; }
        }
        // window of the next row (src/ksw.c:965-970)
        j = beg;
        while (j < end && H16(j) == 0 && E16(j) == 0) ++j;
        beg = j;
        j = end;
        while (j >= beg && H16(j) == 0 && E16(j) == 0) --j;
        end = j + 2 < qlen ? j + 2 : qlen;
    }
#undef PAIR_SCALAR_CELL
#undef H16
#undef E16
    r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = best_ie + 1; r.gscore = gscore; r.max_off = max_off;
}
