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
// order with the reference's window [beg, end): when the window starts at an odd / ends at an even
// column, the outside half of the first / last pair is masked (inputs zeroed, stored state kept,
// excluded from the row maximum), so the state of a cell outside the window is never changed
// (src/ksw.c:909-970 semantics are kept exactly; nothing is speculated).
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
#define PAIR_CORE(HE, SEL)                                                                       \
        const uint32_t S_ = b200_prmt(tlo, tab_n, (SEL));                                        \
        const uint32_t M_ = __viaddmin_s16x2((HE).x, S_, (HE).x * 32u);                          \
        const uint32_t t2_ = __viaddmax_s16x2_relu(M_, noe_ins2, noe_ins2);                      \
        const uint32_t Fh_ = __viaddmax_s16x2(c, ne_ins2, t2_ << 16);                            \
        const uint32_t F_ = __byte_perm(c, Fh_, 0x7632);                                         \
        c = __viaddmax_s16x2(Fh_, ne_ins2, t2_);                                                 \
        const uint32_t h_ = __vimax3_s16x2(M_, (HE).y, F_);                                      \
        const uint32_t t1_ = __viaddmax_s16x2_relu(M_, noe_del2, noe_del2);                      \
        const uint32_t En_ = __viaddmax_s16x2((HE).y, ne_del2, t1_);

#define PAIR_MAX(HK, PIDX, KOFF)                                                                 \
        if (KEYED) m2 = (KOFF) ? __viaddmax_u16x2((HK) * 64u + pp, (KOFF), m2) : __vmaxu2(m2, (HK) * 64u + pp); \
        else {                                                                                   \
            bool pH_, pL_;                                                                       \
            m2 = __vibmax_s16x2((HK), m2, &pH_, &pL_);                                           \
            pjL = pL_ ? (PIDX) : pjL;                                                            \
            pjH = pH_ ? (PIDX) : pjH;                                                            \
        }

// a pair wholly inside the window
#define PAIR_STEP(HP, PIDX, SEL, KOFF)                                                           \
    {                                                                                            \
        const uint2 he_ = *(HP);                                                                 \
        PAIR_CORE(he_, SEL)                                                                      \
        PAIR_MAX(h_, PIDX, KOFF)                                                                 \
        uint2 o_;                                                                                \
        o_.y = En_;                                                                              \
        o_.x = __byte_perm(hprev, h_, 0x5432);                                                   \
        *(HP) = o_;                                                                              \
        hprev = h_;                                                                              \
    }

// the first / last pair of the window.  LO_OUT: the low column (beg - 1) lies outside: its inputs are zeroed, which
// makes its H, F and gap-open terms 0 = the reference's initial h1 and f of the row, and its stored state is kept.
// HI_OUT: the high column (end) lies outside: it is computed and discarded, except that eh[end] = {H(i,end-1), 0}
// (src/ksw.c:940) is exactly what the pair store leaves there.
#define PAIR_STEP_EDGE(HP, PIDX, SEL, LO_OUT, HI_OUT)                                            \
    {                                                                                            \
        const uint2 old_ = *(HP);                                                                \
        const uint32_t inm_ = (LO_OUT) ? 0xffff0000u : 0xffffffffu;                              \
        const uint32_t outm_ = (HI_OUT) ? (inm_ & 0x0000ffffu) : inm_;                           \
        uint2 he_;                                                                               \
        he_.x = old_.x & inm_; he_.y = old_.y & inm_;                                            \
        if (LO_OUT) hprev = old_.x << 16;                                                        \
        PAIR_CORE(he_, SEL)                                                                      \
        const uint32_t hk_ = (HI_OUT) ? (h_ & 0x0000ffffu) : h_;                                 \
        PAIR_MAX(hk_, PIDX, 0u)                                                                  \
        uint2 o_;                                                                                \
        o_.y = (En_ & outm_) | (old_.y & ~inm_);                                                 \
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
    const int oe_ins = P.o_ins + P.e_ins;
    uint16_t *const hw = reinterpret_cast<uint16_t *>(HEp);
#define H16(j) hw[((j) >> 1) * (NT * 4) + ((j) & 1)]
#define E16(j) hw[((j) >> 1) * (NT * 4) + 2 + ((j) & 1)]
    const uint16_t *const qs16 = reinterpret_cast<const uint16_t *>(QSp);   // selector pair of column pair p: qs16[(p >> 1) * (NT * 2) + (p & 1)]
    // stage the query as PRMT selector bytes, four columns per word, through column qlen (columns >= qlen: N)
    for (int j8 = 0; j8 <= qlen; j8 += 8) {
        uint32_t wv = 0;
        if (!BYTES && j8 < qlen) wv = J.qp[(qo + j8) >> 3];
        uint32_t s0 = 0, s1 = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            uint32_t cde = 4u;
            if (j8 + u < qlen) { cde = BYTES ? (uint32_t)J.qb[qo + j8 + u] : (wv >> (28 - 4 * u)) & 15u; cde = cde > 4u ? 4u : cde; }
            const uint32_t sb = cde * 17u + 0x80u;
            if (u < 4) s0 |= sb << (8 * u); else s1 |= sb << (8 * (u - 4));
        }
        QSp[(j8 >> 2) * NT] = s0;
        if (j8 + 4 <= qlen) QSp[((j8 >> 2) + 1) * NT] = s1;
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
        int h1 = 0, m = 0, mj = -1;
        if (beg == 0) { h1 = h0 - (P.o_del + P.e_del * (i + 1)); h1 = h1 < 0 ? 0 : h1; }
        if (beg < end) {
            const int p0 = beg >> 1, p1 = (end - 1) >> 1;          // first and last column pair of the window
            const bool lo_out = (beg & 1) != 0, hi_out = (end & 1) != 0;
            uint32_t hprev = (uint32_t)h1 << 16, c = 0u, m2 = 0u;   // c.hi = F(i, beg) = 0
            uint32_t pp = (uint32_t)p0 * 0x00010001u;
            int pjL = p0, pjH = p0;
            uint2 *hp = HEp + p0 * NT;
            const uint16_t *qp = qs16 + (p0 >> 1) * (NT * 2) + (p0 & 1);
            PAIR_STEP_EDGE(hp, p0, (uint32_t)*qp, lo_out, hi_out && p0 == p1)
            int p = p0 + 1;
            pp += 0x00010001u; hp += NT; qp += (p0 & 1) ? (NT * 2 - 1) : 1;
            for (; p + 4 <= p1; p += 4, pp += 0x00040004u, hp += 4 * NT, qp += 2 * (NT * 2)) {
                const int o1 = (p & 1) ? (NT * 2 - 1) : 1;          // selectors of pairs p .. p+3 (two per word, words NT apart)
                const uint32_t s0 = qp[0], s1 = qp[o1], s2 = qp[NT * 2], s3 = qp[NT * 2 + o1];
                PAIR_STEP(hp, p, s0, 0u)
                PAIR_STEP(hp + NT, p + 1, s1, 0x00010001u)
                PAIR_STEP(hp + 2 * NT, p + 2, s2, 0x00020002u)
                PAIR_STEP(hp + 3 * NT, p + 3, s3, 0x00030003u)
            }
            for (; p < p1; ++p, pp += 0x00010001u, hp += NT) {
                const uint32_t s0 = *qp;
                qp += (p & 1) ? (NT * 2 - 1) : 1;
                PAIR_STEP(hp, p, s0, 0u)
            }
            if (p1 > p0) { PAIR_STEP_EDGE(hp, p1, (uint32_t)*qp, false, hi_out) }
            h1 = hi_out ? (int)(hprev & 0xffffu) : (int)(hprev >> 16);
            // row maximum; among equal maxima the last column wins (src/ksw.c:928)
            int sL, cL, sH, cH;
            if (KEYED) {
                const uint32_t kl = m2 & 0xffffu, kh = m2 >> 16;
                sL = (int)(kl >> 6); cL = (int)(kl & 63u) * 2; sH = (int)(kh >> 6); cH = (int)(kh & 63u) * 2 + 1;
            } else {
                sL = (int)(m2 & 0xffffu); cL = pjL * 2; sH = (int)(m2 >> 16); cH = pjH * 2 + 1;
            }
            const bool hi_wins = sH > sL || (sH == sL && cH > cL);
            m = hi_wins ? sH : sL; mj = hi_wins ? cH : cL;
        }
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
        int j = beg;
        while (j < end && H16(j) == 0 && E16(j) == 0) ++j;
        beg = j;
        j = end;
        while (j >= beg && H16(j) == 0 && E16(j) == 0) --j;
        end = j + 2 < qlen ? j + 2 : qlen;
    }
#undef H16
#undef E16
    r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = best_ie + 1; r.gscore = gscore; r.max_off = max_off;
}
