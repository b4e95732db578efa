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
// Instruction budget of a column pair.  Measured on the B200 (bwa_b200_measure_int_alu): the ALU pipe (PRMT, VIADDMNMX,
// VIMNMX3) and the FMA pipe (IMAD) each issue 0.5 warp instructions per clock per SMSP, both together 0.68, IMAD.HI
// half of IMAD -- so the kernel is bound by the instructions it issues, whatever their pipe:
//   ALU: PRMT (score lookup), VIADDMNMX (M), VIADDMNMX.RELU (gap open, once when the insertion and deletion
//        penalties are equal), 2 x VIADDMNMX + PRMT (F chain and its two halves), VIMNMX3 (h), VIADDMNMX (E),
//        PRMT (H(i, j-1) realigned for the stored diagonal), half a VIMNMX3.U16x2 (row maximum of two pairs)  = 9.5
//   FMA: H * 32 (zero test of the diagonal), t << 16, key = h * 64 + index
//   LSU: LDS.64 {H, E}, LDS.U16 selectors, STS.64
//
// Per-lane state (lane stride NT elements, conflict-free), SLOT = pair index, or pair index modulo the ring size:
//   HE[s] = uint2 { H(i-1, 2p-1) | H(i-1, 2p) << 16 ,  E(i, 2p) | E(i, 2p+1) << 16 }
//           i.e. the reference's eh[j] = {H(i-1,j-1), E(i,j)} for j = 2p (low halves) and 2p+1 (high halves)
//   QS[s] = the two PRMT selector bytes of the pair, one per column: code * 17 + 0x80
//           (low nibble picks the score byte, high nibble replicates its sign into the upper byte)
//
// Band-sized state (RING): with a band, row i only touches columns [i - w, i + w + 1] (src/ksw.c:902-907, and
// eh[end] at :940), and beg never decreases, so a column left of i - w is never read again.  The state is then a
// ring of R >= w + 2 pairs instead of qlen / 2 + 1.  A column that enters the band for the first time must read what
// the reference's untouched eh[] holds there -- the first-row initialisation (src/ksw.c:880-883) -- so the pair that
// enters at row i is (re)initialised in closed form before the row is evaluated; columns that were inside, fell
// out of [beg, end) and come back keep their stale values, because they never left the ring.
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
// a * b + c as IMAD (the FMA pipe; callers pass multipliers the compiler cannot see through)
__device__ __forceinline__ uint32_t b200_mad(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t b200_umulhi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
#else
#define B200_DEV static inline
struct uint2 { uint32_t x, y; };
static inline uint32_t b200_umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
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
    int32_t  ring;               // pairs in the state ring; 0 = one slot per pair of the longest query (no ring)
    uint32_t ring_magic;         // ceil(2^32 / ring): p / ring == umulhi(p, ring_magic) whenever p * ring < 2^32
};

constexpr int PAIR_MAX_SCORE = 1023;   // H * 32 must stay below 2^15 (zero test of the diagonal, see PAIR_CORE)
constexpr int PAIR_MAX_Q = 512;        // longest query of the class
constexpr int PAIR_WIDE_MAX_SCORE = 32700;   // WIDE: scores as s16 with the zero test min(H + s, min(H, 1023) * 32), exact below 2^15 - 32
constexpr int PAIR_WIDE_MAX_Q = 65535;       // WIDE: the row maximum is a 32-bit key per column, score << 16 | column
constexpr int PAIR_CHUNK = 64;         // (score, pair index within the chunk) is one 16-bit key: score < 2^10, index < 2^6

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
    S->ring = 0; S->ring_magic = 0;
    return ok;
}
static inline uint32_t pair_ring_magic(int ring) { return ring > 1 ? (uint32_t)(((1ull << 32) + (uint64_t)ring - 1) / (uint64_t)ring) : 0u; }
static inline bool pair_same_gap(const bwa_b200_ext_params_t *p) { return p->o_del == p->o_ins && p->e_del == p->e_ins; }
// state slots per lane for queries up to max_q: the ring when the band makes it smaller
static inline int pair_slots(const bwa_b200_ext_params_t *p, int max_q, PairParams *S)
{
    const int full = max_q / 2 + 1;
    int ring = 0;
    if (p->use_band && p->w >= 0 && p->w + 2 < full) ring = p->w + 2;
    if (S) { S->ring = ring; S->ring_magic = pair_ring_magic(ring); }
    return ring ? ring : full;
}

// ---- jobs answered in closed form -----------------------------------------------------------------------------------------------
// What is left of a read beside a maximal exact match usually differs from the reference in the base that ended the match and in
// little else.  When the query equals the head of its target except for a FEW substituted bases r_1 < ... < r_k (k <= CF_KMAX, no
// base outside A/C/G/T, target at least as long as the query), ksw_extend2's six outputs follow from h0, qlen and the r_m, without
// a matrix -- provided the main diagonal is the strict maximum of every row and of every column.  With match a, mismatch -b,
// g = min(o_del + e_del, o_ins + e_ins) > a + b and D(i) = h0 + (i + 1) a - (a + b) * #{differing bases at or before i}:
//   * H(i, i) >= D(i) as long as D stays positive (h0 > k b), so `M = M ? M + s : 0` (src/ksw.c:924) never cuts the diagonal; the
//     cell after a non-zero cell is always inside [beg, end) (src/ksw.c:959-965) and, with w >= 0, inside the band.  Every value the
//     recurrence produces is the score of an alignment path from the origin (or 0): gaps opening from M only (:926-938), the first
//     row's and column's gap from the origin (:880-883, :905) and cells that fell out of [beg, end) only remove paths.
//   * Measure a path to (i, j) against D(i) (it enters at most i + 1 rows by a diagonal step) or against D(j) (at most j + 1 columns):
//     it gains (a + b) for every difference whose diagonal step (r_m - 1, r_m - 1) -> (r_m, r_m) it does not take, loses (a + b) for
//     every mismatch it meets elsewhere, and loses its gaps, each at least g.  Cut it into its stretches on the main diagonal (gain 0)
//     and its EXCURSIONS off it.  An excursion avoids a run r_m .. r_m' of c neighbouring differences and is off the diagonal in every
//     row between them; its first gap opens at or before row r_m.  A gap of more than dmax_k bases (dmax_k: the longest that costs no
//     more than k (a + b)) costs more than all k differences give back, so inside an excursion worth considering every gap and every
//     shift |j - i| is at most dmax_k.
//   * THE TEST: between each two neighbouring differences, on the rows lo = r_m + dmax_k + 1 <= c < hi = r_m+1 (there must be one),
//     every diagonal shifted by s, 1 <= |s| <= dmax_k, holds a mismatch q[c + s] != t[c] or leaves the matrix.  An excursion that
//     avoids both r_m and r_m+1 either walks one such diagonal through all of [lo, hi) and meets that mismatch, or takes a gap step
//     on those rows -- a gap that opened after row r_m (reaching back to r_m would take more than dmax_k bases) and before r_m+1, so
//     it is a gap of its own, neither the excursion's first one nor the one counted for another pair of neighbours.  Either way the
//     c - 1 pairs of neighbours inside the excursion cost it (a + b) each at least, its first gap costs g, and the excursion ends
//     (a + b) c - (a + b)(c - 1) - g < 0 below the diagonal.  A path with an excursion is therefore strictly below D(i) (and D(j)):
//     H(i, i) = D(i), every other cell of row i and of column i is smaller.  k <= 1 needs no test; random sequence passes it within a
//     base or two per diagonal, tandem repeats and differences closer than dmax_k + 2 do not, and go to the kernels.
//   * The running maximum starts at h0 (src/ksw.c:896) and moves only on m > max (:947), always with mj = i, so max_off = 0 and
//     (max_i, max_j) is the first of D's peaks -- the row before each difference and the last row -- that holds the largest value, if it
//     exceeds h0; gscore = D(qlen - 1) at row qlen - 1.  The z-drop test (:950-957) sees max - m <= k b with equal row and column
//     distance, so it cannot fire when zdrop <= 0 or k b <= zdrop.  Rows after the last query row score below D of their column, hence
//     below max and, in the last column, below D(qlen - 1).
//   * With a band, the columns beyond i + w still hold the first row's values (src/ksw.c:880-883) when row i reaches them: a gap of
//     more than dmax_k bases, out of the running for w > dmax_k + 1.
// Every step is checked against the oracle on jobs built around the conditions (tests/test_ext_pair_host.py).
constexpr int CF_KMAX = 6, CF_DMAX_CAP = 24;
struct ClosedParams { int32_t ok, a, b, zdrop, kcap; int32_t dmax[CF_KMAX + 1]; };     // dmax[k]: the longest gap that costs no more than k mismatches; kcap: most differences taken
static inline ClosedParams closed_params_from(const bwa_b200_ext_params_t *p)
{
    ClosedParams C{};
    const int a = p->mat[0], b = -p->mat[1];
    if (a < 1 || b < 1) return C;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (p->mat[i * 5 + j] != (i == j ? a : -b)) return C;
    const int oe_del = p->o_del + p->e_del, oe_ins = p->o_ins + p->e_ins, g = oe_del < oe_ins ? oe_del : oe_ins;
    if (p->e_del < 1 || p->e_ins < 1 || p->o_del < 0 || p->o_ins < 0 || g <= a + b) return C;
    auto dmax_of = [&](int k) {
        const int d_del = (k * (a + b) - p->o_del) / p->e_del, d_ins = (k * (a + b) - p->o_ins) / p->e_ins;
        const int d = d_del > d_ins ? d_del : d_ins;
        return d < 0 ? 0 : d;
    };
    if (dmax_of(2) > 16) return C;
    if (p->use_band && p->w < dmax_of(2) + 2) return C;
    C.ok = 1; C.a = a; C.b = b; C.zdrop = p->zdrop; C.kcap = 2;
    for (int k = 2; k <= CF_KMAX; ++k) {       // more differences as long as the gaps they pay for stay short and inside the band
        const int d = dmax_of(k);
        if (k > 2 && (d > CF_DMAX_CAP || (p->use_band && p->w < d + 2))) break;
        C.dmax[k] = d; C.kcap = k;
    }
    return C;
}
// The test in three steps, so that key_kernel can spread the second one over a warp: the shape (the differing bases and the cheap
// conditions), one shifted diagonal of one stretch, the six outputs.
struct ClosedShape { int32_t k; int32_t rr[CF_KMAX]; };
// 0: not a closed-form job; 1: it is, nothing left to test; 2: it is if every shifted diagonal of every stretch is broken
template <bool BYTES>
B200_DEV int closed_form_shape(const ClosedParams &C, const JobView &J, uint32_t a, ClosedShape &S)
{
    const uint32_t ql = J.qlen[a], tl = J.tlen[a];
    const int h0 = (int)J.h0[a];
    if (!C.ok || ql == 0 || tl < ql || ql > 0x00ffffffu || h0 > 0x00ffffff) return 0;
    // the differing bases: at most kcap, every compared base in A/C/G/T
    const int kcap = C.kcap;
    int k = 0;
    for (int m = 0; m < CF_KMAX; ++m) S.rr[m] = 0;
    if (BYTES) {
        const uint8_t *qb = J.qb + J.qoff[a], *tb = J.tb + J.toff[a];
        for (uint32_t i = 0; i < ql; ++i) {
            if (qb[i] > 3 || tb[i] > 3) return 0;
            if (qb[i] != tb[i]) { if (k == kcap) return 0; S.rr[k++] = (int)i; }
        }
    } else {
        const uint32_t *qp = J.qp + (J.qoff[a] >> 3), *tp = J.tp + (J.toff[a] >> 3);
        const uint32_t nw = (ql + 7) >> 3;
        for (uint32_t w = 0; w < nw; ++w) {
            const uint32_t qw = qp[w], tw = tp[w], rem = ql - 8 * w;
            const uint32_t m = rem < 8 ? ~(0xffffffffu >> (4 * rem)) : 0xffffffffu;
            if ((qw | tw) & 0xccccccccu & m) return 0;
            uint32_t x = (qw ^ tw) & m;
            for (int n = 0; x; ++n, x <<= 4)
                if (x >> 28) { if (k == kcap) return 0; S.rr[k++] = (int)(8 * w) + n; }
        }
    }
    S.k = k;
    if (h0 <= k * C.b || (C.zdrop > 0 && k * C.b > C.zdrop)) return 0;
    const int dmax = C.dmax[k];
    if (k < 2 || dmax <= 0) return 1;
    for (int m = 0; m + 1 < k; ++m)
        if (S.rr[m] + dmax + 1 >= S.rr[m + 1]) return 0;      // no row left between two neighbouring differences to break a diagonal on
    return 2;
}
// some row c of [lo, hi) has q[c + s] != t[c] or lies outside the matrix (c + s >= lo - dmax > 0)
template <bool BYTES>
B200_DEV bool closed_form_diag_broken(const JobView &J, uint32_t a, int ql, int lo, int hi, int s)
{
    if (hi - 1 + s >= ql) return true;
    if (BYTES) {
        const uint8_t *qb = J.qb + J.qoff[a], *tb = J.tb + J.toff[a];
        for (int c = lo; c < hi; ++c)
            if (qb[c + s] != tb[c]) return true;
    } else {
        const uint32_t *qp = J.qp + (J.qoff[a] >> 3), *tp = J.tp + (J.toff[a] >> 3);
        for (int c = lo; c < hi; ++c) {
            const int j = c + s;
            if (((qp[j >> 3] >> (28 - 4 * (j & 7))) & 15u) != ((tp[c >> 3] >> (28 - 4 * (c & 7))) & 15u)) return true;
        }
    }
    return false;
}
// check number x of a job with k differences (0 <= x < closed_form_checks): stretch x / (2 dmax), shift -1, +1, -2, +2 ...
B200_DEV int closed_form_checks(const ClosedParams &C, int k) { return (k - 1) * 2 * C.dmax[k]; }
template <bool BYTES>
B200_DEV bool closed_form_check(const ClosedParams &C, const JobView &J, uint32_t a, int ql, int k, const int32_t *rr, int x)
{
    const int dmax = C.dmax[k], m = x / (2 * dmax), y = x - m * 2 * dmax, d = (y >> 1) + 1;
    return closed_form_diag_broken<BYTES>(J, a, ql, rr[m] + dmax + 1, rr[m + 1], (y & 1) ? d : -d);
}
B200_DEV void closed_form_result(const ClosedParams &C, int h0, int ql, const ClosedShape &S, bwa_b200_ext_result_t *r)
{
    // D at its peaks, in row order; the first one holding the largest value is where the maximum was last raised
    const int ab = C.a + C.b, last = ql - 1, k = S.k;
    int best = h0, row = -1;
    for (int m = 0; m < k; ++m) {
        if (S.rr[m] < 1) continue;
        const int v = h0 + S.rr[m] * C.a - m * ab;               // D(rr[m] - 1): m differences before it
        if (v > best) { best = v; row = S.rr[m] - 1; }
    }
    const int g = h0 + ql * C.a - k * ab;
    if (g > best) { best = g; row = last; }
    r->score = best; r->qle = row + 1; r->tle = row + 1;
    r->gscore = g; r->gtle = (int32_t)ql; r->max_off = 0;
}
template <bool BYTES>
B200_DEV bool closed_form_job(const ClosedParams &C, const JobView &J, uint32_t a, bwa_b200_ext_result_t *r)
{
    ClosedShape S;
    const int st = closed_form_shape<BYTES>(C, J, a, S);
    if (st == 0) return false;
    const int ql = (int)J.qlen[a];
    if (st == 2)
        for (int x = 0, n = closed_form_checks(C, S.k); x < n; ++x)
            if (!closed_form_check<BYTES>(C, J, a, ql, S.k, S.rr, x)) return false;
    closed_form_result(C, (int)J.h0[a], ql, S, r);
    return true;
}

// One column pair (2p, 2p+1) of row i.  SEL: low 16 bits = the two PRMT selector bytes of the pair.
//   Hd   = {H(i-1,2p-1), H(i-1,2p)}                     the diagonal of both columns
//   M    = min(Hd + score, Hd * 32)                      `M = M ? M + s : 0` (src/ksw.c:924): Hd == 0 gives M <= 0,
//                                                        which every later use treats like 0; Hd > 0 gives Hd + score
//   c    : chain register, HIGH half = F(i,2p); Fh.hi = F(i,2p+1); new c.hi = F(i,2p+2)
//   max(x - pen, 0) is the RELU form of VIADDMNMX with the (negative) penalty as its own third operand
#define PAIR_CORE(HE, SEL)                                                                       \
        const uint32_t S_ = b200_prmt(tlo, tab_n, (SEL));                                        \
        const uint32_t M_ = __viaddmin_s16x2((HE).x, S_, (WIDE ? __vmins2((HE).x, 0x03ff03ffu) : (HE).x) * 32u); \
        const uint32_t t2_ = __viaddmax_s16x2_relu(M_, noe_ins2, noe_ins2);                      \
        const uint32_t t1_ = SAME_GAP ? t2_ : __viaddmax_s16x2_relu(M_, noe_del2, noe_del2);     \
        const uint32_t Fh_ = __viaddmax_s16x2(c, ne_ins2, t2_ << 16);                            \
        const uint32_t F_ = __byte_perm(c, Fh_, 0x7632);                                         \
        c = __viaddmax_s16x2(Fh_, ne_ins2, t2_);                                                 \
        const uint32_t h_ = __vimax3_s16x2(M_, (HE).y, F_);                                      \
        const uint32_t En_ = __viaddmax_s16x2((HE).y, ne_del2, t1_);

// a pair wholly inside the window: HEV = its loaded state, KEY receives h * 64 + PPK (PPK = the pair's index within the
// (WIDE: the larger of the two 32-bit keys score << 16 | column, PPK = the low column's offset, a literal, or the column itself)
// chunk, in both halves -- or a literal offset, added to the index later); the store leaves {H(i, 2p-1) | H(i, 2p) << 16, E(i+1, .)} = what the next row reads
#define PAIR_STEP(HP, HEV, SEL, PPK, KEY)                                                        \
    {                                                                                            \
        PAIR_CORE(HEV, SEL)                                                                      \
        if (WIDE) KEY = max(__byte_perm((PPK), h_, 0x5410), __byte_perm((PPK), h_, 0x7610) + 1u); \
        else KEY = h_ * k64 + (PPK);                                                             \
        uint2 o_;                                                                                \
        o_.y = En_;                                                                              \
        o_.x = __byte_perm(hprev, h_, 0x5432);                                                   \
        *(HP) = o_;                                                                              \
        hprev = h_;                                                                              \
    }

// the first / last pair of the window.  LO_OUT: the low column (beg - 1) lies outside: its inputs are zeroed, which
// makes its H, F and gap-open terms 0 = the reference's initial h1 and f of the row, and its stored state is kept.
// HI_OUT: the high column (end) lies outside: it is computed and discarded, except that eh[end] = {H(i,end-1), 0}
// (src/ksw.c:940) is exactly what the pair store leaves there.  OUT receives what was stored.
#define PAIR_STEP_EDGE(HP, SEL, PPK, LO_OUT, HI_OUT, OUT)                                        \
    {                                                                                            \
        const uint2 old_ = *(HP);                                                                \
        const uint32_t inm_ = (LO_OUT) ? 0xffff0000u : 0xffffffffu;                              \
        const uint32_t outm_ = (HI_OUT) ? (inm_ & 0x0000ffffu) : inm_;                           \
        uint2 he_;                                                                               \
        he_.x = old_.x & inm_; he_.y = old_.y & inm_;                                            \
        if (LO_OUT) hprev = old_.x << 16;                                                        \
        PAIR_CORE(he_, SEL)                                                                      \
        const uint32_t hk_ = (HI_OUT) ? (h_ & 0x0000ffffu) : h_;                                 \
        if (WIDE) m2 = max(m2, max(__byte_perm((PPK), hk_, 0x5410), __byte_perm((PPK), hk_, 0x7610) + 1u)); \
        else m2 = __vmaxu2(m2, hk_ * 64u + (PPK));                                               \
        (OUT).y = (En_ & outm_) | (old_.y & ~inm_);                                              \
        (OUT).x = __byte_perm(hprev, h_, 0x5432);                                                \
        *(HP) = (OUT);                                                                           \
        hprev = h_;                                                                              \
    }

// One job.  HEp / QSp are this lane's element 0 of the [slot][lane] arrays; tab = S.tab in memory that can be indexed.
// RING: the state is a ring of S.ring pairs (see the header); CHUNKED: a row may hold more than PAIR_CHUNK pairs, so the row
// maximum is folded chunk by chunk; U: pairs per trip of the longest unrolled loop (4 or 8); WIDE: scores up to PAIR_WIDE_MAX_SCORE and
// queries up to PAIR_WIDE_MAX_Q (the row maximum as 32-bit keys; needs RING or a state array for the whole query, never CHUNKED).
template <bool BYTES, int NT, bool SAME_GAP, bool RING, bool CHUNKED, int U, bool WIDE = false>
B200_DEV void pair_job(const ExtParams &P, const PairParams &S, const uint32_t *const tab, const JobView &J, uint32_t a, int qlen, int tlen, int h0,
                       uint2 *const HEp, uint16_t *const QSp, bwa_b200_ext_result_t &r, unsigned long long &my_cells)
{
    const uint32_t qo = J.qoff[a], to = J.toff[a];
    const int oe_ins = P.o_ins + P.e_ins;
    const int R = RING ? S.ring : 0;
    const uint32_t rmagic = S.ring_magic;
    uint16_t *const hw = reinterpret_cast<uint16_t *>(HEp);
#define SLOT(p) (RING ? (int)((uint32_t)(p) - b200_umulhi((uint32_t)(p), rmagic) * (uint32_t)R) : (int)(p))
#define H16(j) hw[SLOT((j) >> 1) * (NT * 4) + ((j) & 1)]
#define E16(j) hw[SLOT((j) >> 1) * (NT * 4) + 2 + ((j) & 1)]
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
    // Staging of pair p: the query as two PRMT selector bytes (columns >= qlen: N), and the first row of eh[]:
    // H(-1,-1) = h0, then one gap open, then extensions, clipped at 0 (src/ksw.c:880-883); E = 0
    uint32_t qword = 0;
    int qword_g = -1;                                       // packed query word held in qword
    auto stage = [&](int p) {
        uint32_t c0 = 4u, c1 = 4u;
        const int j0 = 2 * p;
        if (BYTES) {
            if (j0 < qlen) c0 = J.qb[qo + j0];
            if (j0 + 1 < qlen) c1 = J.qb[qo + j0 + 1];
        } else if (j0 < qlen) {
            if ((j0 >> 3) != qword_g) { qword_g = j0 >> 3; qword = J.qp[(qo >> 3) + qword_g]; }
            c0 = (qword >> (28 - 4 * (j0 & 7))) & 15u;
            if (j0 + 1 < qlen) c1 = (qword >> (24 - 4 * (j0 & 7))) & 15u;
        }
        c0 = c0 > 4u ? 4u : c0; c1 = c1 > 4u ? 4u : c1;
        const int s = SLOT(p);
        QSp[s * NT] = (uint16_t)((c0 * 17u + 0x80u) | (c1 * 17u + 0x80u) << 8);
        int v0 = j0 == 0 ? h0 : h0 - oe_ins - (j0 - 1) * P.e_ins;
        int v1 = h0 - oe_ins - j0 * P.e_ins;
        v0 = v0 > 0 ? v0 : 0; v1 = v1 > 0 ? v1 : 0;
        uint2 o;
        o.x = (uint32_t)v0 | (uint32_t)v1 << 16;
        o.y = 0u;
        HEp[s * NT] = o;
    };
    const int p_last = qlen >> 1;                           // pair of column qlen (eh[qlen] is written, never evaluated)
    int pinit = 0;                                          // pairs [0, pinit) are staged
    {
        const int first = RING ? ((w + 1) >> 1) : p_last;   // row 0 reaches column min(qlen, w + 1)
        const int upto = first < p_last ? first : p_last;
        for (; pinit <= upto; ++pinit) stage(pinit);
    }
    int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
    int beg = 0, end = qlen;
    uint32_t tword = 0, cells = 0;
    // per-batch constants, made to depend on a run-time zero so that they stay in registers (ptxas otherwise
    // re-reads the kernel parameter bank inside the column loop) and so that IMAD stays IMAD
    const uint32_t rz = (uint32_t)tlen >> 31;
    const uint32_t tab_n = S.tab_n + rz, noe_del2 = S.noe_del2 + rz, ne_del2 = S.ne_del2 + rz, noe_ins2 = S.noe_ins2 + rz, ne_ins2 = S.ne_ins2 + rz;
    const uint32_t k64 = 64u + rz;
    (void)noe_del2; (void)k64;
    constexpr uint32_t PPI = WIDE ? 2u : 0x00010001u;         // what one pair adds to pp: its low column (WIDE) / its index in both halves
#define KMAX3(a, b, c) (WIDE ? __vimax3_u32((a), (b), (c)) : __vimax3_u16x2((a), (b), (c)))
#define KMAX2(a, b) (WIDE ? max((a), (b)) : __vmaxu2((a), (b)))
#define KMERGE(k, base, acc) (WIDE ? max((acc), (k) + (base)) : __viaddmax_u16x2((k), (base), (acc)))
    int h1_edge = h0 - P.o_del;                             // h0 - (o_del + e_del * (i + 1)), the first column of row i while beg == 0
    for (int i = 0; i < tlen; ++i) {
        int tbv;
        if (BYTES) tbv = J.tb[to + i];
        else {
            if ((i & 7) == 0) tword = J.tp[(to + i) >> 3];
            tbv = (int)(tword >> 28);
            tword <<= 4;
        }
        tbv = tbv > 4 ? 4 : tbv;
        if (P.use_band) {
            if (beg < i - w) beg = i - w;
            if (end > i + w + 1) end = i + w + 1;
            if (end > qlen) end = qlen;
        }
        if (RING) {                                         // the pair that enters the band at this row (see the header)
            int need = (i + w + 1) >> 1;
            need = need < p_last ? need : p_last;
            for (; pinit <= need; ++pinit) stage(pinit);
        }
        const uint32_t tlo = tab[tbv];                    // row of the matrix (shared memory on the device: one broadcast load)
        h1_edge -= P.e_del;
        int h1 = 0, m = 0, mj = -1;
        if (beg == 0) h1 = h1_edge < 0 ? 0 : h1_edge;
        bool beg_live = false;                              // eh[beg] is known to be non-zero after the row
        if (beg < end) {
            const int p0 = beg >> 1, p1 = (end - 1) >> 1;          // first and last column pair of the window
            const bool lo_out = (beg & 1) != 0, hi_out = (end & 1) != 0;
            uint32_t hprev = (uint32_t)h1 << 16, c = 0u, m2 = 0u;   // c.hi = F(i, beg) = 0
            const int s0 = SLOT(p0);
            uint2 *hp = HEp + s0 * NT;
            const uint16_t *qp = QSp + s0 * NT;
            int pw = RING ? p0 - s0 + R : 0x7fffffff;               // first pair at or after p0 that sits in slot 0
            int pbase = p0, climit = CHUNKED ? p0 + PAIR_CHUNK : 0x7fffffff;     // the chunk of the row maximum's keys
            uint32_t pp = WIDE ? 2u * (uint32_t)p0 : 0u;            // (p - pbase) in both halves; WIDE: the pair's low column
            auto fold = [&]() {                                     // (m, mj) <- chunk maximum; the later column wins ties (src/ksw.c:928)
                if (WIDE) { m = (int)(m2 >> 16); mj = (int)(m2 & 0xffffu); return; }
                const uint32_t kl = m2 & 0xffffu, kh = m2 >> 16;
                const int sL = (int)(kl >> 6), cL = (pbase + (int)(kl & 63u)) * 2, sH = (int)(kh >> 6), cH = (pbase + (int)(kh & 63u)) * 2 + 1;
                const bool hi_wins = sH > sL || (sH == sL && cH > cL);
                const int cs = hi_wins ? sH : sL, cc = hi_wins ? cH : cL;
                if (!CHUNKED || cs >= m) { m = cs; mj = cc; }
            };
            uint2 o0;
            PAIR_STEP_EDGE(hp, (uint32_t)*qp, pp, lo_out, hi_out && p0 == p1, o0)
            beg_live = lo_out ? ((o0.x | o0.y) >> 16) != 0u : ((o0.x | o0.y) & 0xffffu) != 0u;
            int p = p0 + 1;
            pp += PPI; hp += NT; qp += NT;
            if (RING && p == pw) { hp -= R * NT; qp -= R * NT; pw += R; }
            while (p < p1) {
                int stop = p1;
                if (RING) stop = stop < pw ? stop : pw;
                if (CHUNKED) stop = stop < climit ? stop : climit;
                // unrolled trips of 8, then 4 pairs: keys carry the pair's offset within the trip as a literal; the trip's maximum
                // gets the trip's index (pp) added inside the VIADDMNMX that merges it into the row maximum.  Eight pairs per trip
                // measured 12 % faster than four on the bins whose shared-memory state leaves ten warps per SM (more independent
                // work between dependent instructions); sixteen measured slower again.
                if (U == 8 && p + 8 <= stop) {
                    const int n8 = (stop - p) >> 3;
                    uint2 *const hp_end = hp + n8 * (8 * NT);
                    p += n8 * 8;
                    do {
                        const uint2 e0 = hp[0], e1 = hp[NT], e2 = hp[2 * NT], e3 = hp[3 * NT], e4 = hp[4 * NT], e5 = hp[5 * NT], e6 = hp[6 * NT], e7 = hp[7 * NT];
                        const uint32_t s0_ = qp[0], s1_ = qp[NT], s2_ = qp[2 * NT], s3_ = qp[3 * NT], s4_ = qp[4 * NT], s5_ = qp[5 * NT], s6_ = qp[6 * NT], s7_ = qp[7 * NT];
                        uint32_t k0, k1, k2, k3, k4, k5, k6, k7;
                        PAIR_STEP(hp, e0, s0_, 0u, k0)
                        PAIR_STEP(hp + NT, e1, s1_, PPI, k1)
                        PAIR_STEP(hp + 2 * NT, e2, s2_, 2u * PPI, k2)
                        PAIR_STEP(hp + 3 * NT, e3, s3_, 3u * PPI, k3)
                        PAIR_STEP(hp + 4 * NT, e4, s4_, 4u * PPI, k4)
                        PAIR_STEP(hp + 5 * NT, e5, s5_, 5u * PPI, k5)
                        PAIR_STEP(hp + 6 * NT, e6, s6_, 6u * PPI, k6)
                        PAIR_STEP(hp + 7 * NT, e7, s7_, 7u * PPI, k7)
                        k0 = KMAX3(k0, k1, k2);
                        k3 = KMAX3(k3, k4, k5);
                        k0 = KMAX3(k0, k3, k6);
                        k0 = KMAX2(k0, k7);
                        m2 = KMERGE(k0, pp, m2);
                        pp += 8u * PPI; hp += 8 * NT; qp += 8 * NT;
                    } while (hp != hp_end);
                }
                if (p + 4 <= stop) {
                    const int n4 = (stop - p) >> 2;
                    uint2 *const hp_end = hp + n4 * (4 * NT);
                    p += n4 * 4;
                    do {
                        const uint2 e0 = hp[0], e1 = hp[NT], e2 = hp[2 * NT], e3 = hp[3 * NT];
                        const uint32_t s0_ = qp[0], s1_ = qp[NT], s2_ = qp[2 * NT], s3_ = qp[3 * NT];
                        uint32_t k0, k1, k2, k3;
                        PAIR_STEP(hp, e0, s0_, 0u, k0)
                        PAIR_STEP(hp + NT, e1, s1_, PPI, k1)
                        PAIR_STEP(hp + 2 * NT, e2, s2_, 2u * PPI, k2)
                        PAIR_STEP(hp + 3 * NT, e3, s3_, 3u * PPI, k3)
                        k0 = KMAX3(k0, k1, k2);
                        k0 = KMAX2(k0, k3);
                        m2 = KMERGE(k0, pp, m2);
                        pp += 4u * PPI; hp += 4 * NT; qp += 4 * NT;
                    } while (hp != hp_end);
                }
                for (; p < stop; ++p, pp += PPI, hp += NT, qp += NT) {
                    const uint2 e0 = hp[0];
                    uint32_t k0;
                    PAIR_STEP(hp, e0, (uint32_t)*qp, pp, k0)
                    m2 = KMAX2(m2, k0);
                }
                if (RING && p == pw) { hp -= R * NT; qp -= R * NT; pw += R; }
                if (CHUNKED && p == climit) { fold(); m2 = 0u; pp = 0u; pbase = p; climit += PAIR_CHUNK; }
            }
            if (p1 > p0) { uint2 o1; PAIR_STEP_EDGE(hp, (uint32_t)*qp, pp, false, hi_out, o1) (void)o1; }
            h1 = hi_out ? (int)(hprev & 0xffffu) : (int)(hprev >> 16);
            fold();
            cells += (uint32_t)(end - beg);
        }
        // ---- row i is complete (src/ksw.c:940-959)
        H16(end) = (uint16_t)h1; E16(end) = 0;           // eh[end] = {h1, 0}
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
        // window of the next row (src/ksw.c:965-970); the common answers are already in registers: eh[beg] as stored by
        // the first pair, and eh[end] = {h1, 0}
        if (!beg_live) {
            int j = beg;
            while (j < end && H16(j) == 0 && E16(j) == 0) ++j;
            beg = j;
        }
        int j = end;
        if (h1 == 0) {
            --j;
            while (j >= beg && H16(j) == 0 && E16(j) == 0) --j;
        }
        end = j + 2 < qlen ? j + 2 : qlen;
    }
#undef H16
#undef E16
#undef SLOT
#undef KMAX3
#undef KMAX2
#undef KMERGE
    my_cells += cells;
    r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = best_ie + 1; r.gscore = gscore; r.max_off = max_off;
}
