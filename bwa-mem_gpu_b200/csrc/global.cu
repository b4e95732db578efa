// global.cu -- banded global alignment with backtrack -> CIGAR, score and NM, sm_100a  (SURVEY 8f row 4).
//
// Parity target: the reference's CPU path  mem_reg2aln -> bwa_gen_cigar2 -> ksw_global2  (src/bwa.c:111-216,
// src/ksw.c:1120-1241 = bwa_index/ksw.c:504-606): the remaining dynamic programming of the reference's worker2.
//
//   global_kernel<R, BLOCK>   one job per lane, 32 jobs of similar band width per warp, rows in order (the recurrence of
//                             ksw_global2 cell for cell, int32).  Per-column state {H(i-1,j-1), E(i,j)} lives in a ring of R
//                             slots in shared memory, [slot][lane] (a row touches columns [i-w, i+w+1), so 2w+2 slots are
//                             live).  The backtrack matrix keeps 4 bits per cell -- the three direction fields of the
//                             reference's byte (h: 2 bits, e: 1, f: 1) -- eight cells per word, written [row][word][lane] so
//                             that the 32 lanes of a warp store one 128-byte line; warps are persistent and reuse their
//                             slab, which therefore stays in L2 for the backtrack that follows immediately.
//                             The backtrack walks (i, k) exactly as the reference, merges operations like push_cigar, counts
//                             mismatches for NM on the way, and writes the operations right-aligned into the job's row.
//   compact_kernel            rows -> flat CIGAR array at scanned offsets.
//
// Bound: INT ALU (about 20 integer operations per cell plus one 64-bit shared-memory load and store); HBM traffic is the
// sequences once and 0.5 byte per cell of backtrack state that mostly stays in L2.
#include "internal.h"
#include <algorithm>
#include <vector>
#include <cub/cub.cuh>

namespace {

constexpr int32_t MINF = -0x40000000;      // MINUS_INF, src/ksw.c

struct GArgs {
    const uint8_t *qseq, *tseq;
    const uint32_t *qoff, *qlen, *toff, *tlen, *w;
    int8_t mat[25];
    int32_t o_del, e_del, o_ins, e_ins;
    int32_t *score, *nm;
    uint32_t *n_cigar, *rows;               // rows: cig_stride operations per job, right-aligned
    uint32_t cig_stride;
    uint32_t aligned8;                      // every sequence starts on an 8-byte boundary and is padded to a multiple of 8
    unsigned long long *counters;           // [0] cells, [1] widest CIGAR row needed
};

// stage a byte-per-base sequence into shared memory as 4-bit codes, 8 per word, first base in the high nibble, words `stride` apart.
// aligned8: the sequence starts on an 8-byte boundary and is padded to a multiple of 8 (the GASAL host layout) -> one 64-bit load per
// word and SIMD-in-register packing instead of eight byte loads
__device__ __forceinline__ uint32_t pack8(uint2 v)
{
    const uint32_t a = __vminu4(v.x, 0x04040404u), b = __vminu4(v.y, 0x04040404u);     // codes above 4 are read as 4
    const uint32_t ta = (a << 4) | (a >> 8), tb = (b << 4) | (b >> 8);                  // byte 0 = b0 b1, byte 2 = b2 b3 (nibbles)
    return (__byte_perm(ta, 0u, 0x4402u) << 16) | __byte_perm(tb, 0u, 0x4402u);
}
__device__ __forceinline__ void stage_seq(uint32_t *dst, int stride, const uint8_t *src, int len, bool aligned8)
{
    if (aligned8) {
        const uint2 *s8 = reinterpret_cast<const uint2 *>(src);
        for (int j8 = 0; j8 < len; j8 += 8) {
            uint32_t wv = pack8(s8[j8 >> 3]);
            if (j8 + 8 > len) wv = (wv & ~(0xffffffffu >> (4 * (len - j8)))) | (0x44444444u >> (4 * (len - j8)));   // past the end: code 4
            dst[(j8 >> 3) * stride] = wv;
        }
    } else {
        for (int j8 = 0; j8 < len; j8 += 8) {
            uint32_t wv = 0;
            for (int u = 0; u < 8; ++u) { const uint32_t c = j8 + u < len ? src[j8 + u] : 4u; wv |= (c > 4u ? 4u : c) << (28 - 4 * u); }
            dst[(j8 >> 3) * stride] = wv;
        }
    }
}

template <int R, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
global_kernel(GArgs a, const uint32_t *__restrict__ perm, uint32_t n, uint32_t tlen_max, uint32_t qw_max, uint32_t tw_max, uint32_t *__restrict__ z_all)
{
    // dynamic shared memory: eh ring [R][BLOCK] of int2, then the staged sequences as 4-bit codes, 8 per word, first base in
    // the high nibble: query [qw_max][BLOCK], target [tw_max][BLOCK]
    extern __shared__ int2 eh_ring[];
    __shared__ uint32_t smat_lo[5], smat_hi[5];            // biased score bytes of matrix row t: lo = query codes 0..3, hi = code 4
    __shared__ int s_bias;
    int2 (*eh)[BLOCK] = reinterpret_cast<int2 (*)[BLOCK]>(eh_ring);
    uint32_t *const qs = reinterpret_cast<uint32_t *>(eh_ring + (size_t)R * BLOCK) + threadIdx.x;       // qs[g * BLOCK]
    uint32_t *const ts = qs + (size_t)qw_max * BLOCK;                                                   // ts[g * BLOCK]
    constexpr int WPR = R / 8;              // backtrack words per row
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    if (tid == 0) {
        int mn = 0;
        for (int i = 0; i < 25; ++i) mn = mn < a.mat[i] ? mn : a.mat[i];
        s_bias = -mn;
        for (int t = 0; t < 5; ++t) {
            uint32_t lo = 0;
            for (int q = 0; q < 4; ++q) lo |= (uint32_t)(uint8_t)(a.mat[t * 5 + q] - mn) << (8 * q);
            smat_lo[t] = lo; smat_hi[t] = (uint32_t)(uint8_t)(a.mat[t * 5 + 4] - mn);
        }
    }
    __syncthreads();
    const int bias = s_bias;
    const uint32_t gwarp = blockIdx.x * (BLOCK / 32) + (tid >> 5), n_warps = gridDim.x * (BLOCK / 32);
    uint32_t *const z = z_all + (uint64_t)gwarp * tlen_max * WPR * 32;
    const int oe_del = a.o_del + a.e_del, oe_ins = a.o_ins + a.e_ins;
    unsigned long long cells = 0;

    for (uint32_t chunk = gwarp; (uint64_t)chunk * 32 < n; chunk += n_warps) {
        const uint32_t idx = chunk * 32 + lane;
        const bool valid = idx < n;
        const uint32_t job = valid ? perm[idx] : 0u;
        const int qlen = valid ? (int)a.qlen[job] : 0, tlen = valid ? (int)a.tlen[job] : 0, w = valid ? (int)a.w[job] : 0;
        const uint8_t *q = a.qseq + (valid ? a.qoff[job] : 0u), *t = a.tseq + (valid ? a.toff[job] : 0u);
        // stage both sequences (codes above 4 are read as 4)
        stage_seq(qs, BLOCK, q, qlen, a.aligned8 != 0);
        stage_seq(ts, BLOCK, t, tlen, a.aligned8 != 0);
        // first row (src/ksw.c:1141-1147); only columns 0 .. min(w + 1, qlen) can be read before they are rewritten
        eh[0][tid] = make_int2(0, MINF);
        {
            const int top = w + 1 < qlen ? w + 1 : qlen;
            for (int j = 1; j <= top; ++j) eh[j & (R - 1)][tid] = j <= w ? make_int2(-(a.o_ins + a.e_ins * j), MINF) : make_int2(MINF, MINF);
        }
        int warp_tl = tlen, warp_w = w;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            warp_tl = max(warp_tl, __shfl_xor_sync(0xffffffffu, warp_tl, o));
            warp_w = max(warp_w, __shfl_xor_sync(0xffffffffu, warp_w, o));
        }
        const int warp_nc = 2 * warp_w + 1;          // no row of any lane has more columns
        uint32_t tword = 0;
        for (int i = 0; i < warp_tl; ++i) {
            const bool act = i < tlen;
            const int beg = i > w ? i - w : 0;
            const int end = i + w + 1 < qlen ? i + w + 1 : qlen;
            const int ncol = act && end > beg ? end - beg : 0;
            if ((i & 7) == 0 && act) tword = ts[(i >> 3) * BLOCK];
            const int tb = act ? (int)((tword >> (28 - 4 * (i & 7))) & 15u) : 4;
            const uint32_t mlo = smat_lo[tb], mhi = smat_hi[tb];
            int32_t f = MINF, h1 = beg == 0 ? -(a.o_del + a.e_del * (i + 1)) : MINF;
            uint32_t zw = 0, qw = ncol ? qs[(beg >> 3) * BLOCK] : 0u;
            uint32_t *zrow = z + (uint64_t)i * WPR * 32 + lane;
#pragma unroll 4
            for (int c = 0; c < warp_nc; ++c) {
                if (c < ncol) {
                    const int j = beg + c;
                    if ((j & 7) == 0 && c) qw = qs[(j >> 3) * BLOCK];
                    const uint32_t qc = (qw >> (28 - 4 * (j & 7))) & 15u;
                    int2 *p = &eh[j & (R - 1)][tid];
                    const int2 pe = *p;
                    // max with the "first operand wins ties" predicate in one DPX instruction (VIMNMX with predicate output)
                    int32_t m = pe.x + (int)(__byte_perm(mlo, mhi, qc) & 0xffu) - bias, e = pe.y, h;
                    bool m_ge_e, h_ge_f, t_ge_e, t_ge_f;
                    uint32_t d;
                    h = __vibmax_s32(m, e, &m_ge_e);
                    d = m_ge_e ? 0u : 1u;
                    h = __vibmax_s32(h, f, &h_ge_f);
                    d = h_ge_f ? d : 2u;
                    e = __vibmax_s32(m - oe_del, e - a.e_del, &t_ge_e);       // E(i+1,j); extension wins only when strictly larger
                    d |= t_ge_e ? 0u : 4u;
                    *p = make_int2(h1, e);
                    h1 = h;
                    f = __vibmax_s32(m - oe_ins, f - a.e_ins, &t_ge_f);
                    d |= t_ge_f ? 0u : 8u;
                    zw |= d << (4 * (c & 7));
                }
                if ((c & 7) == 7) { zrow[(c >> 3) * 32] = zw; zw = 0; }
            }
            if (warp_nc & 7) zrow[(warp_nc >> 3) * 32] = zw;
            if (act) eh[end & (R - 1)][tid] = make_int2(h1, MINF);
            cells += (unsigned long long)ncol;
        }
        if (!valid) continue;
        // score = eh[qlen].h (src/ksw.c:1207).  When the last row does not reach column qlen the reference reads the
        // value of the first-row initialisation, which is MINUS_INF there (qlen > w), or was never in the ring (tlen == 0)
        int32_t score;
        if (tlen == 0) score = qlen == 0 ? 0 : (qlen <= w ? -(a.o_ins + a.e_ins * qlen) : MINF);
        else score = tlen + w >= qlen ? eh[qlen & (R - 1)][tid].x : MINF;
        // backtrack (src/ksw.c:1208-1231)
        int i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        uint32_t which = 0, n_ops = 0, cur_op = 0, cur_len = 0, nmm = 0, gap = 0, edge_del = 0;
        uint32_t *row = a.rows + (uint64_t)job * a.cig_stride;
        auto flush = [&]() {                 // the operation is complete: store it (right-aligned, the list is built backwards)
            if (cur_len == 0) return;
            if (n_ops < a.cig_stride) row[a.cig_stride - 1 - n_ops] = cur_len << 4 | cur_op;
            if (cur_op) gap += cur_len;
            if (cur_op == 2 && n_ops == 0) edge_del += cur_len;        // will be the last operation of the CIGAR
            ++n_ops;
        };
        auto push = [&](uint32_t op, uint32_t len) {                   // push_cigar
            if (cur_len && op == cur_op) cur_len += len;
            else { flush(); cur_op = op; cur_len = len; }
        };
        while (i >= 0 && k >= 0) {
            int c = k - (i > w ? i - w : 0);
            c = c < 0 ? 0 : (c > R - 1 ? R - 1 : c);        // never outside the row's words (the reference would read out of bounds)
            const uint32_t code = (z[((uint64_t)i * WPR + (c >> 3)) * 32 + lane] >> (4 * (c & 7))) & 15u;
            which = which == 0 ? (code & 3u) : (which == 1 ? ((code >> 2) & 1u) : ((code >> 3) & 1u) * 2u);
            if (which == 0) {
                const uint32_t qc = (qs[(k >> 3) * BLOCK] >> (28 - 4 * (k & 7))) & 15u, tc = (ts[(i >> 3) * BLOCK] >> (28 - 4 * (i & 7))) & 15u;
                nmm += qc != tc; push(0, 1); --i; --k;
            }
            else if (which == 1) { push(2, 1); --i; }
            else { push(1, 1); --k; }
        }
        if (i >= 0) push(2, (uint32_t)(i + 1));
        if (k >= 0) push(1, (uint32_t)(k + 1));
        if (cur_len && cur_op == 2 && n_ops > 0) edge_del += cur_len;   // the first operation of the CIGAR (and not also the last)
        flush();
        a.score[job] = score;
        a.n_cigar[job] = n_ops;
        a.nm[job] = (int32_t)(nmm + gap - edge_del);                    // src/bwa.c:178-205
        if (n_ops > a.cig_stride) atomicMax(a.counters + 1, (unsigned long long)n_ops);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if (lane == 0 && cells) atomicAdd(a.counters, cells);
}

// global_band_kernel<W, BLOCK>: the same recurrence for the narrow bands (w <= W, W = 7 or 15 -- almost every job of a read-length
// workload), with the column state in REGISTERS.  In band coordinates c = j - i + w a row's columns are c = 0 .. 2w whatever the row:
// the diagonal neighbour (i-1, j-1) has the same c, the upper one (i-1, j) has c + 1, so {H(i-1, .), E(i, .)} are two register arrays
// updated in place by a fully unrolled loop over c -- no shared-memory ring, no index arithmetic; the 4-bit query codes of the
// row's window sit in a 2 x 64-bit shift register that moves one base per row.  Out-of-range cells (j < 0, j >= qlen) are
// predicated off and leave the arrays as the reference leaves eh[]; the first-column boundary -(o_del + e_del (i+1)) is planted
// where the next row reads its diagonal.  Backtrack state is indexed by c.
template <int W, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
global_band_kernel(GArgs a, const uint32_t *__restrict__ perm, uint32_t n, uint32_t tlen_max, uint32_t qw_max, uint32_t tw_max, uint32_t *__restrict__ z_all)
{
    extern __shared__ int2 eh_ring[];                       // only the staged sequences here: query [qw_max][BLOCK], target [tw_max][BLOCK]
    __shared__ uint32_t smat_lo[5], smat_hi[5];
    __shared__ int s_bias;
    constexpr int NC = 2 * W + 1, WPR = (NC + 7) / 8;
    uint32_t *const qs = reinterpret_cast<uint32_t *>(eh_ring) + threadIdx.x;
    uint32_t *const ts = qs + (size_t)qw_max * BLOCK;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    if (tid == 0) {
        int mn = 0;
        for (int i = 0; i < 25; ++i) mn = mn < a.mat[i] ? mn : a.mat[i];
        s_bias = -mn;
        for (int t = 0; t < 5; ++t) {
            uint32_t lo = 0;
            for (int q = 0; q < 4; ++q) lo |= (uint32_t)(uint8_t)(a.mat[t * 5 + q] - mn) << (8 * q);
            smat_lo[t] = lo; smat_hi[t] = (uint32_t)(uint8_t)(a.mat[t * 5 + 4] - mn);
        }
    }
    __syncthreads();
    const int bias = s_bias;
    const uint32_t gwarp = blockIdx.x * (BLOCK / 32) + (tid >> 5), n_warps = gridDim.x * (BLOCK / 32);
    uint32_t *const z = z_all + (uint64_t)gwarp * tlen_max * WPR * 32;
    const int oe_del = a.o_del + a.e_del, oe_ins = a.o_ins + a.e_ins;
    unsigned long long cells = 0;

    for (uint32_t chunk = gwarp; (uint64_t)chunk * 32 < n; chunk += n_warps) {
        const uint32_t idx = chunk * 32 + lane;
        const bool valid = idx < n;
        const uint32_t job = valid ? perm[idx] : 0u;
        const int qlen = valid ? (int)a.qlen[job] : 0, tlen = valid ? (int)a.tlen[job] : 0, w = valid ? (int)a.w[job] : 0;
        const uint8_t *q = a.qseq + (valid ? a.qoff[job] : 0u), *t = a.tseq + (valid ? a.toff[job] : 0u);
        stage_seq(qs, BLOCK, q, qlen, a.aligned8 != 0);
        stage_seq(ts, BLOCK, t, tlen, a.aligned8 != 0);
        auto qcode = [&](int j) -> uint32_t { return (qs[(j >> 3) * BLOCK] >> (28 - 4 * (j & 7))) & 15u; };
        // first row of the reference (src/ksw.c:1141-1147) in band coordinates: row 0 reads H(-1, j-1) at c = j + w
        int32_t B[NC], E[NC + 1];
#pragma unroll
        for (int c = 0; c < NC; ++c) { const int j = c - w; B[c] = j == 0 ? 0 : (j > 0 && j <= w ? -(a.o_ins + a.e_ins * j) : MINF); E[c] = MINF; }
        E[NC] = MINF;
        // query window: code of column c at bits 4c of {qlo, qhi}; row 0 holds j = 0 .. w at c = w .. 2w
        uint64_t qlo = 0, qhi = 0;
        for (int j = 0; j <= w && j < qlen; ++j) {
            const int c = j + w;
            if (c < 16) qlo |= (uint64_t)qcode(j) << (4 * c); else qhi |= (uint64_t)qcode(j) << (4 * (c - 16));
        }
        int warp_tl = tlen, warp_w = w;
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            warp_tl = max(warp_tl, __shfl_xor_sync(0xffffffffu, warp_tl, o));
            warp_w = max(warp_w, __shfl_xor_sync(0xffffffffu, warp_w, o));
        }
        uint32_t tword = 0;
        for (int i = 0; i < warp_tl; ++i) {
            const bool act = i < tlen;
            const int cbeg = w - i > 0 ? w - i : 0;                               // j >= 0
            int cend = qlen - i + w < 2 * w + 1 ? qlen - i + w : 2 * w + 1;        // j < qlen, c <= 2w
            if (!act || cend < cbeg) cend = cbeg;
            if ((i & 7) == 0 && act) tword = ts[(i >> 3) * BLOCK];
            const int tb = act ? (int)((tword >> (28 - 4 * (i & 7))) & 15u) : 4;
            const uint32_t mlo = smat_lo[tb], mhi = smat_hi[tb];
            int32_t f = MINF;
            uint32_t zw[WPR];
#pragma unroll
            for (int k = 0; k < WPR; ++k) zw[k] = 0;
            const int warp_nc = 2 * warp_w + 1;          // jobs are ordered by band width: a warp's lanes agree on it (almost) always
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (c >= warp_nc) break;                 // warp-uniform: the unrolled cells beyond the warp's band are skipped, not predicated
                if (c >= cbeg && c < cend) {
                    const uint32_t qc = c < 16 ? (uint32_t)(qlo >> (4 * c)) & 15u : (uint32_t)(qhi >> (4 * (c - 16))) & 15u;
                    const int32_t m = B[c] + (int)(__byte_perm(mlo, mhi, qc) & 0xffu) - bias;
                    bool m_ge_e, h_ge_f, t_ge_e, t_ge_f;
                    int32_t h = __vibmax_s32(m, E[c + 1], &m_ge_e);
                    uint32_t d = m_ge_e ? 0u : 1u;
                    h = __vibmax_s32(h, f, &h_ge_f);
                    d = h_ge_f ? d : 2u;
                    E[c] = __vibmax_s32(m - oe_del, E[c + 1] - a.e_del, &t_ge_e);
                    d |= t_ge_e ? 0u : 4u;
                    f = __vibmax_s32(m - oe_ins, f - a.e_ins, &t_ge_f);
                    d |= t_ge_f ? 0u : 8u;
                    B[c] = h;
                    zw[c >> 3] |= d << (4 * (c & 7));
                }
            }
            // the reference ends a row with eh[end] = {H(i, end-1), MINUS_INF}: the E read by the next row's new last column
            // (c = 2w there) is E[2w+1], never written, MINUS_INF; a row that stops at qlen needs nothing
            uint32_t *zrow = z + (uint64_t)i * WPR * 32 + lane;
#pragma unroll
            for (int k = 0; k < WPR; ++k) zrow[k * 32] = zw[k];
            cells += (unsigned long long)(cend - cbeg);
            if (i < warp_w) {
                // first-column boundary: the next row's column j = 0 (c = w - i - 1) reads H(i, -1) = -(o_del + e_del (i + 1)) as its diagonal
                const int cb = w - i - 1;
                const int32_t bnd = -(a.o_del + a.e_del * (i + 1));
#pragma unroll
                for (int c = 0; c < W; ++c) if (act && c == cb) B[c] = bnd;
            }
            // the window moves one base: column c takes the code of c + 1, the new base j = i + 1 + w enters at c = 2w
            qlo = (qlo >> 4) | (qhi << 60); qhi >>= 4;
            {
                const int jn = i + 1 + w;
                if (act && jn < qlen) {
                    const uint64_t code = qcode(jn);
                    const int c = 2 * w;
                    if (c < 16) qlo |= code << (4 * c); else qhi |= code << (4 * (c - 16));
                }
            }
        }
        if (!valid) continue;
        // score = eh[qlen].h = H(tlen-1, qlen-1) when the last row reaches column qlen - 1 and is not past it (see global_kernel)
        int32_t score;
        if (tlen == 0) score = qlen == 0 ? 0 : (qlen <= w ? -(a.o_ins + a.e_ins * qlen) : MINF);
        else if (qlen == 0) score = tlen - 1 <= w ? -(a.o_del + a.e_del * tlen) : MINF;     // no column at all: eh[0].h of the last row
        else if (tlen + w < qlen || tlen > qlen + w) score = MINF;
        else {
            const int cs = qlen - tlen + w;
            score = MINF;
#pragma unroll
            for (int c = 0; c < NC; ++c) if (c == cs) score = B[c];
        }
        // backtrack (src/ksw.c:1208-1231), cells addressed by c = k - i + w
        int i = tlen - 1, k = (i + w + 1 < qlen ? i + w + 1 : qlen) - 1;
        uint32_t which = 0, n_ops = 0, cur_op = 0, cur_len = 0, nmm = 0, gap = 0, edge_del = 0;
        uint32_t *row = a.rows + (uint64_t)job * a.cig_stride;
        auto flush = [&]() {
            if (cur_len == 0) return;
            if (n_ops < a.cig_stride) row[a.cig_stride - 1 - n_ops] = cur_len << 4 | cur_op;
            if (cur_op) gap += cur_len;
            if (cur_op == 2 && n_ops == 0) edge_del += cur_len;
            ++n_ops;
        };
        auto push = [&](uint32_t op, uint32_t len) {
            if (cur_len && op == cur_op) cur_len += len;
            else { flush(); cur_op = op; cur_len = len; }
        };
        while (i >= 0 && k >= 0) {
            int c = k - i + w;
            c = c < 0 ? 0 : (c > NC - 1 ? NC - 1 : c);
            const uint32_t code = (z[((uint64_t)i * WPR + (c >> 3)) * 32 + lane] >> (4 * (c & 7))) & 15u;
            which = which == 0 ? (code & 3u) : (which == 1 ? ((code >> 2) & 1u) : ((code >> 3) & 1u) * 2u);
            if (which == 0) {
                const uint32_t qc = qcode(k), tc = (ts[(i >> 3) * BLOCK] >> (28 - 4 * (i & 7))) & 15u;
                nmm += qc != tc; push(0, 1); --i; --k;
            }
            else if (which == 1) { push(2, 1); --i; }
            else { push(1, 1); --k; }
        }
        if (i >= 0) push(2, (uint32_t)(i + 1));
        if (k >= 0) push(1, (uint32_t)(k + 1));
        if (cur_len && cur_op == 2 && n_ops > 0) edge_del += cur_len;
        flush();
        a.score[job] = score;
        a.n_cigar[job] = n_ops;
        a.nm[job] = (int32_t)(nmm + gap - edge_del);
        if (n_ops > a.cig_stride) atomicMax(a.counters + 1, (unsigned long long)n_ops);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) cells += __shfl_xor_sync(0xffffffffu, cells, o);
    if (lane == 0 && cells) atomicAdd(a.counters, cells);
}

__global__ void __launch_bounds__(128)
compact_kernel(uint32_t n, const uint32_t *__restrict__ n_cigar, const uint64_t *__restrict__ off, const uint32_t *__restrict__ rows,
               uint32_t stride, uint32_t *__restrict__ flat)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    const uint32_t m = n_cigar[a];
    if (m > stride) return;
    const uint32_t *src = rows + (uint64_t)a * stride + (stride - m);
    uint32_t *dst = flat + off[a];
    for (uint32_t i = 0; i < m; ++i) dst[i] = src[i];
}

__global__ void total32_kernel(const uint32_t *n_per, const uint64_t *off, uint32_t n, unsigned long long *total)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *total = n ? off[n - 1] + n_per[n - 1] : 0ull;
}

struct U32ToU64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; } };

template <typename T> int grow_dev(T *&p, uint64_t &cap, uint64_t need)
{
    if (need <= cap && p) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    uint64_t c = need + need / 4 + 64;
    if (cudaMalloc(&p, c * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return 1; }
    cap = c;
    return 0;
}

struct ClassCfg { int R, block; uint32_t wmax; };
const ClassCfg CLASSES[5] = {{16, 128, 7}, {32, 128, 15}, {64, 64, 31}, {128, 32, 63}, {256, 32, 127}};

} // namespace

struct bwa_b200_cigar {
    int device = 0, n_sm = 0;
    cudaStream_t stream = nullptr;
    // inputs (host API)
    uint8_t *d_q = nullptr, *d_t = nullptr; uint64_t q_cap = 0, t_cap = 0;
    uint32_t *d_qoff = nullptr, *d_qlen = nullptr, *d_toff = nullptr, *d_tlen = nullptr; uint64_t qoff_cap = 0, qlen_cap = 0, toff_cap = 0, tlen_cap = 0;
    // per job
    uint32_t *d_w = nullptr, *d_perm = nullptr, *d_ncig = nullptr; uint64_t w_cap = 0, perm_cap = 0, ncig_cap = 0;
    int32_t *d_score = nullptr, *d_nm = nullptr; uint64_t score_cap = 0, nm_cap = 0;
    uint64_t *d_off = nullptr; uint64_t off_cap = 0;
    uint32_t *d_rows = nullptr; uint64_t rows_cap = 0; uint32_t cig_stride = 16;
    uint32_t *d_flat = nullptr; uint64_t flat_cap = 0;
    uint32_t *d_z = nullptr; uint64_t z_cap = 0;
    void *d_cub = nullptr; uint64_t cub_cap = 0;
    unsigned long long *d_counters = nullptr, *h_counters = nullptr;   // [0] cells [1] widest row needed [2] total operations
    int grid[5] = {0, 0, 0, 0, 0};
    cudaStream_t side[5] = {};              // band classes of one batch run concurrently (forked from / joined to `stream`)
    cudaEvent_t ev_fork = nullptr, ev_join[5] = {};
    std::vector<uint32_t> perm; uint32_t cls_n[5] = {}, cls_tl[5] = {}, cls_ql[5] = {};
    // bwa_b200_reg2aln_host: the batch's reads and region table on the device
    uint32_t *r_packed = nullptr; uint64_t r_packed_cap = 0; uint64_t *r_woff = nullptr; uint64_t r_woff_cap = 0;
    void *r_jobs = nullptr; uint64_t r_jobs_cap = 0;
    int smem_optin = 0;
    bool use_band = true;
    uint64_t z_off[5] = {};
    uint64_t last_n = 0, last_ops = 0, last_cells = 0, launches = 0;
    b200::Prof prof; int profiling = 0;
    // pinned host results of bwa_b200_global_host_view (grown geometrically, reused batch after batch)
    int32_t *p_score = nullptr, *p_nm = nullptr; uint32_t *p_ncig = nullptr, *p_flat = nullptr; uint64_t *p_off = nullptr;
    uint64_t p_jobs_cap = 0, p_ops_cap = 0;
};

// resident blocks of one class for the shared memory its longest sequences need: at most 24 warps per SM, so that the backtrack
// slabs of the resident warps stay of the order of the L2 size
template <int R, int BLOCK> static int class_grid(int n_sm, size_t smem, int *grid)
{
    int occ = 0;
    B200_CUDA(cudaFuncSetAttribute(global_kernel<R, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, global_kernel<R, BLOCK>, BLOCK, smem));
    if (occ < 1) occ = 1;
    if (occ * (BLOCK / 32) > 24) occ = std::max(1, 24 / (BLOCK / 32));
    *grid = n_sm * occ;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_cigar_create(int device, bwa_b200_cigar_t **out)
{
    if (!out) { b200::set_error("cigar_create: bad argument"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(device));
    bwa_b200_cigar *c = new bwa_b200_cigar();
    c->device = device;
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    c->n_sm = prop.multiProcessorCount;
    B200_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    B200_CUDA(cudaMalloc(&c->d_counters, 4 * sizeof(unsigned long long)));
    B200_CUDA(cudaHostAlloc(&c->h_counters, 4 * sizeof(unsigned long long), cudaHostAllocDefault));
    c->smem_optin = (int)prop.sharedMemPerBlockOptin;
    B200_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    for (int k = 0; k < 5; ++k) {
        B200_CUDA(cudaStreamCreateWithFlags(&c->side[k], cudaStreamNonBlocking));
        B200_CUDA(cudaEventCreateWithFlags(&c->ev_join[k], cudaEventDisableTiming));
    }
    *out = c;
    return BWA_B200_OK;
}

extern "C" void bwa_b200_cigar_destroy(bwa_b200_cigar_t *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaFree(c->d_q); cudaFree(c->d_t); cudaFree(c->d_qoff); cudaFree(c->d_qlen); cudaFree(c->d_toff); cudaFree(c->d_tlen);
    cudaFree(c->d_w); cudaFree(c->d_perm); cudaFree(c->d_ncig); cudaFree(c->d_score); cudaFree(c->d_nm); cudaFree(c->d_off);
    cudaFree(c->d_rows); cudaFree(c->d_flat); cudaFree(c->d_z); cudaFree(c->d_cub); cudaFree(c->d_counters);
    cudaFree(c->r_packed); cudaFree(c->r_woff); cudaFree(c->r_jobs);
    cudaFreeHost(c->h_counters);
    cudaFreeHost(c->p_score); cudaFreeHost(c->p_nm); cudaFreeHost(c->p_ncig); cudaFreeHost(c->p_off); cudaFreeHost(c->p_flat);
    for (int k = 0; k < 5; ++k) { if (c->side[k]) cudaStreamDestroy(c->side[k]); if (c->ev_join[k]) cudaEventDestroy(c->ev_join[k]); }
    if (c->ev_fork) cudaEventDestroy(c->ev_fork);
    cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int bwa_b200_cigar_band(const bwa_b200_ext_params_t *p, int w_, int l_query, int64_t rlen)
{ // src/bwa.c:161-169
    if (!p) return -1;
    int max_ins = (int)((double)(((l_query + 1) >> 1) * p->mat[0] - p->o_ins) / p->e_ins + 1.);
    int max_del = (int)((double)(((l_query + 1) >> 1) * p->mat[0] - p->o_del) / p->e_del + 1.);
    int max_gap = max_ins > max_del ? max_ins : max_del;
    max_gap = max_gap > 1 ? max_gap : 1;
    int diff = (int)(rlen - l_query); if (diff < 0) diff = -diff;
    int w = (max_gap + diff + 1) >> 1;
    w = w < w_ ? w : w_;
    int min_w = diff + 3;
    return w > min_w ? w : min_w;
}

template <int R, int BLOCK> static size_t class_smem(const bwa_b200_cigar *c, int cls)
{
    return (size_t)R * BLOCK * sizeof(int2) + ((size_t)(c->cls_ql[cls] + 7) / 8 + (size_t)(c->cls_tl[cls] + 7) / 8) * BLOCK * 4;
}
// the register-resident kernel of the two narrow classes needs the staged sequences only
template <int BLOCK> static size_t band_smem(const bwa_b200_cigar *c, int cls)
{
    return ((size_t)(c->cls_ql[cls] + 7) / 8 + (size_t)(c->cls_tl[cls] + 7) / 8) * BLOCK * 4;
}
template <int W, int BLOCK> static int band_grid(int n_sm, size_t smem, int *grid)
{
    int occ = 0;
    B200_CUDA(cudaFuncSetAttribute(global_band_kernel<W, BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, global_band_kernel<W, BLOCK>, BLOCK, smem));
    if (occ < 1) occ = 1;
    if (occ * (BLOCK / 32) > 24) occ = std::max(1, 24 / (BLOCK / 32));
    *grid = n_sm * occ;
    return BWA_B200_OK;
}
template <int W, int BLOCK>
static void launch_band(bwa_b200_cigar *c, const GArgs &ga, int cls, uint32_t first, const char *name)
{
    b200::Prof *prof = c->profiling ? &c->prof : nullptr;
    cudaStream_t st = prof ? c->stream : c->side[cls];
    B200_LAUNCH(prof, name, st,
        (global_band_kernel<W, BLOCK><<<c->grid[cls], BLOCK, band_smem<BLOCK>(c, cls), st>>>(ga, c->d_perm + first, c->cls_n[cls], c->cls_tl[cls],
                                                                                            (c->cls_ql[cls] + 7) / 8, (c->cls_tl[cls] + 7) / 8, c->d_z + c->z_off[cls])));
    ++c->launches;
}
template <int R, int BLOCK>
static void launch_class(bwa_b200_cigar *c, const GArgs &ga, int cls, uint32_t first, const char *name)
{
    b200::Prof *prof = c->profiling ? &c->prof : nullptr;
    cudaStream_t st = prof ? c->stream : c->side[cls];      // profiled: one after another on the main stream
    B200_LAUNCH(prof, name, st,
        (global_kernel<R, BLOCK><<<c->grid[cls], BLOCK, class_smem<R, BLOCK>(c, cls), st>>>(ga, c->d_perm + first, c->cls_n[cls], c->cls_tl[cls],
                                                                                                 (c->cls_ql[cls] + 7) / 8, (c->cls_tl[cls] + 7) / 8, c->d_z + c->z_off[cls])));
    ++c->launches;
}

// jobs already on the device (byte per base); host copies of tlen and w drive the binning
static int cigar_run(bwa_b200_cigar *c, const bwa_b200_ext_params_t *p, uint64_t n_jobs, const uint8_t *d_q, const uint32_t *d_qoff,
                     const uint32_t *d_qlen, const uint8_t *d_t, const uint32_t *d_toff, const uint32_t *d_tlen, const uint32_t *d_w,
                     const uint32_t *h_qlen, const uint32_t *h_tlen, const uint32_t *h_w, bool aligned8)
{
    c->last_n = n_jobs; c->last_ops = 0; c->last_cells = 0;
    if (n_jobs == 0) return BWA_B200_OK;
    if (n_jobs > 0xfffffff0ull) { b200::set_error("global: at most 2^32-16 jobs per batch"); return BWA_B200_ERR_ARG; }
    const uint32_t n = (uint32_t)n_jobs;
    // bin the jobs by band-width class
    uint32_t cnt[5] = {0, 0, 0, 0, 0}, tl[5] = {0, 0, 0, 0, 0}, ql[5] = {0, 0, 0, 0, 0};
    for (uint32_t a = 0; a < n; ++a) {
        const uint32_t w = h_w[a];
        int k = 0;
        while (k < 5 && w > CLASSES[k].wmax) ++k;
        if (k == 5) { b200::set_error("global: band %u of job %u is wider than 127", w, a); return BWA_B200_ERR_ARG; }
        ++cnt[k]; tl[k] = std::max(tl[k], h_tlen[a]); ql[k] = std::max(ql[k], h_qlen[a]);
    }
    uint32_t first[6] = {0, 0, 0, 0, 0, 0};
    for (int k = 0; k < 5; ++k) { first[k + 1] = first[k] + cnt[k]; c->cls_n[k] = cnt[k]; c->cls_tl[k] = tl[k] ? tl[k] : 1; c->cls_ql[k] = ql[k] ? ql[k] : 1; }
    {   // shared memory of a class = ring + its longest query and target; the grid follows from the occupancy at that size
        // w <= 15: column state in registers (global_band_kernel); BWA_B200_GLOBAL_RING=1 sends those classes through the shared-memory
        // ring kernel as well (tests run both)
        c->use_band = getenv("BWA_B200_GLOBAL_RING") == nullptr;
        const size_t sm[5] = {c->use_band ? band_smem<128>(c, 0) : class_smem<16, 128>(c, 0), c->use_band ? band_smem<128>(c, 1) : class_smem<32, 128>(c, 1),
                              class_smem<64, 64>(c, 2), class_smem<128, 32>(c, 3), class_smem<256, 32>(c, 4)};
        for (int k = 0; k < 5; ++k)
            if (cnt[k] && sm[k] > (size_t)c->smem_optin) { b200::set_error("global: sequences of %u / %u bases in band class %d do not fit shared memory", ql[k], tl[k], k); return BWA_B200_ERR_CAPACITY; }
        int rc = 0;
        if (cnt[0]) rc |= c->use_band ? band_grid<7, 128>(c->n_sm, sm[0], &c->grid[0]) : class_grid<16, 128>(c->n_sm, sm[0], &c->grid[0]);
        if (cnt[1]) rc |= c->use_band ? band_grid<15, 128>(c->n_sm, sm[1], &c->grid[1]) : class_grid<32, 128>(c->n_sm, sm[1], &c->grid[1]);
        if (cnt[2]) rc |= class_grid<64, 64>(c->n_sm, sm[2], &c->grid[2]);
        if (cnt[3]) rc |= class_grid<128, 32>(c->n_sm, sm[3], &c->grid[3]);
        if (cnt[4]) rc |= class_grid<256, 32>(c->n_sm, sm[4], &c->grid[4]);
        if (rc) return BWA_B200_ERR_CUDA;
    }
    // within a class jobs are ordered by band width (counting sort), so the 32 jobs of a warp have (nearly) the same number of
    // columns per row and no lane idles through another lane's wider band
    c->perm.resize(n);
    {
        uint32_t hist[129];
        memset(hist, 0, sizeof(hist));
        for (uint32_t a = 0; a < n; ++a) ++hist[h_w[a] + 1];
        for (int v = 0; v < 128; ++v) hist[v + 1] += hist[v];                    // hist[w] = first position of band w (classes are contiguous in w)
        for (uint32_t a = 0; a < n; ++a) c->perm[hist[h_w[a]]++] = a;
    }
    uint64_t z_need = 1;                    // the classes run concurrently: each has its own region of backtrack slabs
    for (int k = 0; k < 5; ++k) {
        c->z_off[k] = z_need;
        if (cnt[k]) z_need += (uint64_t)c->grid[k] * (CLASSES[k].block / 32) * c->cls_tl[k] * (CLASSES[k].R / 8) * 32;
    }
    int bad = 0;
    bad |= grow_dev(c->d_perm, c->perm_cap, n); bad |= grow_dev(c->d_ncig, c->ncig_cap, n); bad |= grow_dev(c->d_score, c->score_cap, n);
    bad |= grow_dev(c->d_nm, c->nm_cap, n); bad |= grow_dev(c->d_off, c->off_cap, n); bad |= grow_dev(c->d_z, c->z_cap, z_need);
    bad |= grow_dev(c->d_rows, c->rows_cap, (uint64_t)n * c->cig_stride);
    size_t cub_bytes = 0;
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t *> it(c->d_ncig, U32ToU64());
    B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, it, c->d_off, (int)n, c->stream));
    uint8_t *cubp = (uint8_t *)c->d_cub;
    bad |= grow_dev(cubp, c->cub_cap, cub_bytes + 16);
    c->d_cub = cubp;
    if (bad) { b200::set_error("global: out of device memory"); return BWA_B200_ERR_NOMEM; }
    cudaStream_t st = c->stream;
    B200_CUDA(cudaMemcpyAsync(c->d_perm, c->perm.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (c->profiling) c->prof.reset();
    for (;;) {
        B200_CUDA(cudaMemsetAsync(c->d_counters, 0, 4 * sizeof(unsigned long long), st));
        GArgs ga;
        ga.qseq = d_q; ga.tseq = d_t; ga.qoff = d_qoff; ga.qlen = d_qlen; ga.toff = d_toff; ga.tlen = d_tlen; ga.w = d_w;
        memcpy(ga.mat, p->mat, 25);
        ga.o_del = p->o_del; ga.e_del = p->e_del; ga.o_ins = p->o_ins; ga.e_ins = p->e_ins;
        ga.score = c->d_score; ga.nm = c->d_nm; ga.n_cigar = c->d_ncig; ga.rows = c->d_rows; ga.cig_stride = c->cig_stride; ga.aligned8 = aligned8 ? 1u : 0u;
        ga.counters = c->d_counters;
        const bool fan = !c->profiling;
        if (fan) {
            B200_CUDA(cudaEventRecord(c->ev_fork, st));
            for (int k = 0; k < 5; ++k) if (cnt[k]) B200_CUDA(cudaStreamWaitEvent(c->side[k], c->ev_fork, 0));
        }
        if (cnt[0]) { if (c->use_band) launch_band<7, 128>(c, ga, 0, first[0], "global_band_kernel_w7"); else launch_class<16, 128>(c, ga, 0, first[0], "global_kernel_w7"); }
        if (cnt[1]) { if (c->use_band) launch_band<15, 128>(c, ga, 1, first[1], "global_band_kernel_w15"); else launch_class<32, 128>(c, ga, 1, first[1], "global_kernel_w15"); }
        if (cnt[2]) launch_class<64, 64>(c, ga, 2, first[2], "global_kernel_w31");
        if (cnt[3]) launch_class<128, 32>(c, ga, 3, first[3], "global_kernel_w63");
        if (cnt[4]) launch_class<256, 32>(c, ga, 4, first[4], "global_kernel_w127");
        if (fan)
            for (int k = 0; k < 5; ++k) if (cnt[k]) { B200_CUDA(cudaEventRecord(c->ev_join[k], c->side[k])); B200_CUDA(cudaStreamWaitEvent(st, c->ev_join[k], 0)); }
        B200_CUDA(cudaGetLastError());
        size_t tmp = c->cub_cap;
        B200_CUDA(cub::DeviceScan::ExclusiveSum(c->d_cub, tmp, it, c->d_off, (int)n, st));
        total32_kernel<<<1, 1, 0, st>>>(c->d_ncig, c->d_off, n, c->d_counters + 2);
        c->launches += 2;
        B200_CUDA(cudaMemcpyAsync(c->h_counters, c->d_counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        B200_CUDA(cudaStreamSynchronize(st));
        if (c->h_counters[1] <= c->cig_stride) break;
        // some CIGAR has more operations than a row holds: widen the rows and run the batch again (rare)
        c->cig_stride = (uint32_t)c->h_counters[1] + 8;
        if (grow_dev(c->d_rows, c->rows_cap, (uint64_t)n * c->cig_stride)) { b200::set_error("global: out of device memory"); return BWA_B200_ERR_NOMEM; }
    }
    c->last_cells = c->h_counters[0];
    c->last_ops = c->h_counters[2];
    if (grow_dev(c->d_flat, c->flat_cap, c->last_ops ? c->last_ops : 1)) { b200::set_error("global: out of device memory"); return BWA_B200_ERR_NOMEM; }
    b200::Prof *prof = c->profiling ? &c->prof : nullptr;
    B200_LAUNCH(prof, "compact_kernel", st,
        (compact_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, c->d_ncig, c->d_off, c->d_rows, c->cig_stride, c->d_flat)));
    ++c->launches;
    B200_CUDA(cudaGetLastError());
    return BWA_B200_OK;
}

extern "C" int bwa_b200_global_device(bwa_b200_cigar_t *c, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                                      const uint8_t *dev_qseq, const uint32_t *dev_qoff, const uint32_t *dev_qlen,
                                      const uint8_t *dev_tseq, const uint32_t *dev_toff, const uint32_t *dev_tlen,
                                      const uint32_t *host_qlen, const uint32_t *host_tlen, const uint32_t *host_w, int aligned8)
{
    if (!c || !p || (n_jobs && (!dev_qseq || !dev_qoff || !dev_qlen || !dev_tseq || !dev_toff || !dev_tlen || !host_qlen || !host_tlen || !host_w))) {
        b200::set_error("global_device: bad argument"); return BWA_B200_ERR_ARG;
    }
    B200_CUDA(cudaSetDevice(c->device));
    if (n_jobs) {
        if (grow_dev(c->d_w, c->w_cap, n_jobs)) { b200::set_error("global: out of device memory"); return BWA_B200_ERR_NOMEM; }
        B200_CUDA(cudaMemcpyAsync(c->d_w, host_w, n_jobs * 4, cudaMemcpyHostToDevice, c->stream));
    }
    const bool al = aligned8 != 0 && ((uintptr_t)dev_qseq & 7) == 0 && ((uintptr_t)dev_tseq & 7) == 0;
    return cigar_run(c, p, n_jobs, dev_qseq, dev_qoff, dev_qlen, dev_tseq, dev_toff, dev_tlen, c->d_w, host_qlen, host_tlen, host_w, al);
}

extern "C" int bwa_b200_global_device_view(bwa_b200_cigar_t *c, bwa_b200_cigars_t *v)
{
    if (!c || !v) return BWA_B200_ERR_ARG;
    B200_CUDA(cudaSetDevice(c->device));
    B200_CUDA(cudaStreamSynchronize(c->stream));
    v->n_jobs = c->last_n; v->n_ops = c->last_ops; v->score = c->d_score; v->nm = c->d_nm; v->n_cigar = c->d_ncig; v->cigar_off = c->d_off; v->cigar = c->d_flat;
    return BWA_B200_OK;
}

extern "C" void bwa_b200_cigars_free(bwa_b200_cigars_t *r)
{
    if (!r) return;
    free(r->score); free(r->nm); free(r->n_cigar); free(r->cigar_off); free(r->cigar);
    memset(r, 0, sizeof(*r));
}

// host job batch -> device, kernels; results left on the device (c->d_score, d_nm, d_ncig, d_off, d_flat; c->last_ops operations)
static int global_host_run(bwa_b200_cigar_t *c, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                           const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                           const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen, const uint32_t *w)
{
    bool al = true;                        // the GASAL host layout: 8-byte aligned, padded sequences -> 64-bit loads in the kernel
    for (uint64_t a = 0; a < n_jobs; ++a) {
        if ((uint64_t)qoff[a] + qlen[a] > q_bytes || (uint64_t)toff[a] + tlen[a] > t_bytes) { b200::set_error("global_host: job %llu reaches past its sequence buffer", (unsigned long long)a); return BWA_B200_ERR_ARG; }
        if ((qoff[a] & 7) || (toff[a] & 7) || (uint64_t)qoff[a] + ((uint64_t)qlen[a] + 7) / 8 * 8 > q_bytes || (uint64_t)toff[a] + ((uint64_t)tlen[a] + 7) / 8 * 8 > t_bytes) al = false;
    }
    B200_CUDA(cudaSetDevice(c->device));
    int bad = 0;
    bad |= grow_dev(c->d_q, c->q_cap, q_bytes ? q_bytes : 1); bad |= grow_dev(c->d_t, c->t_cap, t_bytes ? t_bytes : 1);
    bad |= grow_dev(c->d_qoff, c->qoff_cap, n_jobs); bad |= grow_dev(c->d_qlen, c->qlen_cap, n_jobs);
    bad |= grow_dev(c->d_toff, c->toff_cap, n_jobs); bad |= grow_dev(c->d_tlen, c->tlen_cap, n_jobs); bad |= grow_dev(c->d_w, c->w_cap, n_jobs);
    if (bad) { b200::set_error("global: out of device memory"); return BWA_B200_ERR_NOMEM; }
    cudaStream_t st = c->stream;
    if (q_bytes) B200_CUDA(cudaMemcpyAsync(c->d_q, qseq, q_bytes, cudaMemcpyHostToDevice, st));
    if (t_bytes) B200_CUDA(cudaMemcpyAsync(c->d_t, tseq, t_bytes, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(c->d_qoff, qoff, n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(c->d_qlen, qlen, n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(c->d_toff, toff, n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(c->d_tlen, tlen, n_jobs * 4, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(c->d_w, w, n_jobs * 4, cudaMemcpyHostToDevice, st));
    return cigar_run(c, p, n_jobs, c->d_q, c->d_qoff, c->d_qlen, c->d_t, c->d_toff, c->d_tlen, c->d_w, qlen, tlen, w, al);
}

static int global_download(bwa_b200_cigar_t *c, uint64_t n_jobs, int32_t *score, int32_t *nm, uint32_t *n_cigar, uint64_t *cigar_off, uint32_t *cigar)
{
    cudaStream_t st = c->stream;
    const uint64_t ops = c->last_ops;
    B200_CUDA(cudaMemcpyAsync(score, c->d_score, n_jobs * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(nm, c->d_nm, n_jobs * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(n_cigar, c->d_ncig, n_jobs * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(cigar_off, c->d_off, n_jobs * 8, cudaMemcpyDeviceToHost, st));
    if (ops) B200_CUDA(cudaMemcpyAsync(cigar, c->d_flat, ops * 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaStreamSynchronize(st));
    return BWA_B200_OK;
}

extern "C" int bwa_b200_global_host(bwa_b200_cigar_t *c, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                                    const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                                    const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen,
                                    const uint32_t *w, bwa_b200_cigars_t *out)
{
    if (!c || !p || !out || (n_jobs && (!qseq || !qoff || !qlen || !tseq || !toff || !tlen || !w))) { b200::set_error("global_host: bad argument"); return BWA_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    out->n_jobs = n_jobs;
    if (n_jobs == 0) return BWA_B200_OK;
    int rc = global_host_run(c, p, n_jobs, qseq, q_bytes, qoff, qlen, tseq, t_bytes, toff, tlen, w);
    if (rc) return rc;
    const uint64_t ops = c->last_ops;
    out->n_ops = ops;
    out->score = (int32_t *)malloc(n_jobs * 4); out->nm = (int32_t *)malloc(n_jobs * 4); out->n_cigar = (uint32_t *)malloc(n_jobs * 4);
    out->cigar_off = (uint64_t *)malloc(n_jobs * 8); out->cigar = (uint32_t *)malloc((ops ? ops : 1) * 4);
    if (!out->score || !out->nm || !out->n_cigar || !out->cigar_off || !out->cigar) { bwa_b200_cigars_free(out); b200::set_error("global_host: out of host memory"); return BWA_B200_ERR_NOMEM; }
    rc = global_download(c, n_jobs, out->score, out->nm, out->n_cigar, out->cigar_off, out->cigar);
    if (rc) bwa_b200_cigars_free(out);
    return rc;
}

// the same with the results in pinned buffers owned by the handle: the per-batch path of a driver (no allocation, no pageable
// staging; with the inputs in pinned memory as well, every copy is one asynchronous DMA)
extern "C" int bwa_b200_global_host_view(bwa_b200_cigar_t *c, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                                         const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                                         const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen,
                                         const uint32_t *w, bwa_b200_cigars_t *view)
{
    if (!c || !p || !view || (n_jobs && (!qseq || !qoff || !qlen || !tseq || !toff || !tlen || !w))) { b200::set_error("global_host_view: bad argument"); return BWA_B200_ERR_ARG; }
    memset(view, 0, sizeof(*view));
    view->n_jobs = n_jobs;
    if (n_jobs == 0) return BWA_B200_OK;
    int rc = global_host_run(c, p, n_jobs, qseq, q_bytes, qoff, qlen, tseq, t_bytes, toff, tlen, w);
    if (rc) return rc;
    const uint64_t ops = c->last_ops;
    if (n_jobs > c->p_jobs_cap) {
        cudaFreeHost(c->p_score); cudaFreeHost(c->p_nm); cudaFreeHost(c->p_ncig); cudaFreeHost(c->p_off);
        c->p_score = c->p_nm = nullptr; c->p_ncig = nullptr; c->p_off = nullptr; c->p_jobs_cap = 0;
        const uint64_t cap = n_jobs + n_jobs / 4 + 256;
        B200_CUDA(cudaHostAlloc(&c->p_score, cap * 4, cudaHostAllocDefault)); B200_CUDA(cudaHostAlloc(&c->p_nm, cap * 4, cudaHostAllocDefault));
        B200_CUDA(cudaHostAlloc(&c->p_ncig, cap * 4, cudaHostAllocDefault)); B200_CUDA(cudaHostAlloc(&c->p_off, cap * 8, cudaHostAllocDefault));
        c->p_jobs_cap = cap;
    }
    if (ops > c->p_ops_cap || !c->p_flat) {
        cudaFreeHost(c->p_flat); c->p_flat = nullptr; c->p_ops_cap = 0;
        const uint64_t cap = ops + ops / 4 + 1024;
        B200_CUDA(cudaHostAlloc(&c->p_flat, cap * 4, cudaHostAllocDefault));
        c->p_ops_cap = cap;
    }
    rc = global_download(c, n_jobs, c->p_score, c->p_nm, c->p_ncig, c->p_off, c->p_flat);
    if (rc) return rc;
    view->n_ops = ops; view->score = c->p_score; view->nm = c->p_nm; view->n_cigar = c->p_ncig; view->cigar_off = c->p_off; view->cigar = c->p_flat;
    return BWA_B200_OK;
}

extern "C" void *bwa_b200_cigar_stream(bwa_b200_cigar_t *c) { return c ? (void *)c->stream : nullptr; }
extern "C" uint64_t bwa_b200_cigar_launches(const bwa_b200_cigar_t *c) { return c ? c->launches : 0; }
extern "C" uint64_t bwa_b200_cigar_last_cells(const bwa_b200_cigar_t *c) { return c ? c->last_cells : 0; }
extern "C" int bwa_b200_cigar_profile(bwa_b200_cigar_t *c, int enable) { if (!c) return BWA_B200_ERR_ARG; c->profiling = enable != 0; return BWA_B200_OK; }
extern "C" int bwa_b200_cigar_kernel_times(bwa_b200_cigar_t *c, const char **names, float *ms, int cap)
{
    if (!c) return BWA_B200_ERR_ARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    int n = 0;
    for (size_t i = 0; i < c->prof.used && n < cap; ++i, ++n) {
        float t = 0;
        cudaEventElapsedTime(&t, c->prof.recs[i].a, c->prof.recs[i].b);
        names[n] = c->prof.recs[i].name; ms[n] = t;
    }
    return n;
}

// ============================================================================ mem_reg2aln, batched
// The caller of the CIGAR path (src/bwamem.c:2344-2438).  Per alignment region: band inference (infer_bw), bwa_gen_cigar2 with up to
// three band-doubling retries, squeeze of a leading / trailing deletion, soft clips, forward position and contig.  Here a retry is a
// wave: every region still improving goes through global_kernel again with a doubled band.  Sequences never come from the host:
// r2a_cut_kernel cuts query and reference window of every region from the packed reads and the resident 2-bit reference, in the
// orientation bwa_gen_cigar2 aligns them (both reversed for a reverse-strand hit, src/bwa.c:145-150).
namespace {

struct R2AJob { uint32_t read; int32_t qb, qe; int64_t rb, re; uint32_t qoff, toff; };

__global__ void __launch_bounds__(128)
r2a_cut_kernel(uint32_t n, const R2AJob *__restrict__ jobs, const uint32_t *__restrict__ packed, const uint64_t *__restrict__ word_off,
               const uint32_t *__restrict__ pac, uint64_t l_pac, uint8_t *__restrict__ qout, uint8_t *__restrict__ tout)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
    if (warp >= n) return;
    const R2AJob J = jobs[warp];
    const int ql = J.qe - J.qb, tl = (int)(J.re - J.rb);
    const bool rev = J.rb >= (int64_t)l_pac;
    const uint64_t wo = word_off[J.read];
    for (int j = (int)lane; j < ql; j += 32) {
        const int pos = rev ? J.qe - 1 - j : J.qb + j;
        uint32_t c = (packed[wo + ((uint32_t)pos >> 3)] >> (28 - 4 * (pos & 7))) & 15u;
        qout[J.qoff + j] = (uint8_t)(c > 4u ? 4u : c);
    }
    const uint64_t t0 = rev ? 2 * l_pac - (uint64_t)J.re : (uint64_t)J.rb;      // forward-strand start of the window
    for (int j = (int)lane; j < tl; j += 32) {
        const uint64_t g = t0 + (uint64_t)j;
        const uint32_t b = (pac[g >> 4] >> ((~g & 15u) << 1)) & 3u;
        tout[J.toff + j] = (uint8_t)(rev ? 3u - b : b);
    }
}

int infer_bw(int l1, int l2, int score, int a, int q, int r)
{ // src/bwamem.c:1486-1494
    if (l1 == l2 && l1 * a - score < (q + r - a) << 1) return 0;
    int w = (int)((double)((l1 < l2 ? l1 : l2) * a - score - q) / r + 2.);
    const int d = l1 > l2 ? l1 - l2 : l2 - l1;
    return w < d ? d : w;
}

} // namespace

extern "C" int bwa_b200_reg2aln_host(bwa_b200_cigar_t *c, const bwa_b200_index_t *idx, int32_t n_ctg, const int64_t *ctg_off,
                                     const uint32_t *packed, const uint64_t *word_off, const uint32_t *read_len, uint64_t n_reads,
                                     const bwa_b200_aln_in_t *alns, uint64_t n_alns, const bwa_b200_ext_params_t *p, int32_t match_score,
                                     bwa_b200_aln_out_t *out, uint32_t **cigar, uint64_t *n_ops)
{
    if (!c || !idx || !p || !out || !cigar || !n_ops || n_ctg < 1 || !ctg_off || (n_alns && (!alns || !packed || !word_off || !read_len))) {
        b200::set_error("reg2aln: bad argument"); return BWA_B200_ERR_ARG;
    }
    *cigar = nullptr; *n_ops = 0;
    if (!idx->d_pac) { b200::set_error("reg2aln: the index has no reference attached (bwa_b200_index_attach_ref)"); return BWA_B200_ERR_ARG; }
    if (idx->device != c->device) { b200::set_error("reg2aln: index and CIGAR handle live on different devices"); return BWA_B200_ERR_ARG; }
    if (n_alns > 0x7ffffff0ull) { b200::set_error("reg2aln: too many regions"); return BWA_B200_ERR_ARG; }
    const int64_t l_pac = (int64_t)idx->l_pac;
    const int opt_w = p->w, a_sc = match_score;
    // ---- jobs: every mapped region the reference would hand to ksw_global2
    struct St { int w2, last_sc, iter, score, nm, n_waves; bool job, done; std::vector<uint32_t> cig; };
    std::vector<St> st(n_alns);
    std::vector<R2AJob> jobs; std::vector<uint32_t> job_aln;
    uint64_t qbytes = 0, tbytes = 0;
    for (uint64_t k = 0; k < n_alns; ++k) {
        const bwa_b200_aln_in_t &r = alns[k];
        St &s = st[k];
        s.w2 = 0; s.last_sc = -(1 << 30); s.iter = 0; s.score = 0; s.nm = -1; s.n_waves = 0; s.job = false; s.done = true;
        if (r.rb < 0 || r.re < 0) continue;                                   // unmapped record
        if (r.read >= n_reads || r.qb < 0 || r.qe > (int32_t)read_len[r.read]) { b200::set_error("reg2aln: region %llu does not lie on its read", (unsigned long long)k); return BWA_B200_ERR_ARG; }
        if (r.re > 2 * l_pac) { b200::set_error("reg2aln: region %llu reaches past the reference", (unsigned long long)k); return BWA_B200_ERR_ARG; }
        const int lq = r.qe - r.qb; const int64_t rl = r.re - r.rb;
        int tmp = infer_bw(lq, (int)rl, r.truesc, a_sc, p->o_del, p->e_del);
        int w2 = infer_bw(lq, (int)rl, r.truesc, a_sc, p->o_ins, p->e_ins);
        w2 = w2 > tmp ? w2 : tmp;
        if (w2 > opt_w) w2 = w2 < r.w ? w2 : r.w;
        s.w2 = w2; s.done = false;
        if (lq <= 0 || r.rb >= r.re || (r.rb < l_pac && r.re > l_pac)) continue;   // bwa_gen_cigar2 rejects it (src/bwa.c:124): no DP, score stays 0
        s.job = true;
        R2AJob j; j.read = r.read; j.qb = r.qb; j.qe = r.qe; j.rb = r.rb; j.re = r.re; j.qoff = (uint32_t)qbytes; j.toff = (uint32_t)tbytes;
        qbytes += ((uint64_t)lq + 7) / 8 * 8; tbytes += ((uint64_t)rl + 7) / 8 * 8;
        if (qbytes > 0xfffffff0ull || tbytes > 0xfffffff0ull) { b200::set_error("reg2aln: more than 4 GB of sequence in one batch"); return BWA_B200_ERR_CAPACITY; }
        jobs.push_back(j); job_aln.push_back((uint32_t)k);
    }
    B200_CUDA(cudaSetDevice(c->device));
    cudaStream_t stq = c->stream;
    const uint64_t nj = jobs.size();
    if (nj) {
        // reads + job table up, sequences cut on the device
        const uint64_t n_words = word_off[n_reads];
        int bad = 0;
        bad |= grow_dev(c->r_packed, c->r_packed_cap, n_words ? n_words : 1); bad |= grow_dev(c->r_woff, c->r_woff_cap, n_reads + 1);
        bad |= grow_dev(c->d_q, c->q_cap, qbytes ? qbytes : 1); bad |= grow_dev(c->d_t, c->t_cap, tbytes ? tbytes : 1);
        uint8_t *jb = (uint8_t *)c->r_jobs;
        bad |= grow_dev(jb, c->r_jobs_cap, nj * sizeof(R2AJob));
        c->r_jobs = jb;
        bad |= grow_dev(c->d_qoff, c->qoff_cap, nj); bad |= grow_dev(c->d_qlen, c->qlen_cap, nj); bad |= grow_dev(c->d_toff, c->toff_cap, nj);
        bad |= grow_dev(c->d_tlen, c->tlen_cap, nj); bad |= grow_dev(c->d_w, c->w_cap, nj);
        if (bad) { b200::set_error("reg2aln: out of device memory"); return BWA_B200_ERR_NOMEM; }
        B200_CUDA(cudaMemcpyAsync(c->r_packed, packed, n_words * 4, cudaMemcpyHostToDevice, stq));
        B200_CUDA(cudaMemcpyAsync(c->r_woff, word_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, stq));
        B200_CUDA(cudaMemcpyAsync(c->r_jobs, jobs.data(), nj * sizeof(R2AJob), cudaMemcpyHostToDevice, stq));
        B200_CUDA(cudaMemsetAsync(c->d_q, 4, qbytes ? qbytes : 1, stq));
        B200_CUDA(cudaMemsetAsync(c->d_t, 4, tbytes ? tbytes : 1, stq));
        r2a_cut_kernel<<<(unsigned)((nj * 32 + 127) / 128), 128, 0, stq>>>((uint32_t)nj, (const R2AJob *)c->r_jobs, c->r_packed, c->r_woff, idx->d_pac,
                                                                          (uint64_t)l_pac, c->d_q, c->d_t);
        ++c->launches;
        B200_CUDA(cudaGetLastError());
    }
    // ---- waves (the do-while of src/bwamem.c:2375-2392, all regions at once)
    std::vector<uint32_t> act, h_qoff, h_qlen, h_toff, h_tlen, h_w, h_nc, h_flat;
    std::vector<int32_t> h_score, h_nm; std::vector<uint64_t> h_off;
    for (int wave = 0; wave < 3; ++wave) {
        act.clear(); h_qoff.clear(); h_qlen.clear(); h_toff.clear(); h_tlen.clear(); h_w.clear();
        for (uint64_t jx = 0; jx < nj; ++jx) {
            St &s = st[job_aln[jx]];
            if (s.done) continue;
            s.w2 = s.w2 < opt_w << 2 ? s.w2 : opt_w << 2;
            const int lq = jobs[jx].qe - jobs[jx].qb; const int64_t rl = jobs[jx].re - jobs[jx].rb;
            // ungapped shortcut of bwa_gen_cigar2 (src/bwa.c:151-160) == the DP with a band of 0: one diagonal, all M
            const int band = (lq == rl && s.w2 == 0) ? 0 : bwa_b200_cigar_band(p, s.w2, lq, rl);
            act.push_back((uint32_t)jx); h_qoff.push_back(jobs[jx].qoff); h_qlen.push_back((uint32_t)lq); h_toff.push_back(jobs[jx].toff);
            h_tlen.push_back((uint32_t)rl); h_w.push_back((uint32_t)band);
        }
        // regions bwa_gen_cigar2 rejects go through the same control flow with score 0 and no CIGAR
        for (uint64_t k = 0; k < n_alns; ++k) {
            St &s = st[k];
            if (s.done || s.job) continue;
            s.w2 = s.w2 < opt_w << 2 ? s.w2 : opt_w << 2;
            ++s.n_waves;
            if (s.score == s.last_sc || s.w2 == opt_w << 2) { s.done = true; continue; }
            s.last_sc = s.score; s.w2 <<= 1;
            if (!(++s.iter < 3 && s.score < alns[k].truesc - a_sc)) s.done = true;
        }
        const uint64_t na = act.size();
        if (na == 0) break;
        B200_CUDA(cudaMemcpyAsync(c->d_qoff, h_qoff.data(), na * 4, cudaMemcpyHostToDevice, stq));
        B200_CUDA(cudaMemcpyAsync(c->d_qlen, h_qlen.data(), na * 4, cudaMemcpyHostToDevice, stq));
        B200_CUDA(cudaMemcpyAsync(c->d_toff, h_toff.data(), na * 4, cudaMemcpyHostToDevice, stq));
        B200_CUDA(cudaMemcpyAsync(c->d_tlen, h_tlen.data(), na * 4, cudaMemcpyHostToDevice, stq));
        B200_CUDA(cudaMemcpyAsync(c->d_w, h_w.data(), na * 4, cudaMemcpyHostToDevice, stq));
        int rc = cigar_run(c, p, na, c->d_q, c->d_qoff, c->d_qlen, c->d_t, c->d_toff, c->d_tlen, c->d_w, h_qlen.data(), h_tlen.data(), h_w.data(), true);
        if (rc) return rc;
        const uint64_t ops = c->last_ops;
        h_score.resize(na); h_nm.resize(na); h_nc.resize(na); h_off.resize(na); h_flat.resize(ops ? ops : 1);
        B200_CUDA(cudaMemcpyAsync(h_score.data(), c->d_score, na * 4, cudaMemcpyDeviceToHost, stq));
        B200_CUDA(cudaMemcpyAsync(h_nm.data(), c->d_nm, na * 4, cudaMemcpyDeviceToHost, stq));
        B200_CUDA(cudaMemcpyAsync(h_nc.data(), c->d_ncig, na * 4, cudaMemcpyDeviceToHost, stq));
        B200_CUDA(cudaMemcpyAsync(h_off.data(), c->d_off, na * 8, cudaMemcpyDeviceToHost, stq));
        if (ops) B200_CUDA(cudaMemcpyAsync(h_flat.data(), c->d_flat, ops * 4, cudaMemcpyDeviceToHost, stq));
        B200_CUDA(cudaStreamSynchronize(stq));
        for (uint64_t x = 0; x < na; ++x) {
            const uint64_t k = job_aln[act[x]];
            St &s = st[k];
            s.score = h_score[x]; s.nm = h_nm[x]; ++s.n_waves;
            s.cig.assign(h_flat.begin() + h_off[x], h_flat.begin() + h_off[x] + h_nc[x]);
            if (s.score == s.last_sc || s.w2 == opt_w << 2) { s.done = true; continue; }
            s.last_sc = s.score; s.w2 <<= 1;
            if (!(++s.iter < 3 && s.score < alns[k].truesc - a_sc)) s.done = true;
        }
    }
    // ---- position, deletion squeeze, soft clips (src/bwamem.c:2402-2432)
    uint64_t total = 0;
    for (uint64_t k = 0; k < n_alns; ++k) {
        const bwa_b200_aln_in_t &r = alns[k];
        St &s = st[k];
        bwa_b200_aln_out_t &o = out[k];
        memset(&o, 0, sizeof(o));
        if (r.rb < 0 || r.re < 0) { o.rid = -1; o.pos = -1; o.cigar_off = total; continue; }
        int64_t pos = r.rb < l_pac ? r.rb : r.re - 1;
        const int is_rev = pos >= l_pac;
        if (is_rev) pos = (l_pac << 1) - 1 - pos;
        std::vector<uint32_t> &cg = s.cig;
        if (!cg.empty()) {
            if ((cg.front() & 0xf) == 2) { pos += cg.front() >> 4; cg.erase(cg.begin()); }
            else if ((cg.back() & 0xf) == 2) cg.pop_back();
        }
        const int l_query = (int)read_len[r.read];
        if (r.qb != 0 || r.qe != l_query) {
            const int clip5 = is_rev ? l_query - r.qe : r.qb, clip3 = is_rev ? r.qb : l_query - r.qe;
            if (clip5) cg.insert(cg.begin(), (uint32_t)clip5 << 4 | 3u);
            if (clip3) cg.push_back((uint32_t)clip3 << 4 | 3u);
        }
        int rid = -1;
        if (pos < l_pac) {               // bns_pos2rid, src/bntseq.c:349-363
            int left = 0, mid = 0, right = n_ctg;
            while (left < right) {
                mid = (left + right) >> 1;
                if (pos >= ctg_off[mid]) { if (mid == n_ctg - 1) break; if (pos < ctg_off[mid + 1]) break; left = mid + 1; }
                else right = mid;
            }
            rid = mid;
        }
        o.rid = rid; o.pos = pos - (rid >= 0 ? ctg_off[rid] : 0); o.is_rev = is_rev; o.score = s.score; o.nm = s.nm;
        o.n_cigar = (int32_t)cg.size(); o.band = s.w2; o.n_waves = s.n_waves; o.cigar_off = total;
        total += cg.size();
    }
    uint32_t *flat = (uint32_t *)malloc((total ? total : 1) * 4);
    if (!flat) { b200::set_error("reg2aln: out of host memory"); return BWA_B200_ERR_NOMEM; }
    for (uint64_t k = 0; k < n_alns; ++k)
        if (!st[k].cig.empty() && out[k].rid != -1) memcpy(flat + out[k].cigar_off, st[k].cig.data(), st[k].cig.size() * 4);
    *cigar = flat; *n_ops = total;
    return BWA_B200_OK;
}
