// extend.cu -- ksw_extend2-equivalent banded seed extension, sm_100a.
//
// Parity target: ksw_extend2 (src/ksw.c:864-986; stock bwa_index/ksw.c:380-479): affine gaps,
// band, z-drop, end bonus; outputs max score, qle, tle, gtle, gscore, max_off -- bit-exact,
// including the reference's row-window semantics: row i evaluates columns [beg_i, end_i) only,
// the window is re-derived from zeros of the previous row, and state outside the window keeps
// whatever an earlier row (or the initial row) left there.  Those semantics are row-sequential,
// so the inter-query kernel below keeps rows in order per job and never speculates across rows.
//
//   ext_pair_kernel    one job per lane, two query columns per s16x2 register (ext_pair_core.cuh):
//                      M, E and the gap-open terms of a row are data-parallel along the row, only F
//                      carries from column to column, so a column pair costs ~11 packed ALU
//                      instructions (VIADDMNMX / VIMNMX3 / PRMT) instead of ~2 x 15 scalar ones.
//                      State {H(i-1,j-1), E(i,j)} of a column pair is one uint2 in shared memory,
//                      laid out [slot][lane]; with a band the slots are a ring of w + 2 pairs (the state is
//                      sized by the band, not by the query).  Takes every job whose score bound
//                      h0 + qlen * max(mat) is at most 1023 and whose query is at most 512 bases.
//   ext_inter_kernel   one job per lane, one column per step in 32-bit arithmetic: the general
//                      fallback (scores up to 2^15, queries up to 1024, any matrix).  The per-column
//                      state lives in shared memory as one 32-bit word per column, [column][lane];
//                      the query is staged next to it as 4-bit codes.
//   ext_intra_kernel   one job per WARP, for everything the two kernels above do not take (queries longer than
//                      1024 bases, scores that do not fit 16 bits): the 32 lanes compute 32 consecutive columns
//                      of a row at once.  M and E of a row depend on the previous row only; F is the max-plus
//                      prefix scan F(j+1) = max(F(j) - e_ins, max(M(j) - oe_ins, 0)), done with warp shuffles
//                      (5 steps per 32 columns) and carried from chunk to chunk.  int32 throughout, {H, E} per
//                      column in a per-warp slab in global memory (coalesced 256-byte rows), so there is no
//                      length limit other than the slab size (BWA_B200_EXT_INTRA_MAX_Q, default 16384 bases).
//   Jobs are sorted by (class, query length) on the device and launched per length bin, so lanes
//   of a warp run similar trip counts and each bin gets exactly the shared memory its longest
//   query needs.
//
// This is not a port of GASAL2's KSW kernel (one thread per alignment with eh[] in local
// memory, no band, zdrop = 0, fixed clip penalty): band, z-drop, end bonus and separate
// insertion/deletion penalties are honoured, and all six outputs are returned.
#include "internal.h"
#include "ext_pair_core.cuh"
#include <cub/cub.cuh>
#include <algorithm>

namespace {

#include "ext_wave.cuh"

constexpr int N_BINS = 7;
__constant__ int c_bin_hi[N_BINS] = {16, 32, 64, 128, 256, 512, 1024};   // max qlen of each bin

constexpr int N_PBINS = 16;              // bins of the column-pair kernel (PAIR_MAX_Q = the last one)
__constant__ int c_pbin_hi[N_PBINS] = {16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 448, 512};
constexpr uint32_t CLS_BIT = 1u << 19;   // key bits 19-20 = the job's class: 0 column-pair kernel, 1 32-bit per-lane kernel, 2 ext_wave_kernel
constexpr uint32_t CLS_WAVE = 2u << 19;  // (one job per warp, s16x2), 3 ext_intra_kernel (one job per warp, int32, no band needed)
constexpr uint32_t CLS_INTRA = 3u << 19;

constexpr uint32_t CLS_DONE = 4u << 19;  // answered by key_kernel in closed form (closed_form_job): sorted behind every kernel's range

// sort key: query length | class << 19.  Class 0: eligible for the column-pair s16x2 kernel (score bound at most 1023, query at most
// 512, eligible matrix / penalties); 2: ext_wave_kernel (banded batch, query of WAVE_MIN_Q + 1 .. 65535 bases, score bound at most
// WAVE_MAX_SCORE); 1: the 32-bit per-lane kernel (query at most 1024, scores below 2^15); 3: ext_intra_kernel (everything else);
// 4: the job's result is written here (counted in counters[1]) and no kernel sees it
template <bool BYTES>
__global__ void key_kernel(uint32_t n, JobView J, ClosedParams C, int max_score, int simd_ok, int wave_ok, uint32_t *keys, uint32_t *vals,
                           bwa_b200_ext_result_t *__restrict__ res, unsigned long long *__restrict__ counters)
{
    const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31u;
    // closed-form jobs (closed_form_job, ext_pair_core.cuh).  The shape test is per lane; the shifted diagonals of the jobs that need
    // them are tested by the whole warp, one job at a time, a diagonal per lane: a lane on its own would scan 8 .. 240 of them while
    // the lanes whose jobs need none wait (0.38 ms for C2's batch that way).
    ClosedShape S;
    int st = 0;
    if (a < n && C.ok) st = closed_form_shape<BYTES>(C, J, a, S);
    uint32_t pend = __ballot_sync(0xffffffffu, st == 2);
    while (pend) {
        const int src = __ffs(pend) - 1;
        pend &= pend - 1;
        const uint32_t a_s = __shfl_sync(0xffffffffu, a, src);
        const int k_s = __shfl_sync(0xffffffffu, S.k, src);
        int32_t rr_s[CF_KMAX];
#pragma unroll
        for (int m = 0; m < CF_KMAX; ++m) rr_s[m] = __shfl_sync(0xffffffffu, S.rr[m], src);
        const int ql_s = (int)J.qlen[a_s], n_checks = closed_form_checks(C, k_s);
        bool ok = true;
        for (int x = (int)lane; x < n_checks && ok; x += 32) ok = closed_form_check<BYTES>(C, J, a_s, ql_s, k_s, rr_s, x);
        const bool all_ok = __all_sync(0xffffffffu, ok);
        if ((int)lane == src) st = all_ok ? 1 : 0;
    }
    const bool done = st == 1;
    if (a < n) {
        uint32_t q = J.qlen[a];
        uint64_t bound = (uint64_t)J.h0[a] + (uint64_t)q * (uint64_t)(max_score > 0 ? max_score : 0);
        uint32_t k = q > 0x7ffffu ? 0x7ffffu : q;
        if (done) { bwa_b200_ext_result_t r; closed_form_result(C, (int)J.h0[a], (int)q, S, &r); res[a] = r; k |= CLS_DONE; }
        else if (simd_ok && bound <= (uint64_t)PAIR_MAX_SCORE && q <= (uint32_t)PAIR_MAX_Q) { }
        else if (wave_ok && q > (uint32_t)WAVE_MIN_Q && q <= 0xffffu && bound <= (uint64_t)WAVE_MAX_SCORE) k |= CLS_WAVE;
        else if (q <= 1024u && bound < 32767ull) k |= CLS_BIT;
        else k |= CLS_INTRA;
        keys[a] = k;
        vals[a] = a;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, done);
    if (m && lane == 0) atomicAdd(counters + 1, (unsigned long long)__popc(m));
}

// Sorted positions of the bin boundaries.  range[0..N_PBINS] bound the bins of the column-pair kernel (bin b =
// [range[b], range[b+1])), range[N_PBINS+1 .. N_PBINS+1+N_BINS] those of the 32-bit kernel.
__device__ __forceinline__ uint32_t lower_bound_key(const uint32_t *sorted_keys, uint32_t n, uint32_t thr)
{ // first position whose key is >= thr
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (sorted_keys[mid] < thr) lo = mid + 1; else hi = mid; }
    return lo;
}
__global__ void range_kernel(uint32_t n, const uint32_t *sorted_keys, uint32_t *range, int *err_flag, int intra_ok)
{
    const int k = threadIdx.x;
    if (k > N_PBINS + N_BINS + 3) return;
    uint32_t thr;
    if (k <= N_PBINS) thr = k == 0 ? 0u : (uint32_t)c_pbin_hi[k - 1] + 1u;
    else if (k <= N_PBINS + N_BINS + 1) thr = CLS_BIT | (k == N_PBINS + 1 ? 0u : (uint32_t)c_bin_hi[k - N_PBINS - 2] + 1u);
    else thr = k == N_PBINS + N_BINS + 2 ? CLS_INTRA : CLS_DONE;          // the last slot: where the jobs no kernel takes begin
    const uint32_t lo = lower_bound_key(sorted_keys, n, thr);
    range[k] = lo;
    // class-0 keys beyond the last pair bin cannot exist (key_kernel); if one does, it is reported, not dropped
    if (k == N_PBINS && lo < lower_bound_key(sorted_keys, n, CLS_BIT)) atomicExch(err_flag, 2);
    // keys at or beyond range[N_PBINS + N_BINS + 1] run one per warp: ext_wave_kernel up to range[N_PBINS + N_BINS + 2] (class 2 exists only
    // when that kernel is launched), ext_intra_kernel beyond; a caller that ruled the latter out (no slabs, no launch) and was wrong gets
    // an error, not missing results
    if (k == N_PBINS + N_BINS + 2 && !intra_ok && lo < lower_bound_key(sorted_keys, n, CLS_DONE)) atomicExch(err_flag, 3);
}

// One DP cell.  State word p = H(i-1,j-1) | E(i,j) << 16.  The reference's `M = M ? M + s : 0`
// (src/ksw.c:924) is computed as min(H + s, H << 16): for H > 0 the second term is huge, for H == 0
// it is 0 and M <= 0, which every later use treats exactly like 0 (h = max(M,e,f) with e,f >= 0;
// t = max(M - oe, 0)).  Scores come from a row of the matrix kept as biased bytes (score + bias) in
// {mlo, mhi} and are picked with one PRMT whose selector nibble is the query code (the other
// selector nibbles are neighbouring codes, all < 8, and their bytes are masked off).
#define EXT_CELL(J, PE)                                                                          \
    {                                                                                            \
        const uint32_t p_ = *(PE);                                                               \
        const int x_ = (int)(p_ << 16);                                                          \
        const int hh_ = (int)(p_ & 0xffffu), e_ = (int)(p_ >> 16);                               \
        qw = __funnelshift_l(qw, qw, 4);                                                         \
        const int scb_ = (int)(__byte_perm(mlo, mhi, qw) & 0xffu);                               \
        const int M_ = __viaddmin_s32(hh_ + scb_, nbias, x_);                                    \
        const int h_ = __vimax3_s32(M_, e_, f);                                                  \
        const uint32_t kj_ = ((uint32_t)h_ << 16) + (uint32_t)(J);                               \
        key = key > kj_ ? key : kj_;                                                             \
        const int t1_ = __viaddmax_s32(M_, noe_del, 0);                                          \
        const int en_ = __viaddmax_s32(e_, ne_del, t1_);                                         \
        const int t2_ = __viaddmax_s32(M_, noe_ins, 0);                                          \
        f = __viaddmax_s32(f, ne_ins, t2_);                                                      \
        *(PE) = (uint32_t)h1 + ((uint32_t)en_ << 16);                                            \
        h1 = h_;                                                                                 \
    }

template <bool BYTES, int NT>
__global__ void __launch_bounds__(NT)
ext_inter_kernel(ExtParams P, JobView J, const uint32_t *__restrict__ order, const uint32_t *__restrict__ range, int bin,
                 int max_q, bwa_b200_ext_result_t *__restrict__ res, unsigned long long *__restrict__ cells_total,
                 int *__restrict__ err_flag)
{
    extern __shared__ uint32_t smem[];
    __shared__ uint32_t smat[10];                          // biased matrix rows: [t] = bytes q0..q3, [5 + t] = byte q4
    const int tid = threadIdx.x;
    uint32_t *const ehp = smem + tid;                      // eh[j] = ehp[j * NT]   (h | e << 16)
    uint32_t *const qsp = smem + (size_t)(max_q + 1) * NT + tid;   // 4-bit query codes, 8 per word
    const uint32_t lo = range[bin], hi = range[bin + 1];
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    const int noe_del = -oe_del, noe_ins = -oe_ins, ne_del = -P.e_del, ne_ins = -P.e_ins, nbias = -P.bias;
    unsigned long long my_cells = 0;
    if (tid < 5) {
        uint32_t w = 0;
        for (int q = 0; q < 4; ++q) w |= (uint32_t)(uint8_t)(P.mat[tid * 5 + q] + P.bias) << (8 * q);
        smat[tid] = w;
        smat[5 + tid] = (uint32_t)(uint8_t)(P.mat[tid * 5 + 4] + P.bias);
    }
    __syncthreads();

    for (uint32_t base = lo + blockIdx.x * NT; base < hi; base += gridDim.x * NT) {
        const uint32_t pos = base + tid;
        if (pos < hi) {
            const uint32_t a = order[hi - 1u - (pos - lo)];       // longest queries of the bin first: a short tail
            const int qlen = (int)J.qlen[a], tlen = (int)J.tlen[a], h0 = (int)J.h0[a];
            const uint32_t qo = J.qoff[a], to = J.toff[a];
            if (qlen == 0) {               // absent job (pipeline slots): ksw_extend2 is never called, score stays h0
                bwa_b200_ext_result_t r0; r0.score = h0; r0.qle = 0; r0.tle = 0; r0.gtle = 0; r0.gscore = -1; r0.max_off = 0;
                res[a] = r0;
                continue;
            }
            if (qlen > max_q || h0 < 1) { atomicExch(err_flag, 1); continue; }

            // stage the query
            for (int j8 = 0; j8 < qlen; j8 += 8) {
                uint32_t wv;
                if (BYTES) {
                    wv = 0;
                    for (int u = 0; u < 8; ++u) { uint32_t c = j8 + u < qlen ? J.qb[qo + j8 + u] : 4u; wv |= (c > 4u ? 4u : c) << (28 - 4 * u); }
                } else {
                    wv = J.qp[(qo + j8) >> 3];        // device-packed input: codes 0..4 (bwa_b200_pack_device)
                }
                qsp[(j8 >> 3) * NT] = wv;
            }
            // first row: H(-1,-1) = h0, then one gap open, then extensions (src/ksw.c:880-883)
            {
                int v = h0;
                ehp[0] = (uint32_t)v;
                v = h0 > oe_ins ? h0 - oe_ins : 0;
                for (int j = 1; j <= qlen; ++j) {
                    ehp[j * NT] = (uint32_t)v;
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
            for (int i = 0; i < tlen; ++i) {
                int tbv;
                if (BYTES) { tbv = J.tb[to + i]; tbv = tbv > 4 ? 4 : tbv; }
                else {
                    if ((i & 7) == 0) tword = J.tp[(to + i) >> 3];
                    tbv = (int)((tword >> (28 - 4 * (i & 7))) & 15u);
                    tbv = tbv > 4 ? 4 : tbv;
                }
                const uint32_t mlo = smat[tbv], mhi = smat[5 + tbv];
                if (P.use_band) {
                    if (beg < i - w) beg = i - w;
                    if (end > i + w + 1) end = i + w + 1;
                    if (end > qlen) end = qlen;
                }
                int h1 = 0, f = 0;
                if (beg == 0) { h1 = h0 - (P.o_del + P.e_del * (i + 1)); h1 = h1 < 0 ? 0 : h1; }
                uint32_t key = 0;                      // (row max << 16) + column, last column wins ties
                uint32_t qw = 0;
                int j = beg;
                uint32_t *pe = ehp + j * NT;
                // head: up to the next multiple of 8
                if ((j & 7) && j < end) {
                    qw = qsp[(j >> 3) * NT];
                    qw = __funnelshift_l(qw, qw, 4 * (j & 7));
                    const int stop = min(end, (j | 7) + 1);
                    for (; j < stop; ++j, pe += NT) EXT_CELL(j, pe)
                }
                // whole words of the query: 8 cells per iteration, no bounds checks
                for (; j + 8 <= end; j += 8, pe += 8 * NT) {
                    qw = qsp[(j >> 3) * NT];
#pragma unroll
                    for (int u = 0; u < 8; ++u) EXT_CELL(j + u, pe + u * NT)
                }
                // tail
                if (j < end) {
                    qw = qsp[(j >> 3) * NT];
                    for (; j < end; ++j, pe += NT) EXT_CELL(j, pe)
                }
                my_cells += (unsigned long long)(end > beg ? end - beg : 0);
                ehp[end * NT] = (uint32_t)h1;          // H(i, end-1); E = 0
                const int m = (int)(key >> 16), mj = beg < end ? (int)(key & 0xffffu) : -1;
                if (end == qlen && beg < end) {        // `j == qlen` after the column loop
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
                j = beg;
                while (j < end && ehp[j * NT] == 0u) ++j;
                beg = j;
                j = end;
                while (j >= beg && ehp[j * NT] == 0u) --j;
                end = j + 2 < qlen ? j + 2 : qlen;
            }
            bwa_b200_ext_result_t r;
            r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = best_ie + 1; r.gscore = gscore; r.max_off = max_off;
            res[a] = r;
        }
    }
    // one atomic per warp for the evaluated-cell counter
    for (int o = 16; o > 0; o >>= 1) my_cells += __shfl_down_sync(0xffffffffu, my_cells, o);
    if ((tid & 31) == 0 && my_cells) atomicAdd(cells_total, my_cells);
}

// ext_intra_kernel: one job per warp (see the header).  Row-order semantics are those of ext_inter_kernel; the row itself is
// evaluated 32 columns at a time.  All row-level state (beg, end, best, ...) is kept redundantly in every lane.
constexpr int INTRA_WARPS = 4;
template <bool BYTES>
__global__ void __launch_bounds__(INTRA_WARPS * 32)
ext_intra_kernel(ExtParams P, JobView J, const uint32_t *__restrict__ order, const uint32_t *__restrict__ first,
                 int max_q, int2 *__restrict__ slab_all, bwa_b200_ext_result_t *__restrict__ res,
                 unsigned long long *__restrict__ cells_total, int *__restrict__ err_flag)
{
    __shared__ int8_t smat[32];
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 25) smat[tid] = P.mat[tid];
    __syncthreads();
    const uint32_t gw = blockIdx.x * INTRA_WARPS + (tid >> 5), n_warps = gridDim.x * INTRA_WARPS;
    int2 *const eh = slab_all + (uint64_t)gw * (uint64_t)(max_q + 1);      // eh[j] = {H(i-1,j-1), E(i,j)}
    const int oe_del = P.o_del + P.e_del, oe_ins = P.o_ins + P.e_ins;
    unsigned long long my_cells = 0;
    const uint32_t n = first[1];                       // the jobs behind are answered already (CLS_DONE)
    for (uint32_t pos = *first + gw; pos < n; pos += n_warps) {
        const uint32_t a = order[pos];
        const int qlen = (int)J.qlen[a], tlen = (int)J.tlen[a], h0 = (int)J.h0[a];
        const uint32_t qo = J.qoff[a], to = J.toff[a];
        if (qlen == 0) {
            if (lane == 0) { bwa_b200_ext_result_t r0; r0.score = h0; r0.qle = 0; r0.tle = 0; r0.gtle = 0; r0.gscore = -1; r0.max_off = 0; res[a] = r0; }
            continue;
        }
        if (qlen > max_q || h0 < 1) { if (lane == 0) atomicExch(err_flag, 1); continue; }
        auto qcode = [&](int j) -> int {
            int c;
            if (BYTES) c = J.qb[qo + j];
            else c = (int)((J.qp[(qo + j) >> 3] >> (28 - 4 * (j & 7))) & 15u);
            return c > 4 ? 4 : c;
        };
        // first row (src/ksw.c:880-883): H(-1,-1) = h0, one gap open, then extensions, clipped at 0
        for (int j = lane; j <= qlen; j += 32) {
            long long v = j == 0 ? (long long)h0 : (long long)h0 - oe_ins - (long long)(j - 1) * P.e_ins;
            eh[j] = make_int2(v > 0 ? (int)v : 0, 0);
        }
        __syncwarp();
        int w = P.w;
        {   // band clamp (src/ksw.c:885-893)
            int max_ins = (int)((double)(qlen * P.max_score + P.end_bonus - P.o_ins) / P.e_ins + 1.);
            max_ins = max_ins > 1 ? max_ins : 1;
            w = w < max_ins ? w : max_ins;
            int max_del = (int)((double)(qlen * P.max_score + P.end_bonus - P.o_del) / P.e_del + 1.);
            max_del = max_del > 1 ? max_del : 1;
            w = w < max_del ? w : max_del;
        }
        int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
        int beg = 0, end = qlen;
        for (int i = 0; i < tlen; ++i) {
            int tbv;
            if (BYTES) tbv = J.tb[to + i];
            else tbv = (int)((J.tp[(to + i) >> 3] >> (28 - 4 * (i & 7))) & 15u);
            tbv = tbv > 4 ? 4 : tbv;
            const int8_t *srow = smat + tbv * 5;
            if (P.use_band) {
                if (beg < i - w) beg = i - w;
                if (end > i + w + 1) end = i + w + 1;
                if (end > qlen) end = qlen;
            }
            int hc = 0, fc = 0;                 // carries into the next chunk: H(i, j-1) and F(i, j) of its first column
            if (beg == 0) { hc = h0 - (P.o_del + P.e_del * (i + 1)); hc = hc < 0 ? 0 : hc; }
            unsigned long long key = 0;         // (row max << 32) + column, last column wins ties
            for (int base = beg; base < end; base += 32) {
                const int j = base + lane;
                const bool in = j < end;
                const int2 pe = in ? eh[j] : make_int2(0, 0);
                const int M = in && pe.x ? pe.x + srow[qcode(j)] : 0;           // M = M ? M + s : 0   (src/ksw.c:924)
                const int g = M - oe_ins > 0 ? M - oe_ins : 0;
                int pm = g + lane * P.e_ins;                                    // prefix max of g(k) + k * e_ins
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(0xffffffffu, pm, o); if (lane >= o) pm = pm > up ? pm : up; }
                const int pprev = __shfl_up_sync(0xffffffffu, pm, 1);
                int f = fc - lane * P.e_ins;
                if (lane > 0) { const int t = pprev - (lane - 1) * P.e_ins; f = f > t ? f : t; }
                int h = M > pe.y ? M : pe.y;
                h = h > f ? h : f;
                int t1 = M - oe_del; t1 = t1 > 0 ? t1 : 0;
                int e2 = pe.y - P.e_del; e2 = e2 > t1 ? e2 : t1;
                const int hup = __shfl_up_sync(0xffffffffu, h, 1);
                if (in) {
                    eh[j] = make_int2(lane == 0 ? hc : hup, e2);
                    const unsigned long long kj = ((unsigned long long)(uint32_t)h << 32) | (uint32_t)j;
                    key = key > kj ? key : kj;
                }
                const int last = end - 1 - base < 31 ? end - 1 - base : 31;        // last column of this chunk
                hc = __shfl_sync(0xffffffffu, h, last);
                const int p31 = __shfl_sync(0xffffffffu, pm, 31);
                const int c1 = p31 - 31 * P.e_ins, c2 = fc - 32 * P.e_ins;
                fc = c1 > c2 ? c1 : c2;
            }
            if (lane == 0) eh[end] = make_int2(hc, 0);       // H(i, end-1); E = 0
            __syncwarp();
#pragma unroll
            for (int o = 16; o; o >>= 1) { const unsigned long long ok = __shfl_xor_sync(0xffffffffu, key, o); key = key > ok ? key : ok; }
            if (lane == 0) my_cells += (unsigned long long)(end > beg ? end - beg : 0);
            const int m = (int)(key >> 32), mj = beg < end ? (int)(uint32_t)key : -1;
            if (end == qlen && beg < end) {        // `j == qlen` after the column loop
                best_ie = gscore > hc ? best_ie : i;
                gscore = gscore > hc ? gscore : hc;
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
            // shrink the window to the non-zero part of the row (src/ksw.c:965-970)
            int nb = end;
            for (int base = beg; base < end; base += 32) {
                const int j = base + lane;
                int2 v = make_int2(0, 0);
                if (j < end) v = eh[j];
                const uint32_t bal = __ballot_sync(0xffffffffu, (v.x | v.y) != 0);
                if (bal) { nb = base + __ffs(bal) - 1; break; }
            }
            int jj = nb - 1;
            for (int top = end; top >= nb; top -= 32) {
                const int j = top - lane;
                int2 v = make_int2(0, 0);
                if (j >= nb) v = eh[j];
                const uint32_t bal = __ballot_sync(0xffffffffu, (v.x | v.y) != 0);
                if (bal) { jj = top - (__ffs(bal) - 1); break; }
            }
            beg = nb;
            end = jj + 2 < qlen ? jj + 2 : qlen;
            __syncwarp();
        }
        if (lane == 0) {
            bwa_b200_ext_result_t r;
            r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = best_ie + 1; r.gscore = gscore; r.max_off = max_off;
            res[a] = r;
        }
        __syncwarp();
    }
    if (lane == 0 && my_cells) atomicAdd(cells_total, my_cells);
}

// ext_pair_kernel: one job per lane, two query columns per s16x2 register (ext_pair_core.cuh); n_slots = state slots per lane
// WIDE: the class of ext_wave_kernel (scores to 32700, any query length, ring state), one job per lane, taken when the batch holds at
// least min_jobs such jobs
template <bool BYTES, int NT, bool SAME_GAP, bool RING, bool CHUNKED, int U, bool WIDE = false>
__global__ void __launch_bounds__(NT)
ext_pair_kernel(ExtParams P, PairParams S, JobView J, const uint32_t *__restrict__ order, const uint32_t *__restrict__ range, int bin,
                int max_q, int n_slots, bwa_b200_ext_result_t *__restrict__ res, unsigned long long *__restrict__ cells_total,
                int *__restrict__ err_flag, uint32_t min_jobs = 0)
{
    extern __shared__ uint2 smem2[];
    __shared__ uint32_t stab[8];                                                                 // S.tab, indexed by the target base
    const int tid = threadIdx.x;
    if (tid < 5) stab[tid] = S.tab[tid];
    __syncthreads();
    uint2 *const HEp = smem2 + tid;                                                              // HE[s] = HEp[s * NT]
    uint16_t *const QSp = reinterpret_cast<uint16_t *>(smem2 + (size_t)n_slots * NT) + tid;      // QS[s] = QSp[s * NT]
    const uint32_t lo = range[bin], hi = range[bin + 1];
    if (hi - lo < min_jobs) return;
    unsigned long long my_cells = 0;
    for (uint32_t base = lo + blockIdx.x * NT; base < hi; base += gridDim.x * NT) {
        const uint32_t pos = base + tid;
        if (pos < hi) {
            const uint32_t a = order[hi - 1u - (pos - lo)];       // longest queries of the bin first: a short tail
            const int qlen = (int)J.qlen[a], tlen = (int)J.tlen[a], h0 = (int)J.h0[a];
            bwa_b200_ext_result_t r;
            if (qlen == 0) {               // absent job (pipeline slots): ksw_extend2 is never called, score stays h0
                r.score = h0; r.qle = 0; r.tle = 0; r.gtle = 0; r.gscore = -1; r.max_off = 0;
                res[a] = r;
                continue;
            }
            if (qlen > max_q || h0 < 1 || (WIDE && (long long)h0 + (long long)qlen * P.max_score > PAIR_WIDE_MAX_SCORE)) { atomicExch(err_flag, 1); continue; }
            pair_job<BYTES, NT, SAME_GAP, RING, CHUNKED, U, WIDE>(P, S, stab, J, a, qlen, tlen, h0, HEp, QSp, r, my_cells);
            res[a] = r;
        }
    }
    for (int o = 16; o > 0; o >>= 1) my_cells += __shfl_down_sync(0xffffffffu, my_cells, o);
    if ((tid & 31) == 0 && my_cells) atomicAdd(cells_total, my_cells);
}

// (score, qend, tend) after the local-vs-to-end rule (src/bwamem.c:1892-1901)
__global__ void triple_kernel(uint32_t n, const bwa_b200_ext_result_t *res, const uint32_t *qlen, int pen_clip,
                              int32_t *score, int32_t *qend, int32_t *tend)
{
    uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n) return;
    bwa_b200_ext_result_t r = res[a];
    bool local = r.gscore <= 0 || r.gscore <= r.score - pen_clip;
    score[a] = local ? r.score : r.gscore;
    qend[a] = local ? r.qle : (int32_t)qlen[a];
    tend[a] = local ? r.tle : r.gtle;
}

__global__ void pack_kernel(const uint8_t *__restrict__ bytes, uint64_t n_words, uint64_t n_bytes, uint32_t *__restrict__ packed)
{
    uint64_t wi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (wi >= n_words) return;
    uint32_t wv = 0;
    for (int u = 0; u < 8; ++u) {
        uint64_t p = wi * 8 + u;
        uint32_t c = p < n_bytes ? bytes[p] : 4u;
        wv |= (c > 4u ? 4u : c) << (28 - 4 * u);
    }
    packed[wi] = wv;
}

} // namespace

// the instantiations of ext_pair_kernel: blocks of 32 lanes only where a row can be longer than a chunk (the long, unbanded bins)
using pair_kern_t = void (*)(ExtParams, PairParams, JobView, const uint32_t *, const uint32_t *, int, int, int, bwa_b200_ext_result_t *,
                             unsigned long long *, int *, uint32_t);
template <bool BYTES, bool SG, bool RG, bool CH>
static pair_kern_t pair_kernel_pick(int nt, int u)
{
    if (nt == 32) {
        if (!CH) return nullptr;
        (void)u; return (pair_kern_t)ext_pair_kernel<BYTES, 32, SG, RG, true, 8>;
    }
    return (pair_kern_t)ext_pair_kernel<BYTES, 64, SG, RG, CH, 8>;
}
template <bool BYTES>
static const void *pair_kernel_ptr(int nt, bool sg, bool rg, bool ch, int u)
{
    pair_kern_t k;
    if (sg) k = rg ? (ch ? pair_kernel_pick<BYTES, true, true, true>(nt, u) : pair_kernel_pick<BYTES, true, true, false>(nt, u))
                   : (ch ? pair_kernel_pick<BYTES, true, false, true>(nt, u) : pair_kernel_pick<BYTES, true, false, false>(nt, u));
    else    k = rg ? (ch ? pair_kernel_pick<BYTES, false, true, true>(nt, u) : pair_kernel_pick<BYTES, false, true, false>(nt, u))
                   : (ch ? pair_kernel_pick<BYTES, false, false, true>(nt, u) : pair_kernel_pick<BYTES, false, false, false>(nt, u));
    return (const void *)k;
}

static int ext_grow_jobs(bwa_b200_extender *e, uint64_t n)
{
    if (n <= e->max_jobs) return BWA_B200_OK;
    uint64_t cap = n + n / 4 + 256;
    cudaFree(e->d_qoff); cudaFree(e->d_qlen); cudaFree(e->d_toff); cudaFree(e->d_tlen); cudaFree(e->d_h0);
    cudaFree(e->d_keys); cudaFree(e->d_keys2); cudaFree(e->d_vals); cudaFree(e->d_order); cudaFree(e->d_res); cudaFree(e->d_tri); cudaFree(e->d_cub);
    e->d_qoff = e->d_qlen = e->d_toff = e->d_tlen = e->d_h0 = e->d_keys = e->d_keys2 = e->d_vals = e->d_order = nullptr;
    e->d_res = nullptr; e->d_tri = nullptr; e->d_cub = nullptr; e->max_jobs = 0;
    B200_CUDA(cudaMalloc(&e->d_qoff, cap * 4)); B200_CUDA(cudaMalloc(&e->d_qlen, cap * 4));
    B200_CUDA(cudaMalloc(&e->d_toff, cap * 4)); B200_CUDA(cudaMalloc(&e->d_tlen, cap * 4));
    B200_CUDA(cudaMalloc(&e->d_h0, cap * 4));
    B200_CUDA(cudaMalloc(&e->d_keys, cap * 4)); B200_CUDA(cudaMalloc(&e->d_keys2, cap * 4));
    B200_CUDA(cudaMalloc(&e->d_vals, cap * 4)); B200_CUDA(cudaMalloc(&e->d_order, cap * 4));
    B200_CUDA(cudaMalloc(&e->d_res, cap * sizeof(bwa_b200_ext_result_t)));
    B200_CUDA(cudaMalloc(&e->d_tri, cap * 3 * 4));
    B200_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, e->cub_bytes, e->d_keys, e->d_keys2, e->d_vals, e->d_order, (int)cap, 0, 21, e->stream));
    B200_CUDA(cudaMalloc(&e->d_cub, e->cub_bytes + 16));
    e->max_jobs = cap;
    return BWA_B200_OK;
}

static int ext_grow_seq(bwa_b200_extender *e, uint64_t qb, uint64_t tb)
{
    if (qb > e->max_q) { cudaFree(e->d_q); e->d_q = nullptr; e->max_q = 0; uint64_t c = qb + qb / 4 + 64; B200_CUDA(cudaMalloc(&e->d_q, c)); e->max_q = c; }
    if (tb > e->max_t) { cudaFree(e->d_t); e->d_t = nullptr; e->max_t = 0; uint64_t c = tb + tb / 4 + 64; B200_CUDA(cudaMalloc(&e->d_t, c)); e->max_t = c; }
    return BWA_B200_OK;
}

extern "C" void bwa_b200_fill_scmat(int a, int b, int8_t mat[25])
{ // bwa_fill_scmat (src/bwa.c): match a, mismatch -b, anything against N -1
    int k = 0;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[k++] = (int8_t)(i == j ? a : -b); mat[k++] = -1; }
    for (int j = 0; j < 5; ++j) mat[k++] = -1;
}

extern "C" void bwa_b200_ext_params_default(bwa_b200_ext_params_t *p)
{
    if (!p) return;
    bwa_b200_fill_scmat(1, 4, p->mat);
    p->o_del = p->o_ins = 6; p->e_del = p->e_ins = 1;
    p->w = 100; p->end_bonus = 5; p->zdrop = 100; p->use_band = 1; p->pen_clip = 5;
}

extern "C" int bwa_b200_extender_create(int device, uint64_t max_jobs, uint64_t max_query_bytes, uint64_t max_target_bytes,
                                        bwa_b200_extender_t **out)
{
    if (!out) { b200::set_error("extender_create: bad argument"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(device));
    bwa_b200_extender *e = new bwa_b200_extender();
    e->device = device;
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, device));
    e->n_sm = prop.multiProcessorCount;
    e->smem_optin = (int)prop.sharedMemPerBlockOptin;
    B200_CUDA(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    e->n_side = getenv("BWA_B200_EXT_SIDE_STREAMS") ? atoi(getenv("BWA_B200_EXT_SIDE_STREAMS")) : 4;
    if (e->n_side < 0) e->n_side = 0;
    if (e->n_side > 8) e->n_side = 8;
    for (int k = 0; k < e->n_side; ++k) {
        B200_CUDA(cudaStreamCreateWithFlags(&e->side[k], cudaStreamNonBlocking));
        B200_CUDA(cudaEventCreateWithFlags(&e->ev_join[k], cudaEventDisableTiming));
    }
    B200_CUDA(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    B200_CUDA(cudaMalloc(&e->d_range, (N_PBINS + N_BINS + 4) * 4));
    B200_CUDA(cudaMalloc(&e->d_cells, 16));
    B200_CUDA(cudaMalloc(&e->d_err, 4));
    e->no_closed_form = getenv("BWA_B200_EXT_NO_CLOSED") != nullptr;
    e->intra_max_q = getenv("BWA_B200_EXT_INTRA_MAX_Q") ? atoi(getenv("BWA_B200_EXT_INTRA_MAX_Q")) : 16384;
    if (e->intra_max_q < 1024) e->intra_max_q = 1024;
    {   // a warp's row is one dependent chain of 32-column chunks (loads, scan, carries): with 4 warps per SM the kernel issued 17 % of
        // the time (ncu, profiles/r01_ncu_full_intra_v9.txt); 24 resident warps per SM hide most of that latency
        int occ = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ext_intra_kernel<true>, INTRA_WARPS * 32, 0));
        if (occ < 1) occ = 1;
        if (occ > 6) occ = 6;
        e->intra_grid = e->n_sm * occ;
    }
    // the slabs of ext_intra_kernel (0.47 GB on a 148-SM part) are allocated by the first batch that can reach that kernel (ext_launch)
    B200_CUDA(cudaMemset(e->d_cells, 0, 16));
    B200_CUDA(cudaMemset(e->d_err, 0, 4));
    B200_CUDA(cudaHostAlloc(&e->h_cells, 16, cudaHostAllocDefault));
    B200_CUDA(cudaHostAlloc(&e->h_err, 4, cudaHostAllocDefault));
    *e->h_cells = 0; *e->h_err = 0;
    {   // dynamic shared memory opt-in: device limit minus the kernels' static shared memory
        cudaFuncAttributes fa;
        B200_CUDA(cudaFuncGetAttributes(&fa, ext_inter_kernel<true, 128>));
        e->smem_optin -= (int)fa.sharedSizeBytes + 16;
        const void *ks[6] = {(const void *)ext_inter_kernel<true, 128>, (const void *)ext_inter_kernel<false, 128>,
                             (const void *)ext_inter_kernel<true, 64>, (const void *)ext_inter_kernel<false, 64>,
                             (const void *)ext_inter_kernel<true, 32>, (const void *)ext_inter_kernel<false, 32>};
        for (const void *k : ks) B200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, e->smem_optin));
        for (int v = 0; v < 64; ++v) {          // every instantiation (the unroll bit is a no-op: one unroll factor is shipped)
            const void *k = pair_kernel_ptr<true>(v & 1 ? 32 : 64, (v & 2) != 0, (v & 4) != 0, (v & 8) != 0, v & 16 ? 16 : 8);
            if (v & 32) k = pair_kernel_ptr<false>(v & 1 ? 32 : 64, (v & 2) != 0, (v & 4) != 0, (v & 8) != 0, v & 16 ? 16 : 8);
            if (k) B200_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, e->smem_optin));
        }
    }
    e->pair_no_ring = getenv("BWA_B200_PAIR_NO_RING") != nullptr;
    int rc = ext_grow_jobs(e, max_jobs ? max_jobs : 1024);
    if (rc) return rc;
    rc = ext_grow_seq(e, max_query_bytes ? max_query_bytes : 1024, max_target_bytes ? max_target_bytes : 1024);
    if (rc) return rc;
    *out = e;
    return BWA_B200_OK;
}

extern "C" void bwa_b200_extender_destroy(bwa_b200_extender_t *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    cudaStreamSynchronize(e->stream);
    cudaFree(e->d_q); cudaFree(e->d_t); cudaFree(e->d_qoff); cudaFree(e->d_qlen); cudaFree(e->d_toff); cudaFree(e->d_tlen);
    cudaFree(e->d_h0); cudaFree(e->d_keys); cudaFree(e->d_keys2); cudaFree(e->d_vals); cudaFree(e->d_order); cudaFree(e->d_range);
    cudaFree(e->d_res); cudaFree(e->d_tri); cudaFree(e->d_cells); cudaFree(e->d_err); cudaFree(e->d_cub); cudaFree(e->d_intra);
    cudaFreeHost(e->h_cells); cudaFreeHost(e->h_err);
    for (int k = 0; k < e->n_side; ++k) { cudaStreamSynchronize(e->side[k]); cudaStreamDestroy(e->side[k]); cudaEventDestroy(e->ev_join[k]); }
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->own_stream) cudaStreamDestroy(e->stream);
    delete e;
}

extern "C" void *bwa_b200_extender_stream(bwa_b200_extender_t *e) { return e ? (void *)e->stream : nullptr; }
extern "C" uint64_t bwa_b200_extender_launches(const bwa_b200_extender_t *e) { return e ? e->launches : 0; }

static void to_dev_params(const bwa_b200_ext_params_t *p, ExtParams *d)
{
    memset(d, 0, sizeof(*d));
    memcpy(d->mat, p->mat, 25);
    d->o_del = p->o_del; d->e_del = p->e_del; d->o_ins = p->o_ins; d->e_ins = p->e_ins;
    d->w = p->w; d->end_bonus = p->end_bonus; d->zdrop = p->zdrop; d->use_band = p->use_band; d->pen_clip = p->pen_clip;
    int mx = 0;
    for (int i = 0; i < 25; ++i) mx = mx > p->mat[i] ? mx : p->mat[i];
    d->max_score = mx;
    int mn = 0;
    for (int i = 0; i < 25; ++i) mn = mn < p->mat[i] ? mn : p->mat[i];
    d->bias = -mn;
}

// sort by query length, derive bin ranges, launch one kernel per bin (empty bins exit at once)
template <bool BYTES>
// may_intra: some job of the batch may be beyond the per-lane kernels (query longer than 1024 bases or scores beyond 16 bits); callers
// that know the batch cannot hold one (short reads) pass false and neither the launch nor the slab allocation happens.  may_wave:
// some job may be outside the column-pair class with a query beyond WAVE_MIN_Q bases (ext_wave_kernel takes those of a banded batch)
static int ext_launch(bwa_b200_extender *e, const bwa_b200_ext_params_t *p, uint32_t n, const JobView &J,
                      bwa_b200_ext_result_t *d_res, bool may_intra = true, bool may_wave = true)
{
    if (p->e_del <= 0 || p->e_ins <= 0) { b200::set_error("extend: gap extension penalties must be positive"); return BWA_B200_ERR_ARG; }
    ExtParams P;
    to_dev_params(p, &P);
    if (e->phase_prof) e->phase_prof->begin("ext_phase", e->stream);
    B200_CUDA(cudaMemsetAsync(e->d_cells, 0, 16, e->stream));
    if (e->prof) e->prof->begin("ext_sort", e->stream);
    // the column-pair s16x2 kernel takes any matrix with 0 < max <= 31 and one score for a query N
    // (bwa_fill_scmat, src/bwa.c, is of that form); any other matrix runs entirely in the 32-bit kernel
    PairParams S;
    memset(&S, 0, sizeof(S));
    int simd_ok = pair_params_from(p, &S);
    if (getenv("BWA_B200_EXT_NO_SIMD")) simd_ok = 0;
    // ext_wave_kernel: banded batches whose ring (w + 2 pairs per job, WAVE_WARPS jobs per block) fits shared memory
    PairParams SW = S;
    bool wave_ok = simd_ok && may_wave && p->use_band && p->w >= 0 && p->w + 2 + WAVE_AHEAD <= WAVE_MAX_RING && !getenv("BWA_B200_EXT_NO_WAVE");
    size_t wave_smem = 0;
    if (wave_ok) {
        SW.ring = p->w + 2 + WAVE_AHEAD; SW.ring_magic = pair_ring_magic(SW.ring);
        wave_smem = (size_t)WAVE_WARPS * SW.ring * 10 + 16;
        if (wave_smem > (size_t)e->smem_optin) wave_ok = false;
    }
    ClosedParams CF = closed_params_from(p);
    if (e->no_closed_form) CF.ok = 0;
    key_kernel<BYTES><<<(n + 255) / 256, 256, 0, e->stream>>>(n, J, CF, P.max_score, simd_ok, wave_ok ? 1 : 0, e->d_keys, e->d_vals, d_res, e->d_cells);
    size_t tmp = e->cub_bytes;
    B200_CUDA(cub::DeviceRadixSort::SortPairs(e->d_cub, tmp, e->d_keys, e->d_keys2, e->d_vals, e->d_order, (int)n, 0, 22, e->stream));
    range_kernel<<<1, 32, 0, e->stream>>>(n, e->d_keys2, e->d_range, e->d_err, may_intra ? 1 : 0);
    if (e->prof) e->prof->end(e->stream);
    e->launches += 3;
    static const int bin_hi[N_BINS] = {16, 32, 64, 128, 256, 512, 1024};
    static const char *bin_name[N_BINS] = {"ext_inter_kernel_q16", "ext_inter_kernel_q32", "ext_inter_kernel_q64", "ext_inter_kernel_q128",
                                           "ext_inter_kernel_q256", "ext_inter_kernel_q512", "ext_inter_kernel_q1024"};
    static const int pbin_hi[N_PBINS] = {16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 448, 512};
    static const char *pbin_name[N_PBINS] = {"ext_pair_kernel_q16", "ext_pair_kernel_q32", "ext_pair_kernel_q48", "ext_pair_kernel_q64",
                                             "ext_pair_kernel_q80", "ext_pair_kernel_q96", "ext_pair_kernel_q112", "ext_pair_kernel_q128",
                                             "ext_pair_kernel_q160", "ext_pair_kernel_q192", "ext_pair_kernel_q224", "ext_pair_kernel_q256",
                                             "ext_pair_kernel_q320", "ext_pair_kernel_q384", "ext_pair_kernel_q448", "ext_pair_kernel_q512"};
    // Bins are independent: outside profiling they go to side streams (forked from and joined to e->stream with
    // events), so that the tail of one bin -- a few long jobs on a few SMs -- overlaps the next bins.
    const bool fan = e->prof == nullptr && e->n_side > 0;
    int next_side = 0;
    if (fan) {
        B200_CUDA(cudaEventRecord(e->ev_fork, e->stream));
        for (int k = 0; k < e->n_side; ++k) B200_CUDA(cudaStreamWaitEvent(e->side[k], e->ev_fork, 0));
    }
    auto bin_stream = [&]() -> cudaStream_t { if (!fan) return e->stream; cudaStream_t st = e->side[next_side]; next_side = (next_side + 1) % e->n_side; return st; };
    // column-pair s16x2 kernel, longest bins first
    const bool same_gap = pair_same_gap(p);
    for (int b = N_PBINS - 1; b >= 0 && simd_ok; --b) {
        const int L = pbin_hi[b];
        // The column state of a lane lives in shared memory: with a band it is a ring of w + 2 column pairs whatever the query
        // length, otherwise one slot per pair of the bin's longest query, which then bounds the resident lanes per SM.  Blocks of
        // 64 lanes waste up to 63 lanes' worth of it on the long bins; there, blocks of 32 lanes fit more lanes (q384: 96 vs 64).
        bwa_b200_ext_params_t pr = *p;
        if (e->pair_no_ring) pr.use_band = 0;                    // sizing only: the kernel still applies the band
        const int n_slots = pair_slots(&pr, L, &S);
        const size_t per_lane = (size_t)n_slots * 10;
        if (per_lane * 32 > (size_t)e->smem_optin) { b200::set_error("extend: query bin %d does not fit shared memory", L); return BWA_B200_ERR_CAPACITY; }
        const bool ring = S.ring != 0, chunked = n_slots > PAIR_CHUNK + 1;      // a row of at most 64 pairs needs no chunk fold
        const pair_kern_t k64 = (pair_kern_t)pair_kernel_ptr<BYTES>(64, same_gap, ring, chunked, e->pair_unroll);
        const pair_kern_t k32 = (pair_kern_t)pair_kernel_ptr<BYTES>(32, same_gap, ring, chunked, e->pair_unroll);
        int occ64 = 0, occ32 = 0;
        if (per_lane * 64 <= (size_t)e->smem_optin) B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, k64, 64, per_lane * 64));
        if (k32) B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, k32, 32, per_lane * 32));
        const bool use32 = occ32 * 32 > occ64 * 64;
        const int nt = use32 ? 32 : 64;
        int occ = use32 ? occ32 : occ64;
        if (occ < 1) occ = 1;
        const size_t smem = per_lane * nt;
        uint32_t max_blocks = (n + nt - 1) / nt;
        uint32_t grid = (uint32_t)(e->n_sm * occ);
        if (grid > max_blocks) grid = max_blocks;
        if (grid < 1) grid = 1;
        cudaStream_t st = bin_stream();
        B200_LAUNCH(e->prof, pbin_name[b], st,
            ((use32 ? k32 : k64)<<<grid, nt, smem, st>>>(P, S, J, e->d_order, e->d_range, b, L, n_slots, d_res, e->d_cells, e->d_err, 0u)));
        e->launches += 1;
    }
    // 32-bit kernel for everything else
    for (int b = 0; b < N_BINS; ++b) {
        const int L = bin_hi[b];
        const size_t per_thread = ((size_t)(L + 1) + (size_t)(L + 7) / 8) * 4;
        const int nt = L <= 128 ? 128 : (L <= 256 ? 64 : 32);
        size_t smem = per_thread * nt;
        if (smem > (size_t)e->smem_optin) { b200::set_error("extend: query bin %d does not fit shared memory", L); return BWA_B200_ERR_CAPACITY; }
        auto kern = nt == 128 ? ext_inter_kernel<BYTES, 128> : (nt == 64 ? ext_inter_kernel<BYTES, 64> : ext_inter_kernel<BYTES, 32>);
        int occ = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nt, smem));
        if (occ < 1) occ = 1;
        uint32_t max_blocks = (n + nt - 1) / nt;
        uint32_t grid = (uint32_t)(e->n_sm * occ);
        if (grid > max_blocks) grid = max_blocks;
        if (grid < 1) grid = 1;
        cudaStream_t st = bin_stream();
        B200_LAUNCH(e->prof, bin_name[b], st,
            (kern<<<grid, nt, smem, st>>>(P, J, e->d_order, e->d_range + (N_PBINS + 1), b, L, d_res, e->d_cells, e->d_err)));
        e->launches += 1;
    }
    if (wave_ok) {
        // The jobs of a banded batch outside the column-pair class (long queries, scores beyond 1023).  A batch with enough of them to
        // fill the machine one job per lane runs ext_pair_kernel<WIDE> (ring of w + 2 pairs per lane); a smaller one runs one job per
        // warp in ext_wave_kernel.  Both are launched; the count decides on the device which of the two returns at once.
        const bool sg = pair_same_gap(p);
        uint32_t wide_min = 0xffffffffu;                    // no wide launch: ext_wave_kernel takes any count
        {
            PairParams SP = S;
            SP.ring = p->w + 2; SP.ring_magic = pair_ring_magic(SP.ring);
            const size_t smem = (size_t)SP.ring * 10 * 32;
            if (smem <= (size_t)e->smem_optin && !getenv("BWA_B200_EXT_NO_WIDE")) {
                pair_kern_t kern = sg ? (pair_kern_t)ext_pair_kernel<BYTES, 32, true, true, false, 8, true> : (pair_kern_t)ext_pair_kernel<BYTES, 32, false, true, false, 8, true>;
                B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, e->smem_optin));
                int occ = 0;
                B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32, smem));
                if (occ < 1) occ = 1;
                wide_min = (uint32_t)std::max(1024, e->n_sm * occ * 32 / 4);       // a quarter of the resident lanes: where the two kernels break even
                if (getenv("BWA_B200_EXT_WIDE_MIN")) wide_min = (uint32_t)std::max(1, atoi(getenv("BWA_B200_EXT_WIDE_MIN")));   // tests: force either kernel
                uint32_t grid = (uint32_t)(e->n_sm * occ), max_blocks = (n + 31) / 32;
                if (grid > max_blocks) grid = max_blocks;
                cudaStream_t st = bin_stream();
                B200_LAUNCH(e->prof, "ext_pair_kernel_wide", st,
                    (kern<<<grid, 32, smem, st>>>(P, SP, J, e->d_order, e->d_range, N_PBINS + 1 + N_BINS, PAIR_WIDE_MAX_Q, SP.ring, d_res, e->d_cells, e->d_err, wide_min)));
                e->launches += 1;
            }
        }
        auto kern = sg ? ext_wave_kernel<BYTES, true> : ext_wave_kernel<BYTES, false>;
        B200_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wave_smem));
        int occ = 0;
        B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WAVE_WARPS * 32, wave_smem));
        if (occ < 1) occ = 1;
        if (occ > 4) occ = 4;
        uint32_t grid = (uint32_t)(e->n_sm * occ), max_blocks = (n + WAVE_WARPS - 1) / WAVE_WARPS;
        if (grid > max_blocks) grid = max_blocks;
        cudaStream_t st = bin_stream();
        B200_LAUNCH(e->prof, "ext_wave_kernel", st,
            (kern<<<grid, WAVE_WARPS * 32, wave_smem, st>>>(P, SW, J, e->d_order, e->d_range + (N_PBINS + 1 + N_BINS), wide_min, d_res, e->d_cells, e->d_err)));
        e->launches += 1;
    }
    if (may_intra && !e->d_intra)
        B200_CUDA(cudaMalloc(&e->d_intra, (size_t)e->intra_grid * INTRA_WARPS * ((size_t)e->intra_max_q + 1) * sizeof(int2)));
    if (may_intra) {   // one warp per job in int32 for what is left (no band, scores beyond 16 bits); exits at once when there is none
        cudaStream_t st = bin_stream();
        B200_LAUNCH(e->prof, "ext_intra_kernel", st,
            (ext_intra_kernel<BYTES><<<e->intra_grid, INTRA_WARPS * 32, 0, st>>>(P, J, e->d_order, e->d_range + (N_PBINS + N_BINS + 2), e->intra_max_q,
                                                                                e->d_intra, d_res, e->d_cells, e->d_err)));
        e->launches += 1;
    }
    if (fan)
        for (int k = 0; k < e->n_side; ++k) {
            B200_CUDA(cudaEventRecord(e->ev_join[k], e->side[k]));
            B200_CUDA(cudaStreamWaitEvent(e->stream, e->ev_join[k], 0));
        }
    if (e->phase_prof) e->phase_prof->end(e->stream);
    B200_CUDA(cudaGetLastError());
    return BWA_B200_OK;
}

// host pages -> device, kernels, results -> host; everything asynchronous on e->stream
extern "C" int bwa_b200_extend_async_paged(bwa_b200_extender_t *e, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                                           const bwa_b200_host_page_t *qpages, int n_qpages, uint64_t q_bytes,
                                           const uint32_t *qoff, const uint32_t *qlen,
                                           const bwa_b200_host_page_t *tpages, int n_tpages, uint64_t t_bytes,
                                           const uint32_t *toff, const uint32_t *tlen,
                                           const uint32_t *h0, bwa_b200_ext_result_t *res6,
                                           int32_t *aln_score, int32_t *query_end, int32_t *target_end)
{
    if (!e || !p || !qpages || !tpages || !qoff || !qlen || !toff || !tlen || !h0) { b200::set_error("extend_async: null argument"); return BWA_B200_ERR_ARG; }
    if (n_jobs == 0) { b200::set_error("extend_async: n_jobs == 0"); return BWA_B200_ERR_ARG; }      // gasal_align.cu:32-35
    if (q_bytes == 0 || t_bytes == 0) { b200::set_error("extend_async: empty batch"); return BWA_B200_ERR_ARG; }
    if (n_jobs >= 0x7fffffffull) { b200::set_error("extend_async: too many jobs"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(e->device));
    int rc = ext_grow_jobs(e, n_jobs);
    if (rc) return rc;
    rc = ext_grow_seq(e, q_bytes, t_bytes);
    if (rc) return rc;
    const uint32_t n = (uint32_t)n_jobs;
    cudaStream_t st = e->stream;
    for (int i = 0; i < n_qpages; ++i) {
        if (!qpages[i].bytes) continue;
        if (qpages[i].offset + qpages[i].bytes > q_bytes) { b200::set_error("extend_async: query page %d exceeds the batch", i); return BWA_B200_ERR_ARG; }
        B200_CUDA(cudaMemcpyAsync(e->d_q + qpages[i].offset, qpages[i].data, qpages[i].bytes, cudaMemcpyHostToDevice, st));
    }
    for (int i = 0; i < n_tpages; ++i) {
        if (!tpages[i].bytes) continue;
        if (tpages[i].offset + tpages[i].bytes > t_bytes) { b200::set_error("extend_async: target page %d exceeds the batch", i); return BWA_B200_ERR_ARG; }
        B200_CUDA(cudaMemcpyAsync(e->d_t + tpages[i].offset, tpages[i].data, tpages[i].bytes, cudaMemcpyHostToDevice, st));
    }
    B200_CUDA(cudaMemcpyAsync(e->d_qoff, qoff, n * 4ull, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(e->d_qlen, qlen, n * 4ull, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(e->d_toff, toff, n * 4ull, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(e->d_tlen, tlen, n * 4ull, cudaMemcpyHostToDevice, st));
    B200_CUDA(cudaMemcpyAsync(e->d_h0, h0, n * 4ull, cudaMemcpyHostToDevice, st));
    JobView J{e->d_q, e->d_t, nullptr, nullptr, e->d_qoff, e->d_qlen, e->d_toff, e->d_tlen, e->d_h0};
    bool may_intra = false, may_wave = false;  // host arrays: the exact tests of key_kernel / range_kernel
    {
        int mx = 0;
        for (int i = 0; i < 25; ++i) mx = mx > p->mat[i] ? mx : p->mat[i];
        for (uint32_t a = 0; a < n && !(may_intra && may_wave); ++a) {
            const uint64_t bound = (uint64_t)h0[a] + (uint64_t)qlen[a] * (uint64_t)mx;
            may_intra |= qlen[a] > 1024u || bound >= 32767ull;
            may_wave |= qlen[a] > (uint32_t)WAVE_MIN_Q && (qlen[a] > (uint32_t)PAIR_MAX_Q || bound > (uint64_t)PAIR_MAX_SCORE);
        }
    }
    rc = ext_launch<true>(e, p, n, J, e->d_res, may_intra, may_wave);
    if (rc) return rc;
    if (res6) B200_CUDA(cudaMemcpyAsync(res6, e->d_res, n * sizeof(bwa_b200_ext_result_t), cudaMemcpyDeviceToHost, st));
    if (aln_score || query_end || target_end) {
        triple_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, e->d_res, e->d_qlen, p->pen_clip, e->d_tri, e->d_tri + e->max_jobs, e->d_tri + 2 * e->max_jobs);
        e->launches += 1;
        if (aln_score) B200_CUDA(cudaMemcpyAsync(aln_score, e->d_tri, n * 4ull, cudaMemcpyDeviceToHost, st));
        if (query_end) B200_CUDA(cudaMemcpyAsync(query_end, e->d_tri + e->max_jobs, n * 4ull, cudaMemcpyDeviceToHost, st));
        if (target_end) B200_CUDA(cudaMemcpyAsync(target_end, e->d_tri + 2 * e->max_jobs, n * 4ull, cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA(cudaMemcpyAsync(e->h_cells, e->d_cells, 16, cudaMemcpyDeviceToHost, st));
    B200_CUDA(cudaMemcpyAsync(e->h_err, e->d_err, 4, cudaMemcpyDeviceToHost, st));
    e->pending = true;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_extend_async(bwa_b200_extender_t *e, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                                     const uint8_t *qseq, uint64_t q_bytes, const uint32_t *qoff, const uint32_t *qlen,
                                     const uint8_t *tseq, uint64_t t_bytes, const uint32_t *toff, const uint32_t *tlen,
                                     const uint32_t *h0, bwa_b200_ext_result_t *res6,
                                     int32_t *aln_score, int32_t *query_end, int32_t *target_end)
{
    if (!qseq || !tseq) { b200::set_error("extend_async: null argument"); return BWA_B200_ERR_ARG; }
    bwa_b200_host_page_t qp{qseq, 0, q_bytes}, tp{tseq, 0, t_bytes};
    return bwa_b200_extend_async_paged(e, p, n_jobs, &qp, 1, q_bytes, qoff, qlen, &tp, 1, t_bytes, toff, tlen, h0, res6,
                                       aln_score, query_end, target_end);
}

extern "C" int bwa_b200_extend_query(bwa_b200_extender_t *e)
{
    if (!e) return BWA_B200_ERR_ARG;
    cudaError_t st = cudaStreamQuery(e->stream);
    if (st == cudaSuccess) { e->pending = false; return 0; }
    if (st == cudaErrorNotReady) return 1;
    b200::set_error("extend_query: %s", cudaGetErrorString(st));
    return BWA_B200_ERR_CUDA;
}

extern "C" int bwa_b200_extend_wait(bwa_b200_extender_t *e)
{
    if (!e) return BWA_B200_ERR_ARG;
    B200_CUDA(cudaStreamSynchronize(e->stream));
    e->pending = false;
    if (*e->h_err) {
        *e->h_err = 0;
        cudaMemsetAsync(e->d_err, 0, 4, e->stream);
        b200::set_error("extend: a job had h0 < 1 (ksw_extend2 asserts h0 > 0, src/ksw.c:869) or a query beyond BWA_B200_EXT_INTRA_MAX_Q bases");
        return BWA_B200_ERR_ARG;
    }
    return BWA_B200_OK;
}

extern "C" uint64_t bwa_b200_extender_last_cells(bwa_b200_extender_t *e)
{
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaMemcpyAsync(e->h_cells, e->d_cells, 16, cudaMemcpyDeviceToHost, e->stream);
    cudaStreamSynchronize(e->stream);
    return *e->h_cells;
}

extern "C" uint64_t bwa_b200_extender_last_closed_form(bwa_b200_extender_t *e)
{
    if (!e) return 0;
    cudaSetDevice(e->device);
    cudaMemcpyAsync(e->h_cells, e->d_cells, 16, cudaMemcpyDeviceToHost, e->stream);
    cudaStreamSynchronize(e->stream);
    return e->h_cells[1];
}

extern "C" int bwa_b200_extender_set_closed_form(bwa_b200_extender_t *e, int enable)
{
    if (!e) return BWA_B200_ERR_ARG;
    e->no_closed_form = !enable;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_extend_device(bwa_b200_extender_t *e, const bwa_b200_ext_params_t *p, uint64_t n_jobs,
                                      const uint32_t *dev_qpacked, const uint32_t *dev_qoff, const uint32_t *dev_qlen,
                                      const uint32_t *dev_tpacked, const uint32_t *dev_toff, const uint32_t *dev_tlen,
                                      const uint32_t *dev_h0, bwa_b200_ext_result_t *dev_res6)
{
    if (!e || !p || !dev_qpacked || !dev_tpacked || !dev_qoff || !dev_qlen || !dev_toff || !dev_tlen || !dev_h0 || !dev_res6) { b200::set_error("extend_device: null argument"); return BWA_B200_ERR_ARG; }
    if (n_jobs == 0 || n_jobs >= 0x7fffffffull) { b200::set_error("extend_device: bad n_jobs"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(e->device));
    int rc = ext_grow_jobs(e, n_jobs);
    if (rc) return rc;
    JobView J{nullptr, nullptr, dev_qpacked, dev_tpacked, dev_qoff, dev_qlen, dev_toff, dev_tlen, dev_h0};
    rc = ext_launch<false>(e, p, (uint32_t)n_jobs, J, dev_res6);
    if (rc) return rc;
    B200_CUDA(cudaMemcpyAsync(e->h_err, e->d_err, 4, cudaMemcpyDeviceToHost, e->stream));
    e->pending = true;
    return BWA_B200_OK;
}

int b200_ext_run_packed(bwa_b200_extender *e, const bwa_b200_ext_params_t *p, uint32_t n,
                        const uint32_t *d_qp, const uint32_t *d_qoff, const uint32_t *d_qlen,
                        const uint32_t *d_tp, const uint32_t *d_toff, const uint32_t *d_tlen,
                        const uint32_t *d_h0, bwa_b200_ext_result_t *d_res, int64_t max_read_len)
{
    int rc = ext_grow_jobs(e, n);
    if (rc) return rc;
    JobView J{nullptr, nullptr, d_qp, d_tp, d_qoff, d_qlen, d_toff, d_tlen, d_h0};
    // queries and seed scores are bounded by the longest read (h0 = seed length * a): short reads never reach ext_intra_kernel
    int mx = 1;
    for (int i = 0; i < 25; ++i) mx = mx > p->mat[i] ? mx : p->mat[i];
    const bool may_intra = max_read_len < 0 || max_read_len > 1024 || (uint64_t)max_read_len * (uint64_t)(2 * mx) >= 32767ull;
    const bool may_wave = max_read_len < 0 || (max_read_len > WAVE_MIN_Q && (max_read_len > PAIR_MAX_Q || (uint64_t)max_read_len * (uint64_t)(2 * mx) > (uint64_t)PAIR_MAX_SCORE));
    rc = ext_launch<false>(e, p, n, J, d_res, may_intra, may_wave);
    if (rc) return rc;
    B200_CUDA(cudaMemcpyAsync(e->h_err, e->d_err, 4, cudaMemcpyDeviceToHost, e->stream));
    e->pending = true;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_pack_device(bwa_b200_extender_t *e, const uint8_t *dev_bytes, uint64_t n_bytes, uint32_t *dev_packed)
{
    if (!e || !dev_bytes || !dev_packed) return BWA_B200_ERR_ARG;
    B200_CUDA(cudaSetDevice(e->device));
    uint64_t n_words = (n_bytes + 7) / 8;
    if (n_words == 0) return BWA_B200_OK;
    pack_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, e->stream>>>(dev_bytes, n_words, n_bytes, dev_packed);
    e->launches += 1;
    B200_CUDA(cudaGetLastError());
    return BWA_B200_OK;
}
