// ext_wave.cuh -- ext_wave_kernel: one ksw_extend2 job per WARP, a row evaluated 64 columns at a time -- two columns per lane in one
// s16x2 register (the arithmetic of ext_pair_core.cuh), the F carry of the row done as a warp max-plus scan over the lanes' column
// pairs, the diagonal H(i, j-1) passed to the next lane by a shuffle.  The intra-query kernel for what the per-lane kernels do not
// take -- queries beyond 512 bases, scores beyond 1023 (up to 2^15 - 64) -- when the batch has a band and holds too few such jobs to
// fill the machine one job per lane (those batches go to ext_pair_kernel<WIDE>, which is several times faster per cell).
//
// State sized by the band, not by the query: a ring of w + 2 column pairs {H(i-1, .), E(i, .)} and their PRMT selectors per warp in
// shared memory (10 bytes per pair: 1 KB per job at w = 100, whatever the read length); a pair that enters the band for the first
// time is initialised in closed form with the reference's first row (see ext_pair_core.cuh, "Band-sized state").  Row-order
// semantics are exactly those of the per-lane kernels: rows in the reference's order over the reference's window [beg, end), the
// window re-derived from the zeros of the row just written (here: from ballots taken while the row is evaluated), columns outside
// the window untouched.
//
// The F scan.  Within a row F(j+1) = max(F(j) - e, t(j)), t(j) = max(M(j) - oe_ins, 0).  A lane holding columns (2p, 2p+1) turns the
// F entering its low column into the F leaving its high column by f_out = max(f_in - 2e, g), g = max(t(2p) - e, t(2p+1)); with
// pm(l) = g(l) + 2e * l, the F entering lane l is max(fc - 2e * l, max_{k < l} pm(k) - 2e * (l - 1)): one inclusive prefix maximum
// over the warp (5 shuffles), fc = the F carried in from the previous 64 columns.
//
// Scores: s16 with the diagonal's zero test as min(H + s, min(H, 1023) * 32) -- exact for H + s below 2^15 - 32 --, the row maximum
// as 32-bit keys (score << 16 | column), so queries up to 65535 bases and score bounds h0 + qlen * max(mat) up to 32700.
#pragma once

constexpr int WAVE_WARPS = 8;            // jobs (warps) per block
constexpr int WAVE_MAX_SCORE = 32700;
constexpr int WAVE_MIN_Q = 256;          // shorter queries outside the column-pair class stay with the 32-bit per-lane kernel (a row would not fill a warp)
constexpr int WAVE_AHEAD = 16;           // pairs staged ahead of the band's leading edge
constexpr int WAVE_MAX_RING = 2048;      // pairs: bands up to 2048 - 2 - WAVE_AHEAD (20 KB of shared memory per job)

template <bool BYTES, bool SAME_GAP>
__global__ void __launch_bounds__(WAVE_WARPS * 32)
ext_wave_kernel(ExtParams P, PairParams S, JobView J, const uint32_t *__restrict__ order, const uint32_t *__restrict__ range, uint32_t max_jobs,
                bwa_b200_ext_result_t *__restrict__ res, unsigned long long *__restrict__ cells_total, int *__restrict__ err_flag)
{
    extern __shared__ uint2 wave_smem[];
    __shared__ uint32_t stab[8];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid < 5) stab[tid] = S.tab[tid];
    __syncthreads();
    const int R = S.ring;
    const uint32_t rmagic = S.ring_magic;
    uint2 *const HE = wave_smem + (size_t)wid * R;                                                         // HE[slot]
    uint16_t *const QS = reinterpret_cast<uint16_t *>(wave_smem + (size_t)WAVE_WARPS * R) + (size_t)wid * R;   // QS[slot]
    uint16_t *const hw = reinterpret_cast<uint16_t *>(HE);
#define WSLOT(p) ((int)((uint32_t)(p) - __umulhi((uint32_t)(p), rmagic) * (uint32_t)R))
    const uint32_t lo = range[0], hi = range[1];
    if (hi - lo >= max_jobs) return;                        // a batch this large fills the machine one job per lane: ext_pair_kernel<WIDE> takes it
    const uint32_t gw = blockIdx.x * WAVE_WARPS + wid, n_warps = gridDim.x * WAVE_WARPS;
    const int oe_ins = P.o_ins + P.e_ins, e_ins = P.e_ins;
    const uint32_t tab_n = S.tab_n, noe_del2 = S.noe_del2, ne_del2 = S.ne_del2, noe_ins2 = S.noe_ins2;
    (void)noe_del2;
    unsigned long long my_cells = 0;
    for (uint32_t pos = lo + gw; pos < hi; pos += n_warps) {
        const uint32_t a = order[pos];
        const int qlen = (int)J.qlen[a], tlen = (int)J.tlen[a], h0 = (int)J.h0[a];
        const uint32_t qo = J.qoff[a], to = J.toff[a];
        if (qlen == 0) {
            if (lane == 0) { bwa_b200_ext_result_t r0; r0.score = h0; r0.qle = 0; r0.tle = 0; r0.gtle = 0; r0.gscore = -1; r0.max_off = 0; res[a] = r0; }
            continue;
        }
        if (h0 < 1 || qlen > 0xffff || (long long)h0 + (long long)qlen * P.max_score > WAVE_MAX_SCORE) { if (lane == 0) atomicExch(err_flag, 1); continue; }
        int w = P.w;
        {   // band clamp (src/ksw.c:885-893)
            int max_ins = (int)((double)(qlen * P.max_score + P.end_bonus - P.o_ins) / P.e_ins + 1.);
            max_ins = max_ins > 1 ? max_ins : 1;
            w = w < max_ins ? w : max_ins;
            int max_del = (int)((double)(qlen * P.max_score + P.end_bonus - P.o_del) / P.e_del + 1.);
            max_del = max_del > 1 ? max_del : 1;
            w = w < max_del ? w : max_del;
        }
        // staging of pair p (one lane): selectors of its two columns (columns >= qlen: N) and the first row of eh[] (src/ksw.c:880-883)
        auto stage = [&](int p) {
            uint32_t c0 = 4u, c1 = 4u;
            const int j0 = 2 * p;
            if (BYTES) {
                if (j0 < qlen) c0 = J.qb[qo + j0];
                if (j0 + 1 < qlen) c1 = J.qb[qo + j0 + 1];
            } else if (j0 < qlen) {
                const uint32_t qword = J.qp[(qo >> 3) + (j0 >> 3)];
                c0 = (qword >> (28 - 4 * (j0 & 7))) & 15u;
                if (j0 + 1 < qlen) c1 = (qword >> (24 - 4 * (j0 & 7))) & 15u;
            }
            c0 = c0 > 4u ? 4u : c0; c1 = c1 > 4u ? 4u : c1;
            const int s = WSLOT(p);
            QS[s] = (uint16_t)((c0 * 17u + 0x80u) | (c1 * 17u + 0x80u) << 8);
            int v0 = j0 == 0 ? h0 : h0 - oe_ins - (j0 - 1) * e_ins;
            int v1 = h0 - oe_ins - j0 * e_ins;
            v0 = v0 > 0 ? v0 : 0; v1 = v1 > 0 ? v1 : 0;
            HE[s] = make_uint2((uint32_t)v0 | (uint32_t)v1 << 16, 0u);
        };
        const int p_last = qlen >> 1;
        int pinit;
        {
            const int first = (w + 1) >> 1;
            const int upto = first < p_last ? first : p_last;
            for (int p = lane; p <= upto; p += 32) stage(p);
            pinit = upto + 1;
        }
        __syncwarp();
        int best = h0, best_i = -1, best_j = -1, best_ie = -1, gscore = -1, max_off = 0;
        int beg = 0, end = qlen;
        int h1_edge = h0 - P.o_del;
        uint32_t tword = 0;                                   // packed: the 8 target bases of rows i .. i | 7; bytes: the next row's base
        if (BYTES) tword = J.tb[to];
        for (int i = 0; i < tlen; ++i) {
            int tbv;
            if (BYTES) { tbv = (int)tword; if (i + 1 < tlen) tword = J.tb[to + i + 1]; }
            else {
                if ((i & 7) == 0) tword = J.tp[(to + i) >> 3];
                tbv = (int)(tword >> 28);
                tword <<= 4;
            }
            tbv = tbv > 4 ? 4 : tbv;
            const uint32_t tlo = stab[tbv];
            if (beg < i - w) beg = i - w;
            if (end > i + w + 1) end = i + w + 1;
            if (end > qlen) end = qlen;
            {   // pairs that enter the band at this row: staged WAVE_AHEAD pairs ahead of need, so that the query loads and the warp
                // barrier are paid once in 2 * WAVE_AHEAD rows (the ring holds w + 2 + WAVE_AHEAD pairs)
                int need = (i + w + 1) >> 1;
                need = need < p_last ? need : p_last;
                if (pinit <= need) {
                    int upto = need + WAVE_AHEAD;
                    upto = upto < p_last ? upto : p_last;
                    for (int p = pinit + lane; p <= upto; p += 32) stage(p);
                    pinit = upto + 1;
                    __syncwarp();
                }
            }
            h1_edge -= P.e_del;
            int h1 = 0, m = 0, mj = -1;
            if (beg == 0) h1 = h1_edge < 0 ? 0 : h1_edge;
            int firstnz = 0x7fffffff, lastnz = -1;          // first / last column of [beg, end) whose stored {H, E} is not zero
            if (beg < end) {
                const int p0 = beg >> 1, p1 = (end - 1) >> 1;
                const bool lo_out = (beg & 1) != 0, hi_out = (end & 1) != 0;
                int fc = 0;                                  // F entering the low column of the chunk's first pair
                uint32_t hc = (uint32_t)h1 << 16;            // high half: H(i, 2p - 1) for the chunk's first pair
                uint32_t key = 0u, hlast = 0u;
                for (int base = p0; base <= p1; base += 32) {
                    const int p = base + lane;
                    const bool in = p <= p1, first = p == p0, last = p == p1;
                    const int s = in ? WSLOT(p) : 0;
                    uint2 old = make_uint2(0u, 0u);
                    uint32_t sel = 0xc4c4u;
                    if (in) { old = HE[s]; sel = QS[s]; }
                    uint32_t inm = (first && lo_out) ? 0xffff0000u : 0xffffffffu;
                    if (!in) inm = 0u;
                    const uint32_t outm = (last && hi_out) ? (inm & 0x0000ffffu) : inm;
                    const uint32_t hx = old.x & inm, ey = old.y & inm;
                    const uint32_t S_ = b200_prmt(tlo, tab_n, sel);
                    const uint32_t M_ = __viaddmin_s16x2(hx, S_, __vmins2(hx, 0x03ff03ffu) * 32u);
                    const uint32_t t2_ = __viaddmax_s16x2_relu(M_, noe_ins2, noe_ins2);
                    const uint32_t t1_ = SAME_GAP ? t2_ : __viaddmax_s16x2_relu(M_, noe_del2, noe_del2);
                    // F over the warp
                    const int t2lo = (int)(t2_ & 0xffffu), t2hi = (int)(t2_ >> 16);
                    const int g = max(t2lo - e_ins, t2hi);
                    int pm = g + 2 * e_ins * lane;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int up = __shfl_up_sync(0xffffffffu, pm, o); if (lane >= o) pm = max(pm, up); }
                    const int pprev = __shfl_up_sync(0xffffffffu, pm, 1);
                    int f_in = fc - 2 * e_ins * lane;
                    if (lane > 0) f_in = max(f_in, pprev - 2 * e_ins * (lane - 1));
                    f_in = max(f_in, 0);
                    const int f1 = max(f_in - e_ins, t2lo);
                    const uint32_t F_ = (uint32_t)f_in | (uint32_t)f1 << 16;
                    const uint32_t h_ = __vimax3_s16x2(M_, ey, F_);
                    const uint32_t En_ = __viaddmax_s16x2(ey, ne_del2, t1_);
                    const int fo = max(f_in - 2 * e_ins, g);
                    fc = __shfl_sync(0xffffffffu, fo, 31);
                    // the diagonal of the next row: H(i, 2p - 1) comes from the lane below
                    uint32_t hup = __shfl_up_sync(0xffffffffu, h_, 1);
                    if (lane == 0) hup = hc;
                    if (first && lo_out) hup = old.x << 16;               // the outside column's stored H is kept
                    uint2 o_;
                    o_.x = __byte_perm(hup, h_, 0x5432);
                    o_.y = (En_ & outm) | (old.y & ~inm);
                    if (in) HE[s] = o_;
                    hc = __shfl_sync(0xffffffffu, h_, 31);
                    if (base + 32 > p1) hlast = __shfl_sync(0xffffffffu, h_, p1 - base);
                    // row maximum: the last column among equal maxima (src/ksw.c:928)
                    const bool vlo = in && !(first && lo_out), vhi = in && !(last && hi_out);
                    const uint32_t kl = vlo ? ((h_ & 0xffffu) << 16 | (uint32_t)(2 * p)) : 0u;
                    const uint32_t kh = vhi ? ((h_ >> 16) << 16 | (uint32_t)(2 * p + 1)) : 0u;
                    key = max(key, max(kl, kh));
                    // zeros of the stored row, for the next window
                    const uint32_t nzw = o_.x | o_.y;
                    const uint32_t bl = __ballot_sync(0xffffffffu, vlo && (nzw & 0xffffu) != 0u);
                    const uint32_t bh = __ballot_sync(0xffffffffu, vhi && (nzw >> 16) != 0u);
                    if (bl | bh) {
                        if (firstnz == 0x7fffffff) {
                            const int cl = bl ? 2 * (base + __ffs(bl) - 1) : 0x7fffffff, ch = bh ? 2 * (base + __ffs(bh) - 1) + 1 : 0x7fffffff;
                            firstnz = cl < ch ? cl : ch;
                        }
                        const int cl = bl ? 2 * (base + 31 - __clz(bl)) : -1, ch = bh ? 2 * (base + 31 - __clz(bh)) + 1 : -1;
                        lastnz = cl > ch ? cl : ch;
                    }
                }
                h1 = hi_out ? (int)(hlast & 0xffffu) : (int)(hlast >> 16);
                key = __reduce_max_sync(0xffffffffu, key);
                m = (int)(key >> 16); mj = (int)(key & 0xffffu);
                if (lane == 0) my_cells += (unsigned long long)(end - beg);
            }
            __syncwarp();
            if (lane == 0) {                                              // eh[end] = {h1, 0}
                const int se = WSLOT(end >> 1);
                hw[se * 4 + (end & 1)] = (uint16_t)h1;
                hw[se * 4 + 2 + (end & 1)] = 0;
            }
            __syncwarp();
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
            const int nb = firstnz != 0x7fffffff ? firstnz : (beg < end ? end : beg);
            const int j = h1 != 0 ? end : (lastnz >= 0 ? lastnz : nb - 1);
            beg = nb;
            end = j + 2 < qlen ? j + 2 : qlen;
        }
        if (lane == 0) {
            bwa_b200_ext_result_t r;
            r.score = best; r.qle = best_j + 1; r.tle = best_i + 1; r.gtle = best_ie + 1; r.gscore = gscore; r.max_off = max_off;
            res[a] = r;
        }
        __syncwarp();
    }
#undef WSLOT
    if (lane == 0 && my_cells) atomicAdd(cells_total, my_cells);
}
