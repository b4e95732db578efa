// seed.cu -- SMEM seeding over the FMD index + SA locate, sm_100a.
//
// Parity target: the reference's CPU path  mem_collect_intv pass 1 -> bwt_smem1 -> bwt_extend ->
// bwt_2occ4, then bwt_sa on the rows mem_chain reads (bwa_index/bwamem.c:121-131,278-283;
// src/bwt.c:340-566; bwa_index/bwt.c:151-172).  This is not a port of GPUSeed: the decomposition
// below keeps the CPU algorithm's exact work (one bwt_extend per step, including its
// interval-merging rule) while giving every lane an independent dependent-load chain.
//
//   fwd_kernel     one lane per read.  Runs the forward phases of successive bwt_smem1 calls
//                  (x0 = 0, x_{i+1} = ret(x_i)) as one loop over the read: one bidirectional
//                  extension (two 32-byte buckets, fetched with one 256-bit load each) per base.
//                  Every change of interval size is recorded as a candidate (x, end, k, s).
//   back_kernel    persistent lanes, one read at a time per lane, claimed per warp and streamed
//                  through shared memory with cp.async (see the kernel).  For each
//                  forward segment the candidates are walked backwards longest-first; a per-lane
//                  "envelope" of the interval sizes of the previous (longer) candidate reproduces
//                  the `ok[c].x[2] != curr->a[curr->n-1].x[2]` merge test and the
//                  `curr->n == 0` / start-coordinate emission test of src/bwt.c:528-553 exactly,
//                  so the number of extensions equals the CPU's.  Each loop iteration performs
//                  exactly one backward extension for every lane, whatever its read/segment.
//   fill_kernel    expands SMEMs to (row, qbeg, qend, score) entries at their final offsets.
//   locate_kernel  persistent lanes, one SA row per lane: LF-walk to a sampled row; symbol and
//                  occurrence count of a step come from the same 32-byte bucket.
//
// HBM traffic per read is dominated by random 32-byte sectors of the bucket array (one sector per
// occurrence lookup); candidates cost 16 B written + read once per size change.
#include "internal.h"
#include <cub/cub.cuh>

using b200::IndexView;

namespace {

constexpr int FWD_THREADS = 128;
constexpr int BACK_THREADS = 128;
constexpr int ENV_SMEM = 12;          // envelope entries kept in shared memory per lane
constexpr int LOC_THREADS = 128;
constexpr uint32_t LOC_CLAIM = 128;    // seeds a warp of locate_kernel claims per queue atomic

using b200::Cand;

// ----------------------------------------------------------------------------- bucket access
// A bucket is one 32-byte sector: counters of A,C,G,T before its first symbol, then two 64-bit bit
// planes (bit p of L / H = low / high bit of symbol p), see index.cu.
struct Bkt { uint32_t c[4]; uint32_t w[4]; };   // w = {L_lo, L_hi, H_lo, H_hi}

// L2 policy for the bucket array: evict_last, so that the streaming traffic of a batch (candidates,
// reads, seeds) does not push occurrence buckets out of L2 when the index is of the order of the L2 size
// With an index far beyond L2 the buckets get the normal priority instead, and evict_last goes to the k-mer table.
__device__ __forceinline__ uint64_t evict_last_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t bucket_policy(const IndexView &ix)
{
    uint64_t pol;
    if (ix.bkt_evict_last) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// ---------------------------------------------------------------------------- k-mer table
// The suffix-array interval of every pattern of up to K bases, as the chain of bwt_extend calls of the reference would compute it
// (kt_level_kernel below IS that chain, run once per pattern at load time).  A lookup replaces one backward or forward extension whose
// result is a pattern of at most K bases: one 8-byte load from a table that stays in L2 (45 MB for K = 11) instead of two dependent
// 32-byte sectors of the bucket array, which for a human-sized index come from HBM.  An entry whose size does not fit 24 bits
// (the shortest patterns of a large genome) says so and the caller takes the bucket path for that step; those few hundred buckets
// are L2-hot anyway.  The reference has a hook for the same idea that it leaves switched off (pre_calc_seed_intervals,
// src/GPUSeed/seed_gen.cu:1169, src/fastmap.c:455).
constexpr uint32_t KT_SAT = 0xffffffu;
// entries of levels 1 .. m-1 = (4^m - 4) / 3 = 0b0101..01 (m ones) - 1: a shift, not a 64-bit division in the inner loop
__device__ __host__ __forceinline__ uint64_t kt_off(int m) { return (uint64_t)(0x55555555u >> (32 - 2 * m)) - 1ull; }
__device__ __forceinline__ uint64_t ld_kt(const uint64_t *kt, int m, uint32_t val, uint64_t pol)
{
    uint64_t e;
    asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(e) : "l"(kt + kt_off(m) + val), "l"(pol));
    return e;
}
// the same under a predicate, without a branch: the load is issued (or not) in line, so that a warp whose lanes are split between
// table steps and bucket steps has all of its requests in flight before anything is waited for
__device__ __forceinline__ uint64_t ld_kt_if(bool doit, const uint64_t *kt, int m, uint32_t val, uint64_t pol)
{
    uint64_t e;
    const uint64_t *p = kt + (doit ? kt_off(m) + val : 0ull);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u64 %0, 0;\n\t@q ld.global.nc.L2::cache_hint.u64 %0, [%1], %3;\n\t}"
                 : "=&l"(e) : "l"(p), "r"((uint32_t)doit), "l"(pol));
    return e;
}
// a bucket under a predicate (zeros when not loaded)
__device__ __forceinline__ Bkt ld_bucket_if(bool doit, const uint32_t *bkt, uint64_t b, uint64_t pol)
{
    Bkt r;
    const uint32_t *p = bkt + (doit ? b * 8 : 0ull);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t"
                 "mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0; mov.u32 %3, 0; mov.u32 %4, 0; mov.u32 %5, 0; mov.u32 %6, 0; mov.u32 %7, 0;\n\t"
                 "@q ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %10;\n\t}"
                 : "=&r"(r.c[0]), "=&r"(r.c[1]), "=&r"(r.c[2]), "=&r"(r.c[3]), "=&r"(r.w[0]), "=&r"(r.w[1]), "=&r"(r.w[2]), "=&r"(r.w[3])
                 : "l"(p), "r"((uint32_t)doit), "l"(pol));
    return r;
}
// reverse complement of a pattern of m bases (2 bits each, first base most significant)
__device__ __forceinline__ uint32_t kt_revcomp(uint32_t val, int m)
{
    uint32_t r = __brev(val);
    r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
    return (r >> (32 - 2 * m)) ^ ((1u << (2 * m)) - 1u);
}

__device__ __forceinline__ Bkt ld_bucket(const uint32_t *bkt, uint64_t b, uint64_t pol)
{ // one 32-byte sector, one LDG.256 on sm_100
    Bkt r;
    const uint32_t *p = bkt + b * 8;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                 : "=r"(r.c[0]), "=r"(r.c[1]), "=r"(r.c[2]), "=r"(r.c[3]), "=r"(r.w[0]), "=r"(r.w[1]), "=r"(r.w[2]), "=r"(r.w[3])
                 : "l"(p), "l"(pol));
    return r;
}

// second bucket of a lookup pair: loaded under a predicate (no branch, no request when both ends
// share a bucket) into registers that do not depend on the first load, then selected, so the two
// sectors of a pair are in flight together
__device__ __forceinline__ Bkt ld_bucket_or(bool doit, const Bkt &other, const uint32_t *bkt, uint64_t b, uint64_t pol)
{
    Bkt r;
    const uint32_t *p = bkt + b * 8;
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %9, 0;\n\t"
                 "mov.u32 %0, 0; mov.u32 %1, 0; mov.u32 %2, 0; mov.u32 %3, 0; mov.u32 %4, 0; mov.u32 %5, 0; mov.u32 %6, 0; mov.u32 %7, 0;\n\t"
                 "@q ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %10;\n\t}"
                 : "=&r"(r.c[0]), "=&r"(r.c[1]), "=&r"(r.c[2]), "=&r"(r.c[3]), "=&r"(r.w[0]), "=&r"(r.w[1]), "=&r"(r.w[2]), "=&r"(r.w[3])
                 : "l"(p), "r"((uint32_t)doit), "l"(pol));
#pragma unroll
    for (int i = 0; i < 4; ++i) { r.c[i] = doit ? r.c[i] : other.c[i]; r.w[i] = doit ? r.w[i] : other.w[i]; }
    return r;
}

// masks selecting the first n (0..64) symbols of a bucket in the low / high plane words
__device__ __forceinline__ void first_n(int n, uint32_t &mlo, uint32_t &mhi)
{
    mlo = ~__funnelshift_lc(0u, 0xffffffffu, (uint32_t)n);          // n >= 32: all ones
    mhi = __funnelshift_rc(0xffffffffu, 0u, (uint32_t)(64 - n));    // n <= 32: zero
}

// occurrences of all four bases among the first n (1..64) symbols of a bucket, plus its counters
__device__ __forceinline__ void bucket_occ4(const Bkt &b, int n, uint32_t cnt[4])
{
    uint32_t mlo, mhi;
    first_n(n, mlo, mhi);
    const uint32_t t = __popc(b.w[0] & b.w[2] & mlo) + __popc(b.w[1] & b.w[3] & mhi);
    const uint32_t g = __popc(~b.w[0] & b.w[2] & mlo) + __popc(~b.w[1] & b.w[3] & mhi);
    const uint32_t c = __popc(b.w[0] & ~b.w[2] & mlo) + __popc(b.w[1] & ~b.w[3] & mhi);
    cnt[0] = b.c[0] + ((uint32_t)n - c - g - t); cnt[1] = b.c[1] + c; cnt[2] = b.c[2] + g; cnt[3] = b.c[3] + t;
}

// v[b] for b in 0..3 without a branch and without a local array: the two bits of b pick through masks (three LOP3).  The nested
// `b == 0 ? .. : b == 1 ? ..` form of this compiled to divergent branches in the extension loops (profiles: BSSY / BRA / BSYNC around
// every counter and L2 pick, lanes split three ways by their base).
__device__ __forceinline__ uint32_t sel4(int b, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3)
{
    const uint32_t m1 = 0u - (uint32_t)(b & 1), m2 = 0u - (uint32_t)((b >> 1) & 1);
    const uint32_t lo = (v0 & ~m1) | (v1 & m1), hi = (v2 & ~m1) | (v3 & m1);
    return (lo & ~m2) | (hi & m2);
}

// occurrences of one base among the first n symbols; nl / nh = all ones where the base's low / high
// bit is ZERO (so plane ^ mask has a one wherever the plane bit matches the base)
__device__ __forceinline__ uint32_t bucket_occ1(const Bkt &b, int n, int base, uint32_t nl, uint32_t nh)
{
    uint32_t mlo, mhi;
    first_n(n, mlo, mhi);
    const uint32_t r = __popc((b.w[0] ^ nl) & ((b.w[2] ^ nh) & mlo)) + __popc((b.w[1] ^ nl) & ((b.w[3] ^ nh) & mhi));
    return sel4(base, b.c[0], b.c[1], b.c[2], b.c[3]) + r;
}

// cumulative counts by select chain (a dynamic index into the kernel-parameter copy of the index
// view would force it onto the local stack)
__device__ __forceinline__ uint64_t L2_at(const IndexView &ix, int b)
{
    return b == 0 ? ix.L2[0] : (b == 1 ? ix.L2[1] : (b == 2 ? ix.L2[2] : (b == 3 ? ix.L2[3] : ix.L2[4])));
}
// the same for a base (0..3), branch-free
template <typename RowT>
__device__ __forceinline__ RowT L2_base(const IndexView &ix, int b)
{
    const uint32_t lo = sel4(b, (uint32_t)ix.L2[0], (uint32_t)ix.L2[1], (uint32_t)ix.L2[2], (uint32_t)ix.L2[3]);
    if (sizeof(RowT) == 4) return (RowT)lo;
    const uint32_t hi = sel4(b, (uint32_t)(ix.L2[0] >> 32), (uint32_t)(ix.L2[1] >> 32), (uint32_t)(ix.L2[2] >> 32), (uint32_t)(ix.L2[3] >> 32));
    return (RowT)((uint64_t)hi << 32 | lo);
}

// ------------------------------------------------------------------------------- fwd_kernel
#ifndef FWD_MIN_BLOCKS
#define FWD_MIN_BLOCKS 6        // measured on B200: C2 (32-bit rows) 6 = 8 -> 1.46 ms, 10 -> 1.64 ms, 16 -> 2.05 ms (spills); C3 (64-bit rows) 6 -> 4.30 ms, 8 -> 4.43 ms
#endif
template <typename RowT, int MINB, bool KT>
__global__ void __launch_bounds__(FWD_THREADS, MINB)
fwd_kernel(IndexView ix, const uint32_t *__restrict__ packed, const uint64_t *__restrict__ word_off,
           const uint32_t *__restrict__ read_len, uint32_t n_reads, int min_seed_len, uint32_t cand_stride,
           Cand *__restrict__ cand, uint32_t *__restrict__ n_cand, unsigned long long *__restrict__ stats)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int len = (int)read_len[r];
    const uint64_t woff = word_off[r];
    Cand *out = cand + (uint64_t)r * cand_stride;
    uint32_t n_out = 0;
    uint32_t n_sect = 0, n_tab = 0;      // bucket sectors / table entries this lane requested (reported through stats, bench.py's roofline)
    if (len < min_seed_len) { n_cand[r] = 0; return; }   // mem_chain: read shorter than a seed

    RowT k = 0, l = 0;
    const RowT primary = (RowT)ix.primary;
    const uint64_t pol = bucket_policy(ix), tpol = evict_last_policy();
    const int K = KT ? (int)ix.kt_K : 0, klo = KT ? (int)ix.kt_lo : 0;
    uint32_t s = 0;
    int x = -1, i = 0;
    uint32_t word = 0;
    uint32_t val = 0;                    // KT: the match read[x, i) as a pattern, while it has at most K bases
    bool active = false, l_ok = true;    // KT: l is not tracked through table steps; it is looked up when a bucket step needs it
    auto push = [&](int end) {
        if (end >= min_seed_len)     // streaming store: candidates are read back once, by back_kernel
            __stcs(reinterpret_cast<uint4 *>(out + n_out++), make_uint4((uint32_t)k, (uint32_t)((uint64_t)k >> 32), s, (uint32_t)x | (uint32_t)end << 16));
    };
    auto start_at = [&](int b, int p) {
        k = (RowT)L2_at(ix, b) + 1; s = (uint32_t)(L2_at(ix, b + 1) - L2_at(ix, b)); l = (RowT)L2_at(ix, 3 - b) + 1; x = p; active = true;
        val = (uint32_t)b; l_ok = true;
    };

    word = __ldg(packed + woff);
    while (i < len) {
        int b = (int)((word >> (28 - 4 * (i & 7))) & 15u);
        if (!active) {                       // looking for the first base of a segment
            if (b < 4) start_at(b, i);
            ++i;
            if ((i & 7) == 0 && i < len) word = __ldg(packed + woff + (uint32_t)(i >> 3));
            continue;
        }
        if (b > 3) {                         // ambiguous base ends the segment (src/bwt.c:516-519)
            push(i);
            active = false;
            ++i;
            if ((i & 7) == 0 && i < len) word = __ldg(packed + woff + (uint32_t)(i >> 3));
            continue;
        }
        // forward extension by complement(b): bwt_extend(ik, ok, 0)  (src/bwt.c:455-470)
        const int cb = 3 - b;
        RowT nk = 0, nl = 0;
        uint32_t ns = 0;
        const int m = i - x + 1;             // bases of the match after this extension
        const bool by_table = KT && m >= klo && m <= K;
        if (KT && m <= K) val = (val << 2) | (uint32_t)b;
        if (KT && !by_table && !l_ok) {      // x[1] of the current match = first row of its reverse complement (once per segment)
            const int m0 = m - 1;
            const uint32_t cur = m <= K ? val >> 2 : val;
            l = (RowT)(ld_kt(ix.kt, m0, kt_revcomp(cur, m0), tpol) >> 24);
        }
        {
            const uint64_t e = ld_kt_if(by_table, ix.kt, m, val, tpol);
            RowT p0 = l - 1, p1 = l - 1 + s;                     // rows; both >= 0
            RowT j0 = p0 - (RowT)(p0 >= primary), j1 = p1 - (RowT)(p1 >= primary);
            Bkt b0 = ld_bucket_if(!by_table, ix.bkt, j0 >> 6, pol);
            const Bkt b1 = ld_bucket_or(!by_table && (j1 >> 6) != (j0 >> 6), b0, ix.bkt, j1 >> 6, pol);   // both ends in one bucket: one sector (src/bwt.c:369)
            n_tab += by_table; n_sect += by_table ? 0u : 1u + (uint32_t)((j1 >> 6) != (j0 >> 6));
            uint32_t tk[4], tl[4];
            bucket_occ4(b0, (int)(j0 & 63) + 1, tk);
            bucket_occ4(b1, (int)(j1 & 63) + 1, tl);
            uint32_t s3 = tl[3] - tk[3], s2 = tl[2] - tk[2], s1 = tl[1] - tk[1], s0 = tl[0] - tk[0];
            ns = sel4(cb, s0, s1, s2, s3);
            const uint32_t tkc = sel4(cb, tk[0], tk[1], tk[2], tk[3]);
            nk = k + (RowT)(l <= primary && l + s - 1 >= primary) + (RowT)sel4(cb, s3 + s2 + s1, s3 + s2, s3, 0u);
            nl = L2_base<RowT>(ix, cb) + 1 + tkc;
            if (by_table) { ns = (uint32_t)(e & KT_SAT); nk = (RowT)(e >> 24); }
        }
        if (ns != s) {
            push(i);
            if (ns == 0) {                   // cannot extend: next bwt_smem1 call starts here
                start_at(b, i);
                ++i;
                if ((i & 7) == 0 && i < len) word = __ldg(packed + woff + (uint32_t)(i >> 3));
                continue;
            }
        }
        k = nk; s = ns;
        if (!by_table) l = nl;
        l_ok = !by_table;
        ++i;
        if ((i & 7) == 0 && i < len) word = __ldg(packed + woff + (uint32_t)(i >> 3));
    }
    if (active) push(len);                   // reached the end of the read (src/bwt.c:521-522)
    n_cand[r] = n_out;
    if (stats) { atomicAdd(stats + 0, (unsigned long long)n_sect); atomicAdd(stats + 1, (unsigned long long)n_tab); }
}

// ------------------------------------------------------------------------------ back_kernel
__device__ __forceinline__ uint32_t seeds_of(uint32_t s, int max_occ)
{ // bwa_index/bwamem.c:278-283
    if (max_occ <= 0) return s;
    uint32_t step = s > (uint32_t)max_occ ? s / (uint32_t)max_occ : 1u;
    uint32_t cnt = (s + step - 1) / step;
    return cnt < (uint32_t)max_occ ? cnt : (uint32_t)max_occ;
}

#ifndef BACK_MIN_BLOCKS
#define BACK_MIN_BLOCKS 7       // measured on B200 (tools/gpu_visit_r2x.sh, r2y.sh), C2 / C3 ms: 12 -> 5.10 / 15.26 (spills), 10 -> 4.03 / 12.19, 8 -> 3.96 / 10.77,
                                // 7 -> 3.95 / 10.27, 6 -> 4.08 / 10.26: the lanes a register cap adds cost more in spills and replays than they hide
#endif
// Work distribution.  A read has one candidate per change of interval size (~30 for a 150 bp read)
// and each is walked only a few steps, so a lane moves to its next candidate every few iterations.
// None of those moves may wait on global memory:
//  * per lane, candidates stream through a 4-entry shared-memory ring filled by cp.async four
//    candidates ahead of use (cp.async has no destination register, so nothing is waited for
//    until cp.async.wait_group, by which time the copy has long landed);
//  * per warp, reads are claimed QG at a time in converged code (one queue atomic per claim) and their
//    n_cand, word_off and last four candidates are staged in a second ring, so a lane that runs out
//    of work finds its next read on chip;
//  * the read word under the pivot is kept per segment (all candidates of a segment start there).
constexpr int QS = 32;                // claimed reads per warp (ring of {r, n_cand, word_off})
constexpr int QG = 16;                // reads claimed per refill
constexpr int KC = 4;                 // per-lane candidate ring / prefetch distance
constexpr int POP_DELAY = 2;          // iterations a lane sits out after taking a read, while its first candidates land

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc)
{
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");   // .cg: no L1 allocation
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// RESEED: the candidates come from fwd2_kernel (pass 2 of mem_collect_intv): bits 48..63 of a candidate's k hold min_intv - 1 of its
// bwt_smem1 call, and a backward extension fails when the interval gets smaller than min_intv (bwa_index/bwt.c:407) instead of empty.
template <typename RowT, int MINB, bool RESEED, bool KT>
__global__ void __launch_bounds__(BACK_THREADS, MINB)
back_kernel(IndexView ix, const uint32_t *__restrict__ packed, const uint64_t *__restrict__ word_off,
            uint32_t n_reads, int min_seed_len, int max_occ,
            uint32_t cand_stride, Cand *__restrict__ cand, const uint32_t *__restrict__ n_cand,
            uint32_t *__restrict__ n_smems, uint32_t *__restrict__ n_seeds,
            uint32_t *__restrict__ env_spill, uint32_t env_stride, unsigned long long *__restrict__ next_read,
            unsigned long long *__restrict__ stats)
{
    __shared__ uint32_t env_s[ENV_SMEM][BACK_THREADS];
    __shared__ uint4 ring[BACK_THREADS / 32][QS];        // {r, n_cand, word_off, -}
    __shared__ uint4 mine[KC][BACK_THREADS];             // candidate ring of each lane, entry = slot & (KC-1)
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t gtid = blockIdx.x * blockDim.x + tid;
    uint4 *const q = ring[tid >> 5];
    const RowT primary = (RowT)ix.primary;
    const uint64_t pol = bucket_policy(ix), tpol = evict_last_policy();
    const int K = KT ? (int)ix.kt_K : 0, klo = KT ? (int)ix.kt_lo : 0;
    uint32_t win = 0, val = 0;           // KT: the first K bases behind the pivot / the current match, as patterns (first base most significant)
    uint32_t n_sect = 0, n_tab = 0;      // bucket sectors / table entries this lane requested

    // warp-uniform queue state: slots [head, tail) hold claimed reads
    uint32_t head = 0, tail = 0;
    bool exhausted = false;

    bool finished = false, has_read = false, need_cand = true, first = true;
    uint32_t r = 0, woff = 0;
    int slot = -1, cur_x = -1, t = 0, t_head = 0, x = 0, end = 0, delay = 0;
    RowT ck = 0;
    uint32_t cs = 0, acc_seeds = 0, bw = 0, bw0 = 0, mi1 = 0;
    int wtop = -1;                       // SMEMs of the read are written downwards from the top of its row: next free slot

    for (;;) {
        // ---------------- converged: claim reads for the warp, hand them to the lanes that need one
        if (!exhausted && tail - head <= (uint32_t)(QS - QG)) {
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(next_read, (unsigned long long)QG);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (lane < QG) {
                const unsigned long long rr = base + lane;
                uint4 m = make_uint4(0xffffffffu, 0, 0, 0);
                if (rr < n_reads) m = make_uint4((uint32_t)rr, n_cand[rr], (uint32_t)word_off[rr], 0);
                q[(tail + lane) & (QS - 1)] = m;
            }
            tail += QG;
            if (base + QG >= n_reads) exhausted = true;
            __syncwarp();
        }
        {
            const bool need = !finished && !has_read;
            const uint32_t ballot = __ballot_sync(0xffffffffu, need);
            if (ballot) {
                const uint32_t cnt = __popc(ballot), avail = tail - head, rank = __popc(ballot & ((1u << lane) - 1u));
                if (need && rank < avail) {
                    const uint4 m = q[(head + rank) & (QS - 1)];
                    r = m.x;
                    if (r == 0xffffffffu) finished = true;
                    else {
                        woff = m.z;
                        slot = (int)m.y - 1;
                        const Cand *src = cand + (uint64_t)r * cand_stride;
#pragma unroll
                        for (int j = 0; j < KC; ++j)
                            if (slot - j >= 0) cp_async16(&mine[(slot - j) & (KC - 1)][tid], src + (slot - j));
                        cp_async_commit();
#pragma unroll
                        for (int j = 1; j < KC; ++j) cp_async_commit();   // empty groups: this one is now KC groups old
                        cur_x = -1; wtop = slot; acc_seeds = 0;
                        has_read = true; need_cand = true; delay = POP_DELAY;
                    }
                } else if (need && exhausted) finished = true;   // nothing claimed, nothing coming
                head += cnt < avail ? cnt : avail;
                __syncwarp();                             // ring slots are read before a later claim rewrites them
            }
        }
        if (__all_sync(0xffffffffu, finished)) break;
        if (delay > 0) { --delay; continue; }

        // ---------------- per lane: next candidate of the current read
        if (has_read && need_cand) {
            if (slot < 0) { n_smems[r] = (uint32_t)((int)n_cand[r] - 1 - wtop); n_seeds[r] = acc_seeds; has_read = false; }
            else {
                cp_async_wait_group<KC - 1>();            // the copy for `slot` is at least KC groups old (or the read's first group)
                uint4 *const m = &mine[slot & (KC - 1)][tid];
                const uint4 c = *m;
                if (slot >= KC) cp_async16(m, cand + (uint64_t)r * cand_stride + (slot - KC));
                cp_async_commit();
                ck = (RowT)(((uint64_t)(RESEED ? c.y & 0xffffu : c.y) << 32) | c.x); cs = c.z; x = (int)(c.w & 0xffffu); end = (int)(c.w >> 16);
                if (RESEED) mi1 = c.y >> 16;
                if (x != cur_x) {
                    cur_x = x; first = true; t_head = 0;
                    if (x > 0) bw0 = __ldg(packed + ((uint64_t)woff + ((uint32_t)(x - 1) >> 3)));
                    if (KT) {            // read[x, x + K) as a pattern; words behind the segment's longest candidate (this one) are not touched
                        const uint32_t w_first = (uint32_t)x >> 3, w_last = (uint32_t)(end - 1) >> 3;
                        uint32_t wd[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) wd[j] = __ldg(packed + ((uint64_t)woff + min(w_first + (uint32_t)j, w_last)));
                        win = 0;
                        for (int qq = 0; qq < K; ++qq) {
                            const uint32_t p = (uint32_t)x + (uint32_t)qq, wi = (p >> 3) - w_first;
                            const uint32_t w = wi == 0 ? wd[0] : (wi == 1 ? wd[1] : wd[2]);
                            win = (win << 2) | ((w >> (28 - 4 * (p & 7u))) & 3u);
                        }
                    }
                }
                t = 0; bw = bw0;
                if (KT) { const int l0 = end - x; val = l0 <= K ? win >> (2 * (K - l0)) : 0u; }
                need_cand = false;
            }
        }
        // ---------------- per lane: one backward extension
        if (has_read && !need_cand) {
            const int i = x - 1 - t;
            int b = 4;
            if (i >= 0) b = (int)((bw >> (28 - 4 * (i & 7))) & 15u);
            bool fail = true;
            RowT nk = 0;
            uint32_t ns = 0;
            uint32_t nval = 0;
            if (b < 4) {       // backward extension by b: only x[0], x[2] are needed downstream
                const int m = end - x + t + 1;            // bases of the match after this extension
                const bool by_table = KT && m >= klo && m <= K;
                if (KT && m <= K) nval = val | (uint32_t)b << (2 * (m - 1));
                const uint64_t e = ld_kt_if(by_table, ix.kt, m, nval, tpol);
                const RowT p0 = ck - 1, p1 = ck - 1 + cs;
                const RowT j0 = p0 - (RowT)(p0 >= primary), j1 = p1 - (RowT)(p1 >= primary);
                const Bkt b1 = ld_bucket_if(!by_table, ix.bkt, j1 >> 6, pol);
                const Bkt b0 = ld_bucket_or(!by_table && (j0 >> 6) != (j1 >> 6), b1, ix.bkt, j0 >> 6, pol);   // one sector when both ends share a bucket (src/bwt.c:312)
                n_tab += by_table; n_sect += by_table ? 0u : 1u + (uint32_t)((j0 >> 6) != (j1 >> 6));
                const uint32_t nl = (b & 1) ? 0u : 0xffffffffu, nh = (b & 2) ? 0u : 0xffffffffu;
                const uint32_t ok = bucket_occ1(b0, (int)(j0 & 63) + 1, b, nl, nh);
                const uint32_t ol = bucket_occ1(b1, (int)(j1 & 63) + 1, b, nl, nh);
                ns = ol - ok;
                nk = L2_base<RowT>(ix, b) + 1 + ok;
                if (by_table) { ns = (uint32_t)(e & KT_SAT); nk = (RowT)(e >> 24); }
                fail = RESEED ? ns <= mi1 : ns == 0;
            }
            uint32_t *const env_g = env_spill + (uint64_t)gtid * env_stride;     // steps >= ENV_SMEM (rare)
            if (fail) {
                // candidate stops after t steps: SMEM iff no longer match of this segment stopped here.  Only SMEMs are written back,
                // compacted downwards from the top of the read's row (slots at or above the current one have been consumed), so the
                // row ends with the read's SMEMs in ascending (start, end) order and the consumers never scan the other candidates
                if (first || t > t_head) {
                    if (end - (x - t) >= min_seed_len) {
                        acc_seeds += seeds_of(cs, max_occ);
                        __stcs(reinterpret_cast<uint4 *>(cand + (uint64_t)r * cand_stride + wtop),
                               make_uint4((uint32_t)ck, (uint32_t)((uint64_t)ck >> 32), cs, (uint32_t)(x - t) | (uint32_t)end << 16));
                        --wtop;
                    }
                    t_head = t; first = false;           // the envelope now covers steps [0, t)
                }
                --slot; need_cand = true;
            } else {
                uint32_t prev = 0;
                if (t < t_head) prev = t < ENV_SMEM ? env_s[t][tid] : env_g[t - ENV_SMEM];
                if (t < t_head && prev == ns) {         // same interval as the longer match: contained
                    --slot; need_cand = true;
                } else {
                    if (t < ENV_SMEM) env_s[t][tid] = ns; else env_g[t - ENV_SMEM] = ns;
                    ck = nk; cs = ns; ++t;
                    if (KT) val = nval;
                    if ((i & 7) == 0 && i > 0) bw = __ldg(packed + ((uint64_t)woff + ((uint32_t)(i - 1) >> 3)));   // previous word of the read
                }
            }
        }
    }
    if (stats) { atomicAdd(stats + 2, (unsigned long long)n_sect); atomicAdd(stats + 3, (unsigned long long)n_tab); }
}

// ------------------------------------------------------------- re-seeding (passes 2 and 3)
// mem_collect_intv after its first pass (bwa_index/bwamem.c:132-161), which the reference GPU path leaves out
// (README.md:93) and stock `bwa mem` performs:
//   pass 2  for every pass-1 SMEM of length >= split_len with at most split_width occurrences:
//           bwt_smem1(x = middle, min_intv = occurrences + 1), keep length >= min_seed_len
//   pass 3  bwt_seed_strategy1 from x = 0: forward extension until the interval is smaller than max_mem_intv and the
//           match at least min_seed_len + 1 long; restart behind it
//   sort    by (start, end)  -- entries with equal (start, end) are equal intervals, so any sort gives the CPU's list
// The merged list of read r is written to row r of a second candidate array (xstride entries per read); a read
// that needs more reports its count through *max_need and the host re-runs the two kernels with a wider row.

// forward extension of the bidirectional interval (k, l, s) by base complement cb: bwt_extend(ik, ok, 0), ok[cb]
template <typename RowT>
__device__ __forceinline__ void fwd_extend(const IndexView &ix, uint64_t pol, RowT primary, RowT k, RowT l, uint32_t s, int cb,
                                           RowT &nk, RowT &nl, uint32_t &ns)
{
    const RowT p0 = l - 1, p1 = l - 1 + s;
    const RowT j0 = p0 - (RowT)(p0 >= primary), j1 = p1 - (RowT)(p1 >= primary);
    const Bkt b0 = ld_bucket(ix.bkt, j0 >> 6, pol);
    const Bkt b1 = ld_bucket_or((j1 >> 6) != (j0 >> 6), b0, ix.bkt, j1 >> 6, pol);
    uint32_t tk[4], tl[4];
    bucket_occ4(b0, (int)(j0 & 63) + 1, tk);
    bucket_occ4(b1, (int)(j1 & 63) + 1, tl);
    const uint32_t s3 = tl[3] - tk[3], s2 = tl[2] - tk[2], s1 = tl[1] - tk[1], s0 = tl[0] - tk[0];
    ns = sel4(cb, s0, s1, s2, s3);
    const uint32_t tkc = sel4(cb, tk[0], tk[1], tk[2], tk[3]);
    nk = k + (RowT)(l <= primary && l + s - 1 >= primary) + (RowT)sel4(cb, s3 + s2 + s1, s3 + s2, s3, 0u);
    nl = L2_base<RowT>(ix, cb) + 1 + tkc;
}

// backward extension of (k, s) by base b: x[0] and x[2] of bwt_extend(ik, ok, 1), ok[b]
template <typename RowT>
__device__ __forceinline__ void back_extend(const IndexView &ix, uint64_t pol, RowT primary, RowT k, uint32_t s, int b, RowT &nk, uint32_t &ns)
{
    const RowT p0 = k - 1, p1 = k - 1 + s;
    const RowT j0 = p0 - (RowT)(p0 >= primary), j1 = p1 - (RowT)(p1 >= primary);
    const Bkt b1 = ld_bucket(ix.bkt, j1 >> 6, pol);
    const Bkt b0 = ld_bucket_or((j0 >> 6) != (j1 >> 6), b1, ix.bkt, j0 >> 6, pol);
    const uint32_t nl = (b & 1) ? 0u : 0xffffffffu, nh = (b & 2) ? 0u : 0xffffffffu;
    const uint32_t ok = bucket_occ1(b0, (int)(j0 & 63) + 1, b, nl, nh);
    const uint32_t ol = bucket_occ1(b1, (int)(j1 & 63) + 1, b, nl, nh);
    ns = ol - ok;
    nk = L2_base<RowT>(ix, b) + 1 + ok;
}

__device__ __forceinline__ int read_base(const uint32_t *__restrict__ packed, uint64_t woff, int i)
{
    return (int)((__ldg(packed + woff + (uint32_t)(i >> 3)) >> (28 - 4 * (i & 7))) & 15u);
}

constexpr int RS_THREADS = 128;

// pass 3, one lane per read: the loop of bwa_index/bwamem.c:145-158 with bwt_seed_strategy1 (bwa_index/bwt.c:434-455) unrolled
// into one forward extension per iteration, like fwd_kernel.  Entries go to the front of the read's row.
template <typename RowT>
__global__ void __launch_bounds__(RS_THREADS)
reseed3_kernel(IndexView ix, const uint32_t *__restrict__ packed, const uint64_t *__restrict__ word_off,
               const uint32_t *__restrict__ read_len, uint32_t n_reads, int min_seed_len, int max_intv,
               uint32_t xstride, Cand *__restrict__ cand2, uint32_t *__restrict__ n_cand2)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int len = (int)read_len[r];
    uint32_t n_out = 0;
    if (len >= min_seed_len && max_intv > 0) {
        const uint64_t woff = word_off[r];
        Cand *out = cand2 + (uint64_t)r * xstride;
        const RowT primary = (RowT)ix.primary;
        const uint64_t pol = bucket_policy(ix);
        RowT k = 0, l = 0;
        uint32_t s = 0, word = __ldg(packed + woff);
        int x = 0;
        bool active = false;
        for (int i = 0; i < len; ++i) {
            if ((i & 7) == 0 && i) word = __ldg(packed + woff + (uint32_t)(i >> 3));
            const int b = (int)((word >> (28 - 4 * (i & 7))) & 15u);
            if (!active) {
                if (b < 4) { k = (RowT)L2_at(ix, b) + 1; s = (uint32_t)(L2_at(ix, b + 1) - L2_at(ix, b)); l = (RowT)L2_at(ix, 3 - b) + 1; x = i; active = true; }
                continue;
            }
            if (b > 3) { active = false; continue; }           // `else return i + 1`: the next call starts behind the N
            RowT nk = 0, nl = 0;
            uint32_t ns = 0;
            if (s) fwd_extend<RowT>(ix, pol, primary, k, l, s, 3 - b, nk, nl, ns);    // an empty interval stays empty
            if (ns < (uint32_t)max_intv && i - x >= min_seed_len) {
                if (ns > 0) {                                   // `if (m.x[2] > 0) kv_push`
                    if (n_out < xstride) out[n_out] = Cand{(uint64_t)nk, ns, (uint16_t)x, (uint16_t)(i + 1)};
                    ++n_out;
                }
                active = false;                                 // returns i + 1
                continue;
            }
            k = nk; l = nl; s = ns;
        }
    }
    n_cand2[r] = n_out;                                        // may exceed xstride: merge_kernel reports it
}

// pass 2, forward phases: one lane per read walks its pass-1 SMEMs; for every long, rare one (bwa_index/bwamem.c:136-137) it runs
// the forward phase of bwt_smem1(x = middle, min_intv = occurrences + 1) -- one forward extension per loop iteration, whatever the
// lane's SMEM -- and writes the candidates to the read's row of a third candidate array.  back_kernel<RESEED> then does the
// backward phases exactly as for pass 1.  Rows hold stride3 candidates; a read that needs more reports it through *max_need.
template <typename RowT>
__global__ void __launch_bounds__(RS_THREADS)
fwd2_kernel(IndexView ix, const uint32_t *__restrict__ packed, const uint64_t *__restrict__ word_off,
            const uint32_t *__restrict__ read_len, uint32_t n_reads, int min_seed_len, int split_len, int split_width,
            uint32_t cand_stride, const Cand *__restrict__ cand, const uint32_t *__restrict__ n_cand, const uint32_t *__restrict__ n_smems,
            uint32_t stride3, Cand *__restrict__ cand3, uint32_t *__restrict__ n_cand3, unsigned long long *__restrict__ max_need)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int len = (int)read_len[r];
    const uint64_t woff = word_off[r];
    const uint32_t n1 = len >= min_seed_len ? n_smems[r] : 0u;                  // pass-1 SMEMs: the last n1 entries of the read's row
    const Cand *row1 = cand + (uint64_t)r * cand_stride + (n_cand[r] - n1);
    Cand *out = cand3 + (uint64_t)r * stride3;
    const RowT primary = (RowT)ix.primary;
    const uint64_t pol = bucket_policy(ix);
    uint32_t n_out = 0, j = 0, s = 0, mi1 = 0;
    RowT k = 0, l = 0;
    int x = 0, i = 0;
    bool have = false;
    auto push = [&](int end) {
        if (end >= min_seed_len) {                              // as fwd_kernel: a match ending before min_seed_len cannot become a seed
            if (n_out < stride3)
                __stcs(reinterpret_cast<uint4 *>(out + n_out), make_uint4((uint32_t)k, (uint32_t)((uint64_t)k >> 32) | mi1 << 16, s, (uint32_t)x | (uint32_t)end << 16));
            ++n_out;
        }
    };
    for (;;) {
        if (!have) {
            bool found = false;
            while (j < n1) {
                const Cand c = row1[j++];
                if ((int)c.end - (int)c.x < split_len || c.s > (uint32_t)split_width) continue;
                x = ((int)c.x + (int)c.end) >> 1; mi1 = c.s;     // min_intv - 1
                found = true;
                break;
            }
            if (!found) break;
            const int b0 = read_base(packed, woff, x);           // inside an SMEM: never ambiguous
            k = (RowT)L2_at(ix, b0) + 1; l = (RowT)L2_at(ix, 3 - b0) + 1; s = (uint32_t)(L2_at(ix, b0 + 1) - L2_at(ix, b0));
            i = x + 1; have = true;
            if (i >= len) { push(len); have = false; }
            continue;
        }
        const int b = read_base(packed, woff, i);
        if (b > 3) { push(i); have = false; continue; }
        RowT nk, nl;
        uint32_t ns;
        fwd_extend<RowT>(ix, pol, primary, k, l, s, 3 - b, nk, nl, ns);
        if (ns != s) {
            push(i);
            if (ns <= mi1) { have = false; continue; }           // ok[c].x[2] < min_intv: stop
        }
        k = nk; l = nl; s = ns; ++i;
        if (i == len) { push(len); have = false; }
    }
    n_cand3[r] = n_out <= stride3 ? n_out : 0u;
    if (n_out > stride3) atomicMax(max_need + 1, (unsigned long long)n_out);
}

// merge: pass-1 SMEMs (cand), pass-2 SMEMs (cand3, after back_kernel<RESEED>) and the pass-3 entries already at the front of the
// read's cand2 row; sort by (start, end); per-read interval and seed counts.
__global__ void __launch_bounds__(RS_THREADS)
merge_kernel(uint32_t n_reads, const uint32_t *__restrict__ read_len, int min_seed_len, int max_occ,
             uint32_t cand_stride, const Cand *__restrict__ cand, const uint32_t *__restrict__ n_cand, const uint32_t *__restrict__ n_smems1,
             uint32_t stride3, const Cand *__restrict__ cand3, const uint32_t *__restrict__ n_cand3, const uint32_t *__restrict__ n_smems3,
             uint32_t xstride, Cand *__restrict__ cand2, uint32_t *__restrict__ n_cand2,
             uint32_t *__restrict__ n_smems, uint32_t *__restrict__ n_seeds, unsigned long long *__restrict__ max_need)
{
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    Cand *row2 = cand2 + (uint64_t)r * xstride;
    uint32_t n_out = n_cand2[r];                               // pass-3 entries (reseed3_kernel); may already exceed the row
    if ((int)read_len[r] >= min_seed_len) {
        const uint32_t n1 = n_smems1[r], n3 = n_smems3[r];                      // SMEMs of pass 1 / pass 2 close their rows
        const Cand *row1 = cand + (uint64_t)r * cand_stride + (n_cand[r] - n1), *row3 = cand3 + (uint64_t)r * stride3 + (n_cand3[r] - n3);
        for (uint32_t j = 0; j < n1; ++j) {
            if (n_out < xstride) row2[n_out] = row1[j];
            ++n_out;
        }
        for (uint32_t j = 0; j < n3; ++j) {
            if (n_out < xstride) row2[n_out] = row3[j];
            ++n_out;
        }
    }
    uint32_t seeds = 0;
    if (n_out <= xstride) {
        for (uint32_t a = 1; a < n_out; ++a) {                 // insertion sort, a dozen entries
            const Cand v = row2[a];
            const uint32_t key = (uint32_t)v.x << 16 | v.end;
            uint32_t b = a;
            while (b > 0) {
                const Cand u = row2[b - 1];
                if (((uint32_t)u.x << 16 | u.end) <= key) break;
                row2[b] = u; --b;
            }
            if (b != a) row2[b] = v;
        }
        for (uint32_t a = 0; a < n_out; ++a) seeds += seeds_of(row2[a].s, max_occ);
        n_cand2[r] = n_out; n_smems[r] = n_out; n_seeds[r] = seeds;
    } else {
        atomicMax(max_need, (unsigned long long)n_out);
        n_cand2[r] = 0; n_smems[r] = 0; n_seeds[r] = 0;
    }
}

// ------------------------------------------------------------------------------ fill_kernel
__global__ void __launch_bounds__(128)
fill_kernel(uint32_t n_reads, int max_occ, uint32_t cand_stride, const Cand *__restrict__ cand,
            const uint32_t *__restrict__ n_cand, const uint32_t *__restrict__ n_smems, const uint64_t *__restrict__ seed_off,
            uint64_t *__restrict__ rbeg, int2 *__restrict__ qq, uint32_t *__restrict__ score, uint64_t cap)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t ns = n_smems[r];
    const Cand *rc = cand + (uint64_t)r * cand_stride + (n_cand[r] - ns);     // the read's SMEMs close its row (back_kernel / merge_kernel)
    uint64_t o = seed_off[r];
    for (uint32_t j = 0; j < ns; ++j) {
        Cand c = rc[j];
        uint32_t step = (max_occ > 0 && c.s > (uint32_t)max_occ) ? c.s / (uint32_t)max_occ : 1u;
        uint32_t cnt = seeds_of(c.s, max_occ);
        for (uint32_t t = 0; t < cnt; ++t, ++o) {
            if (o >= cap) return;
            rbeg[o] = c.k + (uint64_t)t * step;
            qq[o] = make_int2((int)c.x, (int)c.end);
            score[o] = t == 0 ? c.s : 0u;
        }
    }
}

// SMEM-only dump used by tests: (qbeg, qend, k, s) per SMEM in read order
__global__ void smem_dump_kernel(uint32_t n_reads, uint32_t cand_stride, const Cand *__restrict__ cand,
                                 const uint32_t *__restrict__ n_cand, const uint32_t *__restrict__ n_smems, const uint64_t *__restrict__ smem_off,
                                 int32_t *qbeg, int32_t *qend, uint64_t *k, uint64_t *s, uint64_t cap)
{
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint32_t ns = n_smems[r];
    const Cand *rc = cand + (uint64_t)r * cand_stride + (n_cand[r] - ns);
    uint64_t o = smem_off[r];
    for (uint32_t j = 0; j < ns; ++j, ++o) {
        Cand c = rc[j];
        if (o < cap) { qbeg[o] = c.x; qend[o] = c.end; k[o] = c.k; s[o] = c.s; }
    }
}

// ---------------------------------------------------------------------------- locate_kernel
template <typename RowT>
__global__ void __launch_bounds__(LOC_THREADS)
locate_kernel(IndexView ix, uint64_t *__restrict__ rbeg, const unsigned long long *__restrict__ total_p,
              uint64_t cap, unsigned long long *__restrict__ next_seed)
{
    const uint64_t total = min((uint64_t)*total_p, cap);
    const RowT mask = (RowT)((1ull << ix.sa_shift) - 1), primary = (RowT)ix.primary;
    const uint64_t pol = bucket_policy(ix);
    bool finished = false, need = true;
    uint64_t idx = 0;
    RowT k = 0;
    uint32_t steps = 0;
    // seeds are claimed per warp, LOC_CLAIM at a time (one queue atomic per claim instead of one per seed:
    // two million same-address atomics serialise in L2), and handed to the lanes that need one
    const uint32_t lane = threadIdx.x & 31u;
    uint64_t pool = 0;
    uint32_t pool_left = 0;
    bool exhausted = false;
    for (;;) {
        {
            const bool want = !finished && need;
            const uint32_t ballot = __ballot_sync(0xffffffffu, want);
            if (ballot) {
                const uint32_t cnt = __popc(ballot);
                if (pool_left < cnt && !exhausted) {          // refill: the few seeds left in the old pool are handed out first
                    // (pool, pool_left) is a contiguous range, so a refill must not strand seeds: take them now
                    const uint32_t rank0 = __popc(ballot & ((1u << lane) - 1u));
                    if (want && rank0 < pool_left) { idx = pool + rank0; k = (RowT)rbeg[idx]; steps = 0; need = false; }
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(next_seed, (unsigned long long)LOC_CLAIM);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    pool = base;
                    pool_left = base < total ? (uint32_t)(total - base < (uint64_t)LOC_CLAIM ? total - base : (uint64_t)LOC_CLAIM) : 0u;
                    if (base + LOC_CLAIM >= total) exhausted = true;
                }
                const bool want2 = !finished && need;
                const uint32_t ballot2 = __ballot_sync(0xffffffffu, want2);
                const uint32_t rank = __popc(ballot2 & ((1u << lane) - 1u)), cnt2 = __popc(ballot2);
                if (want2) {
                    if (rank < pool_left) { idx = pool + rank; k = (RowT)rbeg[idx]; steps = 0; need = false; }
                    else if (exhausted) finished = true;
                }
                const uint32_t take = cnt2 < pool_left ? cnt2 : pool_left;
                pool += take; pool_left -= take;
            }
        }
        if (__all_sync(0xffffffffu, finished)) break;
        if (!finished) {
            if ((k & mask) == 0) {
                uint64_t j = (uint64_t)(k >> ix.sa_shift), pos;
                if (j == 0) pos = (uint64_t)steps - 1;                      // sa[0] == -1 (bwa_index/bwt.c:160-163)
                else {
                    uint64_t hi = 0;
                    if (ix.pack_mask) {
                        uint32_t per = 32u / ix.pack_size;
                        hi = (ix.sa_hi[j / per] >> ((uint32_t)(j % per) * ix.pack_size)) & ix.pack_mask;
                    }
                    pos = steps + ((uint64_t)ix.sa[j] | hi << 32);
                }
                rbeg[idx] = pos;
                need = true;
            } else if (k == primary) {                               // bwt_invPsi: row of '$'
                k = 0; ++steps;
            } else {
                RowT j = k - (RowT)(k > primary);
                Bkt b = ld_bucket(ix.bkt, j >> 6, pol);
                const int off = (int)(j & 63);
                const uint32_t lw = off < 32 ? b.w[0] : b.w[1], hw = off < 32 ? b.w[2] : b.w[3];
                const int sym = (int)(((lw >> (off & 31)) & 1u) | (((hw >> (off & 31)) & 1u) << 1));
                k = L2_base<RowT>(ix, sym) + bucket_occ1(b, off + 1, sym, (sym & 1) ? 0u : 0xffffffffu, (sym & 2) ? 0u : 0xffffffffu);
                ++steps;
            }
        }
    }
}

// Random 32-byte-sector gather: the roofline denominator for the seeding kernels.
// Every lane issues `iters` independent 256-bit loads at hashed sector addresses of a buffer, RS_UNROLL of them in flight
// at a time (2048 resident lanes per SM x 8 = 16 K sectors in flight per SM: the memory system, not the issue rate, is the limit;
// the range reduction is a multiply-high, not a 64-bit modulo).
constexpr int RS_UNROLL = 8;
__global__ void __launch_bounds__(256)
random_sector_kernel(const uint32_t *__restrict__ buf, uint64_t n_sectors, int iters, uint32_t *__restrict__ sink)
{
    uint64_t x = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    uint32_t acc = 0;
    const uint64_t pol = evict_last_policy();
    for (int i = 0; i < iters; i += RS_UNROLL) {
        Bkt b[RS_UNROLL];
#pragma unroll
        for (int u = 0; u < RS_UNROLL; ++u) {
            x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
            b[u] = ld_bucket(buf, __umul64hi(x * 0x94D049BB133111EBull, n_sectors), pol);
        }
#pragma unroll
        for (int u = 0; u < RS_UNROLL; ++u) acc += b[u].c[0] ^ b[u].w[3];
    }
    if (acc == 0x7fffffffu) sink[0] = acc;
}

__global__ void total_kernel(const uint32_t *n_per, const uint64_t *off, uint32_t n, unsigned long long *total)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *total = n ? off[n - 1] + n_per[n - 1] : 0ull;
}

struct U32ToU64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return (uint64_t)v; } };

using FwdFn = void (*)(IndexView, const uint32_t *, const uint64_t *, const uint32_t *, uint32_t, int, uint32_t, Cand *, uint32_t *, unsigned long long *);
using BackFn = void (*)(IndexView, const uint32_t *, const uint64_t *, uint32_t, int, int, uint32_t, Cand *, const uint32_t *, uint32_t *,
                        uint32_t *, uint32_t *, uint32_t, unsigned long long *, unsigned long long *);
// register budget variants (blocks of 128 lanes per SM): more resident lanes = more sectors in flight; KT = with the k-mer table
template <typename RowT, bool KT> FwdFn fwd_variant_t(int minb)
{
    return minb >= 12 ? fwd_kernel<RowT, 12, KT> : (minb >= 10 ? fwd_kernel<RowT, 10, KT> : (minb >= 8 ? fwd_kernel<RowT, 8, KT> : fwd_kernel<RowT, 6, KT>));
}
FwdFn fwd_variant(bool narrow, int minb, bool kt)
{
    if (narrow) return kt ? fwd_variant_t<uint32_t, true>(minb) : fwd_variant_t<uint32_t, false>(minb);
    return kt ? fwd_variant_t<uint64_t, true>(minb) : fwd_variant_t<uint64_t, false>(minb);
}
template <typename RowT, bool RESEED, bool KT> BackFn back_variant_t(int minb)
{
    return minb >= 12 ? back_kernel<RowT, 12, RESEED, KT> : (minb >= 10 ? back_kernel<RowT, 10, RESEED, KT> : (minb >= 8 ? back_kernel<RowT, 8, RESEED, KT>
         : (minb >= 7 ? back_kernel<RowT, 7, RESEED, KT> : back_kernel<RowT, 6, RESEED, KT>)));
}
template <bool RESEED> BackFn back_variant_r(bool narrow, int minb, bool kt)
{
    if (narrow) return kt ? back_variant_t<uint32_t, RESEED, true>(minb) : back_variant_t<uint32_t, RESEED, false>(minb);
    return kt ? back_variant_t<uint64_t, RESEED, true>(minb) : back_variant_t<uint64_t, RESEED, false>(minb);
}
BackFn back_variant(bool narrow, int minb, bool kt, bool reseed = false) { return reseed ? back_variant_r<true>(narrow, minb, kt) : back_variant_r<false>(narrow, minb, kt); }

// ---- k-mer table construction: level m from level m - 1, one lane per parent pattern P: the four backward extensions c.P of
// bwt_extend(ik, ok, 1) (src/bwt.c:455-470), exact 64-bit {k, s} kept in scratch arrays, then packed into the table
template <typename RowT>
__global__ void __launch_bounds__(256)
kt_level_kernel(IndexView ix, int m, const uint64_t *__restrict__ pk, const uint64_t *__restrict__ ps, uint64_t *__restrict__ ck, uint64_t *__restrict__ cs)
{
    const uint64_t n_parent = 1ull << (2 * (m - 1));
    const uint64_t P = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= n_parent) return;
    const RowT primary = (RowT)ix.primary;
    const uint64_t pol = bucket_policy(ix);
    const RowT k = (RowT)pk[P], s = (RowT)ps[P];
    const RowT p0 = k - 1, p1 = k - 1 + s;
    const RowT j0 = p0 - (RowT)(p0 >= primary), j1 = p1 - (RowT)(p1 >= primary);
    const Bkt b0 = ld_bucket(ix.bkt, j0 >> 6, pol);
    const Bkt b1 = ld_bucket(ix.bkt, j1 >> 6, pol);
    uint32_t tk[4], tl[4];
    bucket_occ4(b0, (int)(j0 & 63) + 1, tk);
    bucket_occ4(b1, (int)(j1 & 63) + 1, tl);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint64_t child = ((uint64_t)c << (2 * (m - 1))) | P;
        ck[child] = L2_at(ix, c) + 1 + tk[c];
        cs[child] = (uint64_t)(uint32_t)(tl[c] - tk[c]);
    }
}

__global__ void kt_level1_kernel(IndexView ix, uint64_t *ck, uint64_t *cs)
{
    const int c = threadIdx.x;
    if (c < 4) { ck[c] = L2_at(ix, c) + 1; cs[c] = L2_at(ix, c + 1) - L2_at(ix, c); }     // bwt_set_intv
}

__global__ void __launch_bounds__(256)
kt_pack_kernel(int m, const uint64_t *__restrict__ k, const uint64_t *__restrict__ s, uint64_t *__restrict__ kt, unsigned long long *__restrict__ level_max)
{
    const uint64_t n = 1ull << (2 * m);
    const uint64_t P = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= n) return;
    const bool fits = s[P] < (uint64_t)KT_SAT && k[P] < (1ull << 40);
    kt[kt_off(m) + P] = fits ? (k[P] << 24 | s[P]) : (uint64_t)KT_SAT;
    unsigned long long v = fits ? s[P] : ~0ull;
    for (int o = 16; o; o >>= 1) { const unsigned long long u = __shfl_xor_sync(0xffffffffu, v, o); v = u > v ? u : v; }
    if ((threadIdx.x & 31) == 0) atomicMax(level_max + m, v);          // the widest interval of the level (all ones: one does not fit)
}

} // namespace

// level sizes beyond a 32-bit occurrence difference cannot arise: the bucket counters are 32-bit (index.cu refuses larger texts)
int b200_index_build_kmer_table(bwa_b200_index *idx, int K)
{
    B200_CUDA(cudaSetDevice(idx->device));
    if (idx->d_kt) { cudaFree(idx->d_kt); idx->d_kt = nullptr; }
    idx->kt_K = 0; idx->v.kt = nullptr; idx->v.kt_K = 0;
    if (K <= 0) return BWA_B200_OK;
    if (K > 14) K = 14;
    const uint64_t total = kt_off(K + 1), top = 1ull << (2 * K);
    uint64_t *d_kt = nullptr, *d_k[2] = {nullptr, nullptr}, *d_s[2] = {nullptr, nullptr};
    B200_CUDA(cudaMalloc(&d_kt, total * 8));
    for (int j = 0; j < 2; ++j) { B200_CUDA(cudaMalloc(&d_k[j], top * 8)); B200_CUDA(cudaMalloc(&d_s[j], top * 8)); }
    const bool narrow = idx->v.seq_len < 0xfffffff0ull;
    unsigned long long *d_max = nullptr, h_max[16];
    B200_CUDA(cudaMalloc(&d_max, sizeof(h_max)));
    B200_CUDA(cudaMemset(d_max, 0, sizeof(h_max)));
    kt_level1_kernel<<<1, 32>>>(idx->v, d_k[1], d_s[1]);
    kt_pack_kernel<<<1, 256>>>(1, d_k[1], d_s[1], d_kt, d_max);
    for (int m = 2; m <= K; ++m) {
        const int src = (m - 1) & 1, dst = m & 1;
        const uint64_t n_parent = 1ull << (2 * (m - 1));
        const unsigned grid = (unsigned)((n_parent + 255) / 256);
        if (narrow) kt_level_kernel<uint32_t><<<grid, 256>>>(idx->v, m, d_k[src], d_s[src], d_k[dst], d_s[dst]);
        else kt_level_kernel<uint64_t><<<grid, 256>>>(idx->v, m, d_k[src], d_s[src], d_k[dst], d_s[dst]);
        kt_pack_kernel<<<(unsigned)((4 * n_parent + 255) / 256), 256>>>(m, d_k[dst], d_s[dst], d_kt, d_max);
    }
    cudaError_t er = cudaMemcpy(h_max, d_max, sizeof(h_max), cudaMemcpyDeviceToHost);
    cudaFree(d_max);
    for (int j = 0; j < 2; ++j) { cudaFree(d_k[j]); cudaFree(d_s[j]); }
    if (er != cudaSuccess || (er = cudaGetLastError()) != cudaSuccess) { cudaFree(d_kt); b200::set_error("k-mer table build failed: %s", cudaGetErrorString(er)); return BWA_B200_ERR_CUDA; }
    // the table serves levels kt_lo .. K: the deepest run of levels in which every interval fits an entry.  BWA_B200_KMER_SAT lowers
    // the size regarded as fitting (tests: make the short levels defer to the buckets on a small genome)
    uint64_t sat = KT_SAT;
    if (const char *ev = getenv("BWA_B200_KMER_SAT")) { const long long v = atoll(ev); if (v > 0 && (uint64_t)v < sat) sat = (uint64_t)v; }
    int lo = K + 1;
    while (lo > 1 && h_max[lo - 1] < sat) --lo;
    if (lo > K) { cudaFree(d_kt); return BWA_B200_OK; }                 // nothing fits: no table
    idx->kt_lo = lo; idx->v.kt_lo = (uint32_t)lo;
    idx->d_kt = d_kt; idx->kt_K = K;
    idx->v.kt = d_kt; idx->v.kt_K = (uint32_t)K;
    return BWA_B200_OK;
}

// =============================================================================== host side
static int seeder_ensure_cand(bwa_b200_seeder *s, uint64_t n_reads, uint32_t max_len, int min_seed_len)
{
    uint32_t stride = max_len >= (uint32_t)min_seed_len ? max_len - (uint32_t)min_seed_len + 2 : 2;
    uint64_t need = n_reads * (uint64_t)stride;
    if (need > s->cand_cap) {
        if (s->d_cand) B200_CUDA(cudaFree(s->d_cand));
        s->d_cand = nullptr;
        B200_CUDA(cudaMalloc(&s->d_cand, need * sizeof(Cand)));
        s->cand_cap = need;
    }
    s->cand_stride = stride;
    uint32_t env_stride = max_len > ENV_SMEM ? max_len - ENV_SMEM : 1;
    if (env_stride > s->env_stride) {
        if (s->d_env) B200_CUDA(cudaFree(s->d_env));
        s->d_env = nullptr;
        B200_CUDA(cudaMalloc(&s->d_env, (uint64_t)s->back_grid * BACK_THREADS * env_stride * 4));
        s->env_stride = env_stride;
    }
    return BWA_B200_OK;
}

static int seeder_ensure_reseed(bwa_b200_seeder *s, uint64_t n_reads, uint32_t max_len, uint32_t want_x, uint32_t want_3)
{
    // BWA_B200_RESEED_ROW0 starts the rows at that width instead (tests use a tiny one to drive every caller through the widen-and-redo path)
    const char *row0 = getenv("BWA_B200_RESEED_ROW0");
    uint32_t xs = row0 ? (uint32_t)std::max(1, atoi(row0)) : std::max<uint32_t>(24u, max_len / 4);
    uint32_t s3 = row0 ? (uint32_t)std::max(1, atoi(row0)) : std::max<uint32_t>(48u, max_len / 2);
    xs = std::max(xs, std::max(want_x, s->xstride));
    s3 = std::max(s3, std::max(want_3, s->stride3));
    if (n_reads * (uint64_t)xs > s->cand2_cap) {
        if (s->d_cand2) B200_CUDA(cudaFree(s->d_cand2));
        s->d_cand2 = nullptr; s->cand2_cap = 0;
        B200_CUDA(cudaMalloc(&s->d_cand2, n_reads * (uint64_t)xs * sizeof(Cand)));
        s->cand2_cap = n_reads * (uint64_t)xs;
    }
    if (n_reads * (uint64_t)s3 > s->cand3_cap) {
        if (s->d_cand3) B200_CUDA(cudaFree(s->d_cand3));
        s->d_cand3 = nullptr; s->cand3_cap = 0;
        B200_CUDA(cudaMalloc(&s->d_cand3, n_reads * (uint64_t)s3 * sizeof(Cand)));
        s->cand3_cap = n_reads * (uint64_t)s3;
    }
    s->xstride = xs; s->stride3 = s3;
    return BWA_B200_OK;
}

static int seeder_ensure_out(bwa_b200_seeder *s, uint64_t n_seeds)
{
    if (n_seeds <= s->seed_cap) return BWA_B200_OK;
    uint64_t cap = n_seeds + n_seeds / 8 + 1024;
    if (s->d_rbeg) { cudaFree(s->d_rbeg); cudaFree(s->d_qq); cudaFree(s->d_score); }
    s->d_rbeg = nullptr; s->d_qq = nullptr; s->d_score = nullptr; s->seed_cap = 0;
    B200_CUDA(cudaMalloc(&s->d_rbeg, cap * 8));
    B200_CUDA(cudaMalloc(&s->d_qq, cap * 8));
    B200_CUDA(cudaMalloc(&s->d_score, cap * 4));
    s->seed_cap = cap;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_seeder_create(const bwa_b200_index_t *idx, uint64_t max_reads, uint64_t max_words,
                                      bwa_b200_seeder_t **out)
{
    if (!idx || !out || max_reads == 0) { b200::set_error("seeder_create: bad argument"); return BWA_B200_ERR_ARG; }
    if (max_reads > 0x80000000ull) { b200::set_error("seeder_create: at most 2^31 reads per batch"); return BWA_B200_ERR_ARG; }
    if (max_words >= 0xffffffffull) { b200::set_error("seeder_create: at most 2^32-2 packed words (32 G bases) per batch"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(idx->device));
    bwa_b200_seeder *s = new bwa_b200_seeder();
    s->idx = idx; s->device = idx->device;
    s->max_reads = max_reads; s->max_words = max_words;
    cudaDeviceProp prop;
    B200_CUDA(cudaGetDeviceProperties(&prop, idx->device));
    s->n_sm = prop.multiProcessorCount;
    int occ_b = 0, occ_l = 0;
    s->fwd_minb = getenv("BWA_B200_FWD_MINB") ? atoi(getenv("BWA_B200_FWD_MINB")) : FWD_MIN_BLOCKS;
    s->back_minb = getenv("BWA_B200_BACK_MINB") ? atoi(getenv("BWA_B200_BACK_MINB")) : BACK_MIN_BLOCKS;
    // 32-bit row arithmetic when every BWT row fits; BWA_B200_WIDE_ROWS forces the 64-bit kernels (human-sized
    // indexes) so that tests can cover them on small genomes
    const bool narrow_rows = idx->v.seq_len < 0xfffffff0ull && !getenv("BWA_B200_WIDE_ROWS");
    s->narrow_rows = narrow_rows;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, back_variant(narrow_rows, s->back_minb, idx->v.kt_K > 0), BACK_THREADS, 0));
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l, locate_kernel<uint64_t>, LOC_THREADS, 0));
    s->back_grid = s->n_sm * (occ_b > 0 ? occ_b : 1);
    s->loc_grid = s->n_sm * (occ_l > 0 ? occ_l : 1);
    B200_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    if (const char *pv = getenv("BWA_B200_L2_PERSIST")) {
        // keep the occurrence buckets resident in L2 (set-aside + access policy window on this stream)
        double ratio = atof(pv);
        if (ratio > 0 && prop.persistingL2CacheMaxSize > 0) {
            size_t bytes = ((idx->n_words + 7) / 8 * 8 + 8) * 4;
            size_t win = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, prop.persistingL2CacheMaxSize);
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof(av));
            av.accessPolicyWindow.base_ptr = (void *)idx->d_bkt;
            av.accessPolicyWindow.num_bytes = win;
            double fit = (double)prop.persistingL2CacheMaxSize / (double)win;
            av.accessPolicyWindow.hitRatio = (float)(ratio * (fit < 1.0 ? fit : 1.0));
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            cudaError_t er = cudaStreamSetAttribute(s->stream, cudaStreamAttributeAccessPolicyWindow, &av);
            fprintf(stderr, "[b200] L2 persist: window %zu MB, set-aside %d MB, hitRatio %.2f (%s)\n", win >> 20,
                    prop.persistingL2CacheMaxSize >> 20, av.accessPolicyWindow.hitRatio, cudaGetErrorString(er));
        }
    }
    B200_CUDA(cudaMalloc(&s->d_packed, (max_words ? max_words : 1) * 4));
    B200_CUDA(cudaMalloc(&s->d_len, max_reads * 4));
    B200_CUDA(cudaMalloc(&s->d_woff, (max_reads + 1) * 8));
    B200_CUDA(cudaMalloc(&s->d_ncand, max_reads * 4));
    B200_CUDA(cudaMalloc(&s->d_nsmems, max_reads * 4));
    B200_CUDA(cudaMalloc(&s->d_nseeds, max_reads * 4));
    B200_CUDA(cudaMalloc(&s->d_seed_off, max_reads * 8));
    B200_CUDA(cudaMalloc(&s->d_smem_off, max_reads * 8));
    B200_CUDA(cudaMalloc(&s->d_counters, 8 * sizeof(unsigned long long)));
    B200_CUDA(cudaHostAlloc(&s->h_counters, 8 * sizeof(unsigned long long), cudaHostAllocDefault));
    B200_CUDA(cudaMalloc(&s->d_ncand2, max_reads * 4));
    B200_CUDA(cudaMalloc(&s->d_ncand3, max_reads * 4));
    B200_CUDA(cudaMalloc(&s->d_rs_dummy, max_reads * 8));
    B200_CUDA(cudaMalloc(&s->d_nsmems1, max_reads * 4));
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t *> it(s->d_nseeds, U32ToU64());
    B200_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, s->cub_bytes, it, s->d_seed_off, (int)max_reads, s->stream));
    B200_CUDA(cudaMalloc(&s->d_cub, s->cub_bytes + 16));
    int rc = seeder_ensure_out(s, max_reads * 4);
    if (rc) return rc;
    *out = s;
    return BWA_B200_OK;
}

extern "C" void bwa_b200_seeder_destroy(bwa_b200_seeder_t *s)
{
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    cudaFree(s->d_packed); cudaFree(s->d_len); cudaFree(s->d_woff); cudaFree(s->d_cand); cudaFree(s->d_ncand);
    cudaFree(s->d_nsmems); cudaFree(s->d_nseeds); cudaFree(s->d_env); cudaFree(s->d_seed_off); cudaFree(s->d_smem_off);
    cudaFree(s->d_counters); cudaFree(s->d_cub); cudaFree(s->d_rbeg); cudaFree(s->d_qq); cudaFree(s->d_score);
    cudaFree(s->d_stats); cudaFree(s->d_cand2); cudaFree(s->d_ncand2); cudaFree(s->d_cand3); cudaFree(s->d_ncand3); cudaFree(s->d_rs_dummy); cudaFree(s->d_nsmems1);
    cudaFreeHost(s->h_counters);
    cudaStreamDestroy(s->stream);
    delete s;
}

extern "C" void bwa_b200_seed_params_default(bwa_b200_seed_params_t *p)
{ // mem_opt_init, bwa_index/bwamem.c:56-62
    if (!p) return;
    p->min_seed_len = 19; p->max_occ = 500; p->reseed = 0; p->split_factor = 1.5f; p->split_width = 10; p->max_mem_intv = 20;
}

extern "C" void *bwa_b200_seeder_stream(bwa_b200_seeder_t *s) { return s ? (void *)s->stream : nullptr; }

// request counters of pass 1: enable != 0 makes fwd_kernel / back_kernel count the bucket sectors and k-mer table entries they ask
// for (one atomic per lane at exit); out = {fwd sectors, fwd table entries, back sectors, back table entries} of the last batch
extern "C" int bwa_b200_seeder_request_counts(bwa_b200_seeder_t *s, int enable, uint64_t out[4])
{
    if (!s) return BWA_B200_ERR_ARG;
    B200_CUDA(cudaSetDevice(s->device));
    B200_CUDA(cudaStreamSynchronize(s->stream));
    if (out) {
        for (int i = 0; i < 4; ++i) out[i] = 0;
        if (s->d_stats) B200_CUDA(cudaMemcpy(out, s->d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    if (enable && !s->d_stats) { B200_CUDA(cudaMalloc(&s->d_stats, 4 * sizeof(unsigned long long))); B200_CUDA(cudaMemset(s->d_stats, 0, 4 * sizeof(unsigned long long))); }
    if (!enable && s->d_stats) { cudaFree(s->d_stats); s->d_stats = nullptr; }
    return BWA_B200_OK;
}
extern "C" uint64_t bwa_b200_seeder_launches(const bwa_b200_seeder_t *s) { return s ? s->launches : 0; }

static int seeder_fill_locate(bwa_b200_seeder *s)
{
    const IndexView &ix = s->idx->v;
    const uint32_t n = (uint32_t)s->last_n_reads;
    cudaStream_t st = s->stream;
    const bool rs = s->last_p.reseed != 0;
    B200_LAUNCH(s->prof, "fill_kernel", st,
        (fill_kernel<<<(n + 127) / 128, 128, 0, st>>>(n, s->last_p.max_occ, rs ? s->xstride : s->cand_stride, rs ? s->d_cand2 : s->d_cand,
                                                       rs ? s->d_ncand2 : s->d_ncand, s->d_nsmems, s->d_seed_off, s->d_rbeg, s->d_qq, s->d_score, s->seed_cap)));
    if (s->narrow_rows)
        B200_LAUNCH(s->prof, "locate_kernel", st,
            (locate_kernel<uint32_t><<<s->loc_grid, LOC_THREADS, 0, st>>>(ix, s->d_rbeg, s->d_counters + 2, s->seed_cap, s->d_counters + 1)));
    else
        B200_LAUNCH(s->prof, "locate_kernel", st,
            (locate_kernel<uint64_t><<<s->loc_grid, LOC_THREADS, 0, st>>>(ix, s->d_rbeg, s->d_counters + 2, s->seed_cap, s->d_counters + 1)));
    s->launches += 2;
    B200_CUDA(cudaGetLastError());
    return BWA_B200_OK;
}

// passes 2 and 3 of mem_collect_intv over the current batch: the merged, sorted interval rows and the per-read counts
static int seeder_reseed(bwa_b200_seeder *s)
{
    const IndexView &ix = s->idx->v;
    const uint32_t n = (uint32_t)s->last_n_reads;
    const bwa_b200_seed_params_t &p = s->last_p;
    cudaStream_t st = s->stream;
    const int split_len = (int)(p.min_seed_len * p.split_factor + .499f);      // bwa_index/bwamem.c:118
    const unsigned grid = (n + RS_THREADS - 1) / RS_THREADS;
    unsigned long long *cnt = s->d_counters;
    if (!s->rs_saved) {                   // first time for this batch: d_nsmems still holds back_kernel's pass-1 counts
        B200_CUDA(cudaMemcpyAsync(s->d_nsmems1, s->d_nsmems, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
        s->rs_saved = true;
    }
    B200_CUDA(cudaMemsetAsync(cnt + 4, 0, 3 * sizeof(unsigned long long), st));   // [4] [5] widest rows needed, [6] read queue of the second back_kernel
    if (s->narrow_rows) {
        B200_LAUNCH(s->prof, "reseed3_kernel", st,
            (reseed3_kernel<uint32_t><<<grid, RS_THREADS, 0, st>>>(ix, s->cur_packed, s->cur_woff, s->cur_len, n, p.min_seed_len, p.max_mem_intv, s->xstride, s->d_cand2, s->d_ncand2)));
        B200_LAUNCH(s->prof, "fwd2_kernel", st,
            (fwd2_kernel<uint32_t><<<grid, RS_THREADS, 0, st>>>(ix, s->cur_packed, s->cur_woff, s->cur_len, n, p.min_seed_len, split_len, p.split_width, s->cand_stride, s->d_cand,
                                                                s->d_ncand, s->d_nsmems1, s->stride3, s->d_cand3, s->d_ncand3, cnt + 4)));
    } else {
        B200_LAUNCH(s->prof, "reseed3_kernel", st,
            (reseed3_kernel<uint64_t><<<grid, RS_THREADS, 0, st>>>(ix, s->cur_packed, s->cur_woff, s->cur_len, n, p.min_seed_len, p.max_mem_intv, s->xstride, s->d_cand2, s->d_ncand2)));
        B200_LAUNCH(s->prof, "fwd2_kernel", st,
            (fwd2_kernel<uint64_t><<<grid, RS_THREADS, 0, st>>>(ix, s->cur_packed, s->cur_woff, s->cur_len, n, p.min_seed_len, split_len, p.split_width, s->cand_stride, s->d_cand,
                                                                s->d_ncand, s->d_nsmems1, s->stride3, s->d_cand3, s->d_ncand3, cnt + 4)));
    }
    // backward phases of the pass-2 calls: the pass-1 kernel, with the per-candidate min_intv (its per-read counts are not used)
    B200_LAUNCH(s->prof, "back_kernel_pass2", st,
        (back_variant(s->narrow_rows, s->back_minb, ix.kt_K > 0, true)<<<s->back_grid, BACK_THREADS, 0, st>>>(ix, s->cur_packed, s->cur_woff, n, p.min_seed_len, p.max_occ, s->stride3,
                                                                                                 s->d_cand3, s->d_ncand3, s->d_rs_dummy, s->d_rs_dummy + s->max_reads, s->d_env,
                                                                                                 s->env_stride, cnt + 6, (unsigned long long *)nullptr)));
    B200_LAUNCH(s->prof, "merge_kernel", st,
        (merge_kernel<<<grid, RS_THREADS, 0, st>>>(n, s->cur_len, p.min_seed_len, p.max_occ, s->cand_stride, s->d_cand, s->d_ncand, s->d_nsmems1, s->stride3, s->d_cand3, s->d_ncand3, s->d_rs_dummy,
                                                   s->xstride, s->d_cand2, s->d_ncand2, s->d_nsmems, s->d_nseeds, cnt + 4)));
    s->launches += 4;
    B200_CUDA(cudaGetLastError());
    return BWA_B200_OK;
}

// per-read seed offsets and the seed total
static int seeder_scan(bwa_b200_seeder *s)
{
    const uint32_t n = (uint32_t)s->last_n_reads;
    cudaStream_t st = s->stream;
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t *> it(s->d_nseeds, U32ToU64());
    size_t tmp = s->cub_bytes;
    if (s->prof) s->prof->begin("scan", st);
    B200_CUDA(cub::DeviceScan::ExclusiveSum(s->d_cub, tmp, it, s->d_seed_off, (int)n, st));
    total_kernel<<<1, 1, 0, st>>>(s->d_nseeds, s->d_seed_off, n, s->d_counters + 2);
    if (s->prof) s->prof->end(st);
    s->launches += 2;
    B200_CUDA(cudaGetLastError());
    return BWA_B200_OK;
}

// enqueue fwd -> back -> [re-seeding] -> scan -> total -> fill -> locate on s->stream; no host synchronisation
int b200_seeder_run(bwa_b200_seeder *s, const uint32_t *d_packed, const uint64_t *d_woff, const uint32_t *d_len,
                    uint64_t n_reads, uint32_t max_len, const bwa_b200_seed_params_t *p)
{
    if (n_reads > s->max_reads) { b200::set_error("seed: %llu reads > capacity %llu", (unsigned long long)n_reads, (unsigned long long)s->max_reads); return BWA_B200_ERR_CAPACITY; }
    if (max_len > 65535) { b200::set_error("seed: reads longer than 65535 bases are not supported"); return BWA_B200_ERR_ARG; }
    if (p->min_seed_len < 1) { b200::set_error("seed: min_seed_len < 1"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(s->device));
    s->last_n_reads = n_reads; s->last_total = 0; s->last_p = *p; s->filled = false; s->rs_saved = false;
    if (n_reads == 0) return BWA_B200_OK;
    if (!s->idx->v.sa) { b200::set_error("seed: index has no suffix array samples"); return BWA_B200_ERR_ARG; }
    int rc = seeder_ensure_cand(s, n_reads, max_len, p->min_seed_len);
    if (rc) return rc;
    const IndexView &ix = s->idx->v;
    const uint32_t n = (uint32_t)n_reads;
    cudaStream_t st = s->stream;
    B200_CUDA(cudaMemsetAsync(s->d_counters, 0, 8 * sizeof(unsigned long long), st));
    if (s->d_stats) B200_CUDA(cudaMemsetAsync(s->d_stats, 0, 4 * sizeof(unsigned long long), st));
    // 32-bit row arithmetic when every BWT row fits (seq_len < 2^32), 64-bit otherwise (human-sized)
    const bool narrow = s->narrow_rows;
    const unsigned fwd_grid = (n + FWD_THREADS - 1) / FWD_THREADS;
    B200_LAUNCH(s->prof, "fwd_kernel", st,
        (fwd_variant(narrow, s->fwd_minb, ix.kt_K > 0)<<<fwd_grid, FWD_THREADS, 0, st>>>(ix, d_packed, d_woff, d_len, n, p->min_seed_len, s->cand_stride, s->d_cand, s->d_ncand, s->d_stats)));
    B200_LAUNCH(s->prof, "back_kernel", st,
        (back_variant(narrow, s->back_minb, ix.kt_K > 0)<<<s->back_grid, BACK_THREADS, 0, st>>>(ix, d_packed, d_woff, n, p->min_seed_len, p->max_occ, s->cand_stride,
                                                                                  s->d_cand, s->d_ncand, s->d_nsmems, s->d_nseeds, s->d_env, s->env_stride, s->d_counters + 0, s->d_stats)));
    s->launches += 2;
    s->cur_packed = d_packed; s->cur_woff = d_woff; s->cur_len = d_len; s->cur_max_len = max_len;
    if (p->reseed) {
        rc = seeder_ensure_reseed(s, n_reads, max_len, 0, 0);
        if (rc) return rc;
        rc = seeder_reseed(s);
        if (rc) return rc;
    }
    rc = seeder_scan(s);
    if (rc) return rc;
    if (s->seed_cap > 0) {                 // output arrays exist: keep going without a host round trip
        rc = seeder_fill_locate(s);
        if (rc) return rc;
        s->filled = true;
    }
    return BWA_B200_OK;
}

// read the seed total back; if the arrays were too small (or absent) grow them and redo fill+locate
int b200_seeder_finish(bwa_b200_seeder *s)
{
    s->redone = false;
    if (s->last_n_reads == 0) return BWA_B200_OK;
    B200_CUDA(cudaMemcpyAsync(s->h_counters, s->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
    B200_CUDA(cudaStreamSynchronize(s->stream));
    while (s->last_p.reseed && (s->h_counters[4] > s->xstride || s->h_counters[5] > s->stride3)) {
        // a read produced more intervals / pass-2 candidates than a row holds: widen the rows and redo the re-seeding passes (rare)
        int rc = seeder_ensure_reseed(s, s->last_n_reads, s->cur_max_len, (uint32_t)s->h_counters[4] + 8, (uint32_t)s->h_counters[5] + 8);
        if (rc) return rc;
        rc = seeder_reseed(s);
        if (rc) return rc;
        rc = seeder_scan(s);
        if (rc) return rc;
        s->filled = false; s->redone = true;
        B200_CUDA(cudaMemcpyAsync(s->h_counters, s->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s->stream));
        B200_CUDA(cudaStreamSynchronize(s->stream));
    }
    s->last_total = s->h_counters[2];
    if (s->filled && s->last_total <= s->seed_cap) return BWA_B200_OK;
    s->redone = true;
    int rc = seeder_ensure_out(s, s->last_total);
    if (rc) return rc;
    if (s->last_total) {
        B200_CUDA(cudaMemsetAsync(s->d_counters + 1, 0, sizeof(unsigned long long), s->stream));
        rc = seeder_fill_locate(s);
        if (rc) return rc;
    }
    s->filled = true;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_seed_device(bwa_b200_seeder_t *s, const uint32_t *dev_packed, const uint64_t *dev_word_off,
                                    const uint32_t *dev_read_len, uint64_t n_reads, const bwa_b200_seed_params_t *p)
{
    if (!s || !p || (n_reads && (!dev_packed || !dev_word_off || !dev_read_len))) { b200::set_error("seed_device: bad argument"); return BWA_B200_ERR_ARG; }
    // the longest read bounds the candidate stride; reduce it on device
    uint32_t max_len = 0;
    if (n_reads) {
        B200_CUDA(cudaSetDevice(s->device));
        size_t tmp = 0;
        uint32_t *d_max = (uint32_t *)(s->d_counters + 3);
        B200_CUDA(cub::DeviceReduce::Max(nullptr, tmp, dev_read_len, d_max, (int)n_reads, s->stream));
        if (tmp > s->cub_bytes) { cudaFree(s->d_cub); s->d_cub = nullptr; B200_CUDA(cudaMalloc(&s->d_cub, tmp + 16)); s->cub_bytes = tmp; }
        B200_CUDA(cub::DeviceReduce::Max(s->d_cub, tmp, dev_read_len, d_max, (int)n_reads, s->stream));
        B200_CUDA(cudaMemcpyAsync(s->h_counters + 3, d_max, 4, cudaMemcpyDeviceToHost, s->stream));
        B200_CUDA(cudaStreamSynchronize(s->stream));
        max_len = *(uint32_t *)(s->h_counters + 3);
        s->launches += 1;
    }
    int rc = b200_seeder_run(s, dev_packed, dev_word_off, dev_read_len, n_reads, max_len, p);
    if (rc) return rc;
    return b200_seeder_finish(s);
}

extern "C" int bwa_b200_seed_device_result(bwa_b200_seeder_t *s, bwa_b200_seeds_t *v)
{
    if (!s || !v) return BWA_B200_ERR_ARG;
    B200_CUDA(cudaSetDevice(s->device));
    B200_CUDA(cudaStreamSynchronize(s->stream));
    v->n_reads = s->last_n_reads; v->n_seeds = s->last_total;
    v->rbeg = s->d_rbeg; v->qbeg_qend = (int32_t *)s->d_qq; v->score = s->d_score;
    v->n_seeds_per_read = s->d_nseeds; v->seed_off = s->d_seed_off;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_seed_host(bwa_b200_seeder_t *s, const uint32_t *packed, const uint64_t *word_off,
                                  const uint32_t *read_len, uint64_t n_reads, const bwa_b200_seed_params_t *p,
                                  bwa_b200_seeds_t *out)
{
    if (!s || !p || !out || (n_reads && (!packed || !word_off || !read_len))) { b200::set_error("seed_host: bad argument"); return BWA_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    out->n_reads = n_reads;
    if (n_reads > s->max_reads) { b200::set_error("seed_host: %llu reads > capacity %llu", (unsigned long long)n_reads, (unsigned long long)s->max_reads); return BWA_B200_ERR_CAPACITY; }
    uint64_t n_words = n_reads ? word_off[n_reads] : 0;
    if (n_words > s->max_words) { b200::set_error("seed_host: %llu words > capacity %llu", (unsigned long long)n_words, (unsigned long long)s->max_words); return BWA_B200_ERR_CAPACITY; }
    B200_CUDA(cudaSetDevice(s->device));
    uint32_t max_len = 0;
    for (uint64_t r = 0; r < n_reads; ++r) max_len = read_len[r] > max_len ? read_len[r] : max_len;
    if (n_reads) {
        B200_CUDA(cudaMemcpyAsync(s->d_packed, packed, n_words * 4, cudaMemcpyHostToDevice, s->stream));
        B200_CUDA(cudaMemcpyAsync(s->d_woff, word_off, (n_reads + 1) * 8, cudaMemcpyHostToDevice, s->stream));
        B200_CUDA(cudaMemcpyAsync(s->d_len, read_len, n_reads * 4, cudaMemcpyHostToDevice, s->stream));
    }
    int rc = b200_seeder_run(s, s->d_packed, s->d_woff, s->d_len, n_reads, max_len, p);
    if (rc) return rc;
    rc = b200_seeder_finish(s);
    if (rc) return rc;
    const uint64_t tot = s->last_total;
    out->n_seeds = tot;
    out->rbeg = (uint64_t *)malloc((tot ? tot : 1) * 8);
    out->qbeg_qend = (int32_t *)malloc((tot ? tot : 1) * 8);
    out->score = (uint32_t *)malloc((tot ? tot : 1) * 4);
    out->n_seeds_per_read = (uint32_t *)malloc((n_reads ? n_reads : 1) * 4);
    out->seed_off = (uint64_t *)malloc((n_reads ? n_reads : 1) * 8);
    if (!out->rbeg || !out->qbeg_qend || !out->score || !out->n_seeds_per_read || !out->seed_off) { bwa_b200_seeds_free(out); b200::set_error("seed_host: out of host memory"); return BWA_B200_ERR_NOMEM; }
    if (n_reads) {
        if (tot) {
            B200_CUDA(cudaMemcpyAsync(out->rbeg, s->d_rbeg, tot * 8, cudaMemcpyDeviceToHost, s->stream));
            B200_CUDA(cudaMemcpyAsync(out->qbeg_qend, s->d_qq, tot * 8, cudaMemcpyDeviceToHost, s->stream));
            B200_CUDA(cudaMemcpyAsync(out->score, s->d_score, tot * 4, cudaMemcpyDeviceToHost, s->stream));
        }
        B200_CUDA(cudaMemcpyAsync(out->n_seeds_per_read, s->d_nseeds, n_reads * 4, cudaMemcpyDeviceToHost, s->stream));
        B200_CUDA(cudaMemcpyAsync(out->seed_off, s->d_seed_off, n_reads * 8, cudaMemcpyDeviceToHost, s->stream));
        B200_CUDA(cudaStreamSynchronize(s->stream));
    }
    return BWA_B200_OK;
}

extern "C" void bwa_b200_seeds_free(bwa_b200_seeds_t *r)
{
    if (!r) return;
    free(r->rbeg); free(r->qbeg_qend); free(r->score); free(r->n_seeds_per_read); free(r->seed_off);
    memset(r, 0, sizeof(*r));
}

extern "C" int bwa_b200_seed_device_smems(bwa_b200_seeder_t *s, uint64_t n_reads, uint32_t *host_n_smems,
                                          int32_t *host_qbeg, int32_t *host_qend, uint64_t *host_k, uint64_t *host_s,
                                          uint64_t cap, uint64_t *total)
{
    if (!s || n_reads != s->last_n_reads) { b200::set_error("seed_device_smems: no matching batch"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(s->device));
    if (n_reads == 0) { if (total) *total = 0; return BWA_B200_OK; }
    const uint32_t n = (uint32_t)n_reads;
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t *> it(s->d_nsmems, U32ToU64());
    size_t tmp = s->cub_bytes;
    B200_CUDA(cub::DeviceScan::ExclusiveSum(s->d_cub, tmp, it, s->d_smem_off, (int)n, s->stream));
    total_kernel<<<1, 1, 0, s->stream>>>(s->d_nsmems, s->d_smem_off, n, s->d_counters + 3);
    B200_CUDA(cudaMemcpyAsync(s->h_counters + 3, s->d_counters + 3, 8, cudaMemcpyDeviceToHost, s->stream));
    B200_CUDA(cudaStreamSynchronize(s->stream));
    uint64_t tot = s->h_counters[3];
    if (total) *total = tot;
    B200_CUDA(cudaMemcpy(host_n_smems, s->d_nsmems, n_reads * 4, cudaMemcpyDeviceToHost));
    if (tot == 0) return BWA_B200_OK;
    if (tot > cap) { b200::set_error("seed_device_smems: %llu SMEMs > cap", (unsigned long long)tot); return BWA_B200_ERR_CAPACITY; }
    int32_t *d_qb, *d_qe; uint64_t *d_k, *d_s;
    B200_CUDA(cudaMalloc(&d_qb, tot * 4)); B200_CUDA(cudaMalloc(&d_qe, tot * 4));
    B200_CUDA(cudaMalloc(&d_k, tot * 8)); B200_CUDA(cudaMalloc(&d_s, tot * 8));
    const bool rs = s->last_p.reseed != 0;
    smem_dump_kernel<<<(n + 127) / 128, 128, 0, s->stream>>>(n, rs ? s->xstride : s->cand_stride, rs ? s->d_cand2 : s->d_cand, rs ? s->d_ncand2 : s->d_ncand,
                                                             s->d_nsmems, s->d_smem_off, d_qb, d_qe, d_k, d_s, tot);
    B200_CUDA(cudaStreamSynchronize(s->stream));
    B200_CUDA(cudaMemcpy(host_qbeg, d_qb, tot * 4, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(host_qend, d_qe, tot * 4, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(host_k, d_k, tot * 8, cudaMemcpyDeviceToHost));
    B200_CUDA(cudaMemcpy(host_s, d_s, tot * 8, cudaMemcpyDeviceToHost));
    cudaFree(d_qb); cudaFree(d_qe); cudaFree(d_k); cudaFree(d_s);
    return BWA_B200_OK;
}

// measured random-sector read throughput in GB/s over a buffer of `bytes` (0 on failure): best of `reps` launches after one
// warm-up launch; `iters` loads per lane (rounded up to a multiple of 8) -- 2048 gives about 40 GB of traffic per launch
extern "C" double bwa_b200_measure_random_sector_gbs(int device, uint64_t bytes, int iters, int reps)
{
    if (cudaSetDevice(device) != cudaSuccess) return 0;
    if (const char *ev = getenv("BWA_B200_L2_FETCH")) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(ev));
    uint32_t *buf = nullptr, *sink = nullptr;
    if (bytes < 4096 || cudaMalloc(&buf, bytes) != cudaSuccess) { cudaGetLastError(); return 0; }
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    int blocks = prop.multiProcessorCount * 8;
    iters = (iters + RS_UNROLL - 1) / RS_UNROLL * RS_UNROLL;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(a);
        random_sector_kernel<<<blocks, 256>>>(buf, bytes / 32, iters, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        double gbs = (double)blocks * 256 * iters * 32.0 / (ms * 1e-3) / 1e9;
        if (r > 0 && gbs > best) best = gbs;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    cudaFree(buf); cudaFree(sink);
    return best;
}
