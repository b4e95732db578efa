// index.cu -- FMD index: file loader, HBM upload, peer replication, host-side builder.
//
// File formats are the reference's (bwa_index/bwt.c:461-487 writers; seed_gen.cu:1386-1468
// readers; bwa_index/bwtindex.c:174-197 bucket layout).  In HBM a bucket keeps the file's size and
// position -- 32 bytes = one DRAM/L2 sector per 64 BWT symbols, so every occurrence lookup costs
// one sector -- but its 2-bit symbols are re-arranged once, at upload, into two 64-bit bit planes
// {cnt[4], L_lo, L_hi, H_lo, H_hi}: bit p of L/H = low/high bit of symbol p.  "Occurrences among the
// first n symbols" then needs one 64-bit mask and 2 POPC per base instead of 4 masked words.
#include "common.h"
#include <algorithm>
#include <array>
#include <cstdarg>
#include <thread>
#include <vector>
#include <atomic>
#include <chrono>

namespace b200 {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
} // namespace b200

extern "C" const char *bwa_b200_last_error(void) { return b200::g_err; }
extern "C" int bwa_b200_version(void) { return 100; }

extern "C" int bwa_b200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" void *bwa_b200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void bwa_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

// in place: {cnt[4], sym[4]} (symbol i at bits (15-(i&15))*2 of word i>>4) -> {cnt[4], L_lo, L_hi, H_lo, H_hi}
__global__ void planes_kernel(uint32_t *bkt, uint64_t n_buckets)
{
    uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_buckets) return;
    uint32_t *w = bkt + b * 8 + 4;
    uint32_t s[4] = {w[0], w[1], w[2], w[3]};
    uint32_t L[2] = {0, 0}, H[2] = {0, 0};
    for (int p = 0; p < 64; ++p) {
        uint32_t v = (s[p >> 4] >> ((15 - (p & 15)) * 2)) & 3u;
        L[p >> 5] |= (v & 1u) << (p & 31);
        H[p >> 5] |= (v >> 1) << (p & 31);
    }
    w[0] = L[0]; w[1] = L[1]; w[2] = H[0]; w[3] = H[1];
}

static int ilog2(uint64_t x) { int r = 0; while ((1ull << r) < x) ++r; return r; }

// a half-built index is freed when a CUDA call fails on the way (B200_CUDA returns from the function)
struct IndexGuard {
    bwa_b200_index *p;
    ~IndexGuard() { if (p) bwa_b200_index_free(p); }
    bwa_b200_index *release() { bwa_b200_index *q = p; p = nullptr; return q; }
};

extern "C" int bwa_b200_index_from_host(uint64_t primary, const uint64_t L2[5], const uint32_t *bwt_words, uint64_t n_words,
                                        const uint32_t *sa, const uint32_t *sa_hi, uint64_t n_sa, int sa_intv, int pack_size,
                                        int device, bwa_b200_index_t **out)
{
    if (!L2 || !bwt_words || !out || n_words < 8) { b200::set_error("index_from_host: bad argument"); return BWA_B200_ERR_ARG; }
    if (sa && (sa_intv <= 0 || (sa_intv & (sa_intv - 1)))) { b200::set_error("SA interval %d is not a power of two", sa_intv); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(device));
    if (const char *ev = getenv("BWA_B200_L2_FETCH")) {      // experiment knob: L2 fetch granularity from DRAM (32 / 64 / 128 bytes)
        cudaError_t er = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(ev));
        size_t got = 0; cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        fprintf(stderr, "[b200] L2 fetch granularity -> %zu (%s)\n", got, cudaGetErrorString(er));
    }
    bwa_b200_index *idx = new bwa_b200_index();
    IndexGuard guard{idx};
    idx->device = device;
    idx->n_words = n_words;
    idx->n_sa = sa ? n_sa : 0;
    idx->sa_intv = sa_intv;
    idx->pack_size = pack_size;
    uint64_t seq_len = L2[4];
    // pad the bucket array to a whole bucket so 256-bit loads of the last one stay in bounds
    uint64_t padded = (n_words + 7) / 8 * 8 + 8;
    B200_CUDA(cudaMalloc(&idx->d_bkt, padded * 4));
    B200_CUDA(cudaMemset(idx->d_bkt, 0, padded * 4));
    B200_CUDA(cudaMemcpy(idx->d_bkt, bwt_words, n_words * 4, cudaMemcpyHostToDevice));
    {
        uint64_t n_buckets = (seq_len + 63) / 64;
        planes_kernel<<<(unsigned)((n_buckets + 255) / 256), 256>>>(idx->d_bkt, n_buckets);
        B200_CUDA(cudaGetLastError());
        B200_CUDA(cudaDeviceSynchronize());
    }
    if (sa) {
        B200_CUDA(cudaMalloc(&idx->d_sa, n_sa * 4));
        B200_CUDA(cudaMemcpy(idx->d_sa, sa, n_sa * 4, cudaMemcpyHostToDevice));
        idx->n_hi = (uint64_t)pack_size * n_sa / 32 + 1;
        B200_CUDA(cudaMalloc(&idx->d_sa_hi, idx->n_hi * 4));
        if (sa_hi) B200_CUDA(cudaMemcpy(idx->d_sa_hi, sa_hi, idx->n_hi * 4, cudaMemcpyHostToDevice));
        else B200_CUDA(cudaMemset(idx->d_sa_hi, 0, idx->n_hi * 4));
    }
    b200::IndexView &v = idx->v;
    v.bkt = idx->d_bkt; v.sa = idx->d_sa; v.sa_hi = idx->d_sa_hi;
    v.primary = primary; v.seq_len = seq_len;
    for (int i = 0; i < 5; ++i) v.L2[i] = L2[i];
    v.sa_shift = sa ? (uint32_t)ilog2((uint64_t)sa_intv) : 0;
    v.pack_size = (uint32_t)pack_size;
    // bwa_index/bwt.c:82-112: no high bits are kept when seq_len < 2^32
    v.pack_mask = (seq_len >> 32) == 0 ? 0u : (pack_size >= 32 ? 0xffffffffu : ((1u << pack_size) - 1));
    // bucket loads keep evict_last only while the bucket array is of the order of the L2 size; beyond, that priority goes to the k-mer table
    v.bkt_evict_last = padded * 4 <= (192ull << 20) ? 1u : 0u;
    {   // k-mer interval table (seed.cu).  Measured on B200 (profiles/r02_kmer_table_ab.txt): with an index of the order of the L2 size
        // the table only adds instructions to kernels that already issue 57 % of their cycles (100 Mb genome: back_kernel 4.0 -> 5.0 ms),
        // with the buckets in HBM it removes the sectors that bound them (1 Gb genome: 10.7 -> 7.5 ms at K = 13).  BWA_B200_KMER_K = 0 .. 14
        // overrides.
        const char *ev = getenv("BWA_B200_KMER_K");
        const uint64_t bkt_bytes = padded * 4;
        int K = ev ? atoi(ev) : (bkt_bytes <= (256ull << 20) ? 0 : (bkt_bytes < (1ull << 30) ? 12 : 13));
        while (K > 0 && (1ull << (2 * K)) > seq_len) --K;        // a tiny text does not need 4^K patterns
        int rc = b200_index_build_kmer_table(idx, K);
        if (rc) return rc;                                       // the guard frees the index
    }
    *out = guard.release();
    return BWA_B200_OK;
}

extern "C" int bwa_b200_index_set_kmer_table(bwa_b200_index_t *idx, int K)
{
    if (!idx || K < 0) { b200::set_error("index_set_kmer_table: bad argument"); return BWA_B200_ERR_ARG; }
    return b200_index_build_kmer_table(idx, K);
}

extern "C" int bwa_b200_index_load(const char *bwt_path, const char *sa_path, int device, bwa_b200_index_t **out)
{
    if (!bwt_path || !out) { b200::set_error("index_load: bad argument"); return BWA_B200_ERR_ARG; }
    FILE *fp = fopen(bwt_path, "rb");
    if (!fp) { b200::set_error("cannot open %s", bwt_path); return BWA_B200_ERR_IO; }
    fseek(fp, 0, SEEK_END);
    long fsz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    if (fsz < 40 + 32) { fclose(fp); b200::set_error("%s: too short", bwt_path); return BWA_B200_ERR_FORMAT; }
    uint64_t primary, L2[5] = {0, 0, 0, 0, 0};
    uint64_t n_words = (uint64_t)(fsz - 40) >> 2;
    std::vector<uint32_t> words(n_words);
    bool ok = fread(&primary, 8, 1, fp) == 1 && fread(L2 + 1, 8, 4, fp) == 4 && fread(words.data(), 4, n_words, fp) == n_words;
    fclose(fp);
    if (!ok) { b200::set_error("%s: short read", bwt_path); return BWA_B200_ERR_IO; }
    uint64_t seq_len = L2[4];
    // 32-bit-occ / 64-symbol layout has 4*ceil(n/64) + ceil(n/16) + 4 words
    uint64_t expect = 4 * ((seq_len + 63) / 64) + (seq_len + 15) / 16 + 4;
    if (n_words != expect) {
        b200::set_error("%s: %llu payload words, expected %llu for the 32-bit-occ 64-symbol bucket layout (seq_len %llu); "
                        "was the index built with `bwa index -s bwt` (OCC_INTV_SHIFT 6)?", bwt_path,
                        (unsigned long long)n_words, (unsigned long long)expect, (unsigned long long)seq_len);
        return BWA_B200_ERR_FORMAT;
    }
    if (!sa_path) return bwa_b200_index_from_host(primary, L2, words.data(), n_words, nullptr, nullptr, 0, 0, 0, device, out);
    fp = fopen(sa_path, "rb");
    if (!fp) { b200::set_error("cannot open %s", sa_path); return BWA_B200_ERR_IO; }
    uint64_t hdr[7];
    if (fread(hdr, 8, 7, fp) != 7) { fclose(fp); b200::set_error("%s: short header", sa_path); return BWA_B200_ERR_IO; }
    if (hdr[0] != primary) { fclose(fp); b200::set_error("SA-BWT inconsistency: primary is not the same."); return BWA_B200_ERR_FORMAT; }
    if (hdr[6] != seq_len) { fclose(fp); b200::set_error("SA-BWT inconsistency: seq_len is not the same."); return BWA_B200_ERR_FORMAT; }
    int sa_intv = (int)hdr[5];
    if (sa_intv <= 0 || (sa_intv & (sa_intv - 1))) { fclose(fp); b200::set_error("%s: bad SA interval", sa_path); return BWA_B200_ERR_FORMAT; }
    uint64_t n_sa = (seq_len + (uint64_t)sa_intv) / (uint64_t)sa_intv;
    std::vector<uint32_t> sa(n_sa);
    sa[0] = 0xffffffffu;
    uint8_t ps = 0;
    ok = fread(sa.data() + 1, 4, n_sa - 1, fp) == n_sa - 1 && fread(&ps, 1, 1, fp) == 1;
    if (!ok || ps == 0 || ps > 32) { fclose(fp); b200::set_error("%s: short read / bad pack_size", sa_path); return BWA_B200_ERR_FORMAT; }
    uint64_t n_hi = (uint64_t)ps * n_sa / 32 + 1;
    std::vector<uint32_t> hi(n_hi, 0);
    size_t got = fread(hi.data(), 4, n_hi, fp); (void)got;
    fclose(fp);
    return bwa_b200_index_from_host(primary, L2, words.data(), n_words, sa.data(), hi.data(), n_sa, sa_intv, ps, device, out);
}

extern "C" int bwa_b200_index_clone_to(const bwa_b200_index_t *src, int device, bwa_b200_index_t **out)
{
    if (!src || !out) { b200::set_error("index_clone_to: bad argument"); return BWA_B200_ERR_ARG; }
    B200_CUDA(cudaSetDevice(device));
    bwa_b200_index *idx = new bwa_b200_index(*src);
    IndexGuard guard{idx};
    idx->device = device;
    idx->d_bkt = idx->d_sa = idx->d_sa_hi = nullptr;
    idx->d_pac = nullptr;
    idx->d_kt = nullptr; idx->v.kt = nullptr;
    uint64_t padded = (src->n_words + 7) / 8 * 8 + 8;
    B200_CUDA(cudaMalloc(&idx->d_bkt, padded * 4));
    B200_CUDA(cudaMemcpyPeer(idx->d_bkt, device, src->d_bkt, src->device, padded * 4));
    if (src->d_sa) {
        B200_CUDA(cudaMalloc(&idx->d_sa, src->n_sa * 4));
        B200_CUDA(cudaMemcpyPeer(idx->d_sa, device, src->d_sa, src->device, src->n_sa * 4));
        B200_CUDA(cudaMalloc(&idx->d_sa_hi, src->n_hi * 4));
        B200_CUDA(cudaMemcpyPeer(idx->d_sa_hi, device, src->d_sa_hi, src->device, src->n_hi * 4));
    }
    idx->v.bkt = idx->d_bkt; idx->v.sa = idx->d_sa; idx->v.sa_hi = idx->d_sa_hi;
    if (src->d_kt) {
        const uint64_t bytes = (((1ull << (2 * (src->kt_K + 1))) - 4) / 3) * 8;
        B200_CUDA(cudaMalloc(&idx->d_kt, bytes));
        B200_CUDA(cudaMemcpyPeer(idx->d_kt, device, src->d_kt, src->device, bytes));
        idx->v.kt = idx->d_kt;
    }
    if (src->d_pac) {          // the attached 2-bit forward reference (bwa_b200_index_attach_ref) travels with the index
        const uint64_t words = (src->l_pac + 15) / 16 + 1;
        B200_CUDA(cudaMalloc(&idx->d_pac, words * 4));
        B200_CUDA(cudaMemcpyPeer(idx->d_pac, device, src->d_pac, src->device, words * 4));
    } else idx->l_pac = 0;
    *out = guard.release();
    return BWA_B200_OK;
}

extern "C" int bwa_b200_index_info(const bwa_b200_index_t *idx, bwa_b200_index_info_t *info)
{
    if (!idx || !info) return BWA_B200_ERR_ARG;
    info->primary = idx->v.primary; info->seq_len = idx->v.seq_len;
    for (int i = 0; i < 5; ++i) info->L2[i] = idx->v.L2[i];
    info->n_buckets = (idx->v.seq_len + 63) / 64;
    info->n_sa = idx->n_sa; info->sa_intv = idx->sa_intv; info->pack_size = idx->pack_size;
    info->device = idx->device;
    info->hbm_bytes = ((idx->n_words + 7) / 8 * 8 + 8) * 4 + idx->n_sa * 4 + idx->n_hi * 4 + (idx->kt_K ? (((1ull << (2 * (idx->kt_K + 1))) - 4) / 3) * 8 : 0);
    return BWA_B200_OK;
}

extern "C" void bwa_b200_index_free(bwa_b200_index_t *idx)
{
    if (!idx) return;
    cudaSetDevice(idx->device);
    cudaFree(idx->d_bkt); cudaFree(idx->d_sa); cudaFree(idx->d_sa_hi); cudaFree(idx->d_pac); cudaFree(idx->d_kt);
    delete idx;
}

// ------------------------------------------------------------------------------------------
// Host-side index construction.  Suffix array of T = fwd + revcomp(fwd) by a 7-mer partition
// followed by per-partition sorts keyed on a 2-bit packed copy of T (32 bases per 64-bit
// window), then BWT, occurrence buckets and SA samples in the reference's formats.
// Suffix indexes are 32-bit while 2*l_pac < 2^32 - 1 and 64-bit beyond (human-sized genomes:
// the SA samples then carry their high bits in the packed array of bwa_index/bwt.c:78-147).
// Every pass over the rows is spread over the host threads: the BWT pass is one random read of
// the text per row, which a single thread cannot feed at 6 G rows.
// ------------------------------------------------------------------------------------------
namespace {

struct Packed {
    std::vector<uint64_t> w;   // base i at bits 62-2*(i%32) of word i/32; zero padded
    uint64_t n = 0;
    inline uint64_t window(uint64_t i) const
    { // 32 bases starting at i
        uint64_t q = i >> 5, r = (i & 31) * 2;
        uint64_t hi = w[q] << r;
        uint64_t lo = r ? (w[q + 1] >> (64 - r)) : 0;
        return hi | lo;
    }
    inline uint32_t base(uint64_t i) const { return (uint32_t)(w[i >> 5] >> (62 - 2 * (i & 31))) & 3u; }
    // suffix order with an implicit sentinel smaller than every base
    inline bool less(uint64_t a, uint64_t b) const
    {
        uint64_t lim = n - std::max(a, b);
        for (uint64_t off = 0; off < lim; off += 32) {
            uint64_t x = window(a + off), y = window(b + off);
            if (x != y) return x < y;
        }
        return a > b; // common prefix as long as the shorter one: the shorter suffix is smaller
    }
};

// f(first, last, thread) over [0, n) cut into n_threads pieces whose boundaries are multiples of `align`
template <class F> void parallel_for(uint64_t n, int n_threads, F f, uint64_t align = 1)
{
    if (n_threads <= 1 || n < 2) { f(0, n, 0); return; }
    std::vector<std::thread> th;
    uint64_t chunk = (n + n_threads - 1) / n_threads;
    chunk = (chunk + align - 1) / align * align;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t a = std::min(n, chunk * t), b = std::min(n, a + chunk);
        th.emplace_back([=] { f(a, b, t); });
    }
    for (auto &x : th) x.join();
}

int write_file(const std::string &path, const std::vector<std::pair<const void *, size_t>> &parts)
{
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) { b200::set_error("cannot write %s", path.c_str()); return BWA_B200_ERR_IO; }
    for (auto &p : parts)
        if (p.second && fwrite(p.first, 1, p.second, fp) != p.second) { fclose(fp); b200::set_error("short write %s", path.c_str()); return BWA_B200_ERR_IO; }
    fclose(fp);
    return BWA_B200_OK;
}

// bucket arrays: `cw` count words (u32 x 4 or u64 x 4 = 8 words) in front of every `sym` BWT symbols, one more set at the end
// (bwa_index/bwtindex.c:151-197).  raw = 16 symbols per word; pre[t] = counts before thread t's piece.
template <class CntT>
void bucket_layout(std::vector<uint32_t> &out, const std::vector<uint32_t> &raw, uint64_t n, uint64_t sym, int n_threads,
                   const std::vector<std::array<uint64_t, 4>> &pre, uint64_t piece)
{
    const uint64_t cw = 4 * sizeof(CntT) / 4, n_raw = (n + 15) / 16, wpb = sym / 16;
    out.assign(cw * ((n + sym - 1) / sym) + n_raw + cw, 0);
    parallel_for(n, n_threads, [&](uint64_t a, uint64_t b, int t) {
        if (a >= b) return;
        CntT c[4] = {(CntT)pre[t][0], (CntT)pre[t][1], (CntT)pre[t][2], (CntT)pre[t][3]};
        for (uint64_t i = a; i < b; i += 16) {                       // a is a multiple of `piece`, itself a multiple of sym
            uint64_t o = (i / sym) * (cw + wpb) + cw + (i % sym) / 16;
            if (i % sym == 0) memcpy(&out[o - cw], c, sizeof(c));
            const uint32_t v = raw[i >> 4];
            out[o] = v;
            const int m = (int)std::min<uint64_t>(16, n - i);
            for (int j = 0; j < m; ++j) ++c[(v >> ((15 - j) << 1)) & 3u];
        }
        if (b == n) memcpy(&out[out.size() - cw], c, sizeof(c));
    }, piece);
}

template <class IdxT>
int build_impl(const uint8_t *fwd, uint64_t l_pac, int sa_intv, const char *prefix, int also_stock_layout, int n_threads)
{
    const uint64_t n = 2 * l_pac;
    const bool verbose = getenv("BWA_B200_BUILD_TIMES") != nullptr;              // phase times on stderr
    auto t_last = std::chrono::steady_clock::now();
    auto phase = [&](const char *what) {
        auto now = std::chrono::steady_clock::now();
        if (verbose) fprintf(stderr, "[build_index] %-28s %.2f s\n", what, std::chrono::duration<double>(now - t_last).count());
        t_last = now;
    };
    // packed text, validation and base counts in one pass over the forward strand
    Packed P;
    P.n = n;
    P.w.assign(n / 32 + 3, 0);
    std::vector<std::array<uint64_t, 5>> cnt_t(n_threads, std::array<uint64_t, 5>{0, 0, 0, 0, 0});
    parallel_for(n, n_threads, [&](uint64_t a, uint64_t b, int t) {
        auto &c = cnt_t[t];
        for (uint64_t i = a; i < b; ++i) {
            uint32_t v = i < l_pac ? fwd[i] : 3u - fwd[n - 1 - i];
            if (i < l_pac && v > 3) { if (!c[4]) c[4] = i + 1; v = 0; }
            v &= 3u;
            ++c[v];
            P.w[i >> 5] |= (uint64_t)v << (62 - 2 * (i & 31));
        }
    }, 32);
    uint64_t L2[5] = {0, 0, 0, 0, 0};
    for (auto &c : cnt_t) {
        if (c[4]) { b200::set_error("build_index: code %d at %llu (only A,C,G,T)", fwd[c[4] - 1], (unsigned long long)(c[4] - 1)); return BWA_B200_ERR_ARG; }
        for (int k = 0; k < 4; ++k) L2[k + 1] += c[k];
    }
    for (int k = 0; k < 4; ++k)
        if (L2[k + 1] > 0xffffffffull) { b200::set_error("build_index: %llu occurrences of one base do not fit the 32-bit bucket counts", (unsigned long long)L2[k + 1]); return BWA_B200_ERR_CAPACITY; }
    for (int c = 1; c <= 4; ++c) L2[c] += L2[c - 1];
    phase("pack text");

    // Partition by the first K bases (zero padded past the end): per-thread histograms over contiguous pieces of the text, so the
    // scatter needs no atomics and leaves every partition in text order.  Each partition (n / 4^K suffixes, cache-sized) is then
    // sorted on its own: the 32 bases behind the shared K-mer travel with each suffix as the sort key, fetched once in text order,
    // and the text is touched again only on a tie.  Zero padding past the end sorts like the sentinel or like A; a tie falls back to
    // the full comparison, so the order is the strict suffix order and the result does not depend on the thread count.
    const int K = 7;
    const uint64_t NB = 1ull << (2 * K);
    auto key = [&](uint64_t i) { return P.window(i) >> (64 - 2 * K); };
    std::vector<uint64_t> start(NB + 1, 0);
    std::vector<IdxT> sa(n);
    {
        std::vector<uint64_t> hist((size_t)n_threads * NB, 0);
        parallel_for(n, n_threads, [&](uint64_t a, uint64_t b, int t) {
            uint64_t *h = &hist[(size_t)t * NB];
            for (uint64_t i = a; i < b; ++i) ++h[key(i)];
        });
        uint64_t run = 0;
        for (uint64_t k = 0; k < NB; ++k) {
            start[k] = run;
            for (int t = 0; t < n_threads; ++t) { uint64_t c = hist[(size_t)t * NB + k]; hist[(size_t)t * NB + k] = run; run += c; }
        }
        start[NB] = run;
        parallel_for(n, n_threads, [&](uint64_t a, uint64_t b, int t) {
            uint64_t *h = &hist[(size_t)t * NB];
            for (uint64_t i = a; i < b; ++i) sa[h[key(i)]++] = (IdxT)i;      // (gathering a cache line per partition first: no gain measured)
        });
        phase("k-mer partition");
    }
    {
        std::atomic<uint64_t> next(0);
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&] {
                typedef std::pair<uint64_t, IdxT> KeyIdx;
                std::vector<KeyIdx> buf, buf2;
                constexpr int RB = 12;                       // large partitions: one counting pass on the key's top 12 bits first
                std::vector<uint32_t> cnt((1u << RB) + 1);
                auto lt = [&](const KeyIdx &x, const KeyIdx &y) { return x.first != y.first ? x.first < y.first : P.less(x.second, y.second); };
                for (;;) {
                    const uint64_t k = next.fetch_add(1);
                    if (k >= NB) break;
                    const uint64_t a = start[k], m = start[k + 1] - a;
                    if (m < 2) continue;
                    buf.resize(m);
                    for (uint64_t e = 0; e < m; ++e) { const IdxT p = sa[a + e]; buf[e] = KeyIdx(P.window((uint64_t)p + K), p); }
                    if (m < 4096 || m > 0xffffffffull) {
                        std::sort(buf.begin(), buf.end(), lt);
                        for (uint64_t e = 0; e < m; ++e) sa[a + e] = buf[e].second;
                        continue;
                    }
                    std::fill(cnt.begin(), cnt.end(), 0u);
                    for (uint64_t e = 0; e < m; ++e) ++cnt[(buf[e].first >> (64 - RB)) + 1];
                    for (uint32_t d = 0; d < (1u << RB); ++d) cnt[d + 1] += cnt[d];
                    buf2.resize(m);
                    for (uint64_t e = 0; e < m; ++e) buf2[cnt[buf[e].first >> (64 - RB)]++] = buf[e];      // cnt[d] is now the END of digit d
                    uint32_t b0 = 0;
                    for (uint32_t d = 0; d < (1u << RB); ++d) {
                        const uint32_t b1 = cnt[d];
                        if (b1 - b0 > 1) std::sort(buf2.begin() + b0, buf2.begin() + b1, lt);
                        b0 = b1;
                    }
                    for (uint64_t e = 0; e < m; ++e) sa[a + e] = buf2[e].second;
                }
            });
        for (auto &x : th) x.join();
        phase("partition sorts");
    }
    // ---- BWT with '$' removed, primary.  Rows: 0 = empty suffix (char T[n-1]); row r >= 1 = sa[r-1]; the row whose suffix is
    // the whole text (primary) has no symbol, so symbol j >= 1 belongs to row j + (j >= primary)
    uint64_t primary = 0;
    parallel_for(n, n_threads, [&](uint64_t a, uint64_t b, int) {
        for (uint64_t r = a; r < b; ++r) if (sa[r] == 0) primary = r + 1;       // exactly one writer
    });
    const uint64_t n_raw = (n + 15) / 16;
    std::vector<uint32_t> raw(n_raw, 0);
    const uint64_t piece = 128 * (((n + n_threads - 1) / n_threads + 127) / 128);     // whole buckets of either layout per thread
    std::vector<std::array<uint64_t, 4>> pre(n_threads + 1, std::array<uint64_t, 4>{0, 0, 0, 0});
    parallel_for(n, n_threads, [&](uint64_t a, uint64_t b, int t) {
        uint64_t c[4] = {0, 0, 0, 0};
        for (uint64_t j = a; j < b; ++j) {
            uint32_t v;
            if (j == 0) v = P.base(n - 1);
            else v = P.base((uint64_t)sa[j - 1 + (j >= primary)] - 1);
            ++c[v];
            raw[j >> 4] |= v << ((~j & 15) << 1);
        }
        for (int k = 0; k < 4; ++k) pre[t + 1][k] = c[k];
    }, piece);
    for (int t = 1; t <= n_threads; ++t) for (int k = 0; k < 4; ++k) pre[t][k] += pre[t - 1][k];
    phase("bwt");

    int rc;
    {   // ---- GPU layout: u32 counts + 4 words per 64 symbols, trailing counts (bwtindex.c:174-197)
        std::vector<uint32_t> g;
        bucket_layout<uint32_t>(g, raw, n, 64, n_threads, pre, piece);
        rc = write_file(std::string(prefix) + ".bwt", {{&primary, 8}, {L2 + 1, 32}, {g.data(), g.size() * 4}});
        if (rc) return rc;
    }
    if (also_stock_layout) { // u64 counts + 8 words per 128 symbols (bwtindex.c:151-172)
        std::vector<uint32_t> s;
        bucket_layout<uint64_t>(s, raw, n, 128, n_threads, pre, piece);
        rc = write_file(std::string(prefix) + ".bwt128", {{&primary, 8}, {L2 + 1, 32}, {s.data(), s.size() * 4}});
        if (rc) return rc;
    }
    std::vector<uint32_t>().swap(raw);
    phase("bucket layouts + write");
    // ---- SA samples: low 32 bits, and the high bits packed `pack_size` to a sample (bwa_index/bwt.c:78-147, 472-487)
    const uint64_t n_sa = (n + (uint64_t)sa_intv) / (uint64_t)sa_intv;
    uint8_t pack_size = 1;
    uint32_t pack_mask = 0;
    {
        const uint32_t upper = (uint32_t)(n >> 32);
        int msb = 0;
        while (msb < 32 && (upper >> msb)) ++msb;                 // position of the highest set bit, 1-based
        if (msb == 1) { pack_size = 1; pack_mask = 1u; }
        else if (msb == 2) { pack_size = 2; pack_mask = 3u; }
        else if (msb > 2 && msb <= 4) { pack_size = 4; pack_mask = 0xfu; }
        else if (msb > 4 && msb <= 8) { pack_size = 8; pack_mask = 0xffu; }
        else if (msb > 8 && msb <= 16) { pack_size = 16; pack_mask = 0xffffu; }
        else if (msb > 16) { pack_size = 32; pack_mask = 0xffffffffu; }
    }
    const uint64_t pack_div = 32 / pack_size;
    std::vector<uint32_t> smp(n_sa);
    std::vector<uint32_t> hi((uint64_t)pack_size * n_sa / 32 + 1, 0);
    parallel_for(n_sa, n_threads, [&](uint64_t a, uint64_t b, int) {
        for (uint64_t j = std::max<uint64_t>(a, 1); j < b; ++j) {
            const uint64_t v = (uint64_t)sa[j * (uint64_t)sa_intv - 1];
            smp[j] = (uint32_t)v;
            if (pack_mask) hi[j / pack_div] |= ((uint32_t)(v >> 32) & pack_mask) << ((j % pack_div) * pack_size);
        }
    }, 32);
    smp[0] = 0xffffffffu;
    hi[0] |= pack_mask;                                            // row 0 reads as -1 (bwa_index/bwt.c:144-146)
    uint64_t intv64 = (uint64_t)sa_intv, seq_len = n;
    rc = write_file(std::string(prefix) + ".sa", {{&primary, 8}, {L2 + 1, 32}, {&intv64, 8}, {&seq_len, 8},
                                                  {smp.data() + 1, (n_sa - 1) * 4}, {&pack_size, 1}, {hi.data(), hi.size() * 4}});
    phase("sa samples + write");
    return rc;
}

} // namespace

extern "C" int bwa_b200_build_index(const uint8_t *fwd, uint64_t l_pac, int sa_intv, const char *prefix,
                                    int also_stock_layout, int n_threads)
{
    if (!fwd || !prefix || l_pac == 0 || sa_intv <= 0 || (sa_intv & (sa_intv - 1))) { b200::set_error("build_index: bad argument"); return BWA_B200_ERR_ARG; }
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const char *wide = getenv("BWA_B200_BUILD_WIDE");              // tests: 64-bit suffix indexes on a small text
    if (2 * l_pac >= 0xffffffffull || (wide && wide[0] == '1'))
        return build_impl<uint64_t>(fwd, l_pac, sa_intv, prefix, also_stock_layout, n_threads);
    return build_impl<uint32_t>(fwd, l_pac, sa_intv, prefix, also_stock_layout, n_threads);
}
