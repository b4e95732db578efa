// pack.cu -- host-side packing of reads into the 4-bit wire/HBM layout.
//
// The reference ships one byte per base over PCIe and packs on the GPU (pack_4bit_fow,
// seed_gen.cu:1088-1108; gasal_pack_kernel, GASAL2/src/kernels/pack_rc_seqs.h:13-53).  Here the
// C host code packs before the copy, halving H2D bytes.  Layout: 8 bases per u32, base 0 in
// bits 31..28; codes A0 C1 G2 T3, anything else 4 (nst_nt4_table, src/bntseq.c); every read
// starts on a word boundary and is padded with code 4.
#include "common.h"
#include <algorithm>
#include <thread>
#include <vector>

namespace {

struct Nt4 {
    uint8_t t[256];
    Nt4()
    {
        memset(t, 4, sizeof(t));
        t['A'] = t['a'] = 0; t['C'] = t['c'] = 1; t['G'] = t['g'] = 2; t['T'] = t['t'] = 3;
    }
};
const Nt4 g_nt4;

template <class F> void par(uint64_t n, int n_threads, F f)
{
    if (n_threads <= 1 || n < 1024) { f(0, n); return; }
    std::vector<std::thread> th;
    uint64_t chunk = (n + n_threads - 1) / n_threads;
    for (int t = 0; t < n_threads; ++t) {
        uint64_t a = std::min(n, chunk * t), b = std::min(n, a + chunk);
        if (a < b) th.emplace_back([=] { f(a, b); });
    }
    for (auto &x : th) x.join();
}

template <bool ASCII>
int pack_impl(const uint8_t *src, const uint64_t *base_off, uint64_t n_reads, uint32_t *packed, uint64_t *word_off,
              uint32_t *read_len, int n_threads)
{
    if (!src || !base_off || !packed || !word_off || !read_len) { b200::set_error("pack: null argument"); return BWA_B200_ERR_ARG; }
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    uint64_t w = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        uint64_t len = base_off[r + 1] - base_off[r];
        if (len > 0xffffffffull) { b200::set_error("pack: read %llu too long", (unsigned long long)r); return BWA_B200_ERR_ARG; }
        read_len[r] = (uint32_t)len;
        word_off[r] = w;
        w += (len + 7) / 8;
    }
    word_off[n_reads] = w;
    par(n_reads, n_threads, [=](uint64_t a, uint64_t b) {
        for (uint64_t r = a; r < b; ++r) {
            const uint8_t *s = src + base_off[r];
            uint32_t len = read_len[r];
            uint32_t *dst = packed + word_off[r];
            for (uint32_t i = 0; i < len; i += 8) {
                uint32_t word = 0;
                for (uint32_t j = 0; j < 8; ++j) {
                    uint32_t c = 4;
                    if (i + j < len) c = ASCII ? g_nt4.t[s[i + j]] : (s[i + j] > 3 ? 4u : s[i + j]);
                    word |= c << (28 - 4 * j);
                }
                dst[i >> 3] = word;
            }
        }
    });
    return BWA_B200_OK;
}

// the compact wire layout (bwa_b200_align_host_compact): 2 bits per base, 16 bases per word, base 0 in bits 31..30; anything that is
// not A/C/G/T is written as 0 and listed in n_list as (read << 32 | position)
template <bool ASCII>
int pack2_impl(const uint8_t *src, const uint64_t *base_off, uint64_t n_reads, uint32_t *packed2, uint32_t *read_len,
               uint64_t *n_list, uint64_t n_cap, uint64_t *n_n, int n_threads)
{
    if (!src || !base_off || !packed2 || !n_n || (n_cap && !n_list)) { b200::set_error("pack2: null argument"); return BWA_B200_ERR_ARG; }
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    std::vector<uint64_t> woff(n_reads + 1);
    uint64_t w = 0;
    for (uint64_t r = 0; r < n_reads; ++r) {
        const uint64_t len = base_off[r + 1] - base_off[r];
        if (len > 0xffffffffull) { b200::set_error("pack2: read %llu too long", (unsigned long long)r); return BWA_B200_ERR_ARG; }
        if (read_len) read_len[r] = (uint32_t)len;
        woff[r] = w;
        w += (len + 15) / 16;
    }
    woff[n_reads] = w;
    const int nt = n_threads;
    std::vector<std::vector<uint64_t>> found(nt);       // per thread, reads ascending: concatenation keeps the list sorted
    const uint64_t chunk = (n_reads + nt - 1) / nt;
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) {
        const uint64_t a = std::min(n_reads, chunk * t), b = std::min(n_reads, a + chunk);
        if (a >= b) continue;
        th.emplace_back([=, &found, &woff] {
            for (uint64_t r = a; r < b; ++r) {
                const uint8_t *s = src + base_off[r];
                const uint32_t len = (uint32_t)(base_off[r + 1] - base_off[r]);
                uint32_t *dst = packed2 + woff[r];
                for (uint32_t i = 0; i < len; i += 16) {
                    uint32_t word = 0;
                    for (uint32_t j = 0; j < 16 && i + j < len; ++j) {
                        uint32_t c = ASCII ? g_nt4.t[s[i + j]] : (uint32_t)s[i + j];
                        if (c > 3) { found[t].push_back(r << 32 | (uint64_t)(i + j)); c = 0; }
                        word |= c << (30 - 2 * j);
                    }
                    dst[i >> 4] = word;
                }
            }
        });
    }
    for (auto &x : th) x.join();
    uint64_t k = 0;
    for (int t = 0; t < nt; ++t) for (uint64_t e : found[t]) { if (k < n_cap) n_list[k] = e; ++k; }
    *n_n = k;
    if (k > n_cap) { b200::set_error("pack2: %llu bases are not A/C/G/T, room for %llu", (unsigned long long)k, (unsigned long long)n_cap); return BWA_B200_ERR_CAPACITY; }
    return BWA_B200_OK;
}

} // namespace

extern "C" size_t bwa_b200_packed2_words(const uint32_t *read_len, uint64_t n_reads, uint32_t uniform_len)
{
    if (!read_len) return (size_t)n_reads * (((size_t)uniform_len + 15) / 16);
    size_t w = 0;
    for (uint64_t r = 0; r < n_reads; ++r) w += ((size_t)read_len[r] + 15) / 16;
    return w;
}
extern "C" int bwa_b200_pack2_codes(const uint8_t *codes, const uint64_t *base_off, uint64_t n_reads, uint32_t *packed2, uint32_t *read_len,
                                    uint64_t *n_list, uint64_t n_cap, uint64_t *n_n, int n_threads)
{
    return pack2_impl<false>(codes, base_off, n_reads, packed2, read_len, n_list, n_cap, n_n, n_threads);
}
extern "C" int bwa_b200_pack2_ascii(const char *bases, const uint64_t *base_off, uint64_t n_reads, uint32_t *packed2, uint32_t *read_len,
                                    uint64_t *n_list, uint64_t n_cap, uint64_t *n_n, int n_threads)
{
    return pack2_impl<true>((const uint8_t *)bases, base_off, n_reads, packed2, read_len, n_list, n_cap, n_n, n_threads);
}

extern "C" size_t bwa_b200_packed_words(const uint32_t *read_len, uint64_t n_reads)
{
    size_t w = 0;
    for (uint64_t r = 0; r < n_reads; ++r) w += ((size_t)read_len[r] + 7) / 8;
    return w;
}

extern "C" int bwa_b200_pack_ascii(const char *bases, const uint64_t *base_off, uint64_t n_reads,
                                   uint32_t *packed, uint64_t *word_off, uint32_t *read_len, int n_threads)
{
    return pack_impl<true>((const uint8_t *)bases, base_off, n_reads, packed, word_off, read_len, n_threads);
}

extern "C" int bwa_b200_pack_codes(const uint8_t *codes, const uint64_t *base_off, uint64_t n_reads,
                                   uint32_t *packed, uint64_t *word_off, uint32_t *read_len, int n_threads)
{
    return pack_impl<false>(codes, base_off, n_reads, packed, word_off, read_len, n_threads);
}
