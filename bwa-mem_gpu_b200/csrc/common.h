// common.h -- shared declarations of the B200 hot-path library (internal).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <cuda_runtime.h>
#include "bwamem_b200.h"

namespace b200 {

void set_error(const char *fmt, ...);

#define B200_CUDA(call)                                                                     \
    do {                                                                                    \
        cudaError_t e__ = (call);                                                           \
        if (e__ != cudaSuccess) {                                                           \
            b200::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return BWA_B200_ERR_CUDA;                                                       \
        }                                                                                   \
    } while (0)

// Device view of the FMD index, passed by value to kernels (lives in the constant bank).
struct IndexView {
    const uint32_t *bkt;     // 32-byte buckets per 64 BWT symbols: {cnt[4], L_lo, L_hi, H_lo, H_hi} (bit planes, see index.cu)
    const uint32_t *sa;      // low 32 bits of the sampled suffix array
    const uint32_t *sa_hi;   // packed high bits
    uint64_t primary, seq_len;
    uint64_t L2[5];
    uint32_t sa_shift;       // log2(sa_intv)
    uint32_t pack_size, pack_mask;
    // k-mer interval table (seed.cu, "k-mer table"): the suffix-array interval {first row k, size s} of every pattern of 1 .. kt_K
    // bases, one u64 per pattern = k << 24 | min(s, 0xffffff); level m starts at entry (4^m - 4) / 3 and is indexed by the pattern
    // with its first base most significant.  kt_K = 0: no table.  bkt_evict_last: L2 policy of the bucket loads (small indexes only).
    // Levels below kt_lo hold at least one pattern whose size does not fit the entry; steps landing there take the bucket path (the
    // choice depends on the level only, never on loaded data, so a warp issues its table loads and its bucket loads together).
    const uint64_t *kt;
    uint32_t kt_K, kt_lo, bkt_evict_last;
};

} // namespace b200

struct bwa_b200_index {
    int device = 0;
    b200::IndexView v{};
    uint32_t *d_bkt = nullptr, *d_sa = nullptr, *d_sa_hi = nullptr;
    uint32_t *d_pac = nullptr;      // optional 2-bit forward reference (16 bases per word)
    uint64_t l_pac = 0;
    uint64_t n_words = 0, n_sa = 0, n_hi = 0;
    int sa_intv = 0, pack_size = 0;
    uint64_t *d_kt = nullptr;       // k-mer interval table (optional)
    int kt_K = 0, kt_lo = 0;
};

// seed.cu: (re)build the k-mer interval table of a resident index on its device; K = 0 drops it
int b200_index_build_kmer_table(bwa_b200_index *idx, int K);
