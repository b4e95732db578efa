// compat_gpuseed.cu -- the reference's GPUSeed entry points (src/GPUSeed/seed_gen.h:92-106),
// implemented over the B200 seeding path.  Error behaviour mirrors the reference: a missing file
// or an inconsistent SA prints a message and exits (seed_gen.cu:1391-1418,1446-1449).
#include "common.h"
#include "seed_gen.h"
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

std::mutex g_mu;
std::map<const uint32_t *, bwa_b200_index_t *> g_indices;   // keyed on bwt_gpu.bwt (device pointer)
thread_local int t_device = 0;
thread_local int t_max_occ = 0;

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "[b200 GPUSeed] %s: %s\n", what, bwa_b200_last_error());
    exit(EXIT_FAILURE);
}

} // namespace

extern "C" void gpuseed_b200_set_device(int device) { t_device = device; }
extern "C" void gpuseed_b200_set_max_occ(int max_occ) { t_max_occ = max_occ; }
thread_local int t_reseed = 0, t_split_width = 10, t_max_mem_intv = 20;
thread_local float t_split_factor = 1.5f;
extern "C" void gpuseed_b200_set_reseed(int enable, float split_factor, int split_width, int max_mem_intv)
{
    t_reseed = enable != 0; t_split_factor = split_factor; t_split_width = split_width; t_max_mem_intv = max_mem_intv;
}

extern "C" bwt_t_gpu *bwt_restore_bwt_gpu(const char *fn)
{
    FILE *fp = fopen(fn, "rb");
    if (fp == NULL) { fprintf(stderr, "Unable to open .bwt file.\n"); exit(1); }
    bwt_t_gpu *bwt = (bwt_t_gpu *)calloc(1, sizeof(bwt_t_gpu));
    fseek(fp, 0, SEEK_END);
    bwt->bwt_size = (bwtint_t_gpu)(ftell(fp) - 5 * (long)sizeof(bwtint_t_gpu)) >> 2;
    fseek(fp, 0, SEEK_SET);
    bwt->bwt = (uint32_t *)bwa_b200_host_alloc(bwt->bwt_size * sizeof(uint32_t));
    if (!bwt->bwt) bwt->bwt = (uint32_t *)malloc(bwt->bwt_size * sizeof(uint32_t));   // no device yet: pageable
    bwt->L2 = (bwtint_t_gpu *)calloc(5, sizeof(bwtint_t_gpu));
    bool ok = fread(&bwt->primary, sizeof(bwtint_t_gpu), 1, fp) == 1 && fread(bwt->L2 + 1, sizeof(bwtint_t_gpu), 4, fp) == 4 &&
              fread(bwt->bwt, 4, bwt->bwt_size, fp) == bwt->bwt_size;
    fclose(fp);
    if (!ok) { fprintf(stderr, "Unable to read .bwt file.\n"); exit(1); }
    bwt->seq_len = bwt->L2[4];
    return bwt;
}

extern "C" void bwt_restore_sa_gpu(const char *fn, bwt_t_gpu *bwt)
{
    FILE *fp = fopen(fn, "rb");
    if (fp == NULL) { fprintf(stderr, "Unable to open .sa file.\n"); exit(1); }
    uint64_t hdr[7];
    if (fread(hdr, 8, 7, fp) != 7) { fprintf(stderr, "Unable to read .sa file.\n"); exit(1); }
    if (hdr[0] != bwt->primary) { fprintf(stderr, "SA-BWT inconsistency: primary is not the same.\n"); exit(EXIT_FAILURE); }
    if (hdr[6] != bwt->seq_len) { fprintf(stderr, "SA-BWT inconsistency: seq_len is not the same.\n"); exit(EXIT_FAILURE); }
    bwt->sa_intv = (int)hdr[5];
    bwt->n_sa = (bwt->seq_len + bwt->sa_intv) / bwt->sa_intv;
    bwt->sa = (uint32_t *)malloc(bwt->n_sa * sizeof(uint32_t));
    bwt->sa[0] = (uint32_t)-1;
    bool ok = fread(bwt->sa + 1, sizeof(uint32_t), bwt->n_sa - 1, fp) == bwt->n_sa - 1 && fread(&bwt->pack_size, 1, 1, fp) == 1;
    if (!ok) { fprintf(stderr, "Unable to read .sa file.\n"); exit(1); }
    size_t n_hi = (size_t)bwt->pack_size * bwt->n_sa / 32 + 1;
    bwt->sa_upper_bits = (uint32_t *)calloc(n_hi, sizeof(uint32_t));
    size_t got = fread(bwt->sa_upper_bits, sizeof(uint32_t), n_hi, fp); (void)got;
    fclose(fp);
}

extern "C" void bwt_destroy_gpu(bwt_t_gpu *bwt)
{
    if (bwt == 0) return;
    free(bwt->sa); free(bwt->sa_upper_bits); free(bwt->L2);
    cudaPointerAttributes at;
    if (bwt->bwt && cudaPointerGetAttributes(&at, bwt->bwt) == cudaSuccess && at.type == cudaMemoryTypeHost) bwa_b200_host_free(bwt->bwt);
    else { cudaGetLastError(); free(bwt->bwt); }
    free(bwt);
}

extern "C" bwt_t_gpu gpu_cpy_wrapper(bwt_t_gpu *bwt)
{
    bwa_b200_index_t *idx = nullptr;
    int rc = bwa_b200_index_from_host(bwt->primary, bwt->L2, bwt->bwt, bwt->bwt_size, bwt->sa, bwt->sa_upper_bits, bwt->n_sa,
                                      bwt->sa_intv, bwt->pack_size, t_device, &idx);
    if (rc) die("gpu_cpy_wrapper");
    bwt_t_gpu g;
    memset(&g, 0, sizeof(g));
    g.primary = bwt->primary; g.seq_len = bwt->seq_len; g.bwt_size = bwt->bwt_size; g.sa_intv = bwt->sa_intv;
    g.n_sa = bwt->n_sa; g.pack_size = bwt->pack_size;
    g.bwt = idx->d_bkt; g.sa = idx->d_sa; g.sa_upper_bits = idx->d_sa_hi;
    g.L2 = nullptr;                                  // the reference keeps L2 in __constant__ memory
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_indices[g.bwt] = idx;
    }
    bwt_destroy_gpu(bwt);                            // as the reference does (seed_gen.cu:1553)
    return g;
}

extern "C" void pre_calc_seed_intervals_wrapper(uint2 *, int, bwt_t_gpu) { /* unused by gase_aln (flag = 0, src/fastmap.c:455) */ }

extern "C" void free_gpuseed_data(gpuseed_storage_vector *d)
{
    if (!d) return;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_indices.find(d->bwt_gpu.bwt);
    if (it != g_indices.end()) { bwa_b200_index_free(it->second); g_indices.erase(it); }
    d->bwt_gpu.bwt = d->bwt_gpu.sa = d->bwt_gpu.sa_upper_bits = nullptr;
}

// one FASTA line = one read (seed_gen.cu:1698-1728); header lines start with '>'
extern "C" mem_seed_v_gpu *seed_gpu(gpuseed_storage_vector *d)
{
    if (!d->is_smem) {
        fprintf(stderr, "[b200 GPUSeed] MEM mode (is_smem = 0, `-g`) is not provided by the B200 library; use SMEM seeding.\n");
        exit(EXIT_FAILURE);
    }
    bwa_b200_index_t *idx = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_indices.find(d->bwt_gpu.bwt);
        if (it != g_indices.end()) idx = it->second;
    }
    if (!idx) { fprintf(stderr, "[b200 GPUSeed] seed_gpu: bwt_gpu was not produced by gpu_cpy_wrapper\n"); exit(EXIT_FAILURE); }
    FILE *fp = fopen(d->read_file, "r");
    if (!fp) { fprintf(stderr, "[b200 GPUSeed] cannot open %s\n", d->read_file); exit(EXIT_FAILURE); }
    fseek(fp, (long)d->file_bytes_skip, SEEK_SET);

    mem_seed_v_gpu *res = (mem_seed_v_gpu *)calloc(1, sizeof(mem_seed_v_gpu));
    std::vector<uint32_t> per_read;
    uint64_t n_seeds = 0, cap_seeds = 0, file_bytes = 0;
    const uint64_t BATCH_BASES = 64ull << 20;
    bwa_b200_seeder_t *seeder = nullptr;
    uint64_t seeder_reads = 0, seeder_words = 0;
    std::vector<char> bases;
    std::vector<uint64_t> off;
    char *line = nullptr;
    size_t line_cap = 0;
    bool done = false;
    bwa_b200_seed_params_t sp{d->min_seed_size, t_max_occ, t_reseed, t_split_factor, t_split_width, t_max_mem_intv};
    while (!done) {
        bases.clear(); off.assign(1, 0);
        while (bases.size() < BATCH_BASES) {
            ssize_t got = getline(&line, &line_cap, fp);
            if (got < 0) { done = true; break; }
            file_bytes += (uint64_t)got;
            if (line[0] == '>') continue;
            size_t len = (size_t)got;
            while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) --len;
            bases.insert(bases.end(), line, line + len);
            off.push_back(bases.size());
        }
        uint64_t n = off.size() - 1;
        if (n == 0) continue;
        std::vector<uint32_t> rl(n);
        std::vector<uint64_t> woff(n + 1);
        uint64_t words = 0;
        for (uint64_t r = 0; r < n; ++r) words += (off[r + 1] - off[r] + 7) / 8;
        std::vector<uint32_t> packed(words ? words : 1);
        if (bwa_b200_pack_ascii(bases.data(), off.data(), n, packed.data(), woff.data(), rl.data(), 0)) die("seed_gpu/pack");
        if (!seeder || n > seeder_reads || words > seeder_words) {
            if (seeder) bwa_b200_seeder_destroy(seeder);
            seeder_reads = n; seeder_words = words ? words : 1;
            if (bwa_b200_seeder_create(idx, seeder_reads, seeder_words, &seeder)) die("seed_gpu/seeder_create");
        }
        bwa_b200_seeds_t out;
        if (bwa_b200_seed_host(seeder, packed.data(), woff.data(), rl.data(), n, &sp, &out)) die("seed_gpu");
        if (n_seeds + out.n_seeds > cap_seeds) {
            cap_seeds = (n_seeds + out.n_seeds) * 3 / 2 + 1024;
            res->rbeg = (bwtint_t_gpu *)realloc(res->rbeg, cap_seeds * sizeof(bwtint_t_gpu));
            res->qbeg = (int2 *)realloc(res->qbeg, cap_seeds * sizeof(int2));
            res->score = (uint32_t *)realloc(res->score, cap_seeds * sizeof(uint32_t));
            if (!res->rbeg || !res->qbeg || !res->score) { fprintf(stderr, "Realloc rbeg error\n"); exit(EXIT_FAILURE); }
        }
        memcpy(res->rbeg + n_seeds, out.rbeg, out.n_seeds * sizeof(uint64_t));
        memcpy(res->qbeg + n_seeds, out.qbeg_qend, out.n_seeds * sizeof(int2));
        memcpy(res->score + n_seeds, out.score, out.n_seeds * sizeof(uint32_t));
        per_read.insert(per_read.end(), out.n_seeds_per_read, out.n_seeds_per_read + n);
        n_seeds += out.n_seeds;
        bwa_b200_seeds_free(&out);
    }
    free(line);
    fclose(fp);
    if (seeder) bwa_b200_seeder_destroy(seeder);
    size_t nr = per_read.size();
    res->n_ref_pos_fow_rev_results = (uint32_t *)malloc((nr ? nr : 1) * sizeof(uint32_t));
    res->n_ref_pos_fow_rev_prefix_sums = (uint32_t *)malloc((nr ? nr : 1) * sizeof(uint32_t));
    uint32_t run = 0;
    for (size_t r = 0; r < nr; ++r) { res->n_ref_pos_fow_rev_results[r] = per_read[r]; res->n_ref_pos_fow_rev_prefix_sums[r] = run; run += per_read[r]; }
    if (!res->rbeg) {            // empty file: still hand back freeable arrays
        res->rbeg = (bwtint_t_gpu *)malloc(8); res->qbeg = (int2 *)malloc(8); res->score = (uint32_t *)malloc(4);
    }
    res->file_bytes_skip = d->file_bytes_skip + file_bytes;
    return res;
}
