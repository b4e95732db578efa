/*
 * oracle/cpu_compat.cpp -- TEST INFRASTRUCTURE ONLY (never linked into libbwamem_b200.so, never timed as the product).
 *
 * The two library boundaries of the reference's `gase_aln` driver (include/compat/seed_gen.h = src/GPUSeed/seed_gen.h:92-106,
 * include/compat/gasal_b200_compat.h = GASAL2's ctors/host_batch/interfaces/res/gasal_align headers) implemented on the CPU
 * with the REFERENCE's own functions, so that the unmodified driver can be linked twice -- once over libbwamem_b200.so
 * (oracle/_ref/bwa-gasal2-b200) and once over this file (oracle/_ref/bwa-gasal2-cpu) -- and the two SAM files compared byte
 * for byte (tests/test_gpu_sam.py):
 *   seeding    bwt_smem1 + bwt_sa of the reference's CPU bwa (bwa_index/bwt.c:151-172,365-432), through
 *              oracle/_ref/libbwaref.so (dlopen'ed RTLD_LOCAL | RTLD_DEEPBIND: the driver binary carries the fork's own,
 *              different, functions of the same names).  Contract of seed_gpu: one FASTA line = one read
 *              (seed_gen.cu:1698-1728); every SMEM of length >= min_seed_size contributes ALL its rows, SMEMs in query
 *              order, `score` = occurrence count on the first row of a group (seed_gen.h:68-75, src/bwamem.c:415-431).
 *   extension  the driver's own ksw_extend2 (src/ksw.c:864-986, linked in the same binary) followed by the local-vs-to-end
 *              rule, exactly the loop of the driver's decoy_cpu_align (src/bwamem.c:1791-1907), with the parameter block of
 *              gasal_b200_set_ext_params (defaults = the fork's: no band, zdrop 0, end bonus = clip penalty = 5).
 * Host arrays are plain malloc memory; a "launch" runs synchronously and gasal_is_aln_async_done reports it finished.
 * The stock-layout BWT next to the GPU-layout one is expected at <prefix>.bwt128 (bwa_b200_build_index with
 * also_stock_layout, or `bwa7 index`'s .bwt copied there).
 */
#include <dlfcn.h>
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "gasal_b200_compat.h"
#include "seed_gen.h"

extern "C" int ksw_extend2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat, int o_del, int e_del,
                           int o_ins, int e_ins, int w, int end_bonus, int zdrop, int h0, int *qle, int *tle, int *gtle, int *gscore,
                           int *max_off, int opt_ext);

namespace {

struct RefLib {
    void *so = nullptr;
    void *(*load)(const char *, const char *) = nullptr;
    void (*release)(void *) = nullptr;
    int (*smem1)(void *, int, const uint8_t *, int, int, uint64_t *, int *) = nullptr;
    uint64_t (*sa)(void *, uint64_t) = nullptr;
} g_ref;

void ref_open()
{
    if (g_ref.so) return;
    const char *path = getenv("BWA_B200_LIBBWAREF");
    std::string p = path ? path : "";
    if (p.empty()) {                       // next to this binary: oracle/_ref/
        char exe[4096];
        ssize_t n = readlink("/proc/self/exe", exe, sizeof(exe) - 1);
        if (n > 0) { exe[n] = 0; p = exe; p = p.substr(0, p.rfind('/')) + "/libbwaref.so"; }
    }
    g_ref.so = dlopen(p.c_str(), RTLD_NOW | RTLD_LOCAL | RTLD_DEEPBIND);
    if (!g_ref.so) { fprintf(stderr, "[cpu_compat] cannot load %s: %s\n", p.c_str(), dlerror()); exit(EXIT_FAILURE); }
    g_ref.load = (void *(*)(const char *, const char *))dlsym(g_ref.so, "ref_load");
    g_ref.release = (void (*)(void *))dlsym(g_ref.so, "ref_free");
    g_ref.smem1 = (int (*)(void *, int, const uint8_t *, int, int, uint64_t *, int *))dlsym(g_ref.so, "ref_smem1");
    g_ref.sa = (uint64_t (*)(void *, uint64_t))dlsym(g_ref.so, "ref_sa");
    if (!g_ref.load || !g_ref.smem1 || !g_ref.sa) { fprintf(stderr, "[cpu_compat] libbwaref.so lacks the shim entry points\n"); exit(EXIT_FAILURE); }
}

std::string g_bwt_path, g_sa_path;
void *g_index = nullptr;

bwa_b200_ext_params_t g_params;
bool g_params_init = false;

void fill_scmat(int a, int b, int8_t mat[25])
{ // bwa_fill_scmat (src/bwa.c:83-96)
    int k = 0;
    for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) mat[k++] = i == j ? a : -b; mat[k++] = -1; }
    for (int j = 0; j < 5; ++j) mat[k++] = -1;
}

void params_init()
{
    if (g_params_init) return;
    memset(&g_params, 0, sizeof(g_params));
    fill_scmat(1, 4, g_params.mat);
    g_params.o_del = g_params.o_ins = 6; g_params.e_del = g_params.e_ins = 1;
    g_params.w = 300; g_params.zdrop = 0; g_params.use_band = 0; g_params.pen_clip = 5; g_params.end_bonus = 5;
    g_params_init = true;
}

template <class T> T *host(size_t n) { return (T *)calloc(n ? n : 1, sizeof(T)); }

} // namespace

// ============================================================================ GPUSeed boundary
extern "C" void gpuseed_b200_set_device(int) {}
static int t_max_occ = 0;
extern "C" void gpuseed_b200_set_max_occ(int max_occ) { t_max_occ = max_occ; }
extern "C" void gpuseed_b200_set_reseed(int enable, float, int, int)
{
    if (enable) { fprintf(stderr, "[cpu_compat] re-seeding is not part of this checker\n"); exit(EXIT_FAILURE); }
}

extern "C" bwt_t_gpu *bwt_restore_bwt_gpu(const char *fn)
{
    FILE *fp = fopen(fn, "rb");
    if (fp == NULL) { fprintf(stderr, "Unable to open .bwt file.\n"); exit(1); }
    bwt_t_gpu *bwt = (bwt_t_gpu *)calloc(1, sizeof(bwt_t_gpu));
    bwt->L2 = (bwtint_t_gpu *)calloc(5, sizeof(bwtint_t_gpu));
    if (fread(&bwt->primary, 8, 1, fp) != 1 || fread(bwt->L2 + 1, 8, 4, fp) != 4) { fprintf(stderr, "Unable to read .bwt file.\n"); exit(1); }
    fclose(fp);
    bwt->seq_len = bwt->L2[4];
    g_bwt_path = std::string(fn) + "128";                   // stock layout beside the GPU layout
    return bwt;
}

extern "C" void bwt_restore_sa_gpu(const char *fn, bwt_t_gpu *bwt)
{
    FILE *fp = fopen(fn, "rb");
    if (fp == NULL) { fprintf(stderr, "Unable to open .sa file.\n"); exit(1); }
    uint64_t hdr[7];
    if (fread(hdr, 8, 7, fp) != 7) { fprintf(stderr, "Unable to read .sa file.\n"); exit(1); }
    fclose(fp);
    if (hdr[0] != bwt->primary) { fprintf(stderr, "SA-BWT inconsistency: primary is not the same.\n"); exit(EXIT_FAILURE); }
    if (hdr[6] != bwt->seq_len) { fprintf(stderr, "SA-BWT inconsistency: seq_len is not the same.\n"); exit(EXIT_FAILURE); }
    bwt->sa_intv = (int)hdr[5];
    g_sa_path = fn;
}

extern "C" void bwt_destroy_gpu(bwt_t_gpu *bwt)
{
    if (!bwt) return;
    free(bwt->L2); free(bwt);
}

extern "C" bwt_t_gpu gpu_cpy_wrapper(bwt_t_gpu *bwt)
{
    ref_open();
    g_index = g_ref.load(g_bwt_path.c_str(), g_sa_path.c_str());
    if (!g_index) { fprintf(stderr, "[cpu_compat] cannot load %s / %s\n", g_bwt_path.c_str(), g_sa_path.c_str()); exit(EXIT_FAILURE); }
    bwt_t_gpu g;
    memset(&g, 0, sizeof(g));
    g.primary = bwt->primary; g.seq_len = bwt->seq_len; g.sa_intv = bwt->sa_intv;
    bwt_destroy_gpu(bwt);
    return g;
}

extern "C" void pre_calc_seed_intervals_wrapper(uint2 *, int, bwt_t_gpu) {}

extern "C" void free_gpuseed_data(gpuseed_storage_vector *)
{
    if (g_index && g_ref.release) g_ref.release(g_index);
    g_index = nullptr;
}

extern "C" mem_seed_v_gpu *seed_gpu(gpuseed_storage_vector *d)
{
    if (!d->is_smem) { fprintf(stderr, "[cpu_compat] MEM mode is not provided\n"); exit(EXIT_FAILURE); }
    FILE *fp = fopen(d->read_file, "r");
    if (!fp) { fprintf(stderr, "[cpu_compat] cannot open %s\n", d->read_file); exit(EXIT_FAILURE); }
    fseek(fp, (long)d->file_bytes_skip, SEEK_SET);
    std::vector<uint64_t> rbeg;
    std::vector<int2> qq;
    std::vector<uint32_t> score, per_read;
    std::vector<uint8_t> q;
    std::vector<uint64_t> iv;
    char *line = nullptr;
    size_t cap = 0;
    uint64_t file_bytes = 0;
    ssize_t got;
    while ((got = getline(&line, &cap, fp)) >= 0) {
        file_bytes += (uint64_t)got;
        if (line[0] == '>') continue;
        size_t len = (size_t)got;
        while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) --len;
        q.resize(len);
        for (size_t i = 0; i < len; ++i) {                 // nst_nt4_table
            switch (line[i]) {
                case 'A': case 'a': q[i] = 0; break;
                case 'C': case 'c': q[i] = 1; break;
                case 'G': case 'g': q[i] = 2; break;
                case 'T': case 't': q[i] = 3; break;
                default: q[i] = 4;
            }
        }
        uint32_t n_here = 0;
        if ((int)len >= d->min_seed_size) {
            iv.resize(5 * (len + 1));
            int x = 0;
            while (x < (int)len) {                         // mem_collect_intv pass 1 (bwa_index/bwamem.c:121-131)
                if (q[x] < 4) {
                    int n = 0;
                    x = g_ref.smem1(g_index, (int)len, q.data(), x, 1, iv.data(), &n);
                    for (int i = 0; i < n; ++i) {
                        const uint64_t k = iv[5 * i], s = iv[5 * i + 2];
                        const int beg = (int)iv[5 * i + 3], end = (int)iv[5 * i + 4];
                        if (end - beg < d->min_seed_size) continue;
                        uint64_t step = 1, count = s;
                        if (t_max_occ > 0) { step = s > (uint64_t)t_max_occ ? s / t_max_occ : 1; count = (s + step - 1) / step; if (count > (uint64_t)t_max_occ) count = t_max_occ; }
                        for (uint64_t t = 0; t < count; ++t) {
                            rbeg.push_back(g_ref.sa(g_index, k + t * step));
                            int2 v; v.x = beg; v.y = end;
                            qq.push_back(v);
                            score.push_back(t == 0 ? (uint32_t)s : 0u);
                            ++n_here;
                        }
                    }
                } else ++x;
            }
        }
        per_read.push_back(n_here);
    }
    free(line);
    fclose(fp);
    mem_seed_v_gpu *res = (mem_seed_v_gpu *)calloc(1, sizeof(mem_seed_v_gpu));
    const size_t ns = rbeg.size(), nr = per_read.size();
    res->rbeg = (bwtint_t_gpu *)malloc((ns ? ns : 1) * 8);
    res->qbeg = (int2 *)malloc((ns ? ns : 1) * sizeof(int2));
    res->score = (uint32_t *)malloc((ns ? ns : 1) * 4);
    memcpy(res->rbeg, rbeg.data(), ns * 8); memcpy(res->qbeg, qq.data(), ns * sizeof(int2)); memcpy(res->score, score.data(), ns * 4);
    res->n_ref_pos_fow_rev_results = (uint32_t *)malloc((nr ? nr : 1) * 4);
    res->n_ref_pos_fow_rev_prefix_sums = (uint32_t *)malloc((nr ? nr : 1) * 4);
    uint32_t run = 0;
    for (size_t r = 0; r < nr; ++r) { res->n_ref_pos_fow_rev_results[r] = per_read[r]; res->n_ref_pos_fow_rev_prefix_sums[r] = run; run += per_read[r]; }
    res->file_bytes_skip = d->file_bytes_skip + file_bytes;
    return res;
}

// ============================================================================== GASAL2 boundary
extern "C" void gasal_b200_set_ext_params(int w, int zdrop, int end_bonus, int o_ins, int e_ins, int pen_clip, int use_band)
{
    params_init();
    g_params.w = w; g_params.zdrop = zdrop; g_params.end_bonus = end_bonus; g_params.o_ins = o_ins; g_params.e_ins = e_ins;
    g_params.pen_clip = pen_clip; g_params.use_band = use_band;
}
extern "C" void gasal_b200_get_ext_params(bwa_b200_ext_params_t *out) { params_init(); *out = g_params; }

void gasal_copy_subst_scores(gasal_subst_scores *subst)
{
    params_init();
    fill_scmat(subst->match, subst->mismatch, g_params.mat);
    bool ins_follow = g_params.o_ins == g_params.o_del && g_params.e_ins == g_params.e_del;
    g_params.o_del = subst->gap_open; g_params.e_del = subst->gap_extend;
    if (ins_follow) { g_params.o_ins = subst->gap_open; g_params.e_ins = subst->gap_extend; }
}

Parameters::Parameters(int argc_, char **argv_)
{
    sa = 1; sb = 4; gapo = 6; gape = 1;
    start_pos = WITHOUT_START; print_out = 0; n_threads = 1; k_band = 0;
    secondBest = FALSE; isPacked = false; isReverseComplement = false;
    semiglobal_skipping_head = TARGET; semiglobal_skipping_tail = TARGET;
    algo = UNKNOWN;
    argc = argc_; argv = argv_;
}
Parameters::~Parameters() {}
void Parameters::print() {}

gasal_res_t *gasal_res_new_host(uint32_t max_n_alns, Parameters *)
{
    gasal_res_t *res = (gasal_res_t *)calloc(1, sizeof(gasal_res_t));
    res->aln_score = host<int32_t>(max_n_alns);
    res->query_batch_end = host<int32_t>(max_n_alns);
    res->target_batch_end = host<int32_t>(max_n_alns);
    return res;
}
void gasal_res_destroy_host(gasal_res_t *res)
{
    if (!res) return;
    free(res->aln_score); free(res->query_batch_end); free(res->target_batch_end); free(res);
}

host_batch_t *gasal_host_batch_new(uint32_t batch_bytes, uint32_t offset)
{
    host_batch_t *res = (host_batch_t *)calloc(1, sizeof(host_batch_t));
    res->data = host<uint8_t>(batch_bytes);
    res->page_size = batch_bytes;
    res->offset = offset;
    return res;
}
void gasal_host_batch_destroy(host_batch_t *res)
{
    while (res) { host_batch_t *next = res->next; free(res->data); free(res); res = next; }
}
host_batch_t *gasal_host_batch_getlast(host_batch_t *arg) { while (arg->next) arg = arg->next; return arg; }
void gasal_host_batch_reset(gasal_gpu_storage_t *s)
{
    host_batch_t *heads[2] = {s->extensible_host_unpacked_query_batch, s->extensible_host_unpacked_target_batch};
    for (host_batch_t *p : heads)
        for (; p; p = p->next) { p->data_size = 0; p->offset = 0; p->is_locked = 0; }
}

static uint32_t append(gasal_gpu_storage_t *s, uint32_t idx, const char *data, uint32_t size, data_source SRC, bool padded)
{ // GASAL2/src/host_batch.cpp:79-236: pages chain, a page that cannot take the sequence locks and the next one continues
    host_batch_t *page = SRC == QUERY ? s->extensible_host_unpacked_query_batch : s->extensible_host_unpacked_target_batch;
    uint32_t *total = SRC == QUERY ? &s->host_max_query_batch_bytes : &s->host_max_target_batch_bytes;
    const uint32_t pad = padded ? (8 - size % 8) % 8 : 0, need = size + pad;
    while (page->is_locked) page = page->next;
    if (page->page_size - page->data_size < need) {
        if (page->next == NULL) {
            uint32_t grow = page->page_size * 2;
            while (grow < need) grow *= 2;
            page->next = gasal_host_batch_new(grow, page->offset + page->data_size);
            *total += grow;
        } else page->next->offset = page->offset + page->data_size;
        page->is_locked = 1;
        page = page->next;
    }
    memcpy(page->data + (idx - page->offset), data, size);
    memset(page->data + (idx - page->offset) + size, N_CODE, pad);
    page->data_size += need;
    return idx + need;
}
uint32_t gasal_host_batch_fill(gasal_gpu_storage_t *s, uint32_t idx, const char *data, uint32_t size, data_source SRC) { return append(s, idx, data, size, SRC, true); }
uint32_t gasal_host_batch_add(gasal_gpu_storage_t *s, uint32_t idx, const char *data, uint32_t size, data_source SRC) { return append(s, idx, data, size, SRC, false); }
uint32_t gasal_host_batch_addbase(gasal_gpu_storage_t *s, uint32_t idx, const char base, data_source SRC) { return append(s, idx, &base, 1, SRC, false); }
void gasal_host_batch_print(host_batch_t *) {}
void gasal_host_batch_printall(host_batch_t *) {}

gasal_gpu_storage_v gasal_init_gpu_storage_v(int n_streams)
{
    gasal_gpu_storage_v v;
    v.n = n_streams;
    v.a = (gasal_gpu_storage_t *)calloc(n_streams, sizeof(gasal_gpu_storage_t));
    return v;
}

void gasal_init_streams(gasal_gpu_storage_v *vec, int host_max_query_batch_bytes, int gpu_max_query_batch_bytes,
                        int host_max_target_batch_bytes, int gpu_max_target_batch_bytes, int host_max_n_alns, int gpu_max_n_alns, Parameters *params)
{
    for (int i = 0; i < vec->n; ++i) {
        gasal_gpu_storage_t *s = &vec->a[i];
        s->extensible_host_unpacked_query_batch = gasal_host_batch_new(host_max_query_batch_bytes, 0);
        s->extensible_host_unpacked_target_batch = gasal_host_batch_new(host_max_target_batch_bytes, 0);
        s->host_query_batch_offsets = host<uint32_t>(host_max_n_alns);
        s->host_target_batch_offsets = host<uint32_t>(host_max_n_alns);
        s->host_query_batch_lens = host<uint32_t>(host_max_n_alns);
        s->host_target_batch_lens = host<uint32_t>(host_max_n_alns);
        s->host_seed_scores = host<uint32_t>(host_max_n_alns);
        s->host_res = gasal_res_new_host(host_max_n_alns, params);
        s->host_max_query_batch_bytes = host_max_query_batch_bytes;
        s->host_max_target_batch_bytes = host_max_target_batch_bytes;
        s->gpu_max_query_batch_bytes = gpu_max_query_batch_bytes;
        s->gpu_max_target_batch_bytes = gpu_max_target_batch_bytes;
        s->host_max_n_alns = host_max_n_alns;
        s->gpu_max_n_alns = gpu_max_n_alns;
        s->is_free = 1;
        s->id = i;
    }
}
void gasal_gpu_mem_alloc(gasal_gpu_storage_t *, int, int, int, Parameters *) {}
void gasal_gpu_mem_free(gasal_gpu_storage_t *, Parameters *) {}
void gasal_destroy_streams(gasal_gpu_storage_v *vec, Parameters *)
{
    for (int i = 0; i < vec->n; ++i) {
        gasal_gpu_storage_t *s = &vec->a[i];
        gasal_host_batch_destroy(s->extensible_host_unpacked_query_batch);
        gasal_host_batch_destroy(s->extensible_host_unpacked_target_batch);
        free(s->host_query_batch_offsets); free(s->host_target_batch_offsets); free(s->host_query_batch_lens); free(s->host_target_batch_lens);
        free(s->host_seed_scores);
        gasal_res_destroy_host(s->host_res);
    }
}
void gasal_destroy_gpu_storage_v(gasal_gpu_storage_v *vec) { free(vec->a); vec->a = NULL; vec->n = 0; }

void gasal_host_alns_resize(gasal_gpu_storage_t *s, int new_max_alns, Parameters *params)
{
    auto grow = [&](uint32_t *&arr) {
        uint32_t *n = host<uint32_t>(new_max_alns);
        memcpy(n, arr, s->host_max_n_alns * sizeof(uint32_t));
        free(arr);
        arr = n;
    };
    grow(s->host_query_batch_offsets); grow(s->host_target_batch_offsets);
    grow(s->host_query_batch_lens); grow(s->host_target_batch_lens); grow(s->host_seed_scores);
    gasal_res_destroy_host(s->host_res);
    s->host_res = gasal_res_new_host(new_max_alns, params);
    s->host_max_n_alns = new_max_alns;
}
void gasal_set_device(int, bool) {}

// a sequence of `len` (+ padding) bytes starting at batch offset `off`, gathered across pages
static void fetch(host_batch_t *page, uint32_t off, uint32_t len, std::vector<uint8_t> &out)
{
    out.resize(len);
    uint32_t done = 0;
    while (done < len) {
        while (page->next && off + done >= page->next->offset && page->next->data_size) page = page->next;
        const uint32_t in_page = off + done - page->offset;
        uint32_t take = page->data_size - in_page;
        if (take > len - done) take = len - done;
        memcpy(out.data() + done, page->data + in_page, take);
        done += take;
        if (done < len) {
            if (!page->next) { fprintf(stderr, "[cpu_compat] sequence runs past the last page\n"); exit(EXIT_FAILURE); }
            page = page->next;
        }
    }
}

void gasal_aln_async(gasal_gpu_storage_t *s, const uint32_t actual_query_batch_bytes, const uint32_t actual_target_batch_bytes,
                     const uint32_t actual_n_alns, Parameters *params)
{
    if (actual_n_alns <= 0 || actual_query_batch_bytes <= 0 || actual_target_batch_bytes <= 0 || actual_query_batch_bytes % 8 || actual_target_batch_bytes % 8 ||
        actual_n_alns > s->host_max_n_alns || params->algo != KSW) { fprintf(stderr, "[cpu_compat] gasal_aln_async: bad batch\n"); exit(EXIT_FAILURE); }
    params_init();
    const bwa_b200_ext_params_t &p = g_params;
    std::vector<uint8_t> q, t;
    for (uint32_t a = 0; a < actual_n_alns; ++a) {          // decoy_cpu_align, src/bwamem.c:1795-1903
        const uint32_t ql = s->host_query_batch_lens[a], tl = s->host_target_batch_lens[a];
        fetch(s->extensible_host_unpacked_query_batch, s->host_query_batch_offsets[a], ql, q);
        fetch(s->extensible_host_unpacked_target_batch, s->host_target_batch_offsets[a], tl, t);
        int qle, tle, gtle, gscore, max_off;
        const int score = ksw_extend2((int)ql, q.data(), (int)tl, t.data(), 5, p.mat, p.o_del, p.e_del, p.o_ins, p.e_ins, p.w, p.end_bonus, p.zdrop,
                                      (int)s->host_seed_scores[a], &qle, &tle, &gtle, &gscore, &max_off, p.use_band);
        if (gscore <= 0 || gscore <= score - p.pen_clip) {
            s->host_res->aln_score[a] = score; s->host_res->query_batch_end[a] = qle; s->host_res->target_batch_end[a] = tle;
        } else {
            s->host_res->aln_score[a] = gscore; s->host_res->query_batch_end[a] = (int)ql; s->host_res->target_batch_end[a] = gtle;
        }
    }
    s->is_free = 0;
}

int gasal_is_aln_async_done(gasal_gpu_storage_t *s)
{
    if (s->is_free == 1) return -2;
    gasal_host_batch_reset(s);
    s->is_free = 1;
    s->current_n_alns = 0;
    return 0;
}
