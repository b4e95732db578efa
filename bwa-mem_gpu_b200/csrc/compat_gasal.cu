// compat_gasal.cu -- the GASAL2 entry points used by the reference's gase_aln driver, implemented
// over the B200 extension path.  Interface and error behaviour follow GASAL2/src/{ctors,host_batch,
// interfaces,res}.cpp and gasal_align.cu:29-311; the arithmetic follows ksw_extend2 on the CPU.
#include "common.h"
#include "gasal_b200_compat.h"
#include <mutex>
#include <vector>

namespace {

std::mutex g_mu;
bwa_b200_ext_params_t g_params;
bool g_params_init = false;
int g_match = 1, g_mismatch = 4;
thread_local int t_device = 0;

void params_init_locked()
{
    if (g_params_init) return;
    // the fork's defaults: mem_opt_init (src/bwamem.c:101-129) and opt_ext = 0 at the call site
    bwa_b200_ext_params_default(&g_params);
    g_params.w = 300; g_params.zdrop = 0; g_params.use_band = 0; g_params.pen_clip = 5; g_params.end_bonus = 5;
    g_params_init = true;
}

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "[GASAL-B200 ERROR:] %s: %s\n", what, bwa_b200_last_error());
    exit(EXIT_FAILURE);
}

template <class T> T *pinned(size_t n)
{
    void *p = bwa_b200_host_alloc(n * sizeof(T));
    if (!p) die("pinned host allocation");
    return (T *)p;
}

} // namespace

extern "C" void gasal_b200_set_ext_params(int w, int zdrop, int end_bonus, int o_ins, int e_ins, int pen_clip, int use_band)
{
    std::lock_guard<std::mutex> lk(g_mu);
    params_init_locked();
    g_params.w = w; g_params.zdrop = zdrop; g_params.end_bonus = end_bonus; g_params.o_ins = o_ins; g_params.e_ins = e_ins;
    g_params.pen_clip = pen_clip; g_params.use_band = use_band;
}

extern "C" void gasal_b200_get_ext_params(bwa_b200_ext_params_t *out)
{
    std::lock_guard<std::mutex> lk(g_mu);
    params_init_locked();
    *out = g_params;
}

// gasal_align.cu:314-331: match, mismatch, gap open, gap extend.  The reference uploads them to
// __constant__ memory of the current device; here they are process-wide and travel with each launch.
void gasal_copy_subst_scores(gasal_subst_scores *subst)
{
    std::lock_guard<std::mutex> lk(g_mu);
    params_init_locked();
    g_match = subst->match; g_mismatch = subst->mismatch;
    bwa_b200_fill_scmat(subst->match, subst->mismatch, g_params.mat);
    bool ins_follow = g_params.o_ins == g_params.o_del && g_params.e_ins == g_params.e_del;
    g_params.o_del = subst->gap_open; g_params.e_del = subst->gap_extend;
    if (ins_follow) { g_params.o_ins = subst->gap_open; g_params.e_ins = subst->gap_extend; }   // GASAL2 has one gap model
}

// ------------------------------------------------------------------------------ Parameters
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
void Parameters::print()
{
    fprintf(stderr, "sa=%d sb=%d gapo=%d gape=%d start_pos=%d algo=%d k_band=%d\n", sa, sb, gapo, gape, (int)start_pos, (int)algo, k_band);
}

// --------------------------------------------------------------------------------- res.cpp
gasal_res_t *gasal_res_new_host(uint32_t max_n_alns, Parameters *params)
{
    gasal_res_t *res = (gasal_res_t *)calloc(1, sizeof(gasal_res_t));
    if (!res || !params) { fprintf(stderr, "[GASAL ERROR:] gasal_res_new_host: bad argument\n"); exit(EXIT_FAILURE); }
    res->aln_score = pinned<int32_t>(max_n_alns);
    if (params->algo == GLOBAL) return res;
    res->query_batch_end = pinned<int32_t>(max_n_alns);
    res->target_batch_end = pinned<int32_t>(max_n_alns);
    if (params->start_pos == WITH_START) {
        res->query_batch_start = pinned<int32_t>(max_n_alns);
        res->target_batch_start = pinned<int32_t>(max_n_alns);
    }
    return res;
}

void gasal_res_destroy_host(gasal_res_t *res)
{
    if (!res) return;
    bwa_b200_host_free(res->aln_score); bwa_b200_host_free(res->query_batch_end); bwa_b200_host_free(res->target_batch_end);
    bwa_b200_host_free(res->query_batch_start); bwa_b200_host_free(res->target_batch_start);
    free(res);
}

// -------------------------------------------------------------------------- host_batch.cpp
host_batch_t *gasal_host_batch_new(uint32_t batch_bytes, uint32_t offset)
{
    host_batch_t *res = (host_batch_t *)calloc(1, sizeof(host_batch_t));
    res->data = pinned<uint8_t>(batch_bytes);
    res->page_size = batch_bytes;
    res->offset = offset;
    return res;
}

void gasal_host_batch_destroy(host_batch_t *res)
{
    if (res == NULL) { fprintf(stderr, "[GASAL ERROR] Trying to free a NULL pointer\n"); exit(1); }
    while (res) {
        host_batch_t *next = res->next;
        bwa_b200_host_free(res->data);
        free(res);
        res = next;
    }
}

host_batch_t *gasal_host_batch_getlast(host_batch_t *arg)
{
    while (arg->next) arg = arg->next;
    return arg;
}

void gasal_host_batch_reset(gasal_gpu_storage_t *gpu_storage)
{
    host_batch_t *heads[2] = {gpu_storage->extensible_host_unpacked_query_batch, gpu_storage->extensible_host_unpacked_target_batch};
    for (host_batch_t *p : heads)
        for (; p; p = p->next) { p->data_size = 0; p->offset = 0; p->is_locked = 0; }
}

// append `size` bases at running offset idx, pad to a multiple of 8 with N_CODE, return the new
// running offset (GASAL2/src/host_batch.cpp:79-153); a full page locks and the chain grows by a
// page of twice the size
uint32_t gasal_host_batch_fill(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char *data, uint32_t size, data_source SRC)
{
    host_batch_t *page = NULL;
    uint32_t *total = NULL;
    if (SRC == QUERY) { page = gpu_storage->extensible_host_unpacked_query_batch; total = &gpu_storage->host_max_query_batch_bytes; }
    else if (SRC == TARGET) { page = gpu_storage->extensible_host_unpacked_target_batch; total = &gpu_storage->host_max_target_batch_bytes; }
    else { fprintf(stderr, "[GASAL ERROR:] gasal_host_batch_fill: SRC must be QUERY or TARGET\n"); exit(EXIT_FAILURE); }
    const uint32_t pad = (8 - size % 8) % 8, need = size + pad;
    while (page->is_locked) page = page->next;
    if (page->page_size - page->data_size < need) {
        if (page->next == NULL) {
            uint32_t grow = page->page_size * 2;
            while (grow < need) grow *= 2;
            fprintf(stderr, "[GASAL WARNING:] Trying to write %d bytes while only %d remain (%s) (block size %d, filled %d bytes).\n"
                            "                 Allocating a new block of size %d, total size available reaches %d. Doing this repeadtedly slows down the execution.\n",
                    need, page->page_size - page->data_size, SRC == QUERY ? "query" : "target", page->page_size, page->data_size, grow, *total + grow);
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

// unpadded append of `size` bases at running offset idx (GASAL2/src/host_batch.cpp:156-236); same page rules as fill
uint32_t gasal_host_batch_add(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char *data, uint32_t size, data_source SRC)
{
    host_batch_t *page = NULL;
    uint32_t *total = NULL;
    if (SRC == QUERY) { page = gpu_storage->extensible_host_unpacked_query_batch; total = &gpu_storage->host_max_query_batch_bytes; }
    else if (SRC == TARGET) { page = gpu_storage->extensible_host_unpacked_target_batch; total = &gpu_storage->host_max_target_batch_bytes; }
    else { fprintf(stderr, "[GASAL ERROR:] gasal_host_batch_add: SRC must be QUERY or TARGET\n"); exit(EXIT_FAILURE); }
    while (page->is_locked) page = page->next;
    if (page->page_size - page->data_size < size) {
        if (page->next == NULL) {
            uint32_t grow = page->page_size * 2;
            while (grow < size) grow *= 2;
            page->next = gasal_host_batch_new(grow, page->offset + page->data_size);
            *total += grow;
        } else page->next->offset = page->offset + page->data_size;
        page->is_locked = 1;
        page = page->next;
    }
    memcpy(page->data + (idx - page->offset), data, size);
    page->data_size += size;
    return idx + size;
}

uint32_t gasal_host_batch_addbase(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char base, data_source SRC)
{
    return gasal_host_batch_add(gpu_storage, idx, &base, 1, SRC);
}

void gasal_host_batch_print(host_batch_t *res)
{
    fprintf(stderr, "[GASAL PRINT] Page data: offset=%d, next_offset=%d, data size=%d, page size=%d\n", res->offset,
            res->next ? (int)res->next->offset : -1, res->data_size, res->page_size);
}

void gasal_host_batch_printall(host_batch_t *res)
{
    for (; res; res = res->next) { gasal_host_batch_print(res); if (res->next) fprintf(stderr, "+--->"); }
}

// ------------------------------------------------------------------------------- ctors.cpp
gasal_gpu_storage_v gasal_init_gpu_storage_v(int n_streams)
{
    gasal_gpu_storage_v v;
    v.n = n_streams;
    v.a = (gasal_gpu_storage_t *)calloc(n_streams, sizeof(gasal_gpu_storage_t));
    return v;
}

void gasal_init_streams(gasal_gpu_storage_v *vec, int host_max_query_batch_bytes, int gpu_max_query_batch_bytes,
                        int host_max_target_batch_bytes, int gpu_max_target_batch_bytes, int host_max_n_alns, int gpu_max_n_alns,
                        Parameters *params)
{
    if (params->algo != KSW || params->start_pos != WITHOUT_START) {
        fprintf(stderr, "[GASAL-B200 ERROR:] only algo = KSW with start_pos = WITHOUT_START is provided (the mode gase_aln uses)\n");
        exit(EXIT_FAILURE);
    }
    for (int i = 0; i < vec->n; ++i) {
        gasal_gpu_storage_t *s = &vec->a[i];
        s->extensible_host_unpacked_query_batch = gasal_host_batch_new(host_max_query_batch_bytes, 0);
        s->extensible_host_unpacked_target_batch = gasal_host_batch_new(host_max_target_batch_bytes, 0);
        s->host_query_batch_offsets = pinned<uint32_t>(host_max_n_alns);
        s->host_target_batch_offsets = pinned<uint32_t>(host_max_n_alns);
        s->host_query_batch_lens = pinned<uint32_t>(host_max_n_alns);
        s->host_target_batch_lens = pinned<uint32_t>(host_max_n_alns);
        s->host_seed_scores = pinned<uint32_t>(host_max_n_alns);
        s->host_query_op = NULL; s->host_target_op = NULL;
        s->host_res = gasal_res_new_host(host_max_n_alns, params);
        s->host_res_second = NULL;
        s->host_max_query_batch_bytes = host_max_query_batch_bytes;
        s->host_max_target_batch_bytes = host_max_target_batch_bytes;
        s->gpu_max_query_batch_bytes = gpu_max_query_batch_bytes;
        s->gpu_max_target_batch_bytes = gpu_max_target_batch_bytes;
        s->host_max_n_alns = host_max_n_alns;
        s->gpu_max_n_alns = gpu_max_n_alns;
        s->current_n_alns = 0;
        if (bwa_b200_extender_create(t_device, gpu_max_n_alns, gpu_max_query_batch_bytes, gpu_max_target_batch_bytes, &s->b200)) die("gasal_init_streams");
        s->str = bwa_b200_extender_stream(s->b200);
        s->is_free = 1;
        s->id = i;
    }
}

void gasal_gpu_mem_alloc(gasal_gpu_storage_t *, int, int, int, Parameters *) { /* device buffers belong to the extender and grow on demand */ }
void gasal_gpu_mem_free(gasal_gpu_storage_t *, Parameters *) {}

void gasal_destroy_streams(gasal_gpu_storage_v *vec, Parameters *)
{
    for (int i = 0; i < vec->n; ++i) {
        gasal_gpu_storage_t *s = &vec->a[i];
        bwa_b200_extender_destroy(s->b200);
        s->b200 = NULL;
        gasal_host_batch_destroy(s->extensible_host_unpacked_query_batch);
        gasal_host_batch_destroy(s->extensible_host_unpacked_target_batch);
        bwa_b200_host_free(s->host_query_batch_offsets); bwa_b200_host_free(s->host_target_batch_offsets);
        bwa_b200_host_free(s->host_query_batch_lens); bwa_b200_host_free(s->host_target_batch_lens);
        bwa_b200_host_free(s->host_seed_scores);
        gasal_res_destroy_host(s->host_res);
    }
}

void gasal_destroy_gpu_storage_v(gasal_gpu_storage_v *vec) { free(vec->a); vec->a = NULL; vec->n = 0; }

// -------------------------------------------------------------------------- interfaces.cpp
void gasal_host_alns_resize(gasal_gpu_storage_t *s, int new_max_alns, Parameters *params)
{ // doubles the per-alignment pinned arrays, keeping their contents (interfaces.cpp:26-78)
    fprintf(stderr, "[GASAL RESIZER] Resizing host_max_n_alns from %d to %d\n", s->host_max_n_alns, new_max_alns);
    auto grow = [&](uint32_t *&arr) {
        uint32_t *n = pinned<uint32_t>(new_max_alns);
        memcpy(n, arr, s->host_max_n_alns * sizeof(uint32_t));
        bwa_b200_host_free(arr);
        arr = n;
    };
    grow(s->host_query_batch_offsets); grow(s->host_target_batch_offsets);
    grow(s->host_query_batch_lens); grow(s->host_target_batch_lens); grow(s->host_seed_scores);
    gasal_res_destroy_host(s->host_res);
    s->host_res = gasal_res_new_host(new_max_alns, params);
    s->host_max_n_alns = new_max_alns;
}

void gasal_set_device(int gpu_select, bool isPrintingProp)
{
    t_device = gpu_select;
    if (cudaSetDevice(gpu_select) != cudaSuccess) { fprintf(stderr, "[GASAL-B200 ERROR:] cannot select device %d\n", gpu_select); exit(EXIT_FAILURE); }
    if (isPrintingProp) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, gpu_select) == cudaSuccess) fprintf(stderr, "[GASAL-B200 INFO:] device %d: %s, %d SMs\n", gpu_select, p.name, p.multiProcessorCount);
    }
}

// ------------------------------------------------------------------------- gasal_align.cu
void gasal_aln_async(gasal_gpu_storage_t *s, const uint32_t actual_query_batch_bytes, const uint32_t actual_target_batch_bytes,
                     const uint32_t actual_n_alns, Parameters *params)
{
    // same preconditions, same fatal behaviour as gasal_align.cu:32-67
    if (actual_n_alns <= 0) { fprintf(stderr, "[GASAL ERROR:] actual_n_alns <= 0\n"); exit(EXIT_FAILURE); }
    if (actual_query_batch_bytes <= 0) { fprintf(stderr, "[GASAL ERROR:] actual_query_batch_bytes <= 0\n"); exit(EXIT_FAILURE); }
    if (actual_target_batch_bytes <= 0) { fprintf(stderr, "[GASAL ERROR:] actual_target_batch_bytes <= 0\n"); exit(EXIT_FAILURE); }
    if (actual_query_batch_bytes % 8) { fprintf(stderr, "[GASAL ERROR:] actual_query_batch_bytes=%d is not a multiple of 8\n", actual_query_batch_bytes); exit(EXIT_FAILURE); }
    if (actual_target_batch_bytes % 8) { fprintf(stderr, "[GASAL ERROR:] actual_target_batch_bytes=%d is not a multiple of 8\n", actual_target_batch_bytes); exit(EXIT_FAILURE); }
    if (actual_query_batch_bytes > s->host_max_query_batch_bytes) { fprintf(stderr, "[GASAL ERROR:] actual_query_batch_bytes(%d) > host_max_query_batch_bytes(%d)\n", actual_query_batch_bytes, s->host_max_query_batch_bytes); exit(EXIT_FAILURE); }
    if (actual_target_batch_bytes > s->host_max_target_batch_bytes) { fprintf(stderr, "[GASAL ERROR:] actual_target_batch_bytes(%d) > host_max_target_batch_bytes(%d)\n", actual_target_batch_bytes, s->host_max_target_batch_bytes); exit(EXIT_FAILURE); }
    if (actual_n_alns > s->host_max_n_alns) { fprintf(stderr, "[GASAL ERROR:] actual_n_alns(%d) > host_max_n_alns(%d)\n", actual_n_alns, s->host_max_n_alns); exit(EXIT_FAILURE); }
    if (params->algo != KSW) { fprintf(stderr, "[GASAL-B200 ERROR:] only algo = KSW is provided\n"); exit(EXIT_FAILURE); }

    std::vector<bwa_b200_host_page_t> qp, tp;
    for (host_batch_t *p = s->extensible_host_unpacked_query_batch; p; p = p->next)
        if (p->data_size) qp.push_back({p->data, p->offset, p->data_size});
    for (host_batch_t *p = s->extensible_host_unpacked_target_batch; p; p = p->next)
        if (p->data_size) tp.push_back({p->data, p->offset, p->data_size});
    bwa_b200_ext_params_t ep;
    gasal_b200_get_ext_params(&ep);
    if (bwa_b200_extend_async_paged(s->b200, &ep, actual_n_alns, qp.data(), (int)qp.size(), actual_query_batch_bytes,
                                    s->host_query_batch_offsets, s->host_query_batch_lens, tp.data(), (int)tp.size(),
                                    actual_target_batch_bytes, s->host_target_batch_offsets, s->host_target_batch_lens,
                                    s->host_seed_scores, NULL, s->host_res->aln_score, s->host_res->query_batch_end,
                                    s->host_res->target_batch_end))
        die("gasal_aln_async");
    s->is_free = 0;
}

// 0: finished (pages reset, storage free again); -1: still running; -2: nothing was launched
int gasal_is_aln_async_done(gasal_gpu_storage_t *s)
{
    if (s->is_free == 1) return -2;
    int st = bwa_b200_extend_query(s->b200);
    if (st == 1) return -1;
    if (st < 0) die("gasal_is_aln_async_done");
    if (bwa_b200_extend_wait(s->b200)) die("gasal_is_aln_async_done");
    gasal_host_batch_reset(s);
    s->is_free = 1;
    s->current_n_alns = 0;
    return 0;
}
