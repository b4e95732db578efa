// A miniature of the reference driver's use of the two library boundaries, written against the
// reference's own names (src/fastmap.c:417-511, src/bwamem.c:1102-1167,2106-2181), compiled with
// g++ against include/compat and linked to libbwamem_b200.so.  tests/test_compat_driver.py runs it
// on the GPU box and checks every number against the CPU oracle.
//
//   driver_like <index prefix> <reads.fa> <jobs.bin> <out.bin> <min_seed_len>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "seed_gen.h"
#include "gasal.h"
#include "args_parser.h"
#include "host_batch.h"
#include "ctors.h"
#include "interfaces.h"
#include "res.h"
#include "gasal_align.h"

struct gpu_batch { gasal_gpu_storage_t *gpu_storage; uint32_t n_query_batch, n_target_batch, n_seqs; };

// the fill_extension pattern of src/bwamem.c:1102-1167
static void fill_extension(gpu_batch *cur, const uint8_t *ref_seq, const uint8_t *read_seq, int ref_len, int read_len, int seed_score)
{
    cur->gpu_storage->current_n_alns++;
    if (cur->gpu_storage->current_n_alns > cur->gpu_storage->host_max_n_alns) {
        Parameters *args = new Parameters(0, NULL);
        args->algo = KSW; args->start_pos = WITHOUT_START;
        gasal_host_alns_resize(cur->gpu_storage, cur->gpu_storage->host_max_n_alns * 2, args);
        delete args;
    }
    cur->gpu_storage->host_target_batch_offsets[cur->n_seqs] = cur->n_target_batch;
    cur->gpu_storage->host_query_batch_offsets[cur->n_seqs] = cur->n_query_batch;
    cur->n_target_batch = gasal_host_batch_fill(cur->gpu_storage, cur->n_target_batch, (const char *)ref_seq, ref_len, TARGET);
    cur->n_query_batch = gasal_host_batch_fill(cur->gpu_storage, cur->n_query_batch, (const char *)read_seq, read_len, QUERY);
    cur->gpu_storage->host_query_batch_lens[cur->n_seqs] = read_len;
    cur->gpu_storage->host_target_batch_lens[cur->n_seqs] = ref_len;
    cur->gpu_storage->host_seed_scores[cur->n_seqs] = seed_score;
    cur->n_seqs++;
}

int main(int argc, char **argv)
{
    if (argc < 6) { fprintf(stderr, "usage\n"); return 2; }
    std::string prefix = argv[1];
    FILE *out = fopen(argv[4], "wb");

    // ---- seeding boundary (src/fastmap.c:432-465)
    gpuseed_storage_vector *gd = (gpuseed_storage_vector *)calloc(1, sizeof(gpuseed_storage_vector));
    gd->query_file = argv[1]; gd->read_file = argv[2]; gd->file_bytes_skip = 0;
    gd->min_seed_size = atoi(argv[5]); gd->is_smem = 1;
    gd->bwt = bwt_restore_bwt_gpu((prefix + ".bwt").c_str());
    bwt_restore_sa_gpu((prefix + ".sa").c_str(), gd->bwt);
    gd->bwt_gpu = gpu_cpy_wrapper(gd->bwt);
    gd->pre_calc_seed_len = 13; gd->pre_calc_seed_intervals_flag = 0;
    if (argc > 6 && atoi(argv[6])) gpuseed_b200_set_reseed(1, 1.5f, 10, 20);      // optional: the seed set of stock bwa mem
    mem_seed_v_gpu *seeds = seed_gpu(gd);
    free_gpuseed_data(gd);
    free(gd);
    // count reads = lines not starting with '>'
    uint64_t n_reads = 0;
    { FILE *f = fopen(argv[2], "r"); char *l = NULL; size_t c = 0; while (getline(&l, &c, f) >= 0) if (l[0] != '>') ++n_reads; free(l); fclose(f); }
    uint64_t n_seeds = n_reads ? seeds->n_ref_pos_fow_rev_prefix_sums[n_reads - 1] + seeds->n_ref_pos_fow_rev_results[n_reads - 1] : 0;
    fwrite(&n_reads, 8, 1, out); fwrite(&n_seeds, 8, 1, out);
    fwrite(seeds->n_ref_pos_fow_rev_results, 4, n_reads, out);
    fwrite(seeds->n_ref_pos_fow_rev_prefix_sums, 4, n_reads, out);
    fwrite(seeds->rbeg, 8, n_seeds, out);
    fwrite(seeds->qbeg, 8, n_seeds, out);
    // score is defined on the first seed of every SMEM group only (seed_gen.cu:540): walk like mem_chain does
    std::vector<uint32_t> sc(n_seeds, 0);
    for (uint64_t r = 0; r < n_reads; ++r) {
        uint32_t o = seeds->n_ref_pos_fow_rev_prefix_sums[r];
        for (uint32_t i = 0; i < seeds->n_ref_pos_fow_rev_results[r]; i += seeds->score[i + o]) sc[i + o] = seeds->score[i + o];
    }
    fwrite(sc.data(), 4, n_seeds, out);
    free(seeds->rbeg); free(seeds->qbeg); free(seeds->score);
    free(seeds->n_ref_pos_fow_rev_results); free(seeds->n_ref_pos_fow_rev_prefix_sums); free(seeds);

    // ---- extension boundary (src/fastmap.c:417-430,473-511; src/bwamem.c:2106-2181)
    FILE *jf = fopen(argv[3], "rb");
    uint32_t n_jobs = 0;
    if (fread(&n_jobs, 4, 1, jf) != 1) return 3;
    std::vector<uint32_t> ql(n_jobs), tl(n_jobs), h0(n_jobs);
    if (fread(ql.data(), 4, n_jobs, jf) != n_jobs || fread(tl.data(), 4, n_jobs, jf) != n_jobs || fread(h0.data(), 4, n_jobs, jf) != n_jobs) return 3;
    gasal_subst_scores sub; sub.match = 1; sub.mismatch = 4; sub.gap_open = 6; sub.gap_extend = 1;
    gasal_copy_subst_scores(&sub);
    Parameters *args = new Parameters(0, NULL);
    args->algo = KSW; args->start_pos = WITHOUT_START;
    gasal_gpu_storage_v vec = gasal_init_gpu_storage_v(2);
    // deliberately small so that pages chain and the per-alignment arrays are resized
    gasal_init_streams(&vec, 4096, 4096, 8192, 8192, 64, 64, args);
    std::vector<int32_t> score(n_jobs), qend(n_jobs), tend(n_jobs);
    uint32_t done = 0;
    int which = 0;
    std::vector<uint8_t> q, t;
    while (done < n_jobs) {
        gasal_gpu_storage_t *st = &vec.a[which];
        which ^= 1;
        if (gasal_is_aln_async_done(st) != -2) { fprintf(stderr, "storage not free\n"); return 4; }
        gpu_batch cur = {st, 0, 0, 0};
        uint32_t take = n_jobs - done < 700 ? n_jobs - done : 700, first = done;
        for (uint32_t a = 0; a < take; ++a, ++done) {
            q.resize(ql[done]); t.resize(tl[done]);
            if ((ql[done] && fread(q.data(), 1, ql[done], jf) != ql[done]) || (tl[done] && fread(t.data(), 1, tl[done], jf) != tl[done])) return 3;
            fill_extension(&cur, t.data(), q.data(), (int)tl[done], (int)ql[done], (int)h0[done]);
        }
        gasal_aln_async(st, cur.n_query_batch, cur.n_target_batch, cur.n_seqs, args);
        int rc;
        while ((rc = gasal_is_aln_async_done(st)) == -1) ;
        if (rc != 0) { fprintf(stderr, "unexpected poll result %d\n", rc); return 5; }
        for (uint32_t a = 0; a < take; ++a) {
            score[first + a] = st->host_res->aln_score[a];
            qend[first + a] = st->host_res->query_batch_end[a];
            tend[first + a] = st->host_res->target_batch_end[a];
        }
        if (st->is_free != 1 || st->current_n_alns != 0) { fprintf(stderr, "storage state not reset\n"); return 6; }
    }
    fwrite(&n_jobs, 4, 1, out);
    fwrite(score.data(), 4, n_jobs, out); fwrite(qend.data(), 4, n_jobs, out); fwrite(tend.data(), 4, n_jobs, out);
    gasal_destroy_streams(&vec, args);
    gasal_destroy_gpu_storage_v(&vec);
    delete args;
    fclose(jf); fclose(out);
    return 0;
}
