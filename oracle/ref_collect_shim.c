/*
 * oracle/ref_collect_shim.c -- TEST INFRASTRUCTURE ONLY.
 * mem_collect_intv (bwa_index/bwamem.c:114-162: SMEM pass 1, re-seeding passes 2 and 3, sort) is a static
 * function of the reference's bwamem.c, so this translation unit includes that UNMODIFIED file from the scratch
 * copy oracle/build_ref.sh compiles in (it takes the place of bwamem.o inside oracle/_ref/libbwaref.so) and adds
 * two exported entry points that call it.  No reference code is restated here.
 */
#include "bwamem.c"

/* out: 5 x u64 per interval = x0, x1, x2, start, end; returns the number of intervals (all of them are counted,
 * the first `cap` are stored) */
int ref_collect_intv(void *h, int len, const uint8_t *q, int min_seed_len, float split_factor, int split_width,
                     int max_mem_intv, uint64_t *out, int cap)
{
    mem_opt_t *opt = mem_opt_init();
    opt->min_seed_len = min_seed_len; opt->split_factor = split_factor; opt->split_width = split_width; opt->max_mem_intv = max_mem_intv;
    smem_aux_t *a = smem_aux_init();
    mem_collect_intv(opt, (const bwt_t *)h, len, q, a);
    int n = (int)a->mem.n;
    for (int i = 0; i < n && i < cap; ++i) {
        const bwtintv_t *p = &a->mem.a[i];
        out[5 * i + 0] = p->x[0]; out[5 * i + 1] = p->x[1]; out[5 * i + 2] = p->x[2];
        out[5 * i + 3] = p->info >> 32; out[5 * i + 4] = (uint32_t)p->info;
    }
    smem_aux_destroy(a);
    free(opt);
    return n;
}

/* the same over a batch of reads (codes 0..4, concatenated); n_smems[r] per read, intervals back to back.
 * Returns the total, or -1 when cap is exceeded. */
int64_t ref_collect_batch(void *h, const uint8_t *reads, const uint64_t *read_off, int64_t n_reads, int min_seed_len,
                          float split_factor, int split_width, int max_mem_intv, uint32_t *n_smems, uint64_t *out, int64_t cap)
{
    mem_opt_t *opt = mem_opt_init();
    opt->min_seed_len = min_seed_len; opt->split_factor = split_factor; opt->split_width = split_width; opt->max_mem_intv = max_mem_intv;
    smem_aux_t *a = smem_aux_init();
    int64_t tot = 0;
    for (int64_t r = 0; r < n_reads; ++r) {
        int len = (int)(read_off[r + 1] - read_off[r]);
        n_smems[r] = 0;
        if (len < min_seed_len) continue;                      /* mem_chain, bwa_index/bwamem.c:263 */
        mem_collect_intv(opt, (const bwt_t *)h, len, reads + read_off[r], a);
        if (tot + (int64_t)a->mem.n > cap) { tot = -1; break; }
        for (size_t i = 0; i < a->mem.n; ++i, ++tot) {
            const bwtintv_t *p = &a->mem.a[i];
            out[5 * tot + 0] = p->x[0]; out[5 * tot + 1] = p->x[1]; out[5 * tot + 2] = p->x[2];
            out[5 * tot + 3] = p->info >> 32; out[5 * tot + 4] = (uint32_t)p->info;
        }
        n_smems[r] = (uint32_t)a->mem.n;
    }
    smem_aux_destroy(a);
    free(opt);
    return tot;
}
