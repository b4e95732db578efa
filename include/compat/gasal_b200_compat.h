/*
 * include/compat/gasal_b200_compat.h -- source-compatible replacement for the part of the GASAL2
 * API that the reference's `gase_aln` driver uses (src/bntseq.h:35-40 includes gasal.h,
 * args_parser.h, host_batch.h, ctors.h, interfaces.h, res.h, gasal_align.h; the one-line headers of
 * those names next to this file forward here).
 *
 * Same names, same C++ linkage, same argument meaning and error behaviour as
 *   GASAL2/src/gasal.h:72-146        types the driver touches field by field
 *   GASAL2/src/args_parser.h:24-68   Parameters (only the members the driver sets)
 *   GASAL2/src/ctors.h:5-15          gasal_init_gpu_storage_v / gasal_init_streams / gasal_destroy_*
 *   GASAL2/src/host_batch.h:13       gasal_host_batch_fill (+ new / destroy / reset / getlast)
 *   GASAL2/src/interfaces.h:9-14     gasal_host_alns_resize, gasal_set_device
 *   GASAL2/src/gasal_align.h:96-102  gasal_copy_subst_scores, gasal_aln_async, gasal_is_aln_async_done
 * implemented over the B200 extension path (bwa-mem_gpu_b200/csrc/extend.cu).  Only algo == KSW
 * with WITHOUT_START is provided: it is the only mode gase_aln selects (src/fastmap.c:427-430).
 *
 * Results follow ksw_extend2 on the CPU (src/ksw.c:864-986) followed by the local-vs-to-end rule
 * (src/bwamem.c:1892-1901), i.e. what the driver's own decoy_cpu_align() writes into host_res --
 * not the divergences of GASAL2's KSW kernel (no band, zdrop 0, fixed clip penalty).  The band,
 * z-drop, end bonus, insertion penalties and clip penalty, which the GASAL2 boundary cannot carry,
 * are set with gasal_b200_set_ext_params(); defaults are the fork's mem_opt_init values
 * (w = 300 unused because the fork passes opt_ext = 0, zdrop = 0, pen_clip = 5).
 */
#ifndef GASAL_B200_COMPAT_H
#define GASAL_B200_COMPAT_H

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include "bwamem_b200.h"

#ifndef N_CODE
#define N_CODE 4            /* padding base appended by gasal_host_batch_fill (GASAL2 Makefile N_CODE) */
#endif

enum comp_start { WITHOUT_START, WITH_START };
enum Bool { FALSE, TRUE };
enum data_source { NONE, QUERY, TARGET, BOTH };
enum algo_type { UNKNOWN, GLOBAL, SEMI_GLOBAL, LOCAL, MICROLOCAL, BANDED, KSW };
enum operation_on_seq { FORWARD_NATURAL, REVERSE_NATURAL, FORWARD_COMPLEMENT, REVERSE_COMPLEMENT };

/* pinned host page of unpacked bases; pages chain when a batch outgrows the first one */
struct host_batch {
    uint8_t *data;
    uint32_t page_size;
    uint32_t data_size;
    uint32_t offset;
    int is_locked;
    struct host_batch *next;
};
typedef struct host_batch host_batch_t;

struct gasal_res {
    int32_t *aln_score;
    int32_t *query_batch_end;
    int32_t *target_batch_end;
    int32_t *query_batch_start;     /* NULL for WITHOUT_START */
    int32_t *target_batch_start;
};
typedef struct gasal_res gasal_res_t;

/* per-stream state.  The members below are the ones the driver reads or writes directly
 * (src/bwamem.c:1105-1161,2005-2032,2207-2211); device-side members of the original struct have
 * no meaning here and are folded into the opaque `b200` handle. */
typedef struct {
    host_batch_t *extensible_host_unpacked_query_batch;
    host_batch_t *extensible_host_unpacked_target_batch;

    uint32_t *host_query_batch_offsets;
    uint32_t *host_target_batch_offsets;
    uint32_t *host_query_batch_lens;
    uint32_t *host_target_batch_lens;
    uint32_t *host_seed_scores;

    uint8_t *host_query_op;
    uint8_t *host_target_op;

    gasal_res_t *host_res;
    gasal_res_t *host_res_second;

    uint32_t gpu_max_query_batch_bytes;
    uint32_t gpu_max_target_batch_bytes;
    uint32_t host_max_query_batch_bytes;
    uint32_t host_max_target_batch_bytes;
    uint32_t gpu_max_n_alns;
    uint32_t host_max_n_alns;
    uint32_t current_n_alns;
    void *str;                  /* cudaStream_t of this storage */
    int is_free;
    int id;

    bwa_b200_extender_t *b200;  /* B200 extension workspace bound to this stream */
} gasal_gpu_storage_t;

typedef struct {
    int n;
    gasal_gpu_storage_t *a;
} gasal_gpu_storage_v;

typedef struct {
    int32_t match;
    int32_t mismatch;
    int32_t gap_open;
    int32_t gap_extend;
} gasal_subst_scores;

class Parameters {
public:
    Parameters(int argc, char **argv);
    ~Parameters();
    void print();

    int32_t sa, sb, gapo, gape;
    comp_start start_pos;
    int print_out;
    int n_threads;
    int32_t k_band;
    Bool secondBest;
    bool isPacked;
    bool isReverseComplement;
    data_source semiglobal_skipping_head;
    data_source semiglobal_skipping_tail;
    algo_type algo;
    std::string query_batch_fasta_filename;
    std::string target_batch_fasta_filename;

private:
    int argc;
    char **argv;
};

/* ctors.h */
gasal_gpu_storage_v gasal_init_gpu_storage_v(int n_streams);
void gasal_init_streams(gasal_gpu_storage_v *gpu_storage_vec, int host_max_query_batch_bytes, int gpu_max_query_batch_bytes,
                        int host_max_target_batch_bytes, int gpu_max_target_batch_bytes, int host_max_n_alns, int gpu_max_n_alns,
                        Parameters *params);
void gasal_gpu_mem_alloc(gasal_gpu_storage_t *gpu_storage, int gpu_max_query_batch_bytes, int gpu_max_target_batch_bytes,
                         int gpu_max_n_alns, Parameters *params);
void gasal_gpu_mem_free(gasal_gpu_storage_t *gpu_storage, Parameters *params);
void gasal_destroy_streams(gasal_gpu_storage_v *gpu_storage_vec, Parameters *params);
void gasal_destroy_gpu_storage_v(gasal_gpu_storage_v *gpu_storage_vec);

/* host_batch.h */
host_batch_t *gasal_host_batch_new(uint32_t batch_bytes, uint32_t offset);
void gasal_host_batch_destroy(host_batch_t *res);
host_batch_t *gasal_host_batch_getlast(host_batch_t *arg);
void gasal_host_batch_reset(gasal_gpu_storage_t *gpu_storage);
uint32_t gasal_host_batch_fill(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char *data, uint32_t size, data_source SRC);
uint32_t gasal_host_batch_add(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char *data, uint32_t size, data_source SRC);
uint32_t gasal_host_batch_addbase(gasal_gpu_storage_t *gpu_storage, uint32_t idx, const char base, data_source SRC);
void gasal_host_batch_print(host_batch_t *res);
void gasal_host_batch_printall(host_batch_t *res);

/* interfaces.h */
void gasal_host_alns_resize(gasal_gpu_storage_t *gpu_storage, int new_max_alns, Parameters *params);
void gasal_set_device(int gpu_select = 0, bool isPrintingProp = true);

/* res.h */
gasal_res_t *gasal_res_new_host(uint32_t max_n_alns, Parameters *params);
void gasal_res_destroy_host(gasal_res_t *res);

/* gasal_align.h */
void gasal_copy_subst_scores(gasal_subst_scores *subst);
void gasal_aln_async(gasal_gpu_storage_t *gpu_storage, const uint32_t actual_query_batch_bytes,
                     const uint32_t actual_target_batch_bytes, const uint32_t actual_n_alns, Parameters *params);
int gasal_is_aln_async_done(gasal_gpu_storage_t *gpu_storage);

/* B200 addition: everything ksw_extend2 takes that the GASAL2 boundary cannot express.  Process-wide,
 * like gasal_copy_subst_scores; the match/mismatch/gap scores given there are merged in. */
extern "C" void gasal_b200_set_ext_params(int w, int zdrop, int end_bonus, int o_ins, int e_ins, int pen_clip, int use_band);
extern "C" void gasal_b200_get_ext_params(bwa_b200_ext_params_t *out);

#endif
