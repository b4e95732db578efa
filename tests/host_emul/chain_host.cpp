// Host build of the per-read chaining / job-construction source (bwa-mem_gpu_b200/csrc/chain_core.cuh).
// TEST INFRASTRUCTURE: lets the exact kernel source run against the oracle on the CPU box
// (tests/test_chain_host.py); never shipped and never used as a compute path.
//   g++ -O2 -shared -fPIC -I include -I bwa-mem_gpu_b200/csrc tests/host_emul/chain_host.cpp -o tests/host_emul/libchain_host.so
#include <vector>
#include <string.h>
#include "chain_core.cuh"

using namespace b200chain;

// one read: seeds -> chains (chains / cseeds need ns entries) -> regions (regs needs ns entries).
// pac / rd: the 2-bit reference and the read's 4-bit words, needed for reads mem_flt_chained_seeds acts on.
// triples: optional {aln_score, query_end, target_end} per job, SHORT batch then LONG batch of this read, applied with
// region_finish.  Returns the chain count or a negative error.
extern "C" int chain_host_read(const bwa_b200_chain_params_t *P, int n_ctg, const int64_t *ctg_off, const int32_t *ctg_len, const int32_t *ctg_alt,
                               int64_t l_pac, int l_query, uint32_t ns, const uint64_t *rbeg, const int32_t *qq, const uint32_t *score, int layout_all,
                               bwa_b200_chain_t *chains, bwa_b200_chain_seed_t *cseeds, bwa_b200_region_t *regs, int32_t counts[3],
                               const int32_t *short3, const int32_t *long3, const uint32_t *pac, const uint32_t *rd)
{
    Contigs ctg{ctg_off, ctg_len, ctg_alt, n_ctg, l_pac};
    const size_t n = ns ? ns : 1;
    std::vector<ChainW> ch(n);
    std::vector<int32_t> nxt(n), sq(2 * n), ord(n), kidx(n);
    std::vector<KbNode> nodes((size_t)nodes_needed(ns));
    ReadIO io{rbeg, qq, score, ns, l_query, layout_all, ch.data(), nxt.data(), sq.data(), ord.data(), kidx.data(), nodes.data(), (int32_t)nodes.size(), chains, cseeds};
    const int nc = chain_read(*P, ctg, io);
    counts[0] = counts[1] = counts[2] = 0;
    if (nc < 0) return nc;
    if (nc > 0 && flt_seeds_applies(*P, l_query)) {      // what seedsw_kernel + chain_long_kernel do (pac / rd: the device's packed forms)
        if (!pac || !rd) return -2;
        const int n_cs = chains[nc - 1].seed_off + chains[nc - 1].n;
        int16_t H[SEEDSW_MAX], E[SEEDSW_MAX];
        uint8_t qs[SEEDSW_MAX];
        for (int i = 0; i < n_cs; ++i) cseeds[i].score = seed_sw(*P, ctg, pac, rd, l_query, cseeds[i], H, E, qs, 1);
        flt_seeds_apply(*P, l_query, nc, chains, cseeds);
    }
    std::vector<uint64_t> srt(n);
    AlnIO ao{l_query, nc, chains, cseeds, srt.data(), regs};
    int n_short = 0, n_long = 0;
    const int nr = chain2aln_read(*P, ctg, ao, &n_short, &n_long);
    counts[0] = nr; counts[1] = n_short; counts[2] = n_long;
    if (short3 && long3)
        for (int i = 0; i < nr; ++i)
            region_finish(regs[i], l_query, regs[i].job_long >= 0 ? long3 + 3 * regs[i].job_long : long3, regs[i].job_short >= 0 ? short3 + 3 * regs[i].job_short : short3);
    return nc;
}

// job sequences of one read's regions for one batch (0 SHORT, 1 LONG), cut with the kernels' word functions;
// outputs the jobs' {qlen, tlen, h0} and their packed words back to back.  Returns the job count.
extern "C" int chain_host_cut(const uint32_t *pac, int64_t pac_words, int64_t l_pac, const uint32_t *rd, int64_t rd_words, int l_query,
                              int n_regs, const bwa_b200_region_t *regs, int batch, uint32_t *lens3, uint32_t *qwords, uint32_t *twords,
                              uint32_t *n_qw, uint32_t *n_tw)
{
    int n = 0;
    uint32_t nq = 0, nt = 0;
    read_jobs(regs, n_regs, l_query, 7u, l_pac, [&](int, int is_long, uint32_t ql, uint32_t tl, uint32_t h0, const JobAux &aux) {
        if (is_long != batch) return;
        lens3[3 * n] = ql; lens3[3 * n + 1] = tl; lens3[3 * n + 2] = h0; ++n;
        for (uint32_t w = 0; w < (ql + 7) / 8; ++w) qwords[nq++] = cut_query_word(rd, rd_words, aux, w, ql);
        for (uint32_t w = 0; w < (tl + 7) / 8; ++w) twords[nt++] = cut_target_word(pac, pac_words, l_pac, aux, w, tl);
    });
    *n_qw = nq; *n_tw = nt;
    return n;
}
