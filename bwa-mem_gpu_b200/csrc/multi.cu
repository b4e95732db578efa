// multi.cu -- in-library dispatcher over the GPUs of one box.
//
// The reference runs one GPU: its driver never selects a device (src/fastmap.c:143 has gasal_set_device commented out).  Reads
// and extension jobs are independent units, so the batch shards with no exchange: a host dispatcher deals chunks of reads to
// worker threads, every device holds a replica of the index (bwa_b200_index_clone_to: peer copies over NVLink, the disk and
// PCIe are touched once) and each worker owns an aligner -- its own stream set, arenas and pinned staging.  Two workers per
// device keep two chunks in flight there, so one chunk's copies overlap the other's kernels (the reference's NB_STREAMS = 2,
// src/fastmap.c:31).  Chunks are claimed from an atomic counter (a slow chunk does not hold a device's queue), the per-read
// region counts land in read order, and the region records of a chunk land in one segment of a pinned arena reserved when the
// chunk's count is known; chunk_region_off[] tells where.  No collective: nothing is exchanged between devices.
#include "internal.h"
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

int b200_align_compact(bwa_b200_aligner *a, const uint32_t *packed2, const uint32_t *read_len, uint32_t uniform_len, uint64_t n_reads,
                       const uint64_t *n_list, uint64_t n_n, const bwa_b200_seed_params_t *sp, const bwa_b200_chain_params_t *cp,
                       const bwa_b200_ext_params_t *ep, uint32_t *dst_nregs, uint64_t *n_regions,
                       const std::function<bwa_b200_region_compact_t *(uint64_t)> &reserve);

struct bwa_b200_multi {
    struct Worker { int device = 0; bwa_b200_aligner_t *al = nullptr; std::thread th; uint64_t chunks = 0; };
    std::vector<int> devices;
    std::vector<bwa_b200_index_t *> replicas;       // replicas[0] is the caller's index, the others are owned
    std::vector<Worker> workers;
    uint64_t chunk_reads = 0;
    // Two batches can be in flight (bwa_b200_multi_submit_compact / _wait): the workers drain the older one first and move on to the
    // next without returning to the caller, so the last chunk's D2H of one batch and the first chunk's H2D of the next overlap kernels.
    struct Slot {
        // the batch (parameter blocks copied at submit; the read buffers stay the caller's until wait returns)
        const uint32_t *packed2 = nullptr; const uint32_t *read_len = nullptr; uint32_t uniform_len = 0; uint64_t n_reads = 0;
        const uint64_t *n_list = nullptr; uint64_t n_n = 0;
        bwa_b200_seed_params_t sp; bwa_b200_chain_params_t cp; bwa_b200_ext_params_t ep;
        std::vector<uint64_t> word0, n0;            // per chunk: first 2-bit word, first entry of the N list
        uint64_t n_chunks = 0;
        // pinned outputs
        uint32_t *p_nregs = nullptr; uint64_t nregs_cap = 0;
        bwa_b200_region_compact_t *p_regions = nullptr; uint64_t region_cap = 0;
        std::vector<uint64_t> chunk_off, chunk_regs;
        // progress (next_chunk / done_chunks / state under mu)
        uint64_t next_chunk = 0, done_chunks = 0, seq = 0;
        std::atomic<uint64_t> arena_used{0};
        std::atomic<int> rc{0};
        std::string err;
        int state = 0;                              // 0 free, 1 in flight, 2 finished, 3 results handed out (reusable)
    } slot[2];
    std::mutex mu;
    std::condition_variable cv_go, cv_done;
    uint64_t next_seq = 1; bool quit = false;
};

static void multi_worker(bwa_b200_multi *m, int wi)
{
    bwa_b200_multi::Worker &w = m->workers[wi];
    cudaSetDevice(w.device);
    for (;;) {
        bwa_b200_multi::Slot *c = nullptr;
        uint64_t k = 0;
        {
            std::unique_lock<std::mutex> lk(m->mu);
            m->cv_go.wait(lk, [&] {
                if (m->quit) return true;
                c = nullptr;
                for (auto &s : m->slot)            // the older batch first
                    if (s.state == 1 && s.next_chunk < s.n_chunks && (!c || s.seq < c->seq)) c = &s;
                return c != nullptr;
            });
            if (m->quit) return;
            k = c->next_chunk++;
        }
        if (!c->rc.load()) {
            const uint64_t r0 = k * m->chunk_reads, nr = std::min(m->chunk_reads, c->n_reads - r0);
            const uint64_t nl0 = c->n0[k], nl1 = c->n0[k + 1];
            std::vector<uint64_t> local_n;                        // the chunk's slice of the N list, read indexes made chunk-relative
            if (nl1 > nl0) {
                local_n.assign(c->n_list + nl0, c->n_list + nl1);
                for (uint64_t &e : local_n) e -= r0 << 32;
            }
            uint64_t n_regions = 0;
            auto reserve = [&](uint64_t n) -> bwa_b200_region_compact_t * {
                const uint64_t off = c->arena_used.fetch_add(n);
                c->chunk_off[k] = off; c->chunk_regs[k] = n;
                return off + n <= c->region_cap ? c->p_regions + off : nullptr;
            };
            const int rc = b200_align_compact(w.al, c->packed2 + c->word0[k], c->read_len ? c->read_len + r0 : nullptr, c->uniform_len, nr,
                                              local_n.empty() ? nullptr : local_n.data(), local_n.size(), &c->sp, &c->cp, &c->ep,
                                              c->p_nregs + r0, &n_regions, reserve);
            if (rc) {
                int expect = 0;
                if (c->rc.compare_exchange_strong(expect, rc)) { std::lock_guard<std::mutex> lk(m->mu); c->err = bwa_b200_last_error(); }
            } else ++w.chunks;
        }
        {
            std::lock_guard<std::mutex> lk(m->mu);
            if (++c->done_chunks == c->n_chunks) { c->state = 2; m->cv_done.notify_all(); }
        }
    }
}

extern "C" int bwa_b200_multi_create(const bwa_b200_index_t *idx, const int *devices, int n_devices, int workers_per_device,
                                     uint64_t chunk_reads, uint32_t max_read_len, bwa_b200_multi_t **out)
{
    if (!idx || !out || n_devices < 1 || !devices || workers_per_device < 1 || workers_per_device > 8 || !chunk_reads || !max_read_len) {
        b200::set_error("multi_create: bad argument"); return BWA_B200_ERR_ARG;
    }
    if (devices[0] != idx->device) { b200::set_error("multi_create: devices[0] must be the device the index lives on"); return BWA_B200_ERR_ARG; }
    bwa_b200_multi *m = new bwa_b200_multi();
    m->chunk_reads = chunk_reads;
    m->devices.assign(devices, devices + n_devices);
    m->replicas.push_back(const_cast<bwa_b200_index_t *>(idx));
    int rc = BWA_B200_OK;
    for (int d = 1; d < n_devices && !rc; ++d) {
        bwa_b200_index_t *rep = nullptr;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, devices[d], idx->device);
        if (can) { cudaSetDevice(devices[d]); cudaDeviceEnablePeerAccess(idx->device, 0); cudaGetLastError(); }
        rc = bwa_b200_index_clone_to(idx, devices[d], &rep);
        if (!rc) m->replicas.push_back(rep);
    }
    const uint64_t max_words = chunk_reads * (((uint64_t)max_read_len + 7) / 8);
    m->workers.resize((size_t)n_devices * workers_per_device);
    for (size_t wi = 0; wi < m->workers.size() && !rc; ++wi) {
        const int d = (int)(wi % (size_t)n_devices);               // workers of a device are not neighbours in the claim order
        m->workers[wi].device = devices[d];
        rc = bwa_b200_aligner_create(m->replicas[d], chunk_reads, max_words, &m->workers[wi].al);
    }
    if (rc) { bwa_b200_multi_destroy(m); return rc; }
    for (size_t wi = 0; wi < m->workers.size(); ++wi) m->workers[wi].th = std::thread(multi_worker, m, (int)wi);
    cudaSetDevice(idx->device);
    *out = m;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_multi_set_contigs(bwa_b200_multi_t *m, int32_t n, const int64_t *offset, const int32_t *len, const int32_t *is_alt)
{
    if (!m) return BWA_B200_ERR_ARG;
    for (auto &w : m->workers) {
        const int rc = bwa_b200_aligner_set_contigs(w.al, n, offset, len, is_alt);
        if (rc) return rc;
    }
    return BWA_B200_OK;
}

extern "C" void bwa_b200_multi_destroy(bwa_b200_multi_t *m)
{
    if (!m) return;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        m->quit = true;
    }
    m->cv_go.notify_all();
    for (auto &w : m->workers) if (w.th.joinable()) w.th.join();
    for (auto &w : m->workers) if (w.al) bwa_b200_aligner_destroy(w.al);
    for (size_t d = 1; d < m->replicas.size(); ++d) bwa_b200_index_free(m->replicas[d]);
    for (auto &sl : m->slot) { cudaFreeHost(sl.p_nregs); cudaFreeHost(sl.p_regions); }
    delete m;
}

extern "C" int bwa_b200_multi_n_workers(const bwa_b200_multi_t *m) { return m ? (int)m->workers.size() : 0; }
extern "C" uint64_t bwa_b200_multi_worker_chunks(const bwa_b200_multi_t *m, int worker) { return m && worker >= 0 && worker < (int)m->workers.size() ? m->workers[worker].chunks : 0; }
extern "C" uint64_t bwa_b200_multi_launches(const bwa_b200_multi_t *m)
{
    uint64_t n = 0;
    if (m) for (auto &w : m->workers) n += bwa_b200_aligner_launches(w.al);
    return n;
}

// (re)start slot c on its stored batch: counters reset, workers woken
static void multi_launch(bwa_b200_multi *m, bwa_b200_multi::Slot &c)
{
    c.chunk_off.assign(c.n_chunks, 0); c.chunk_regs.assign(c.n_chunks, 0);
    c.arena_used = 0; c.rc = 0;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        c.next_chunk = 0; c.done_chunks = 0; c.err.clear();
        c.seq = m->next_seq++;
        c.state = 1;
    }
    m->cv_go.notify_all();
}

extern "C" int bwa_b200_multi_submit_compact(bwa_b200_multi_t *m, const uint32_t *packed2, const uint32_t *read_len, uint32_t uniform_len,
                                             uint64_t n_reads, const uint64_t *n_list, uint64_t n_n, const bwa_b200_seed_params_t *sp,
                                             const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep, int *ticket)
{
    if (!m || !sp || !cp || !ep || !ticket || !n_reads || !packed2 || (n_n && !n_list)) { b200::set_error("multi_submit_compact: bad argument"); return BWA_B200_ERR_ARG; }
    bwa_b200_multi::Slot *c = nullptr;
    {   // a free slot, else the one whose results were handed out longest ago
        std::lock_guard<std::mutex> lk(m->mu);
        for (auto &s : m->slot) if (s.state == 0) { c = &s; break; }
        if (!c) for (auto &s : m->slot) if (s.state == 3 && (!c || s.seq < c->seq)) c = &s;
        if (c) c->state = 0;
    }
    if (!c) { b200::set_error("multi_submit_compact: two batches are in flight already; wait for one"); return BWA_B200_ERR_CAPACITY; }
    c->packed2 = packed2; c->read_len = read_len; c->uniform_len = uniform_len; c->n_reads = n_reads; c->n_list = n_list; c->n_n = n_n;
    c->sp = *sp; c->cp = *cp; c->ep = *ep;
    c->n_chunks = (n_reads + m->chunk_reads - 1) / m->chunk_reads;
    c->word0.assign(c->n_chunks + 1, 0); c->n0.assign(c->n_chunks + 1, 0);
    {   // where every chunk's words and N entries start (the N list must be sorted by read)
        uint64_t wsum = 0, ni = 0;
        for (uint64_t k = 0; k < c->n_chunks; ++k) {
            const uint64_t r0 = k * m->chunk_reads, r1 = std::min(n_reads, r0 + m->chunk_reads);
            c->word0[k] = wsum; c->n0[k] = ni;
            if (read_len) for (uint64_t r = r0; r < r1; ++r) wsum += ((uint64_t)read_len[r] + 15) >> 4;
            else wsum += (r1 - r0) * (((uint64_t)uniform_len + 15) >> 4);
            while (ni < n_n && (n_list[ni] >> 32) < r1) {
                if ((n_list[ni] >> 32) < r0) { b200::set_error("multi_align_compact: the N list must be sorted by read"); return BWA_B200_ERR_ARG; }
                ++ni;
            }
        }
        c->word0[c->n_chunks] = wsum; c->n0[c->n_chunks] = ni;
        if (ni != n_n) { b200::set_error("multi_align_compact: the N list names reads beyond the batch or is not sorted"); return BWA_B200_ERR_ARG; }
    }
    cudaSetDevice(m->devices[0]);
    if (n_reads > c->nregs_cap) {
        cudaFreeHost(c->p_nregs); c->p_nregs = nullptr; c->nregs_cap = 0;
        B200_CUDA(cudaHostAlloc(&c->p_nregs, (n_reads + n_reads / 8) * 4, cudaHostAllocPortable));
        c->nregs_cap = n_reads + n_reads / 8;
    }
    // mem_chain2aln rarely keeps more than a few regions per read: room for four, doubled by wait (and the batch repeated) when it needs more
    const uint64_t cap = std::max<uint64_t>(c->region_cap, 4 * n_reads + 1024);
    if (cap > c->region_cap) {
        cudaFreeHost(c->p_regions); c->p_regions = nullptr; c->region_cap = 0;
        B200_CUDA(cudaHostAlloc(&c->p_regions, cap * sizeof(bwa_b200_region_compact_t), cudaHostAllocPortable));
        c->region_cap = cap;
    }
    multi_launch(m, *c);
    *ticket = (int)(c - m->slot);
    return BWA_B200_OK;
}

extern "C" int bwa_b200_multi_wait(bwa_b200_multi_t *m, int ticket, bwa_b200_multi_result_t *out)
{
    if (!m || !out || ticket < 0 || ticket > 1) { b200::set_error("multi_wait: bad argument"); return BWA_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    bwa_b200_multi::Slot &c = m->slot[ticket];
    for (int attempt = 0;; ++attempt) {
        {
            std::unique_lock<std::mutex> lk(m->mu);
            if (c.state != 1 && c.state != 2) { b200::set_error("multi_wait: no batch behind this ticket"); return BWA_B200_ERR_ARG; }
            m->cv_done.wait(lk, [&] { return c.state == 2; });
        }
        if (c.rc.load() == BWA_B200_ERR_CAPACITY && c.arena_used.load() > c.region_cap && attempt < 5) {     // more regions than the arena holds: grow, repeat
            cudaSetDevice(m->devices[0]);
            const uint64_t cap = std::max<uint64_t>(c.region_cap << 1, c.arena_used.load() + 1024);
            cudaFreeHost(c.p_regions); c.p_regions = nullptr; c.region_cap = 0;
            B200_CUDA(cudaHostAlloc(&c.p_regions, cap * sizeof(bwa_b200_region_compact_t), cudaHostAllocPortable));
            c.region_cap = cap;
            multi_launch(m, c);
            continue;
        }
        break;
    }
    const int rc = c.rc.load();
    {
        std::lock_guard<std::mutex> lk(m->mu);
        c.state = rc ? 0 : 3;
    }
    if (rc) { b200::set_error("multi_align_compact: %s", c.err.c_str()); return rc; }
    out->n_reads = c.n_reads; out->n_regions = c.arena_used.load(); out->n_chunks = c.n_chunks; out->chunk_reads = m->chunk_reads;
    out->n_regions_per_read = c.p_nregs; out->chunk_region_off = c.chunk_off.data(); out->regions = c.p_regions;
    return BWA_B200_OK;
}

extern "C" int bwa_b200_multi_align_compact(bwa_b200_multi_t *m, const uint32_t *packed2, const uint32_t *read_len, uint32_t uniform_len,
                                            uint64_t n_reads, const uint64_t *n_list, uint64_t n_n, const bwa_b200_seed_params_t *sp,
                                            const bwa_b200_chain_params_t *cp, const bwa_b200_ext_params_t *ep, bwa_b200_multi_result_t *out)
{
    if (!m || !sp || !cp || !ep || !out || (n_reads && !packed2) || (n_n && !n_list)) { b200::set_error("multi_align_compact: bad argument"); return BWA_B200_ERR_ARG; }
    memset(out, 0, sizeof(*out));
    if (n_reads == 0) return BWA_B200_OK;
    int ticket = -1;
    const int rc = bwa_b200_multi_submit_compact(m, packed2, read_len, uniform_len, n_reads, n_list, n_n, sp, cp, ep, &ticket);
    if (rc) return rc;
    return bwa_b200_multi_wait(m, ticket, out);
}
