#!/usr/bin/env python
"""bench.py -- reads/s end-to-end (seed -> chain -> extend) on synthetic reads, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...

Headline workload (N = 1): BASELINE.json configs[1] -- 1 M synthetic 150 bp reads (1 % substitutions, 0.1 % indels) against a
synthetic 100 Mb genome.  A step is one pass of the pipeline the reference's `gase_aln` worker runs over a batch
(src/bwamem.c:2055-2093, 2286-2306): SMEM seeding (+ SA locate) of every read, mem_chain / mem_chain_flt, mem_chain2aln's
extension jobs for every kept chain, ksw_extend2 on all of them, the region arithmetic -- bwa_b200_align_*.  Reads shard across
ranks (weak scaling: every rank owns a full batch and a replica of the index); no data-path collective.

`value` is measured with inputs resident in HBM (CUDA events on the aligner's stream); `e2e` goes through the host-buffer C-ABI
call (bwa_b200_align_host_view) with H2D and D2H inside the timed region.  `--impl reference` times the same step over the
reference's own CPU functions (oracle/_ref: bwt_smem1 / bwt_sa of bwa_index, the fork's mem_chain .. mem_chain2aln and ksw_extend2)
on all host threads, in a process that never loads the product library.  The oracle is used only there, for the cpu_baseline legs
and for the work counters.  Other legs ride in sub_metrics: the fused one-seed step (round 1's headline), the chained step with
re-seeding, the CIGAR path, BASELINE config 3 (3.1 Gb genome, `c3`), config 4 (extension-only sweep, `c4_extension_sweep`), config 5
(seeding only on the config-3 index, `c5_seeding`) and the reference's CPU `bwa mem` wall time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from tools import synth  # noqa: E402

METRIC = "reads_per_s_end_to_end_seed_extend"
UNIT = "reads/s"
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
# contig lengths of GRCh38 chr1..22, X, Y in Mb: the proportions of config 3's 24 contigs
HG38_MB = [248.96, 242.19, 198.30, 190.21, 181.54, 170.81, 159.35, 145.14, 138.39, 133.80, 135.09, 133.28, 114.36, 107.04, 101.99, 90.34,
           83.26, 80.37, 58.62, 64.44, 46.71, 50.82, 156.04, 57.23]


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            pk = json.load(open(path))
            if isinstance(pk, dict) and float(pk.get("hbm_gbs", 0)) > 0:
                return pk, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


def workload_name(reads, read_len, genome):
    return f"{reads} synthetic {read_len}bp reads (1% sub, 0.1% indel) vs synthetic {genome} bp genome, per GPU"


def contig_lens(genome_bases, n_ctg):
    if n_ctg <= 1:
        return [genome_bases]
    tot = sum(HG38_MB[:n_ctg])
    lens = [int(genome_bases * x / tot) for x in HG38_MB[:n_ctg]]
    lens[-1] = genome_bases - sum(lens[:-1])
    return lens


class ClockSampler:
    """samples SM clock and throttle reasons through NVML while the timed region runs"""

    def __init__(self, dev):
        self.dev, self.rows, self.stop_flag, self.thr = dev, [], False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            uuid = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.dev).uuid)
            except Exception:
                pass
            h = None
            if uuid:
                for cand in ("GPU-" + uuid, uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand.encode() if isinstance(cand, str) else cand)
                        break
                    except Exception:
                        h = None
            if h is None:
                h = pynvml.nvmlDeviceGetHandleByIndex(self.dev)
            self.nv, self.h = pynvml, h
            self.thr = threading.Thread(target=self._pump, daemon=True)
            self.thr.start()
        except Exception as ex:  # noqa: BLE001
            log("clock sampler unavailable:", ex)

    def _pump(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((sm, mx, rs))
            except Exception:
                pass
            time.sleep(0.01)

    def stop(self):
        self.stop_flag = True
        if self.thr:
            self.thr.join(timeout=1)
        sm = sorted(r[0] for r in self.rows)
        reasons = set()
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        for r in self.rows:
            for bit, nm in bits.items():
                if r[2] & bit:
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max((r[1] for r in self.rows), default=None),
                "reasons": sorted(reasons), "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------------------ data
BUILD_SNIPPET = """
import sys, os
sys.path.insert(0, {root!r})
import __graft_entry__ as ge
from tools import synth
pkg = ge.load_package()
g = synth.make_genome({genome}, seed=synth.GENOME_SEED)
pkg.build_index(g, {prefix!r} + ".tmp", sa_intv={sa_intv}, also_stock_layout=True, n_threads=0)
for ext in (".bwt", ".bwt128", ".sa"):
    os.replace({prefix!r} + ".tmp" + ext, {prefix!r} + ext)
"""


def index_prefix(genome_bases, sa_intv=16):
    cache = os.environ.get("BWA_B200_CACHE", "/tmp/bwa_b200_bench")
    os.makedirs(cache, exist_ok=True)
    return os.path.join(cache, f"g{genome_bases}_s{synth.GENOME_SEED}" + ("" if sa_intv == 16 else f"_sa{sa_intv}"))


def have_index(prefix):
    return all(os.path.exists(prefix + e) for e in (".sa", ".bwt", ".bwt128"))


def prepare_data(genome_bases, n_reads, read_len, rank, dist, in_subprocess=False, sa_intv=16):
    """genome + index (built once per box by rank 0, shared through /tmp) and this rank's read batch.  in_subprocess: the index is
    built by a child process, so that the calling process (the reference arm) never maps the product library"""
    prefix = index_prefix(genome_bases, sa_intv)
    t0 = time.time()
    genome = synth.make_genome(genome_bases, seed=synth.GENOME_SEED)
    if rank == 0 and not have_index(prefix):
        if in_subprocess:
            subprocess.check_call([sys.executable, "-c", BUILD_SNIPPET.format(root=ROOT, genome=genome_bases, prefix=prefix, sa_intv=sa_intv)])
        else:
            import __graft_entry__ as ge
            pkg = ge.load_package()
            pkg.build_index(genome, prefix + ".tmp", sa_intv=sa_intv, also_stock_layout=True, n_threads=0)
            for ext in (".bwt", ".bwt128", ".sa"):
                os.replace(prefix + ".tmp" + ext, prefix + ext)
    if dist is not None:
        dist.barrier()
    t1 = time.time()
    reads, pos, strand = synth.make_reads(genome, n_reads, read_len, seed=synth.READS_SEED + rank)
    log(f"[rank {rank}] genome {genome_bases} + index {t1 - t0:.1f}s, {n_reads} reads {time.time() - t1:.1f}s")
    return genome, prefix, reads


# ----------------------------------------------------------------------------------------------- CPU reference legs
class CpuChained:
    """the chained step on the host cores with the reference's own functions (oracle/chain_py.ref_chained_pipeline): bwt_smem1 / bwt_sa of
    bwa_index through libbwaref.so, mem_chain .. mem_chain2aln and ksw_extend2 of the fork through libforkmem.so"""

    def __init__(self, genome, prefix, lens, w=100):
        from oracle import chain_py as CP, oracle_py as O
        self.CP, self.O = CP, O
        self.kind = "reference" if (O.have_ref() and CP.have_fork()) else None
        assert self.kind, "oracle/_ref is not built (libbwaref.so / libforkmem.so): run __graft_entry__.build() where /root/reference exists"
        self.h = O.ref_lib().ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
        assert self.h, "ref_load failed"
        self.ctg = CP.Contigs(tuple(lens))
        self.opt = CP.default_opt(w=w)
        self.kp = O.make_params()
        self.pac = CP.make_pac(genome)
        self.threads = O.default_threads()

    def run(self, reads2d):
        n, L = reads2d.shape
        flat = np.ascontiguousarray(reads2d).reshape(-1)
        off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
        return self.CP.ref_chained_pipeline(self.h, self.opt, self.ctg, self.pac, flat, off, self.kp, 19, self.threads)

    def close(self):
        self.O.ref_lib().ref_free(self.h)


def same_regions(cpu_out, gpu_out, n):
    """GPU regions (REGION_DTYPE) of the first n reads == the CPU arm's, field by field"""
    if not (cpu_out["n_regs"][:n] == gpu_out["n_regions"][:n]).all():
        return False
    tot = int(cpu_out["n_regs"][:n].sum())
    a, b = cpu_out["regs"][:tot], gpu_out["regions"][:tot]
    return all(bool((a[f] == b[f]).all()) for f in ("rb", "re", "qb", "qe", "score", "truesc", "rid", "seedcov", "seedlen0"))


def run_reference(args, rank):
    """--impl reference: the chained step on the CPU, all host threads, the full batch per step when that fits the time budget"""
    if rank != 0:
        return
    genome, prefix, reads = prepare_data(args.genome, args.reads, args.read_len, 0, None, in_subprocess=True)
    cpu = CpuChained(genome, prefix, [args.genome])
    n = reads.shape[0]
    probe = min(n, 50_000)
    cpu.run(reads[:probe])
    t0 = time.perf_counter()
    cpu.run(reads[:probe])
    rate = probe / (time.perf_counter() - t0)
    budget_s = args.ref_budget
    sample = n if n * (args.steps + args.warmup) / rate <= budget_s else max(probe, int(rate * budget_s / (args.steps + args.warmup)))
    sample = min(sample, n)
    for _ in range(args.warmup):
        cpu.run(reads[:sample])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = cpu.run(reads[:sample])
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    loaded = [ln.split()[-1] for ln in open("/proc/self/maps") if "libbwamem_b200" in ln]
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": workload_name(args.reads, args.read_len, args.genome), "pipeline": "seed -> chain -> extend (pass-1 SMEMs)",
                       "reads_per_step": sample, "min_seed_len": 19, "max_occ": 500, "band_w": 100, "zdrop": 100, "sa_intv": 16},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind,
                             "sample": f"{'all' if sample == n else 'first'} {sample} reads of the batch per step, {cpu.threads} host threads",
                             "regions_per_step": int(len(out["regs"])), "product_library_loaded": bool(loaded)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_bwa_mem_cpu(genome, prefix, reads, read_len, label, n_sub, threads):
    """the reference's CPU program `bwa mem -t <all cores>` (oracle/_ref/bwa7p = bwa_index/ with the .sa loader fix) on a FASTA of
    the first n_sub reads: [M::mem_process_seqs] real seconds and whole-process wall seconds (index load included)"""
    import re
    import tempfile
    exe = os.path.join(REF_DIR, "bwa7p")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/bwa7p not built"}
    work = tempfile.mkdtemp(prefix="bwamem_")
    sp = os.path.join(work, "ref")
    fasta = os.path.join(work, "ref.fa")
    t0 = time.time()
    synth.genome_to_fasta(genome, fasta)
    subprocess.check_call([exe, "fa2pac", "-f", fasta, sp], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.unlink(fasta)
    for ext, src in ((".bwt", ".bwt128"), (".sa", ".sa")):
        os.symlink(prefix + src, sp + ext)
    prep = time.time() - t0
    fa = os.path.join(work, "reads.fa")
    pos = np.zeros(n_sub, np.int64)
    synth.reads_to_fasta(reads[:n_sub], pos, pos.astype(np.uint8), fa)
    res = {"reads": n_sub, "workload": label, "fasta_pac_seconds": round(prep, 2)}
    for t in (threads, 1):
        n_here = n_sub if t > 1 else min(n_sub, 20_000)
        if t == 1 and n_here < n_sub:
            synth.reads_to_fasta(reads[:n_here], pos[:n_here], pos[:n_here].astype(np.uint8), fa)
        t0 = time.time()
        p = subprocess.run([exe, "mem", "-t", str(t), sp, fa], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
        wall = time.time() - t0
        real = sum(float(x) for x in re.findall(r"Processed \d+ reads in [\d.]+ CPU sec, ([\d.]+) real sec", p.stderr))
        key = "all_cores" if t > 1 else "one_core"
        res[key] = {"threads": t, "reads": n_here, "rc": p.returncode, "wall_seconds": round(wall, 3), "mem_process_seqs_real_seconds": round(real, 3),
                    "reads_per_s_process_seqs": n_here / real if real > 0 else None, "reads_per_s_wall": n_here / wall}
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    return res


# --------------------------------------------------------------------------------------------------- GPU legs
class Batch:
    """one rank's read batch: packed on the host, resident on the device, pinned for the host-buffer calls"""

    def __init__(self, pkg, reads):
        import torch
        self.n, self.L = reads.shape
        flat = reads.reshape(-1)
        off = (np.arange(self.n + 1, dtype=np.uint64) * np.uint64(self.L))
        self.packed, self.woff, self.rl = pkg.pack_codes(flat, off)
        self.d_packed = torch.from_numpy(self.packed.view(np.int32)).cuda()
        self.d_woff = torch.from_numpy(self.woff.view(np.int64)).cuda()
        self.d_rl = torch.from_numpy(self.rl.view(np.int32)).cuda()
        self.pin = {}
        for name, arr in (("packed", self.packed), ("woff", self.woff), ("rl", self.rl)):
            t = torch.empty(arr.nbytes, dtype=torch.uint8).pin_memory()
            t.numpy()[:] = arr.view(np.uint8)
            self.pin[name] = t
        self.h2d = int(self.packed.nbytes + self.woff.nbytes + self.rl.nbytes)
        # the compact wire layout (2 bits per base, reads of one length: no per-read array crosses the bus)
        p2, _, nl = pkg.pack2_codes(flat, off, with_lengths=False)
        self.n_n = int(nl.size)
        for name, arr in (("packed2", p2), ("nlist", nl if nl.size else np.zeros(1, np.uint64))):
            t = torch.empty(arr.nbytes, dtype=torch.uint8).pin_memory()
            t.numpy()[:] = arr.view(np.uint8)
            self.pin[name] = t
        self.h2d_compact = int(p2.nbytes + nl.nbytes)


def max_over_ranks(x, dist):
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def run_chained(args, pkg, idx, bt, flush, dist, world, lens, reseed=False, steps=None, sampler=None):
    """bwa_b200_align_*: seeding, mem_chain / mem_chain_flt / mem_chain2aln on the device, every extension job of every kept chain,
    the region arithmetic.  Device-resident timing with CUDA events on the aligner's stream, then the host-buffer call with one and
    with two batches in flight (two aligner handles, two host threads: the reference's NB_STREAMS = 2 pattern, src/fastmap.c:31,473-511)."""
    import torch
    steps = steps or args.steps
    n, L = bt.n, bt.L

    def make():
        a = pkg.Aligner(idx, n, int(bt.d_packed.numel()))
        if len(lens) > 1:
            a.set_contigs(np.concatenate([[0], np.cumsum(lens[:-1])]), lens)
        return a
    al = make()
    sp, cp, ep = pkg.seed_params(19, 500, reseed), pkg.chain_params(w=100), pkg.ext_params()
    stream = torch.cuda.ExternalStream(al.stream)

    def step():
        al.align_device(bt.d_packed.data_ptr(), bt.d_woff.data_ptr(), bt.d_rl.data_ptr(), n, L, sp, cp, ep)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = al.launches
    for k in range(steps):
        with torch.cuda.stream(stream):
            flush.zero_()                # evict L2 between timed iterations (outside the event pair)
        ev[k][0].record(stream)
        step()                           # returns when the regions are in HBM (one host round trip inside: the job count)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clocks = sampler.stop() if sampler else None
    ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dist)
    launches = al.launches - l0
    v = al.view()
    al.profile(int(os.environ.get("BWA_B200_BENCH_PROFILE_MODE", "2")))     # 2: one event pair around the extension launch set (bins overlap); 1: every bin alone
    kt = {}
    for k in range(min(steps, 5)):
        with torch.cuda.stream(stream):
            flush.zero_()
        step()
        for name, x in al.kernel_times():
            kt.setdefault(name, []).append(x)
    al.profile(False)
    # ---- host buffers in, host buffers out.  (1) the full 112-byte records through bwa_b200_align_host_view, one batch in flight;
    # (2) the compact boundary through the in-library dispatcher (bwa_b200_multi_align_compact): 2-bit reads in, 40-byte records out,
    # the batch dealt in chunks to two workers of this rank's device, so that one chunk's copies run under the other's kernels
    reps = max(2, min(steps, 10))
    pin = bt.pin
    host = al.align_host_view(pin["packed"].data_ptr(), pin["woff"].data_ptr(), pin["rl"].data_ptr(), n, sp, cp, ep, copy=True)
    n_reg = int(host["regions"].size)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        al.align_host_view(pin["packed"].data_ptr(), pin["woff"].data_ptr(), pin["rl"].data_ptr(), n, sp, cp, ep, copy=False)
    e2e_one = world * n * reps / max_over_ranks(time.perf_counter() - t0, dist)
    al.destroy()
    al = None
    chunk = max(1, (n + args.e2e_chunks - 1) // args.e2e_chunks)
    multi = pkg.MultiAligner(idx, [torch.cuda.current_device()], args.e2e_workers, chunk, L)
    if len(lens) > 1:
        multi.set_contigs(np.concatenate([[0], np.cumsum(lens[:-1])]), lens)

    def step_compact(copy=False):
        return multi.align_compact(pin["packed2"].data_ptr(), None, L, n, pin["nlist"].data_ptr() if bt.n_n else None, bt.n_n, sp, cp, ep, copy=copy, gather=copy)
    comp = step_compact(copy=True)
    step_compact()
    assert int(comp["regions"].size) == n_reg
    cu = pkg.unpack_compact(comp)
    same_c = bool((cu["n_regions"] == host["n_regions"]).all()) and all(bool((cu[f] == host["regions"][f]).all()) for f in ("rb", "re", "qb", "qe", "score", "truesc", "rid", "seedcov", "seedlen0", "w"))
    assert same_c, "compact boundary returned different regions than the full-record call"
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    l0m = multi.launches
    t0 = time.perf_counter()
    for _ in range(reps):
        step_compact()
    e2e_sync_s = max_over_ranks(time.perf_counter() - t0, dist)
    e2e_launches = (multi.launches - l0m) // reps
    # the same steps as a driver that streams batches issues them: the next batch is submitted before the previous one is waited for
    # (bwa_b200_multi_submit_compact / _wait, two batches in flight; every batch's H2D and D2H still inside the timed region)
    def submit():
        return multi.submit_compact(pin["packed2"].data_ptr(), None, L, n, pin["nlist"].data_ptr() if bt.n_n else None, bt.n_n, sp, cp, ep)
    chk = multi.wait(submit(), copy=True, gather=True)
    assert chk["regions"].tobytes() == comp["regions"].tobytes(), "a streamed batch returned different regions than the synchronous call"
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    tk = submit()
    for _ in range(reps - 1):
        tn = submit()
        multi.wait(tk, copy=False, gather=False)
        tk = tn
    multi.wait(tk, copy=False, gather=False)
    e2e_s = max_over_ranks(time.perf_counter() - t0, dist)
    multi.destroy()
    kavg = {k: float(np.mean(x)) for k, x in kt.items()}
    ext_ms = kavg.get("ext_phase", 0.0)
    d2h_full = int(n_reg * pkg.REGION_DTYPE.itemsize + n * 12)
    d2h = int(n_reg * pkg.REGION_COMPACT_DTYPE.itemsize + n * 4)
    res = {"reads_per_s": world * n * steps / (ms / 1e3), "ms_per_step": ms / steps, "steps": steps,
           "e2e_reads_per_s": world * n * reps / e2e_s, "e2e_h2d_bytes_per_step": bt.h2d_compact, "e2e_d2h_bytes_per_step": d2h,
           "e2e_chunks_per_step": int((n + chunk - 1) // chunk), "e2e_gpu_launches": int(e2e_launches), "e2e_identical_to_full_records": same_c,
           "e2e_one_batch_at_a_time_reads_per_s": world * n * reps / e2e_sync_s, "e2e_batches_in_flight": 2,
           "e2e_full_records_one_batch_in_flight": e2e_one, "e2e_full_records_h2d_bytes_per_step": bt.h2d, "e2e_full_records_d2h_bytes_per_step": d2h_full,
           "regions_per_step": int(v.n_regions), "jobs_short": int(v.n_jobs_short), "jobs_long": int(v.n_jobs_long), "seeds": int(v.n_seeds),
           "cells_per_step": int(v.cells), "extension_GCUPS": (v.cells / (ext_ms / 1e3) / 1e9) if ext_ms > 0 else None,
           "closed_form_jobs": int(v.closed_form_jobs),
           "closed_form_note": "extension jobs whose query equals the head of the target except for a few substituted bases (up to six at the default penalties, far enough apart) are answered "
                               "without a matrix (proof and conditions: csrc/ext_pair_core.cuh closed_form_job; identical results); they are "
                               "not in cells_per_step nor in extension_GCUPS",
           "gpu_launches": int(launches), "kernel_ms": kavg,
           "params": "chain w=100 max_occ=500 (mem_opt_init otherwise); extension w=100 zdrop=100 end_bonus=5 banded"
                     + ("; re-seeding split_factor 1.5 split_width 10 max_mem_intv 20" if reseed else "; SMEM pass 1 only")}
    return res, host, clocks


def run_fused(args, pkg, idx, bt, flush, dist, world):
    """round 1's headline, kept as a sub-metric: seeding, then the left / right extension of each read's longest seed only
    (bwa_b200_seed_extend_*), device-resident and through the host-buffer call"""
    import torch
    n, L = bt.n, bt.L
    pl = pkg.Pipeline(idx, n, bt.packed.size, L)
    sp, ep = pkg.SeedParams(19, 500), pkg.ext_params()
    d_out = torch.empty(n * 72, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(pl.stream)

    def step_device():
        pl.run_device(bt.d_packed.data_ptr(), bt.d_woff.data_ptr(), bt.d_rl.data_ptr(), n, L, sp, ep, d_out.data_ptr())

    pl.profile(False)
    for _ in range(args.warmup):
        step_device()
        pl.sync()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    steps = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    l0 = pl.launches
    for k in range(steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        pl.sync()
    torch.cuda.synchronize()
    launches = pl.launches - l0
    ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dist)
    ktimes, kbins = {}, {}
    for mode, dst in ((2, ktimes), (1, kbins)):      # 2: one event pair around the extension launch set; 1: a pair around every launch
        pl.profile(mode)
        for k in range(min(steps, 5)):
            with torch.cuda.stream(stream):
                flush.zero_()
            step_device()
            pl.sync()
            for name, x in pl.kernel_times():
                dst.setdefault(name, []).append(x)
    pl.profile(False)
    tot = pl.totals()
    h_out = torch.empty(n * 72, dtype=torch.uint8).pin_memory()
    pin = bt.pin

    def step_host():
        pkg.check(pkg.lib().bwa_b200_seed_extend_host(pl.h, pin["packed"].data_ptr(), pin["woff"].data_ptr(), pin["rl"].data_ptr(), n, sp, ep, h_out.data_ptr()))
    step_host()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    reps = max(2, min(steps, 10))
    t0 = time.perf_counter()
    for _ in range(reps):
        step_host()
    e2e = world * n * reps / max_over_ranks(time.perf_counter() - t0, dist)
    out_np = h_out.numpy().view(pkg.READ_RESULT_DTYPE).copy()
    kavg = {k: float(np.mean(v)) for k, v in ktimes.items()}
    ext_ms = kavg.get("ext_phase", 0.0)
    res = {"reads_per_s": world * n * steps / (ms / 1e3), "ms_per_step": ms / steps, "e2e_reads_per_s_one_batch_in_flight": e2e,
           "e2e_h2d_bytes_per_step": bt.h2d, "e2e_d2h_bytes_per_step": int(n * 72), "gpu_launches": int(launches),
           "seeds_per_step": tot["seeds"], "ext_jobs_per_step": tot["jobs"], "cells_per_step": tot["cells"],
           "extension_GCUPS": tot["cells"] / (ext_ms / 1e3) / 1e9 if ext_ms > 0 else None,
           "reads_with_seed": int((out_np["seed_qbeg"] >= 0).sum()), "kernel_ms": kavg,
           "kernel_ms_bins_serialised": {k: float(np.mean(v)) for k, v in kbins.items()}}
    pl.destroy()
    return res, out_np


def run_cigar(args, pkg, flush):
    """CIGAR path: ksw_global2 with backtrack + NM over a batch of end-to-end jobs shaped like the output stage
    of the C2 workload (150 bp queries, 3 % substitutions, 1 % short indels, band = |tlen - qlen| + 3 as bwa_gen_cigar2 gives)."""
    import torch
    from oracle import oracle_py as O
    base = synth.make_global_jobs(65_536, qlen_range=(150, 150), seed=2027)       # generated once, tiled 4 x (generation is a Python loop)
    reps_t = 4
    n = reps_t * 65_536
    jobs = {k: np.tile(base[k], reps_t) for k in ("qseq", "tseq", "qlen", "tlen", "w")}
    jobs["qoff"] = np.concatenate([base["qoff"] + np.uint32(r * base["qseq"].size) for r in range(reps_t)]).astype(np.uint32)
    jobs["toff"] = np.concatenate([base["toff"] + np.uint32(r * base["tseq"].size) for r in range(reps_t)]).astype(np.uint32)
    ep = pkg.ext_params()
    cg = pkg.Cigar(torch.cuda.current_device())
    dev = {k: torch.from_numpy(jobs[k].view(np.uint8 if k in ("qseq", "tseq") else np.int32)).cuda() for k in ("qseq", "tseq", "qoff", "toff", "qlen", "tlen")}
    stream = torch.cuda.ExternalStream(cg.stream)

    def step():
        cg.global_device(ep, n, dev["qseq"].data_ptr(), dev["qoff"].data_ptr(), dev["qlen"].data_ptr(), dev["tseq"].data_ptr(), dev["toff"].data_ptr(),
                         dev["tlen"].data_ptr(), jobs["qlen"], jobs["tlen"], jobs["w"], aligned8=True)      # make_global_jobs pads to 8

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    reps = max(3, min(args.steps, 10))
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    l0 = cg.launches
    for k in range(reps):
        with torch.cuda.stream(stream):
            flush.zero_()
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / reps
    launches = (cg.launches - l0) // reps
    cells = cg.last_cells
    cg.profile(True)
    step()
    kt = dict(cg.kernel_times())
    cg.profile(False)
    # host to host: pinned job arrays in, the handle's pinned result buffers out (bwa_b200_global_host_view); the malloc'ing call beside it
    pinj = {}
    for k in ("qseq", "tseq", "qoff", "toff", "qlen", "tlen", "w"):
        a = np.ascontiguousarray(jobs[k], np.uint8 if k in ("qseq", "tseq") else np.uint32)
        tt = torch.empty(a.nbytes, dtype=torch.uint8).pin_memory()
        tt.numpy()[:] = a.view(np.uint8)
        pinj[k] = tt

    def step_host():
        return cg.global_host_view(n, pinj["qseq"].data_ptr(), jobs["qseq"].size, pinj["qoff"].data_ptr(), pinj["qlen"].data_ptr(), pinj["tseq"].data_ptr(),
                                   jobs["tseq"].size, pinj["toff"].data_ptr(), pinj["tlen"].data_ptr(), pinj["w"].data_ptr(), ep)
    step_host()
    t0 = time.perf_counter()
    for _ in range(5):
        step_host()
    e2e_s = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    got = cg.global_host(jobs, ep)
    e2e_malloc_s = time.perf_counter() - t0
    view = step_host()
    assert all(bool((view[k] == got[k]).all()) for k in ("score", "nm", "n_cigar", "cigar")), "global_host_view differs from global_host"
    kms = sum(v for k, v in kt.items() if k.startswith("global_"))
    res = {"jobs": n, "ms_per_batch": ms, "jobs_per_s": n / (ms / 1e3), "cells": int(cells), "GCUPS": cells / (kms / 1e3) / 1e9 if kms else None,
           "kernel_ms": kt, "e2e_jobs_per_s": n / e2e_s, "e2e_jobs_per_s_pageable_malloc_call": n / e2e_malloc_s,
           "e2e_h2d_bytes": int(sum(v.numel() for v in pinj.values())), "e2e_d2h_bytes": int(n * 20 + got["cigar"].size * 4), "gpu_launches": int(launches), "cigar_ops": int(got["cigar"].size),
           "workload": "262144 jobs (65536 distinct, tiled 4 x), 150 bp queries, 3% substitutions, 1% indels of 1-4 bases, band |tlen - qlen| + 3"}
    if not args.no_cpu_baseline:
        sample = 65_536
        sj = {k: (v[:sample] if k not in ("qseq", "tseq") else v) for k, v in jobs.items()}
        threads = O.default_threads()
        O.global_batch(sj, O.make_params(), cig_stride=64, n_threads=threads)
        t0 = time.perf_counter()
        want = O.global_batch(sj, O.make_params(), cig_stride=64, n_threads=threads)
        cdt = time.perf_counter() - t0
        same = bool((got["score"][:sample] == want["score"]).all() and (got["nm"][:sample] == want["nm"]).all() and (got["n_cigar"][:sample] == want["n_cigar"]).all())
        assert same, "CIGAR kernel output differs from the oracle on the bench sample"
        res["cpu_baseline"] = {"value": sample / cdt, "unit": "jobs/s", "cores": threads, "kind": "port", "sample": f"first {sample} jobs",
                               "gpu_output_identical_on_sample": same}
    cg.destroy()
    return res


def run_mate_sw(args, pkg):
    """Mate rescue's local alignment (second half): ksw_align2 as mem_matesw calls it (src/bwamem_pair.c:159) -- a 150 bp
    mate against a window of 300-700 reference bases, xtra = KSW_XSUBO | KSW_XSTART | KSW_XBYTE | min_seed_len * a -- host buffers in,
    kswr_t records out through bwa_b200_sw_align2_host."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle_py as O
    base = synth.make_sw_jobs(16_384, qlen_range=(150, 150), tlen_range=(300, 700), seed=2031)       # generated once, tiled (generation is a Python loop)
    reps_t = 8
    n = reps_t * 16_384
    jobs = {k: np.tile(base[k], reps_t) for k in ("qseq", "tseq", "qlen", "tlen", "xtra")}
    jobs["qoff"] = np.concatenate([base["qoff"] + np.uint32(r * base["qseq"].size) for r in range(reps_t)]).astype(np.uint32)
    jobs["toff"] = np.concatenate([base["toff"] + np.uint32(r * base["tseq"].size) for r in range(reps_t)]).astype(np.uint32)
    ep = pkg.ext_params()
    la = pkg.LocalAligner(torch.cuda.current_device())
    got = la.align2_host(jobs, ep)
    la.align2_host(jobs, ep)
    reps = 3
    kms = []
    t0 = time.perf_counter()
    for _ in range(reps):
        la.align2_host(jobs, ep)
        kms.append(la.last_kernel_ms)
    e2e_s = (time.perf_counter() - t0) / reps
    cells = int((jobs["qlen"].astype(np.int64) * jobs["tlen"].astype(np.int64)).sum())
    res = {"jobs": n, "kernel_ms": float(np.mean(kms)), "jobs_per_s_kernel": n / (np.mean(kms) / 1e3), "e2e_jobs_per_s": n / e2e_s,
           "cells_first_pass": cells, "GCUPS_first_pass_cells_over_kernel_time": cells / (np.mean(kms) / 1e3) / 1e9, "gpu_launches": 1,
           "e2e_h2d_bytes": int(jobs["qseq"].size + jobs["tseq"].size + 5 * 4 * n), "e2e_d2h_bytes": int(n * 28),
           "workload": "131072 jobs (16384 distinct, tiled 8 x): 150 bp query, 300-700 bp target holding a 4 %-diverged copy of part of it (10 % none), "
                       "xtra = XSUBO | XSTART | XBYTE | 19; cells = qlen x tlen of the first pass (the second pass on the reversed prefixes is extra work)"}
    if not args.no_cpu_baseline and O.have_ref():
        sample = 16_384
        threads = O.default_threads()
        cuts = np.linspace(0, sample, threads + 1).astype(int)

        def part(k):
            sl = slice(cuts[k], cuts[k + 1])
            sj = {key: (v if key in ("qseq", "tseq") else v[sl]) for key, v in jobs.items()}
            return O.fork_sw_align2_batch(sj, O.make_params())
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(part, range(threads)))
            t0 = time.perf_counter()
            want = np.concatenate(list(ex.map(part, range(threads))))
            cdt = time.perf_counter() - t0
        same = bool(all((got[f][:sample] == want[f]).all() for f in ("score", "te", "qe", "score2", "te2", "tb", "qb")))
        assert same, "ksw_align2 on the device differs from the reference's ksw_align2 on the bench sample"
        res["cpu_baseline"] = {"value": sample / cdt, "unit": "jobs/s", "cores": threads, "kind": "reference",
                               "sample": f"first {sample} jobs, the reference's own ksw_align2 (SSE2 ksw_u8 + second pass), one slice per thread",
                               "gpu_output_identical_on_sample": same}
    la.destroy()
    return res


def run_long_reads(args, pkg, idx, genome, prefix, flush, lens):
    """Reads long enough for mem_flt_chained_seeds to act (src/bwamem.c:970-990: every seed of a kept chain scored by mem_seed_sw's local
    alignment, :774-808, the low ones dropped) through the same bwa_b200_align_* step: 20 000 x 2 000 bp against the C2 index, 5 %
    substitutions and one deletion of 1-5 bases per read; regions compared with the reference's own functions on a sample."""
    n_l, L_l = 20_000, 2_000
    rng = np.random.Generator(np.random.PCG64(4321))
    pos = rng.integers(0, genome.size - L_l - 8, size=n_l, dtype=np.int64)
    dpos = rng.integers(100, L_l - 100, size=n_l); dlen = rng.integers(1, 6, size=n_l)
    col = np.arange(L_l, dtype=np.int64)[None, :]
    reads = genome[pos[:, None] + col + (col >= dpos[:, None]) * dlen[:, None]].astype(np.uint8)
    sub = rng.random(reads.shape) < 0.05
    reads = np.where(sub, (reads + rng.integers(1, 4, size=reads.shape, dtype=np.uint8)) & 3, reads).astype(np.uint8)
    odd = np.arange(n_l) % 2 == 1                       # every other read from the reverse strand
    reads[odd] = (3 - reads[odd])[:, ::-1]
    bt = Batch(pkg, reads)
    res, host, _ = run_chained(args, pkg, idx, bt, flush, None, 1, lens, reseed=False, steps=3)
    res["workload"] = f"{n_l} x {L_l} bp reads (5% substitutions, one 1-5 base deletion each, both strands) vs the {args.genome} bp genome"
    if not args.no_cpu_baseline:
        sample = 2_000
        cpu = CpuChained(genome, prefix, lens)
        cpu.run(reads[:200])
        t0 = time.perf_counter()
        out = cpu.run(reads[:sample])
        cdt = time.perf_counter() - t0
        same = same_regions(out, host, sample)
        assert same, "long reads: GPU regions differ from the CPU reference on the bench sample"
        res["cpu_baseline"] = {"value": sample / cdt, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind,
                               "sample": f"first {sample} reads, {cpu.threads} host threads, the fork's mem_chain .. mem_flt_chained_seeds (mem_seed_sw) .. mem_chain2aln",
                               "gpu_output_identical_on_sample": same}
        cpu.close()
    return res


def seeding_roofline(pkg, local, genome, prefix, reads, kavg, n, index_bytes, pk, pk_src, idx=None, bt=None):
    """roofline block of the dominant seeding kernel: algorithmic bytes = the bucket sectors / LF steps / SA samples the reference's CPU
    algorithm touches on the same reads (instrumented oracle), over the kernel's live CUDA-event time"""
    from oracle import oracle_py as O
    oi = O.OracleIndex(prefix + ".bwt", prefix + ".sa")
    s_n = min(n, 20_000)
    L = reads.shape[1]
    sf = reads[:s_n].reshape(-1).copy()
    soff = (np.arange(s_n + 1) * L).astype(np.uint64)
    _, fc, _ = O.pipeline(oi, genome, sf, soff, O.make_params(), 19, 500, O.default_threads())
    oi.close()
    per_read = {k: v / s_n for k, v in fc.items()}
    alg = {"fwd_kernel": 32.0 * per_read["n_bucket_fwd"],
           "back_kernel": 32.0 * (per_read["n_bucket"] - per_read["n_bucket_fwd"] - per_read["n_lf"]),
           "locate_kernel": 32.0 * per_read["n_lf"] + 4.0 * per_read["n_located"]}
    seed_k = {k: kavg[k] for k in alg if k in kavg}
    if not seed_k:
        return None, per_read
    dom = max(seed_k, key=seed_k.get)
    achieved = alg[dom] * n / (seed_k[dom] / 1e3) / 1e9
    rs = float(pkg.lib().bwa_b200_measure_random_sector_gbs(local, int(index_bytes), 2048, 2))
    all_alg = sum(alg[k] for k in seed_k) * n / (sum(seed_k.values()) / 1e3) / 1e9
    issued = None
    if idx is not None and bt is not None:     # what the kernels themselves ask for: one more pass with the request counters on
        sd = pkg.Seeder(idx, bt.n, int(bt.d_packed.numel()))
        sd.request_counts(True)
        sd.seed_device(bt.d_packed.data_ptr(), bt.d_woff.data_ptr(), bt.d_rl.data_ptr(), bt.n, 19, 500)
        issued = sd.request_counts(False)
        sd.destroy()
    roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": None,
            "peak_source": pk_src, "algorithmic_bytes_per_read": alg[dom], "algorithmic_bytes_per_launch": alg[dom] * n, "ms_per_launch": seed_k[dom],
            "share_of_step": seed_k[dom] / sum(kavg.values()) if kavg else None,
            "random_sector_peak_at_index_footprint_gbs": rs, "index_bucket_bytes": int(index_bytes),
            "requests_per_launch": issued,
            "requested_bucket_gbs": (issued[dom.split("_")[0] + "_sectors"] * 32.0 / (seed_k[dom] / 1e3) / 1e9) if issued and dom in ("fwd_kernel", "back_kernel") else None,
            "frac_of_random_sector": (issued[dom.split("_")[0] + "_sectors"] * 32.0 / (seed_k[dom] / 1e3) / 1e9 / rs) if issued and rs and dom in ("fwd_kernel", "back_kernel") else None,
            "algorithmic_over_random_sector": achieved / rs if rs else None,
            "all_seeding_kernels": {"algorithmic_gbs": all_alg, "ms": sum(seed_k.values()), "kernel_ms": seed_k,
                                    "algorithmic_bytes_per_read": {k: alg[k] for k in seed_k}},
            "note": "achieved = sectors the reference's CPU algorithm touches / kernel time (SURVEY 8d); requested_bucket_gbs = 32-byte bucket sectors the "
                    "kernel itself requested (counted in the kernel) / the same time, and frac_of_random_sector = that over the measured random-sector "
                    "gather rate at the index's footprint; the k-mer interval table answers the steps on short patterns, so requested < algorithmic"}
    return roof, per_read


def run_c4(args, pkg, local, rank, world, dist, flush, int_peak_gcups_s16x2):
    """BASELINE config 4: extension-only sweep -- ksw_extend2 batches of one query length and one band each, query 100..300 bp x band
    16..100, z-drop 100 and 0, end bonus 5, h0 in [19, 150], targets of qlen + min(qlen, 2w) bases.  Every rank runs the
    whole grid on its own GPU (weak scaling; ms = max over ranks).  Jobs are packed in HBM before the timed region; GCUPS counts the
    cells ksw_extend2 evaluates (counted on the device, equal to the oracle's count on the checked sample).  Rank 0 also runs the
    reference's own ksw_extend2 on all host cores over the first 2^14 jobs of every point and asserts the results are identical."""
    import torch
    from oracle import oracle_py as O
    ex = pkg.Extender(local)
    st = torch.cuda.ExternalStream(ex.stream)
    bn = 1 << 14
    t = max(1, args.c4_jobs // bn)
    n = bn * t
    rows = []
    ref_ok = (not args.no_cpu_baseline) and rank == 0 and O.have_ref()
    for qlen in (100, 150, 200, 250, 300):
        for w in (16, 32, 50, 64, 100):
            base = synth.make_ext_jobs(bn, w=w, seed=777 + qlen + w, qlen_range=(qlen, qlen), h0_range=(19, 150))
            dq = torch.from_numpy(base["qseq"]).cuda().repeat(t)
            dt = torch.from_numpy(base["tseq"]).cuda().repeat(t)
            qp = torch.empty((dq.numel() + 7) // 8, dtype=torch.int32, device="cuda")
            tp = torch.empty((dt.numel() + 7) // 8, dtype=torch.int32, device="cuda")
            ex.pack_device(dq.data_ptr(), dq.numel(), qp.data_ptr())
            ex.pack_device(dt.data_ptr(), dt.numel(), tp.data_ptr())
            rep = torch.arange(t, device="cuda", dtype=torch.int64).repeat_interleave(bn)
            dev = {k: torch.from_numpy(base[k].astype(np.int64)).cuda().repeat(t) for k in ("qoff", "toff", "qlen", "tlen", "h0")}
            dev["qoff"] += rep * int(base["qseq"].size)
            dev["toff"] += rep * int(base["tseq"].size)
            assert int(dev["qoff"].max()) < 2 ** 32 and int(dev["toff"].max()) < 2 ** 32
            dev = {k: v.to(torch.int32) if k in ("qlen", "tlen", "h0") else (v & 0xffffffff).to(torch.int64).to(torch.int32) for k, v in dev.items()}
            res = torch.zeros(n * 6, dtype=torch.int32, device="cuda")
            del dq, dt, rep
            for zdrop in (100, 0):
                ep = pkg.ext_params(w=w, zdrop=zdrop)

                def fn():
                    ex.extend_device(ep, n, qp.data_ptr(), dev["qoff"].data_ptr(), dev["qlen"].data_ptr(), tp.data_ptr(), dev["toff"].data_ptr(),
                                     dev["tlen"].data_ptr(), dev["h0"].data_ptr(), res.data_ptr())
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                ms = 0.0
                for _ in range(args.c4_reps):
                    with torch.cuda.stream(st):
                        flush.zero_()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(st); fn(); b.record(st)
                    torch.cuda.synchronize()
                    ms += a.elapsed_time(b)
                ex.wait()
                ms = max_over_ranks(ms / args.c4_reps, dist)
                cells = int(ex.last_cells())
                row = {"qlen": qlen, "w": w, "zdrop": zdrop, "jobs": n, "ms": ms, "cells": cells, "GCUPS": world * cells / (ms / 1e3) / 1e9,
                       "Mjobs_per_s": world * n / (ms / 1e3) / 1e6, "frac_s16x2": (cells / (ms / 1e3) / 1e9) / int_peak_gcups_s16x2 if int_peak_gcups_s16x2 else None}
                if ref_ok:
                    got = res[:bn * 6].cpu().numpy().reshape(bn, 6)
                    kp = O.make_params(w=w, zdrop=zdrop)
                    want = np.zeros((bn, 6), np.int32)
                    t0 = time.perf_counter()
                    O.ref_lib().ref_ksw_batch(bn, base["qseq"], base["qoff"], base["qlen"], base["tseq"], base["toff"], base["tlen"], base["h0"], O._mat(kp),
                                              kp.o_del, kp.e_del, kp.o_ins, kp.e_ins, kp.w, kp.end_bonus, kp.zdrop, want.reshape(-1), O.default_threads())
                    cdt = time.perf_counter() - t0
                    assert (got == want).all(), f"c4: ksw_extend2 results differ from the reference at qlen {qlen} w {w} zdrop {zdrop}"
                    row["cpu_reference_GCUPS"] = (cells / t) / cdt / 1e9
                    row["identical_to_reference_on_sample"] = True
                rows.append(row)
            del qp, tp, dev, res
            torch.cuda.empty_cache()
    ex.destroy()
    out = {"workload": f"{n} jobs per point (2^14 distinct, tiled), every rank the whole grid; h0 in [19,150], 5% substitutions, 1% indels, 1% of jobs with N, "
                       "target = query + min(qlen, 2w) bases, end bonus 5, scoring 1/4/6/1", "scaling": "weak", "points": rows,
           "min_frac_s16x2": min((r["frac_s16x2"] for r in rows if r["frac_s16x2"]), default=None),
           "min_GCUPS_per_gpu": min(r["GCUPS"] for r in rows) / world, "max_GCUPS_per_gpu": max(r["GCUPS"] for r in rows) / world}
    if ref_ok:
        out["cpu_baseline"] = {"kind": "reference", "cores": O.default_threads(), "unit": "GCUPS",
                               "value": float(np.median([r["cpu_reference_GCUPS"] for r in rows])),
                               "sample": "the reference's ksw_extend2 (bwa_index/ksw.c) over the first 2^14 jobs of every point, all host threads; median over the grid",
                               "gpu_output_identical_on_sample": True}
    return out


def run_c5(args, pkg, idx, genome, prefix, local, rank, world, dist, flush, pk, pk_src):
    """BASELINE config 5: seeding only -- SMEMs (min_seed_len 19) + SA locate of 250 bp reads against the 3.1 Gb index, reads/s with
    the reads resident in HBM, pass 1 as the reference's GPU path runs it and with re-seeding; the seeding roofline; identity with the
    reference's own bwt_smem1 / bwt_sa on a sample and that CPU loop's rate on all host threads."""
    import torch
    reads, _, _ = synth.make_reads(genome, args.c5_reads, 250, seed=778 + rank)
    n, L = reads.shape
    packed, woff, rl = pkg.pack_codes(reads.reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(L)))
    d_packed = torch.from_numpy(packed.view(np.int32)).cuda(); d_woff = torch.from_numpy(woff.view(np.int64)).cuda(); d_rl = torch.from_numpy(rl.view(np.int32)).cuda()
    sd = pkg.Seeder(idx, n, packed.size)
    st = torch.cuda.ExternalStream(sd.stream)
    out = {"workload": f"{n} synthetic 250bp reads (1% sub, 0.1% indel) per GPU vs the {args.c3_genome} bp genome", "scaling": "weak", "modes": {}}
    for name, par in (("pass1", pkg.seed_params(19, 500)), ("reseed", pkg.seed_params(19, 500, True))):
        def fn():
            sd.seed_device(d_packed.data_ptr(), d_woff.data_ptr(), d_rl.data_ptr(), n, params=par)
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        ms = 0.0
        reps = 3
        for _ in range(reps):
            with torch.cuda.stream(st):
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st); fn(); b.record(st)
            torch.cuda.synchronize()
            ms += a.elapsed_time(b)
        ms = max_over_ranks(ms / reps, dist)
        out["modes"][name] = {"ms": ms, "Mreads_per_s": world * n / (ms / 1e3) / 1e6, "seeds": int(sd.device_result().n_seeds)}
    sd.destroy()
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import chain_py as CP, oracle_py as O
        s_n = min(n, args.c5_cpu_sample)
        flat = reads[:s_n].reshape(-1).copy()
        off = (np.arange(s_n + 1) * L).astype(np.uint64)
        h = O.ref_lib().ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
        CP.ref_seed_arrays(h, flat[:L * 2000], off[:2001], 19, 500, O.default_threads())
        t0 = time.perf_counter()
        want = CP.ref_seed_arrays(h, flat, off, 19, 500, O.default_threads())
        cdt = time.perf_counter() - t0
        O.ref_lib().ref_free(h)
        sp, swoff, srl = pkg.pack_codes(flat, off)
        sd2 = pkg.Seeder(idx, s_n, sp.size)
        got = sd2.seed_host(sp, swoff, srl, 19, 500)
        sd2.destroy()
        same = bool((got["n_seeds"] == want["n_seeds"]).all() and (got["rbeg"] == want["rbeg"]).all() and (got["qq"] == want["qq"]).all()
                    and (got["score"] == want["score"]).all())
        assert same, "c5: seeds differ from the reference's bwt_smem1 / bwt_sa on the bench sample"
        out["cpu_baseline"] = {"value": s_n / cdt / 1e6, "unit": "Mreads/s", "cores": O.default_threads(), "kind": "reference",
                               "sample": f"first {s_n} reads, the reference's bwt_smem1 + bwt_sa (pass 1, max_occ 500) on all host threads",
                               "gpu_output_identical_on_sample": same}
    return out


def run_c3(args, pkg, local, rank, world, dist, flush, pk, pk_src):
    """BASELINE config 3: 150 bp reads against a 3.1 Gb genome in 24 contigs (sizes in the proportions of GRCh38), index replicated per GPU.
    The chained step device-resident and host-to-host, its kernel times, the seeding roofline in the HBM regime, and identity with the
    CPU reference on a sample."""
    import torch
    lens = contig_lens(args.c3_genome, 24)
    genome, prefix, reads = prepare_data(args.c3_genome, args.c3_reads, args.read_len, rank, dist)
    t0 = time.time()
    idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", local)
    idx.attach_ref(genome)
    info = idx.info()
    load_s = time.time() - t0
    bt = Batch(pkg, reads)
    steps = max(3, min(args.steps, 5))
    res, host, _ = run_chained(args, pkg, idx, bt, flush, dist, world, lens, reseed=False, steps=steps)
    res["workload"] = workload_name(args.c3_reads, args.read_len, args.c3_genome) + ", 24 contigs"
    res["index_hbm_bytes"] = int(info.hbm_bytes)
    res["index_load_seconds"] = round(load_s, 1)
    res["scaling"] = "weak"
    if rank == 0:
        bkt_bytes = int(info.n_buckets) * 32
        roof, per_read = seeding_roofline(pkg, local, genome, prefix, reads, res["kernel_ms"], bt.n, bkt_bytes, pk, pk_src, idx, bt)
        res["roofline"] = roof
        res["oracle_work_per_read"] = per_read
        if not args.no_cpu_baseline:
            sample = min(bt.n, args.c3_cpu_sample)
            cpu = CpuChained(genome, prefix, lens)
            cpu.run(reads[:min(sample, 5000)])
            t0 = time.perf_counter()
            out = cpu.run(reads[:sample])
            cdt = time.perf_counter() - t0
            same = same_regions(out, host, sample)
            assert same, "c3: GPU regions differ from the CPU reference on the bench sample"
            res["cpu_baseline"] = {"value": sample / cdt, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind,
                                   "sample": f"first {sample} reads of the batch, {cpu.threads} host threads", "gpu_output_identical_on_sample": same}
            cpu.close()
    del bt
    torch.cuda.empty_cache()
    c5 = None
    if not args.no_c5:
        c5 = run_c5(args, pkg, idx, genome, prefix, local, rank, world, dist, flush, pk, pk_src)
    idx.free()
    torch.cuda.empty_cache()
    return res, c5


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--cpu-sample", type=int, default=200_000)
    ap.add_argument("--ref-budget", type=float, default=170.0, help="seconds the reference arm may spend in its warm-up + timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chain", action="store_true", help="only the fused one-seed step (kernel A/B runs)")
    ap.add_argument("--no-extras", action="store_true", help="skip the re-seeding, CIGAR and bwa-mem sub-metrics")
    ap.add_argument("--no-c3", action="store_true", help="skip BASELINE config 3 (3.1 Gb genome)")
    ap.add_argument("--e2e-chunks", type=int, default=2, help="chunks the host-to-host step deals its batch in")
    ap.add_argument("--e2e-workers", type=int, default=2, help="worker threads (aligner + stream each) of the rank's device in the host-to-host step")
    ap.add_argument("--c3-genome", type=int, default=3_100_000_000)
    ap.add_argument("--c3-reads", type=int, default=1_250_000, help="reads per GPU of config 3 (10 M reads over 8 GPUs)")
    ap.add_argument("--c3-cpu-sample", type=int, default=50_000)
    ap.add_argument("--no-c4", action="store_true", help="skip BASELINE config 4 (extension-only sweep)")
    ap.add_argument("--c4-jobs", type=int, default=1 << 22, help="jobs per point of the extension sweep")
    ap.add_argument("--c4-reps", type=int, default=3)
    ap.add_argument("--no-c5", action="store_true", help="skip BASELINE config 5 (seeding only, 250 bp reads, the config-3 index)")
    ap.add_argument("--c5-reads", type=int, default=500_000, help="250 bp reads per GPU of config 5 (4 M reads over 8 GPUs)")
    ap.add_argument("--c5-cpu-sample", type=int, default=20_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dref = dist if world > 1 else None

    pkg = ge.load_package()
    pkg.build()
    genome, prefix, reads = prepare_data(args.genome, args.reads, args.read_len, rank, dref)
    idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", local)
    idx.attach_ref(genome)
    info = idx.info()
    bt = Batch(pkg, reads)
    n, L = bt.n, bt.L
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")      # > L2 (126 MB)
    pk, pk_src = peaks()
    lens = [args.genome]

    fused, fused_out = run_fused(args, pkg, idx, bt, flush, dref, world)
    if args.no_chain:
        # kernel A/B mode: the fused step only, one line in the old shape
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": fused["reads_per_s"], "unit": UNIT, "n_gpus": world, "ms_per_step": fused["ms_per_step"],
                              "e2e": {"value": fused["e2e_reads_per_s_one_batch_in_flight"]}, "sub_metrics": {"kernel_ms": fused["kernel_ms"], "fused": fused}}), flush=True)
        if dref:
            dist.destroy_process_group()
        return

    sampler = ClockSampler(local)
    chained, host, clocks = run_chained(args, pkg, idx, bt, flush, dref, world, lens, reseed=False, sampler=sampler)
    chained_rs = cigar = None
    if not args.no_extras:
        chained_rs, _, _ = run_chained(args, pkg, idx, bt, flush, dref, world, lens, reseed=True, steps=max(3, min(args.steps, 10)))

    c3 = c4 = c5 = None
    ia = pkg.measure_int_alu(local)      # measured issue rates (warp-instructions per clock per SM) and the SM clock under load
    r_ = ia["warp_inst_per_clk_per_sm"]
    alu_rate = max(r_["IADD3"], r_["PRMT"], r_["VIADDMNMX.S16x2"], r_["VIMNMX3.S16x2"])
    int_peak_gops = ia["n_sm"] * alu_rate * 32 * ia["sm_mhz"] / 1e3          # thread-level integer ALU operations per second / 1e9
    if not args.no_c4:
        c4 = run_c4(args, pkg, local, rank, world, dref, flush, 2 * int_peak_gops / 15.0)
    if not args.no_c3:
        try:
            c3, c5 = run_c3(args, pkg, local, rank, world, dref, flush, pk, pk_src)
        except AssertionError:
            raise
        except Exception as ex:  # noqa: BLE001  (e.g. not enough host memory for the 6.2 G-row index build on a small box)
            log("c3 leg failed:", repr(ex))
            c3 = {"unavailable": repr(ex)[:300]}

    if rank != 0:
        if dref:
            dist.destroy_process_group()
        return
    mate_sw = long_reads = None
    if not args.no_extras:
        cigar = run_cigar(args, pkg, flush)
        mate_sw = run_mate_sw(args, pkg)
        long_reads = run_long_reads(args, pkg, idx, genome, prefix, flush, lens)

    # ---- rooflines: dominant seeding kernel (HBM sectors) and the extension launch set (INT ALU), both from live CUDA-event times
    kavg = chained["kernel_ms"]
    roofline, per_read = seeding_roofline(pkg, local, genome, prefix, reads, kavg, n, int(info.n_buckets) * 32, pk, pk_src, idx, bt)
    if roofline:
        try:   # DRAM bytes per launch of the same kernel on the same workload, from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if (n, L, args.genome) == (1_000_000, 150, 100_000_000) and roofline["kernel"] in tr["bytes_per_launch"]:
                roofline["traffic"] = tr["bytes_per_launch"][roofline["kernel"]]
                roofline["traffic_unit"] = "bytes/launch (dram read+write, ncu)"
                roofline["traffic_source"] = tr["source"]
        except Exception as ex:  # noqa: BLE001
            log("no ncu traffic figure:", ex)
    ext_ms = kavg.get("ext_phase", 0.0)
    gcups = chained["cells_per_step"] / (ext_ms / 1e3) / 1e9 if ext_ms > 0 else None
    ext_roof = {"bound": "int_alu", "achieved_gcups": gcups, "ops_per_cell": 15,
                "measured": ia, "alu_warp_inst_per_clk_per_sm": alu_rate,
                "peak_gcups_int32": int_peak_gops / 15.0, "frac_int32": gcups / (int_peak_gops / 15.0) if gcups else None,
                "peak_gcups_s16x2": 2 * int_peak_gops / 15.0, "frac_s16x2": gcups / (2 * int_peak_gops / 15.0) if gcups else None,
                "frac": gcups / (2 * int_peak_gops / 15.0) if gcups else None,
                "cells_per_step": chained["cells_per_step"], "ms_per_step": ext_ms,
                "peak_source": "bwa_b200_measure_int_alu, live: fastest of IADD3 / PRMT / VIADDMNMX.S16x2 / VIMNMX3.S16x2 x 32 lanes x SMs x measured SM clock; "
                               "x 2 for s16x2, / 15 integer operations per cell (SURVEY 8d)",
                "timing": "CUDA events around the whole extension launch set of the chained step (sort + all length bins, overlapping on side streams)"}

    cpu_baseline = bwa_mem = None
    if not args.no_cpu_baseline:
        sample = min(n, args.cpu_sample)
        cpu = CpuChained(genome, prefix, lens)
        cpu.run(reads[:min(sample, 20_000)])
        t0 = time.perf_counter()
        out = cpu.run(reads[:sample])
        cdt = time.perf_counter() - t0
        same = same_regions(out, host, sample)
        assert same, "GPU regions differ from the CPU reference on the bench sample"
        cpu_baseline = {"value": sample / cdt, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind,
                        "sample": f"first {sample} reads of the batch, {cpu.threads} host threads, same pipeline (seed -> chain -> extend) on the CPU",
                        "gpu_output_identical_on_sample": same}
        cpu.close()
        # the fused sub-metric against its own CPU counterpart (the reference's bwt_smem1 / bwt_sa / ksw_extend2 on each read's longest seed)
        from oracle import oracle_py as O
        if O.have_ref():
            s2 = min(n, 100_000)
            h = O.ref_lib().ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
            cf = reads[:s2].reshape(-1).copy()
            coff = (np.arange(s2 + 1) * L).astype(np.uint64)
            ref_out = O.ref_pipeline(h, genome, cf, coff, O.make_params(), 19, 500, O.default_threads())
            same_f = bool(ref_out.tobytes() == fused_out[:s2].tobytes())
            assert same_f, "fused step output differs from the CPU reference on the bench sample"
            fused["gpu_output_identical_on_sample"] = same_f
        if not args.no_extras:
            threads = os.cpu_count() or 1
            try:
                g1 = synth.make_genome(5_000_000, seed=synth.GENOME_SEED)
                p1 = index_prefix(5_000_000)
                if not have_index(p1):
                    pkg.build_index(g1, p1 + ".tmp", sa_intv=16, also_stock_layout=True, n_threads=0)
                    for ext in (".bwt", ".bwt128", ".sa"):
                        os.replace(p1 + ".tmp" + ext, p1 + ext)
                r1, _, _ = synth.make_reads(g1, 10_000, 150, seed=synth.READS_SEED)
                bwa_mem = {"c1": run_bwa_mem_cpu(g1, p1, r1, 150, "BASELINE config 1: 10k x 150bp vs 5 Mb", 10_000, threads),
                           "c2_subsample": run_bwa_mem_cpu(genome, prefix, reads, L, "BASELINE config 2 subsample: first 100k reads vs 100 Mb", min(n, 100_000), threads),
                           "program": "oracle/_ref/bwa7p mem (the reference's bwa_index/ sources, OCC_INTV_SHIFT 7, .sa loader fixed), default options "
                                      "(re-seeding on, w=100, zdrop=100), SAM to /dev/null"}
            except Exception as ex:  # noqa: BLE001
                bwa_mem = {"unavailable": repr(ex)[:300]}

    line = {
        "metric": METRIC, "value": chained["reads_per_s"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": chained["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": workload_name(n, L, args.genome), "pipeline": "seed -> chain -> extend (pass-1 SMEMs)",
                   "reads_per_step": n, "min_seed_len": 19, "max_occ": 500, "band_w": 100, "zdrop": 100, "sa_intv": 16,
                   "kmer_table_K": os.environ.get("BWA_B200_KMER_K", "auto: none below 256 MB of buckets, 12 below 1 GB, 13 beyond"),
                   "index_hbm_bytes": int(info.hbm_bytes), "l2_policy": "512 MB memset between timed steps (L2 flush) and inputs + workspace > L2",
                   "parallelism": f"reads sharded over {world} rank(s), index replicated, no collective"},
        "clocks": clocks, "gpu_launches": chained["gpu_launches"],
        "e2e": {"value": chained["e2e_reads_per_s"], "unit": UNIT, "h2d_bytes_per_step": chained["e2e_h2d_bytes_per_step"],
                "d2h_bytes_per_step": chained["e2e_d2h_bytes_per_step"], "chunks_per_step": chained["e2e_chunks_per_step"],
                "identical_to_full_records": chained["e2e_identical_to_full_records"],
                "full_records_one_batch_in_flight": chained["e2e_full_records_one_batch_in_flight"],
                "one_batch_at_a_time": chained["e2e_one_batch_at_a_time_reads_per_s"], "batches_in_flight": chained["e2e_batches_in_flight"],
                "how": "bwa_b200_multi_submit_compact / _wait from pinned host buffers, as a driver that streams batches calls them (the next batch "
                       "submitted before the previous one is waited for: two in flight, the reference's NB_STREAMS pattern, src/fastmap.c:31): 2-bit reads "
                       "in, 40-byte region records out, every batch dealt in chunks to %d worker threads of the rank's device, every chunk's H2D and D2H "
                       "inside the timed region; one_batch_at_a_time = the same through the synchronous bwa_b200_multi_align_compact" % args.e2e_workers,
                "workers": args.e2e_workers},
        "roofline": roofline, "roofline_extension": ext_roof, "cpu_baseline": cpu_baseline,
        "sub_metrics": {"chained": chained, "chained_reseed": chained_rs, "fused_one_seed": fused, "cigar": cigar, "mate_rescue_sw": mate_sw, "long_reads": long_reads, "c3": c3, "c4_extension_sweep": c4, "c5_seeding": c5, "bwa_mem_cpu": bwa_mem,
                        "extension_GCUPS": gcups, "oracle_work_per_read": per_read,
                        "seeding_Mreads_per_s": n / (sum(kavg[k] for k in ("fwd_kernel", "back_kernel", "fill_kernel", "locate_kernel") if k in kavg) / 1e3) / 1e6},
    }
    print(json.dumps(line), flush=True)
    if dref:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
