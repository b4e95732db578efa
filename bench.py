#!/usr/bin/env python
"""bench.py -- reads/s end-to-end (seed + extend) on synthetic reads, one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nproc-per-node N bench.py --gpus N ...

Workload at N=1: BASELINE.json configs[1] -- 1M synthetic 150 bp reads (1 % sub, 0.1 % indel) vs a
synthetic 100 Mb genome.  A step is one pass of the hot path over the batch: SMEM seeding (+ SA
locate) of every read, on-device cut of the left/right extension jobs of each read's longest seed,
and ksw_extend2 on all of them (bwa_b200_seed_extend_*).  Reads shard across ranks (weak scaling:
every rank owns a full batch and a replica of the index); no data-path collective.

`value` is measured with inputs resident in HBM (CUDA events on the pipeline stream); `e2e` goes
through the host-buffer C-ABI call with H2D/D2H inside the timed region.  The oracle is used only
for the cpu_baseline leg and for `--impl reference`.
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
import __graft_entry__ as ge  # noqa: E402
from tools import synth  # noqa: E402

METRIC = "reads_per_s_end_to_end_seed_extend"
UNIT = "reads/s"


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


def prepare_data(args, rank, world, dist):
    """genome + index (built once per box, shared through /tmp) and this rank's read batch"""
    cache = os.environ.get("BWA_B200_CACHE", "/tmp/bwa_b200_bench")
    os.makedirs(cache, exist_ok=True)
    prefix = os.path.join(cache, f"g{args.genome}_s{synth.GENOME_SEED}")
    t0 = time.time()
    genome = synth.make_genome(args.genome, seed=synth.GENOME_SEED)
    pkg = ge.load_package()
    if rank == 0 and not (os.path.exists(prefix + ".sa") and os.path.exists(prefix + ".bwt") and os.path.exists(prefix + ".bwt128")):
        pkg.build_index(genome, prefix + ".tmp", sa_intv=16, also_stock_layout=True, n_threads=0)
        for ext in (".bwt", ".bwt128", ".sa"):
            os.replace(prefix + ".tmp" + ext, prefix + ext)
    if dist is not None:
        dist.barrier()
    t1 = time.time()
    reads, pos, strand = synth.make_reads(genome, args.reads, args.read_len, seed=synth.READS_SEED + rank)
    log(f"[rank {rank}] genome+index {t1 - t0:.1f}s, reads {time.time() - t1:.1f}s")
    return genome, prefix, reads


def run_reference(args, rank, world, dist):
    """CPU arm: the reference's own bwt_smem1 / bwt_sa / ksw_extend2 (oracle/_ref, kind "reference") or,
    if that library did not travel, the oracle port; all host threads; bounded sample per step."""
    from oracle import oracle_py as O
    if rank != 0:
        return
    genome, prefix, reads = prepare_data(args, 0, 1, None)
    sample = min(args.reads, args.cpu_sample)
    f = reads[:sample].reshape(-1).copy()
    off = (np.arange(sample + 1) * args.read_len).astype(np.uint64)
    params = O.make_params()
    threads = O.default_threads()
    if O.have_ref():
        h = O.ref_lib().ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
        assert h, "ref_load failed"
        kind = "reference"
        step = lambda: O.ref_pipeline(h, genome, f, off, params, 19, 500, threads)  # noqa: E731
    else:
        oi = O.OracleIndex(prefix + ".bwt", prefix + ".sa")
        kind = "port"
        step = lambda: O.pipeline(oi, genome, f, off, params, 19, 500, threads)  # noqa: E731
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"{args.reads} synthetic {args.read_len}bp reads (1% sub, 0.1% indel) vs synthetic {args.genome} bp genome",
                       "sample_reads_per_step": sample, "min_seed_len": 19, "max_occ": 500, "band_w": 100, "zdrop": 100},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": f"first {sample} reads of the batch per step, {threads} host threads"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_chained(args, pkg, idx, d_packed, d_woff, d_rl, pin, n, L, flush, dref, world, reseed=False):
    """The same batch through bwa_b200_align_*: seeding, then mem_chain / mem_chain_flt / mem_chain2aln on the device, every
    extension job of every kept chain (not just the longest seed's), the region arithmetic -- SURVEY 8f row 1.  Extension runs
    with the same band / z-drop as the headline step.  Reported under sub_metrics.chained, timed like the headline."""
    import torch
    import torch.distributed as dist
    al = pkg.Aligner(idx, n, int(d_packed.numel()))
    # reseed: seeding also runs passes 2 and 3 of mem_collect_intv (the seed set of stock `bwa mem`)
    sp, cp, ep = pkg.seed_params(19, 500, reseed), pkg.chain_params(w=100), pkg.ext_params()
    stream = torch.cuda.ExternalStream(al.stream)

    def step():
        al.align_device(d_packed.data_ptr(), d_woff.data_ptr(), d_rl.data_ptr(), n, L, sp, cp, ep)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dref:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    l0 = al.launches
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        ev[k][0].record(stream)
        step()                       # returns when the regions are in HBM (one host round trip inside: the job count)
        ev[k][1].record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = al.launches - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dref:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    v = al.view()
    al.profile(2)
    kt = {}
    for k in range(min(args.steps, 5)):
        with torch.cuda.stream(stream):
            flush.zero_()
        step()
        for name, x in al.kernel_times():
            kt.setdefault(name, []).append(x)
    al.profile(False)
    # pinned host buffers in, regions out into the aligner's pinned result buffers (bwa_b200_align_host_view)
    reps = max(1, min(args.steps, 10))
    n_reg = int(al.align_host_view(pin["packed"].data_ptr(), pin["woff"].data_ptr(), pin["rl"].data_ptr(), n, sp, cp, ep, copy=False)["regions"].size)
    torch.cuda.synchronize()
    if dref:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        al.align_host_view(pin["packed"].data_ptr(), pin["woff"].data_ptr(), pin["rl"].data_ptr(), n, sp, cp, ep, copy=False)
    e2e_one = world * n * reps / (time.perf_counter() - t0)
    # two batches in flight: two aligner handles, two host threads (as the headline e2e)
    al2 = pkg.Aligner(idx, n, int(d_packed.numel()))
    errs = []

    def worker(a, k):
        try:
            for _ in range(k):
                a.align_host_view(pin["packed"].data_ptr(), pin["woff"].data_ptr(), pin["rl"].data_ptr(), n, sp, cp, ep, copy=False)
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    worker(al2, 1)
    torch.cuda.synchronize()
    if dref:
        dist.barrier()
    th = [threading.Thread(target=worker, args=(al, (reps + 1) // 2)), threading.Thread(target=worker, args=(al2, reps // 2))]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    e2e_s = time.perf_counter() - t0
    assert not errs, errs
    al2.destroy()
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dref:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    kavg = {k: float(np.mean(x)) for k, x in kt.items()}
    ext_ms = kavg.get("ext_phase", 0.0)
    res = {"reads_per_s": world * n * args.steps / (ms / 1e3), "ms_per_step": ms / args.steps,
           "e2e_reads_per_s": world * n * reps / e2e_s, "e2e_one_batch_in_flight": e2e_one, "e2e_d2h_bytes_per_step": int(n_reg * pkg.REGION_DTYPE.itemsize + n * 12),
           "regions_per_step": int(v.n_regions), "jobs_short": int(v.n_jobs_short), "jobs_long": int(v.n_jobs_long), "seeds": int(v.n_seeds),
           "cells_per_step": int(v.cells), "extension_GCUPS": (v.cells / (ext_ms / 1e3) / 1e9) if ext_ms > 0 else None,
           "gpu_launches": int(launches), "kernel_ms": kavg,
           "params": "chain w=100 max_occ=500 (mem_opt_init otherwise); extension w=100 zdrop=100 end_bonus=5 banded"
                     + ("; re-seeding split_factor 1.5 split_width 10 max_mem_intv 20" if reseed else "; SMEM pass 1 only")}
    al.destroy()
    return res


def run_cigar(args, pkg, flush):
    """CIGAR path: ksw_global2 with backtrack + NM over a batch of end-to-end jobs shaped like the output stage
    of the C2 workload (150 bp queries, 3 % substitutions, 1 % short indels, band = |tlen - qlen| + 3 as bwa_gen_cigar2 gives).
    Device-resident timing with CUDA events; the host-to-host call; the oracle on the host cores beside it."""
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
    t0 = time.perf_counter()
    for _ in range(3):
        got = cg.global_host(jobs, ep)
    e2e_s = (time.perf_counter() - t0) / 3
    kms = sum(v for k, v in kt.items() if k.startswith("global_"))
    res = {"jobs": n, "ms_per_batch": ms, "jobs_per_s": n / (ms / 1e3), "cells": int(cells), "GCUPS": cells / (kms / 1e3) / 1e9 if kms else None,
           "kernel_ms": kt, "e2e_jobs_per_s": n / e2e_s, "gpu_launches": int(launches), "cigar_ops": int(got["cigar"].size),
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
        res["cpu_baseline"] = {"value": sample / cdt, "unit": "jobs/s", "cores": threads, "kind": "port", "sample": f"first {sample} jobs",
                               "gpu_output_identical_on_sample": same}
    cg.destroy()
    return res


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chain", action="store_true", help="skip the seeds -> chains -> jobs -> extension -> regions stage (sub_metrics.chained)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world, None)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dref = dist if world > 1 else None

    pkg = ge.load_package()
    pkg.build()
    genome, prefix, reads = prepare_data(args, rank, world, dref)
    n, L = reads.shape
    flat = reads.reshape(-1)
    off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(L))
    packed, woff, rl = pkg.pack_codes(flat, off)

    idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", local)
    idx.attach_ref(genome)
    info = idx.info()
    pl = pkg.Pipeline(idx, n, packed.size, L)
    sp, ep = pkg.SeedParams(19, 500), pkg.ext_params()

    # resident inputs
    d_packed = torch.from_numpy(packed.view(np.int32)).cuda()
    d_woff = torch.from_numpy(woff.view(np.int64)).cuda()
    d_rl = torch.from_numpy(rl.view(np.int32)).cuda()
    d_out = torch.empty(n * 72, dtype=torch.uint8, device="cuda")
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")      # > L2 (126 MB)
    stream = torch.cuda.ExternalStream(pl.stream)

    def step_device():
        pl.run_device(d_packed.data_ptr(), d_woff.data_ptr(), d_rl.data_ptr(), n, L, sp, ep, d_out.data_ptr())

    pl.profile(False)
    for _ in range(args.warmup):
        step_device()
        pl.sync()
    torch.cuda.synchronize()
    if dref:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = pl.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for k in range(args.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                       # evict L2 between timed iterations (outside the event pair)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        pl.sync()
    torch.cuda.synchronize()
    if dref:
        dist.barrier()
    clocks = sampler.stop()
    gpu_launches = pl.launches - launches0
    # per-kernel durations: the same steps again with CUDA events on the pipeline stream.  Mode 2 = an event pair around
    # every seeding / glue kernel and ONE pair around the extension launch set, whose length bins overlap on side streams
    # exactly as in the timed steps above; mode 1 = a pair around every launch, which runs the bins one after another
    # (the per-bin breakdown).
    ktimes, ktimes_bins = {}, {}
    for mode, dst in ((2, ktimes), (1, ktimes_bins)):
        pl.profile(mode)
        for k in range(args.steps if mode == 2 else min(args.steps, 5)):
            with torch.cuda.stream(stream):
                flush.zero_()
            step_device()
            pl.sync()
            for name, ms in pl.kernel_times():
                dst.setdefault(name, []).append(ms)
    torch.cuda.synchronize()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    tot = pl.totals()
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dref:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * n * args.steps / (total_ms / 1e3)

    # ---- e2e: host buffers through the C ABI (H2D + kernels + D2H per step)
    pin = {}
    for name, arr in (("packed", packed), ("woff", woff), ("rl", rl)):
        tns = torch.empty(arr.nbytes, dtype=torch.uint8).pin_memory()
        tns.numpy()[:] = arr.view(np.uint8)
        pin[name] = tns
    h_out = torch.empty(n * 72, dtype=torch.uint8).pin_memory()
    out_np = h_out.numpy().view(pkg.READ_RESULT_DTYPE)

    def step_host():
        pkg.check(pkg.lib().bwa_b200_seed_extend_host(pl.h, pin["packed"].data_ptr(), pin["woff"].data_ptr(), pin["rl"].data_ptr(),
                                                     n, sp, ep, h_out.data_ptr()))

    pl.profile(False)
    step_host()
    torch.cuda.synchronize()
    if dref:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dref:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_sync = world * n * args.steps / float(t.item())
    # the same call with two batches in flight -- two pipeline handles driven by two host threads, which is how the reference's
    # driver uses its own boundary (NB_STREAMS = 2 storages per worker thread, src/fastmap.c:31,473-511): every step still pays
    # its H2D and D2H inside the timed region, but they overlap the other handle's kernels
    pl2 = pkg.Pipeline(idx, n, packed.size, L)
    h_out2 = torch.empty(n * 72, dtype=torch.uint8).pin_memory()

    errs = []

    def worker(handle, out_t, k):
        try:
            for _ in range(k):
                pkg.check(pkg.lib().bwa_b200_seed_extend_host(handle, pin["packed"].data_ptr(), pin["woff"].data_ptr(), pin["rl"].data_ptr(),
                                                             n, sp, ep, out_t.data_ptr()))
        except Exception as ex:  # noqa: BLE001
            errs.append(ex)

    worker(pl2.h, h_out2, 1)
    assert h_out2.numpy().tobytes() == h_out.numpy().tobytes()
    torch.cuda.synchronize()
    if dref:
        dist.barrier()
    k1, k2 = (args.steps + 1) // 2, args.steps // 2
    th = [threading.Thread(target=worker, args=(pl.h, h_out, k1)), threading.Thread(target=worker, args=(pl2.h, h_out2, k2))]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    torch.cuda.synchronize()
    assert not errs, errs
    e2e_s2 = time.perf_counter() - t0
    t = torch.tensor([e2e_s2], dtype=torch.float64, device="cuda")
    if dref:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(t.item())
    pl2.destroy()
    h2d = int(packed.nbytes + woff.nbytes + rl.nbytes)
    d2h = int(n * 72)
    mapped = int((out_np["seed_qbeg"] >= 0).sum())

    chained = chained_rs = None
    if not args.no_chain:
        chained = run_chained(args, pkg, idx, d_packed, d_woff, d_rl, pin, n, L, flush, dref, world)
        chained_rs = run_chained(args, pkg, idx, d_packed, d_woff, d_rl, pin, n, L, flush, dref, world, reseed=True)

    if rank != 0:
        if dref:
            dist.destroy_process_group()
        return
    cigar = run_cigar(args, pkg, flush) if not args.no_chain else None

    # ---- roofline of the dominant kernel (CUDA-event time per launch, live, over the timed steps)
    pk, pk_src = peaks()
    kavg = {k: float(np.mean(v)) for k, v in ktimes.items()}
    step_kernel_ms = sum(kavg.values())
    from oracle import oracle_py as O
    oi = O.OracleIndex(prefix + ".bwt", prefix + ".sa")
    s_n = min(n, 20_000)
    sf = reads[:s_n].reshape(-1).copy()
    soff = (np.arange(s_n + 1) * L).astype(np.uint64)
    _, fc, kc = O.pipeline(oi, genome, sf, soff, O.make_params(), 19, 500, O.default_threads())
    per_read = {k: v / s_n for k, v in fc.items()}
    alg = {   # algorithmic bytes per read of each seeding kernel: 32 B per distinct bucket / LF step, 4 B per SA sample
        "fwd_kernel": 32.0 * per_read["n_bucket_fwd"],
        "back_kernel": 32.0 * (per_read["n_bucket"] - per_read["n_bucket_fwd"] - per_read["n_lf"]),
        "locate_kernel": 32.0 * per_read["n_lf"] + 4.0 * per_read["n_located"],
    }
    seed_k = {k: kavg[k] for k in alg if k in kavg}
    dom = max(seed_k, key=seed_k.get) if seed_k else None
    roofline = None
    if dom:
        achieved = alg[dom] * n / (seed_k[dom] / 1e3) / 1e9
        roofline = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / pk["hbm_gbs"], "traffic": None, "peak_source": pk_src,
                    "algorithmic_bytes_per_read": alg[dom], "ms_per_launch": seed_k[dom], "share_of_step": seed_k[dom] / step_kernel_ms}
    rs_hbm = float(pkg.lib().bwa_b200_measure_random_sector_gbs(local, 8 << 30, 64, 3))
    rs_l2 = float(pkg.lib().bwa_b200_measure_random_sector_gbs(local, 64 << 20, 64, 3))
    rs_mid = float(pkg.lib().bwa_b200_measure_random_sector_gbs(local, 1 << 30, 64, 3))
    if roofline:
        # random 32-byte-sector gather rates measured live (seed.cu random_sector_kernel): 8 GB buffer (HBM, beyond the
        # TLB reach), 1 GB buffer (HBM, inside the TLB reach: the regime of a 100 Mb .. 1 Gb index) and 64 MB (L2)
        roofline["random_sector_peak_hbm_gbs"] = rs_hbm
        roofline["random_sector_peak_hbm_1gb_gbs"] = rs_mid
        roofline["random_sector_peak_l2_gbs"] = rs_l2
        roofline["frac_of_random_sector_hbm"] = roofline["achieved"] / rs_hbm if rs_hbm else None
        roofline["frac_of_random_sector_hbm_1gb"] = roofline["achieved"] / rs_mid if rs_mid else None
        try:   # DRAM bytes per launch of the same kernel on the same workload, from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if (n, L, args.genome) == (1_000_000, 150, 100_000_000) and dom in tr["bytes_per_launch"]:
                roofline["traffic"] = tr["bytes_per_launch"][dom]
                roofline["traffic_unit"] = "bytes/launch (dram read+write, ncu)"
                roofline["traffic_gbs"] = tr["bytes_per_launch"][dom] / (seed_k[dom] / 1e3) / 1e9
                roofline["algorithmic_bytes_per_launch"] = alg[dom] * n
                roofline["traffic_source"] = tr["source"]
        except Exception as ex:  # noqa: BLE001
            log("no ncu traffic figure:", ex)
    kbins = {k: float(np.mean(v)) for k, v in ktimes_bins.items()}
    ext_ms = kavg.get("ext_phase", 0.0)          # sort + every bin, bins overlapping: what the extension costs inside a step
    ext_ms_serial = sum(v for k, v in kbins.items() if k.startswith("ext_inter_kernel") or k.startswith("ext_pair_kernel"))
    gcups = tot["cells"] / (ext_ms / 1e3) / 1e9 if ext_ms > 0 else None
    seed_ms = sum(seed_k.values())
    # INT-ALU roofline of the extension kernel: 15 integer ops per cell, 64 int lanes/clk/SM on the ALU pipe
    int_peak_gops = 148 * 64 * (pk.get("sm_max_mhz", 1965.0) / 1e3)
    ext_roof = {"bound": "int_alu", "achieved_gcups": gcups, "peak_gcups_int32": int_peak_gops / 15.0,
                "frac_int32": gcups / (int_peak_gops / 15.0) if gcups else None,
                "peak_gcups_s16x2": 2 * int_peak_gops / 15.0, "frac_s16x2": gcups / (2 * int_peak_gops / 15.0) if gcups else None,
                "ops_per_cell": 15,
                "cells_per_step": tot["cells"], "ms_per_step": ext_ms,
                "timing": "CUDA events around the whole extension launch set (sort + all length bins, overlapping on side streams)",
                "gcups_bins_serialised": tot["cells"] / (ext_ms_serial / 1e3) / 1e9 if ext_ms_serial > 0 else None,
                "ms_bins_serialised": ext_ms_serial}

    cpu_baseline = None
    if not args.no_cpu_baseline:
        sample = min(n, args.cpu_sample)
        cf = reads[:sample].reshape(-1).copy()
        coff = (np.arange(sample + 1) * L).astype(np.uint64)
        threads = O.default_threads()
        if O.have_ref():
            h = O.ref_lib().ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
            kind, fn = "reference", (lambda: O.ref_pipeline(h, genome, cf, coff, O.make_params(), 19, 500, threads))
        else:
            kind, fn = "port", (lambda: O.pipeline(oi, genome, cf, coff, O.make_params(), 19, 500, threads))
        fn()
        t0 = time.perf_counter()
        ref_out = fn()
        cdt = time.perf_counter() - t0
        ref_out = ref_out[0] if isinstance(ref_out, tuple) else ref_out
        same = bool(ref_out.tobytes() == out_np[:sample].tobytes())
        cpu_baseline = {"value": sample / cdt, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": f"first {sample} reads of the batch, {threads} host threads, same pipeline on CPU",
                        "gpu_output_identical_on_sample": same}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"{n} synthetic {L}bp reads (1% sub, 0.1% indel) vs synthetic {args.genome} bp genome, per GPU",
                   "min_seed_len": 19, "max_occ": 500, "band_w": 100, "zdrop": 100, "sa_intv": 16,
                   "index_hbm_bytes": int(info.hbm_bytes), "l2_policy": "512 MB memset between timed steps (L2 flush) and inputs + workspace > L2",
                   "parallelism": f"reads sharded over {world} rank(s), index replicated, no collective"},
        "clocks": clocks, "gpu_launches": int(gpu_launches),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "batches_in_flight": 2,
                "value_one_batch_in_flight": e2e_sync,
                "how": "bwa_b200_seed_extend_host with pinned host buffers, two handles driven by two host threads (the reference's NB_STREAMS = 2 pattern)"},
        "roofline": roofline, "roofline_extension": ext_roof, "cpu_baseline": cpu_baseline,
        "sub_metrics": {"seeding_Mreads_per_s": n / (seed_ms / 1e3) / 1e6 if seed_ms else None, "extension_GCUPS": gcups,
                        "seeds_per_step": tot["seeds"], "ext_jobs_per_step": tot["jobs"], "reads_with_seed": mapped,
                        "kernel_ms": kavg, "kernel_ms_bins_serialised": kbins, "oracle_work_per_read": per_read, "chained": chained, "chained_reseed": chained_rs, "cigar": cigar},
    }
    print(json.dumps(line), flush=True)
    if dref:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
