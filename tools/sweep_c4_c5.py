#!/usr/bin/env python
"""BASELINE.json configs[3] and [4] on one GPU, as tables (not bench lines):
  C4  extension-only sweep: ksw_extend2 batches, query 100..300 bp x band 16..100, zdrop 100, end bonus 5 -> GCUPS, jobs/s
  C5  seeding-only sweep:   SMEM seeding (min_seed_len 19) + SA locate of 250 bp reads -> Mreads/s   (index size = --genome)
Device-resident timing with CUDA events (inputs in HBM), 512 MB L2 flush between timed repetitions.
  python tools/sweep_c4_c5.py [--jobs N] [--reads N] [--genome BASES] > profiles/...json"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from tools import synth


def timed(stream, fn, flush, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(reps):
        with torch.cuda.stream(stream):
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream)
        torch.cuda.synchronize()
        ms += a.elapsed_time(b)
    return ms / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=1 << 19)
    ap.add_argument("--reads", type=int, default=500_000)
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only-long", action="store_true", help="only the long-read extension points")
    ap.add_argument("--only-c4", action="store_true", help="skip the seeding sweep")
    ap.add_argument("--zdrop", type=int, default=100)
    args = ap.parse_args()
    pkg = ge.load_package(); pkg.build()
    torch.cuda.set_device(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out = {"c4_extension": [], "c5_seeding": None}
    # ---- C4
    ex = pkg.Extender(0)
    st = torch.cuda.ExternalStream(ex.stream)
    base_n = 1 << 14
    points = [(q, w, base_n, args.jobs) for q in (100, 150, 200, 250, 300) for w in (16, 32, 50, 64, 100)]
    points += [(2000, 100, 512, 8192), (10000, 100, 128, 2048),          # long reads, small batches: one job per warp (ext_wave_kernel)
               (1000, 100, 1024, 131072), (2000, 100, 512, 65536), (10000, 100, 128, 16384)]   # large batches: one job per lane (ext_pair_kernel<WIDE>)
    if args.only_long:
        points = points[-5:]
    for qlen, w, bn, total in points:
        if True:
            base = synth.make_ext_jobs(bn, w=w, seed=777 + qlen + w, qlen_range=(qlen, qlen), h0_range=(19, 150))
            t = total // bn
            jobs = {k: np.tile(base[k], t) for k in ("qseq", "tseq", "qlen", "tlen", "h0")}
            jobs["qoff"] = np.concatenate([base["qoff"] + np.uint32(r * base["qseq"].size) for r in range(t)]).astype(np.uint32)
            jobs["toff"] = np.concatenate([base["toff"] + np.uint32(r * base["tseq"].size) for r in range(t)]).astype(np.uint32)
            n = jobs["qlen"].size
            dq = torch.from_numpy(jobs["qseq"]).cuda(); dt = torch.from_numpy(jobs["tseq"]).cuda()
            qp = torch.empty((dq.numel() + 7) // 8, dtype=torch.int32, device="cuda"); tp = torch.empty((dt.numel() + 7) // 8, dtype=torch.int32, device="cuda")
            ex.pack_device(dq.data_ptr(), dq.numel(), qp.data_ptr()); ex.pack_device(dt.data_ptr(), dt.numel(), tp.data_ptr())
            dev = {k: torch.from_numpy(jobs[k].view(np.int32)).cuda() for k in ("qoff", "toff", "qlen", "tlen", "h0")}
            res = torch.zeros(n * 6, dtype=torch.int32, device="cuda")
            ep = pkg.ext_params(w=w, zdrop=args.zdrop)

            def fn():
                ex.extend_device(ep, n, qp.data_ptr(), dev["qoff"].data_ptr(), dev["qlen"].data_ptr(), tp.data_ptr(), dev["toff"].data_ptr(),
                                 dev["tlen"].data_ptr(), dev["h0"].data_ptr(), res.data_ptr())
            ms = timed(st, fn, flush, args.reps)
            ex.wait()
            cells = int(pkg.lib().bwa_b200_extender_last_cells(ex.h))
            out["c4_extension"].append({"qlen": qlen, "w": w, "jobs": n, "ms": ms, "cells": cells, "GCUPS": cells / (ms / 1e3) / 1e9, "Mjobs_per_s": n / (ms / 1e3) / 1e6})
            print(out["c4_extension"][-1], file=sys.stderr, flush=True)
            del dq, dt, qp, tp, dev, res
    ex.destroy()
    if args.only_long or args.only_c4:
        print(json.dumps(out))
        return
    # ---- C5
    cache = os.environ.get("BWA_B200_CACHE", "/tmp/bwa_b200_bench"); os.makedirs(cache, exist_ok=True)
    prefix = os.path.join(cache, f"g{args.genome}_s{synth.GENOME_SEED}")
    g = synth.make_genome(args.genome, seed=synth.GENOME_SEED)
    if not (os.path.exists(prefix + ".sa") and os.path.exists(prefix + ".bwt")):
        pkg.build_index(g, prefix, sa_intv=16, also_stock_layout=True, n_threads=0)
    idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    reads, _, _ = synth.make_reads(g, args.reads, 250, seed=778)
    n, L = reads.shape
    packed, woff, rl = pkg.pack_codes(reads.reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(L)))
    d_packed = torch.from_numpy(packed.view(np.int32)).cuda(); d_woff = torch.from_numpy(woff.view(np.int64)).cuda(); d_rl = torch.from_numpy(rl.view(np.int32)).cuda()
    sd = pkg.Seeder(idx, n, packed.size)
    st = torch.cuda.ExternalStream(sd.stream)
    rows = []
    for name, par in (("pass1", pkg.seed_params(19, 500)), ("reseed", pkg.seed_params(19, 500, True))):
        fn = lambda: sd.seed_device(d_packed.data_ptr(), d_woff.data_ptr(), d_rl.data_ptr(), n, params=par)  # noqa: E731
        ms = timed(st, fn, flush, args.reps)
        tot = int(sd.device_result().n_seeds)
        rows.append({"mode": name, "reads": n, "read_len": L, "genome": args.genome, "ms": ms, "Mreads_per_s": n / (ms / 1e3) / 1e6, "seeds": tot})
        print(rows[-1], file=sys.stderr, flush=True)
    out["c5_seeding"] = rows
    print(json.dumps(out))


if __name__ == "__main__":
    main()
