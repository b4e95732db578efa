#!/usr/bin/env python
"""bwa_b200_multi_align_compact from ONE process over every visible GPU (what the reference's single driver process would call):
the C2 batch times the number of devices, dealt in chunks; host buffers in -> host buffers out, reads/s; and the same call over one
device for the scaling ratio.  Identity of the two results is checked.
  python tools/multi_single_process.py [--reads-per-gpu N] [--chunk N]"""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from tools import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--reads-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=500_000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    pkg = ge.load_package(); pkg.build()
    n_dev = torch.cuda.device_count()
    cache = os.environ.get("BWA_B200_CACHE", "/tmp/bwa_b200_bench"); os.makedirs(cache, exist_ok=True)
    prefix = os.path.join(cache, f"g{args.genome}_s{synth.GENOME_SEED}")
    g = synth.make_genome(args.genome, seed=synth.GENOME_SEED)
    if not (os.path.exists(prefix + ".sa") and os.path.exists(prefix + ".bwt")):
        pkg.build_index(g, prefix, sa_intv=16, also_stock_layout=True, n_threads=0)
    idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(g)
    n = args.reads_per_gpu * n_dev
    reads, _, _ = synth.make_reads(g, n, 150, seed=synth.READS_SEED)
    L = reads.shape[1]
    p2, _, nl = pkg.pack2_codes(reads.reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(L)), with_lengths=False)
    pin = torch.empty(p2.nbytes, dtype=torch.uint8).pin_memory(); pin.numpy()[:] = p2.view(np.uint8)
    sp, cp, ep = pkg.seed_params(19, 500), pkg.chain_params(w=100), pkg.ext_params()
    out = {"devices": n_dev, "reads": n, "chunk": args.chunk, "runs": []}
    sig = None
    for devs in ([0], list(range(n_dev))):
        t0 = time.time()
        m = pkg.MultiAligner(idx, devs, 2, args.chunk, L)
        setup = time.time() - t0
        nn = args.reads_per_gpu * len(devs)
        res = m.align_compact(pin.data_ptr(), None, L, nn, None, 0, sp, cp, ep, copy=True)
        s = (res["n_regions"][:args.reads_per_gpu].tobytes(), res["regions"][:int(res["n_regions"][:args.reads_per_gpu].sum())].tobytes())
        if sig is None:
            sig = s
        same = s == sig
        m.align_compact(pin.data_ptr(), None, L, nn, None, 0, sp, cp, ep, copy=False, gather=False)
        t0 = time.perf_counter()
        for _ in range(args.reps):
            m.align_compact(pin.data_ptr(), None, L, nn, None, 0, sp, cp, ep, copy=False, gather=False)
        dt = (time.perf_counter() - t0) / args.reps
        out["runs"].append({"devices": devs, "reads": nn, "ms": dt * 1e3, "reads_per_s": nn / dt, "setup_s_incl_index_replicas": round(setup, 2),
                            "worker_chunks": m.worker_chunks(), "first_gpu_share_identical": same})
        print(out["runs"][-1], file=sys.stderr, flush=True)
        m.destroy()
        if n_dev == 1:
            break
    if len(out["runs"]) == 2:
        out["scaling_efficiency"] = out["runs"][1]["reads_per_s"] / (n_dev * out["runs"][0]["reads_per_s"])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
