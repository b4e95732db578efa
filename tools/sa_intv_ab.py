#!/usr/bin/env python
"""SA sampling interval A/B ("denser SA sampling, since HBM is not scarce"): the same reads seeded and located against
indexes of the same genome built with sa_intv 16 / 8 / 4 -- locate_kernel time, LF steps saved, HBM spent, seeds identical.
  python tools/sa_intv_ab.py [--genome BASES] [--reads N]"""
import argparse, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as ge
from tools import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome", type=int, default=100_000_000)
    ap.add_argument("--reads", type=int, default=1_000_000)
    args = ap.parse_args()
    pkg = ge.load_package(); pkg.build()
    torch.cuda.set_device(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    g = synth.make_genome(args.genome, seed=synth.GENOME_SEED)
    reads, _, _ = synth.make_reads(g, args.reads, 150, seed=synth.READS_SEED)
    n, L = reads.shape
    packed, woff, rl = pkg.pack_codes(reads.reshape(-1), (np.arange(n + 1, dtype=np.uint64) * np.uint64(L)))
    d_packed = torch.from_numpy(packed.view(np.int32)).cuda(); d_woff = torch.from_numpy(woff.view(np.int64)).cuda(); d_rl = torch.from_numpy(rl.view(np.int32)).cuda()
    cache = os.environ.get("BWA_B200_CACHE", "/tmp/bwa_b200_bench"); os.makedirs(cache, exist_ok=True)
    rows, ref = [], None
    for intv in (16, 8, 4):
        prefix = os.path.join(cache, f"g{args.genome}_s{synth.GENOME_SEED}" + ("" if intv == 16 else f"_sa{intv}"))
        if not (os.path.exists(prefix + ".sa") and os.path.exists(prefix + ".bwt")):
            pkg.build_index(g, prefix, sa_intv=intv, also_stock_layout=False, n_threads=0)
        idx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", 0)
        idx.attach_ref(g)
        al = pkg.Aligner(idx, n, int(d_packed.numel()))
        sp, cp, ep = pkg.seed_params(19, 500), pkg.chain_params(w=100), pkg.ext_params()
        st = torch.cuda.ExternalStream(al.stream)
        for _ in range(3):
            al.align_device(d_packed.data_ptr(), d_woff.data_ptr(), d_rl.data_ptr(), n, L, sp, cp, ep)
        al.profile(2)
        kt = {}
        for _ in range(5):
            with torch.cuda.stream(st):
                flush.zero_()
            al.align_device(d_packed.data_ptr(), d_woff.data_ptr(), d_rl.data_ptr(), n, L, sp, cp, ep)
            for name, x in al.kernel_times():
                kt.setdefault(name, []).append(x)
        al.profile(False)
        host = al.align_host_view(packed.ctypes.data, woff.ctypes.data, rl.ctypes.data, n, sp, cp, ep, copy=True)
        sig = (host["n_regions"].tobytes(), host["regions"]["rb"].tobytes(), host["regions"]["score"].tobytes())
        if ref is None:
            ref = sig
        info = idx.info()
        rows.append({"sa_intv": intv, "locate_kernel_ms": float(np.mean(kt["locate_kernel"])), "step_ms": float(sum(np.mean(v) for v in kt.values())),
                     "index_hbm_bytes": int(info.hbm_bytes), "regions_identical_to_intv16": sig == ref})
        print(rows[-1], file=sys.stderr, flush=True)
        al.destroy(); idx.free()
    print(json.dumps({"genome": args.genome, "reads": n, "rows": rows}))


if __name__ == "__main__":
    main()
