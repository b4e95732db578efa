"""Builds an index with more than 2^32 rows with the host builder (64-bit suffix indexes, packed SA high bits) and checks it
against ground truth through the oracle: reads cut from known places of both strands must come back as one seed at that place.
Reverse-strand reads from the start of the genome sit beyond row position 2^32 of the text fwd + revcomp(fwd), so their
positions need the high bits (bwa_index/bwt.c:151-172).  CPU only; needs ~50 GB of host memory at the default size.

    python tools/check_wide_index.py [l_pac=2200000000] [n_reads=4000]

With WIDE_GPU=1 the same reads also go through the CUDA seeder on cuda:0 (64-bit row kernels, packed high bits in locate_kernel)
and must give the same seeds as the oracle and the ground truth.
"""
import importlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("bwa-mem_gpu_b200")
from oracle import oracle_py  # noqa: E402  (checker only)


def main():
    l_pac = int(sys.argv[1]) if len(sys.argv) > 1 else 2_200_000_000
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
    prefix = os.environ.get("WIDE_PREFIX", "/tmp/wide_idx")
    L = 100
    n = 2 * l_pac
    rng = np.random.default_rng(20261018)
    fwd = rng.integers(0, 4, l_pac, dtype=np.uint8)
    t = time.time()
    pkg.build_index(fwd, prefix, sa_intv=16, also_stock_layout=False, n_threads=0)
    t_build = time.time() - t
    # reads: even = forward strand anywhere; odd = reverse strand, from the part whose text position is >= 2^32 when there is one
    hi_span = n - (1 << 32) - L if n > (1 << 32) + L else 0
    pos = np.empty(n_reads, np.int64)
    reads = np.empty((n_reads, L), np.uint8)
    expect = np.empty(n_reads, np.uint64)
    for i in range(n_reads):
        if i & 1:
            p = int(rng.integers(0, hi_span if hi_span else l_pac - L))
            reads[i] = 3 - fwd[p:p + L][::-1]
            expect[i] = n - p - L
        else:
            p = int(rng.integers(0, l_pac - L))
            reads[i] = fwd[p:p + L]
            expect[i] = p
        pos[i] = p
    # a second set with substitutions (150 bp, 2 %): several SMEMs per read, checked GPU against oracle only
    L2 = 150
    noisy = np.empty((n_reads, L2), np.uint8)
    for i in range(n_reads):
        p = int(rng.integers(0, hi_span if (i & 1 and hi_span) else l_pac - L2))
        r = fwd[p:p + L2].copy()
        hit = rng.random(L2) < 0.02
        r[hit] = (r[hit] + rng.integers(1, 4, int(hit.sum()), dtype=np.uint8)) & 3
        noisy[i] = 3 - r[::-1] if i & 1 else r
    del fwd
    oi = oracle_py.OracleIndex(prefix + ".bwt", prefix + ".sa")
    off = (np.arange(n_reads + 1) * L).astype(np.uint64)
    res = oi.seed_batch(reads.reshape(-1).copy(), off, 19, 500)
    off2 = (np.arange(n_reads + 1) * L2).astype(np.uint64)
    res2 = oi.seed_batch(noisy.reshape(-1).copy(), off2, 19, 500)
    ok = bool((res["n_seeds"] == 1).all()) and res["total"] == n_reads
    if ok:
        ok = bool((res["rbeg"] == expect).all() and (res["qbeg"] == 0).all() and (res["qend"] == L).all())
    info = dict(l_pac=l_pac, rows=n, pack_size=int(oi.idx.pack_size), build_s=round(t_build, 1), reads=n_reads,
                reads_beyond_2_32=int((expect >= (1 << 32)).sum()), all_located_at_truth=ok,
                noisy_seeds=int(res2["total"]), noisy_seeds_beyond_2_32=int((res2["rbeg"] >= (1 << 32)).sum()))
    oi.close()
    print(json.dumps(info), flush=True)
    if os.environ.get("WIDE_GPU") == "1":
        gidx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", 0)
        packed, woff, rl = pkg.pack_codes(reads.reshape(-1).copy(), off)
        sd = pkg.Seeder(gidx, n_reads, packed.size)
        got = sd.seed_host(packed, woff, rl, 19, 500)
        gpu_ok = bool(got["total"] == res["total"] and (got["n_seeds"] == res["n_seeds"]).all() and (got["rbeg"] == res["rbeg"]).all()
                      and (got["qq"][:, 0] == res["qbeg"]).all() and (got["qq"][:, 1] == res["qend"]).all())
        packed2, woff2, rl2 = pkg.pack_codes(noisy.reshape(-1).copy(), off2)
        sd2 = pkg.Seeder(gidx, n_reads, packed2.size)
        got2 = sd2.seed_host(packed2, woff2, rl2, 19, 500)
        gpu_ok2 = bool(got2["total"] == res2["total"] and (got2["n_seeds"] == res2["n_seeds"]).all() and (got2["rbeg"] == res2["rbeg"]).all()
                       and (got2["qq"][:, 0] == res2["qbeg"]).all() and (got2["qq"][:, 1] == res2["qend"]).all()
                       and (got2["score"] == res2["score"]).all())
        sd2.destroy()
        info["gpu_equals_oracle"] = gpu_ok
        info["gpu_equals_oracle_noisy"] = gpu_ok2
        gpu_ok = gpu_ok and gpu_ok2
        info["gpu_hbm_bytes"] = int(gidx.info().hbm_bytes)
        ok = ok and gpu_ok
        sd.destroy()
        gidx.free()
        print(json.dumps(info), flush=True)
    for e in (".bwt", ".sa"):
        os.remove(prefix + e)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
