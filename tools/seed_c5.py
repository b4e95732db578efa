"""BASELINE config 5 at full index size on one GPU: seeding only (pass 1 + locate, min_seed_len 19) of 250 bp reads (1 % substitutions)
against a synthetic human-sized genome -- index built on the box by bwa_b200_build_index (64-bit suffix indexes, packed SA high bits),
64-bit row kernels.  Timed host to host through bwa_b200_seed_host (H2D of the packed reads and D2H of the seeds included);
a subset is checked against the oracle.

    python tools/seed_c5.py [l_pac=3100000000] [n_reads=1000000] [read_len=250]
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
    l_pac = int(sys.argv[1]) if len(sys.argv) > 1 else 3_100_000_000
    n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    L = int(sys.argv[3]) if len(sys.argv) > 3 else 250
    n_check = min(n_reads, int(os.environ.get("C5_CHECK", "20000")))
    prefix = os.environ.get("WIDE_PREFIX", "/tmp/c5_idx")
    rng = np.random.default_rng(778)
    t = time.time()
    fwd = rng.integers(0, 4, l_pac, dtype=np.uint8)
    t_gen = time.time() - t
    t = time.time()
    pkg.build_index(fwd, prefix, sa_intv=16, also_stock_layout=False, n_threads=0)
    t_build = time.time() - t
    t = time.time()
    p = rng.integers(0, l_pac - L, n_reads)
    reads = fwd[p[:, None] + np.arange(L)[None, :]]
    del fwd
    hit = rng.random(reads.shape) < 0.01
    reads[hit] = (reads[hit] + rng.integers(1, 4, int(hit.sum()), dtype=np.uint8)) & 3
    rc = np.arange(n_reads) & 1 == 1
    reads[rc] = 3 - reads[rc][:, ::-1]
    flat = np.ascontiguousarray(reads).reshape(-1)
    off = (np.arange(n_reads + 1, dtype=np.uint64) * np.uint64(L))
    packed, woff, rl = pkg.pack_codes(flat, off)
    t_reads = time.time() - t
    info = dict(l_pac=l_pac, rows=2 * l_pac, n_reads=n_reads, read_len=L, genome_s=round(t_gen, 1), build_s=round(t_build, 1), reads_s=round(t_reads, 1))
    print(json.dumps(info), flush=True)
    if os.environ.get("WIDE_GPU") == "1":
        t = time.time()
        gidx = pkg.Index.load(prefix + ".bwt", prefix + ".sa", 0)
        info["index_load_s"] = round(time.time() - t, 1)
        info["index_hbm_bytes"] = int(gidx.info().hbm_bytes)
        sd = pkg.Seeder(gidx, n_reads, packed.size)
        times = []
        for _ in range(4):
            t = time.time()
            got = sd.seed_host(packed, woff, rl, 19, 500)
            times.append(time.time() - t)
        info["seed_host_s"] = [round(x, 4) for x in times]
        info["seeding_Mreads_per_s_host_to_host"] = round(n_reads / min(times[1:]) / 1e6, 2)
        info["seeds"] = int(got["total"])
        info["seeds_beyond_2_32"] = int((got["rbeg"] >= (1 << 32)).sum())
        info["launches"] = int(sd.launches)
        print(json.dumps(info), flush=True)
    oi = oracle_py.OracleIndex(prefix + ".bwt", prefix + ".sa")
    res = oi.seed_batch(flat[:n_check * L].copy(), off[:n_check + 1].copy(), 19, 500)
    oi.close()
    info["oracle_checked_reads"] = n_check
    info["oracle_seeds"] = int(res["total"])
    ok = True
    if os.environ.get("WIDE_GPU") == "1":
        m = int(res["total"])
        ok = bool((got["n_seeds"][:n_check] == res["n_seeds"]).all() and (got["rbeg"][:m] == res["rbeg"]).all()
                  and (got["qq"][:m, 0] == res["qbeg"]).all() and (got["qq"][:m, 1] == res["qend"]).all() and (got["score"][:m] == res["score"]).all())
        info["gpu_equals_oracle_on_subset"] = ok
        sd.destroy()
        gidx.free()
    print(json.dumps(info), flush=True)
    for e in (".bwt", ".sa"):
        os.remove(prefix + e)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
