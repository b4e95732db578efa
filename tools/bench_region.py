"""First timing of the region-finishing stage (bwa_b200_finish_regions_host, host to host incl. its allocations): the golden
region sets tiled to a batch.    python tools/bench_region.py [tiles=200]"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bwa-mem_gpu_b200")
from tools import synth  # noqa: E402

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 200
gold = np.load(os.path.join(ROOT, "tests", "golden", "region_golden.npz"))
lens = tuple(int(x) for x in gold["contigs"])
g = synth.make_genome(sum(lens), seed=int(gold["genome_seed"]))
pkg.build_index(g, "/tmp/breg", sa_intv=16, n_threads=4)
idx = pkg.Index.load("/tmp/breg.bwt", None, 0)
idx.attach_ref(g)
reads = np.tile(gold["reads"], (tiles, 1))
n, L = reads.shape
per = np.diff(gold["in_off"])
off = np.concatenate([[0], np.cumsum(np.tile(per, tiles))]).astype(np.uint64)
regs = np.tile(gold["regs_in"], tiles)
packed, woff, rl = pkg.pack_codes(reads.reshape(-1).copy(), (np.arange(n + 1) * L).astype(np.uint64))
opt = pkg.region_opt()
ts = []
for _ in range(4):
    t = time.time()
    got, n_pri = pkg.finish_regions(idx, packed, woff, rl, regs, off, opt, ctg_alt=gold["alt"].astype(np.int32))
    ts.append(time.time() - t)
print(json.dumps(dict(reads=n, regions=int(regs.size), seconds=[round(x, 4) for x in ts],
                      Mreads_per_s=round(n / min(ts[1:]) / 1e6, 2), Mregions_per_s=round(regs.size / min(ts[1:]) / 1e6, 2),
                      note="includes cudaMalloc/cudaFree, H2D/D2H and the Python-side slicing of the result")))
