"""N > 1 host path on CPU: two gloo ranks shard a read batch, each runs its block (here through the
CPU oracle pipeline, standing in for the per-GPU pass), rank 0 gathers; result == single-process run."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, prefix, gpath, rpath, outpath):
    import importlib.util
    import torch.distributed as dist
    from oracle import oracle_py as O
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(ROOT, "bwa-mem_gpu_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = np.load(gpath)
    reads = np.load(rpath)
    n, L = reads.shape
    lo, hi = sh.shard_bounds(n, rank, world)
    oi = O.OracleIndex(prefix + ".bwt", prefix + ".sa")
    f = reads[lo:hi].reshape(-1).copy()
    off = (np.arange(hi - lo + 1) * L).astype(np.uint64)
    rec, _, _ = O.pipeline(oi, g, f, off, O.make_params(), 19, 500, n_threads=1)
    allrec = sh.gather_records(rec, n, dist, rank, world)
    if rank == 0:
        np.save(outpath, allrec)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sharding", os.path.join(ROOT, "bwa-mem_gpu_b200", "sharding.py"))
    sh = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sh)
    for n in (0, 1, 7, 1000, 1001):
        for w in (1, 2, 3, 8):
            b = [sh.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_two_rank_gloo_matches_single_process(pkg, oracle, small_index):
    import torch.multiprocessing as mp
    from tools import synth
    g, prefix = small_index
    reads, _, _ = synth.make_reads(g, 1001, 150, seed=21)
    tmp = tempfile.mkdtemp()
    gpath, rpath, outpath = os.path.join(tmp, "g.npy"), os.path.join(tmp, "r.npy"), os.path.join(tmp, "out.npy")
    np.save(gpath, g)
    np.save(rpath, reads)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, prefix, gpath, rpath, outpath), nprocs=2, join=True)
    got = np.load(outpath)
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    want, _, _ = oracle.pipeline(oi, g, reads.reshape(-1).copy(), (np.arange(1002) * 150).astype(np.uint64), oracle.make_params(), 19, 500, 2)
    assert got.tobytes() == want.tobytes()
    oi.close()
