"""CPU tests: oracle vs golden vectors (generated from the reference's own CPU functions, see
tests/golden/make_golden.py), host index builder vs golden hashes, oracle internal consistency."""
import hashlib
import os

import numpy as np
import pytest

from tools import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.mark.parametrize("wide,threads", [("0", 4), ("1", 4), ("1", 7), ("0", 1)])
def test_builder_matches_reference_hashes(pkg, tmp_path, monkeypatch, wide, threads):
    """index files byte-identical to the reference's `bwa index` output (hashes recorded by make_golden.py); wide = the builder's
    64-bit suffix-index instantiation (the one genomes with 2*l_pac >= 2^32 take), forced onto the same small texts; the passes are
    cut by thread, so odd thread counts move the cuts"""
    monkeypatch.setenv("BWA_B200_BUILD_WIDE", wide)
    gold = np.load(os.path.join(GOLD, "index_hashes.npz"), allow_pickle=True)
    for n, rep, seed, intv, h_bwt, h_sa, h_128 in gold["rows"]:
        g = synth.make_genome(int(n), seed=int(seed), repeats=bool(int(rep)))
        prefix = str(tmp_path / f"g{n}")
        pkg.build_index(g, prefix, sa_intv=int(intv), also_stock_layout=True, n_threads=threads)
        assert sha(prefix + ".bwt") == h_bwt
        assert sha(prefix + ".sa") == h_sa
        assert sha(prefix + ".bwt128") == h_128


def test_oracle_smems_match_reference_golden(oracle, pkg, tmp_path):
    gold = np.load(os.path.join(GOLD, "seed_golden.npz"))
    g = synth.make_repeat_genome(int(gold["genome_len"]), seed=int(gold["genome_seed"]))
    prefix = str(tmp_path / "g")
    pkg.build_index(g, prefix, sa_intv=int(gold["sa_intv"]), n_threads=4)
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    reads = gold["reads"]
    n, L = reads.shape
    off = (np.arange(n + 1) * L).astype(np.uint64)
    res = oi.smem_batch(reads.reshape(-1).copy(), off, 19)
    assert (res["n_smems"] == gold["n_smems"]).all()
    for key in ("qbeg", "qend", "k", "s"):
        assert (res[key] == gold[key]).all(), key
    sb = oi.seed_batch(reads.reshape(-1).copy(), off, 19, int(gold["max_occ"]), n_threads=2)
    assert (sb["n_seeds"] == gold["n_seeds"]).all()
    assert (sb["rbeg"] == gold["rbeg"]).all()
    assert (sb["score"] == gold["score"]).all()
    oi.close()


@pytest.mark.parametrize("name", ["ksw_band", "ksw_noband", "ksw_narrow", "ksw_asym"])
def test_oracle_ksw_matches_reference_golden(oracle, name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    jobs = {k: gold[k] for k in ("qseq", "tseq", "qoff", "toff", "qlen", "tlen", "h0")}
    kw = {k: int(v) for k, v in zip(gold["param_names"], gold["param_values"])}
    res, cnt = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=2)
    assert (res == gold["res6"]).all()
    assert cnt["cells"] > 0


def test_oracle_edge_cases(oracle):
    """tlen == 0, 1-base queries, all-N query, h0 == 1"""
    p = oracle.make_params()
    q = np.array([0, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 4, 2, 4, 4, 4, 4, 4, 4, 4], np.uint8)
    t = np.array([0, 1, 2, 3, 0, 1, 2, 3, 0, 0, 0, 0, 0, 0, 0, 0], np.uint8)
    jobs = dict(qseq=q, tseq=t, qoff=np.array([0, 8, 16], np.uint32), toff=np.array([0, 0, 8], np.uint32),
                qlen=np.array([1, 8, 1], np.uint32), tlen=np.array([0, 8, 8], np.uint32), h0=np.array([5, 1, 30], np.uint32))
    res, _ = oracle.ksw_batch(jobs, p, n_threads=1)
    assert list(res[0]) == [5, 0, 0, 0, -1, 0]          # no target rows: score = h0
    assert res[1][0] == 1                                # all-N query never beats h0
    assert res[2][0] == 30


@pytest.mark.parametrize("wide", ["0", "1"])
def test_builder_and_oracle_locate_reads_at_their_true_place(pkg, oracle, tmp_path, wide):
    """ground truth instead of a hash: error-free reads cut from known places of both strands come back as one seed at that place
    (tools/check_wide_index.py at a small size; the same tool checks an index beyond 2^32 rows)"""
    import json
    import subprocess
    import sys
    env = dict(os.environ, BWA_B200_BUILD_WIDE=wide, WIDE_PREFIX=str(tmp_path / "w"))
    env.pop("WIDE_GPU", None)
    out = subprocess.run([sys.executable, os.path.join(pkg.ROOT, "tools", "check_wide_index.py"), "2500000", "300"], env=env,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    info = json.loads(out.stdout.strip().splitlines()[-1])
    assert info["all_located_at_truth"] and info["rows"] == 5000000 and info["noisy_seeds"] > 300
