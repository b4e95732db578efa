"""GPU parity tests of the re-seeding passes (mem_collect_intv passes 2 and 3): the CUDA path through
the C ABI against the oracle and against golden vectors produced by the reference's own mem_collect_intv.  Bit-exact."""
import os

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gpu(pkg):
    assert pkg.lib().bwa_b200_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    return pkg


@pytest.fixture(scope="module")
def dev_index(gpu, small_index, oracle):
    g, prefix = small_index
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    yield g, idx, oi
    idx.free()
    oi.close()


def reseed_compare(gpu, oracle, idx, oi, flat, off, min_seed_len, max_occ, sf=1.5, sw=10, mmi=20):
    packed, woff, rl = gpu.pack_codes(flat, off)
    n = rl.size
    sd = gpu.Seeder(idx, max(n, 1), max(packed.size, 1))
    p = gpu.seed_params(min_seed_len, max_occ, True, sf, sw, mmi)
    rs = oracle.reseed(sf, sw, mmi)
    got = sd.seed_host(packed, woff, rl, params=p)
    sm = sd.smems(n, max(4096, int(flat.size) * 4))
    sd.destroy()
    wsm = oi.smem_batch(flat, off, min_seed_len, rs=rs)
    want = oi.seed_batch(flat, off, min_seed_len, max_occ, n_threads=4, rs=rs)
    assert (sm["n_smems"] == wsm["n_smems"]).all()
    for key in ("qbeg", "qend", "k", "s"):
        assert (sm[key] == wsm[key]).all(), key
    assert got["total"] == want["total"]
    assert (got["n_seeds"] == want["n_seeds"]).all() and (got["seed_off"] == want["seed_off"]).all()
    assert (got["qq"][:, 0] == want["qbeg"]).all() and (got["qq"][:, 1] == want["qend"]).all()
    assert (got["score"] == want["score"]).all() and (got["rbeg"] == want["rbeg"]).all()
    return sm, wsm


def test_reseed_matches_oracle(gpu, oracle, dev_index):
    g, idx, oi = dev_index
    reads, _, _ = synth.make_reads(g, 4000, 150, seed=51, n_rate=0.002)
    flat, off = reads.reshape(-1).copy(), (np.arange(4001) * 150).astype(np.uint64)
    sm, _ = reseed_compare(gpu, oracle, idx, oi, flat, off, 19, 500)
    p1 = oi.smem_batch(flat, off, 19)
    assert sm["n_smems"].sum() > 2 * p1["n_smems"].sum()              # the extra passes did add intervals
    reseed_compare(gpu, oracle, idx, oi, flat, off, 19, 7, sf=1.0, sw=3, mmi=0)       # pass 3 off
    reseed_compare(gpu, oracle, idx, oi, flat, off, 19, 20, sf=2.0, sw=50, mmi=5)


def test_reseed_ragged_edge_reads_and_row_overflow(gpu, oracle, dev_index):
    g, idx, oi = dev_index
    rng = np.random.default_rng(13)
    base, _, _ = synth.make_reads(g, 400, 250, seed=19, sub_rate=0.02, n_rate=0.003)
    rl = [base[i, :int(rng.integers(1, 251))] for i in range(400)]
    rl += [np.full(40, 4, np.uint8), np.zeros(5, np.uint8), np.zeros(300, np.uint8), np.array([1], np.uint8), g[1000:1019].copy(),
           g[5000:5600].copy(), synth.revcomp(g[7000:7400].copy()), g[-100:].copy(), g[:64].copy()]
    lens = np.array([len(r) for r in rl], np.uint64)
    off = np.zeros(len(rl) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    flat = np.concatenate(rl).astype(np.uint8)
    reseed_compare(gpu, oracle, idx, oi, flat, off, 19, 500)
    # short seeds + permissive thresholds: far more intervals per read than the initial row width (the widen-and-redo path)
    sm, _ = reseed_compare(gpu, oracle, idx, oi, flat, off, 8, 3, sf=1.0, sw=1000, mmi=1000)
    assert sm["n_smems"].max() > 64


def test_reseed_wide_rows_path(gpu, oracle, dev_index, monkeypatch):
    g, idx, oi = dev_index
    monkeypatch.setenv("BWA_B200_WIDE_ROWS", "1")
    reads, _, _ = synth.make_reads(g, 2000, 150, seed=61, n_rate=0.002)
    reseed_compare(gpu, oracle, idx, oi, reads.reshape(-1).copy(), (np.arange(2001) * 150).astype(np.uint64), 19, 500)


def test_reseed_golden_from_reference(gpu, tmp_path):
    """CUDA path against the interval lists and seeds of the reference's mem_collect_intv + bwt_sa (make_reseed_golden.py)"""
    gold = np.load(os.path.join(GOLD, "reseed_golden.npz"))
    g = synth.make_repeat_genome(int(gold["genome_len"]), seed=int(gold["genome_seed"]))
    prefix = str(tmp_path / "g")
    gpu.build_index(g, prefix, sa_intv=int(gold["sa_intv"]), n_threads=4)
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    flat, off = gold["reads"], gold["read_off"]
    packed, woff, rl = gpu.pack_codes(flat, off)
    for si, (sf, sw, mmi) in enumerate(gold["settings"]):
        sd = gpu.Seeder(idx, rl.size, packed.size)
        got = sd.seed_host(packed, woff, rl, params=gpu.seed_params(19, int(gold["max_occ"]), True, float(sf), int(sw), int(mmi)))
        sm = sd.smems(rl.size, int(flat.size) * 4)
        sd.destroy()
        want = gold[f"intv_{si}"]
        assert (sm["n_smems"] == gold[f"n_smems_{si}"]).all()
        assert (sm["qbeg"] == want[:, 0]).all() and (sm["qend"] == want[:, 1]).all()
        assert (sm["k"] == want[:, 2]).all() and (sm["s"] == want[:, 3]).all()
        if si == 0:
            assert (got["n_seeds"] == gold["n_seeds"]).all()
            assert (got["rbeg"] == gold["rbeg"]).all() and (got["score"] == gold["score"]).all()
    idx.free()


@pytest.mark.parametrize("K,sat", [(0, None), (5, None), (9, 100)])
def test_reseed_kmer_table_variants(gpu, oracle, dev_index, monkeypatch, K, sat):
    """re-seeding (back_kernel<RESEED> takes table steps too) under several k-mer table shapes"""
    g, idx, oi = dev_index
    if sat is not None:
        monkeypatch.setenv("BWA_B200_KMER_SAT", str(sat))
    idx.set_kmer_table(K)
    try:
        reads, _, _ = synth.make_reads(g, 2000, 150, seed=61 + K, n_rate=0.002)
        flat, off = reads.reshape(-1).copy(), (np.arange(2001) * 150).astype(np.uint64)
        reseed_compare(gpu, oracle, idx, oi, flat, off, 19, 500)
        reseed_compare(gpu, oracle, idx, oi, flat, off, 19, 20, sf=2.0, sw=50, mmi=5)
    finally:
        monkeypatch.delenv("BWA_B200_KMER_SAT", raising=False)
        idx.set_kmer_table(9)
