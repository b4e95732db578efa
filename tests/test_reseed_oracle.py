"""CPU tests of the re-seeding oracle (passes 2 and 3 of mem_collect_intv): the restatement in
oracle/fmd_oracle.c against golden vectors produced by the reference's own mem_collect_intv
(tests/golden/make_reseed_golden.py) and, when oracle/_ref is present, against that function live on other inputs."""
import os

import numpy as np
import pytest

from tools import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _index(pkg, tmp_path, g, intv):
    prefix = str(tmp_path / "g")
    pkg.build_index(g, prefix, sa_intv=intv, also_stock_layout=True, n_threads=4)
    return prefix


def test_reseed_oracle_matches_reference_golden(oracle, pkg, tmp_path):
    gold = np.load(os.path.join(GOLD, "reseed_golden.npz"))
    g = synth.make_repeat_genome(int(gold["genome_len"]), seed=int(gold["genome_seed"]))
    prefix = _index(pkg, tmp_path, g, int(gold["sa_intv"]))
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    flat, off = gold["reads"], gold["read_off"]
    for si, (sf, sw, mmi) in enumerate(gold["settings"]):
        rs = oracle.reseed(float(sf), int(sw), int(mmi))
        res = oi.smem_batch(flat, off, 19, rs=rs)
        want = gold[f"intv_{si}"]
        assert (res["n_smems"] == gold[f"n_smems_{si}"]).all()
        assert (res["qbeg"] == want[:, 0]).all() and (res["qend"] == want[:, 1]).all()
        assert (res["k"] == want[:, 2]).all() and (res["s"] == want[:, 3]).all()
    sb = oi.seed_batch(flat, off, 19, int(gold["max_occ"]), n_threads=2, rs=oracle.reseed())
    assert (sb["n_seeds"] == gold["n_seeds"]).all()
    assert (sb["rbeg"] == gold["rbeg"]).all() and (sb["score"] == gold["score"]).all()
    # more intervals than pass 1 alone, and pass 1 is a subset
    p1 = oi.smem_batch(flat, off, 19)
    assert p1["n_smems"].sum() < gold["n_smems_0"].sum()
    oi.close()


def test_reseed_oracle_matches_reference_live(oracle, pkg, tmp_path):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    g = synth.make_genome(300_000, seed=77, repeats=True)
    prefix = _index(pkg, tmp_path, g, 8)
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    h = oracle.ref_lib().ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
    assert h
    r1, _, _ = synth.make_reads(g, 1500, 150, seed=31)
    r2, _, _ = synth.make_reads(g, 500, 101, seed=32, n_rate=0.005, sub_rate=0.02)
    for reads in (r1, r2):
        n, L = reads.shape
        flat = reads.reshape(-1).copy()
        off = (np.arange(n + 1) * L).astype(np.uint64)
        for sf, sw, mmi in ((1.5, 10, 20), (1.2, 2, 100)):
            n_smems, iv = oracle.ref_collect_batch(h, flat, off, 19, sf, sw, mmi)
            res = oi.smem_batch(flat, off, 19, rs=oracle.reseed(sf, sw, mmi))
            assert (res["n_smems"] == n_smems).all()
            assert (res["qbeg"] == iv[:, 3]).all() and (res["qend"] == iv[:, 4]).all()
            assert (res["k"] == iv[:, 0]).all() and (res["s"] == iv[:, 2]).all()
    oracle.ref_lib().ref_free(h)
    oi.close()
