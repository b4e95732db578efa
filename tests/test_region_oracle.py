"""CPU tests of the region-finishing restatement (oracle/region_oracle.c: mem_sort_dedup_patch with mem_patch_reg, is_alt,
mem_mark_primary_se, mem_approx_mapq_se) against golden vectors from the reference fork's own functions
(tests/golden/make_region_golden.py) and, when oracle/_ref is present, against those functions live on fresh cases."""
import importlib.util
import os

import numpy as np
import pytest

from tools import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _maker():
    spec = importlib.util.spec_from_file_location("mkreg", os.path.join(GOLD, "make_region_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def test_region_oracle_matches_fork_golden(oracle):
    from oracle import chain_py as CP, region_py as RP
    mk = _maker()
    gold = np.load(os.path.join(GOLD, "region_golden.npz"))
    ctg = CP.Contigs(tuple(int(x) for x in gold["contigs"]), alt=tuple(int(x) for x in gold["alt"]))
    g = synth.make_genome(ctg.l_pac, seed=int(gold["genome_seed"]))
    reads, regs_in, in_off = gold["reads"], gold["regs_in"], gold["in_off"]
    assert regs_in.dtype == RP.REGION_DT
    merged = secondary = dropped = 0
    for oi, kw in enumerate(mk.OPTS):
        opt = RP.default_opt(**kw)
        want, woff, wpri = gold[f"out_{oi}"], gold[f"out_off_{oi}"], gold[f"n_pri_{oi}"]
        for i in range(len(reads)):
            a, n_pri = RP.oracle_finish(opt, ctg, g, reads[i], regs_in[in_off[i]:in_off[i + 1]], i)
            w = want[woff[i]:woff[i + 1]]
            assert n_pri == wpri[i], (oi, i)
            assert RP.equal(a, w), (oi, i)
            merged += int((a["n_comp"] > 1).sum()); secondary += int((a["secondary"] >= 0).sum())
            dropped += int(in_off[i + 1] - in_off[i]) - len(a)
    assert merged > 50 and secondary > 1000 and dropped > 500          # the cases did reach the branches they were built for


def test_region_oracle_edge_cases(oracle):
    from oracle import chain_py as CP, region_py as RP
    ctg = CP.Contigs((5000,))
    g = synth.make_genome(5000, seed=3)
    opt = RP.default_opt()
    q = g[100:250].copy()
    a, n_pri = RP.oracle_finish(opt, ctg, g, q, np.zeros(0, RP.REGION_DT), 0)
    assert len(a) == 0 and n_pri == 0
    one = np.zeros(1, RP.REGION_DT)
    one["rb"], one["re"], one["qb"], one["qe"], one["score"], one["truesc"], one["seedcov"], one["secondary"] = 100, 250, 0, 150, 150, 150, 150, -1
    a, n_pri = RP.oracle_finish(opt, ctg, g, q, one, 7)
    assert len(a) == 1 and n_pri == 1 and int(a["secondary"][0]) == -1 and int(a["mapq"][0]) == 60      # a unique perfect hit
    low = one.copy(); low["score"] = 19                                                                   # sub defaults to min_seed_len * a
    a, _ = RP.oracle_finish(opt, ctg, g, q, low, 7)
    assert int(a["mapq"][0]) == 0
    two = np.concatenate([one, one])                                                                      # identical hits collapse
    a, _ = RP.oracle_finish(opt, ctg, g, q, two, 7)
    assert len(a) == 1


def test_region_oracle_matches_fork_live(oracle):
    from oracle import chain_py as CP, region_py as RP
    if not CP.have_fork():
        pytest.skip("oracle/_ref/libforkmem.so not built (no /root/reference here)")
    mk = _maker()
    ctg, g, reads, cases = mk.make_inputs(n_reads=700, seed=2026)
    pac = CP.make_pac(g)
    for kw in mk.OPTS + (dict(w=5, mask_level=0.9), ):
        opt = RP.default_opt(**kw)
        for i, regs in enumerate(cases):
            a, pa = RP.oracle_finish(opt, ctg, g, reads[i], regs, 1000 + i)
            b, pb = RP.fork_finish(opt, ctg, pac, reads[i], regs, 1000 + i)
            assert pa == pb and RP.equal(a, b), (kw, i)
