"""GPU parity of the region-finishing stage (bwa_b200_finish_regions_host: mem_sort_dedup_patch with mem_patch_reg, is_alt,
mem_mark_primary_se, mapq) against the reference fork's golden vectors and the oracle on fresh cases."""
import importlib.util
import os

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gpu(pkg):
    if pkg.lib().bwa_b200_device_count() < 1:
        pytest.fail("no CUDA device: the product path has no CPU fallback")
    return pkg


def _maker():
    spec = importlib.util.spec_from_file_location("mkreg", os.path.join(GOLD, "make_region_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def _index(gpu, g, tmp_path):
    prefix = str(tmp_path / "g")
    gpu.build_index(g, prefix, sa_intv=16, n_threads=4)
    idx = gpu.Index.load(prefix + ".bwt", None, 0)
    idx.attach_ref(g)
    return idx


def _opt(gpu, RP, kw):
    o = RP.default_opt(**kw)
    return o, gpu.region_opt(**{k: getattr(o, k) for k, _ in o._fields_})


def test_finish_regions_matches_fork_golden(gpu, oracle, tmp_path):
    from oracle import chain_py as CP, region_py as RP
    mk = _maker()
    gold = np.load(os.path.join(GOLD, "region_golden.npz"))
    ctg = CP.Contigs(tuple(int(x) for x in gold["contigs"]), alt=tuple(int(x) for x in gold["alt"]))
    g = synth.make_genome(ctg.l_pac, seed=int(gold["genome_seed"]))
    idx = _index(gpu, g, tmp_path)
    reads, regs_in, in_off = gold["reads"], gold["regs_in"], gold["in_off"]
    n, L = reads.shape
    packed, woff, rl = gpu.pack_codes(reads.reshape(-1).copy(), (np.arange(n + 1) * L).astype(np.uint64))
    assert gpu.ALNREG_DTYPE == RP.REGION_DT
    for oi, kw in enumerate(mk.OPTS):
        _, opt = _opt(gpu, RP, kw)
        got, n_pri = gpu.finish_regions(idx, packed, woff, rl, regs_in, in_off, opt, ctg_alt=ctg.alt, first_read_id=0)
        want, woff_, wpri = gold[f"out_{oi}"], gold[f"out_off_{oi}"], gold[f"n_pri_{oi}"]
        assert (n_pri == wpri).all()
        for i in range(n):
            assert RP.equal(got[i], want[woff_[i]:woff_[i + 1]]), (oi, i)
    idx.free()


def test_finish_regions_matches_oracle_fresh_cases_and_edges(gpu, oracle, tmp_path):
    from oracle import region_py as RP
    mk = _maker()
    ctg, g, reads, cases = mk.make_inputs(n_reads=3000, seed=4711)
    cases[0] = np.zeros(0, RP.REGION_DT)                               # a read without regions
    cases[1] = cases[1][:1] if len(cases[1]) else cases[1]             # and one with a single region
    idx = _index(gpu, g, tmp_path)
    n, L = reads.shape
    packed, woff, rl = gpu.pack_codes(reads.reshape(-1).copy(), (np.arange(n + 1) * L).astype(np.uint64))
    off = np.concatenate([[0], np.cumsum([len(c) for c in cases])]).astype(np.uint64)
    flat = np.concatenate(cases)
    for kw in mk.OPTS + (dict(w=5, mask_level=0.9), ):
        o, opt = _opt(gpu, RP, kw)
        got, n_pri = gpu.finish_regions(idx, packed, woff, rl, flat, off, opt, ctg_alt=ctg.alt, first_read_id=123456)
        merged = 0
        for i in range(n):
            b, pb = RP.oracle_finish(o, ctg, g, reads[i], cases[i], 123456 + i)
            assert int(n_pri[i]) == pb and RP.equal(got[i], b), (kw, i)
            merged += int((b["n_comp"] > 1).sum())
        assert merged > 20
    # nothing to do / no reference attached
    got, n_pri = gpu.finish_regions(idx, packed[:0], np.zeros(1, np.uint64), rl[:0], flat[:0], np.zeros(1, np.uint64), opt)
    assert got == [] and n_pri.size == 0
    idx.free()
