"""GPU parity of ksw_align2 (mate rescue / mem_seed_sw): the CUDA path through the C ABI (bwa_b200_sw_align2_host)
against the oracle and against golden vectors produced by the reference's own SSE2 functions.  Bit-exact, every field of kswr_t."""
import os
import sys

import numpy as np
import pytest

from oracle import oracle_py as O
from tools import synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_sw_oracle import CASES, GOLD  # noqa: E402


@pytest.fixture(scope="module")
def sw(pkg):
    assert pkg.lib().bwa_b200_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    h = pkg.LocalAligner(0)
    yield h
    h.destroy()


@pytest.mark.parametrize("name,jk,pk", CASES, ids=[c[0] for c in CASES])
def test_sw_align2_matches_oracle_and_reference_golden(pkg, oracle, sw, name, jk, pk):
    jobs = synth.make_sw_jobs(**jk)
    got = sw.align2_host(jobs, pkg.ext_params(**pk))
    want = O.sw_align2_batch(jobs, O.make_params(**pk), n_threads=4)
    bad = [i for i in range(len(got)) if tuple(got[i]) != tuple(want[i])]
    assert not bad, (bad[:5], got[bad[:3]], want[bad[:3]])
    if name != "overflow_byte":
        gold = np.load(GOLD)[name]
        for k, f in enumerate(got.dtype.names):
            assert (got[f] == gold[:, k]).all(), f


def test_sw_align2_larger_batch_and_argument_checks(pkg, oracle, sw):
    jobs = synth.make_sw_jobs(20_000, qlen_range=(100, 150), tlen_range=(300, 600), seed=77)
    l0 = sw.launches
    got = sw.align2_host(jobs, pkg.ext_params())
    assert sw.launches in (l0 + 1, l0 + 2)          # the register-resident byte kernel, plus the replay kernel when some job is outside its class
    want = O.sw_align2_batch(jobs, O.make_params())
    assert got.tobytes() == want.tobytes()
    assert (got["qb"] >= 0).mean() > 0.5 and (got["score2"] > 0).sum() > 0      # the workload reaches the KSW_XSTART pass and the second-best score
    bad = dict(jobs)
    bad["qlen"] = jobs["qlen"].copy(); bad["qlen"][5] = 0
    with pytest.raises(pkg.B200Error):
        sw.align2_host(bad, pkg.ext_params())
    assert sw.align2_host({k: v[:0] for k, v in jobs.items()}, pkg.ext_params()).size == 0


@pytest.mark.parametrize("pk", [dict(), dict(a=2, b=3, o_del=4, e_del=2, o_ins=7, e_ins=1), dict(o_del=2, e_del=1, o_ins=2, e_ins=1), dict(a=1, b=9, o_del=11, e_del=3, o_ins=9, e_ins=2)])
def test_sw_byte_kernel_in_registers_fuzz(pkg, oracle, sw, pk):
    """sw_stripe_kernel (one job per eight lanes, the striped vectors in registers): every query length of its class (1 .. 256 bases: 1 .. 16
    vectors per row), targets from one base to beyond a thousand, byte saturation, jobs of mixed shapes sharing a warp"""
    X = O.SW_XSUBO | O.SW_XSTART | O.SW_XBYTE | 19
    n_ov = 0
    for seed, jk in ((41, dict(qlen_range=(1, 256), tlen_range=(1, 900))), (42, dict(qlen_range=(100, 160), tlen_range=(300, 1200), sub_rate=0.01, indel_rate=0.002)),
                     (43, dict(qlen_range=(1, 40), tlen_range=(1, 60), none_frac=0.4)), (44, dict(qlen_range=(200, 256), tlen_range=(200, 700), sub_rate=0.0, indel_rate=0.0, none_frac=0.0))):
        jobs = synth.make_sw_jobs(3000, seed=seed, xtra=X, **jk)
        jobs["xtra"] = jobs["xtra"].copy()
        jobs["xtra"][::7] = O.SW_XBYTE | O.SW_XSTART            # mem_seed_sw-like flags in byte mode
        jobs["xtra"][3::11] = O.SW_XBYTE                         # the score only
        got = sw.align2_host(jobs, pkg.ext_params(**pk))
        want = O.sw_align2_batch(jobs, O.make_params(**pk), n_threads=4)
        ov = want["score"] == 255                                # byte overflow: only the first pass is defined (see test_sw_oracle)
        assert (got["score"] == want["score"]).all() and (got["te"] == want["te"]).all()
        g, w = got[~ov], want[~ov]
        bad = [i for i in range(len(g)) if tuple(g[i]) != tuple(w[i])]
        assert not bad, (seed, bad[:5], g[bad[:3]], w[bad[:3]], jobs["qlen"][~ov][bad[:3]], jobs["tlen"][~ov][bad[:3]])
        n_ov = n_ov + int(ov.sum())
    if pk.get("a", 1) == 2:
        assert n_ov > 20                                         # the byte kernel's saturation path is exercised
