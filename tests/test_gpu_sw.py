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
    assert sw.launches == l0 + 1
    want = O.sw_align2_batch(jobs, O.make_params())
    assert got.tobytes() == want.tobytes()
    assert (got["qb"] >= 0).mean() > 0.5 and (got["score2"] > 0).sum() > 0      # the workload reaches the KSW_XSTART pass and the second-best score
    bad = dict(jobs)
    bad["qlen"] = jobs["qlen"].copy(); bad["qlen"][5] = 0
    with pytest.raises(pkg.B200Error):
        sw.align2_host(bad, pkg.ext_params())
    assert sw.align2_host({k: v[:0] for k, v in jobs.items()}, pkg.ext_params()).size == 0
