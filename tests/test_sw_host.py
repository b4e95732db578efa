"""The ksw_align2 kernel's source (csrc/sw_core.cuh) built for the host, against the oracle and the reference's golden vectors: checks
the replay of the striped kernels (lane-local F, lazy-F rounds, the run list behind score2, the reversed second pass) bit for bit
without a GPU.  The GPU parity test (tests/test_gpu_sw.py) runs the same source on the device."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import oracle_py as O
from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "host_emul")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_sw_oracle import CASES, GOLD  # noqa: E402


@pytest.fixture(scope="module")
def emul(pkg):
    so = os.path.join(HERE, "libsw_host.so")
    srcs = [os.path.join(HERE, "sw_host.cpp"), os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc", "sw_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc"), srcs[0], "-o", so])
    L = C.CDLL(so)
    L.sw_host_run.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 8

    def run(jobs, ep):
        n = jobs["qlen"].size
        res = np.zeros(n, O.SW_RES_DT)
        L.sw_host_run(C.addressof(ep), n, *[jobs[k].ctypes.data for k in ("qseq", "qoff", "qlen", "tseq", "toff", "tlen", "xtra")], res.ctypes.data)
        return res
    return run


@pytest.mark.parametrize("name,jk,pk", CASES, ids=[c[0] for c in CASES])
def test_sw_kernel_source_on_host(pkg, oracle, emul, name, jk, pk):
    jobs = synth.make_sw_jobs(**jk)
    got = emul(jobs, pkg.ext_params(**pk))
    want = O.sw_align2_batch(jobs, O.make_params(**pk), n_threads=4)
    bad = [i for i in range(len(got)) if tuple(got[i]) != tuple(want[i])]
    assert not bad, (bad[:5], got[bad[:3]], want[bad[:3]])
    if name != "overflow_byte":
        gold = np.load(GOLD)[name]
        for k, f in enumerate(got.dtype.names):
            assert (got[f] == gold[:, k]).all(), f
