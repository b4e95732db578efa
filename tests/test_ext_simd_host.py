"""The two-row s16x2 extension kernel's source (ext_simd_core.cuh), built for the host with the DPX
intrinsics emulated, against the oracle.  Runs on the CPU box: it checks the kernel's row-window
logic (speculative second row, completion / roll-back) bit for bit without a GPU.  The GPU parity
tests run the same source through the real instructions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "host_emul")


@pytest.fixture(scope="module")
def emul(pkg):
    so = os.path.join(HERE, "libextsimd_host.so")
    srcs = [os.path.join(HERE, "ext_simd_host.cpp"), os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc", "ext_simd_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc"), srcs[0], "-o", so])
    L = C.CDLL(so)
    L.ext_simd_host_run.restype = C.c_longlong
    L.ext_simd_host_run.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 9

    def run(jobs, ep):
        n = jobs["qlen"].size
        res = np.zeros((n, 6), np.int32)
        skipped = np.zeros(n, np.uint8)
        cells = L.ext_simd_host_run(C.addressof(ep), n, jobs["qseq"].ctypes.data, jobs["qoff"].ctypes.data, jobs["qlen"].ctypes.data,
                                    jobs["tseq"].ctypes.data, jobs["toff"].ctypes.data, jobs["tlen"].ctypes.data, jobs["h0"].ctypes.data,
                                    res.ctypes.data, skipped.ctypes.data)
        return res, skipped.astype(bool), cells
    return run


KW = [dict(w=100, zdrop=100), dict(w=16, zdrop=100), dict(w=8, zdrop=0), dict(w=50, zdrop=30), dict(w=300, zdrop=0, use_band=0),
      dict(w=33, zdrop=100, o_del=3, e_del=1, o_ins=5, e_ins=2, a=3, b=2), dict(w=2, zdrop=10), dict(w=1, zdrop=0, end_bonus=0)]
SETS = [(61, dict(qlen_range=(1, 260), h0_range=(1, 250))),
        (62, dict(qlen_range=(1, 120), sub_rate=0.25, indel_rate=0.08, n_job_frac=0.3, h0_range=(1, 40))),
        (63, dict(qlen_range=(100, 500), h0_range=(100, 400), sub_rate=0.02, indel_rate=0.02)),
        (64, dict(qlen_range=(1, 40), h0_range=(1, 300), sub_rate=0.5, indel_rate=0.2)),
        (65, dict(qlen_range=(1, 12), h0_range=(1, 12), sub_rate=0.4, indel_rate=0.3, n_job_frac=0.5))]


@pytest.mark.parametrize("kw", KW)
def test_two_row_source_matches_oracle(pkg, oracle, emul, kw):
    for seed, extra in SETS:
        jobs = synth.make_ext_jobs(3000, w=kw["w"], seed=seed, **extra)
        want, _ = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        res, skipped, cells = emul(jobs, pkg.ext_params(**kw))
        assert cells >= 0
        ok = ~skipped
        assert ok.sum() > 100
        bad = np.nonzero((res[ok] != want[ok]).any(axis=1))[0]
        assert bad.size == 0, (seed, np.nonzero(ok)[0][bad[:3]], res[ok][bad[:3]], want[ok][bad[:3]])
        # evaluated-cell count of the eligible jobs equals the oracle's
        sub = {k: v for k, v in jobs.items()}
        if ok.all():
            _, cnt = oracle.ksw_batch(sub, oracle.make_params(**kw), n_threads=4)
            assert cells == cnt["cells"]
