"""The column-pair s16x2 extension kernel's source (ext_pair_core.cuh), built for the host with the DPX
intrinsics emulated, against the oracle.  Runs on the CPU box: it checks the kernel's row logic
(window head/tail cells, F carry across packed column pairs, keyed row maximum) bit for bit without
a GPU.  The GPU parity tests run the same source through the real instructions."""
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
    so = os.path.join(HERE, "libextpair_host.so")
    srcs = [os.path.join(HERE, "ext_pair_host.cpp"), os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc", "ext_pair_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc"), srcs[0], "-o", so])
    L = C.CDLL(so)
    L.ext_pair_host_run.restype = C.c_longlong
    L.ext_pair_host_run.argtypes = [C.c_void_p, C.c_int, C.c_uint64] + [C.c_void_p] * 9
    L.ext_pair_host_run_wide.restype = C.c_longlong
    L.ext_pair_host_run_wide.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 9

    def run(jobs, ep, variant):
        n = jobs["qlen"].size
        res = np.zeros((n, 6), np.int32)
        skipped = np.zeros(n, np.uint8)
        cells = L.ext_pair_host_run(C.addressof(ep), int(variant), n, jobs["qseq"].ctypes.data, jobs["qoff"].ctypes.data, jobs["qlen"].ctypes.data,
                                    jobs["tseq"].ctypes.data, jobs["toff"].ctypes.data, jobs["tlen"].ctypes.data, jobs["h0"].ctypes.data,
                                    res.ctypes.data, skipped.ctypes.data)
        return res, skipped.astype(bool), cells
    run.lib = L
    return run


KW = [dict(w=100, zdrop=100), dict(w=16, zdrop=100), dict(w=8, zdrop=0), dict(w=50, zdrop=30), dict(w=300, zdrop=0, use_band=0),
      dict(w=33, zdrop=100, o_del=3, e_del=1, o_ins=5, e_ins=2, a=3, b=2), dict(w=2, zdrop=10), dict(w=1, zdrop=0, end_bonus=0)]
SETS = [(61, dict(qlen_range=(1, 260), h0_range=(1, 250))),
        (62, dict(qlen_range=(1, 120), sub_rate=0.25, indel_rate=0.08, n_job_frac=0.3, h0_range=(1, 40))),
        (63, dict(qlen_range=(100, 500), h0_range=(100, 400), sub_rate=0.02, indel_rate=0.02)),
        (64, dict(qlen_range=(1, 40), h0_range=(1, 300), sub_rate=0.5, indel_rate=0.2)),
        (65, dict(qlen_range=(1, 12), h0_range=(1, 12), sub_rate=0.4, indel_rate=0.3, n_job_frac=0.5)),
        (66, dict(qlen_range=(100, 128), h0_range=(800, 895), sub_rate=0.01, indel_rate=0.005))]   # scores up to the 1023 bound


@pytest.mark.parametrize("variant", [0, 1, 2, 3], ids=["u8", "u4", "ring_u8", "ring_u4"])
@pytest.mark.parametrize("kw", KW)
def test_pair_source_matches_oracle(pkg, oracle, emul, kw, variant):
    for seed, extra in SETS:
        jobs = synth.make_ext_jobs(700 if seed == 63 else 2000, w=kw["w"], seed=seed, **extra)       # the long queries are the slow ones on the host
        want, _ = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        res, skipped, cells = emul(jobs, pkg.ext_params(**kw), variant)
        assert cells >= 0
        ok = ~skipped
        if seed not in (63, 66):
            assert ok.sum() > 100
        bad = np.nonzero((res[ok] != want[ok]).any(axis=1))[0]
        assert bad.size == 0, (seed, np.nonzero(ok)[0][bad[:3]], res[ok][bad[:3]], want[ok][bad[:3]])
        if ok.all():      # evaluated-cell count equals the oracle's
            _, cnt = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
            assert cells == cnt["cells"]


def test_pair_source_general_matrix(pkg, oracle, emul):
    """any byte matrix with one score for a query N is eligible (the PRMT table is a row of the matrix)"""
    jobs = synth.make_ext_jobs(3000, w=100, seed=71, qlen_range=(1, 128), h0_range=(1, 150), n_job_frac=0.2)
    P = oracle.make_params(w=100, zdrop=100)
    mat = np.array([[2, -3, -1, -3, -2], [-3, 2, -3, -1, -2], [-1, -3, 2, -3, -2], [-3, -1, -3, 2, -2], [-1, -4, -1, -2, -2]], np.int8)
    ep = pkg.ext_params(w=100, zdrop=100)
    for i in range(25):
        P.mat[i] = int(mat.reshape(-1)[i])
        ep.mat[i] = int(mat.reshape(-1)[i])
    want, _ = oracle.ksw_batch(jobs, P, n_threads=4)
    for variant in (0, 3):
        res, skipped, cells = emul(jobs, ep, variant)
        ok = ~skipped
        assert cells >= 0 and ok.sum() > 1000
        assert (res[ok] == want[ok]).all()


def test_ring_slot_division_is_exact():
    """SLOT(p) = p - umulhi(p, ceil(2^32 / R)) * R is p mod R for every pair index (queries up to 65535 bases) and every ring size"""
    p = np.concatenate([np.arange(0, 3000), np.arange(32000, 32800)]).astype(np.uint64)
    for R in list(range(2, 300)) + [511, 512, 513, 700, 1023, 2047, 2048]:
        magic = ((1 << 32) + R - 1) // R
        assert magic < (1 << 32)
        assert ((p - ((p * magic) >> 32) * R) == p % R).all(), R


@pytest.mark.parametrize("kw", [dict(w=100, zdrop=100), dict(w=16, zdrop=100), dict(w=2, zdrop=10), dict(w=1, zdrop=0, end_bonus=0),
                                dict(w=33, zdrop=100, o_del=3, e_del=1, o_ins=5, e_ins=2, a=3, b=2), dict(w=300, zdrop=0)],
                         ids=lambda k: f"w{k['w']}")
def test_pair_source_wide_scores_long_queries(pkg, oracle, emul, kw):
    """the WIDE instantiation (lane per job for large batches of long banded jobs): scores beyond 1023, queries beyond 512, 32-bit
    row-maximum keys, ring state, against the oracle -- results and evaluated cells"""
    sets = [(81, dict(qlen_range=(1, 260), h0_range=(1, 250))),
            (82, dict(qlen_range=(300, 3000), h0_range=(1, 900), sub_rate=0.05, indel_rate=0.02, n_job_frac=0.1)),
            (83, dict(qlen_range=(100, 700), h0_range=(900, 3000), sub_rate=0.01, indel_rate=0.005)),
            (84, dict(qlen_range=(1, 60), h0_range=(1, 5000), sub_rate=0.4, indel_rate=0.2, n_job_frac=0.3))]
    for seed, extra in sets:
        n = 150 if seed == 82 else 1500
        jobs = synth.make_ext_jobs(n, w=kw["w"], seed=seed, **extra)
        want, cnt = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        res = np.zeros((n, 6), np.int32)
        skipped = np.zeros(n, np.uint8)
        ep = pkg.ext_params(**kw)
        cells = emul.lib.ext_pair_host_run_wide(C.addressof(ep), n, jobs["qseq"].ctypes.data, jobs["qoff"].ctypes.data, jobs["qlen"].ctypes.data,
                                                jobs["tseq"].ctypes.data, jobs["toff"].ctypes.data, jobs["tlen"].ctypes.data, jobs["h0"].ctypes.data,
                                                res.ctypes.data, skipped.ctypes.data)
        assert cells >= 0
        ok = ~skipped.astype(bool)
        assert ok.sum() > 0.9 * n
        bad = np.nonzero((res[ok] != want[ok]).any(axis=1))[0]
        assert bad.size == 0, (seed, np.nonzero(ok)[0][bad[:3]], res[ok][bad[:3]], want[ok][bad[:3]])
        if ok.all():
            assert cells == cnt["cells"]


CF_KW = [(dict(w=100, zdrop=100), True), (dict(w=8, zdrop=0), True), (dict(w=6, zdrop=4), True), (dict(w=11, zdrop=12), True), (dict(w=300, zdrop=0, use_band=0), True),
         (dict(w=100, zdrop=0, end_bonus=0), True),                              # short queries: the band clamp of src/ksw.c:885-893 falls below dmax + 2
         (dict(w=20, zdrop=50, a=2, b=3), True), (dict(w=20, zdrop=10, a=2, b=5, o_del=7, e_del=2, o_ins=8, e_ins=1), True),
         (dict(w=100, zdrop=3), True),                                            # only jobs without a difference pass the z-drop condition
         (dict(w=5, zdrop=100), False),                                           # band narrower than the gaps that have to be ruled out + 2
         (dict(w=33, zdrop=100, o_del=3, e_del=1, o_ins=5, e_ins=2, a=3, b=2), False),   # a gap is cheaper than a mismatch + a match
         (dict(w=50, zdrop=100, a=1, b=6), False)]


@pytest.mark.parametrize("kw,eligible", CF_KW)
def test_closed_form_jobs_match_oracle(pkg, oracle, emul, kw, eligible):
    """closed_form_job (ext_pair_core.cuh): every job it takes has exactly ksw_extend2's six outputs, it takes exactly the jobs of the
    documented shape (a few substitutions, shifted diagonals ruled out between them), in both sequence forms, and only under the documented
    parameter conditions"""
    L = emul.lib
    L.ext_closed_form_host.restype = C.c_longlong
    L.ext_closed_form_host.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 9
    ep = pkg.ext_params(**kw)
    dmax = synth.closed_form_eligible(**kw)
    assert (dmax is not None) == eligible
    n_taken = n_two = n_rep = 0
    n_more = {3: 0, 4: 0, 5: 0, 6: 0}
    for seed, extra in ((71, dict()), (72, dict(qlen_range=(1, 12), h0_range=(1, 30))), (73, dict(qlen_range=(100, 600), h0_range=(19, 250))),
                        (74, dict(qlen_range=(20, 160), h0_range=(19, 150)))):
        jobs = synth.make_flank_jobs(2500, seed=seed, w=kw["w"], **extra)
        n = jobs["qlen"].size
        res = np.zeros((n, 6), np.int32); flags = np.zeros(n, np.uint8)
        taken = L.ext_closed_form_host(C.addressof(ep), n, jobs["qseq"].ctypes.data, jobs["qoff"].ctypes.data, jobs["qlen"].ctypes.data,
                                       jobs["tseq"].ctypes.data, jobs["toff"].ctypes.data, jobs["tlen"].ctypes.data, jobs["h0"].ctypes.data,
                                       res.ctypes.data, flags.ctypes.data)
        if not eligible:
            assert taken == -1
            continue
        assert taken >= 0, taken
        want, _ = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        got = flags != 0
        assert (flags[got] == 3).all()
        mask = synth.closed_form_mask(jobs, kw.get("a", 1), kw.get("b", 4), dmax, kw["zdrop"])
        assert (got == mask).all(), np.nonzero(got != mask)[0][:5]
        bad = np.nonzero((res != want).any(axis=1) & got)[0]
        assert bad.size == 0, (bad[:5], res[bad[:3]], want[bad[:3]], jobs["qlen"][bad[:3]], jobs["h0"][bad[:3]])
        n_taken += int(got.sum())
        # how many of the taken jobs have two differences, and how many two-difference jobs were refused (repeats, adjacent differences)
        for k in np.nonzero(jobs["tlen"] >= jobs["qlen"])[0]:
            ql = int(jobs["qlen"][k])
            q = jobs["qseq"][int(jobs["qoff"][k]):int(jobs["qoff"][k]) + ql]; t = jobs["tseq"][int(jobs["toff"][k]):int(jobs["toff"][k]) + ql]
            if (q < 4).all() and (t < 4).all() and int((q != t).sum()) == 2:
                n_two += int(got[k]); n_rep += int(not got[k])
            if (q < 4).all() and (t < 4).all() and int((q != t).sum()) in n_more:
                n_more[int((q != t).sum())] += int(got[k])
    if eligible and (kw["zdrop"] == 0 or kw["zdrop"] >= 2 * kw.get("b", 4)):
        assert n_taken > 2400 and n_two > 500 and n_rep > 100, (n_taken, n_two, n_rep)
        print(kw, n_taken, n_two, n_rep, n_more)
        for kk, need in ((3, 150), (4, 60), (5, 20), (6, 5)):
            if kk in dmax and (kw["zdrop"] == 0 or kw["zdrop"] >= kk * kw.get("b", 4)):
                assert n_more[kk] > need, (kk, n_more)
    elif eligible:
        assert n_taken > 150


def test_closed_form_on_general_jobs(pkg, oracle, emul):
    # the usual fuzz sets: whatever the shortcut takes there must be right as well
    L = emul.lib
    L.ext_closed_form_host.restype = C.c_longlong
    L.ext_closed_form_host.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 9
    kw = dict(w=100, zdrop=100)
    ep = pkg.ext_params(**kw)
    for seed, extra in SETS:
        jobs = synth.make_ext_jobs(3000, w=100, seed=seed, **extra)
        n = jobs["qlen"].size
        res = np.zeros((n, 6), np.int32); flags = np.zeros(n, np.uint8)
        taken = L.ext_closed_form_host(C.addressof(ep), n, jobs["qseq"].ctypes.data, jobs["qoff"].ctypes.data, jobs["qlen"].ctypes.data,
                                       jobs["tseq"].ctypes.data, jobs["toff"].ctypes.data, jobs["tlen"].ctypes.data, jobs["h0"].ctypes.data,
                                       res.ctypes.data, flags.ctypes.data)
        assert taken >= 0
        want, _ = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        got = flags != 0
        assert (got == synth.closed_form_mask(jobs, 1, 4, None, 100)).all()
        assert (res[got] == want[got]).all()


@pytest.mark.parametrize("kw", [dict(w=100, zdrop=100), dict(w=21, zdrop=0), dict(w=40, zdrop=60, a=2, b=3),
                                dict(w=60, zdrop=100, o_del=4, e_del=2, o_ins=9, e_ins=3)])
def test_closed_form_on_repeats_with_many_differences(pkg, oracle, emul, kw):
    """the jobs the proof is about: up to six substitutions at spacings around dmax_k + 2, over tandem repeats whose shifted diagonals
    are clean but for a break or two -- whatever closed_form_job takes there must equal ksw_extend2, and it must take some of every k"""
    L = emul.lib
    L.ext_closed_form_host.restype = C.c_longlong
    L.ext_closed_form_host.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 9
    ep = pkg.ext_params(**kw)
    dmax = synth.closed_form_eligible(**kw)
    by_k = {}
    for seed in (31, 32):
        jobs = synth.make_repeat_flank_jobs(6000, seed, kw)
        n = jobs["qlen"].size
        res = np.zeros((n, 6), np.int32); flags = np.zeros(n, np.uint8)
        taken = L.ext_closed_form_host(C.addressof(ep), n, jobs["qseq"].ctypes.data, jobs["qoff"].ctypes.data, jobs["qlen"].ctypes.data,
                                       jobs["tseq"].ctypes.data, jobs["toff"].ctypes.data, jobs["tlen"].ctypes.data, jobs["h0"].ctypes.data,
                                       res.ctypes.data, flags.ctypes.data)
        assert taken >= 0
        want, _ = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        got = flags != 0
        assert (got == synth.closed_form_mask(jobs, kw.get("a", 1), kw.get("b", 4), dmax, kw["zdrop"])).all()
        bad = np.nonzero((res != want).any(axis=1) & got)[0]
        assert bad.size == 0, (bad[:5], res[bad[:3]], want[bad[:3]])
        for k in np.nonzero(got)[0]:
            ql = int(jobs["qlen"][k])
            q = jobs["qseq"][int(jobs["qoff"][k]):int(jobs["qoff"][k]) + ql]; t = jobs["tseq"][int(jobs["toff"][k]):int(jobs["toff"][k]) + ql]
            by_k[int((q != t).sum())] = by_k.get(int((q != t).sum()), 0) + 1
    for k in dmax:
        assert by_k.get(k, 0) > (3 if k == 6 else 20), by_k
