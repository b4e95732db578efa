"""Live check of the oracle (oracle/*.c restatements) against the REFERENCE's own functions compiled from /root/reference into
oracle/_ref/libbwaref.so (oracle/build_ref.sh): bwt_occ4 / bwt_sa / bwt_extend / bwt_smem1 (src/bwt.c:150-566 there), the seed rows
of mem_collect_intv pass 1 + mem_chain's SA sampling, and ksw_extend2 (src/ksw.c:860-980).  The golden vectors in tests/golden/ pin
the same functions on fixed inputs; this fuzzes them on fresh ones wherever the reference was built (skipped elsewhere)."""
import ctypes as C

import numpy as np
import pytest

from oracle import chain_py as CP
from oracle import oracle_py as O
from tools import synth

pytestmark = pytest.mark.skipif(not O.have_ref(), reason="oracle/_ref not built (no /root/reference here)")


def _ref_ksw(jobs, kw):
    p = O.make_params(**kw)
    n = jobs["qlen"].size
    out = np.zeros((n, 6), np.int32)
    mat = np.frombuffer(bytes(p.mat), dtype=np.int8).copy()
    O.ref_lib().ref_ksw_batch(n, jobs["qseq"], jobs["qoff"], jobs["qlen"], jobs["tseq"], jobs["toff"], jobs["tlen"], jobs["h0"], mat,
                              p.o_del, p.e_del, p.o_ins, p.e_ins, p.w, p.end_bonus, p.zdrop, out.reshape(-1), 4)
    return out


@pytest.mark.parametrize("kw", [dict(w=100, zdrop=100), dict(w=16, zdrop=100), dict(w=8, zdrop=0), dict(w=50, zdrop=30, end_bonus=0),
                                dict(w=33, zdrop=100, o_del=3, e_del=1, o_ins=5, e_ins=2, a=3, b=2),
                                dict(w=40, zdrop=60, a=2, b=5, o_del=7, e_del=2, o_ins=8, e_ins=1)])
def test_ksw_extend2_oracle_equals_reference(oracle, kw):
    sets = [synth.make_ext_jobs(3000, w=kw["w"], seed=901, qlen_range=(1, 300), h0_range=(1, 200)),
            synth.make_ext_jobs(3000, w=kw["w"], seed=902, qlen_range=(1, 120), sub_rate=0.25, indel_rate=0.08, n_job_frac=0.3, h0_range=(1, 40)),
            synth.make_ext_jobs(300, w=kw["w"], seed=903, qlen_range=(200, 700), h0_range=(19, 150)),
            synth.make_flank_jobs(3000, seed=904, w=kw["w"]),
            synth.make_repeat_flank_jobs(3000, 905, kw)]
    for jobs in sets:
        got, _ = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        want = _ref_ksw(jobs, kw)
        bad = np.nonzero((got != want).any(axis=1))[0]
        assert bad.size == 0, (kw, bad[:5], got[bad[:3]], want[bad[:3]])


@pytest.fixture(scope="module")
def both_indexes(pkg, tmp_path_factory):
    d = tmp_path_factory.mktemp("ovr")
    g = synth.make_repeat_genome(400_000, seed=91)             # 30 diverged copies of a 1 kb unit: intervals above max_occ
    prefix = str(d / "g")
    pkg.build_index(g, prefix, sa_intv=8, also_stock_layout=True, n_threads=4)
    oi = O.OracleIndex(prefix + ".bwt", prefix + ".sa")
    h = O.ref_lib().ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
    assert h
    yield g, oi, h
    O.ref_lib().ref_free(h)
    oi.close()


def test_occ_sa_oracle_equal_reference(both_indexes):
    g, oi, h = both_indexes
    R = O.ref_lib()
    rng = np.random.default_rng(5)
    n_rows = 2 * g.size + 1
    assert oi.seq_len == 2 * g.size
    ks = np.concatenate([rng.integers(0, n_rows, 3000), [0, 1, n_rows - 2, n_rows - 1], np.arange(120, 136), np.arange(n_rows // 2 - 3, n_rows // 2 + 3)])
    cnt = np.zeros(4, np.uint64)
    for k in ks:
        k = int(k)
        assert R.ref_sa(h, k) == oi.sa(k), k
        if k < n_rows - 1:
            R.ref_occ4(h, k, cnt)
            assert (cnt == oi.occ4(k)).all(), k


def test_seed_rows_oracle_equal_reference(both_indexes):
    """pass-1 SMEMs and their sampled SA rows per read: the oracle's seed arrays against the reference's own bwt_smem1 / bwt_sa"""
    g, oi, h = both_indexes
    r1, _, _ = synth.make_reads(g, 3000, 150, seed=41)
    r2, _, _ = synth.make_reads(g, 1000, 101, seed=42, n_rate=0.005, sub_rate=0.03)
    r3, _, _ = synth.make_reads(g, 600, 250, seed=43, sub_rate=0.002)
    seen_sampled = 0
    for reads in (r1, r2, r3):
        n, L = reads.shape
        flat = reads.reshape(-1).copy()
        off = (np.arange(n + 1) * L).astype(np.uint64)
        for max_occ in (500, 8):
            got = oi.seed_batch(flat, off, 19, max_occ, n_threads=4)
            want = CP.ref_seed_arrays(h, flat, off, 19, max_occ, n_threads=4)
            assert got["total"] == want["total"]
            assert (got["n_seeds"] == want["n_seeds"]).all()
            assert (got["rbeg"] == want["rbeg"]).all()
            assert (got["qbeg"] == want["qq"][:, 0]).all() and (got["qend"] == want["qq"][:, 1]).all()
            assert (got["score"] == want["score"]).all()
            seen_sampled += int((got["score"] > max_occ).sum())
    assert seen_sampled > 0                      # the max_occ sampling rule was exercised (repeats in the genome)


INTV_DT = np.dtype([("k", "<u8"), ("l", "<u8"), ("s", "<u8"), ("beg", "<i4"), ("end", "<i4")])


def test_smem1_and_extend_oracle_equal_reference(both_indexes):
    """bwt_smem1 call by call (every start position the reference's loop visits, min_intv 1 and larger: the re-seeding form) and
    bwt_extend in both directions on the intervals it returns"""
    g, oi, h = both_indexes
    R, L = O.ref_lib(), O.lib()
    vp = C.c_void_p
    L.fmd_smem1.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp]
    L.fmd_smem1.restype = C.c_int
    L.fmd_extend.argtypes = [vp, vp, vp, C.c_int, vp]
    reads, _, _ = synth.make_reads(g, 300, 150, seed=51, sub_rate=0.02, n_rate=0.004)
    n_calls = n_ext = 0
    for q in reads:
        q = np.ascontiguousarray(q)
        for min_intv in (1, 3, 12):
            x = 0
            while x < q.size:
                if q[x] > 3:
                    x += 1
                    continue
                want = np.zeros(5 * (q.size + 1), np.uint64); wn = C.c_int(0)
                got = np.zeros(q.size + 1, INTV_DT); gn = C.c_int(0)
                xr = R.ref_smem1(h, q.size, q, x, min_intv, want, C.byref(wn))
                xo = L.fmd_smem1(C.byref(oi.idx), q.size, q.ctypes.data, x, min_intv, got.ctypes.data, C.addressof(gn), None)
                assert xr == xo and wn.value == gn.value, (x, min_intv)
                w5 = want[:5 * wn.value].reshape(-1, 5); gg = got[:gn.value]
                assert (gg["k"] == w5[:, 0]).all() and (gg["l"] == w5[:, 1]).all() and (gg["s"] == w5[:, 2]).all()
                assert (gg["beg"] == w5[:, 3]).all() and (gg["end"] == w5[:, 4]).all()
                n_calls += 1
                if n_ext < 400:
                    for iv in gg[:2]:
                        for is_back in (0, 1):
                            ok_r = np.zeros(12, np.uint64); ok_o = np.zeros(4, INTV_DT)
                            R.ref_extend(h, np.array([iv["k"], iv["l"], iv["s"]], np.uint64), ok_r, is_back)
                            one = np.zeros(1, INTV_DT); one[0] = iv
                            L.fmd_extend(C.byref(oi.idx), one.ctypes.data, ok_o.ctypes.data, is_back, None)
                            r3 = ok_r.reshape(4, 3)
                            assert (ok_o["k"] == r3[:, 0]).all() and (ok_o["l"] == r3[:, 1]).all() and (ok_o["s"] == r3[:, 2]).all()
                            n_ext += 1
                x = xr
    assert n_calls > 1500 and n_ext >= 400
