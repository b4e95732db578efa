"""CPU tests of the CIGAR-path oracle (ksw_global2 / bwa_gen_cigar2 restated in oracle/global_oracle.c) against
golden vectors from the reference's own functions (tests/golden/make_global_golden.py) and, when oracle/_ref is present, against
those functions live on other inputs."""
import ctypes as C
import os

import numpy as np
import pytest

from tools import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_global_oracle_matches_reference_golden(oracle):
    gold = np.load(os.path.join(GOLD, "global_golden.npz"))
    p = oracle.make_params()
    for si in range(3):
        jobs = {k: gold[f"s{si}_{k}"] for k in ("qseq", "tseq", "qoff", "toff", "qlen", "tlen", "w")}
        stride = gold[f"s{si}_cigar"].shape[1]
        res = oracle.global_batch(jobs, p, cig_stride=stride, n_threads=2)
        assert (res["score"] == gold[f"s{si}_score"]).all()
        assert (res["n_cigar"] == gold[f"s{si}_n_cigar"]).all()
        assert (res["cigar"] == gold[f"s{si}_cigar"]).all()
        assert res["cells"] > 0
    g = synth.make_genome(int(gold["gc_genome_len"]), seed=int(gold["gc_genome_seed"]))
    qoff = gold["gc_qoff"]
    for i in range(gold["gc_rb"].size):
        q = gold["gc_query"][qoff[i]:qoff[i + 1]]
        got = oracle.gen_cigar2(p, int(gold["gc_w"][i]), g, q, int(gold["gc_rb"][i]), int(gold["gc_re"][i]))
        assert got is not None
        sc, nm, cig = got
        n = int(gold["gc_n_cigar"][i])
        assert sc == gold["gc_score"][i] and nm == gold["gc_nm"][i] and cig.size == n
        assert (cig == gold["gc_cigar"][i, :n]).all()


def test_global_oracle_matches_reference_live(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (no /root/reference here)")
    R = oracle.ref_lib()
    for kw, pkw in ((dict(n_jobs=400, qlen_range=(1, 200), seed=401, w_extra=(0, 10)), {}),
                    (dict(n_jobs=300, qlen_range=(50, 250), seed=402, sub_rate=0.1, indel_rate=0.04, w_extra=(0, 3)), dict(a=2, b=3, o_del=4, e_del=2, o_ins=5, e_ins=1))):
        p = oracle.make_params(**pkw)
        mat = np.frombuffer(bytes(p.mat), dtype=np.int8).copy()
        jobs = synth.make_global_jobs(**kw)
        res = oracle.global_batch(jobs, p, cig_stride=128, n_threads=2)
        for a in range(jobs["qlen"].size):
            q = jobs["qseq"][jobs["qoff"][a]:jobs["qoff"][a] + jobs["qlen"][a]].copy()
            t = jobs["tseq"][jobs["toff"][a]:jobs["toff"][a] + jobs["tlen"][a]].copy()
            k = C.c_int(0)
            row = np.zeros(128, np.uint32)
            sc = R.ref_ksw_global2(q.size, q, t.size, t, mat, p.o_del, p.e_del, p.o_ins, p.e_ins, int(jobs["w"][a]), C.byref(k), row, 128)
            assert sc == res["score"][a] and k.value == res["n_cigar"][a]
            assert (row[:k.value] == res["cigar"][a, :k.value]).all()


def test_global_band_rule(oracle):
    p = oracle.make_params()
    # src/bwa.c:161-169 by hand: l_query 100, rlen 104, w_ 100 -> max_gap = int((50 - 6) / 1 + 1) = 45; w = (45 + 4 + 1) >> 1 = 25; min_w = 7
    assert oracle.global_band(p, 100, 100, 104) == 25
    assert oracle.global_band(p, 5, 100, 104) == 7          # w_ below min_w: min_w wins
    assert oracle.global_band(p, 100, 10, 10) == 3          # max_gap 1 -> w = 1, min_w = 3
