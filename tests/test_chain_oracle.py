"""CPU: the chaining / extension-job oracle (oracle/chain_oracle.c) against the unmodified reference fork
(oracle/_ref/libforkmem.so = src/bwamem.c's mem_chain, mem_chain_flt, mem_flt_chained_seeds, mem_chain2aln),
and against the committed golden file made from it (tests/golden/make_chain_golden.py)."""
import os

import numpy as np
import pytest

from oracle import chain_py as CP
from tools import chain_cases as CC

LENS = (30000, 1500, 20000)
MAXOCC = 50
GOLD = os.path.join(os.path.dirname(__file__), "golden", "chain_golden.npz")


def _check_against_fork(seed, n_reads):
    ctg = CP.Contigs(LENS, alt=[0, 1, 0])
    opt = CP.default_opt(max_occ=MAXOCC)
    fwd, cases = CC.make_cases(seed, n_reads, LENS, MAXOCC)
    pac = CP.make_pac(fwd)
    seen = {"gt9": 0, "dup": 0, "sides0": 0, "skipped": 0}
    for query, rb, qq, sc in cases:
        fc, fs, fr, fj, fseq = CP.fork_read(opt, ctg, pac, query, rb, qq, sc)
        seen["gt9"] += len(fc) > 9
        seen["dup"] += len(set(fc["pos"])) < len(fc)
        seen["sides0"] += int((fr["align_sides"] == 0).sum())
        seen["skipped"] += len(fs) - len(fr)
        for layout_all in (1, 0):
            a = (rb, qq, sc) if layout_all else CC.to_compact(rb, qq, sc, MAXOCC)
            oc, osd = CP.oracle_chains(opt, ctg, len(query), a[0], a[1], a[2], layout_all)
            assert oc.tobytes() == fc.tobytes() and osd.tobytes() == fs.tobytes()
        regs, jobs, seqs = CP.oracle_chain2aln(opt, ctg, fwd, query, oc, osd)
        assert regs.tobytes() == fr.tobytes()
        for s in (0, 1):
            assert jobs[s].tobytes() == fj[s].tobytes()
            assert seqs[s][0].tobytes() == fseq[s][0].tobytes() and seqs[s][1].tobytes() == fseq[s][1].tobytes()
    return seen


@pytest.mark.skipif(not CP.have_fork(), reason="oracle/_ref/libforkmem.so not built")
def test_chain_oracle_equals_reference_fork():
    seen = {"gt9": 0, "dup": 0, "sides0": 0, "skipped": 0}
    for seed in (11, 12):
        for k, v in _check_against_fork(seed, 350).items():
            seen[k] += v
    # the cases must actually reach the hard branches: split chain trees, equal keys, whole-read seeds, skipped seeds
    assert seen["gt9"] > 20 and seen["dup"] > 10 and seen["sides0"] > 10 and seen["skipped"] > 100


@pytest.mark.skipif(not CP.have_fork(), reason="oracle/_ref/libforkmem.so not built")
def test_chain_oracle_default_options_single_contig():
    # the bench configuration: one contig, max_occ 500, the fork's w = 300
    ctg = CP.Contigs((40000,))
    opt = CP.default_opt()
    fwd, cases = CC.make_cases(5, 200, (40000,), 500)
    pac = CP.make_pac(fwd)
    for query, rb, qq, sc in cases:
        fc, fs, fr, fj, _ = CP.fork_read(opt, ctg, pac, query, rb, qq, sc)
        oc, osd = CP.oracle_chains(opt, ctg, len(query), rb, qq, sc, 1)
        assert oc.tobytes() == fc.tobytes() and osd.tobytes() == fs.tobytes()
        regs, jobs, _ = CP.oracle_chain2aln(opt, ctg, fwd, query, oc, osd)
        assert regs.tobytes() == fr.tobytes() and jobs[0].tobytes() == fj[0].tobytes() and jobs[1].tobytes() == fj[1].tobytes()


@pytest.mark.skipif(not CP.have_fork(), reason="oracle/_ref/libforkmem.so not built")
@pytest.mark.parametrize("lens,max_occ,seed", [((30000, 1500, 20000), 50, 41), ((40000,), 500, 42)])
def test_chain_oracle_long_reads_equal_reference_fork(lens, max_occ, seed):
    # reads of 760 bases and more: the fork runs mem_flt_chained_seeds / mem_seed_sw (ksw_align2) on them, src/bwamem.c:774-808,970-990
    ctg = CP.Contigs(lens, alt=[0, 1, 0][:len(lens)])
    opt = CP.default_opt(max_occ=max_occ)
    fwd, cases = CC.make_long_cases(seed, 60, lens, max_occ)
    pac = CP.make_pac(fwd)
    moved = 0
    for query, rb, qq, sc in cases:
        fc, fs, fr, fj, fseq = CP.fork_read(opt, ctg, pac, query, rb, qq, sc)
        oc, osd = CP.oracle_chains(opt, ctg, len(query), rb, qq, sc, 1, fwd, query)
        assert oc.tobytes() == fc.tobytes() and osd.tobytes() == fs.tobytes()
        regs, jobs, seqs = CP.oracle_chain2aln(opt, ctg, fwd, query, oc, osd)
        assert regs.tobytes() == fr.tobytes()
        for s in (0, 1):
            assert jobs[s].tobytes() == fj[s].tobytes()
            assert seqs[s][0].tobytes() == fseq[s][0].tobytes() and seqs[s][1].tobytes() == fseq[s][1].tobytes()
        moved += int((osd["score"] != osd["len"]).sum())
    assert moved > 200


def test_chain_oracle_equals_golden():
    g = np.load(GOLD)
    ctg = CP.Contigs(g["contig_lens"], alt=g["contig_alt"])
    opt = CP.default_opt(max_occ=int(g["max_occ"]))
    fwd, cases = CC.make_cases(int(g["seed"]), int(g["n_reads"]), tuple(int(x) for x in g["contig_lens"]), int(g["max_occ"]))
    chains, cseeds, regs, js, jl = [], [], [], [], []
    for query, rb, qq, sc in cases:
        oc, osd = CP.oracle_chains(opt, ctg, len(query), rb, qq, sc, 1)
        r, jobs, _ = CP.oracle_chain2aln(opt, ctg, fwd, query, oc, osd)
        chains.append(oc); cseeds.append(osd); regs.append(r); js.append(jobs[0]); jl.append(jobs[1])
    for name, parts in (("chains", chains), ("cseeds", cseeds), ("regs", regs), ("jobs_short", js), ("jobs_long", jl)):
        got = np.concatenate(parts)
        assert got.tobytes() == g[name].tobytes(), name


def test_chain_oracle_long_reads_equal_golden():
    # the seed filter of long reads against vectors made from the fork (tests/golden/make_chain_golden.py, second part)
    g = np.load(os.path.join(os.path.dirname(GOLD), "chain_long_golden.npz"))
    ctg = CP.Contigs(g["contig_lens"], alt=g["contig_alt"])
    opt = CP.default_opt(max_occ=int(g["max_occ"]))
    fwd, cases = CC.make_long_cases(int(g["seed"]), int(g["n_reads"]), tuple(int(x) for x in g["contig_lens"]), int(g["max_occ"]))
    chains, cseeds, regs, js, jl = [], [], [], [], []
    for query, rb, qq, sc in cases:
        oc, osd = CP.oracle_chains(opt, ctg, len(query), rb, qq, sc, 1, fwd, query)
        r, jobs, _ = CP.oracle_chain2aln(opt, ctg, fwd, query, oc, osd)
        chains.append(oc); cseeds.append(osd); regs.append(r); js.append(jobs[0]); jl.append(jobs[1])
    for name, parts in (("chains", chains), ("cseeds", cseeds), ("regs", regs), ("jobs_short", js), ("jobs_long", jl)):
        assert np.concatenate(parts).tobytes() == g[name].tobytes(), name
    assert (g["cseeds"]["score"] != g["cseeds"]["len"]).sum() > 300


def test_regs_finish_arithmetic():
    # src/bwamem.c:2286-2306: score = left + right - seedlen (two sides), ends measured from the seed
    regs = np.zeros(3, dtype=CP.REG_DT)
    regs["seedlen0"] = [30, 40, 150]; regs["align_sides"] = [2, 1, 0]; regs["where_is_long"] = [1, 0, 0]
    regs["query_seed_begin"] = [50, 110, 0]; regs["target_seed_begin"] = [1000, 2000, 3000]; regs["score"] = [30, 40, 150]
    long_t = np.array([[60, 70, 72], [90, 110, 111]], np.int32)      # reg0 right (long), reg1 left (long)
    short_t = np.array([[55, 50, 49]], np.int32)                      # reg0 left (short)
    out = CP.oracle_regs_finish(150, regs, short_t, long_t)
    assert (out["score"][0], out["qb"][0], out["qe"][0], out["rb"][0], out["re"][0]) == (55 + 60 - 30, 0, 150, 1000 - 49, 1000 + 30 + 72)
    assert (out["score"][1], out["qb"][1], out["qe"][1], out["rb"][1], out["re"][1]) == (90, 0, 150, 2000 - 111, 2040)
    assert (out["score"][2], out["qb"][2], out["qe"][2], out["rb"][2], out["re"][2]) == (150, 0, 150, 3000, 3150)


@pytest.mark.skipif(not (CP.have_fork() and CP.O.have_ref()), reason="oracle/_ref not built")
def test_reference_chained_pipeline_equals_oracle_batch(pkg, small_index):
    """bench.py's CPU arm for the chained step (the reference's bwt_smem1 / bwt_sa, the fork's mem_chain .. mem_chain2aln and
    ksw_extend2, multi-threaded) gives the oracle's regions, read by read"""
    from oracle import oracle_py as O
    from tools import synth
    g, prefix = small_index
    n, L = 600, 150
    reads, _, _ = synth.make_reads(g, n, L, seed=91, sub_rate=0.02, n_rate=0.002)
    flat = reads.reshape(-1).copy()
    off = (np.arange(n + 1) * L).astype(np.uint64)
    ctg = CP.Contigs((g.size,))
    opt = CP.default_opt(w=100)
    kp = O.make_params()
    h = O.ref_lib().ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
    assert h
    got = CP.ref_chained_pipeline(h, opt, ctg, CP.make_pac(g), flat, off, kp, 19, n_threads=3)
    oi = O.OracleIndex(prefix + ".bwt", prefix + ".sa")
    sd = oi.seed_batch(flat, off, 19, 0, n_threads=2)            # all rows of every SMEM group
    assert sd["total"] == got["n_seeds"]
    want = CP.oracle_align_batch(opt, ctg, g, reads, sd["rbeg"], np.stack([sd["qbeg"], sd["qend"]], axis=1).astype(np.int32), sd["score"],
                                 sd["n_seeds"], sd["seed_off"], 1, kp, n_threads=2)
    assert (got["n_regs"] == want["n_regions"]).all() and got["n_jobs"] == len(want["jobs"])
    assert len(got["regs"]) == len(want["aln"]) > n
    for f in ("rb", "re", "qb", "qe", "score", "truesc"):
        assert (got["regs"][f] == want["aln"][f]).all(), f
    assert (got["regs"]["seedcov"] == want["regs"]["seedcov"]).all() and (got["regs"]["rid"] == want["regs"]["rid"]).all()
