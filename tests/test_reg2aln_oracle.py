"""CPU tests of the mem_reg2aln restatement (oracle/global_oracle.c glb_reg2aln: band inference, bwa_gen_cigar2 with band-doubling
retries, deletion squeeze, soft clips, position) against golden vectors from the reference fork's own mem_reg2aln
(tests/golden/make_reg2aln_golden.py) and, when oracle/_ref is present, against that function live."""
import os

import numpy as np
import pytest

from tools import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _opt(CP, v):
    return CP.default_opt(a=int(v[0]), b=int(v[1]), o_del=int(v[2]), e_del=int(v[3]), o_ins=int(v[4]), e_ins=int(v[5]), w=int(v[6]))


def test_reg2aln_oracle_matches_fork_golden(oracle):
    from oracle import chain_py as CP
    gold = np.load(os.path.join(GOLD, "reg2aln_golden.npz"))
    ctg = CP.Contigs(tuple(int(x) for x in gold["contigs"]))
    g = synth.make_genome(ctg.l_pac, seed=int(gold["genome_seed"]))
    reads, regs = gold["reads"], gold["regs"]
    waves = 0
    for oi in range(2):
        opt = _opt(CP, gold[f"opt_{oi}"])
        kp = oracle.make_params(a=opt.a, b=opt.b, o_del=opt.o_del, e_del=opt.e_del, o_ins=opt.o_ins, e_ins=opt.e_ins)
        rec, cig = gold[f"rec_{oi}"], gold[f"cigar_{oi}"]
        for k, (i, qb, qe, rb, re, truesc, w) in enumerate(regs):
            a, c = CP.oracle_reg2aln(opt, kp, ctg, g, reads[i], qb, qe, rb, re, truesc, w)
            assert (int(a["pos"]), int(a["rid"]), int(a["is_rev"])) == tuple(int(x) for x in rec[k, :3]), k
            if rec[k, 1] >= 0:
                assert int(a["nm"]) == rec[k, 3] and int(a["n_cigar"]) == rec[k, 4], k
                assert (c == cig[k, :c.size]).all(), k
                waves = max(waves, int(a["n_waves"]))
    assert waves >= 2                                        # the band-doubling retry did run on some region


def test_reg2aln_oracle_matches_fork_live(oracle):
    from oracle import chain_py as CP
    if not CP.have_fork():
        pytest.skip("oracle/_ref/libforkmem.so not built (no /root/reference here)")
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk", os.path.join(GOLD, "make_reg2aln_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    ctg = CP.Contigs((20000, 25000))
    g = synth.make_genome(ctg.l_pac, seed=31337)
    pac = CP.make_pac(g)
    reads, regs = mk.make_cases(g, 300, 99, read_len=101)
    opt = CP.default_opt(w=50)
    kp = oracle.make_params()
    for (i, qb, qe, rb, re, truesc, w) in regs:
        want = CP.fork_reg2aln(opt, ctg, pac, reads[i], qb, qe, rb, re, truesc, w)
        a, c = CP.oracle_reg2aln(opt, kp, ctg, g, reads[i], qb, qe, rb, re, truesc, w)
        assert (int(a["pos"]), int(a["rid"]), int(a["is_rev"])) == (want["pos"], want["rid"], want["is_rev"])
        if want["rid"] >= 0:
            assert int(a["nm"]) == want["nm"] and (c == want["cigar"]).all()
