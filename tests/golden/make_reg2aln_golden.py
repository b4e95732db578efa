"""Golden vectors for the CIGAR stage at the call-site level: the REFERENCE FORK's own mem_reg2aln (src/bwamem.c:2344-2438, reached
through oracle/fork_mem_shim.cpp inside oracle/_ref/libforkmem.so) on synthetic alignment regions.

Run in the build container only:   python tests/golden/make_reg2aln_golden.py
  reg2aln_golden.npz   genome seed / contigs, reads, regions (read, qb, qe, rb, re, truesc, w) and, per region, pos / rid / is_rev / NM /
                       n_cigar / CIGAR (soft clips included, leading or trailing deletion squeezed out)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import chain_py as CP  # noqa: E402
from tools import synth  # noqa: E402

CTG = (30000, 12000, 18000)
STRIDE = 64


def make_cases(g, n, seed, read_len=150):
    """reads + plausible alignment regions around their true locus (both strands), clipped ends, a spread of truesc / w"""
    rng = np.random.default_rng(seed)
    L = g.size
    reads, pos, strand = synth.make_reads(g, n, read_len, seed=seed, sub_rate=0.02, ins_rate=0.004, del_rate=0.004)
    regs = []
    for i in range(n):
        qb = int(rng.choice([0, 0, 0, 1, 3, 7])); qe = read_len - int(rng.choice([0, 0, 0, 2, 5, 11]))
        d = int(rng.integers(-3, 4))
        rlen = max(1, qe - qb + d)
        p0 = int(pos[i])
        if strand[i] == 0:
            rb = p0 + qb
        else:
            rb = 2 * L - (p0 + read_len) + qb
        rb = max(0, min(rb, 2 * L - rlen))
        re = rb + rlen
        if rb < L < re:
            rb, re = L - rlen, L
        truesc = max(1, (qe - qb) - int(rng.choice([0, 5, 10, 20, 40, 90])))
        regs.append((i, qb, qe, rb, re, truesc, int(rng.choice([100, 100, 10, 3]))))
    regs.append((0, 0, read_len, -1, -1, 50, 100))                          # unmapped record
    regs.append((1, 10, 60, regs[1][3] + 10, regs[1][3] + 60, 50, 100))    # equal lengths, perfect score: the ungapped shortcut (w2 == 0)
    return reads, np.array(regs, np.int64)


def main():
    assert CP.have_fork(), "oracle/_ref/libforkmem.so missing: run oracle/build_ref.sh"
    g = synth.make_genome(sum(CTG), seed=9090)
    ctg = CP.Contigs(CTG)
    pac = CP.make_pac(g)
    reads, regs = make_cases(g, 800, 21)
    out = dict(genome_seed=9090, contigs=np.array(CTG), reads=reads, regs=regs)
    for oi, okw in enumerate((dict(w=100), dict(w=20, a=2, b=5, o_del=7, e_del=2, o_ins=5, e_ins=1))):
        opt = CP.default_opt(**okw)
        rec = np.zeros((len(regs), 5), np.int64); cig = np.zeros((len(regs), STRIDE), np.uint32)
        for k, (i, qb, qe, rb, re, truesc, w) in enumerate(regs):
            r = CP.fork_reg2aln(opt, ctg, pac, reads[i], qb, qe, rb, re, truesc, w)
            rec[k] = (r["pos"], r["rid"], r["is_rev"], r["nm"], r["n_cigar"])
            assert r["n_cigar"] <= STRIDE
            cig[k, :r["n_cigar"]] = r["cigar"]
        out[f"rec_{oi}"] = rec; out[f"cigar_{oi}"] = cig
        out[f"opt_{oi}"] = np.array([opt.a, opt.b, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, opt.w])
        print(f"opt {oi}: {len(regs)} regions, reverse {int(rec[:, 2].sum())}, mean NM {rec[rec[:, 1] >= 0, 3].mean():.2f}, n_cigar max {rec[:, 4].max()}")
    np.savez_compressed(os.path.join(HERE, "reg2aln_golden.npz"), **out)


if __name__ == "__main__":
    main()
