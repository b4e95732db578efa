"""Golden vectors for the stage between the extension results and SAM: the REFERENCE FORK's own mem_sort_dedup_patch
(src/bwamem.c:620-681), is_alt marking (:2321-2325), mem_mark_primary_se (:715-760) and mem_approx_mapq_se (:1690-1716), reached through
fork_finish_regs (oracle/fork_mem_shim.cpp) inside oracle/_ref/libforkmem.so, on the synthetic region sets of oracle/region_py.py.

Run in the build container only:   python tests/golden/make_region_golden.py
  region_golden.npz   genome seed / contigs / ALT flags, reads, per read the input regions and, per option set, the finished regions + n_pri
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import chain_py as CP  # noqa: E402
from oracle import region_py as RP  # noqa: E402
from tools import synth  # noqa: E402

CTG = (30000, 12000, 18000, 9000)
ALT = (0, 0, 1, 1)
OPTS = (dict(), dict(a=2, b=5, o_del=7, e_del=2, o_ins=5, e_ins=1, w=20, mask_level=0.3, mask_level_redun=0.8, mapQ_coef_len=0, mapQ_coef_fac=0,
                     max_chain_gap=200))


def make_inputs(n_reads=500, seed=77):
    ctg = CP.Contigs(CTG, alt=ALT)
    g = synth.make_genome(ctg.l_pac, seed=4242)
    reads, pos, strand = synth.make_reads(g, n_reads, 150, seed=seed, sub_rate=0.02, ins_rate=0.003, del_rate=0.003)
    return ctg, g, reads, RP.make_cases(g, ctg, reads, pos, strand, seed + 1)


def main():
    assert CP.have_fork(), "oracle/_ref/libforkmem.so missing: run oracle/build_ref.sh"
    ctg, g, reads, cases = make_inputs()
    pac = CP.make_pac(g)
    off = np.concatenate([[0], np.cumsum([len(c) for c in cases])]).astype(np.int64)
    out = dict(genome_seed=4242, contigs=np.array(CTG), alt=np.array(ALT), reads=reads, regs_in=np.concatenate(cases), in_off=off)
    for oi, kw in enumerate(OPTS):
        opt = RP.default_opt(**kw)
        res, npri = [], []
        for i, regs in enumerate(cases):
            a, n_pri = RP.fork_finish(opt, ctg, pac, reads[i], regs, i)
            res.append(a); npri.append(n_pri)
        out[f"out_{oi}"] = np.concatenate(res)
        out[f"out_off_{oi}"] = np.concatenate([[0], np.cumsum([len(r) for r in res])]).astype(np.int64)
        out[f"n_pri_{oi}"] = np.array(npri, np.int32)
        o = out[f"out_{oi}"]
        print(f"opt {oi}: {len(out['regs_in'])} regions in, {len(o)} out, {int((o['n_comp'] > 1).sum())} merged, "
              f"{int((o['secondary'] >= 0).sum())} secondary, mapq>0 on {int((o['mapq'] > 0).sum())}")
    np.savez_compressed(os.path.join(HERE, "region_golden.npz"), **out)


if __name__ == "__main__":
    main()
