"""Golden vectors for the chaining / extension-job stage, generated from the UNMODIFIED reference fork
(oracle/_ref/libforkmem.so, built by oracle/build_ref.sh).  Run here, where /root/reference exists:
    python tests/golden/make_chain_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import chain_py as CP  # noqa: E402
from tools import chain_cases as CC  # noqa: E402

LENS, ALT, MAXOCC, SEED, N = (30000, 1500, 20000), (0, 1, 0), 50, 2026, 280
ctg = CP.Contigs(LENS, alt=ALT)
opt = CP.default_opt(max_occ=MAXOCC)
fwd, cases = CC.make_cases(SEED, N, LENS, MAXOCC)
pac = CP.make_pac(fwd)
chains, cseeds, regs, js, jl, nch, nreg = [], [], [], [], [], [], []
for query, rb, qq, sc in cases:
    fc, fs, fr, fj, _ = CP.fork_read(opt, ctg, pac, query, rb, qq, sc)
    chains.append(fc); cseeds.append(fs); regs.append(fr); js.append(fj[0]); jl.append(fj[1]); nch.append(len(fc)); nreg.append(len(fr))
np.savez_compressed(os.path.join(os.path.dirname(__file__), "chain_golden.npz"), contig_lens=np.array(LENS), contig_alt=np.array(ALT),
                    max_occ=MAXOCC, seed=SEED, n_reads=N, chains=np.concatenate(chains), cseeds=np.concatenate(cseeds),
                    regs=np.concatenate(regs), jobs_short=np.concatenate(js), jobs_long=np.concatenate(jl),
                    n_chains=np.array(nch), n_regs=np.array(nreg))
print("reads", N, "chains", sum(nch), "regs", sum(nreg))

# reads of 760 bases and more: the fork's mem_flt_chained_seeds / mem_seed_sw act on them (src/bwamem.c:774-808,970-990)
LSEED, LN = 2027, 80
fwd, cases = CC.make_long_cases(LSEED, LN, LENS, MAXOCC)
pac = CP.make_pac(fwd)
chains, cseeds, regs, js, jl = [], [], [], [], []
for query, rb, qq, sc in cases:
    fc, fs, fr, fj, _ = CP.fork_read(opt, ctg, pac, query, rb, qq, sc)
    chains.append(fc); cseeds.append(fs); regs.append(fr); js.append(fj[0]); jl.append(fj[1])
np.savez_compressed(os.path.join(os.path.dirname(__file__), "chain_long_golden.npz"), contig_lens=np.array(LENS), contig_alt=np.array(ALT),
                    max_occ=MAXOCC, seed=LSEED, n_reads=LN, chains=np.concatenate(chains), cseeds=np.concatenate(cseeds),
                    regs=np.concatenate(regs), jobs_short=np.concatenate(js), jobs_long=np.concatenate(jl))
print("long reads", LN, "chains", sum(len(c) for c in chains), "seeds", sum(len(c) for c in cseeds), "regs", sum(len(r) for r in regs))
