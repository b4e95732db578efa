#!/usr/bin/env python
"""Randomised stress of the CIGAR kernels against the oracle (GPU box): many seeds, scoring schemes, band classes, both the
register-resident and the ring kernel.  Prints one line per configuration; exits non-zero on the first mismatch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
from oracle import oracle_py as O
from tools import synth

pkg = ge.load_package(); pkg.build()
rng = np.random.default_rng(12345)
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 24):
    ring = it % 3 == 2
    if ring:
        os.environ["BWA_B200_GLOBAL_RING"] = "1"
    else:
        os.environ.pop("BWA_B200_GLOBAL_RING", None)
    kw = dict(a=int(rng.integers(1, 4)), b=int(rng.integers(1, 7)), o_del=int(rng.integers(0, 9)), e_del=int(rng.integers(1, 4)),
              o_ins=int(rng.integers(0, 9)), e_ins=int(rng.integers(1, 4)))
    jobs = synth.make_global_jobs(6000, qlen_range=(1, int(rng.choice([60, 150, 300]))), seed=1000 + it, sub_rate=float(rng.choice([0.0, 0.03, 0.15])),
                                  indel_rate=float(rng.choice([0.0, 0.01, 0.06])), w_extra=(0, int(rng.choice([0, 4, 12, 40, 110]))), w_cap=127,
                                  n_frac=float(rng.choice([0.0, 0.05])))
    cg = pkg.Cigar(0)
    got = cg.global_host(jobs, pkg.ext_params(**kw))
    want = O.global_batch(jobs, O.make_params(**kw), cig_stride=700, n_threads=8)
    ok = (got["score"] == want["score"]).all() and (got["nm"] == want["nm"]).all() and (got["n_cigar"] == want["n_cigar"]).all() and cg.last_cells == want["cells"]
    if ok:
        for a in range(jobs["qlen"].size):
            m, o = int(got["n_cigar"][a]), int(got["cigar_off"][a])
            if not (got["cigar"][o:o + m] == want["cigar"][a, :m]).all():
                ok = False; break
    cg.destroy()
    print(f"it {it:2d} ring={int(ring)} {kw} w_max={int(jobs['w'].max())} n_cigar_max={int(got['n_cigar'].max())} {'ok' if ok else 'MISMATCH'}", flush=True)
    bad += not ok
sys.exit(1 if bad else 0)
