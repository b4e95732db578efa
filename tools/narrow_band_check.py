"""CPU check behind DESIGN §3's "next step": for extension jobs whose query equals the head of the target except for k = 2 .. 6
substitutions (h0 > k b), does ksw_extend2 with a band of dmax_k + 1 columns return what it returns with w = 100?  The oracle is run
both ways on flanks of exact matches and on adversarial tandem-repeat flanks (tools/synth).  Not used by the product: the kernels
run every job with the caller's band.

    python tools/narrow_band_check.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py as O
from tools import synth
kw = dict(w=100, zdrop=100)
dm = synth.closed_form_eligible(**kw)
tot = ndiff = 0
for seed in range(20):
    jobs = synth.make_repeat_flank_jobs(20000, 7000 + seed, kw) if seed % 2 else synth.make_flank_jobs(20000, seed=7000 + seed, w=100, qlen_range=(1, 200), h0_range=(5, 200))
    n = jobs['qlen'].size
    # substitution-only jobs with k in 2..6, h0 > k b, k b <= zdrop
    keep = np.zeros(n, bool); kk = np.zeros(n, int)
    for j in range(n):
        ql, tl, h0 = int(jobs['qlen'][j]), int(jobs['tlen'][j]), int(jobs['h0'][j])
        if ql == 0 or tl < ql: continue
        q = jobs['qseq'][jobs['qoff'][j]:jobs['qoff'][j] + ql]; t = jobs['tseq'][jobs['toff'][j]:jobs['toff'][j] + ql]
        if (q > 3).any() or (t > 3).any(): continue
        k = int((q != t).sum())
        if 2 <= k <= 6 and h0 > 4 * k: keep[j] = True; kk[j] = k
    want, _ = O.ksw_batch(jobs, O.make_params(**kw), n_threads=4)
    for k in range(2, 7):
        sel = keep & (kk == k)
        if not sel.any(): continue
        sub = synth.subset_jobs(jobs, sel)
        got, _ = O.ksw_batch(sub, O.make_params(w=dm[k] + 1, zdrop=100), n_threads=4)
        d = (got != want[sel]).any(axis=1)
        tot += int(sel.sum()); ndiff += int(d.sum())
        if d.any() and ndiff <= 3:
            a = np.nonzero(d)[0][0]
            print('DIFF k', k, 'narrow', got[a], 'wide', want[sel][a], 'ql', sub['qlen'][a], 'h0', sub['h0'][a])
print('jobs', tot, 'differ', ndiff)
