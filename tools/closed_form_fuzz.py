"""Offline fuzz of the closed-form answer (closed_form_job, csrc/ext_pair_core.cuh) against the oracle's ksw_extend2 on adversarial
jobs (tools/synth.make_repeat_flank_jobs: tandem repeats with a break or two, substitutions spaced around the dmax_k + 2 boundary),
nine parameter sets, both sequence forms (bytes and packed words must agree).  CPU only; needs tests/host_emul/libextpair_host.so
(built by the test suite's `emul` fixture).

    python tools/closed_form_fuzz.py [jobs_per_batch=20000] [batches_per_parameter_set=2]
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bwa-mem_gpu_b200")
from oracle import oracle_py as O      # noqa: E402
from tools import synth                # noqa: E402

if __name__ == '__main__':
    L = C.CDLL(os.path.join(ROOT, 'tests', 'host_emul', 'libextpair_host.so'))
    L.ext_closed_form_host.restype = C.c_longlong
    L.ext_closed_form_host.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 9
    KWS = [dict(w=100, zdrop=100), dict(w=100, zdrop=0), dict(w=30, zdrop=40), dict(w=16, zdrop=100), dict(w=21, zdrop=0),
           dict(w=40, zdrop=60, a=2, b=3), dict(w=40, zdrop=0, a=2, b=5, o_del=7, e_del=2, o_ins=8, e_ins=1), dict(w=60, zdrop=100, o_del=4, e_del=2, o_ins=9, e_ins=3),
           dict(w=300, zdrop=0, use_band=0)]
    nper = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    for ki, kw in enumerate(KWS):
        ep = pkg.ext_params(**kw)
        tot = {}
        for rep in range(int(sys.argv[2]) if len(sys.argv) > 2 else 2):
            jobs = synth.make_repeat_flank_jobs(nper, 1000 + 17 * ki + rep, kw)
            n = jobs['qlen'].size
            res = np.zeros((n, 6), np.int32); flags = np.zeros(n, np.uint8)
            taken = L.ext_closed_form_host(C.addressof(ep), n, jobs["qseq"].ctypes.data, jobs["qoff"].ctypes.data, jobs["qlen"].ctypes.data,
                                           jobs["tseq"].ctypes.data, jobs["toff"].ctypes.data, jobs["tlen"].ctypes.data, jobs["h0"].ctypes.data,
                                           res.ctypes.data, flags.ctypes.data)
            assert taken >= 0, (kw, taken)
            want, _ = O.ksw_batch(jobs, O.make_params(**kw), n_threads=4)
            got = flags != 0
            bad = np.nonzero((res != want).any(axis=1) & got)[0]
            for a in np.nonzero(got)[0]:
                ql = int(jobs['qlen'][a]); q = jobs['qseq'][jobs['qoff'][a]:jobs['qoff'][a] + ql]; t = jobs['tseq'][jobs['toff'][a]:jobs['toff'][a] + ql]
                kk = int((q != t).sum()); tot[kk] = tot.get(kk, 0) + 1
            if bad.size:
                a = bad[0]
                ql = int(jobs['qlen'][a]); tl = int(jobs['tlen'][a])
                print('BAD', kw, bad.size, 'job', a, 'h0', jobs['h0'][a], 'res', res[a], 'want', want[a])
                print('q', ''.join('ACGT'[x] for x in jobs['qseq'][jobs['qoff'][a]:jobs['qoff'][a] + ql]))
                print('t', ''.join('ACGT'[x] for x in jobs['tseq'][jobs['toff'][a]:jobs['toff'][a] + tl]))
                sys.exit(1)
        print(kw, 'taken by k:', dict(sorted(tot.items())))
