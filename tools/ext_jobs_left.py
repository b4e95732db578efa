"""Which extension jobs of a C2-like read set the closed-form answer (closed_form_job, csrc/ext_pair_core.cuh) does not take, and how
many of the reference's DP cells each kind holds.  CPU only: the oracle's seeds -> chains -> jobs on N reads of a 5 Mb genome, the
closed-form predicate as tools/synth.closed_form_mask states it, ksw_extend2's cell counter per class.

    python tools/ext_jobs_left.py [n_reads]
"""
import importlib
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("bwa-mem_gpu_b200")
from oracle import oracle_py as O, chain_py as CP      # noqa: E402
from tools import synth                                 # noqa: E402


def classify(q, t, ql, tl, h0, b=4):
    if tl < ql:
        return "target shorter than the query"
    if (q > 3).any() or (t[:ql] > 3).any():
        return "N"
    k = int((q != t[:ql]).sum())
    if k <= 3:
        return "%d substitutions, refused (%s)" % (k, "h0 <= k b" if h0 <= b * k else "spacing / repeat")
    if k > 0.2 * ql:
        return "indel (positions disagree beyond it)"
    return "4+ substitutions"


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
    g = synth.make_genome(5_000_000)
    d = tempfile.mkdtemp(prefix="extjobs")
    prefix = os.path.join(d, "g5m")
    pkg.build_index(g, prefix, sa_intv=16, also_stock_layout=True, n_threads=4)
    oi = O.OracleIndex(prefix + ".bwt", prefix + ".sa")
    base, _, _ = synth.make_reads(g, n, 150, seed=77)
    reads = [base[i] for i in range(n)]
    rf = np.concatenate(reads)
    off = np.arange(n + 1, dtype=np.uint64) * 150
    sd = oi.seed_batch(rf, off, 19, 500, n_threads=4)
    ctg = CP.Contigs((g.size,))
    qq = np.stack([sd["qbeg"], sd["qend"]], axis=1).astype(np.int32)
    kp = O.make_params(w=100, zdrop=100, use_band=1)
    want = CP.oracle_align_batch(CP.default_opt(max_occ=500, w=100), ctg, g, reads, sd["rbeg"], qq, sd["score"], sd["n_seeds"], sd["seed_off"], 0,
                                 kp, n_threads=4)
    jobs, qs, ts = want["jobs"], want["qseq"], want["tseq"]
    jd = {k: np.ascontiguousarray(v) for k, v in dict(qseq=qs, tseq=ts, qoff=jobs["qoff"], toff=jobs["toff"], qlen=jobs["qlen"],
                                                      tlen=jobs["tlen"], h0=jobs["h0"]).items()}
    mask = synth.closed_form_mask(jd, 1, 4, None, 100)
    _, cnt = O.ksw_batch(jd, kp, n_threads=4)
    print("jobs %d (%.2f per read), closed form %d = %.1f %%" % (len(jobs), len(jobs) / n, mask.sum(), 100 * mask.mean()))
    cls = {}
    for j in np.nonzero(~mask)[0]:
        ql, tl, h0 = int(jobs["qlen"][j]), int(jobs["tlen"][j]), int(jobs["h0"][j])
        q = qs[int(jobs["qoff"][j]):int(jobs["qoff"][j]) + ql]
        t = ts[int(jobs["toff"][j]):int(jobs["toff"][j]) + tl]
        cls.setdefault(classify(q, t, ql, tl, h0), []).append(j)
    cells = {}
    for c, idx in cls.items():
        keep = np.zeros(len(jobs), bool)
        keep[idx] = True
        _, cn = O.ksw_batch(synth.subset_jobs(jd, keep), kp, n_threads=4)
        cells[c] = cn["cells"]
    tot = sum(cells.values())
    print("cells: all jobs %d, jobs left to the kernels %d" % (cnt["cells"], tot))
    for c in sorted(cls, key=lambda x: -cells[x]):
        print("%-42s jobs %6d  cells %5.1f %%  (%.0f per job, mean query %.0f)" % (c, len(cls[c]), 100 * cells[c] / tot, cells[c] / len(cls[c]),
                                                                                  jobs["qlen"][cls[c]].mean()))
    oi.close()


if __name__ == "__main__":
    main()
