"""Golden vectors for the CIGAR path from the REFERENCE's own ksw_global2 (bwa_index/ksw.c:504) and
bwa_gen_cigar2 (bwa_index/bwa.c:121), reached through oracle/ref_shim.c inside oracle/_ref/libbwaref.so.

Run in the build container only:   python tests/golden/make_global_golden.py
  global_golden.npz   (a) ksw_global2 jobs (tools/synth.make_global_jobs, three settings) with score / n_cigar / CIGAR;
                      (b) bwa_gen_cigar2 calls on a synthetic genome (both strands, w_ in {0, 5, 100}) with score / NM / CIGAR.
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from tools import synth  # noqa: E402

JOB_SETS = [dict(n_jobs=600, qlen_range=(1, 150), seed=31, w_extra=(0, 0)),
            dict(n_jobs=400, qlen_range=(80, 300), seed=32, sub_rate=0.08, indel_rate=0.03, w_extra=(0, 40)),
            dict(n_jobs=300, qlen_range=(20, 120), seed=33, sub_rate=0.2, indel_rate=0.05, w_extra=(0, 6), w_cap=8)]
STRIDE = 96


def ref_global(R, jobs, p):
    n = jobs["qlen"].size
    mat = np.frombuffer(bytes(p.mat), dtype=np.int8).copy()
    score = np.zeros(n, np.int32); nc = np.zeros(n, np.uint32); cig = np.zeros((n, STRIDE), np.uint32)
    for a in range(n):
        k = C.c_int(0)
        q = jobs["qseq"][jobs["qoff"][a]:jobs["qoff"][a] + jobs["qlen"][a]].copy()
        t = jobs["tseq"][jobs["toff"][a]:jobs["toff"][a] + jobs["tlen"][a]].copy()
        row = np.zeros(STRIDE, np.uint32)
        score[a] = R.ref_ksw_global2(q.size, q, t.size, t, mat, p.o_del, p.e_del, p.o_ins, p.e_ins, int(jobs["w"][a]), C.byref(k), row, STRIDE)
        nc[a] = k.value; cig[a] = row
        assert k.value <= STRIDE
    return score, nc, cig


def gen_cigar_cases(g, n, seed):
    rng = np.random.default_rng(seed)
    L = g.size
    reads, pos, strand = synth.make_reads(g, n, 150, seed=seed, sub_rate=0.02, ins_rate=0.004, del_rate=0.004)
    cases = []
    for i in range(n):
        ql = int(rng.integers(30, 151))
        q = reads[i, :ql].copy()
        rlen = ql + int(rng.integers(-3, 4))
        p0 = int(pos[i])
        if strand[i] == 0:
            rb = p0
        else:           # read i is the reverse complement of g[p0 : p0 + 150 + ...): its first ql bases align near the end of that window
            rb = 2 * L - (p0 + 150) + int(rng.integers(-2, 3))
        rb = max(0, min(rb, 2 * L - rlen))
        re = rb + rlen
        if rb < L < re:
            re = L; rb = re - rlen
        cases.append((q, rb, re, int(rng.choice([0, 5, 100]))))
    return cases


def main():
    assert O.have_ref(), "oracle/_ref missing: run oracle/build_ref.sh"
    R = O.ref_lib()
    p = O.make_params()
    out = {}
    for si, kw in enumerate(JOB_SETS):
        jobs = synth.make_global_jobs(**kw)
        score, nc, cig = ref_global(R, jobs, p)
        for k, v in jobs.items():
            out[f"s{si}_{k}"] = v
        out[f"s{si}_score"] = score; out[f"s{si}_n_cigar"] = nc; out[f"s{si}_cigar"] = cig
        print(f"set {si}: {jobs['qlen'].size} jobs, n_cigar max {nc.max()}, w max {jobs['w'].max()}")
    # bwa_gen_cigar2
    g = synth.make_genome(50_000, seed=4711)
    pac = np.zeros((g.size + 3) // 4, np.uint8)
    for sh in range(4):
        part = g[sh::4]
        pac[:part.size] |= (part << ((3 - sh) * 2)).astype(np.uint8)
    mat = np.frombuffer(bytes(p.mat), dtype=np.int8).copy()
    cases = gen_cigar_cases(g, 500, 77)
    qcat = np.concatenate([c[0] for c in cases]); qoff = np.zeros(len(cases) + 1, np.int64); qoff[1:] = np.cumsum([c[0].size for c in cases])
    rb = np.array([c[1] for c in cases], np.int64); re = np.array([c[2] for c in cases], np.int64); w_ = np.array([c[3] for c in cases], np.int32)
    sc = np.zeros(len(cases), np.int32); nm = np.zeros(len(cases), np.int32); nc = np.zeros(len(cases), np.int32); cig = np.zeros((len(cases), STRIDE), np.uint32)
    for i, (q, b, e, w) in enumerate(cases):
        s_, n_ = C.c_int(0), C.c_int(0)
        row = np.zeros(STRIDE, np.uint32)
        nc[i] = R.ref_gen_cigar2(mat, p.o_del, p.e_del, p.o_ins, p.e_ins, w, g.size, pac, q.size, q, int(b), int(e), C.byref(s_), C.byref(n_), row, STRIDE)
        sc[i] = s_.value; nm[i] = n_.value; cig[i] = row
    out.update(gc_genome_len=g.size, gc_genome_seed=4711, gc_query=qcat, gc_qoff=qoff, gc_rb=rb, gc_re=re, gc_w=w_, gc_score=sc, gc_nm=nm,
               gc_n_cigar=nc, gc_cigar=cig)
    print(f"gen_cigar2: {len(cases)} calls, reverse strand {int((rb >= g.size).sum())}, mean NM {nm.mean():.2f}, n_cigar max {nc.max()}")
    np.savez_compressed(os.path.join(HERE, "global_golden.npz"), **out)


if __name__ == "__main__":
    main()
