"""Generates the golden vectors under tests/golden/ from the REFERENCE's own CPU code.

Run in the build container only (needs /root/reference compiled into oracle/_ref by
oracle/build_ref.sh):   python tests/golden/make_golden.py
Outputs are committed; the GPU box and the CPU test-suite only read them.

  index_hashes.npz  sha256 of the .bwt/.sa/.bwt128 the reference's `bwa index` writes
                    (two passes of build_index.sh) for seeded synthetic genomes
  seed_golden.npz   reads + pass-1 SMEMs from bwt_smem1 (bwa_index/bwt.c) and located seeds from
                    bwt_sa with the mem_chain sampling rule (bwa_index/bwamem.c:278-283)
  ksw_*.npz         extension jobs + the six outputs of ksw_extend2 (stock bwa_index/ksw.c:380 for
                    the banded sets, the fork's src/ksw.c:864 with opt_ext=0 for the unbanded one)
"""
import ctypes as C
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402
from tools import synth  # noqa: E402


def sha(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest()


def main():
    assert O.have_ref(), "oracle/_ref missing: run oracle/build_ref.sh"
    tmp = tempfile.mkdtemp()
    # ---- index hashes
    rows = []
    for n, rep, seed, intv in [(1000, 0, 11, 16), (33333, 1, 12, 8), (200000, 1, 13, 16), (777777, 1, 14, 32), (2000000, 0, 15, 16)]:
        g = synth.make_genome(n, seed=seed, repeats=bool(rep))
        fa = os.path.join(tmp, f"g{n}.fa")
        synth.genome_to_fasta(g, fa)
        prefix = os.path.join(tmp, f"ref{n}")
        O.ref_build_index(fa, prefix, intv)
        rows.append((str(n), str(rep), str(seed), str(intv), sha(prefix + ".bwt"), sha(prefix + ".sa"), sha(prefix + ".bwt128")))
    np.savez_compressed(os.path.join(HERE, "index_hashes.npz"), rows=np.array(rows))

    # ---- seeding golden
    genome_len, genome_seed, sa_intv, max_occ = 60000, 4242, 16, 10
    g = synth.make_repeat_genome(genome_len, seed=genome_seed)
    fa = os.path.join(tmp, "seed.fa")
    synth.genome_to_fasta(g, fa)
    prefix = os.path.join(tmp, "seed")
    O.ref_build_index(fa, prefix, sa_intv)
    R = O.ref_lib()
    h = R.ref_load((prefix + ".bwt128").encode(), (prefix + ".sa").encode())
    assert h
    r1, _, _ = synth.make_reads(g, 300, 150, seed=99)
    r2, _, _ = synth.make_reads(g, 150, 150, seed=100, n_rate=0.01, sub_rate=0.03)
    r3, _, _ = synth.make_reads(g, 50, 150, seed=101, sub_rate=0.0, ins_rate=0.0, del_rate=0.0)
    rnd = np.random.default_rng(5).integers(0, 4, size=(20, 150), dtype=np.uint8)      # unrelated to the genome
    poly = np.zeros((4, 150), np.uint8); poly[1] = 3; poly[2] = 4; poly[3, ::2] = 1     # poly-A, poly-T, all-N, AC repeat
    reads = np.concatenate([r1, r2, r3, rnd, poly]).astype(np.uint8)
    n_smems, qb, qe, kk, ss, n_seeds, rbeg, score = [], [], [], [], [], [], [], []
    buf = np.zeros(5 * 400, np.uint64)
    cnt = C.c_int()
    for r in range(reads.shape[0]):
        q = reads[r].copy()
        x, m, ns = 0, 0, 0
        while x < q.size:
            if q[x] < 4:
                x = R.ref_smem1(h, q.size, q, x, 1, buf, C.byref(cnt))
                for i in range(cnt.value):
                    k_, _, s_, b_, e_ = (int(v) for v in buf[5 * i:5 * i + 5])
                    if e_ - b_ < 19:
                        continue
                    m += 1
                    qb.append(b_); qe.append(e_); kk.append(k_); ss.append(s_)
                    step = s_ // max_occ if s_ > max_occ else 1
                    k = c = 0
                    while k < s_ and c < max_occ:
                        rbeg.append(int(R.ref_sa(h, k_ + k)))
                        score.append(s_ if c == 0 else 0)
                        k += step; c += 1; ns += 1
            else:
                x += 1
        n_smems.append(m); n_seeds.append(ns)
    np.savez_compressed(os.path.join(HERE, "seed_golden.npz"), genome_len=genome_len, genome_seed=genome_seed,
                        sa_intv=sa_intv, max_occ=max_occ, reads=reads,
                        n_smems=np.array(n_smems, np.uint32), qbeg=np.array(qb, np.int32), qend=np.array(qe, np.int32),
                        k=np.array(kk, np.uint64), s=np.array(ss, np.uint64), n_seeds=np.array(n_seeds, np.uint32),
                        rbeg=np.array(rbeg, np.uint64), score=np.array(score, np.uint32))
    print("seed golden:", sum(n_smems), "SMEMs", len(rbeg), "seeds")

    # ---- extension golden
    F = O.fork_lib()
    sets = {
        "ksw_band":   (dict(w=100, zdrop=100), dict(n_jobs=400, qlen_range=(1, 160), seed=21, n_job_frac=0.1, h0_range=(1, 150))),
        "ksw_noband": (dict(w=300, zdrop=0, use_band=0), dict(n_jobs=300, qlen_range=(1, 160), seed=22, n_job_frac=0.1, h0_range=(1, 150))),
        "ksw_narrow": (dict(w=5, zdrop=20), dict(n_jobs=400, qlen_range=(1, 120), seed=23, sub_rate=0.15, indel_rate=0.05, n_job_frac=0.2, h0_range=(1, 60))),
        "ksw_asym":   (dict(w=40, zdrop=60, o_del=4, e_del=2, o_ins=7, e_ins=1, a=2, b=3, end_bonus=0, pen_clip=3),
                       dict(n_jobs=300, qlen_range=(1, 100), seed=24, sub_rate=0.1, indel_rate=0.03, h0_range=(1, 100))),
    }
    for name, (kw, jkw) in sets.items():
        jobs = synth.make_ext_jobs(w=kw["w"], **jkw)
        p = O.make_params(**kw)
        mat = np.frombuffer(bytes(p.mat), dtype=np.int8).copy()
        n = jobs["qlen"].size
        res6 = np.zeros((n, 6), np.int32)
        out = np.zeros(6, np.int32)
        for a in range(n):
            q = jobs["qseq"][jobs["qoff"][a]:jobs["qoff"][a] + jobs["qlen"][a]].copy()
            t = jobs["tseq"][jobs["toff"][a]:jobs["toff"][a] + jobs["tlen"][a]].copy()
            if p.use_band:
                R.ref_ksw_extend2(q.size, q, t.size, t, mat, p.o_del, p.e_del, p.o_ins, p.e_ins, p.w, p.end_bonus, p.zdrop,
                                  int(jobs["h0"][a]), out)
                chk = out.copy()
            F.fork_ksw_extend2(q.size, q, t.size, t, mat, p.o_del, p.e_del, p.o_ins, p.e_ins, p.w, p.end_bonus, p.zdrop,
                               int(jobs["h0"][a]), p.use_band, out)
            if p.use_band:
                assert (chk == out).all()
            res6[a] = out
        np.savez_compressed(os.path.join(HERE, name + ".npz"), res6=res6, param_names=np.array(list(kw.keys())),
                            param_values=np.array(list(kw.values())), **jobs)
        print(name, n, "jobs")


if __name__ == "__main__":
    main()
