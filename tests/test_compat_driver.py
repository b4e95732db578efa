"""The reference driver's call sequence, in C++ against include/compat (same names as GPUSeed / GASAL2),
run as a separate process; results vs the CPU oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_driver(out_dir):
    exe = os.path.join(out_dir, "driver_like")
    cmd = ["g++", "-std=c++14", "-O1", "-fpermissive", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "include", "compat"),
           os.path.join(ROOT, "tests", "compat", "driver_like.cpp"), "-o", exe,
           "-L", os.path.join(ROOT, "bwa-mem_gpu_b200"), "-lbwamem_b200", "-Wl,-rpath," + os.path.join(ROOT, "bwa-mem_gpu_b200")]
    subprocess.check_call(cmd)
    return exe


def test_compat_headers_compile_and_link(pkg, tmp_path):
    """CPU: the reference-named API compiles with g++ exactly as the driver's .c files are compiled"""
    assert os.path.exists(build_driver(str(tmp_path)))


@pytest.mark.gpu
def test_driver_like_flow_matches_oracle(pkg, oracle, small_index, tmp_path):
    g, prefix = small_index
    exe = build_driver(str(tmp_path))
    reads, pos, strand = synth.make_reads(g, 1500, 150, seed=8, n_rate=0.002)
    fa = str(tmp_path / "reads.fa")
    synth.reads_to_fasta(reads, pos, strand, fa)
    jobs = synth.make_ext_jobs(2000, w=300, seed=12, qlen_range=(1, 150), h0_range=(1, 150), n_job_frac=0.05)
    jb = str(tmp_path / "jobs.bin")
    with open(jb, "wb") as fh:
        n = jobs["qlen"].size
        fh.write(struct.pack("<I", n))
        fh.write(jobs["qlen"].tobytes()); fh.write(jobs["tlen"].tobytes()); fh.write(jobs["h0"].tobytes())
        for a in range(n):
            fh.write(jobs["qseq"][jobs["qoff"][a]:jobs["qoff"][a] + jobs["qlen"][a]].tobytes())
            fh.write(jobs["tseq"][jobs["toff"][a]:jobs["toff"][a] + jobs["tlen"][a]].tobytes())
    out = str(tmp_path / "out.bin")
    subprocess.check_call([exe, prefix, fa, jb, out, "19"])
    buf = open(out, "rb").read()
    n_reads, n_seeds = struct.unpack_from("<QQ", buf, 0)
    o = 16
    per = np.frombuffer(buf, np.uint32, n_reads, o); o += 4 * n_reads
    pre = np.frombuffer(buf, np.uint32, n_reads, o); o += 4 * n_reads
    rbeg = np.frombuffer(buf, np.uint64, n_seeds, o); o += 8 * n_seeds
    qq = np.frombuffer(buf, np.int32, 2 * n_seeds, o).reshape(-1, 2); o += 8 * n_seeds
    score = np.frombuffer(buf, np.uint32, n_seeds, o); o += 4 * n_seeds
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    want = oi.seed_batch(reads.reshape(-1).copy(), (np.arange(1501) * 150).astype(np.uint64), 19, 0, n_threads=4)   # all occurrences
    assert n_reads == 1500 and n_seeds == want["total"]
    assert (per == want["n_seeds"]).all() and (pre == want["seed_off"].astype(np.uint32)).all()
    assert (rbeg == want["rbeg"]).all() and (qq[:, 0] == want["qbeg"]).all() and (qq[:, 1] == want["qend"]).all()
    assert (score == want["score"]).all()
    # the same driver with re-seeding switched on through the compat layer: stock bwa mem's seed set in GPUSeed's layout
    out2 = str(tmp_path / "out2.bin")
    subprocess.check_call([exe, prefix, fa, jb, out2, "19", "1"])
    b2 = open(out2, "rb").read()
    n_reads2, n_seeds2 = struct.unpack_from("<QQ", b2, 0)
    want2 = oi.seed_batch(reads.reshape(-1).copy(), (np.arange(1501) * 150).astype(np.uint64), 19, 0, n_threads=4, rs=oracle.reseed())
    assert n_reads2 == 1500 and n_seeds2 == want2["total"] > n_seeds
    o2 = 16
    per2 = np.frombuffer(b2, np.uint32, n_reads2, o2); o2 += 8 * n_reads2
    rbeg2 = np.frombuffer(b2, np.uint64, n_seeds2, o2); o2 += 8 * n_seeds2
    qq2 = np.frombuffer(b2, np.int32, 2 * n_seeds2, o2).reshape(-1, 2)
    assert (per2 == want2["n_seeds"]).all() and (rbeg2 == want2["rbeg"]).all() and (qq2[:, 0] == want2["qbeg"]).all() and (qq2[:, 1] == want2["qend"]).all()
    (nj,) = struct.unpack_from("<I", buf, o); o += 4
    sc = np.frombuffer(buf, np.int32, nj, o); o += 4 * nj
    qe = np.frombuffer(buf, np.int32, nj, o); o += 4 * nj
    te = np.frombuffer(buf, np.int32, nj, o)
    # the fork's defaults: no band (opt_ext = 0), zdrop 0, clip 5 -- what decoy_cpu_align computes (src/bwamem.c:1886-1901)
    res6, _ = oracle.ksw_batch(jobs, oracle.make_params(w=300, zdrop=0, use_band=0, pen_clip=5), n_threads=4)
    wsc, wqe, wte = oracle.gasal_triple(res6, jobs["qlen"], 5)
    assert (sc == wsc).all() and (qe == wqe).all() and (te == wte).all()
    oi.close()
