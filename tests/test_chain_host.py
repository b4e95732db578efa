"""The per-read chaining / job-construction source of the CUDA kernels (chain_core.cuh), built for the host,
against the oracle (oracle/chain_oracle.c, itself pinned to the reference fork).  Runs on the CPU box; the GPU
parity tests run the same source in the kernels."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import chain_py as CP
from tools import chain_cases as CC

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "host_emul")

REGION_DT = np.dtype([("rb", "<i8"), ("re", "<i8"), ("rb_est", "<i8"), ("re_est", "<i8"), ("target_seed_begin", "<i8"),
                      ("qb", "<i4"), ("qe", "<i4"), ("score", "<i4"), ("truesc", "<i4"),
                      ("qb_est", "<i4"), ("qe_est", "<i4"), ("rid", "<i4"), ("align_sides", "<i4"), ("where_is_long", "<i4"),
                      ("query_seed_begin", "<i4"), ("seedlen0", "<i4"), ("seedcov", "<i4"), ("w", "<i4"), ("frac_rep", "<f4"),
                      ("left_tlen", "<i4"), ("right_tlen", "<i4"), ("job_short", "<i4"), ("job_long", "<i4")], align=True)
REG_FIELDS = ["rb_est", "re_est", "target_seed_begin", "qb_est", "qe_est", "rid", "align_sides", "where_is_long",
              "query_seed_begin", "seedlen0", "seedcov", "w", "frac_rep"]


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "libchain_host.so")
    srcs = [os.path.join(HERE, "chain_host.cpp"), os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc", "chain_core.cuh"),
            os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc", "sw_core.cuh"), os.path.join(ROOT, "include", "bwamem_b200.h")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc"), srcs[0], "-o", so])
    L = C.CDLL(so)
    L.chain_host_read.restype = C.c_int
    L.chain_host_read.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_uint32] + [C.c_void_p] * 3 + \
                                 [C.c_int] + [C.c_void_p] * 8

    def run(opt, ctg, l_query, rb, qq, sc, layout_all, short3=None, long3=None, pac2=None, rd4=None):
        n = max(len(rb), 1)
        rb = np.ascontiguousarray(rb, np.uint64); qq = np.ascontiguousarray(qq, np.int32); sc = np.ascontiguousarray(sc, np.uint32)
        chains = np.zeros(n, CP.CHAIN_DT); cs = np.zeros(n, CP.CSEED_DT); regs = np.zeros(n, REGION_DT); counts = np.zeros(3, np.int32)
        s3 = np.ascontiguousarray(short3, np.int32) if short3 is not None else None
        l3 = np.ascontiguousarray(long3, np.int32) if long3 is not None else None
        nc = L.chain_host_read(C.addressof(opt), ctg.n, ctg.off.ctypes.data, ctg.len.ctypes.data, ctg.alt.ctypes.data, ctg.l_pac, l_query,
                               len(rb), rb.ctypes.data, qq.ctypes.data, sc.ctypes.data, int(layout_all), chains.ctypes.data, cs.ctypes.data,
                               regs.ctypes.data, counts.ctypes.data, s3.ctypes.data if s3 is not None else None,
                               l3.ctypes.data if l3 is not None else None, pac2.ctypes.data if pac2 is not None else None,
                               rd4.ctypes.data if rd4 is not None else None)
        assert nc >= 0, nc
        chains = chains[:nc]
        return chains, cs[:int(chains["n"].sum()) if nc else 0], regs[:counts[0]], counts
    L.chain_host_cut.restype = C.c_int
    L.chain_host_cut.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 5
    run.lib = L
    return run


def pack2(fwd):
    """2-bit forward reference as the device holds it: 16 bases per word, base 0 in the top two bits, one spare word"""
    n = len(fwd)
    f = np.zeros((n + 15) // 16 * 16 + 16, np.uint32)
    f[:n] = fwd & 3
    f = f.reshape(-1, 16)
    return np.ascontiguousarray((f << (2 * (15 - np.arange(16, dtype=np.uint32)))[None, :]).sum(axis=1).astype(np.uint32))


def pack4(codes):
    """4-bit packing of one sequence: 8 bases per word, base 0 in the high nibble; a partial last word is N-padded"""
    n = len(codes)
    f = np.full((n + 7) // 8 * 8, 4, np.uint32)
    f[:n] = codes
    f = f.reshape(-1, 8)
    return np.ascontiguousarray((f << (4 * (7 - np.arange(8, dtype=np.uint32)))[None, :]).sum(axis=1).astype(np.uint32))


def host_cut(emul, pac2, l_pac, query, regs, batch):
    rd = pack4(query)
    rd[-1] &= np.uint32((0xffffffff << (4 * ((-len(query)) % 8))) & 0xffffffff)      # the read's own padding is zero, as pack_codes leaves it
    cap = max(1, len(regs)) * 2
    lens3 = np.zeros(3 * cap, np.uint32); qw = np.zeros(cap * 64, np.uint32); tw = np.zeros(cap * 128, np.uint32)
    nq = C.c_uint32(0); nt = C.c_uint32(0)
    regs = np.ascontiguousarray(regs)
    n = emul.lib.chain_host_cut(pac2.ctypes.data, len(pac2), l_pac, rd.ctypes.data, len(rd), len(query), len(regs), regs.ctypes.data, batch,
                                lens3.ctypes.data, qw.ctypes.data, tw.ctypes.data, C.addressof(nq), C.addressof(nt))
    return lens3[:3 * n].reshape(-1, 3), qw[:nq.value], tw[:nt.value]


def test_region_struct_layout():
    assert REGION_DT.itemsize == 112 and CP.CHAIN_DT.itemsize == 40 and CP.CSEED_DT.itemsize == 24


@pytest.mark.parametrize("lens,max_occ,seed", [((30000, 1500, 20000), 50, 21), ((30000, 1500, 20000), 50, 22), ((40000,), 500, 23)])
def test_chain_source_matches_oracle(emul, lens, max_occ, seed):
    ctg = CP.Contigs(lens, alt=[0, 1, 0][:len(lens)])
    opt = CP.default_opt(max_occ=max_occ)
    fwd, cases = CC.make_cases(seed, 300, lens, max_occ)
    pac2 = pack2(fwd)
    rng = np.random.default_rng(seed)
    n_multi = 0
    for query, rb, qq, sc in cases:
        for layout_all in (1, 0):
            a = (rb, qq, sc) if layout_all else CC.to_compact(rb, qq, sc, max_occ)
            oc, osd = CP.oracle_chains(opt, ctg, len(query), a[0], a[1], a[2], layout_all)
            oregs, ojobs, oseqs = CP.oracle_chain2aln(opt, ctg, fwd, query, oc, osd)
            # made-up extension results: the region arithmetic must agree too
            s3 = rng.integers(0, 200, size=(max(len(ojobs[0]), 1), 3)).astype(np.int32)
            l3 = rng.integers(0, 200, size=(max(len(ojobs[1]), 1), 3)).astype(np.int32)
            chains, cs, regs, counts = emul(opt, ctg, len(query), a[0], a[1], a[2], layout_all, s3, l3)
            assert chains.tobytes() == oc.tobytes() and cs.tobytes() == osd.tobytes()
            assert len(regs) == len(oregs) and counts[1] == len(ojobs[0]) and counts[2] == len(ojobs[1])
            for f in REG_FIELDS:
                assert (regs[f] == oregs[f]).all(), f
            # jobs: lengths and order
            ks = regs["job_short"] >= 0
            assert (regs["job_short"][ks] == np.arange(ks.sum())).all() and (regs["job_long"][regs["job_long"] >= 0] == np.arange(counts[2])).all()
            lq, rq = regs["query_seed_begin"], len(query) - regs["query_seed_begin"] - regs["seedlen0"]
            left_is_long = regs["where_is_long"] == 0
            two = regs["align_sides"] == 2
            any_ = regs["align_sides"] > 0
            long_q = np.where(left_is_long, lq, rq)[any_]; long_t = np.where(left_is_long, regs["left_tlen"], regs["right_tlen"])[any_]
            short_q = np.where(left_is_long, rq, lq)[two]; short_t = np.where(left_is_long, regs["right_tlen"], regs["left_tlen"])[two]
            assert (long_q == ojobs[1]["qlen"]).all() and (long_t == ojobs[1]["tlen"]).all()
            assert (short_q == ojobs[0]["qlen"]).all() and (short_t == ojobs[0]["tlen"]).all()
            assert (regs["seedlen0"][any_] == ojobs[1]["h0"]).all()
            # job sequences, cut a word at a time from the 2-bit reference and the packed read
            for batch in (0, 1):
                lens3, qw, tw = host_cut(emul, pac2, ctg.l_pac, query, regs, batch)
                oj = ojobs[batch]
                assert (lens3[:, 0] == oj["qlen"]).all() and (lens3[:, 1] == oj["tlen"]).all() and (lens3[:, 2] == oj["h0"]).all()
                assert qw.tobytes() == pack4(oseqs[batch][0]).tobytes() and tw.tobytes() == pack4(oseqs[batch][1]).tobytes()
            aln = CP.oracle_regs_finish(len(query), oregs, s3, l3)
            for f in ("rb", "re", "qb", "qe", "score", "truesc"):
                assert (regs[f] == aln[f]).all(), f
            n_multi += len(oc) > 9
    assert n_multi > 10


@pytest.mark.parametrize("lens,max_occ,seed,min_chain_weight", [((30000, 1500, 20000), 50, 31, 0), ((40000,), 500, 32, 0), ((30000, 1500, 20000), 50, 33, 30)])
def test_long_reads_seed_filter_matches_oracle(emul, lens, max_occ, seed, min_chain_weight):
    # reads mem_flt_chained_seeds acts on: seed_sw (the score of ksw_i16 around every chain seed) + flt_seeds_apply of the device source
    ctg = CP.Contigs(lens, alt=[0, 1, 0][:len(lens)])
    opt = CP.default_opt(max_occ=max_occ)
    opt.min_chain_weight = min_chain_weight
    fwd, cases = CC.make_long_cases(seed, 40, lens, max_occ)
    pac2 = pack2(fwd)
    dropped = moved = short_circuit = 0
    for query, rb, qq, sc in cases:
        rd4 = pack4(query)
        for layout_all in (1, 0):
            a = (rb, qq, sc) if layout_all else CC.to_compact(rb, qq, sc, max_occ)
            oc, osd = CP.oracle_chains(opt, ctg, len(query), a[0], a[1], a[2], layout_all, fwd, query)
            oregs, ojobs, _ = CP.oracle_chain2aln(opt, ctg, fwd, query, oc, osd)
            chains, cs, regs, counts = emul(opt, ctg, len(query), a[0], a[1], a[2], layout_all, None, None, pac2, rd4)
            assert chains.tobytes() == oc.tobytes() and cs.tobytes() == osd.tobytes()
            assert len(regs) == len(oregs) and counts[1] == len(ojobs[0]) and counts[2] == len(ojobs[1])
            for f in REG_FIELDS:
                assert (regs[f] == oregs[f]).all(), f
        # the unfiltered chains of the same read: the filter must have done something over the set
        L = CP._bind_oracle()
        n = len(rb)
        ch0 = np.zeros(max(n, 1), CP.CHAIN_DT); cs0 = np.zeros(max(n, 1), CP.CSEED_DT); nc0 = C.c_int32(0)
        rbc = np.ascontiguousarray(rb, np.uint64); qqc = np.ascontiguousarray(qq, np.int32); scc = np.ascontiguousarray(sc, np.uint32)
        L.chain_oracle_read_any(C.byref(opt), ctg.l_pac, ctg.n, ctg.off.ctypes.data, ctg.len.ctypes.data, ctg.alt.ctypes.data, len(query), n,
                                rbc.ctypes.data, qqc.ctypes.data, scc.ctypes.data, 1, C.byref(nc0), ch0.ctypes.data, cs0.ctypes.data)
        dropped += int(ch0[:nc0.value]["n"].sum()) - len(osd)
        moved += int((osd["score"] != osd["len"]).sum())
        short_circuit += int((osd["len"] >= 200).sum())
    assert dropped > 20 and moved > 50 and short_circuit > 0
