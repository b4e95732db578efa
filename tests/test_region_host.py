"""The per-read region-finishing source meant for the CUDA kernels (region_core.cuh: sort / dedup / patch, primary marking, mapq),
built for the host, against the fork's golden vectors and the oracle (oracle/region_oracle.c, itself pinned to the reference fork).
Runs on the CPU box."""
import ctypes as C
import importlib.util
import os
import subprocess

import numpy as np
import pytest

from tools import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "host_emul")
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def emul():
    from oracle import chain_py as CP, region_py as RP
    so = os.path.join(HERE, "libregion_host.so")
    srcs = [os.path.join(HERE, "region_host.cpp"), os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc", "region_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(x) > os.path.getmtime(so) for x in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "bwa-mem_gpu_b200", "csrc"), srcs[0], "-o", so])
    L = C.CDLL(so)
    L.region_host_read.restype = C.c_int
    L.region_host_read.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int)]

    def run(opt, ctg, fwd, query, regs, rid_id):
        a = np.ascontiguousarray(regs.copy())
        q = np.ascontiguousarray(query, dtype=np.uint8)
        n_pri = C.c_int(0)
        n = L.region_host_read(C.addressof(opt), ctg.l_pac, ctg.alt.ctypes.data, fwd.ctypes.data, len(q), q.ctypes.data, len(a), a.ctypes.data,
                               rid_id, C.byref(n_pri))
        return a[:n], int(n_pri.value)
    return run


def _maker():
    spec = importlib.util.spec_from_file_location("mkreg", os.path.join(GOLD, "make_region_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return mk


def test_region_core_matches_fork_golden(emul, oracle):
    from oracle import chain_py as CP, region_py as RP
    mk = _maker()
    gold = np.load(os.path.join(GOLD, "region_golden.npz"))
    ctg = CP.Contigs(tuple(int(x) for x in gold["contigs"]), alt=tuple(int(x) for x in gold["alt"]))
    g = synth.make_genome(ctg.l_pac, seed=int(gold["genome_seed"]))
    reads, regs_in, in_off = gold["reads"], gold["regs_in"], gold["in_off"]
    for oi, kw in enumerate(mk.OPTS):
        opt = RP.default_opt(**kw)
        want, woff, wpri = gold[f"out_{oi}"], gold[f"out_off_{oi}"], gold[f"n_pri_{oi}"]
        for i in range(len(reads)):
            a, n_pri = emul(opt, ctg, g, reads[i], regs_in[in_off[i]:in_off[i + 1]], i)
            assert n_pri == wpri[i] and RP.equal(a, want[woff[i]:woff[i + 1]]), (oi, i)


def test_region_core_matches_oracle_fresh_cases(emul, oracle):
    from oracle import region_py as RP
    mk = _maker()
    ctg, g, reads, cases = mk.make_inputs(n_reads=1500, seed=909)
    merged = 0
    for kw in mk.OPTS + (dict(w=5, mask_level=0.9), dict(mapQ_coef_len=100, mapQ_coef_fac=4)):
        opt = RP.default_opt(**kw)
        for i, regs in enumerate(cases):
            a, pa = emul(opt, ctg, g, reads[i], regs, 5000 + i)
            b, pb = RP.oracle_finish(opt, ctg, g, reads[i], regs, 5000 + i)
            assert pa == pb and RP.equal(a, b), (kw, i)
            merged += int((a["n_comp"] > 1).sum())
    assert merged > 100


def _tie_heavy_case(g, ctg, rng, n_regs, L=150):
    """many regions with few distinct ends / scores: the sorts run far beyond their insertion-sort range and meet equal keys all the time"""
    from oracle import region_py as RP
    l = ctg.l_pac
    regs = np.zeros(n_regs, RP.REGION_DT)
    ends = rng.integers(200, l - 200, max(4, n_regs // 40))
    for k in range(n_regs):
        qb = int(rng.integers(0, 60)); qe = int(rng.integers(qb + 30, L + 1))
        re = int(rng.choice(ends)) + (l if k & 1 else 0)
        ln = qe - qb + int(rng.integers(-2, 3))
        rb = re - ln
        if rb < 0 or (rb < l < re):
            rb, re = 300, 300 + ln
        regs[k]["rb"], regs[k]["re"], regs[k]["qb"], regs[k]["qe"] = rb, re, qb, qe
        regs[k]["score"] = regs[k]["truesc"] = int(rng.choice([30, 31, 45, 60]))
        regs[k]["rid"] = RP._rid(ctg, rb, re)
        regs[k]["w"] = int(rng.choice([0, 10, 100])); regs[k]["seedcov"] = int(rng.integers(19, 100)); regs[k]["secondary"] = -1
    return regs


def test_region_core_ties_ambiguous_bases_and_large_sets(emul, oracle):
    """N bases in the read (scored -1 by the patch alignment) and thousands of regions per read with equal keys everywhere;
    the fork itself is compared when oracle/_ref is present"""
    from oracle import chain_py as CP, region_py as RP
    mk = _maker()
    ctg, g, reads, cases = mk.make_inputs(n_reads=400, seed=31)
    rng = np.random.default_rng(8)
    reads = reads.copy()
    reads[rng.random(reads.shape) < 0.02] = 4
    pac = CP.make_pac(g) if CP.have_fork() else None
    opt = RP.default_opt()
    for i, regs in enumerate(cases):
        a, pa = emul(opt, ctg, g, reads[i], regs, i)
        b, pb = RP.oracle_finish(opt, ctg, g, reads[i], regs, i)
        assert pa == pb and RP.equal(a, b), i
        if pac is not None:
            c, pc = RP.fork_finish(opt, ctg, pac, reads[i], regs, i)
            assert pc == pb and RP.equal(c, b), i
    for n_regs in (17, 64, 700, 3000):
        regs = _tie_heavy_case(g, ctg, rng, n_regs)
        for kw in (dict(), dict(max_chain_gap=50, mask_level_redun=0.5)):
            opt = RP.default_opt(**kw)
            a, pa = emul(opt, ctg, g, reads[0], regs, 77)
            b, pb = RP.oracle_finish(opt, ctg, g, reads[0], regs, 77)
            assert pa == pb and RP.equal(a, b), (n_regs, kw)
            if pac is not None:
                c, pc = RP.fork_finish(opt, ctg, pac, reads[0], regs, 77)
                assert pc == pb and RP.equal(c, b), (n_regs, kw)


def test_region_sorts_alone_incl_comb_sort_fallback(emul, oracle):
    """ks_introsort's comb-sort fallback only runs when quicksort degenerates, which the stage tests never cause: the four sort
    instances (and the comb sort by itself) of the device source, of the oracle and -- when oracle/_ref is present -- of the fork, on
    arrays full of equal keys; seedlen0 carries each record's identity so that the permutation itself is compared"""
    import ctypes as C
    from oracle import chain_py as CP, region_py as RP, oracle_py as O
    L = C.CDLL(os.path.join(HERE, "libregion_host.so"))
    L.region_host_sort.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    OL = O.lib()
    OL.region_combsort.argtypes = OL.region_introsort.argtypes = [C.c_int, C.c_int, C.c_void_p]
    OL.region_combsort.restype = OL.region_introsort.restype = None
    FL = None
    if CP.have_fork():
        FL = CP.fork_lib()
        FL.fork_sort_regs.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
        FL.fork_sort_regs.restype = None
    rng = np.random.default_rng(12)
    for n in (1, 2, 3, 9, 10, 11, 16, 17, 18, 40, 129, 1000, 5000):
        for distinct in (2, 7, 10 ** 6):
            a = np.zeros(n, RP.REGION_DT)
            a["re"] = rng.integers(0, distinct, n); a["rb"] = rng.integers(0, distinct, n); a["qb"] = rng.integers(0, distinct, n)
            a["score"] = rng.integers(0, distinct, n); a["is_alt"] = rng.integers(0, 2, n); a["hash"] = rng.integers(0, distinct, n).astype(np.uint64)
            a["seedlen0"] = np.arange(n)
            for comb in (0, 1):
                for which in range(4):
                    x, y = a.copy(), a.copy()
                    L.region_host_sort(comb, which, n, x.ctypes.data)
                    (OL.region_combsort if comb else OL.region_introsort)(which, n, y.ctypes.data)
                    assert (x["seedlen0"] == y["seedlen0"]).all(), (n, distinct, comb, which)
                    if FL is not None:
                        z = a.copy()
                        FL.fork_sort_regs(comb, which, n, z.ctypes.data)
                        assert (x["seedlen0"] == z["seedlen0"]).all(), (n, distinct, comb, which, "fork")


def test_region_core_on_regions_of_real_reads(emul, oracle, small_index):
    """the stage's real input distribution: regions that seeding -> chaining -> extension (the oracle pipeline, itself pinned to the
    reference) produce for reads with errors on a genome with repeats, re-seeding on (more chains, more overlapping regions);
    device source == oracle (== the fork when oracle/_ref is present)"""
    from oracle import chain_py as CP, region_py as RP
    g, prefix = small_index
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    reads, _, _ = synth.make_reads(g, 1500, 150, seed=77, sub_rate=0.03, ins_rate=0.004, del_rate=0.004)
    n, L = reads.shape
    rf = reads.reshape(-1).copy()
    off = (np.arange(n + 1) * L).astype(np.uint64)
    sd = oi.seed_batch(rf, off, 19, 500, n_threads=4, rs=oracle.reseed())
    oi.close()
    ctg = CP.Contigs((g.size,))
    kp = oracle.make_params(w=100, zdrop=100, use_band=1)
    qq = np.stack([sd["qbeg"], sd["qend"]], axis=1).astype(np.int32)
    want = CP.oracle_align_batch(CP.default_opt(max_occ=500, w=100), ctg, g, list(reads), sd["rbeg"], qq, sd["score"], sd["n_seeds"], sd["seed_off"], 0, kp)
    regs, aln, nr = want["regs"], want["aln"], want["n_regions"]
    a = np.zeros(len(regs), RP.REGION_DT)
    for k in ("rb", "re", "qb", "qe", "score", "truesc"):
        a[k] = aln[k]
    for k in ("rid", "w", "seedcov", "seedlen0", "frac_rep"):
        a[k] = regs[k]
    a["secondary"] = a["secondary_all"] = -1
    ro = np.concatenate([[0], np.cumsum(nr)]).astype(np.int64)
    pac = CP.make_pac(g) if CP.have_fork() else None
    opt = RP.default_opt()
    multi = dropped = secondary = 0
    for i in range(n):
        mine = a[ro[i]:ro[i + 1]]
        keep = mine[(mine["qe"] > mine["qb"]) & (mine["re"] > mine["rb"])]
        x, px = emul(opt, ctg, g, reads[i], keep, i)
        y, py = RP.oracle_finish(opt, ctg, g, reads[i], keep, i)
        assert px == py and RP.equal(x, y), i
        if pac is not None:
            z, pz = RP.fork_finish(opt, ctg, pac, reads[i], keep, i)
            assert pz == py and RP.equal(z, y), i
        multi += int(len(keep) > 1); dropped += len(keep) - len(y); secondary += int((y["secondary"] >= 0).sum())
    assert multi > 50 and dropped > 20 and secondary > 10
