"""GPU parity of the seeds -> chains -> extension jobs -> extension -> regions stage (bwa_b200_align_*), called through
the C ABI, against the CPU oracle of the reference fork's mem_chain / mem_chain_flt / mem_chain2aln / ksw_extend2 /
result gathering (oracle/chain_oracle.c, pinned to the fork's own code by tests/test_chain_oracle.py).  Bit-exact."""
import numpy as np
import pytest

from oracle import chain_py as CP
from tools import chain_cases as CC
from tools import synth

pytestmark = pytest.mark.gpu

REG_FIELDS = ["rb_est", "re_est", "target_seed_begin", "qb_est", "qe_est", "rid", "align_sides", "where_is_long",
              "query_seed_begin", "seedlen0", "seedcov", "w", "frac_rep"]
ALN_FIELDS = ["rb", "re", "qb", "qe", "score", "truesc"]


def pack4(codes):
    n = len(codes)
    f = np.full((n + 7) // 8 * 8, 4, np.uint32)
    f[:n] = codes
    f = f.reshape(-1, 8)
    return np.ascontiguousarray((f << (4 * (7 - np.arange(8, dtype=np.uint32)))[None, :]).sum(axis=1).astype(np.uint32))


def flat(reads_list):
    lens = np.array([len(r) for r in reads_list], np.uint64)
    off = np.zeros(len(reads_list) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    return np.concatenate(reads_list).astype(np.uint8), off


def check_batch(got, want, detail=True):
    assert (got["n_regions"] == want["n_regions"]).all()
    assert (got["region_off"] == np.concatenate([[0], np.cumsum(want["n_regions"])[:-1]])).all()
    assert len(got["regions"]) == len(want["regs"])
    for f in REG_FIELDS:
        assert (got["regions"][f] == want["regs"][f]).all(), f
    for f in ALN_FIELDS:
        assert (got["regions"][f] == want["aln"][f]).all(), f
    if not detail:
        return
    assert (got["n_chains"] == want["n_chains"]).all()
    assert got["chains"].tobytes() == want["chains"].tobytes()
    assert got["chain_seeds"].tobytes() == want["chain_seeds"].tobytes()
    assert got["n_jobs_short"] == want["n_jobs_short"] and got["n_jobs_long"] == want["n_jobs_long"]
    for f in ("qoff", "qlen", "toff", "tlen", "h0"):
        assert (got["jobs"][f] == want["jobs"][f]).all(), f
    assert got["qpacked"].tobytes() == pack4(want["qseq"]).tobytes()
    assert got["tpacked"].tobytes() == pack4(want["tseq"]).tobytes()
    assert (got["job_res"] == want["job_res"]).all()
    # region -> job links: the k-th region with a LONG job owns LONG job k, likewise SHORT
    jl, js = got["regions"]["job_long"], got["regions"]["job_short"]
    assert (jl[jl >= 0] == np.arange((jl >= 0).sum())).all() and (js[js >= 0] == np.arange((js >= 0).sum())).all()


@pytest.fixture(scope="module")
def gpu(pkg):
    assert pkg.lib().bwa_b200_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    return pkg


@pytest.fixture(scope="module")
def case_index(gpu, tmp_path_factory):
    """three contigs (one ALT), index built by the product's host builder, reference attached"""
    lens = (30000, 1500, 20000)
    fwd, cases = CC.make_cases(31, 400, lens, 50)
    prefix = str(tmp_path_factory.mktemp("cidx") / "g")
    gpu.build_index(fwd, prefix, sa_intv=16, n_threads=4)
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(fwd)
    yield lens, fwd, cases, idx
    idx.free()


@pytest.mark.parametrize("layout_all", [1, 0])
@pytest.mark.parametrize("ext", [dict(w=100, zdrop=100, use_band=1), dict(w=300, zdrop=0, use_band=0)])
def test_align_given_seeds_matches_oracle(gpu, oracle, case_index, layout_all, ext):
    lens, fwd, cases, idx = case_index
    ctg = CP.Contigs(lens, alt=[0, 1, 0])
    opt = CP.default_opt(max_occ=50, w=ext["w"])
    reads = [c[0] for c in cases]
    seeds = [(c[1], c[2], c[3]) if layout_all else CC.to_compact(c[1], c[2], c[3], 50) for c in cases]
    n_seeds = np.array([len(s[0]) for s in seeds], np.uint32)
    seed_off = np.concatenate([[0], np.cumsum(n_seeds)[:-1]]).astype(np.uint64)
    rbeg = np.concatenate([s[0] for s in seeds]); qq = np.concatenate([s[1].reshape(-1, 2) for s in seeds]); score = np.concatenate([s[2] for s in seeds])
    kp = oracle.make_params(w=ext["w"], zdrop=ext["zdrop"], use_band=ext["use_band"])
    want = CP.oracle_align_batch(opt, ctg, fwd, reads, rbeg, qq, score, n_seeds, seed_off, layout_all, kp)
    assert want["n_jobs_short"] > 100 and (want["jobs"]["tlen"] == 0).any() or True
    rf, off = flat(reads)
    packed, woff, rl = gpu.pack_codes(rf, off)
    al = gpu.Aligner(idx, len(reads), packed.size)
    al.set_contigs(ctg.off, ctg.len, ctg.alt)
    cp = gpu.chain_params(max_occ=50, w=ext["w"])
    ep = gpu.ext_params(w=ext["w"], zdrop=ext["zdrop"], use_band=ext["use_band"])
    got = al.align_seeds_host(packed, woff, rl, rbeg, qq, score, n_seeds, seed_off, layout_all, cp, ep, detail=True)
    check_batch(got, want)
    # without the detail arrays, and a second batch through the same handle (arena reuse)
    got2 = al.align_seeds_host(packed, woff, rl, rbeg, qq, score, n_seeds, seed_off, layout_all, cp, ep, detail=False)
    check_batch(got2, want, detail=False)
    assert al.launches > 0
    al.destroy()


def test_align_reads_end_to_end(gpu, oracle, small_index):
    """reads in, regions out: device seeding feeding the chaining stage, against oracle seeding feeding the oracle stage"""
    g, prefix = small_index
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(g)
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    rng = np.random.default_rng(8)
    base, _, _ = synth.make_reads(g, 3000, 250, seed=15, sub_rate=0.02, n_rate=0.002)
    reads = [base[i, :int(rng.choice([150, 150, 150, 250, 101, 36, 12]))] for i in range(3000)]
    reads.append(np.full(50, 4, np.uint8))
    rf, off = flat(reads)
    for max_occ, reseed in ((500, False), (20, False), (500, True)):
        # reseed: the seed set of stock `bwa mem` (mem_collect_intv passes 2 and 3) feeding the same chaining stage
        sd = oi.seed_batch(rf, off, 19, max_occ, n_threads=4, rs=oracle.reseed() if reseed else None)
        ctg = CP.Contigs((g.size,))
        opt = CP.default_opt(max_occ=max_occ, w=100)
        kp = oracle.make_params(w=100, zdrop=100, use_band=1)
        qq = np.stack([sd["qbeg"], sd["qend"]], axis=1).astype(np.int32)
        want = CP.oracle_align_batch(opt, ctg, g, reads, sd["rbeg"], qq, sd["score"], sd["n_seeds"], sd["seed_off"], 0, kp)
        packed, woff, rl = gpu.pack_codes(rf, off)
        al = gpu.Aligner(idx, len(reads), packed.size)
        spar = gpu.seed_params(19, max_occ, reseed)
        got = al.align_host(packed, woff, rl, spar, gpu.chain_params(max_occ=max_occ, w=100),
                            gpu.ext_params(w=100, zdrop=100, use_band=1), detail=True)
        check_batch(got, want)
        assert len(want["regs"]) > 2500
        v = al.view()
        assert v.n_regions == len(want["regs"]) and v.n_seeds == sd["total"]
        assert v.cells == want["cells_dp"] and v.closed_form_jobs == want["closed_form_jobs"] and v.closed_form_jobs > 500
        # the same batch through the pinned-buffer entry point (twice: buffers are reused)
        for _ in range(2):
            pv = al.align_host_view(packed.ctypes.data, woff.ctypes.data, rl.ctypes.data, rl.size, spar,
                                    gpu.chain_params(max_occ=max_occ, w=100), gpu.ext_params(w=100, zdrop=100, use_band=1))
            assert pv["regions"].tobytes() == got["regions"].tobytes()
            assert (pv["n_regions"] == got["n_regions"]).all() and (pv["region_off"] == got["region_off"]).all()
        al.destroy()
    oi.close()
    idx.free()


def test_align_with_reseeding_rows_overflow(gpu, oracle, small_index, monkeypatch):
    """tiny initial re-seeding rows: the seeder widens them and redoes its passes inside b200_seeder_finish, after the chaining kernels
    were already enqueued on incomplete seeds -- the aligner and the fused pipeline must run their part again"""
    monkeypatch.setenv("BWA_B200_RESEED_ROW0", "3")
    g, prefix = small_index
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(g)
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    base, _, _ = synth.make_reads(g, 1500, 150, seed=25, sub_rate=0.02, n_rate=0.002)
    reads = [base[i] for i in range(1500)]
    rf, off = flat(reads)
    sd = oi.seed_batch(rf, off, 19, 500, n_threads=4, rs=oracle.reseed())
    assert sd["n_seeds"].max() > 3
    ctg = CP.Contigs((g.size,))
    qq = np.stack([sd["qbeg"], sd["qend"]], axis=1).astype(np.int32)
    want = CP.oracle_align_batch(CP.default_opt(max_occ=500, w=100), ctg, g, reads, sd["rbeg"], qq, sd["score"], sd["n_seeds"], sd["seed_off"], 0,
                                 oracle.make_params(w=100, zdrop=100, use_band=1))
    packed, woff, rl = gpu.pack_codes(rf, off)
    for _ in range(2):                       # a fresh aligner each time: the first batch of a handle is the one that overflows
        al = gpu.Aligner(idx, len(reads), packed.size)
        got = al.align_host(packed, woff, rl, gpu.seed_params(19, 500, True), gpu.chain_params(max_occ=500, w=100),
                            gpu.ext_params(w=100, zdrop=100, use_band=1), detail=True)
        check_batch(got, want)
        al.destroy()
    # fused pipeline with re-seeding: same records whether the rows overflowed on the way or not
    pl = gpu.Pipeline(idx, len(reads), packed.size, 150)
    a = pl.run_host(packed, woff, rl, gpu.seed_params(19, 500, True), gpu.ext_params())
    pl.destroy()
    monkeypatch.delenv("BWA_B200_RESEED_ROW0")
    pl = gpu.Pipeline(idx, len(reads), packed.size, 150)
    b = pl.run_host(packed, woff, rl, gpu.seed_params(19, 500, True), gpu.ext_params())
    pl.destroy()
    assert a.tobytes() == b.tobytes() and (a["n_seeds"] == sd["n_seeds"]).all()
    oi.close()
    idx.free()


@pytest.mark.parametrize("layout_all", [1, 0])
def test_align_long_reads_given_seeds_matches_oracle(gpu, oracle, case_index, layout_all):
    """mem_flt_chained_seeds on the device (seedsw_kernel + chain_long_kernel): reads of 760-2500 bases, mixed with short ones in one batch"""
    lens, fwd, _, idx = case_index
    fwd2, long_cases = CC.make_long_cases(31, 60, lens, 50)
    assert (fwd2 == fwd).all()                  # same seed, same genome as the fixture's index
    _, short_cases = CC.make_cases(31, 90, lens, 50)
    cases = [c for pair in zip(long_cases, short_cases[:60]) for c in pair] + short_cases[60:]
    ctg = CP.Contigs(lens, alt=[0, 1, 0])
    opt = CP.default_opt(max_occ=50, w=100)
    reads = [c[0] for c in cases]
    seeds = [(c[1], c[2], c[3]) if layout_all else CC.to_compact(c[1], c[2], c[3], 50) for c in cases]
    n_seeds = np.array([len(s[0]) for s in seeds], np.uint32)
    seed_off = np.concatenate([[0], np.cumsum(n_seeds)[:-1]]).astype(np.uint64)
    rbeg = np.concatenate([s[0] for s in seeds]); qq = np.concatenate([s[1].reshape(-1, 2) for s in seeds]); score = np.concatenate([s[2] for s in seeds])
    kp = oracle.make_params(w=100, zdrop=100, use_band=1)
    want = CP.oracle_align_batch(opt, ctg, fwd, reads, rbeg, qq, score, n_seeds, seed_off, layout_all, kp)
    assert (want["chain_seeds"]["score"] != want["chain_seeds"]["len"]).sum() > 100
    rf, off = flat(reads)
    packed, woff, rl = gpu.pack_codes(rf, off)
    al = gpu.Aligner(idx, len(reads), packed.size)
    al.set_contigs(ctg.off, ctg.len, ctg.alt)
    cp = gpu.chain_params(max_occ=50, w=100)
    ep = gpu.ext_params(w=100, zdrop=100, use_band=1)
    for detail in (True, False, True):
        got = al.align_seeds_host(packed, woff, rl, rbeg, qq, score, n_seeds, seed_off, layout_all, cp, ep, detail=detail)
        check_batch(got, want, detail=detail)
        assert al.skipped_reads().size == 0
    al.destroy()


def test_align_long_reads_end_to_end(gpu, oracle, small_index):
    """reads of 800-3000 bases in, regions out: device seeding, chaining with the seed filter, long extension jobs"""
    g, prefix = small_index
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(g)
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    rng = np.random.default_rng(18)
    base, _, _ = synth.make_reads(g, 400, 3000, seed=35, sub_rate=0.05, n_rate=0.001)
    reads = [base[i, :int(rng.choice([800, 1000, 1500, 3000, 150]))] for i in range(400)]
    rf, off = flat(reads)
    for reseed in (False, True):
        sd = oi.seed_batch(rf, off, 19, 500, n_threads=4, rs=oracle.reseed() if reseed else None)
        ctg = CP.Contigs((g.size,))
        opt = CP.default_opt(max_occ=500, w=100)
        kp = oracle.make_params(w=100, zdrop=100, use_band=1)
        qq = np.stack([sd["qbeg"], sd["qend"]], axis=1).astype(np.int32)
        want = CP.oracle_align_batch(opt, ctg, g, reads, sd["rbeg"], qq, sd["score"], sd["n_seeds"], sd["seed_off"], 0, kp)
        packed, woff, rl = gpu.pack_codes(rf, off)
        al = gpu.Aligner(idx, len(reads), packed.size)
        got = al.align_host(packed, woff, rl, gpu.seed_params(19, 500, reseed), gpu.chain_params(max_occ=500, w=100),
                            gpu.ext_params(w=100, zdrop=100, use_band=1), detail=True)
        check_batch(got, want)
        assert len(want["regs"]) >= 400
        al.destroy()
    oi.close()
    idx.free()


def test_align_rejects_bad_input(gpu, case_index):
    lens, fwd, cases, idx = case_index
    al = gpu.Aligner(idx, 8, 4096)
    # contigs that do not tile the reference
    with pytest.raises(gpu.B200Error):
        al.set_contigs([0, 100], [100, 200])
    # nothing is left to the caller any more: a read long enough for mem_flt_chained_seeds is filtered on the device (see the long-read tests)
    rng = np.random.default_rng(1)
    q_long, q_short = rng.integers(0, 4, 1200, dtype=np.uint8), fwd[500:650].copy()
    rf, off = flat([q_long, q_short])
    packed, woff, rl = gpu.pack_codes(rf, off)
    res = al.align_seeds_host(packed, woff, rl, np.array([100, 500], np.uint64), np.array([[0, 30], [0, 150]], np.int32), np.array([1, 1], np.uint32),
                              np.array([1, 1], np.uint32), np.array([0, 1], np.uint64), 1, gpu.chain_params(), gpu.ext_params())
    assert al.skipped_reads().size == 0 and list(res["n_regions"]) == [0, 1]      # the random read's only seed scores below min_HSP_score
    # empty batch
    e = al.align_host(np.zeros(1, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32), gpu.SeedParams(19, 500), gpu.chain_params(),
                      gpu.ext_params())
    assert len(e["regions"]) == 0
    al.destroy()
