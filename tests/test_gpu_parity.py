"""GPU parity tests: the CUDA hot paths, called through the C ABI, against the CPU oracle and the
golden vectors made from the reference's own CPU functions.  Bit-exact: every comparison is ==."""
import os

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def flat(reads_list):
    lens = np.array([len(r) for r in reads_list], np.uint64)
    off = np.zeros(len(reads_list) + 1, np.uint64)
    off[1:] = np.cumsum(lens)
    return (np.concatenate(reads_list).astype(np.uint8) if reads_list else np.zeros(0, np.uint8)), off


@pytest.fixture(scope="module")
def gpu(pkg):
    assert pkg.lib().bwa_b200_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    return pkg


@pytest.fixture(scope="module")
def dev_index(gpu, small_index, oracle):
    g, prefix = small_index
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    oi = oracle.OracleIndex(prefix + ".bwt", prefix + ".sa")
    yield g, idx, oi
    idx.free()
    oi.close()


def seed_compare(gpu, idx, oi, reads_flat, off, min_seed_len, max_occ):
    packed, woff, rl = gpu.pack_codes(reads_flat, off)
    n = rl.size
    sd = gpu.Seeder(idx, max(n, 1), max(packed.size, 1))
    got = sd.seed_host(packed, woff, rl, min_seed_len, max_occ)
    want = oi.seed_batch(reads_flat, off, min_seed_len, max_occ if max_occ > 0 else 0, n_threads=4)
    sm = sd.smems(n, max(1024, int(reads_flat.size)))
    wsm = oi.smem_batch(reads_flat, off, min_seed_len)
    sd.destroy()
    assert (sm["n_smems"] == wsm["n_smems"]).all()
    for key in ("qbeg", "qend", "k", "s"):
        assert (sm[key] == wsm[key]).all(), key
    assert got["total"] == want["total"]
    assert (got["n_seeds"] == want["n_seeds"]).all()
    assert (got["seed_off"] == want["seed_off"]).all()
    assert (got["qq"][:, 0] == want["qbeg"]).all() and (got["qq"][:, 1] == want["qend"]).all()
    assert (got["score"] == want["score"]).all()
    assert (got["rbeg"] == want["rbeg"]).all()
    return got, want


def test_index_info(dev_index):
    g, idx, oi = dev_index
    info = idx.info()
    assert info.seq_len == 2 * g.size == oi.seq_len
    assert info.sa_intv == 16 and info.n_buckets == (info.seq_len + 63) // 64


@pytest.mark.parametrize("max_occ", [500, 7, 0])
def test_seeding_matches_oracle(gpu, dev_index, max_occ):
    g, idx, oi = dev_index
    reads, _, _ = synth.make_reads(g, 3000, 150, seed=5, n_rate=0.002)
    got, want = seed_compare(gpu, idx, oi, reads.reshape(-1).copy(), (np.arange(3001) * 150).astype(np.uint64), 19, max_occ)
    assert got["total"] > 3000


def test_seeding_ragged_and_edge_reads(gpu, dev_index):
    g, idx, oi = dev_index
    rng = np.random.default_rng(3)
    rl = []
    base, _, _ = synth.make_reads(g, 400, 250, seed=9, sub_rate=0.02, n_rate=0.003)
    for i in range(400):
        rl.append(base[i, :int(rng.integers(1, 251))])
    rl.append(np.full(40, 4, np.uint8))                       # all N
    rl.append(np.zeros(5, np.uint8))                          # shorter than a seed
    rl.append(np.zeros(300, np.uint8))                        # poly-A, not in the genome as a whole
    rl.append(np.array([1], np.uint8))                        # single base
    rl.append(g[1000:1019].copy())                            # exactly min_seed_len
    rl.append(g[5000:5600].copy())                            # long exact read
    rl.append(synth.revcomp(g[7000:7400].copy()))
    f, off = flat(rl)
    seed_compare(gpu, idx, oi, f, off, 19, 500)
    seed_compare(gpu, idx, oi, f, off, 10, 3)
    seed_compare(gpu, idx, oi, f, off, 30, 500)


def test_seeding_wide_rows_path(gpu, dev_index, monkeypatch):
    """the 64-bit row kernels (selected for indexes of 2^32 rows and more, i.e. human-sized) forced on a small index"""
    g, idx, oi = dev_index
    monkeypatch.setenv("BWA_B200_WIDE_ROWS", "1")
    reads, _, _ = synth.make_reads(g, 3000, 150, seed=15, n_rate=0.002)
    seed_compare(gpu, idx, oi, reads.reshape(-1).copy(), (np.arange(3001) * 150).astype(np.uint64), 19, 500)
    long_reads, _, _ = synth.make_reads(g, 500, 250, seed=16, sub_rate=0.02)
    seed_compare(gpu, idx, oi, long_reads.reshape(-1).copy(), (np.arange(501) * 250).astype(np.uint64), 19, 20)


def test_seeding_empty_batch(gpu, dev_index):
    _, idx, _ = dev_index
    sd = gpu.Seeder(idx, 16, 16)
    got = sd.seed_host(np.zeros(1, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32))
    assert got["total"] == 0 and got["n_seeds"].size == 0
    sd.destroy()


def test_seeding_golden_from_reference(gpu, tmp_path):
    """CUDA path against SMEMs/seeds produced by the reference's bwt_smem1 / bwt_sa (make_golden.py)"""
    gold = np.load(os.path.join(GOLD, "seed_golden.npz"))
    g = synth.make_repeat_genome(int(gold["genome_len"]), seed=int(gold["genome_seed"]))
    prefix = str(tmp_path / "g")
    gpu.build_index(g, prefix, sa_intv=int(gold["sa_intv"]), n_threads=4)
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    reads = gold["reads"]
    n, L = reads.shape
    off = (np.arange(n + 1) * L).astype(np.uint64)
    packed, woff, rl = gpu.pack_codes(reads.reshape(-1).copy(), off)
    sd = gpu.Seeder(idx, n, packed.size)
    got = sd.seed_host(packed, woff, rl, 19, int(gold["max_occ"]))
    sm = sd.smems(n, 1 << 16)
    assert (sm["n_smems"] == gold["n_smems"]).all()
    for key in ("qbeg", "qend", "k", "s"):
        assert (sm[key] == gold[key]).all(), key
    assert (got["n_seeds"] == gold["n_seeds"]).all()
    assert (got["rbeg"] == gold["rbeg"]).all()
    assert (got["score"] == gold["score"]).all()
    sd.destroy()
    idx.free()


def test_seeding_device_api_and_seed_text_property(gpu, dev_index):
    """device-resident entry point; every located seed must spell the read substring on fwd+revcomp"""
    import torch
    g, idx, oi = dev_index
    reads, _, _ = synth.make_reads(g, 20000, 150, seed=77)
    f = reads.reshape(-1).copy()
    off = (np.arange(20001) * 150).astype(np.uint64)
    packed, woff, rl = gpu.pack_codes(f, off)
    d_packed = torch.from_numpy(packed.view(np.int32)).cuda()
    d_woff = torch.from_numpy(woff.view(np.int64)).cuda()
    d_rl = torch.from_numpy(rl.view(np.int32)).cuda()
    sd = gpu.Seeder(idx, 20000, packed.size)
    sd.seed_device(d_packed.data_ptr(), d_woff.data_ptr(), d_rl.data_ptr(), 20000, 19, 500)
    v = sd.device_result()
    tot = int(v.n_seeds)
    assert tot > 20000
    import ctypes as C

    class DevArr:      # wrap a raw device pointer for torch through the CUDA array interface
        def __init__(self, ptr, n, typestr):
            self.__cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(C.cast(ptr, C.c_void_p).value, False), version=2)

    rbeg = torch.as_tensor(DevArr(v.rbeg, tot, "<i8"), device="cuda")
    qq = torch.as_tensor(DevArr(v.qbeg_qend, tot * 2, "<i4"), device="cuda")
    nper = torch.as_tensor(DevArr(v.n_seeds_per_read, 20000, "<i4"), device="cuda")
    rbeg = rbeg.cpu().numpy().astype(np.uint64)
    qq = qq.cpu().numpy().reshape(-1, 2)
    nper = nper.cpu().numpy().astype(np.uint32)
    want = oi.seed_batch(f, off, 19, 500, n_threads=4)
    assert (nper == want["n_seeds"]).all() and (rbeg == want["rbeg"]).all()
    T = np.concatenate([g, synth.revcomp(g)])
    rid = np.repeat(np.arange(20000), nper)
    for j in np.random.default_rng(0).integers(0, tot, 2000):
        b, e = qq[j]
        assert (T[int(rbeg[j]):int(rbeg[j]) + (e - b)] == reads[rid[j], b:e]).all()
    sd.destroy()


# ------------------------------------------------------------------------------- extension

@pytest.mark.parametrize("name", ["ksw_band", "ksw_noband", "ksw_narrow", "ksw_asym"])
def test_extension_golden_from_reference(gpu, oracle, name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    jobs = {k: np.ascontiguousarray(gold[k]) for k in ("qseq", "tseq", "qoff", "toff", "qlen", "tlen", "h0")}
    kw = {k: int(v) for k, v in zip(gold["param_names"], gold["param_values"])}
    ex = gpu.Extender(0)
    res, tri = ex.extend_host(jobs, gpu.ext_params(**kw))
    assert (res == gold["res6"]).all()
    sc, qe, te = oracle.gasal_triple(gold["res6"], jobs["qlen"], kw.get("pen_clip", 5))
    assert (tri[0] == sc).all() and (tri[1] == qe).all() and (tri[2] == te).all()
    ex.destroy()


@pytest.mark.parametrize("kw", [dict(w=100, zdrop=100), dict(w=16, zdrop=100), dict(w=300, zdrop=0, use_band=0),
                                dict(w=5, zdrop=20), dict(w=50, zdrop=100, o_del=4, e_del=2, o_ins=7, e_ins=1, a=2, b=3, end_bonus=0)])
def test_extension_matches_oracle(gpu, oracle, kw):
    ex = gpu.Extender(0)
    for seed, extra in ((31, dict(qlen_range=(1, 300), h0_range=(1, 200))),
                        (32, dict(qlen_range=(1, 90), sub_rate=0.3, indel_rate=0.1, n_job_frac=0.2, h0_range=(1, 30))),
                        (33, dict(qlen_range=(200, 700), h0_range=(19, 150), pad8=False))):
        jobs = synth.make_ext_jobs(1500 if seed != 33 else 200, w=kw["w"], seed=seed, **extra)
        want, cnt = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        res, _ = ex.extend_host(jobs, gpu.ext_params(**kw))
        bad = np.nonzero((res != want).any(axis=1))[0]
        assert bad.size == 0, (bad[:5], res[bad[:5]], want[bad[:5]])
        assert (ex.last_cells(), ex.last_closed_form()) == synth.dp_cells(oracle, jobs, kw, cnt)
    ex.destroy()


@pytest.mark.parametrize("kw", [dict(w=100, zdrop=100), dict(w=16, zdrop=100), dict(w=8, zdrop=0), dict(w=50, zdrop=30, end_bonus=5),
                                dict(w=300, zdrop=0, use_band=0), dict(w=33, zdrop=100, o_del=3, e_del=1, o_ins=5, e_ins=2, a=3, b=2)])
def test_extension_pair_kernel_stress(gpu, oracle, kw):
    """the s16x2 column-pair kernel (two query columns per register, keyed and predicate row maxima) on many
    shapes: long seeds against short bands (right band clamp binds), noisy jobs (windows shrink from
    both sides, odd/even window edges), N bases; and the same jobs through the 32-bit kernel"""
    ex = gpu.Extender(0)
    for seed, extra in ((61, dict(qlen_range=(1, 260), h0_range=(1, 250))),
                        (62, dict(qlen_range=(1, 120), sub_rate=0.25, indel_rate=0.08, n_job_frac=0.3, h0_range=(1, 40))),
                        (63, dict(qlen_range=(100, 500), h0_range=(100, 400), sub_rate=0.02, indel_rate=0.02)),
                        (64, dict(qlen_range=(1, 40), h0_range=(1, 300), sub_rate=0.5, indel_rate=0.2))):
        jobs = synth.make_ext_jobs(20000, w=kw["w"], seed=seed, **extra)
        want, cnt = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        res, _ = ex.extend_host(jobs, gpu.ext_params(**kw))
        bad = np.nonzero((res != want).any(axis=1))[0]
        assert bad.size == 0, (seed, bad[:5], res[bad[:5]], want[bad[:5]], jobs["qlen"][bad[:5]], jobs["tlen"][bad[:5]], jobs["h0"][bad[:5]])
        assert (ex.last_cells(), ex.last_closed_form()) == synth.dp_cells(oracle, jobs, kw, cnt)
        os.environ["BWA_B200_EXT_NO_SIMD"] = "1"
        try:
            res32, _ = ex.extend_host(jobs, gpu.ext_params(**kw))
        finally:
            del os.environ["BWA_B200_EXT_NO_SIMD"]
        assert (res32 == want).all()
    ex.destroy()


@pytest.mark.parametrize("kw", [dict(w=100, zdrop=100), dict(w=21, zdrop=0), dict(w=40, zdrop=60, a=2, b=3)])
def test_extension_closed_form_flanks_and_repeats(gpu, oracle, kw):
    """the jobs key_kernel answers without a matrix (closed_form_job): flanks of exact matches with up to six substitutions, over random
    sequence and over tandem repeats whose shifted diagonals are clean but for a break or two -- results and the count of jobs taken"""
    ex = gpu.Extender(0)
    taken = 0
    for jobs in (synth.make_flank_jobs(4000, seed=81, w=kw["w"]), synth.make_flank_jobs(3000, seed=82, w=kw["w"], qlen_range=(100, 600), h0_range=(19, 250)),
                 synth.make_repeat_flank_jobs(6000, 83, kw)):
        want, cnt = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
        res, _ = ex.extend_host(jobs, gpu.ext_params(**kw))
        bad = np.nonzero((res != want).any(axis=1))[0]
        assert bad.size == 0, (bad[:5], res[bad[:5]], want[bad[:5]], jobs["qlen"][bad[:5]], jobs["h0"][bad[:5]])
        assert (ex.last_cells(), ex.last_closed_form()) == synth.dp_cells(oracle, jobs, kw, cnt)
        taken += ex.last_closed_form()
    assert taken > 3000
    ex.destroy()


def test_extension_general_matrix(gpu, oracle):
    """a general substitution matrix (transitions cheaper than transversions): PRMT score rows in the pair kernel"""
    jobs = synth.make_ext_jobs(3000, w=100, seed=71, qlen_range=(1, 200), h0_range=(1, 150))
    P = oracle.make_params(w=100, zdrop=100)
    mat = np.array([[2, -3, -1, -3, -1], [-3, 2, -3, -1, -1], [-1, -3, 2, -3, -1], [-3, -1, -3, 2, -1], [-1, -1, -1, -1, -1]], np.int8)
    for i in range(25):
        P.mat[i] = int(mat.reshape(-1)[i])
    want, _ = oracle.ksw_batch(jobs, P, n_threads=4)
    ep = gpu.ext_params(w=100, zdrop=100)
    for i in range(25):
        ep.mat[i] = int(mat.reshape(-1)[i])
    ex = gpu.Extender(0)
    res, _ = ex.extend_host(jobs, ep)
    assert (res == want).all()
    # a matrix whose "query is N" column is not uniform is not eligible for the pair kernel: 32-bit kernel
    P.mat[4] = -2
    ep.mat[4] = -2
    want, _ = oracle.ksw_batch(jobs, P, n_threads=4)
    res, _ = ex.extend_host(jobs, ep)
    assert (res == want).all()
    ex.destroy()


def test_extension_device_packed_path(gpu, oracle):
    import torch
    kw = dict(w=100, zdrop=100)
    jobs = synth.make_ext_jobs(3000, w=100, seed=41, qlen_range=(1, 200), h0_range=(1, 150))
    longj = synth.make_ext_jobs(60, w=100, seed=42, qlen_range=(1030, 1600), h0_range=(1, 150))      # one-warp-per-job kernel, packed input
    jobs = {k: np.concatenate([jobs[k], longj[k]]) for k in ("qseq", "tseq", "qlen", "tlen", "h0")} | \
           {"qoff": np.concatenate([jobs["qoff"], longj["qoff"] + np.uint32(jobs["qseq"].size)]).astype(np.uint32),
            "toff": np.concatenate([jobs["toff"], longj["toff"] + np.uint32(jobs["tseq"].size)]).astype(np.uint32)}
    want, _ = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
    ex = gpu.Extender(0)
    dq = torch.from_numpy(jobs["qseq"]).cuda()
    dt = torch.from_numpy(jobs["tseq"]).cuda()
    qp = torch.empty((dq.numel() + 7) // 8, dtype=torch.int32, device="cuda")
    tp = torch.empty((dt.numel() + 7) // 8, dtype=torch.int32, device="cuda")
    ex.pack_device(dq.data_ptr(), dq.numel(), qp.data_ptr())
    ex.pack_device(dt.data_ptr(), dt.numel(), tp.data_ptr())
    dev = {k: torch.from_numpy(jobs[k].view(np.int32)).cuda() for k in ("qoff", "toff", "qlen", "tlen", "h0")}
    out = torch.zeros(3060 * 6, dtype=torch.int32, device="cuda")
    ex.extend_device(gpu.ext_params(**kw), 3060, qp.data_ptr(), dev["qoff"].data_ptr(), dev["qlen"].data_ptr(),
                     tp.data_ptr(), dev["toff"].data_ptr(), dev["tlen"].data_ptr(), dev["h0"].data_ptr(), out.data_ptr())
    ex.wait()
    assert (out.cpu().numpy().reshape(-1, 6) == want).all()
    ex.destroy()


def test_extension_long_queries_and_wide_scores_intra_kernel(gpu, oracle):
    """jobs the per-lane kernels do not take -- queries beyond 1024 bases, scores beyond 16 bits -- run one per warp in
    ext_intra_kernel (32 columns of a row at once, F by a warp max-plus scan); mixed with ordinary jobs in one batch"""
    ex = gpu.Extender(0)
    for kw in (dict(w=100, zdrop=100), dict(w=300, zdrop=0, use_band=0), dict(w=20, zdrop=50)):
        longj = synth.make_ext_jobs(300, w=kw["w"], seed=91, qlen_range=(1000, 3000), h0_range=(1, 200), sub_rate=0.08, indel_rate=0.02)
        want, cnt = oracle.ksw_batch(longj, oracle.make_params(**kw), n_threads=4)
        res, _ = ex.extend_host(longj, gpu.ext_params(**kw))
        assert (res == want).all()
        assert int(gpu.lib().bwa_b200_extender_last_cells(ex.h)) == synth.dp_cells(oracle, longj, kw, cnt)[0]
    # one very long query (and a target that diverges half way: z-drop / window shrink on a long row)
    rng = np.random.default_rng(5)
    q = rng.integers(0, 4, 15000, dtype=np.uint8)
    t = q.copy(); t[12000:] = rng.integers(0, 4, 3000, dtype=np.uint8)
    pad = lambda x: np.concatenate([x, np.full(-x.size % 8, 4, np.uint8)])
    one = dict(qseq=pad(q), tseq=pad(t), qoff=np.zeros(1, np.uint32), toff=np.zeros(1, np.uint32), qlen=np.array([q.size], np.uint32),
               tlen=np.array([t.size], np.uint32), h0=np.array([30], np.uint32))
    for kw in (dict(w=100, zdrop=100), dict(w=100, zdrop=0)):
        want, _ = oracle.ksw_batch(one, oracle.make_params(**kw), n_threads=1)
        res, _ = ex.extend_host(one, gpu.ext_params(**kw))
        assert (res == want).all() and res[0, 0] > 10000
    # scores beyond 16 bits with ordinary lengths: a = 60, h0 large; mixed with short jobs of the per-lane kernels
    mixed = synth.make_ext_jobs(2000, w=100, seed=92, qlen_range=(1, 1100), h0_range=(1, 150))
    kw = dict(a=60, b=90, o_del=100, e_del=20, o_ins=120, e_ins=30, w=100, zdrop=2000)
    want, _ = oracle.ksw_batch(mixed, oracle.make_params(**kw), n_threads=4)
    res, _ = ex.extend_host(mixed, gpu.ext_params(**kw))
    assert (res == want).all() and res[:, 0].max() > 40000
    ex.destroy()


def test_extension_rejects_bad_jobs(gpu):
    jobs = synth.make_ext_jobs(8, w=100, seed=1, qlen_range=(10, 20))
    jobs["h0"][3] = 0                                  # ksw_extend2 asserts h0 > 0
    ex = gpu.Extender(0)
    with pytest.raises(gpu.B200Error):
        ex.extend_host(jobs, gpu.ext_params())
    jobs = synth.make_ext_jobs(4, w=100, seed=1, qlen_range=(10, 20))
    res, _ = ex.extend_host(jobs, gpu.ext_params())   # the extender recovers after an error
    assert (res[:, 0] >= jobs["h0"].astype(np.int32)).all()
    with pytest.raises(gpu.B200Error):                 # n_jobs == 0 (gasal_aln_async exits, gasal_align.cu:32)
        gpu.check(gpu.lib().bwa_b200_extend_async(ex.h, gpu.ext_params(), 0, jobs["qseq"].ctypes.data, 8, jobs["qoff"].ctypes.data,
                                                  jobs["qlen"].ctypes.data, jobs["tseq"].ctypes.data, 8, jobs["toff"].ctypes.data,
                                                  jobs["tlen"].ctypes.data, jobs["h0"].ctypes.data, None, None, None, None))
    ex.destroy()


# -------------------------------------------------------------------------- fused pipeline

@pytest.mark.parametrize("kw", [dict(w=100, zdrop=100), dict(w=300, zdrop=0, use_band=0), dict(w=10, zdrop=50)])
def test_pipeline_matches_oracle(gpu, oracle, dev_index, kw):
    g, idx, oi = dev_index
    idx.attach_ref(g)
    reads, _, _ = synth.make_reads(g, 6000, 150, seed=55, sub_rate=0.02, n_rate=0.001)
    edge = np.stack([g[:150], synth.revcomp(g[:150]), g[-150:], synth.revcomp(g[-150:])])   # windows clipped at the ends
    reads = np.concatenate([reads, edge, np.full((2, 150), 4, np.uint8)])
    n = reads.shape[0]
    f = reads.reshape(-1).copy()
    off = (np.arange(n + 1) * 150).astype(np.uint64)
    want, fc, kc = oracle.pipeline(oi, g, f, off, oracle.make_params(**kw), 19, 500, n_threads=4)
    packed, woff, rl = gpu.pack_codes(f, off)
    pl = gpu.Pipeline(idx, n, packed.size, 150)
    got = pl.run_host(packed, woff, rl, gpu.SeedParams(19, 500), gpu.ext_params(**kw))
    for key in ("seed_rbeg", "seed_qbeg", "seed_qend", "n_seeds", "h0", "left", "right"):
        bad = np.nonzero((got[key] != want[key]).reshape(n, -1).any(axis=1))[0]
        assert bad.size == 0, (key, bad[:5], got[key][bad[:5]], want[key][bad[:5]])
    tot = pl.totals()
    assert tot["cells"] <= kc["cells"] and tot["seeds"] == fc["n_located"]       # jobs answered in closed form are not in the cell count
    # a second batch through the same pipeline (no reallocation, no stale state)
    got2 = pl.run_host(packed, woff, rl, gpu.SeedParams(19, 500), gpu.ext_params(**kw))
    assert got2.tobytes() == got.tobytes()
    pl.destroy()
    # every job through the kernels: the same records, and the evaluated cells are the oracle's
    os.environ["BWA_B200_EXT_NO_CLOSED"] = "1"
    try:
        pl = gpu.Pipeline(idx, n, packed.size, 150)
    finally:
        del os.environ["BWA_B200_EXT_NO_CLOSED"]
    got3 = pl.run_host(packed, woff, rl, gpu.SeedParams(19, 500), gpu.ext_params(**kw))
    assert got3.tobytes() == got.tobytes() and pl.totals()["cells"] == kc["cells"]
    if synth.closed_form_eligible(**kw) is not None:
        assert tot["cells"] < kc["cells"]
    pl.destroy()


def test_pipeline_host_call_in_slices(gpu, oracle, dev_index, monkeypatch):
    """the host-buffer call cuts a large batch into slices whose copies overlap the kernels: same records as one pass, ragged
    read lengths so that slice boundaries fall at arbitrary word offsets"""
    g, idx, oi = dev_index
    idx.attach_ref(g)
    n = 3 * 32768 + 1234
    base, _, _ = synth.make_reads(g, n, 150, seed=77, sub_rate=0.02, n_rate=0.001)
    rng = np.random.default_rng(7)
    lens = rng.choice([150, 150, 149, 101, 75, 36], size=n)
    reads = [base[i, :lens[i]] for i in range(n)]
    f, off = flat(reads)
    want, _, _ = oracle.pipeline(oi, g, f, off, oracle.make_params(), 19, 500, n_threads=4)
    packed, woff, rl = gpu.pack_codes(f, off)
    pl = gpu.Pipeline(idx, n, packed.size, 150)
    monkeypatch.setenv("BWA_B200_HOST_SLICES", "3")
    got = pl.run_host(packed, woff, rl, gpu.SeedParams(19, 500), gpu.ext_params())
    assert got.tobytes() == want.tobytes()
    monkeypatch.setenv("BWA_B200_HOST_SLICES", "1")
    got1 = pl.run_host(packed, woff, rl, gpu.SeedParams(19, 500), gpu.ext_params())
    assert got1.tobytes() == want.tobytes()
    pl.destroy()


@pytest.mark.parametrize("K,sat", [(0, None), (4, None), (9, None), (9, 40), (7, 3000), (6, 2)])
def test_seeding_kmer_table_variants(gpu, dev_index, monkeypatch, K, sat):
    """the k-mer interval table (seed.cu): every table size, including none, and every size from which an entry defers to the
    occurrence buckets, gives the oracle's SMEMs and seeds -- a table step and a bucket step are interchangeable at any point"""
    g, idx, oi = dev_index
    if sat is not None:
        monkeypatch.setenv("BWA_B200_KMER_SAT", str(sat))
    idx.set_kmer_table(K)
    try:
        reads, _, _ = synth.make_reads(g, 2500, 150, seed=25 + K, sub_rate=0.02, n_rate=0.003)
        seed_compare(gpu, idx, oi, reads.reshape(-1).copy(), (np.arange(2501) * 150).astype(np.uint64), 19, 500)
        rng = np.random.default_rng(K)
        rl = [reads[i, :int(rng.integers(1, 151))] for i in range(300)]
        rl += [np.full(30, 4, np.uint8), np.zeros(200, np.uint8), g[100:130].copy(), synth.revcomp(g[9000:9100].copy()), g[-60:].copy(), synth.revcomp(g[:60].copy())]
        f, off = flat(rl)
        seed_compare(gpu, idx, oi, f, off, 12, 50)
        monkeypatch.setenv("BWA_B200_WIDE_ROWS", "1")
        seed_compare(gpu, idx, oi, f, off, 19, 500)
    finally:
        monkeypatch.delenv("BWA_B200_KMER_SAT", raising=False)
        monkeypatch.delenv("BWA_B200_WIDE_ROWS", raising=False)
        idx.set_kmer_table(9)


@pytest.mark.parametrize("kw", [dict(w=100, zdrop=100), dict(w=1, zdrop=0), dict(w=2, zdrop=20), dict(w=7, zdrop=100), dict(w=33, zdrop=100, o_del=3, e_del=1, o_ins=5, e_ins=2, a=2, b=3),
                                dict(w=64, zdrop=0, end_bonus=0), dict(w=500, zdrop=100), dict(w=2030, zdrop=100)],
                         ids=lambda k: f"w{k['w']}z{k['zdrop']}")
@pytest.mark.parametrize("wide", [False, True], ids=["wave", "wide"])
def test_extension_wave_kernel_banded_long_and_wide(gpu, oracle, kw, wide, monkeypatch):
    """the jobs of a banded batch outside the column-pair class -- queries of 257 .. 3000 bases, scores beyond 1023 -- through both kernels
    that take them: ext_wave_kernel (one job per warp, two columns per lane in s16x2, F by a warp max-plus scan, ring of column pairs in
    shared memory; small batches) and ext_pair_kernel<WIDE> (one job per lane, 32-bit row-maximum keys, ring state; batches that fill
    the machine -- forced here by BWA_B200_EXT_WIDE_MIN).  Mixed with jobs of the other kernels, through the byte-per-base host call and
    the packed device call; bit-exact against the oracle, evaluated cells included"""
    import torch
    if wide and kw["w"] > 600:
        pytest.skip("the ring of this band does not fit a lane's shared memory: ext_wave_kernel takes it whatever the count")
    monkeypatch.setenv("BWA_B200_EXT_WIDE_MIN", "1" if wide else "1000000000")
    ex = gpu.Extender(0)
    jobs = synth.make_ext_jobs(700, w=kw["w"], seed=97 + kw["w"], qlen_range=(200, 3000), h0_range=(1, 900), sub_rate=0.06, indel_rate=0.02, n_job_frac=0.1)
    short = synth.make_ext_jobs(500, w=kw["w"], seed=98, qlen_range=(1, 300), h0_range=(1, 150))
    want, cnt = oracle.ksw_batch(jobs, oracle.make_params(**kw), n_threads=4)
    l0 = ex.launches
    res, _ = ex.extend_host(jobs, gpu.ext_params(**kw))
    assert ex.launches > l0
    bad = np.nonzero((res != want).any(axis=1))[0]
    assert bad.size == 0, (bad[:5], jobs["qlen"][bad[:5]], jobs["tlen"][bad[:5]], jobs["h0"][bad[:5]], res[bad[:3]], want[bad[:3]])
    assert int(gpu.lib().bwa_b200_extender_last_cells(ex.h)) == synth.dp_cells(oracle, jobs, kw, cnt)[0]
    # packed device path, long and short jobs in one batch
    both = {k: np.concatenate([jobs[k], short[k]]) for k in ("qseq", "tseq", "qlen", "tlen", "h0")}
    both["qoff"] = np.concatenate([jobs["qoff"], short["qoff"] + np.uint32(jobs["qseq"].size)]).astype(np.uint32)
    both["toff"] = np.concatenate([jobs["toff"], short["toff"] + np.uint32(jobs["tseq"].size)]).astype(np.uint32)
    want2, _ = oracle.ksw_batch(both, oracle.make_params(**kw), n_threads=4)
    n = both["qlen"].size
    dq = torch.from_numpy(both["qseq"]).cuda(); dt = torch.from_numpy(both["tseq"]).cuda()
    qp = torch.empty((dq.numel() + 7) // 8, dtype=torch.int32, device="cuda"); tp = torch.empty((dt.numel() + 7) // 8, dtype=torch.int32, device="cuda")
    ex.pack_device(dq.data_ptr(), dq.numel(), qp.data_ptr()); ex.pack_device(dt.data_ptr(), dt.numel(), tp.data_ptr())
    dev = {k: torch.from_numpy(both[k].view(np.int32)).cuda() for k in ("qoff", "toff", "qlen", "tlen", "h0")}
    out = torch.zeros(n * 6, dtype=torch.int32, device="cuda")
    ex.extend_device(gpu.ext_params(**kw), n, qp.data_ptr(), dev["qoff"].data_ptr(), dev["qlen"].data_ptr(), tp.data_ptr(), dev["toff"].data_ptr(),
                     dev["tlen"].data_ptr(), dev["h0"].data_ptr(), out.data_ptr())
    ex.wait()
    assert (out.cpu().numpy().reshape(-1, 6) == want2).all()
    ex.destroy()
