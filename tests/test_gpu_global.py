"""GPU parity tests of the CIGAR path (ksw_global2 with backtrack + NM): the CUDA kernel through the C ABI
against the oracle and against golden vectors from the reference's own ksw_global2 / bwa_gen_cigar2.  Bit-exact."""
import os

import numpy as np
import pytest

from tools import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gpu(pkg):
    assert pkg.lib().bwa_b200_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    return pkg


def rows_of(res, stride):
    """flat CIGAR + offsets -> [n, stride] left-aligned rows (the oracle's layout)"""
    n = res["n_cigar"].size
    rows = np.zeros((n, stride), np.uint32)
    for a in range(n):
        m = int(res["n_cigar"][a]); o = int(res["cigar_off"][a])
        rows[a, :m] = res["cigar"][o:o + m]
    return rows


def compare(gpu, oracle, cg, jobs, pkw=None, stride=256):
    pkw = pkw or {}
    got = cg.global_host(jobs, gpu.ext_params(**pkw))
    want = oracle.global_batch(jobs, oracle.make_params(**pkw), cig_stride=stride, n_threads=4)
    assert (want["n_cigar"] <= stride).all()
    assert (got["score"] == want["score"]).all()
    assert (got["n_cigar"] == want["n_cigar"]).all()
    assert (got["nm"] == want["nm"]).all()
    assert (rows_of(got, stride) == want["cigar"]).all()
    assert cg.last_cells == want["cells"]
    off = np.zeros(got["n_cigar"].size, np.uint64)
    off[1:] = np.cumsum(got["n_cigar"].astype(np.uint64))[:-1]
    assert (got["cigar_off"] == off).all() and got["cigar"].size == int(got["n_cigar"].sum())
    return got


def test_global_matches_oracle(gpu, oracle):
    cg = gpu.Cigar(0)
    compare(gpu, oracle, cg, synth.make_global_jobs(20000, qlen_range=(1, 150), seed=501))
    compare(gpu, oracle, cg, synth.make_global_jobs(5000, qlen_range=(80, 300), seed=502, sub_rate=0.08, indel_rate=0.03, w_extra=(0, 40)))
    # every band class up to 127, other scoring
    compare(gpu, oracle, cg, synth.make_global_jobs(3000, qlen_range=(100, 400), seed=503, sub_rate=0.1, indel_rate=0.04, w_extra=(0, 120), w_cap=127),
            dict(a=2, b=3, o_del=4, e_del=2, o_ins=5, e_ins=1))
    # tight bands, many operations per CIGAR (the widen-the-rows path: more than 16 operations)
    got = compare(gpu, oracle, cg, synth.make_global_jobs(3000, qlen_range=(20, 200), seed=504, sub_rate=0.2, indel_rate=0.08, w_extra=(0, 2)))
    assert got["n_cigar"].max() > 16
    assert cg.launches > 0
    cg.destroy()


def test_global_edge_cases(gpu, oracle):
    """one-base sequences, band narrower than the length difference (the last row misses column qlen: score MINUS_INF), all-N"""
    cg = gpu.Cigar(0)
    q = [np.array([2], np.uint8), np.array([0, 1, 2, 3, 0, 1, 2, 3, 0, 1, 2, 3], np.uint8), np.full(20, 4, np.uint8), np.array([1, 1, 1, 1, 1, 1, 1, 1, 1, 1], np.uint8),
         np.array([0, 1, 2, 3, 0, 1, 2, 3], np.uint8)]
    t = [np.array([2], np.uint8), np.array([0, 1, 2], np.uint8), np.full(18, 4, np.uint8), np.array([1], np.uint8), np.array([0, 1, 2, 3, 0, 1, 2, 3, 3, 3], np.uint8)]
    # (a target longer than qlen + w would make the reference's backtrack read cells it never wrote: not a defined case)
    w = np.array([3, 2, 5, 1, 2], np.uint32)
    qoff = np.zeros(len(q), np.uint32); toff = np.zeros(len(q), np.uint32)
    qs, ts = [], []
    for a in range(len(q)):
        qoff[a] = sum(x.size for x in qs); toff[a] = sum(x.size for x in ts)
        qs.append(np.concatenate([q[a], np.full(-q[a].size % 8, 4, np.uint8)])); ts.append(np.concatenate([t[a], np.full(-t[a].size % 8, 4, np.uint8)]))
    jobs = dict(qseq=np.concatenate(qs), tseq=np.concatenate(ts), qoff=qoff, toff=toff, qlen=np.array([x.size for x in q], np.uint32),
                tlen=np.array([x.size for x in t], np.uint32), w=w)
    got = compare(gpu, oracle, cg, jobs)
    assert got["score"][1] == -0x40000000 and got["score"][0] == 1
    # empty batch; band beyond the supported maximum is refused loudly
    e = cg.global_host(dict(qseq=np.zeros(0, np.uint8), tseq=np.zeros(0, np.uint8), qoff=np.zeros(0, np.uint32), toff=np.zeros(0, np.uint32),
                            qlen=np.zeros(0, np.uint32), tlen=np.zeros(0, np.uint32), w=np.zeros(0, np.uint32)), gpu.ext_params())
    assert e["score"].size == 0 and e["cigar"].size == 0
    jobs["w"] = np.array([3, 2, 5, 1, 200], np.uint32)
    with pytest.raises(RuntimeError):
        cg.global_host(jobs, gpu.ext_params())
    cg.destroy()


def test_global_golden_from_reference(gpu, oracle):
    gold = np.load(os.path.join(GOLD, "global_golden.npz"))
    cg = gpu.Cigar(0)
    p = gpu.ext_params()
    for si in range(3):
        jobs = {k: gold[f"s{si}_{k}"] for k in ("qseq", "tseq", "qoff", "toff", "qlen", "tlen", "w")}
        got = cg.global_host(jobs, p)
        stride = gold[f"s{si}_cigar"].shape[1]
        assert (got["score"] == gold[f"s{si}_score"]).all()
        assert (got["n_cigar"] == gold[f"s{si}_n_cigar"]).all()
        assert (rows_of(got, stride) == gold[f"s{si}_cigar"]).all()
    # bwa_gen_cigar2: the caller's part (window fetch, strand reversal, band rule) done here as INTEGRATION.md shows, the DP,
    # backtrack and NM on the device
    g = synth.make_genome(int(gold["gc_genome_len"]), seed=int(gold["gc_genome_seed"]))
    L = g.size
    qoff_g = gold["gc_qoff"]
    qs, ts, ws, keep = [], [], [], []
    for i in range(gold["gc_rb"].size):
        q = gold["gc_query"][qoff_g[i]:qoff_g[i + 1]].copy()
        rb, re, w_ = int(gold["gc_rb"][i]), int(gold["gc_re"][i]), int(gold["gc_w"][i])
        if rb >= L:
            r = (3 - g[2 * L - re:2 * L - rb])[::-1].copy()          # bns_get_seq on the reverse strand
            q, r = q[::-1].copy(), r[::-1].copy()                  # then both reversed (src/bwa.c:145-150)
        else:
            r = g[rb:re].copy()
        if q.size == r.size and w_ == 0:
            continue                                               # the ungapped shortcut needs no DP (src/bwa.c:151-160)
        keep.append(i); qs.append(q); ts.append(r); ws.append(gpu.Cigar.band(p, w_, q.size, r.size))
    assert len(keep) > 300
    qoff = np.zeros(len(qs), np.uint32); toff = np.zeros(len(qs), np.uint32)
    qoff[1:] = np.cumsum([x.size for x in qs])[:-1]; toff[1:] = np.cumsum([x.size for x in ts])[:-1]
    jobs = dict(qseq=np.concatenate(qs), tseq=np.concatenate(ts), qoff=qoff, toff=toff, qlen=np.array([x.size for x in qs], np.uint32),
                tlen=np.array([x.size for x in ts], np.uint32), w=np.array(ws, np.uint32))
    got = cg.global_host(jobs, p)
    keep = np.array(keep)
    assert (got["score"] == gold["gc_score"][keep]).all()
    assert (got["nm"] == gold["gc_nm"][keep]).all()
    assert (got["n_cigar"] == gold["gc_n_cigar"][keep]).all()
    assert (rows_of(got, gold["gc_cigar"].shape[1]) == gold["gc_cigar"][keep]).all()
    cg.destroy()


def test_reg2aln_matches_oracle_and_fork_golden(gpu, oracle, tmp_path):
    """mem_reg2aln over a batch: windows cut on the device, band-doubling retries as waves, squeeze / clips / position on the host --
    against the fork's own mem_reg2aln (golden) and the oracle (band and number of waves too)"""
    from oracle import chain_py as CP
    gold = np.load(os.path.join(GOLD, "reg2aln_golden.npz"))
    ctg = CP.Contigs(tuple(int(x) for x in gold["contigs"]))
    g = synth.make_genome(ctg.l_pac, seed=int(gold["genome_seed"]))
    prefix = str(tmp_path / "g")
    gpu.build_index(g, prefix, sa_intv=16, n_threads=4)
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(g)
    reads, regs = gold["reads"], gold["regs"]
    n, L = reads.shape
    packed, woff, rl = gpu.pack_codes(reads.reshape(-1).copy(), (np.arange(n + 1) * L).astype(np.uint64))
    alns = np.zeros(len(regs), gpu.ALN_IN_DTYPE)
    for k, (i, qb, qe, rb, re, truesc, w) in enumerate(regs):
        alns[k] = (i, qb, qe, rb, re, truesc, w)
    cg = gpu.Cigar(0)
    waves = 0
    for oi in range(2):
        v = gold[f"opt_{oi}"]
        opt = CP.default_opt(a=int(v[0]), b=int(v[1]), o_del=int(v[2]), e_del=int(v[3]), o_ins=int(v[4]), e_ins=int(v[5]), w=int(v[6]))
        kw = dict(a=opt.a, b=opt.b, o_del=opt.o_del, e_del=opt.e_del, o_ins=opt.o_ins, e_ins=opt.e_ins)
        got, flat = cg.reg2aln_host(idx, ctg.off, packed, woff, rl, alns, gpu.ext_params(w=opt.w, **kw), opt.a)
        rec, cig = gold[f"rec_{oi}"], gold[f"cigar_{oi}"]
        kp = oracle.make_params(**kw)
        for k, (i, qb, qe, rb, re, truesc, w) in enumerate(regs):
            o = got[k]
            assert (int(o["pos"]), int(o["rid"]), int(o["is_rev"])) == tuple(int(x) for x in rec[k, :3]), k
            if rec[k, 1] < 0:
                assert o["n_cigar"] == 0
                continue
            mine = flat[int(o["cigar_off"]):int(o["cigar_off"]) + int(o["n_cigar"])]
            assert int(o["nm"]) == rec[k, 3] and int(o["n_cigar"]) == rec[k, 4], k
            assert (mine == cig[k, :mine.size]).all(), k
            a, _ = CP.oracle_reg2aln(opt, kp, ctg, g, reads[i], qb, qe, rb, re, truesc, w)
            assert int(o["score"]) == int(a["score"]) and int(o["band"]) == int(a["band"]) and int(o["n_waves"]) == int(a["n_waves"]), k
            waves = max(waves, int(o["n_waves"]))
    assert waves >= 2                                        # the band-doubling retry did run on some region
    # regions outside the read / the reference are refused
    bad = alns[:1].copy(); bad["qe"] = L + 5
    with pytest.raises(RuntimeError):
        cg.reg2aln_host(idx, ctg.off, packed, woff, rl, bad, gpu.ext_params(), 1)
    cg.destroy()
    idx.free()


def test_global_ring_kernel_on_narrow_bands(gpu, oracle, monkeypatch):
    """bands up to 15 normally run in the register-resident kernel; the shared-memory ring kernel (the one wider bands use) must give
    the same answers on them"""
    monkeypatch.setenv("BWA_B200_GLOBAL_RING", "1")
    cg = gpu.Cigar(0)
    compare(gpu, oracle, cg, synth.make_global_jobs(8000, qlen_range=(1, 150), seed=601))
    compare(gpu, oracle, cg, synth.make_global_jobs(3000, qlen_range=(60, 250), seed=602, sub_rate=0.1, indel_rate=0.04, w_extra=(0, 12), w_cap=15),
            dict(a=2, b=3, o_del=4, e_del=2, o_ins=5, e_ins=1))
    cg.destroy()


def test_align_regions_through_reg2aln(gpu, oracle, small_index):
    """the two stages composed on real data: regions produced by bwa_b200_align_host (chaining + extension on the device) go through
    bwa_b200_reg2aln_host; every alignment equals the oracle's mem_reg2aln on the same region"""
    from oracle import chain_py as CP
    g, prefix = small_index
    idx = gpu.Index.load(prefix + ".bwt", prefix + ".sa", 0)
    idx.attach_ref(g)
    reads, _, _ = synth.make_reads(g, 2000, 150, seed=123, sub_rate=0.03, ins_rate=0.003, del_rate=0.003)
    n, L = reads.shape
    packed, woff, rl = gpu.pack_codes(reads.reshape(-1).copy(), (np.arange(n + 1) * L).astype(np.uint64))
    al = gpu.Aligner(idx, n, packed.size)
    res = al.align_host(packed, woff, rl, gpu.SeedParams(19, 500), gpu.chain_params(w=100), gpu.ext_params(w=100, zdrop=100, use_band=1))
    al.destroy()
    regs = res["regions"]
    read_of = np.repeat(np.arange(n), res["n_regions"])
    keep = (regs["qe"] > regs["qb"]) & (regs["re"] > regs["rb"])
    assert keep.sum() > 1500
    alns = np.zeros(int(keep.sum()), gpu.ALN_IN_DTYPE)
    alns["read"] = read_of[keep]; alns["qb"] = regs["qb"][keep]; alns["qe"] = regs["qe"][keep]; alns["rb"] = regs["rb"][keep]; alns["re"] = regs["re"][keep]
    alns["truesc"] = regs["truesc"][keep]; alns["w"] = regs["w"][keep]
    cg = gpu.Cigar(0)
    ctg = CP.Contigs((g.size,))
    got, flat = cg.reg2aln_host(idx, ctg.off, packed, woff, rl, alns, gpu.ext_params(w=100), 1)
    cg.destroy()
    idx.free()
    opt, kp = CP.default_opt(w=100), oracle.make_params()
    n_gapped = 0
    for k in range(alns.size):
        a = alns[k]
        want, wc = CP.oracle_reg2aln(opt, kp, ctg, g, reads[int(a["read"])], int(a["qb"]), int(a["qe"]), int(a["rb"]), int(a["re"]), int(a["truesc"]), int(a["w"]))
        o = got[k]
        mine = flat[int(o["cigar_off"]):int(o["cigar_off"]) + int(o["n_cigar"])]
        assert (int(o["pos"]), int(o["rid"]), int(o["is_rev"]), int(o["nm"]), int(o["score"]), int(o["n_waves"])) == \
               (int(want["pos"]), int(want["rid"]), int(want["is_rev"]), int(want["nm"]), int(want["score"]), int(want["n_waves"])), k
        assert mine.size == wc.size and (mine == wc).all(), k
        n_gapped += int(((mine & 0xf) == 1).any() or ((mine & 0xf) == 2).any())
    assert n_gapped > 20          # some alignments do carry indels
