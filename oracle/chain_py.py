"""ctypes bindings for the chaining / extension-job oracle (oracle/chain_oracle.c) and for the
reference fork's own host code compiled into oracle/_ref/libforkmem.so (oracle/fork_mem_shim.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import oracle_py as O

CHAIN_DT = np.dtype([("pos", "<i8"), ("rid", "<i4"), ("n", "<i4"), ("w", "<i4"), ("kept", "<i4"), ("first", "<i4"),
                     ("is_alt", "<i4"), ("frac_rep", "<f4"), ("seed_off", "<i4")], align=True)
CSEED_DT = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4"), ("score", "<i4"), ("pad", "<i4")], align=True)
REG_DT = np.dtype([("rb_est", "<i8"), ("re_est", "<i8"), ("target_seed_begin", "<i8"), ("qb_est", "<i4"), ("qe_est", "<i4"),
                   ("rid", "<i4"), ("score", "<i4"), ("truesc", "<i4"), ("align_sides", "<i4"), ("where_is_long", "<i4"),
                   ("query_seed_begin", "<i4"), ("seedlen0", "<i4"), ("seedcov", "<i4"), ("w", "<i4"), ("frac_rep", "<f4")], align=True)
JOB_DT = np.dtype([("qoff", "<u4"), ("qlen", "<u4"), ("toff", "<u4"), ("tlen", "<u4"), ("h0", "<u4")], align=True)
ALN_DT = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("score", "<i4"), ("truesc", "<i4")], align=True)
assert CHAIN_DT.itemsize == 40 and CSEED_DT.itemsize == 24 and REG_DT.itemsize == 72 and JOB_DT.itemsize == 20 and ALN_DT.itemsize == 32


class ChainOpt(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "w", "min_seed_len", "max_occ",
                                         "max_chain_gap", "min_chain_weight", "max_chain_extend")] + \
               [("mask_level", C.c_float), ("drop_ratio", C.c_float)]


def default_opt(**kw) -> ChainOpt:
    o = ChainOpt()
    O.lib().chain_opt_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class Contigs:
    """offsets / lengths / is_alt of the reference sequences (bntseq_t anns)"""

    def __init__(self, lens, alt=None):
        self.len = np.ascontiguousarray(lens, dtype=np.int32)
        self.off = np.ascontiguousarray(np.concatenate([[0], np.cumsum(self.len[:-1], dtype=np.int64)]), dtype=np.int64)
        self.alt = np.ascontiguousarray(alt if alt is not None else np.zeros(len(self.len)), dtype=np.int32)
        self.n = len(self.len)
        self.l_pac = int(self.len.astype(np.int64).sum())


_vp = C.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _bind_oracle():
    L = O.lib()
    if getattr(L, "_chain_bound", False):
        return L
    L.chain_opt_default.argtypes = [C.POINTER(ChainOpt)]
    L.chain_oracle_read.argtypes = [C.POINTER(ChainOpt), C.c_int64, C.c_int, _vp, _vp, _vp, C.c_int, C.c_uint32, _vp, _vp, _vp, C.c_int,
                                    C.POINTER(C.c_int32), _vp, _vp]
    L.chain_oracle_read.restype = C.c_int
    L.chain_oracle_read_any.argtypes = L.chain_oracle_read.argtypes
    L.chain_oracle_read_any.restype = C.c_int
    L.chain_oracle_flt_seeds.argtypes = [C.POINTER(ChainOpt), C.c_int64, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp]
    L.chain_oracle_flt_seeds.restype = C.c_int
    L.chain2aln_oracle_read.argtypes = [C.POINTER(ChainOpt), C.c_int64, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp,
                                        C.POINTER(C.c_int32), _vp, C.c_int, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_uint64]
    L.chain2aln_oracle_read.restype = C.c_int
    L.chain_regs_finish.argtypes = [C.c_int, C.c_int, _vp, _vp, _vp, _vp]
    L._chain_bound = True
    return L


def oracle_chains(opt, ctg: Contigs, l_query, rbeg, qq, score, layout_all, fwd=None, query=None):
    """mem_chain + mem_chain_flt of one read -> (chains[CHAIN_DT], seeds[CSEED_DT]); with the reference and the read given, a read long enough
    for mem_flt_chained_seeds to run goes through it as well (mem_seed_sw on every chained seed), otherwise such a read raises"""
    L = _bind_oracle()
    n = len(rbeg)
    rbeg = np.ascontiguousarray(rbeg, dtype=np.uint64); qq = np.ascontiguousarray(qq, dtype=np.int32); score = np.ascontiguousarray(score, dtype=np.uint32)
    chains = np.zeros(max(n, 1), dtype=CHAIN_DT); cs = np.zeros(max(n, 1), dtype=CSEED_DT)
    nc = C.c_int32(0)
    args = (C.byref(opt), ctg.l_pac, ctg.n, _ptr(ctg.off), _ptr(ctg.len), _ptr(ctg.alt), l_query, n, _ptr(rbeg), _ptr(qq),
            _ptr(score), int(layout_all), C.byref(nc), _ptr(chains), _ptr(cs))
    rc = L.chain_oracle_read(*args)
    if rc == -2 and fwd is not None and query is not None:
        rc = L.chain_oracle_read_any(*args)
        if rc == 0:
            q = np.ascontiguousarray(query, dtype=np.uint8)
            ran = L.chain_oracle_flt_seeds(C.byref(opt), ctg.l_pac, ctg.n, _ptr(ctg.off), _ptr(ctg.len), _ptr(fwd), l_query, _ptr(q), nc.value, _ptr(chains), _ptr(cs))
            assert ran == 1
    if rc:
        raise RuntimeError(f"chain_oracle_read rc={rc}")
    chains = chains[:nc.value]
    ns = int(chains["n"].sum()) if nc.value else 0
    return chains, cs[:ns]


def oracle_chain2aln(opt, ctg: Contigs, fwd, query, chains, cseeds, cap=4096, cap_bytes=1 << 22):
    """mem_chain2aln over the chains of one read -> (regs[REG_DT], [jobs_short, jobs_long], [(q, t) bytes per side])"""
    L = _bind_oracle()
    regs = np.zeros(cap, dtype=REG_DT)
    jobs = [np.zeros(cap, dtype=JOB_DT), np.zeros(cap, dtype=JOB_DT)]
    qb = [np.zeros(cap_bytes, dtype=np.uint8), np.zeros(cap_bytes, dtype=np.uint8)]
    tb = [np.zeros(cap_bytes, dtype=np.uint8), np.zeros(cap_bytes, dtype=np.uint8)]
    qpp = (_vp * 2)(_ptr(qb[0]), _ptr(qb[1])); tpp = (_vp * 2)(_ptr(tb[0]), _ptr(tb[1]))
    nr = C.c_int32(0); nj = (C.c_int32 * 2)(0, 0)
    chains = np.ascontiguousarray(chains); cseeds = np.ascontiguousarray(cseeds)
    query = np.ascontiguousarray(query, dtype=np.uint8)
    rc = L.chain2aln_oracle_read(C.byref(opt), ctg.l_pac, ctg.n, _ptr(ctg.off), _ptr(ctg.len), _ptr(fwd), len(query), _ptr(query),
                                 len(chains), _ptr(chains), _ptr(cseeds), C.byref(nr), _ptr(regs), cap, nj, _ptr(jobs[0]), _ptr(jobs[1]), cap,
                                 qpp, tpp, cap_bytes)
    if rc:
        raise RuntimeError(f"chain2aln_oracle_read rc={rc}")
    out_jobs = [jobs[s][:nj[s]] for s in (0, 1)]
    seqs = []
    for s in (0, 1):
        j = out_jobs[s]
        nq = int(j["qoff"][-1] + (j["qlen"][-1] + 7) // 8 * 8) if len(j) else 0
        nt = int(j["toff"][-1] + (j["tlen"][-1] + 7) // 8 * 8) if len(j) else 0
        seqs.append((qb[s][:nq].copy(), tb[s][:nt].copy()))
    return regs[:nr.value], out_jobs, seqs


def oracle_regs_finish(l_query, regs, short_triples, long_triples):
    L = _bind_oracle()
    out = np.zeros(max(len(regs), 1), dtype=ALN_DT)
    st = np.ascontiguousarray(short_triples, dtype=np.int32).reshape(-1); lt = np.ascontiguousarray(long_triples, dtype=np.int32).reshape(-1)
    if st.size == 0:
        st = np.zeros(3, np.int32)
    if lt.size == 0:
        lt = np.zeros(3, np.int32)
    regs = np.ascontiguousarray(regs)
    L.chain_regs_finish(l_query, len(regs), _ptr(regs), _ptr(st), _ptr(lt), _ptr(out))
    return out[:len(regs)]


# ------------------------------------------------------------------ the reference fork's own code
_fork = None


def have_fork() -> bool:
    return os.path.exists(os.path.join(O.REF_DIR, "libforkmem.so"))


def fork_lib():
    global _fork
    if _fork is None:
        L = C.CDLL(os.path.join(O.REF_DIR, "libforkmem.so"))
        L.fork_mem_read.argtypes = [C.POINTER(ChainOpt), C.c_int64, C.c_int, _vp, _vp, _vp, _vp, C.c_int, _vp, C.c_uint32, _vp, _vp, _vp,
                                    C.POINTER(C.c_int32), _vp, C.c_int, _vp, C.c_int, C.POINTER(C.c_int32), _vp, C.c_int,
                                    C.POINTER(C.c_int32), _vp, _vp, C.c_int]
        L.fork_mem_read.restype = C.c_int
        L.fork_mem_seq.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.fork_mem_seq.restype = C.POINTER(C.c_uint8)
        L.fork_reg2aln.argtypes = [C.POINTER(ChainOpt), C.c_int64, C.c_int, _vp, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int64, C.c_int64,
                                   C.c_int, C.c_int, C.c_int, _vp, _vp, C.c_int]
        L.fork_reg2aln.restype = C.c_int
        _fork = L
    return _fork


def make_pac(fwd) -> np.ndarray:
    """the reference's .pac packing: 4 bases per byte, base l at bits ((~l & 3) << 1) (src/bntseq.c _set_pac)"""
    fwd = np.asarray(fwd, dtype=np.uint8)
    n = len(fwd)
    pad = (-n) % 4
    f = np.concatenate([fwd, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    return np.ascontiguousarray((f[:, 0] << 6) | (f[:, 1] << 4) | (f[:, 2] << 2) | f[:, 3]).astype(np.uint8)


def fork_read(opt, ctg: Contigs, pac, query, rbeg, qq, score, cap=4096):
    """the unmodified fork: mem_chain -> mem_chain_flt -> mem_flt_chained_seeds -> mem_chain2aln per chain.
    Seeds in the reference's full layout (every SMEM group holds all `score` rows)."""
    L = fork_lib()
    n = len(rbeg)
    rbeg = np.ascontiguousarray(rbeg, dtype=np.uint64); qq = np.ascontiguousarray(qq, dtype=np.int32); score = np.ascontiguousarray(score, dtype=np.uint32)
    query = np.ascontiguousarray(query, dtype=np.uint8)
    chains = np.zeros(max(n, 1), dtype=CHAIN_DT); cs = np.zeros(max(n, 1), dtype=CSEED_DT)
    regs = np.zeros(cap, dtype=REG_DT)
    jobs = [np.zeros(cap, dtype=JOB_DT), np.zeros(cap, dtype=JOB_DT)]
    nc = C.c_int32(0); nr = C.c_int32(0); nj = (C.c_int32 * 2)(0, 0)
    rc = L.fork_mem_read(C.byref(opt), ctg.l_pac, ctg.n, _ptr(ctg.off), _ptr(ctg.len), _ptr(ctg.alt), _ptr(pac), len(query), _ptr(query),
                         n, _ptr(rbeg), _ptr(qq), _ptr(score), C.byref(nc), _ptr(chains), len(chains), _ptr(cs), len(cs),
                         C.byref(nr), _ptr(regs), cap, nj, _ptr(jobs[0]), _ptr(jobs[1]), cap)
    if rc:
        raise RuntimeError(f"fork_mem_read rc={rc}")
    chains = chains[:nc.value]
    ns = int(chains["n"].sum()) if nc.value else 0
    seqs = []
    for s in (0, 1):
        pair = []
        for which in (0, 1):
            nb = C.c_uint64(0)
            p = L.fork_mem_seq(s, which, C.byref(nb))
            pair.append(np.ctypeslib.as_array(p, shape=(nb.value,)).copy() if nb.value else np.zeros(0, np.uint8))
        seqs.append(tuple(pair))
    return chains, cs[:ns], regs[:nr.value], [jobs[s][:nj[s]] for s in (0, 1)], seqs


def oracle_align_batch(opt, ctg: Contigs, fwd, reads, rbeg, qq, score, n_seeds, seed_off, layout_all, ksw_params, n_threads=2):
    """The whole seeds -> chains -> jobs -> ksw_extend2 -> regions stage of a read batch on the CPU (the checker of
    bwa_b200_align_*): per read mem_chain / mem_chain_flt / mem_chain2aln, then every job of the batch through the
    ksw oracle in the reference's batch order (SHORT batch, then LONG batch), the local-vs-to-end rule and the
    region arithmetic."""
    per = []
    for r, query in enumerate(reads):
        so, ns = int(seed_off[r]), int(n_seeds[r])
        oc, osd = oracle_chains(opt, ctg, len(query), rbeg[so:so + ns], qq[so:so + ns], score[so:so + ns], layout_all, fwd, query)
        regs, jobs, seqs = oracle_chain2aln(opt, ctg, fwd, query, oc, osd)
        per.append((oc, osd, regs, jobs, seqs))
    out = dict(n_chains=np.array([len(p[0]) for p in per], np.uint32), n_regions=np.array([len(p[2]) for p in per], np.uint32),
               chains=np.concatenate([p[0] for p in per]) if per else np.zeros(0, CHAIN_DT),
               chain_seeds=np.concatenate([p[1] for p in per]) if per else np.zeros(0, CSEED_DT),
               regs=np.concatenate([p[2] for p in per]) if per else np.zeros(0, REG_DT))
    # batch-level job arrays: SHORT batch of every read in read order, then the LONG batch
    jl, ql, tl = [], [], []
    nq = nt = 0
    for side in (0, 1):
        for p in per:
            j = p[3][side].copy()
            j["qoff"] += nq; j["toff"] += nt
            jl.append(j); ql.append(p[4][side][0]); tl.append(p[4][side][1])
            nq += len(p[4][side][0]); nt += len(p[4][side][1])
    jobs = np.concatenate(jl) if jl else np.zeros(0, JOB_DT)
    qseq = np.concatenate(ql) if ql else np.zeros(0, np.uint8)
    tseq = np.concatenate(tl) if tl else np.zeros(0, np.uint8)
    n_short = int(sum(len(p[3][0]) for p in per))
    out.update(jobs=jobs, qseq=qseq, tseq=tseq, n_jobs_short=n_short, n_jobs_long=len(jobs) - n_short)
    if len(jobs):
        jd = dict(qseq=np.ascontiguousarray(qseq if len(qseq) else np.zeros(8, np.uint8)), tseq=np.ascontiguousarray(tseq if len(tseq) else np.zeros(8, np.uint8)),
                  qoff=np.ascontiguousarray(jobs["qoff"]), toff=np.ascontiguousarray(jobs["toff"]),
                  qlen=np.ascontiguousarray(jobs["qlen"]), tlen=np.ascontiguousarray(jobs["tlen"]), h0=np.ascontiguousarray(jobs["h0"]))
        res, cnt = O.ksw_batch(jd, ksw_params, n_threads=n_threads)
        tri = np.stack(O.gasal_triple(res, jobs["qlen"], ksw_params.pen_clip), axis=1).astype(np.int32)
        # what the device extender counts: jobs of the closed-form shape are answered without a matrix (tools/synth.dp_cells)
        from tools import synth
        kw = dict(a=int(ksw_params.mat[0]), b=-int(ksw_params.mat[1]), o_del=ksw_params.o_del, e_del=ksw_params.e_del, o_ins=ksw_params.o_ins,
                  e_ins=ksw_params.e_ins, w=ksw_params.w, zdrop=ksw_params.zdrop, use_band=ksw_params.use_band, end_bonus=ksw_params.end_bonus,
                  pen_clip=ksw_params.pen_clip)
        cells_dp, closed = synth.dp_cells(O, jd, kw, cnt)
    else:
        res, cnt, tri = np.zeros((0, 6), np.int32), dict(cells=0, rows=0, rect=0), np.zeros((0, 3), np.int32)
        cells_dp, closed = 0, 0
    out.update(job_res=res, cells=cnt["cells"], cells_dp=cells_dp, closed_form_jobs=closed)
    alns = []
    i_s, i_l = 0, n_short
    for r, p in enumerate(per):
        ns_, nl_ = len(p[3][0]), len(p[3][1])
        alns.append(oracle_regs_finish(len(reads[r]), p[2], tri[i_s:i_s + ns_], tri[i_l:i_l + nl_]))
        i_s += ns_; i_l += nl_
    out["aln"] = np.concatenate(alns) if alns else np.zeros(0, ALN_DT)
    return out


# ------------------------------------------------------------------ mem_reg2aln (CIGAR stage at the call-site level)
R2A_DT = np.dtype([("pos", "<i8"), ("rid", "<i4"), ("is_rev", "<i4"), ("score", "<i4"), ("nm", "<i4"), ("n_cigar", "<i4"), ("band", "<i4"),
                   ("n_waves", "<i4")], align=True)


def fork_reg2aln(opt, ctg: Contigs, pac, query, qb, qe, rb, re, truesc, ar_w, cap=512):
    """the unmodified fork's mem_reg2aln: dict(pos, rid, is_rev, nm, n_cigar, cigar)"""
    out8 = np.zeros(8, np.int64)
    cig = np.zeros(cap, np.uint32)
    query = np.ascontiguousarray(query, dtype=np.uint8)
    n = fork_lib().fork_reg2aln(C.byref(opt), ctg.l_pac, ctg.n, _ptr(ctg.off), _ptr(ctg.len), _ptr(pac), len(query), _ptr(query), int(qb), int(qe),
                                int(rb), int(re), int(truesc), int(ar_w), 0, _ptr(out8), _ptr(cig), cap)
    assert n <= cap
    return dict(pos=int(out8[0]), rid=int(out8[1]), is_rev=int(out8[2]), nm=int(out8[4]), n_cigar=int(out8[5]), cigar=cig[:n].copy())


def oracle_reg2aln(opt, kp, ctg: Contigs, fwd, query, qb, qe, rb, re, truesc, ar_w, cap=512):
    """oracle/global_oracle.c glb_reg2aln: (R2A_DT record, cigar)"""
    L = O.lib()
    if not getattr(L, "_r2a", False):
        L.glb_reg2aln.argtypes = [_vp] + [C.c_int] * 6 + [C.c_int64, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int,
                                  _vp, _vp, C.c_int]
        L.glb_reg2aln.restype = C.c_int
        L._r2a = True
    mat = np.frombuffer(bytes(kp.mat), dtype=np.int8).copy()
    out = np.zeros(1, R2A_DT)
    cig = np.zeros(cap, np.uint32)
    query = np.ascontiguousarray(query, dtype=np.uint8); fwd = np.ascontiguousarray(fwd, dtype=np.uint8)
    n = L.glb_reg2aln(_ptr(mat), opt.a, opt.o_del, opt.e_del, opt.o_ins, opt.e_ins, opt.w, ctg.l_pac, _ptr(fwd), ctg.n, _ptr(ctg.off), len(query), _ptr(query),
                      int(qb), int(qe), int(rb), int(re), int(truesc), int(ar_w), _ptr(out), _ptr(cig), cap)
    assert n >= 0
    return out[0], cig[:n].copy()


# ------------------------------------------------------------------ CPU arm of the chained step (bench.py, kind "reference")
FORK_ALN_DT = np.dtype([("rb", "<i8"), ("re", "<i8"), ("qb", "<i4"), ("qe", "<i4"), ("score", "<i4"), ("truesc", "<i4"), ("rid", "<i4"),
                        ("seedcov", "<i4"), ("seedlen0", "<i4"), ("w", "<i4"), ("frac_rep", "<f4"), ("pad", "<i4")], align=True)
assert FORK_ALN_DT.itemsize == 56


def ref_seed_arrays(handle, reads, read_off, min_seed_len=19, max_occ=0, n_threads=None):
    """pass-1 SMEM seeds of a batch from the REFERENCE's own bwt_smem1 / bwt_sa (oracle/_ref/libbwaref.so), layout of mem_seed_v_gpu"""
    L = O.ref_lib()
    if not getattr(L, "_seed_arrays", False):
        L.ref_seed_arrays.argtypes = [_vp, _vp, _vp, C.c_int64, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp, _vp, C.c_int64]
        L.ref_seed_arrays.restype = C.c_int64
        L._seed_arrays = True
    n = read_off.size - 1
    reads = np.ascontiguousarray(reads, dtype=np.uint8); read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
    n_seeds = np.zeros(max(n, 1), np.uint32); seed_off = np.zeros(max(n, 1), np.uint64)
    cap = max(1024, 8 * n)
    while True:
        rbeg = np.empty(cap, np.uint64); qq = np.empty(2 * cap, np.int32); score = np.empty(cap, np.uint32)
        tot = L.ref_seed_arrays(handle, _ptr(reads), _ptr(read_off), n, min_seed_len, max_occ, n_threads or O.default_threads(),
                                _ptr(n_seeds), _ptr(seed_off), _ptr(rbeg), _ptr(qq), _ptr(score), cap)
        if tot >= 0:
            break
        cap = -tot + 1024
    return dict(total=int(tot), n_seeds=n_seeds[:n], seed_off=seed_off[:n], rbeg=rbeg[:tot], qq=qq[:2 * tot].reshape(-1, 2), score=score[:tot])


def fork_align_batch(opt, ctg: Contigs, pac, reads, read_off, seeds, ksw_params, n_threads=None):
    """seeds -> chains -> jobs -> ksw_extend2 -> regions of a whole batch with the FORK's own host functions (oracle/_ref/libforkmem.so,
    fork_mem_shim.cpp fork_align_batch).  seeds: dict from ref_seed_arrays (all rows of every SMEM group).  Returns dict(n_regs, reg_off,
    regs[FORK_ALN_DT], n_jobs)."""
    L = fork_lib()
    if not getattr(L, "_align_batch", False):
        L.fork_align_batch.argtypes = [C.POINTER(ChainOpt), C.c_int64, C.c_int, _vp, _vp, _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                       C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int64, C.c_int, C.POINTER(C.c_uint64)]
        L.fork_align_batch.restype = C.c_int64
        L._align_batch = True
    n = read_off.size - 1
    reads = np.ascontiguousarray(reads, dtype=np.uint8); read_off = np.ascontiguousarray(read_off, dtype=np.uint64)
    rbeg = np.ascontiguousarray(seeds["rbeg"], dtype=np.uint64); qq = np.ascontiguousarray(seeds["qq"], dtype=np.int32).reshape(-1)
    score = np.ascontiguousarray(seeds["score"], dtype=np.uint32)
    ns = np.ascontiguousarray(seeds["n_seeds"], dtype=np.uint32); so = np.ascontiguousarray(seeds["seed_off"], dtype=np.uint64)
    if rbeg.size == 0:
        rbeg = np.zeros(1, np.uint64); qq = np.zeros(2, np.int32); score = np.zeros(1, np.uint32)
    n_regs = np.zeros(max(n, 1), np.uint32); reg_off = np.zeros(max(n, 1), np.uint64)
    cap = max(1024, 3 * n)
    nj = C.c_uint64(0)
    while True:
        regs = np.zeros(cap, FORK_ALN_DT)
        tot = L.fork_align_batch(C.byref(opt), ctg.l_pac, ctg.n, _ptr(ctg.off), _ptr(ctg.len), _ptr(ctg.alt), _ptr(pac), n, _ptr(reads), _ptr(read_off),
                                 _ptr(rbeg), _ptr(qq), _ptr(score), _ptr(ns), _ptr(so), ksw_params.w, ksw_params.zdrop, ksw_params.end_bonus,
                                 ksw_params.use_band, ksw_params.pen_clip, _ptr(n_regs), _ptr(reg_off), _ptr(regs), cap, n_threads or O.default_threads(), C.byref(nj))
        if tot >= 0:
            break
        cap = int(n_regs.sum()) + 16
    return dict(n_regs=n_regs[:n], reg_off=reg_off[:n], regs=regs[:tot], n_jobs=int(nj.value))


def ref_chained_pipeline(handle, opt, ctg: Contigs, pac, reads, read_off, ksw_params, min_seed_len=19, n_threads=None):
    """the CPU reference of bwa_b200_align_*: the reference's bwt_smem1 / bwt_sa for the seeds, the fork's mem_chain ... mem_chain2aln and
    ksw_extend2 for everything after, all host threads"""
    seeds = ref_seed_arrays(handle, reads, read_off, min_seed_len, 0, n_threads)
    out = fork_align_batch(opt, ctg, pac, reads, read_off, seeds, ksw_params, n_threads)
    out["n_seeds"] = seeds["total"]
    return out
