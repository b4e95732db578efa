"""ctypes bindings for the CPU oracle (oracle/liboracle.so) and, when present, the
reference's own CPU functions compiled into oracle/_ref/ (see oracle/build_ref.sh).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product path never touches it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
i8p = np.ctypeslib.ndpointer(np.int8, flags="C_CONTIGUOUS")


class FmdIndex(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("L2", C.c_uint64 * 5), ("seq_len", C.c_uint64),
                ("n_words", C.c_uint64), ("bwt", C.POINTER(C.c_uint32)), ("sa_intv", C.c_int),
                ("n_sa", C.c_uint64), ("sa", C.POINTER(C.c_uint32)), ("sa_hi", C.POINTER(C.c_uint32)),
                ("pack_size", C.c_int), ("owns", C.c_int)]


class FmdCounters(C.Structure):
    _fields_ = [("n_extend", C.c_uint64), ("n_bucket", C.c_uint64), ("n_lf", C.c_uint64),
                ("n_located", C.c_uint64), ("n_smem", C.c_uint64),
                ("n_extend_fwd", C.c_uint64), ("n_bucket_fwd", C.c_uint64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class Reseed(C.Structure):
    """fmd_reseed_t: re-seeding passes 2 and 3 of mem_collect_intv (stock defaults 1.5 / 10 / 20)"""
    _fields_ = [("enable", C.c_int32), ("split_factor", C.c_float), ("split_width", C.c_int32), ("max_mem_intv", C.c_int32)]


def reseed(split_factor=1.5, split_width=10, max_mem_intv=20):
    return Reseed(1, split_factor, split_width, max_mem_intv)


class KswParams(C.Structure):
    _fields_ = [("mat", C.c_int8 * 25), ("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32),
                ("e_ins", C.c_int32), ("w", C.c_int32), ("end_bonus", C.c_int32), ("zdrop", C.c_int32),
                ("use_band", C.c_int32), ("pen_clip", C.c_int32)]


class KswCounters(C.Structure):
    _fields_ = [("cells", C.c_uint64), ("rows", C.c_uint64), ("rect", C.c_uint64)]


def build_oracle() -> str:
    """(Re)build oracle/liboracle.so if missing or stale; returns its path."""
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("fmd_oracle.c", "ksw_oracle.c", "pipeline_oracle.c", "chain_oracle.c", "global_oracle.c", "region_oracle.c", "sw_oracle.c", "sw_oracle.h", "region_oracle.h", "global_oracle.h", "fmd_oracle.h", "ksw_oracle.h", "jobs_common.h", "chain_oracle.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.fmd_load.argtypes = [C.POINTER(FmdIndex), C.c_char_p, C.c_char_p]
        L.fmd_load.restype = C.c_int
        L.fmd_free.argtypes = [C.POINTER(FmdIndex)]
        L.fmd_sa.argtypes = [C.POINTER(FmdIndex), C.c_uint64, C.POINTER(FmdCounters)]
        L.fmd_sa.restype = C.c_uint64
        L.fmd_occ4.argtypes = [C.POINTER(FmdIndex), C.c_uint64, u64p, C.POINTER(FmdCounters)]
        L.fmd_seed_batch.argtypes = [C.POINTER(FmdIndex), u8p, u64p, C.c_int64, C.c_int, C.c_int,
                                     u32p, u64p, u64p, i32p, i32p, u32p, C.c_int64, C.c_int, C.POINTER(FmdCounters)]
        L.fmd_seed_batch.restype = C.c_int64
        L.fmd_smem_batch.argtypes = [C.POINTER(FmdIndex), u8p, u64p, C.c_int64, C.c_int,
                                     u32p, i32p, i32p, u64p, u64p, C.c_int64, C.c_int, C.POINTER(FmdCounters)]
        L.fmd_smem_batch.restype = C.c_int64
        L.fmd_seed_batch_rs.argtypes = [C.POINTER(FmdIndex), u8p, u64p, C.c_int64, C.c_int, C.c_int, C.POINTER(Reseed),
                                        u32p, u64p, u64p, i32p, i32p, u32p, C.c_int64, C.c_int, C.POINTER(FmdCounters)]
        L.fmd_seed_batch_rs.restype = C.c_int64
        L.fmd_smem_batch_rs.argtypes = [C.POINTER(FmdIndex), u8p, u64p, C.c_int64, C.c_int, C.POINTER(Reseed),
                                        u32p, i32p, i32p, u64p, u64p, C.c_int64, C.c_int, C.POINTER(FmdCounters)]
        L.fmd_smem_batch_rs.restype = C.c_int64
        L.glb_batch.argtypes = [C.c_int64, u8p, u32p, u32p, u8p, u32p, u32p, u32p, i8p] + [C.c_int] * 4 + [i32p, i32p, u32p, u32p, C.c_int, C.c_int,
                                C.POINTER(C.c_uint64)]
        L.glb_band.argtypes = [i8p] + [C.c_int] * 6 + [C.c_int64]
        L.glb_band.restype = C.c_int
        L.glb_gen_cigar2.argtypes = [i8p] + [C.c_int] * 5 + [C.c_int64, u8p, C.c_int, u8p, C.c_int64, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     u32p, C.c_int]
        L.glb_gen_cigar2.restype = C.c_int
        L.ksw_fill_mat.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int8)]
        L.ksw_extend_batch_oracle.argtypes = [C.c_int64, u8p, u32p, u32p, u8p, u32p, u32p, u32p,
                                              C.POINTER(KswParams), i32p, C.c_int, C.POINTER(KswCounters)]
        L.pipeline_oracle.argtypes = [C.POINTER(FmdIndex), u8p, C.c_int64, u8p, u64p, C.c_int64, C.c_int, C.c_int,
                                      C.POINTER(KswParams), C.c_int, C.c_void_p, C.c_int, C.POINTER(FmdCounters), C.POINTER(KswCounters)]
        _lib = L
    return _lib


READ_RESULT_DTYPE = np.dtype([("seed_rbeg", "<i8"), ("seed_qbeg", "<i4"), ("seed_qend", "<i4"), ("n_seeds", "<i4"), ("h0", "<i4"),
                              ("left", "<i4", (6,)), ("right", "<i4", (6,))])


def pipeline(oi, fwd: np.ndarray, reads: np.ndarray, read_off: np.ndarray, params, min_seed_len=19, max_occ=500, n_threads=None):
    """CPU statement of the fused seed -> extend pass.  Returns (records, fmd counters, ksw counters)."""
    n = read_off.size - 1
    out = np.zeros(max(n, 1), READ_RESULT_DTYPE)
    fc, kc = FmdCounters(), KswCounters()
    a = int(params.mat[0])
    lib().pipeline_oracle(C.byref(oi.idx), np.ascontiguousarray(fwd), fwd.size, np.ascontiguousarray(reads),
                          np.ascontiguousarray(read_off), n, min_seed_len, max_occ, C.byref(params), a,
                          out.ctypes.data, n_threads or default_threads(), C.byref(fc), C.byref(kc))
    return out[:n], fc.as_dict(), dict(cells=int(kc.cells), rows=int(kc.rows), rect=int(kc.rect))


def default_threads() -> int:
    return max(1, len(os.sched_getaffinity(0)))


def make_params(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, w=100, end_bonus=5, zdrop=100,
                use_band=1, pen_clip=5) -> KswParams:
    p = KswParams()
    lib().ksw_fill_mat(a, b, p.mat)
    p.o_del, p.e_del, p.o_ins, p.e_ins = o_del, e_del, o_ins, e_ins
    p.w, p.end_bonus, p.zdrop, p.use_band, p.pen_clip = w, end_bonus, zdrop, use_band, pen_clip
    return p


class OracleIndex:
    def __init__(self, bwt_path: str, sa_path: str | None):
        self.idx = FmdIndex()
        rc = lib().fmd_load(C.byref(self.idx), bwt_path.encode(), sa_path.encode() if sa_path else None)
        if rc != 0:
            raise RuntimeError(f"fmd_load({bwt_path}, {sa_path}) failed: {rc}")

    def close(self):
        lib().fmd_free(C.byref(self.idx))

    @property
    def seq_len(self):
        return int(self.idx.seq_len)

    def sa(self, k: int) -> int:
        return int(lib().fmd_sa(C.byref(self.idx), k, None))

    def occ4(self, k: int):
        out = np.zeros(4, np.uint64)
        lib().fmd_occ4(C.byref(self.idx), k & 0xFFFFFFFFFFFFFFFF, out, None)
        return out

    def seed_batch(self, reads: np.ndarray, read_off: np.ndarray, min_seed_len=19, max_occ=500,
                   n_threads=None, cap=None, rs=None):
        """reads: flat uint8 codes; read_off: uint64[n+1].  Returns dict of arrays + counters."""
        n = read_off.size - 1
        n_threads = n_threads or default_threads()
        cap = cap or max(1024, int(reads.size) * 4)
        while True:
            n_seeds = np.zeros(max(n, 1), np.uint32)
            off = np.zeros(max(n, 1), np.uint64)
            rbeg = np.zeros(cap, np.uint64)
            qbeg = np.zeros(cap, np.int32)
            qend = np.zeros(cap, np.int32)
            score = np.zeros(cap, np.uint32)
            cnt = FmdCounters()
            tot = lib().fmd_seed_batch_rs(C.byref(self.idx), np.ascontiguousarray(reads), np.ascontiguousarray(read_off),
                                          n, min_seed_len, max_occ, C.byref(rs) if rs is not None else None,
                                          n_seeds, off, rbeg, qbeg, qend, score, cap, n_threads, C.byref(cnt))
            if tot >= 0:
                break
            cap *= 4
        return dict(n_seeds=n_seeds[:n], seed_off=off[:n], rbeg=rbeg[:tot], qbeg=qbeg[:tot], qend=qend[:tot],
                    score=score[:tot], total=int(tot), counters=cnt.as_dict())

    def smem_batch(self, reads: np.ndarray, read_off: np.ndarray, min_seed_len=19, cap=None, rs=None):
        n = read_off.size - 1
        cap = cap or max(1024, int(reads.size) * (4 if rs is not None else 1))
        n_smems = np.zeros(max(n, 1), np.uint32)
        qbeg = np.zeros(cap, np.int32)
        qend = np.zeros(cap, np.int32)
        k = np.zeros(cap, np.uint64)
        s = np.zeros(cap, np.uint64)
        cnt = FmdCounters()
        tot = lib().fmd_smem_batch_rs(C.byref(self.idx), np.ascontiguousarray(reads), np.ascontiguousarray(read_off), n,
                                      min_seed_len, C.byref(rs) if rs is not None else None, n_smems, qbeg, qend, k, s, cap, 1, C.byref(cnt))
        assert tot >= 0
        return dict(n_smems=n_smems[:n], qbeg=qbeg[:tot], qend=qend[:tot], k=k[:tot], s=s[:tot],
                    counters=cnt.as_dict())


def _mat(params):
    return np.frombuffer(bytes(params.mat), dtype=np.int8).copy()


def global_batch(jobs: dict, params: KswParams, cig_stride=64, n_threads=None):
    """ksw_global2 + NM over a batch (oracle/global_oracle.c).  jobs: qseq tseq qoff toff qlen tlen w.  Returns dict(score, nm,
    n_cigar, cigar[n, cig_stride], cells)."""
    n = jobs["qlen"].size
    score = np.zeros(n, np.int32); nm = np.zeros(n, np.int32); nc = np.zeros(n, np.uint32)
    cig = np.zeros(max(n, 1) * cig_stride, np.uint32)
    cells = C.c_uint64(0)
    lib().glb_batch(n, jobs["qseq"], jobs["qoff"], jobs["qlen"], jobs["tseq"], jobs["toff"], jobs["tlen"], np.ascontiguousarray(jobs["w"], np.uint32),
                    _mat(params), params.o_del, params.e_del, params.o_ins, params.e_ins, score, nm, nc, cig, cig_stride,
                    n_threads or default_threads(), C.byref(cells))
    return dict(score=score, nm=nm, n_cigar=nc, cigar=cig.reshape(-1, cig_stride)[:n], cells=int(cells.value))


def global_band(params: KswParams, w_, l_query, rlen):
    return int(lib().glb_band(_mat(params), params.o_del, params.e_del, params.o_ins, params.e_ins, int(w_), int(l_query), int(rlen)))


def gen_cigar2(params: KswParams, w_, fwd, query, rb, re, cap=256):
    """bwa_gen_cigar2 restated over forward reference codes: (score, nm, cigar[n_cigar]) or None when rejected"""
    sc, nm = C.c_int(0), C.c_int(0)
    cig = np.zeros(cap, np.uint32)
    q = np.ascontiguousarray(query, np.uint8)
    n = lib().glb_gen_cigar2(_mat(params), params.o_del, params.e_del, params.o_ins, params.e_ins, int(w_), fwd.size, fwd, q.size, q, int(rb), int(re),
                             C.byref(sc), C.byref(nm), cig, cap)
    if n < 0:
        return None
    return int(sc.value), int(nm.value), cig[:n].copy()


def ksw_batch(jobs: dict, params: KswParams, n_threads=None):
    """jobs: dict from tools.synth.make_ext_jobs.  Returns (res[n,6] int32, counters dict).
    Columns: score, qle, tle, gtle, gscore, max_off."""
    n = jobs["qlen"].size
    res = np.zeros((n, 6), np.int32)
    cnt = KswCounters()
    lib().ksw_extend_batch_oracle(n, jobs["qseq"], jobs["qoff"], jobs["qlen"], jobs["tseq"], jobs["toff"],
                                  jobs["tlen"], jobs["h0"], C.byref(params), res.reshape(-1),
                                  n_threads or default_threads(), C.byref(cnt))
    return res, dict(cells=int(cnt.cells), rows=int(cnt.rows), rect=int(cnt.rect))


def gasal_triple(res6: np.ndarray, qlen: np.ndarray, pen_clip: int):
    """local-vs-to-end rule of src/bwamem.c:1892-1901 -> (score, query_end, target_end)."""
    sc, qle, tle, gtle, gsc = (res6[:, i] for i in range(5))
    local = (gsc <= 0) | (gsc <= sc - pen_clip)
    return (np.where(local, sc, gsc).astype(np.int32), np.where(local, qle, qlen.astype(np.int32)).astype(np.int32),
            np.where(local, tle, gtle).astype(np.int32))


# ------------------------------------------------------------------ reference (_ref)

def have_ref() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in ("libbwaref.so", "libforkksw.so", "bwa7", "bwa6"))


_ref = None
_fork = None


def ref_lib():
    global _ref
    if _ref is None:
        L = C.CDLL(os.path.join(REF_DIR, "libbwaref.so"), mode=os.RTLD_LOCAL)
        L.ref_load.argtypes = [C.c_char_p, C.c_char_p]
        L.ref_load.restype = C.c_void_p
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_smem1.argtypes = [C.c_void_p, C.c_int, u8p, C.c_int, C.c_int, u64p, C.POINTER(C.c_int)]
        L.ref_smem1.restype = C.c_int
        L.ref_sa.argtypes = [C.c_void_p, C.c_uint64]
        L.ref_sa.restype = C.c_uint64
        L.ref_occ4.argtypes = [C.c_void_p, C.c_uint64, u64p]
        L.ref_extend.argtypes = [C.c_void_p, u64p, u64p, C.c_int]
        L.ref_ksw_extend2.argtypes = [C.c_int, u8p, C.c_int, u8p, i8p] + [C.c_int] * 8 + [i32p]
        L.ref_ksw_extend2.restype = C.c_int
        L.ref_seed_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.ref_seed_batch.restype = C.c_int64
        L.ref_ksw_global2.argtypes = [C.c_int, u8p, C.c_int, u8p, i8p] + [C.c_int] * 5 + [C.POINTER(C.c_int), u32p, C.c_int]
        L.ref_ksw_global2.restype = C.c_int
        L.ref_gen_cigar2.argtypes = [i8p] + [C.c_int] * 5 + [C.c_int64, u8p, C.c_int, u8p, C.c_int64, C.c_int64, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                     u32p, C.c_int]
        L.ref_gen_cigar2.restype = C.c_int
        L.ref_collect_intv.argtypes = [C.c_void_p, C.c_int, u8p, C.c_int, C.c_float, C.c_int, C.c_int, u64p, C.c_int]
        L.ref_collect_intv.restype = C.c_int
        L.ref_collect_batch.argtypes = [C.c_void_p, u8p, u64p, C.c_int64, C.c_int, C.c_float, C.c_int, C.c_int, u32p, u64p, C.c_int64]
        L.ref_collect_batch.restype = C.c_int64
        L.ref_ksw_batch.argtypes = [C.c_int64, u8p, u32p, u32p, u8p, u32p, u32p, u32p, i8p] + [C.c_int] * 7 + [i32p, C.c_int]
        L.ref_pipeline_batch.argtypes = [C.c_void_p, u8p, C.c_int64, u8p, u64p, C.c_int64, C.c_int, C.c_int, i8p] + [C.c_int] * 8 + [C.c_void_p, C.c_int]
        _ref = L
    return _ref


def ref_collect_batch(handle, reads, read_off, min_seed_len=19, split_factor=1.5, split_width=10, max_mem_intv=20, cap=None):
    """the reference's own mem_collect_intv (bwa_index/bwamem.c:114-162) over a batch: (n_smems[n], intervals[tot, 5] = x0 x1 x2 start end).
    split_width < 0 and max_mem_intv = 0 reduce it to pass 1."""
    n = read_off.size - 1
    cap = cap or max(1024, int(reads.size) * 4)
    n_smems = np.zeros(max(n, 1), np.uint32)
    out = np.zeros(cap * 5, np.uint64)
    tot = ref_lib().ref_collect_batch(handle, np.ascontiguousarray(reads), np.ascontiguousarray(read_off), n, min_seed_len, split_factor,
                                      split_width, max_mem_intv, n_smems, out, cap)
    assert tot >= 0
    return n_smems[:n], out[:tot * 5].reshape(-1, 5)


def ref_pipeline(handle, fwd, reads, read_off, params, min_seed_len=19, max_occ=500, n_threads=None):
    """the fused seed -> extend pass run with the REFERENCE's own bwt_smem1 / bwt_sa / ksw_extend2"""
    n = read_off.size - 1
    out = np.zeros(max(n, 1), READ_RESULT_DTYPE)
    mat = np.frombuffer(bytes(params.mat), dtype=np.int8).copy()
    ref_lib().ref_pipeline_batch(handle, np.ascontiguousarray(fwd), fwd.size, np.ascontiguousarray(reads),
                                 np.ascontiguousarray(read_off), n, min_seed_len, max_occ, mat,
                                 params.o_del, params.e_del, params.o_ins, params.e_ins, params.w, params.end_bonus,
                                 params.zdrop, int(params.mat[0]), out.ctypes.data, n_threads or default_threads())
    return out[:n]


def fork_lib():
    global _fork
    if _fork is None:
        L = C.CDLL(os.path.join(REF_DIR, "libforkksw.so"), mode=os.RTLD_LOCAL)
        L.fork_ksw_extend2.argtypes = [C.c_int, u8p, C.c_int, u8p, i8p] + [C.c_int] * 9 + [i32p]
        L.fork_ksw_extend2.restype = C.c_int
        _fork = L
    return _fork


def ref_build_index(fasta: str, prefix: str, sa_intv: int = 16) -> None:
    """The two passes of the reference's build_index.sh:46-66 with the binaries in oracle/_ref.
    Leaves prefix.bwt (GPU layout), prefix.bwt128 (stock CPU layout), prefix.sa, .pac, .ann, .amb."""
    env = dict(os.environ)
    subprocess.check_call([os.path.join(REF_DIR, "bwa7"), "index", "-s", "sa", "-r", str(sa_intv), "-p", prefix, fasta],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
    os.replace(prefix + ".bwt", prefix + ".bwt128")
    subprocess.check_call([os.path.join(REF_DIR, "bwa6"), "index", "-s", "bwt", "-p", prefix, fasta],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=env)
    os.remove(prefix + ".bwt1")


# ------------------------------------------------------------------ ksw_align2 (mate rescue / mem_seed_sw)
SW_XBYTE, SW_XSTOP, SW_XSUBO, SW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000
SW_RES_DT = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4")])


def _sw_args(jobs, params):
    mat = np.frombuffer(bytes(params.mat), dtype=np.int8).copy()
    a = [np.ascontiguousarray(jobs[k], dtype=(np.uint8 if k in ("qseq", "tseq") else np.uint32)) for k in ("qseq", "qoff", "qlen", "tseq", "toff", "tlen", "xtra")]
    return mat, a


def sw_align2_batch(jobs: dict, params: KswParams, n_threads=None):
    """oracle restatement of ksw_align2 over a batch: jobs = dict(qseq, qoff, qlen, tseq, toff, tlen, xtra); returns SW_RES_DT array"""
    L = lib()
    if not getattr(L, "_sw", False):
        L.sw_align2_batch_oracle.argtypes = [C.c_int64, u8p, u32p, u32p, u8p, u32p, u32p, u32p, C.c_int, i8p] + [C.c_int] * 4 + [C.c_void_p, C.c_int]
        L._sw = True
    mat, (qs, qo, ql, ts, to, tl, xt) = _sw_args(jobs, params)
    n = ql.size
    res = np.zeros(max(n, 1), SW_RES_DT)
    L.sw_align2_batch_oracle(n, qs, qo, ql, ts, to, tl, xt, 5, mat, params.o_del, params.e_del, params.o_ins, params.e_ins, res.ctypes.data, n_threads or default_threads())
    return res[:n]


def fork_sw_align2_batch(jobs: dict, params: KswParams):
    """the fork's own ksw_align2 (SSE2 ksw_u8 / ksw_i16, oracle/_ref/libforkksw.so) over the same batch"""
    L = fork_lib()
    if not getattr(L, "_sw", False):
        L.fork_ksw_align2_batch.argtypes = [C.c_int64, u8p, u32p, u32p, u8p, u32p, u32p, u32p, C.c_int, i8p] + [C.c_int] * 4 + [C.c_void_p]
        L._sw = True
    mat, (qs, qo, ql, ts, to, tl, xt) = _sw_args(jobs, params)
    n = ql.size
    res = np.zeros(max(n, 1), SW_RES_DT)
    L.fork_ksw_align2_batch(n, qs, qo, ql, ts, to, tl, xt, 5, mat, params.o_del, params.e_del, params.o_ins, params.e_ins, res.ctypes.data)
    return res[:n]
