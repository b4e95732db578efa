"""oracle/region_py.py -- TEST INFRASTRUCTURE ONLY.  ctypes bindings of oracle/region_oracle.c (restatement of mem_sort_dedup_patch /
mem_mark_primary_se / mem_approx_mapq_se) and of the same stage in the reference fork's own code (fork_finish_regs in
oracle/_ref/libforkmem.so), plus the synthetic region sets both are run on."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import chain_py as CP
from . import oracle_py as O

REGION_DT = np.dtype([("rb", "<i8"), ("re", "<i8"), ("hash", "<u8"), ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("score", "<i4"),
                      ("truesc", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"),
                      ("seedcov", "<i4"), ("secondary", "<i4"), ("secondary_all", "<i4"), ("seedlen0", "<i4"), ("n_comp", "<i4"),
                      ("is_alt", "<i4"), ("frac_rep", "<f4"), ("mapq", "<i4")], align=True)
assert REGION_DT.itemsize == 96


def equal(a, b) -> bool:
    """every field, hash (set from the read's id by the primary marking) included"""
    return len(a) == len(b) and all(bool((a[k] == b[k]).all()) for k in REGION_DT.names)


class RegionOpt(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "w", "min_seed_len", "max_chain_gap", "mapQ_coef_fac")] + \
               [("mask_level", C.c_float), ("mask_level_redun", C.c_float), ("mapQ_coef_len", C.c_float)]


_vp = C.c_void_p
_bound = False
_fork_bound = False


def _lib():
    global _bound
    L = O.lib()
    if not _bound:
        L.region_opt_default.argtypes = [C.POINTER(RegionOpt)]
        L.region_finish_read.argtypes = [C.POINTER(RegionOpt), C.c_int64, _vp, _vp, _vp, C.c_int, _vp, C.c_int64, C.POINTER(C.c_int)]
        L.region_finish_read.restype = C.c_int
        _bound = True
    return L


def default_opt(**kw) -> RegionOpt:
    o = RegionOpt()
    _lib().region_opt_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def oracle_finish(opt, ctg: CP.Contigs, fwd, query, regs, rid_id):
    """regs: REGION_DT array of one read -> (finished regions, n_pri)"""
    a = np.ascontiguousarray(regs.copy())
    n_pri = C.c_int(0)
    q = np.ascontiguousarray(query, dtype=np.uint8)
    n = _lib().region_finish_read(C.byref(opt), ctg.l_pac, CP._ptr(ctg.alt), CP._ptr(fwd), CP._ptr(q), len(a), CP._ptr(a), rid_id, C.byref(n_pri))
    return a[:n], int(n_pri.value)


class _quiet_stderr:
    """the fork's mem_mark_primary_se prints debugging text to stderr on its ALT paths (src/bwamem.c:731,736)"""

    def __enter__(self):
        self.saved = os.dup(2)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 2)

    def __exit__(self, *exc):
        os.dup2(self.saved, 2)
        os.close(self.null)
        os.close(self.saved)


def fork_finish(opt, ctg: CP.Contigs, pac, query, regs, rid_id):
    global _fork_bound
    L = CP.fork_lib()
    if not _fork_bound:
        L.fork_finish_regs.argtypes = [C.POINTER(RegionOpt), C.c_int64, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int64, C.POINTER(C.c_int)]
        L.fork_finish_regs.restype = C.c_int
        _fork_bound = True
    a = np.ascontiguousarray(regs.copy())
    n_pri = C.c_int(0)
    q = np.ascontiguousarray(query, dtype=np.uint8)
    with _quiet_stderr():
        n = L.fork_finish_regs(C.byref(opt), ctg.l_pac, ctg.n, CP._ptr(ctg.alt), CP._ptr(pac), len(q), CP._ptr(q), len(a), CP._ptr(a), rid_id, C.byref(n_pri))
    return a[:n], int(n_pri.value)


def _rid(ctg: CP.Contigs, rb, re):
    l = ctg.l_pac
    p = rb if rb < l else 2 * l - re          # forward coordinate of the leftmost base
    return int(np.searchsorted(ctg.off, p, side="right") - 1)


def make_cases(g, ctg: CP.Contigs, reads, pos, strand, seed, big_every=25):
    """Per read a set of alignment regions built to reach every branch of the stage: the true locus; shifted copies of it (redundant
    either way); the locus split in two colinear pieces with a small gap / overlap / indel between them (patch candidates that merge
    or are refused on bandwidth, relative bandwidth or score); exact duplicates; hits elsewhere, also on ALT contigs (secondary
    marking, sub / sub_n, alt_sc); now and then a few dozen regions so the sorts leave their insertion-sort range."""
    rng = np.random.default_rng(seed)
    l = ctg.l_pac
    out = []
    for i, q in enumerate(reads):
        L = len(q)
        regs = []

        def add(qb, qe, rb, re, score, **kw):
            if not (0 <= rb < re <= 2 * l) or (rb < l < re) or qb >= qe:
                return
            r = np.zeros((), REGION_DT)
            r["rb"], r["re"], r["qb"], r["qe"], r["score"] = rb, re, qb, qe, score
            r["truesc"] = kw.get("truesc", score)
            r["rid"] = _rid(ctg, rb, re)
            r["w"] = kw.get("w", int(rng.choice([0, 3, 20, 100])))
            r["seedcov"] = int(rng.integers(19, qe - qb + 20))
            r["csub"] = int(rng.choice([0, 0, 0, 15, 40, score]))
            r["seedlen0"] = int(rng.integers(19, 60))
            r["frac_rep"] = float(rng.choice([0.0, 0.0, 0.0, 0.25, 0.8]))
            r["secondary"] = -1
            r["sub"] = int(rng.choice([0, 0, 22]))
            regs.append(r)

        p0 = int(pos[i])
        base = p0 if strand[i] == 0 else 2 * l - (p0 + L)
        qb, qe = int(rng.choice([0, 0, 2, 9])), L - int(rng.choice([0, 0, 3, 12]))
        kind = int(rng.integers(0, 6))
        if kind != 1:                                        # the locus as one region
            add(qb, qe, base + qb, base + qe, qe - qb - int(rng.choice([0, 5, 10, 30])))
        if kind in (1, 2):                                   # two colinear pieces
            m1 = int(rng.integers(30, L - 30))
            gap_q = int(rng.choice([-8, -2, 0, 0, 3, 12]))
            d = int(rng.choice([0, 0, 0, 1, -1, 3, -4, 12, 60, 500]))
            m2 = m1 + gap_q
            add(qb, m1, base + qb, base + m1, m1 - qb - int(rng.choice([0, 4, 9])))
            add(m2, qe, base + m2 + d, base + qe + d, qe - m2 - int(rng.choice([0, 4, 9])))
        if kind in (3, 4):                                   # shifted / shrunk copies: redundant hits
            for _ in range(int(rng.integers(1, 4))):
                s1, s2, dr = int(rng.integers(0, 6)), int(rng.integers(0, 6)), int(rng.integers(-2, 3))
                add(qb + s1, qe - s2, base + qb + s1 + dr, base + qe - s2 + dr, qe - qb - int(rng.choice([0, 5, 10, 30, 31])))
        if kind == 5 and regs:                               # exact duplicates
            regs.append(regs[0].copy())
            regs.append(regs[0].copy())
        n_else = int(rng.integers(20, 45)) if (i % big_every == big_every - 1) else int(rng.choice([0, 0, 1, 2, 4]))
        for _ in range(n_else):                              # hits elsewhere
            a_qb = int(rng.integers(0, L - 40)); a_qe = int(rng.integers(a_qb + 25, L + 1))
            ln = a_qe - a_qb + int(rng.integers(-2, 3))
            rb = int(rng.integers(0, 2 * l - ln))
            add(a_qb, a_qe, rb, rb + ln, int(rng.integers(19, a_qe - a_qb + 1)))
        if regs:
            rng.shuffle(regs)
        out.append(np.array(regs, REGION_DT) if regs else np.zeros(0, REGION_DT))
    return out
