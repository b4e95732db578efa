"""Python-side loader for libbwamem_b200.so (ctypes over the C ABI in include/bwamem_b200.h).

This module holds no compute: every call goes straight into the CUDA library.  It exists so
tests and bench.py can drive the same C entry points the `gase_aln` driver would bind.
The directory name contains '-', so import it through `load_package()` in tests/conftest.py /
__graft_entry__.py (importlib by path, module name `bwa_mem_gpu_b200`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
LIB_PATH = os.path.join(PKG_DIR, "libbwamem_b200.so")

ERR = {0: "OK", -1: "ERR_ARG", -2: "ERR_IO", -3: "ERR_FORMAT", -4: "ERR_CUDA", -5: "ERR_NOMEM", -6: "ERR_CAPACITY"}


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERR.get(code, code)}: {msg}")
        self.code = code


def build(force: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    csrc = os.path.join(PKG_DIR, "csrc")
    srcs = [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cu", ".h", ".cuh"))]
    srcs += [os.path.join(ROOT, "include", "bwamem_b200.h")]
    comp = os.path.join(ROOT, "include", "compat")
    if os.path.isdir(comp):
        srcs += [os.path.join(comp, f) for f in os.listdir(comp)]
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        nvcc = "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            if os.path.exists(LIB_PATH):
                return LIB_PATH
            raise RuntimeError("nvcc not found and libbwamem_b200.so is not built")
        subprocess.check_call(["make", "-C", csrc, "-j8"], stdout=subprocess.DEVNULL)
    return LIB_PATH


class IndexInfo(C.Structure):
    _fields_ = [("primary", C.c_uint64), ("seq_len", C.c_uint64), ("L2", C.c_uint64 * 5), ("n_buckets", C.c_uint64),
                ("n_sa", C.c_uint64), ("sa_intv", C.c_int32), ("pack_size", C.c_int32), ("device", C.c_int32),
                ("hbm_bytes", C.c_uint64)]


class SeedParams(C.Structure):
    """bwa_b200_seed_params_t; SeedParams(19, 500) = pass 1 only, SeedParams(19, 500, 1, 1.5, 10, 20) = stock re-seeding too"""
    _fields_ = [("min_seed_len", C.c_int32), ("max_occ", C.c_int32), ("reseed", C.c_int32), ("split_factor", C.c_float),
                ("split_width", C.c_int32), ("max_mem_intv", C.c_int32)]


def seed_params(min_seed_len=19, max_occ=500, reseed=False, split_factor=1.5, split_width=10, max_mem_intv=20):
    return SeedParams(min_seed_len, max_occ, int(bool(reseed)), split_factor, split_width, max_mem_intv)


class Seeds(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_seeds", C.c_uint64), ("rbeg", C.POINTER(C.c_uint64)),
                ("qbeg_qend", C.POINTER(C.c_int32)), ("score", C.POINTER(C.c_uint32)),
                ("n_seeds_per_read", C.POINTER(C.c_uint32)), ("seed_off", C.POINTER(C.c_uint64))]


class ExtParams(C.Structure):
    _fields_ = [("mat", C.c_int8 * 25), ("o_del", C.c_int32), ("e_del", C.c_int32), ("o_ins", C.c_int32),
                ("e_ins", C.c_int32), ("w", C.c_int32), ("end_bonus", C.c_int32), ("zdrop", C.c_int32),
                ("use_band", C.c_int32), ("pen_clip", C.c_int32)]


class ExtResult(C.Structure):
    _fields_ = [("score", C.c_int32), ("qle", C.c_int32), ("tle", C.c_int32), ("gtle", C.c_int32),
                ("gscore", C.c_int32), ("max_off", C.c_int32)]


class ReadResult(C.Structure):
    _fields_ = [("seed_rbeg", C.c_int64), ("seed_qbeg", C.c_int32), ("seed_qend", C.c_int32), ("n_seeds", C.c_int32),
                ("h0", C.c_int32), ("left", ExtResult), ("right", ExtResult)]


READ_RESULT_DTYPE = np.dtype([("seed_rbeg", "<i8"), ("seed_qbeg", "<i4"), ("seed_qend", "<i4"), ("n_seeds", "<i4"), ("h0", "<i4"),
                              ("left", "<i4", (6,)), ("right", "<i4", (6,))])
assert READ_RESULT_DTYPE.itemsize == C.sizeof(ReadResult) == 72


class RegionOpt(C.Structure):
    """bwa_b200_region_opt_t"""
    _fields_ = [(k, C.c_int32) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "w", "min_seed_len", "max_chain_gap", "mapQ_coef_fac")] + \
               [("mask_level", C.c_float), ("mask_level_redun", C.c_float), ("mapQ_coef_len", C.c_float)]


# bwa_b200_alnreg_t
ALNREG_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("hash", "<u8"), ("qb", "<i4"), ("qe", "<i4"), ("rid", "<i4"), ("score", "<i4"),
                         ("truesc", "<i4"), ("sub", "<i4"), ("alt_sc", "<i4"), ("csub", "<i4"), ("sub_n", "<i4"), ("w", "<i4"),
                         ("seedcov", "<i4"), ("secondary", "<i4"), ("secondary_all", "<i4"), ("seedlen0", "<i4"), ("n_comp", "<i4"),
                         ("is_alt", "<i4"), ("frac_rep", "<f4"), ("mapq", "<i4")], align=True)
assert ALNREG_DTYPE.itemsize == 96

# every symbol include/bwamem_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "bwa_b200_last_error", "bwa_b200_version", "bwa_b200_device_count", "bwa_b200_host_alloc", "bwa_b200_host_free",
    "bwa_b200_index_load", "bwa_b200_index_from_host", "bwa_b200_index_clone_to", "bwa_b200_index_info", "bwa_b200_index_set_kmer_table",
    "bwa_b200_index_free", "bwa_b200_build_index", "bwa_b200_packed_words", "bwa_b200_pack_ascii",
    "bwa_b200_pack_codes", "bwa_b200_seeder_create", "bwa_b200_seeder_destroy", "bwa_b200_seed_host",
    "bwa_b200_seeds_free", "bwa_b200_seed_device", "bwa_b200_seed_device_result", "bwa_b200_seeder_stream",
    "bwa_b200_seed_device_smems", "bwa_b200_seeder_launches", "bwa_b200_seeder_request_counts", "bwa_b200_seed_params_default", "bwa_b200_measure_random_sector_gbs", "bwa_b200_measure_int_alu", "bwa_b200_int_alu_op_name", "bwa_b200_ext_params_default", "bwa_b200_fill_scmat",
    "bwa_b200_extender_create", "bwa_b200_extender_destroy", "bwa_b200_extend_async", "bwa_b200_extend_query",
    "bwa_b200_extend_wait", "bwa_b200_extend_async_paged", "bwa_b200_extend_device", "bwa_b200_pack_device", "bwa_b200_extender_stream",
    "bwa_b200_extender_launches", "bwa_b200_extender_last_cells", "bwa_b200_extender_last_closed_form", "bwa_b200_extender_set_closed_form",
    "bwa_b200_index_attach_ref", "bwa_b200_pipeline_create", "bwa_b200_pipeline_destroy", "bwa_b200_seed_extend_host",
    "bwa_b200_seed_extend_device", "bwa_b200_pipeline_sync", "bwa_b200_pipeline_stream", "bwa_b200_pipeline_launches",
    "bwa_b200_pipeline_totals", "bwa_b200_pipeline_profile", "bwa_b200_pipeline_kernel_times",
    "bwa_b200_chain_params_default", "bwa_b200_alignments_free", "bwa_b200_aligner_create", "bwa_b200_aligner_set_contigs",
    "bwa_b200_aligner_destroy", "bwa_b200_align_host", "bwa_b200_align_host_view", "bwa_b200_align_seeds_host", "bwa_b200_align_device",
    "bwa_b200_align_device_view", "bwa_b200_aligner_skipped_reads", "bwa_b200_aligner_stream", "bwa_b200_aligner_launches", "bwa_b200_aligner_profile",
    "bwa_b200_aligner_kernel_times",
    "bwa_b200_cigar_create", "bwa_b200_cigar_destroy", "bwa_b200_cigar_band", "bwa_b200_global_host", "bwa_b200_global_host_view", "bwa_b200_cigars_free",
    "bwa_b200_global_device", "bwa_b200_global_device_view", "bwa_b200_cigar_stream", "bwa_b200_cigar_launches",
    "bwa_b200_cigar_last_cells", "bwa_b200_cigar_profile", "bwa_b200_cigar_kernel_times", "bwa_b200_reg2aln_host",
    "bwa_b200_region_opt_default", "bwa_b200_finish_regions_host",
    "bwa_b200_sw_create", "bwa_b200_sw_destroy", "bwa_b200_sw_align2_host", "bwa_b200_sw_launches", "bwa_b200_sw_last_kernel_ms",
    "bwa_b200_packed2_words", "bwa_b200_pack2_codes", "bwa_b200_pack2_ascii", "bwa_b200_align_host_compact",
    "bwa_b200_multi_create", "bwa_b200_multi_set_contigs", "bwa_b200_multi_align_compact", "bwa_b200_multi_submit_compact", "bwa_b200_multi_wait", "bwa_b200_multi_n_workers",
    "bwa_b200_multi_worker_chunks", "bwa_b200_multi_launches", "bwa_b200_multi_destroy",
]

ALN_IN_DTYPE = np.dtype([("read", "<u4"), ("qb", "<i4"), ("qe", "<i4"), ("rb", "<i8"), ("re", "<i8"), ("truesc", "<i4"), ("w", "<i4")], align=True)
ALN_OUT_DTYPE = np.dtype([("pos", "<i8"), ("rid", "<i4"), ("is_rev", "<i4"), ("score", "<i4"), ("nm", "<i4"), ("n_cigar", "<i4"), ("band", "<i4"),
                          ("n_waves", "<i4"), ("cigar_off", "<u8")], align=True)


class Cigars(C.Structure):
    _fields_ = [("n_jobs", C.c_uint64), ("n_ops", C.c_uint64), ("score", C.POINTER(C.c_int32)), ("nm", C.POINTER(C.c_int32)),
                ("n_cigar", C.POINTER(C.c_uint32)), ("cigar_off", C.POINTER(C.c_uint64)), ("cigar", C.POINTER(C.c_uint32))]


class ChainParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("a", "b", "o_del", "e_del", "o_ins", "e_ins", "w", "min_seed_len", "max_occ",
                                         "max_chain_gap", "min_chain_weight", "max_chain_extend")] + \
               [("mask_level", C.c_float), ("drop_ratio", C.c_float)]


CHAIN_DTYPE = np.dtype([("pos", "<i8"), ("rid", "<i4"), ("n", "<i4"), ("w", "<i4"), ("kept", "<i4"), ("first", "<i4"),
                        ("is_alt", "<i4"), ("frac_rep", "<f4"), ("seed_off", "<i4")], align=True)
CHAIN_SEED_DTYPE = np.dtype([("rbeg", "<i8"), ("qbeg", "<i4"), ("len", "<i4"), ("score", "<i4"), ("pad", "<i4")], align=True)
REGION_DTYPE = np.dtype([("rb", "<i8"), ("re", "<i8"), ("rb_est", "<i8"), ("re_est", "<i8"), ("target_seed_begin", "<i8"),
                         ("qb", "<i4"), ("qe", "<i4"), ("score", "<i4"), ("truesc", "<i4"),
                         ("qb_est", "<i4"), ("qe_est", "<i4"), ("rid", "<i4"), ("align_sides", "<i4"), ("where_is_long", "<i4"),
                         ("query_seed_begin", "<i4"), ("seedlen0", "<i4"), ("seedcov", "<i4"), ("w", "<i4"), ("frac_rep", "<f4"),
                         ("left_tlen", "<i4"), ("right_tlen", "<i4"), ("job_short", "<i4"), ("job_long", "<i4")], align=True)
REGION_COMPACT_DTYPE = np.dtype([("rb", "<i8"), ("rlen", "<i4"), ("qb", "<u2"), ("qe", "<u2"), ("score", "<i4"), ("truesc", "<i4"),
                                 ("seedcov", "<i4"), ("rid", "<i4"), ("w", "<u2"), ("seedlen0", "<u2"), ("frac_rep", "<f4")], align=True)
assert REGION_COMPACT_DTYPE.itemsize == 40
JOB_DTYPE = np.dtype([("qoff", "<u4"), ("qlen", "<u4"), ("toff", "<u4"), ("tlen", "<u4"), ("h0", "<u4")], align=True)
assert CHAIN_DTYPE.itemsize == 40 and CHAIN_SEED_DTYPE.itemsize == 24 and REGION_DTYPE.itemsize == 112 and JOB_DTYPE.itemsize == 20


class Alignments(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_reads", "n_regions", "n_chains", "n_chain_seeds", "n_jobs_short", "n_jobs_long",
                                          "q_words", "t_words")] + \
               [(k, C.c_void_p) for k in ("n_regions_per_read", "region_off", "regions", "n_chains_per_read", "chain_off", "chains",
                                          "chain_seed_off", "chain_seeds", "jobs", "qpacked", "tpacked", "job_res")]


class AlignView(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("n_reads", "n_regions", "n_jobs_short", "n_jobs_long", "n_seeds", "cells")] + \
               [(k, C.c_void_p) for k in ("n_regions_per_read", "region_off", "regions")] + [("closed_form_jobs", C.c_uint64)]


_lib = None
vp = C.c_void_p


def lib():
    """Load the CUDA library.  Fails loudly if it is missing: there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is not built; run __graft_entry__.build() (needs nvcc). "
                               "There is no CPU fallback for the hot paths.")
        L = C.CDLL(LIB_PATH)
        L.bwa_b200_last_error.restype = C.c_char_p
        L.bwa_b200_host_alloc.restype = vp
        L.bwa_b200_host_alloc.argtypes = [C.c_size_t]
        L.bwa_b200_host_free.argtypes = [vp]
        L.bwa_b200_index_load.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(vp)]
        L.bwa_b200_index_from_host.argtypes = [C.c_uint64, vp, vp, C.c_uint64, vp, vp, C.c_uint64, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
        L.bwa_b200_index_clone_to.argtypes = [vp, C.c_int, C.POINTER(vp)]
        L.bwa_b200_index_info.argtypes = [vp, C.POINTER(IndexInfo)]
        L.bwa_b200_index_free.argtypes = [vp]
        L.bwa_b200_index_set_kmer_table.argtypes = [vp, C.c_int]
        L.bwa_b200_build_index.argtypes = [vp, C.c_uint64, C.c_int, C.c_char_p, C.c_int, C.c_int]
        L.bwa_b200_packed_words.argtypes = [vp, C.c_uint64]
        L.bwa_b200_packed_words.restype = C.c_size_t
        L.bwa_b200_pack_ascii.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_int]
        L.bwa_b200_pack_codes.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_int]
        L.bwa_b200_seeder_create.argtypes = [vp, C.c_uint64, C.c_uint64, C.POINTER(vp)]
        L.bwa_b200_seeder_destroy.argtypes = [vp]
        L.bwa_b200_seed_host.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(SeedParams), C.POINTER(Seeds)]
        L.bwa_b200_seeds_free.argtypes = [C.POINTER(Seeds)]
        L.bwa_b200_seed_device.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(SeedParams)]
        L.bwa_b200_seed_device_result.argtypes = [vp, C.POINTER(Seeds)]
        L.bwa_b200_seeder_stream.argtypes = [vp]
        L.bwa_b200_seeder_stream.restype = vp
        L.bwa_b200_seed_device_smems.argtypes = [vp, C.c_uint64, vp, vp, vp, vp, vp, C.c_uint64, C.POINTER(C.c_uint64)]
        L.bwa_b200_seeder_launches.argtypes = [vp]
        L.bwa_b200_seeder_launches.restype = C.c_uint64
        L.bwa_b200_seeder_request_counts.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]
        L.bwa_b200_measure_random_sector_gbs.argtypes = [C.c_int, C.c_uint64, C.c_int, C.c_int]
        L.bwa_b200_measure_random_sector_gbs.restype = C.c_double
        L.bwa_b200_measure_int_alu.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
        L.bwa_b200_int_alu_op_name.argtypes = [C.c_int]
        L.bwa_b200_int_alu_op_name.restype = C.c_char_p
        L.bwa_b200_packed2_words.argtypes = [vp, C.c_uint64, C.c_uint32]
        L.bwa_b200_packed2_words.restype = C.c_size_t
        L.bwa_b200_pack2_codes.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
        L.bwa_b200_pack2_ascii.argtypes = [vp, vp, C.c_uint64, vp, vp, vp, C.c_uint64, C.POINTER(C.c_uint64), C.c_int]
        L.bwa_b200_align_host_compact.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_uint64, C.POINTER(SeedParams), C.POINTER(ChainParams),
                                                  C.POINTER(ExtParams), C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp)]
        L.bwa_b200_multi_create.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_uint64, C.c_uint32, C.POINTER(vp)]
        L.bwa_b200_multi_set_contigs.argtypes = [vp, C.c_int32, vp, vp, vp]
        L.bwa_b200_multi_align_compact.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_uint64, C.POINTER(SeedParams), C.POINTER(ChainParams),
                                                   C.POINTER(ExtParams), C.POINTER(MultiResult)]
        L.bwa_b200_multi_submit_compact.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_uint64, C.POINTER(SeedParams), C.POINTER(ChainParams),
                                                    C.POINTER(ExtParams), C.POINTER(C.c_int)]
        L.bwa_b200_multi_wait.argtypes = [vp, C.c_int, C.POINTER(MultiResult)]
        L.bwa_b200_multi_n_workers.argtypes = [vp]
        L.bwa_b200_multi_worker_chunks.argtypes = [vp, C.c_int]
        L.bwa_b200_multi_worker_chunks.restype = C.c_uint64
        L.bwa_b200_multi_launches.argtypes = [vp]
        L.bwa_b200_multi_launches.restype = C.c_uint64
        L.bwa_b200_multi_destroy.argtypes = [vp]
        L.bwa_b200_multi_destroy.restype = None
        L.bwa_b200_ext_params_default.argtypes = [C.POINTER(ExtParams)]
        L.bwa_b200_fill_scmat.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int8)]
        L.bwa_b200_extender_create.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.POINTER(vp)]
        L.bwa_b200_extender_destroy.argtypes = [vp]
        L.bwa_b200_extend_async.argtypes = [vp, C.POINTER(ExtParams), C.c_uint64, vp, C.c_uint64, vp, vp, vp, C.c_uint64,
                                            vp, vp, vp, vp, vp, vp, vp]
        L.bwa_b200_extend_query.argtypes = [vp]
        L.bwa_b200_extend_wait.argtypes = [vp]
        L.bwa_b200_extend_device.argtypes = [vp, C.POINTER(ExtParams), C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp]
        L.bwa_b200_pack_device.argtypes = [vp, vp, C.c_uint64, vp]
        L.bwa_b200_extender_stream.argtypes = [vp]
        L.bwa_b200_extender_stream.restype = vp
        L.bwa_b200_extender_launches.argtypes = [vp]
        L.bwa_b200_extender_launches.restype = C.c_uint64
        L.bwa_b200_extender_last_cells.argtypes = [vp]
        L.bwa_b200_extender_last_cells.restype = C.c_uint64
        L.bwa_b200_extender_last_closed_form.argtypes = [vp]
        L.bwa_b200_extender_last_closed_form.restype = C.c_uint64
        L.bwa_b200_extender_set_closed_form.argtypes = [vp, C.c_int]
        L.bwa_b200_index_attach_ref.argtypes = [vp, vp, C.c_uint64]
        L.bwa_b200_pipeline_create.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(vp)]
        L.bwa_b200_pipeline_destroy.argtypes = [vp]
        L.bwa_b200_seed_extend_host.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(SeedParams), C.POINTER(ExtParams), vp]
        L.bwa_b200_seed_extend_device.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_uint32, C.POINTER(SeedParams), C.POINTER(ExtParams), vp]
        L.bwa_b200_pipeline_sync.argtypes = [vp]
        L.bwa_b200_pipeline_stream.argtypes = [vp]
        L.bwa_b200_pipeline_stream.restype = vp
        L.bwa_b200_pipeline_launches.argtypes = [vp]
        L.bwa_b200_pipeline_launches.restype = C.c_uint64
        L.bwa_b200_pipeline_totals.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.bwa_b200_pipeline_profile.argtypes = [vp, C.c_int]
        L.bwa_b200_pipeline_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
        L.bwa_b200_chain_params_default.argtypes = [C.POINTER(ChainParams)]
        L.bwa_b200_alignments_free.argtypes = [C.POINTER(Alignments)]
        L.bwa_b200_aligner_create.argtypes = [vp, C.c_uint64, C.c_uint64, C.POINTER(vp)]
        L.bwa_b200_aligner_set_contigs.argtypes = [vp, C.c_int32, vp, vp, vp]
        L.bwa_b200_aligner_destroy.argtypes = [vp]
        L.bwa_b200_align_host.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(SeedParams), C.POINTER(ChainParams), C.POINTER(ExtParams),
                                          C.c_int, C.POINTER(Alignments)]
        L.bwa_b200_align_host_view.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(SeedParams), C.POINTER(ChainParams), C.POINTER(ExtParams),
                                               C.POINTER(C.c_uint64), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.bwa_b200_align_seeds_host.argtypes = [vp, vp, vp, vp, C.c_uint64, C.POINTER(Seeds), C.c_int, C.POINTER(ChainParams),
                                                C.POINTER(ExtParams), C.c_int, C.POINTER(Alignments)]
        L.bwa_b200_align_device.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_uint32, C.POINTER(SeedParams), C.POINTER(ChainParams),
                                            C.POINTER(ExtParams)]
        L.bwa_b200_align_device_view.argtypes = [vp, C.POINTER(AlignView)]
        L.bwa_b200_aligner_skipped_reads.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(vp)]
        L.bwa_b200_aligner_stream.argtypes = [vp]
        L.bwa_b200_aligner_stream.restype = vp
        L.bwa_b200_aligner_launches.argtypes = [vp]
        L.bwa_b200_aligner_launches.restype = C.c_uint64
        L.bwa_b200_aligner_profile.argtypes = [vp, C.c_int]
        L.bwa_b200_aligner_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
        L.bwa_b200_sw_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.bwa_b200_sw_destroy.argtypes = [vp]
        L.bwa_b200_sw_align2_host.argtypes = [vp, C.POINTER(ExtParams), C.c_uint64, vp, C.c_uint64, vp, vp, vp, C.c_uint64, vp, vp, vp, vp]
        L.bwa_b200_sw_launches.argtypes = [vp]
        L.bwa_b200_sw_launches.restype = C.c_uint64
        L.bwa_b200_sw_last_kernel_ms.argtypes = [vp]
        L.bwa_b200_sw_last_kernel_ms.restype = C.c_float
        L.bwa_b200_cigar_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.bwa_b200_cigar_destroy.argtypes = [vp]
        L.bwa_b200_cigar_band.argtypes = [C.POINTER(ExtParams), C.c_int, C.c_int, C.c_int64]
        L.bwa_b200_global_host.argtypes = [vp, C.POINTER(ExtParams), C.c_uint64, vp, C.c_uint64, vp, vp, vp, C.c_uint64, vp, vp, vp, C.POINTER(Cigars)]
        L.bwa_b200_global_host_view.argtypes = [vp, C.POINTER(ExtParams), C.c_uint64, vp, C.c_uint64, vp, vp, vp, C.c_uint64, vp, vp, vp, C.POINTER(Cigars)]
        L.bwa_b200_cigars_free.argtypes = [C.POINTER(Cigars)]
        L.bwa_b200_global_device.argtypes = [vp, C.POINTER(ExtParams), C.c_uint64, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int]
        L.bwa_b200_global_device_view.argtypes = [vp, C.POINTER(Cigars)]
        L.bwa_b200_cigar_stream.argtypes = [vp]
        L.bwa_b200_cigar_stream.restype = vp
        L.bwa_b200_cigar_launches.argtypes = [vp]
        L.bwa_b200_cigar_launches.restype = C.c_uint64
        L.bwa_b200_cigar_last_cells.argtypes = [vp]
        L.bwa_b200_cigar_last_cells.restype = C.c_uint64
        L.bwa_b200_cigar_profile.argtypes = [vp, C.c_int]
        L.bwa_b200_cigar_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
        L.bwa_b200_reg2aln_host.argtypes = [vp, vp, C.c_int32, vp, vp, vp, vp, C.c_uint64, vp, C.c_uint64, C.POINTER(ExtParams), C.c_int32, vp,
                                            C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]
        L.bwa_b200_region_opt_default.argtypes = [C.POINTER(RegionOpt)]
        L.bwa_b200_region_opt_default.restype = None
        L.bwa_b200_finish_regions_host.argtypes = [vp, C.c_int32, vp, vp, vp, vp, C.c_uint64, vp, vp, vp, vp, C.c_int64, C.POINTER(RegionOpt)]
        _lib = L
    return _lib


def check(rc: int):
    if rc != 0:
        raise B200Error(rc, lib().bwa_b200_last_error().decode())


def _p(a):
    """raw pointer of a numpy array (kept alive by the caller) or an int device pointer"""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    return a.ctypes.data


def build_index(fwd_codes: np.ndarray, prefix: str, sa_intv: int = 16, also_stock_layout: bool = False, n_threads: int = 0):
    fwd = np.ascontiguousarray(fwd_codes, dtype=np.uint8)
    check(lib().bwa_b200_build_index(_p(fwd), fwd.size, sa_intv, prefix.encode(), int(also_stock_layout), n_threads))


def pack_codes(reads_flat: np.ndarray, base_off: np.ndarray, n_threads: int = 0):
    """codes (0..4) -> (packed u32 words, word_off u64[n+1], read_len u32[n])"""
    reads_flat = np.ascontiguousarray(reads_flat, dtype=np.uint8)
    base_off = np.ascontiguousarray(base_off, dtype=np.uint64)
    n = base_off.size - 1
    lens = (base_off[1:] - base_off[:-1]).astype(np.uint32)
    n_words = int(((lens.astype(np.uint64) + 7) // 8).sum())
    packed = np.zeros(max(n_words, 1), np.uint32)
    woff = np.zeros(n + 1, np.uint64)
    rl = np.zeros(max(n, 1), np.uint32)
    check(lib().bwa_b200_pack_codes(_p(reads_flat), _p(base_off), n, _p(packed), _p(woff), _p(rl), n_threads))
    return packed[:n_words], woff, rl[:n]


def pack_ascii(ascii_bytes: np.ndarray, base_off: np.ndarray, n_threads: int = 0):
    ascii_bytes = np.ascontiguousarray(ascii_bytes, dtype=np.uint8)
    base_off = np.ascontiguousarray(base_off, dtype=np.uint64)
    n = base_off.size - 1
    lens = (base_off[1:] - base_off[:-1]).astype(np.uint32)
    n_words = int(((lens.astype(np.uint64) + 7) // 8).sum())
    packed = np.zeros(max(n_words, 1), np.uint32)
    woff = np.zeros(n + 1, np.uint64)
    rl = np.zeros(max(n, 1), np.uint32)
    check(lib().bwa_b200_pack_ascii(_p(ascii_bytes), _p(base_off), n, _p(packed), _p(woff), _p(rl), n_threads))
    return packed[:n_words], woff, rl[:n]


class Index:
    """Device-resident FMD index (bwt_restore_bwt_gpu + bwt_restore_sa_gpu + gpu_cpy_wrapper)."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def load(cls, bwt_path: str, sa_path: str | None, device: int = 0):
        h = vp()
        check(lib().bwa_b200_index_load(bwt_path.encode(), sa_path.encode() if sa_path else None, device, C.byref(h)))
        return cls(h)

    def clone_to(self, device: int):
        h = vp()
        check(lib().bwa_b200_index_clone_to(self.h, device, C.byref(h)))
        return Index(h)

    def info(self) -> IndexInfo:
        info = IndexInfo()
        check(lib().bwa_b200_index_info(self.h, C.byref(info)))
        return info

    def set_kmer_table(self, K: int):
        """rebuild the k-mer interval table with another K (0 drops it)"""
        check(lib().bwa_b200_index_set_kmer_table(self.h, int(K)))

    def attach_ref(self, fwd_codes: np.ndarray):
        fwd = np.ascontiguousarray(fwd_codes, dtype=np.uint8)
        check(lib().bwa_b200_index_attach_ref(self.h, _p(fwd), fwd.size))

    def free(self):
        if self.h:
            lib().bwa_b200_index_free(self.h)
            self.h = None


class Seeder:
    def __init__(self, index: Index, max_reads: int, max_words: int):
        self.h = vp()
        self.index = index
        check(lib().bwa_b200_seeder_create(index.h, max_reads, max_words, C.byref(self.h)))

    def seed_host(self, packed, word_off, read_len, min_seed_len=19, max_occ=500, params: SeedParams = None):
        """host arrays in, numpy arrays out (copies of the malloc'ed result)."""
        n = read_len.size
        p = params if params is not None else SeedParams(min_seed_len, max_occ)
        out = Seeds()
        check(lib().bwa_b200_seed_host(self.h, _p(packed), _p(word_off), _p(read_len), n, C.byref(p), C.byref(out)))
        tot = int(out.n_seeds)
        res = dict(
            total=tot,
            rbeg=np.ctypeslib.as_array(out.rbeg, shape=(max(tot, 1),))[:tot].copy(),
            qq=np.ctypeslib.as_array(out.qbeg_qend, shape=(max(tot, 1) * 2,))[:2 * tot].copy().reshape(-1, 2),
            score=np.ctypeslib.as_array(out.score, shape=(max(tot, 1),))[:tot].copy(),
            n_seeds=np.ctypeslib.as_array(out.n_seeds_per_read, shape=(max(n, 1),))[:n].copy(),
            seed_off=np.ctypeslib.as_array(out.seed_off, shape=(max(n, 1),))[:n].copy(),
        )
        lib().bwa_b200_seeds_free(C.byref(out))
        return res

    def seed_device(self, d_packed: int, d_word_off: int, d_read_len: int, n_reads: int, min_seed_len=19, max_occ=500, params: SeedParams = None):
        p = params if params is not None else SeedParams(min_seed_len, max_occ)
        check(lib().bwa_b200_seed_device(self.h, d_packed, d_word_off, d_read_len, n_reads, C.byref(p)))

    def device_result(self) -> Seeds:
        out = Seeds()
        check(lib().bwa_b200_seed_device_result(self.h, C.byref(out)))
        return out

    def smems(self, n_reads: int, cap: int):
        n_smems = np.zeros(max(n_reads, 1), np.uint32)
        qb = np.zeros(cap, np.int32)
        qe = np.zeros(cap, np.int32)
        k = np.zeros(cap, np.uint64)
        s = np.zeros(cap, np.uint64)
        tot = C.c_uint64()
        check(lib().bwa_b200_seed_device_smems(self.h, n_reads, _p(n_smems), _p(qb), _p(qe), _p(k), _p(s), cap, C.byref(tot)))
        t = int(tot.value)
        return dict(n_smems=n_smems[:n_reads], qbeg=qb[:t], qend=qe[:t], k=k[:t], s=s[:t])

    def request_counts(self, enable=True):
        """switch the request counting of fwd_kernel / back_kernel on or off; returns the last counted batch's
        dict(fwd_sectors, fwd_table, back_sectors, back_table)"""
        out = (C.c_uint64 * 4)()
        check(lib().bwa_b200_seeder_request_counts(self.h, int(enable), out))
        return dict(fwd_sectors=int(out[0]), fwd_table=int(out[1]), back_sectors=int(out[2]), back_table=int(out[3]))

    @property
    def stream(self) -> int:
        return int(lib().bwa_b200_seeder_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().bwa_b200_seeder_launches(self.h))

    def destroy(self):
        if self.h:
            lib().bwa_b200_seeder_destroy(self.h)
            self.h = None


def ext_params(a=1, b=4, o_del=6, e_del=1, o_ins=6, e_ins=1, w=100, end_bonus=5, zdrop=100, use_band=1, pen_clip=5) -> ExtParams:
    p = ExtParams()
    lib().bwa_b200_fill_scmat(a, b, p.mat)
    p.o_del, p.e_del, p.o_ins, p.e_ins = o_del, e_del, o_ins, e_ins
    p.w, p.end_bonus, p.zdrop, p.use_band, p.pen_clip = w, end_bonus, zdrop, use_band, pen_clip
    return p


class Extender:
    def __init__(self, device: int = 0, max_jobs: int = 1 << 16, max_q: int = 1 << 20, max_t: int = 1 << 21):
        self.h = vp()
        check(lib().bwa_b200_extender_create(device, max_jobs, max_q, max_t, C.byref(self.h)))

    def extend_host(self, jobs: dict, params: ExtParams, want_triple: bool = True):
        """jobs: qseq,tseq (uint8 codes), qoff,toff,qlen,tlen,h0 (uint32).  Returns (res6[n,6], triple[3,n])."""
        n = jobs["qlen"].size
        res = np.zeros((n, 6), np.int32)
        tri = np.zeros((3, n), np.int32) if want_triple else None
        check(lib().bwa_b200_extend_async(self.h, C.byref(params), n, _p(jobs["qseq"]), jobs["qseq"].size, _p(jobs["qoff"]),
                                          _p(jobs["qlen"]), _p(jobs["tseq"]), jobs["tseq"].size, _p(jobs["toff"]),
                                          _p(jobs["tlen"]), _p(jobs["h0"]), _p(res),
                                          _p(tri[0]) if want_triple else None, _p(tri[1]) if want_triple else None,
                                          _p(tri[2]) if want_triple else None))
        check(lib().bwa_b200_extend_wait(self.h))
        return res, tri

    def extend_async(self, params, n, qseq, q_bytes, qoff, qlen, tseq, t_bytes, toff, tlen, h0, res6, sc=None, qe=None, te=None):
        check(lib().bwa_b200_extend_async(self.h, C.byref(params), n, _p(qseq), q_bytes, _p(qoff), _p(qlen), _p(tseq), t_bytes,
                                          _p(toff), _p(tlen), _p(h0), _p(res6), _p(sc), _p(qe), _p(te)))

    def extend_device(self, params, n, d_qp, d_qoff, d_qlen, d_tp, d_toff, d_tlen, d_h0, d_res6):
        check(lib().bwa_b200_extend_device(self.h, C.byref(params), n, d_qp, d_qoff, d_qlen, d_tp, d_toff, d_tlen, d_h0, d_res6))

    def pack_device(self, d_bytes: int, n_bytes: int, d_packed: int):
        check(lib().bwa_b200_pack_device(self.h, d_bytes, n_bytes, d_packed))

    def query(self) -> int:
        return int(lib().bwa_b200_extend_query(self.h))

    def wait(self):
        check(lib().bwa_b200_extend_wait(self.h))

    @property
    def stream(self) -> int:
        return int(lib().bwa_b200_extender_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().bwa_b200_extender_launches(self.h))

    def last_cells(self) -> int:
        return int(lib().bwa_b200_extender_last_cells(self.h))

    def last_closed_form(self) -> int:
        """jobs of the last batch answered in closed form (not in last_cells)"""
        return int(lib().bwa_b200_extender_last_closed_form(self.h))

    def set_closed_form(self, on: bool):
        check(lib().bwa_b200_extender_set_closed_form(self.h, int(on)))

    def destroy(self):
        if self.h:
            lib().bwa_b200_extender_destroy(self.h)
            self.h = None


class Pipeline:
    """Fused seed -> extend pass (bwa_b200_seed_extend_*)."""

    def __init__(self, index: Index, max_reads: int, max_words: int, max_read_len: int):
        self.h = vp()
        self.index = index
        check(lib().bwa_b200_pipeline_create(index.h, max_reads, max_words, max_read_len, C.byref(self.h)))

    def run_host(self, packed, word_off, read_len, seed_params: SeedParams, ext_p: ExtParams, out=None):
        n = read_len.size
        if out is None:
            out = np.zeros(max(n, 1), READ_RESULT_DTYPE)
        check(lib().bwa_b200_seed_extend_host(self.h, _p(packed), _p(word_off), _p(read_len), n, C.byref(seed_params),
                                              C.byref(ext_p), _p(out)))
        return out[:n]

    def run_device(self, d_packed, d_woff, d_len, n, max_read_len, seed_params, ext_p, d_out):
        check(lib().bwa_b200_seed_extend_device(self.h, d_packed, d_woff, d_len, n, max_read_len, C.byref(seed_params),
                                                C.byref(ext_p), d_out))

    def sync(self):
        check(lib().bwa_b200_pipeline_sync(self.h))

    def totals(self):
        out = (C.c_uint64 * 3)()
        check(lib().bwa_b200_pipeline_totals(self.h, out))
        return dict(seeds=int(out[0]), jobs=int(out[1]), cells=int(out[2]))

    def profile(self, on: bool):
        check(lib().bwa_b200_pipeline_profile(self.h, int(on)))

    def kernel_times(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        n = lib().bwa_b200_pipeline_kernel_times(self.h, names, ms, 64)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    @property
    def stream(self) -> int:
        return int(lib().bwa_b200_pipeline_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().bwa_b200_pipeline_launches(self.h))

    def destroy(self):
        if self.h:
            lib().bwa_b200_pipeline_destroy(self.h)
            self.h = None


def chain_params(**kw) -> ChainParams:
    p = ChainParams()
    lib().bwa_b200_chain_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def _take(ptr, n, dtype):
    """copy n records out of a malloc'ed C array"""
    dtype = np.dtype(dtype)
    if not ptr or n == 0:
        return np.zeros(0, dtype)
    return np.frombuffer(C.string_at(ptr, int(n) * dtype.itemsize), dtype=dtype).copy()


def _view(ptr, cnt, dt, copy=True):
    """numpy array over cnt items at a C pointer (a copy, or a view valid as long as the owner keeps the buffer)"""
    dt = np.dtype(dt)
    val = ptr.value if hasattr(ptr, "value") else ptr
    if not val or not cnt:
        return np.zeros(0, dt)
    v = np.frombuffer((C.c_char * (int(cnt) * dt.itemsize)).from_address(val), dtype=dt)
    return v.copy() if copy else v


def pack2_codes(codes: np.ndarray, base_off: np.ndarray, with_lengths: bool = True, n_threads: int = 0):
    """codes (0..3, anything else = N) -> (packed2, read_len or None, n_list): the compact wire layout of bwa_b200_align_host_compact"""
    codes = np.ascontiguousarray(codes, np.uint8); base_off = np.ascontiguousarray(base_off, np.uint64)
    n = base_off.size - 1
    lens = (base_off[1:] - base_off[:-1]).astype(np.uint32)
    words = int(((lens.astype(np.uint64) + np.uint64(15)) // np.uint64(16)).sum())
    packed2 = np.zeros(max(words, 1), np.uint32)
    rl = np.zeros(max(n, 1), np.uint32)
    cap = int((codes > 3).sum()) + 1
    nl = np.zeros(cap, np.uint64)
    nn = C.c_uint64(0)
    check(lib().bwa_b200_pack2_codes(_p(codes), _p(base_off), n, _p(packed2), _p(rl), _p(nl), cap, C.byref(nn), n_threads))
    return packed2[:words], (rl[:n] if with_lengths else None), nl[:nn.value].copy()


def unpack_compact(res: dict) -> dict:
    """the compact records with the names of REGION_DTYPE (re = rb + rlen), and exclusive region offsets per read"""
    r = res["regions"]
    off = np.zeros(res["n_regions"].size, np.uint64)
    if off.size:
        off[1:] = np.cumsum(res["n_regions"][:-1], dtype=np.uint64)
    return dict(n_regions=res["n_regions"], region_off=off, rb=r["rb"].astype(np.int64), re=r["rb"] + r["rlen"], qb=r["qb"].astype(np.int32), qe=r["qe"].astype(np.int32),
                score=r["score"], truesc=r["truesc"], seedcov=r["seedcov"], rid=r["rid"], w=r["w"].astype(np.int32), seedlen0=r["seedlen0"].astype(np.int32),
                frac_rep=r["frac_rep"])


class MultiResult(C.Structure):
    _fields_ = [("n_reads", C.c_uint64), ("n_regions", C.c_uint64), ("n_chunks", C.c_uint64), ("chunk_reads", C.c_uint64),
                ("n_regions_per_read", vp), ("chunk_region_off", vp), ("regions", vp)]


class MultiAligner:
    """every GPU of the box behind one call (bwa_b200_multi_*): chunks of reads dealt to worker threads, the index replicated by peer copies"""

    def __init__(self, index, devices, workers_per_device: int, chunk_reads: int, max_read_len: int):
        self.h = vp()
        self.index = index
        dv = (C.c_int * len(devices))(*devices)
        check(lib().bwa_b200_multi_create(index.h, dv, len(devices), workers_per_device, chunk_reads, max_read_len, C.byref(self.h)))

    def set_contigs(self, offset, length, is_alt=None):
        off = np.ascontiguousarray(offset, dtype=np.int64); ln = np.ascontiguousarray(length, dtype=np.int32)
        alt = np.ascontiguousarray(is_alt, dtype=np.int32) if is_alt is not None else None
        check(lib().bwa_b200_multi_set_contigs(self.h, ln.size, _p(off), _p(ln), _p(alt)))

    def submit_compact(self, packed2_ptr, len_ptr, uniform_len, n, nlist_ptr, n_n, seed_p, chain_p, ext_p) -> int:
        """start a batch and return its ticket; at most two in flight (bwa_b200_multi_submit_compact)"""
        t = C.c_int(-1)
        check(lib().bwa_b200_multi_submit_compact(self.h, packed2_ptr, len_ptr or None, uniform_len, n, nlist_ptr or None, n_n, C.byref(seed_p), C.byref(chain_p),
                                                  C.byref(ext_p), C.byref(t)))
        return int(t.value)

    def wait(self, ticket: int, copy=True, gather=True):
        """block until the batch behind the ticket is done; same dict as align_compact"""
        out = MultiResult()
        check(lib().bwa_b200_multi_wait(self.h, int(ticket), C.byref(out)))
        return self._result(out, int(out.n_reads), copy, gather)

    def align_compact(self, packed2_ptr, len_ptr, uniform_len, n, nlist_ptr, n_n, seed_p, chain_p, ext_p, copy=True, gather=True):
        """regions of the whole batch; gather=True puts the chunks' records in read order (a host copy), False returns them as the chunks landed
        together with chunk_region_off"""
        out = MultiResult()
        check(lib().bwa_b200_multi_align_compact(self.h, packed2_ptr, len_ptr or None, uniform_len, n, nlist_ptr or None, n_n, C.byref(seed_p), C.byref(chain_p),
                                                 C.byref(ext_p), C.byref(out)))
        return self._result(out, n, copy, gather)

    @staticmethod
    def _result(out, n, copy, gather):
        nregs = _view(out.n_regions_per_read, n, np.uint32, copy)
        coff = _view(out.chunk_region_off, out.n_chunks, np.uint64, True)
        regs = _view(out.regions, out.n_regions, REGION_COMPACT_DTYPE, copy and not gather)
        res = dict(n_regions=nregs, chunk_region_off=coff, chunk_reads=int(out.chunk_reads), n_chunks=int(out.n_chunks), regions=regs)
        if gather and out.n_chunks:
            cr = int(out.chunk_reads)
            per_chunk = np.add.reduceat(nregs.astype(np.uint64), np.arange(0, n, cr))
            res["regions"] = np.concatenate([regs[int(coff[k]):int(coff[k]) + int(per_chunk[k])] for k in range(int(out.n_chunks))]) if out.n_regions else regs.copy()
        return res

    @property
    def n_workers(self) -> int:
        return int(lib().bwa_b200_multi_n_workers(self.h))

    def worker_chunks(self):
        return [int(lib().bwa_b200_multi_worker_chunks(self.h, i)) for i in range(self.n_workers)]

    @property
    def launches(self) -> int:
        return int(lib().bwa_b200_multi_launches(self.h))

    def destroy(self):
        if self.h:
            lib().bwa_b200_multi_destroy(self.h)
            self.h = vp()


class Aligner:
    """seeds -> chains -> extension jobs -> extension -> alignment regions on the device (bwa_b200_align_*)."""

    def __init__(self, index: Index, max_reads: int, max_words: int):
        self.h = vp()
        self.index = index
        check(lib().bwa_b200_aligner_create(index.h, max_reads, max_words, C.byref(self.h)))

    def set_contigs(self, offset, length, is_alt=None):
        off = np.ascontiguousarray(offset, dtype=np.int64); ln = np.ascontiguousarray(length, dtype=np.int32)
        alt = np.ascontiguousarray(is_alt, dtype=np.int32) if is_alt is not None else None
        check(lib().bwa_b200_aligner_set_contigs(self.h, ln.size, _p(off), _p(ln), _p(alt)))

    @staticmethod
    def _unpack(out: Alignments, detail: bool):
        n = int(out.n_reads)
        r = dict(n_regions=_take(out.n_regions_per_read, n, np.uint32), region_off=_take(out.region_off, n, np.uint64),
                 regions=_take(out.regions, out.n_regions, REGION_DTYPE))
        if detail:
            nj = int(out.n_jobs_short + out.n_jobs_long)
            r.update(n_chains=_take(out.n_chains_per_read, n, np.uint32), chain_off=_take(out.chain_off, n, np.uint64),
                     chain_seed_off=_take(out.chain_seed_off, n, np.uint64), chains=_take(out.chains, out.n_chains, CHAIN_DTYPE),
                     chain_seeds=_take(out.chain_seeds, out.n_chain_seeds, CHAIN_SEED_DTYPE), jobs=_take(out.jobs, nj, JOB_DTYPE),
                     qpacked=_take(out.qpacked, out.q_words, np.uint32), tpacked=_take(out.tpacked, out.t_words, np.uint32),
                     job_res=_take(out.job_res, nj * 6, np.int32).reshape(-1, 6), n_jobs_short=int(out.n_jobs_short),
                     n_jobs_long=int(out.n_jobs_long))
        lib().bwa_b200_alignments_free(C.byref(out))
        return r

    def align_host(self, packed, word_off, read_len, seed_p: SeedParams, chain_p: ChainParams, ext_p: ExtParams, detail=False):
        out = Alignments()
        check(lib().bwa_b200_align_host(self.h, _p(packed), _p(word_off), _p(read_len), read_len.size, C.byref(seed_p), C.byref(chain_p),
                                        C.byref(ext_p), int(detail), C.byref(out)))
        return self._unpack(out, detail)

    def align_host_view(self, packed_ptr, woff_ptr, len_ptr, n, seed_p, chain_p, ext_p, copy=True):
        """regions only, through the aligner's pinned result buffers (bwa_b200_align_host_view).  copy=False returns
        numpy views of those buffers, valid until the aligner's next call."""
        nr, a, b, c = C.c_uint64(0), vp(), vp(), vp()
        check(lib().bwa_b200_align_host_view(self.h, packed_ptr, woff_ptr, len_ptr, n, C.byref(seed_p), C.byref(chain_p), C.byref(ext_p),
                                             C.byref(nr), C.byref(a), C.byref(b), C.byref(c)))

        def arr(ptr, cnt, dt):
            dt = np.dtype(dt)
            if not ptr.value or cnt == 0:
                return np.zeros(0, dt)
            v = np.frombuffer((C.c_char * (int(cnt) * dt.itemsize)).from_address(ptr.value), dtype=dt)
            return v.copy() if copy else v
        return dict(n_regions=arr(a, n, np.uint32), region_off=arr(b, n, np.uint64), regions=arr(c, nr.value, REGION_DTYPE))

    def align_host_compact(self, packed2_ptr, len_ptr, uniform_len, n, nlist_ptr, n_n, seed_p, chain_p, ext_p, copy=True):
        """the compact boundary (bwa_b200_align_host_compact): 2-bit reads in (len_ptr None / 0: every read has uniform_len bases), 40-byte region
        records out, through the aligner's pinned buffers"""
        nr, a, c = C.c_uint64(0), vp(), vp()
        check(lib().bwa_b200_align_host_compact(self.h, packed2_ptr, len_ptr or None, uniform_len, n, nlist_ptr or None, n_n, C.byref(seed_p), C.byref(chain_p),
                                                C.byref(ext_p), C.byref(nr), C.byref(a), C.byref(c)))
        return dict(n_regions=_view(a, n, np.uint32, copy), regions=_view(c, nr.value, REGION_COMPACT_DTYPE, copy))

    def align_seeds_host(self, packed, word_off, read_len, rbeg, qq, score, n_seeds, seed_off, layout_all, chain_p, ext_p, detail=False):
        rbeg = np.ascontiguousarray(rbeg, np.uint64); qq = np.ascontiguousarray(qq, np.int32); score = np.ascontiguousarray(score, np.uint32)
        n_seeds = np.ascontiguousarray(n_seeds, np.uint32); seed_off = np.ascontiguousarray(seed_off, np.uint64)
        sd = Seeds(read_len.size, rbeg.size, C.cast(_p(rbeg), C.POINTER(C.c_uint64)), C.cast(_p(qq), C.POINTER(C.c_int32)),
                   C.cast(_p(score), C.POINTER(C.c_uint32)), C.cast(_p(n_seeds), C.POINTER(C.c_uint32)), C.cast(_p(seed_off), C.POINTER(C.c_uint64)))
        out = Alignments()
        check(lib().bwa_b200_align_seeds_host(self.h, _p(packed), _p(word_off), _p(read_len), read_len.size, C.byref(sd), int(layout_all),
                                              C.byref(chain_p), C.byref(ext_p), int(detail), C.byref(out)))
        return self._unpack(out, detail)

    def align_device(self, d_packed, d_woff, d_len, n, max_read_len, seed_p, chain_p, ext_p):
        check(lib().bwa_b200_align_device(self.h, d_packed, d_woff, d_len, n, max_read_len, C.byref(seed_p), C.byref(chain_p), C.byref(ext_p)))

    def skipped_reads(self) -> np.ndarray:
        """always empty: reads long enough for mem_flt_chained_seeds are aligned on the device (kept for ABI compatibility)"""
        n, p = C.c_uint64(0), vp()
        check(lib().bwa_b200_aligner_skipped_reads(self.h, C.byref(n), C.byref(p)))
        if not n.value:
            return np.zeros(0, np.uint32)
        return np.frombuffer((C.c_char * (int(n.value) * 4)).from_address(p.value), dtype=np.uint32).copy()

    def view(self) -> AlignView:
        v = AlignView()
        check(lib().bwa_b200_align_device_view(self.h, C.byref(v)))
        return v

    def profile(self, on: bool):
        check(lib().bwa_b200_aligner_profile(self.h, int(on)))

    def kernel_times(self):
        names = (C.c_char_p * 96)()
        ms = (C.c_float * 96)()
        n = lib().bwa_b200_aligner_kernel_times(self.h, names, ms, 96)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    @property
    def stream(self) -> int:
        return int(lib().bwa_b200_aligner_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().bwa_b200_aligner_launches(self.h))

    def destroy(self):
        if self.h:
            lib().bwa_b200_aligner_destroy(self.h)
            self.h = None


class Cigar:
    """banded global alignment with backtrack on the device: CIGAR, score and NM per job (bwa_b200_global_*)."""

    def __init__(self, device: int = 0):
        self.h = vp()
        check(lib().bwa_b200_cigar_create(device, C.byref(self.h)))

    @staticmethod
    def band(ext_p: ExtParams, w_, l_query, rlen) -> int:
        return int(lib().bwa_b200_cigar_band(C.byref(ext_p), int(w_), int(l_query), int(rlen)))

    def global_host(self, jobs: dict, ext_p: ExtParams):
        """jobs: qseq tseq (uint8 codes) qoff toff qlen tlen w (uint32).  Returns dict(score, nm, n_cigar, cigar_off, cigar)."""
        n = jobs["qlen"].size
        arr = {k: np.ascontiguousarray(jobs[k], np.uint8 if k in ("qseq", "tseq") else np.uint32) for k in ("qseq", "tseq", "qoff", "toff", "qlen", "tlen", "w")}
        out = Cigars()
        check(lib().bwa_b200_global_host(self.h, C.byref(ext_p), n, _p(arr["qseq"]), arr["qseq"].size, _p(arr["qoff"]), _p(arr["qlen"]),
                                         _p(arr["tseq"]), arr["tseq"].size, _p(arr["toff"]), _p(arr["tlen"]), _p(arr["w"]), C.byref(out)))
        res = dict(score=_take(out.score, n, np.int32), nm=_take(out.nm, n, np.int32), n_cigar=_take(out.n_cigar, n, np.uint32),
                   cigar_off=_take(out.cigar_off, n, np.uint64), cigar=_take(out.cigar, out.n_ops, np.uint32))
        lib().bwa_b200_cigars_free(C.byref(out))
        return res

    def global_host_view(self, n, qseq_ptr, q_bytes, qoff_ptr, qlen_ptr, tseq_ptr, t_bytes, toff_ptr, tlen_ptr, w_ptr, ext_p: ExtParams, copy=False):
        """bwa_b200_global_host_view: host pointers in (pinned memory makes the copies asynchronous), results as views of the handle's
        pinned buffers (valid until its next call) or copies"""
        out = Cigars()
        check(lib().bwa_b200_global_host_view(self.h, C.byref(ext_p), n, qseq_ptr, q_bytes, qoff_ptr, qlen_ptr, tseq_ptr, t_bytes, toff_ptr, tlen_ptr, w_ptr, C.byref(out)))

        def arr(ptr, cnt, dt):
            return _view(C.cast(ptr, vp), cnt, dt, copy)
        return dict(score=arr(out.score, n, np.int32), nm=arr(out.nm, n, np.int32), n_cigar=arr(out.n_cigar, n, np.uint32),
                    cigar_off=arr(out.cigar_off, n, np.uint64), cigar=arr(out.cigar, out.n_ops, np.uint32))

    def reg2aln_host(self, index, ctg_off, packed, word_off, read_len, alns, ext_p: ExtParams, match_score: int):
        """mem_reg2aln over a batch (bwa_b200_reg2aln_host).  alns: ALN_IN_DTYPE records.  Returns (ALN_OUT_DTYPE records, flat cigar)."""
        alns = np.ascontiguousarray(alns, dtype=ALN_IN_DTYPE)
        ctg_off = np.ascontiguousarray(ctg_off, dtype=np.int64)
        out = np.zeros(alns.size, ALN_OUT_DTYPE)
        cig = C.POINTER(C.c_uint32)()
        n_ops = C.c_uint64(0)
        check(lib().bwa_b200_reg2aln_host(self.h, index.h, ctg_off.size, _p(ctg_off), _p(packed), _p(word_off), _p(read_len), read_len.size, _p(alns),
                                          alns.size, C.byref(ext_p), int(match_score), _p(out), C.byref(cig), C.byref(n_ops)))
        flat = np.ctypeslib.as_array(cig, shape=(max(int(n_ops.value), 1),))[:int(n_ops.value)].copy()
        C.CDLL(None).free(cig)
        return out, flat

    def global_device(self, ext_p, n, d_qseq, d_qoff, d_qlen, d_tseq, d_toff, d_tlen, h_qlen, h_tlen, h_w, aligned8=False):
        check(lib().bwa_b200_global_device(self.h, C.byref(ext_p), n, d_qseq, d_qoff, d_qlen, d_tseq, d_toff, d_tlen, _p(h_qlen), _p(h_tlen), _p(h_w),
                                           int(aligned8)))

    def view(self) -> Cigars:
        v = Cigars()
        check(lib().bwa_b200_global_device_view(self.h, C.byref(v)))
        return v

    def profile(self, on: bool):
        check(lib().bwa_b200_cigar_profile(self.h, int(on)))

    def kernel_times(self):
        names = (C.c_char_p * 16)()
        ms = (C.c_float * 16)()
        n = lib().bwa_b200_cigar_kernel_times(self.h, names, ms, 16)
        return [(names[i].decode(), float(ms[i])) for i in range(n)]

    @property
    def stream(self) -> int:
        return int(lib().bwa_b200_cigar_stream(self.h) or 0)

    @property
    def launches(self) -> int:
        return int(lib().bwa_b200_cigar_launches(self.h))

    @property
    def last_cells(self) -> int:
        return int(lib().bwa_b200_cigar_last_cells(self.h))

    def destroy(self):
        if self.h:
            lib().bwa_b200_cigar_destroy(self.h)
            self.h = None


def region_opt(**kw) -> RegionOpt:
    o = RegionOpt()
    lib().bwa_b200_region_opt_default(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def alnregs_from_regions(regions: np.ndarray) -> np.ndarray:
    """the regions bwa_b200_align_* returns (REGION_DTYPE) as the records bwa_b200_finish_regions_host takes: the mem_alnreg_t fields
    mem_chain2aln and the extension results have filled by then (src/bwamem.c:1263-1472, 2297-2306); sub / csub / sub_n stay 0 as in the
    fork, which never fills them before this stage; secondary = -1"""
    a = np.zeros(regions.size, ALNREG_DTYPE)
    for k in ("rb", "re", "qb", "qe", "rid", "score", "truesc", "w", "seedcov", "seedlen0", "frac_rep"):
        a[k] = regions[k]
    a["secondary"] = -1
    a["secondary_all"] = -1
    return a


def finish_regions(index: Index, packed, word_off, read_len, regs, region_off, opt: RegionOpt, ctg_alt=None, first_read_id=0):
    """mem_sort_dedup_patch -> is_alt -> mem_mark_primary_se -> mapq for a batch (bwa_b200_finish_regions_host).
    regs: ALNREG_DTYPE array, read r's regions at [region_off[r], region_off[r+1]).  Returns (list of per-read arrays, n_pri)."""
    n = read_len.size
    a = np.ascontiguousarray(regs, dtype=ALNREG_DTYPE).copy()
    off = np.ascontiguousarray(region_off, dtype=np.uint64)
    n_out = np.zeros(max(n, 1), np.uint32)
    n_pri = np.zeros(max(n, 1), np.int32)
    alt = np.ascontiguousarray(ctg_alt, dtype=np.int32) if ctg_alt is not None else None
    check(lib().bwa_b200_finish_regions_host(index.h, 0 if alt is None else alt.size, _p(alt) if alt is not None else None, _p(packed), _p(word_off),
                                             _p(read_len), n, _p(off), _p(a), _p(n_out), _p(n_pri), first_read_id, C.byref(opt)))
    return [a[int(off[r]):int(off[r]) + int(n_out[r])] for r in range(n)], n_pri[:n]


def measure_int_alu(device: int = 0) -> dict:
    """issue rates of the extension kernels' integer instructions (warp-instructions per clock per SM), measured on the device"""
    rates = (C.c_double * 16)()
    mhz, n_sm = C.c_double(0), C.c_int(0)
    n = lib().bwa_b200_measure_int_alu(device, rates, 16, C.byref(mhz), C.byref(n_sm))
    if n < 0:
        check(n)
    return {"sm_mhz": mhz.value, "n_sm": n_sm.value,
            "warp_inst_per_clk_per_sm": {lib().bwa_b200_int_alu_op_name(i).decode(): rates[i] for i in range(n)}}


SW_RESULT_DTYPE = np.dtype([("score", "<i4"), ("te", "<i4"), ("qe", "<i4"), ("score2", "<i4"), ("te2", "<i4"), ("tb", "<i4"), ("qb", "<i4")])
KSW_XBYTE, KSW_XSTOP, KSW_XSUBO, KSW_XSTART = 0x10000, 0x20000, 0x40000, 0x80000


class LocalAligner:
    """ksw_align2 on the device (bwa_b200_sw_*): the Smith-Waterman call of mate rescue and mem_seed_sw"""

    def __init__(self, device: int = 0):
        self.h = vp()
        check(lib().bwa_b200_sw_create(device, C.byref(self.h)))

    def align2_host(self, jobs: dict, ext_p: ExtParams) -> np.ndarray:
        """jobs = dict(qseq, qoff, qlen, tseq, toff, tlen, xtra), byte-per-base codes 0..4; returns SW_RESULT_DTYPE records (kswr_t)"""
        a = {k: np.ascontiguousarray(jobs[k], dtype=(np.uint8 if k in ("qseq", "tseq") else np.uint32)) for k in ("qseq", "qoff", "qlen", "tseq", "toff", "tlen", "xtra")}
        n = a["qlen"].size
        out = np.zeros(max(n, 1), SW_RESULT_DTYPE)
        check(lib().bwa_b200_sw_align2_host(self.h, C.byref(ext_p), n, _p(a["qseq"]), a["qseq"].size, _p(a["qoff"]), _p(a["qlen"]),
                                            _p(a["tseq"]), a["tseq"].size, _p(a["toff"]), _p(a["tlen"]), _p(a["xtra"]), _p(out)))
        return out[:n]

    @property
    def launches(self) -> int:
        return int(lib().bwa_b200_sw_launches(self.h))

    @property
    def last_kernel_ms(self) -> float:
        return float(lib().bwa_b200_sw_last_kernel_ms(self.h))

    def destroy(self):
        if self.h:
            lib().bwa_b200_sw_destroy(self.h)
            self.h = None
