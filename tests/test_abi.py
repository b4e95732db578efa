"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes
import os
import re


def test_library_exports_header_symbols(pkg):
    hdr = open(os.path.join(pkg.ROOT, "include", "bwamem_b200.h")).read()
    declared = set(re.findall(r"\b(bwa_b200_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    L = pkg.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert set(pkg.SYMBOLS) == declared


def test_no_cpu_fallback_without_device(pkg):
    """compute entry points must fail loudly when there is no GPU (never route to a CPU path)"""
    L = pkg.lib()
    if L.bwa_b200_device_count() > 0:
        return
    h = ctypes.c_void_p()
    rc = L.bwa_b200_extender_create(0, 16, 1024, 1024, ctypes.byref(h))
    assert rc == -4  # BWA_B200_ERR_CUDA
    assert b"failed" in L.bwa_b200_last_error()


def test_pack_codes_and_ascii(pkg):
    import numpy as np
    codes = np.array([0, 1, 2, 3, 4, 0, 0, 0, 3, 3, 1], np.uint8)   # reads of 9 and 2 bases
    off = np.array([0, 9, 11], np.uint64)
    packed, woff, rl = pkg.pack_codes(codes, off)
    assert list(rl) == [9, 2] and list(woff) == [0, 2, 3]
    assert packed[0] == 0x01234000 and packed[1] == 0x34444444 and packed[2] == 0x31444444
    asc = np.frombuffer(b"ACGTNaaatXc", np.uint8)
    packed2, _, _ = pkg.pack_ascii(asc, off)
    assert packed2[0] == 0x01234000 and packed2[1] == 0x34444444 and packed2[2] == 0x41444444


def test_index_load_rejects_stock_layout(pkg, small_index, tmp_path):
    _, prefix = small_index
    L = pkg.lib()
    h = ctypes.c_void_p()
    rc = L.bwa_b200_index_load((prefix + ".bwt128").encode(), None, 0, ctypes.byref(h))
    assert rc in (-3, -4)     # format error (or no device before the check on a GPU-less box)
    if rc == -3:
        assert b"OCC_INTV_SHIFT 6" in L.bwa_b200_last_error()
    rc = L.bwa_b200_index_load(str(tmp_path / "missing.bwt").encode(), None, 0, ctypes.byref(h))
    assert rc == -2


def test_finish_regions_argument_checks_and_no_fallback(pkg):
    """bad arguments are refused before any device work; with good ones and no device the call fails loudly (no CPU path)"""
    import numpy as np
    L = pkg.lib()
    opt = pkg.region_opt()
    one = np.zeros(1, np.uint64)
    rc = L.bwa_b200_finish_regions_host(None, 0, None, None, one.ctypes.data, one.ctypes.data, 0, one.ctypes.data, None, one.ctypes.data,
                                        one.ctypes.data, 0, ctypes.byref(opt))
    assert rc == -1 and b"bad argument" in L.bwa_b200_last_error()          # BWA_B200_ERR_ARG: no index
    assert (opt.a, opt.b, opt.w, opt.min_seed_len, opt.max_chain_gap, opt.mapQ_coef_fac) == (1, 4, 100, 19, 10000, 3)
    assert abs(opt.mask_level_redun - 0.95) < 1e-6 and opt.mapQ_coef_len == 50.0
    assert pkg.ALNREG_DTYPE.itemsize == 96


def test_alnregs_from_regions_maps_the_alnreg_fields(pkg):
    import numpy as np
    r = np.zeros(3, pkg.REGION_DTYPE)
    r["rb"], r["re"], r["qb"], r["qe"] = [10, 2000, 5], [160, 2150, 90], [0, 3, 1], [150, 150, 86]
    r["rid"], r["score"], r["truesc"], r["w"], r["seedcov"], r["seedlen0"], r["frac_rep"] = [0, 1, 0], [140, 120, 60], [140, 118, 60], 100, [150, 80, 40], [19, 33, 25], [0, .5, 0]
    a = pkg.alnregs_from_regions(r)
    assert a.dtype == pkg.ALNREG_DTYPE and a.size == 3
    for k in ("rb", "re", "qb", "qe", "rid", "score", "truesc", "w", "seedcov", "seedlen0", "frac_rep"):
        assert (a[k] == r[k]).all(), k
    assert (a["secondary"] == -1).all() and (a["sub"] == 0).all() and (a["csub"] == 0).all() and (a["n_comp"] == 0).all() and (a["is_alt"] == 0).all()
