"""CPU suite: the C-ABI library loads, exports every symbol include/*.h declares, and refuses to
compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import glob
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(lz_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_symbols():
    names = declared_symbols()
    assert len(names) >= 40
    for must in ("lz_spmv", "lz_spmm", "lz_vector_lanczos", "lz_block_lanczos", "lz_mm_tt", "lz_mm_tt2", "lz_mm_ts",
                 "lz_sqrtm", "lz_ritz", "lz_csr_create", "lz_ell_create", "lz_vector_lanczos_sharded"):
        assert must in names


def test_library_exports_every_declared_symbol(lz):
    L = lz.lib()
    missing = [n for n in declared_symbols() if not hasattr(L, n)]
    assert not missing, missing
    # and the Python binding knows a signature for each of them
    unbound = [n for n in declared_symbols() if n not in L._signatures]
    assert not unbound, unbound
    assert L.lz_version() >= 100


def test_no_cpu_fallback(lz):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    st = lz.lib().lz_ctx_create(0, None, C.byref(h))
    assert st == -2
    assert b"no CPU fallback" in lz.lib().lz_last_error()
    with pytest.raises(lz.LanczosError):
        lz.Context(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "gpu-implementation-of-signle-and-block-lanczos_b200")
    bad = []
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                txt = open(os.path.join(root, f), errors="replace").read()
                if re.search(r"oracle|liboracle|orc_", txt):
                    bad.append(os.path.join(root, f))
    assert not bad, bad


def test_host_entry_points_validate_arguments():
    """Host-side entry points (no device needed): error codes instead of crashes on bad input."""
    import ctypes as C
    import numpy as np
    import lanczos_b200 as lz
    L = lz.lib()
    a = np.zeros(4); b = np.ones(4); th = np.zeros(2); rs = np.zeros(2)
    assert L.lz_ritz(0, 1, a.ctypes.data, b.ctypes.data, None, 1, th.ctypes.data, rs.ctypes.data) != 0      # m < 1
    assert L.lz_ritz(4, 1, a.ctypes.data, b.ctypes.data, None, 9, th.ctypes.data, rs.ctypes.data) != 0      # k > dim(T)
    assert L.lz_expm_sym(0, a.ctypes.data) != 0 and L.lz_expm_sym(2, None) != 0
    assert L.lz_expm_sym(4096, a.ctypes.data) != 0                                                          # too large for the host solver
    assert L.lz_lanczos_solution(2, 1, a.ctypes.data, b.ctypes.data, None, 1.0, th.ctypes.data) != 0
    lo, hi = C.c_int64(), C.c_int64()
    assert L.lz_partition_rows(100, 7, 2, 0, C.byref(lo), C.byref(hi)) != 0                                 # rows not a multiple of the granule
    assert L.lz_partition_rows(64, 32, 4, 0, C.byref(lo), C.byref(hi)) != 0                                 # fewer granules than ranks
    assert L.lz_partition_rows(64, 8, 4, 3, C.byref(lo), C.byref(hi)) == 0 and (lo.value, hi.value) == (48, 64)
    assert b"lz_" in L.lz_last_error() or len(L.lz_last_error()) > 0
    # expm of a diagonal matrix is the elementwise exponential of the diagonal
    D = np.diag([0.5, -1.0, 2.0])
    assert np.allclose(lz.expm_sym(D), np.diag(np.exp([0.5, -1.0, 2.0])), rtol=1e-14, atol=1e-15)
