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
