import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gpu-implementation-of-signle-and-block-lanczos_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_gold(name):
    return dict(np.load(os.path.join(GOLD, name)))


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as o
    o.lib()
    o.set_threads(1)
    return o


@pytest.fixture(scope="session")
def maxwell10(orc):
    """The reference's Maxwell operator at N = 10 (golden, minted from matrix_a/build_A_ell.hpp)."""
    g = load_gold("maxwell_N10_matrix.npz")
    n, w = int(g["n_rows"]), int(g["width"])
    csr = orc.ell_to_csr(n, w, g["ell_data"], g["ell_idx"])
    return dict(n=n, width=w, ell_data=g["ell_data"], ell_idx=g["ell_idx"], csr=csr)


@pytest.fixture(scope="session")
def lz():
    import lanczos_b200 as m
    m.lib()
    return m


@pytest.fixture(scope="session")
def ctx(lz):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")          # create the primary context torch and the library share
    c = lz.Context(0)
    yield c
    c.close()


def rel_err(a, b, floor=0.0):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
