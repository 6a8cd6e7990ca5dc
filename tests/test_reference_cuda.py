"""The reference's own CUDA drivers (oracle/_ref/ref_cuda_dump_*, built from /root/reference with the
dead texture block stubbed -- oracle/build_ref_cuda.sh) run on the B200 next to the new library:
the oracle, the golden series and the product must all agree with what the reference itself computes
(cuBLAS dot/gemm order, cuSOLVER syevj for the block square roots)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_gold

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def run_ref(orc, tmp_path, nc, mode, N, m):
    exe = os.path.join(REFDIR, "ref_cuda_dump_%d" % nc)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_cuda_dump_%d not built (needs /root/reference at build time)" % nc)
    out = str(tmp_path / "ref.bin")
    r = subprocess.run([exe, mode, str(N), str(m), out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    return orc.read_dump(out)


def test_reference_cuda_vector_agrees_with_golden(orc, tmp_path):
    d, g = run_ref(orc, tmp_path, 4, "vector", 10, 100), load_gold("maxwell_N10_vector_m100.npz")
    assert d["lc"] == int(g["lc"])
    scale = np.maximum(np.abs(g["alpha"][:50]), np.mean(np.abs(g["beta"][1:50])))
    assert np.max(np.abs(d["alpha"][:50] - g["alpha"][:50]) / scale) < 1e-10
    assert np.max(np.abs(d["beta"][:50] - g["beta"][:50]) / np.abs(g["beta"][:50])) < 1e-10
    assert np.max(np.abs(d["q"][:50] - g["q"][:50])) < 1e-10 * np.abs(g["q"]).max()


@pytest.mark.parametrize("nc", [4, 8])
def test_reference_cuda_block_agrees_with_golden(orc, tmp_path, nc):
    m = 25
    d, g = run_ref(orc, tmp_path, nc, "block", 10, m), load_gold("maxwell_N10_block%d_m25.npz" % nc)
    nb = 10 * nc * nc
    assert np.max(np.abs(d["alpha"][:nb] - g["alpha"][:nb])) < 1e-10 * np.abs(g["alpha"]).max()
    assert np.max(np.abs(d["beta"][:nb] - g["beta"][:nb])) < 1e-10 * np.abs(g["beta"][:nb]).max()
    assert np.max(np.abs(d["q"][:10 * nc] - g["q"][:10 * nc])) < 1e-10 * np.abs(g["q"]).max()
    a = d["alpha"].reshape(m, nc, nc).transpose(0, 2, 1)
    b = d["beta"].reshape(m + 1, nc, nc).transpose(0, 2, 1)
    ga = g["alpha"].reshape(m, nc, nc).transpose(0, 2, 1)
    gb = g["beta"].reshape(m + 1, nc, nc).transpose(0, 2, 1)
    th, tg = np.linalg.eigvalsh(orc.assemble_T(a, b)), np.linalg.eigvalsh(orc.assemble_T(ga, gb))
    assert np.max(np.abs(th - tg)) < 1e-8 * np.abs(tg).max()          # north_star: Ritz values to 1e-8
