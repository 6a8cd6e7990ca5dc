"""GPU suite: the alternative code paths behind the environment knobs and the public
reorthogonalisation modes give the same coefficients as the default path and the oracle.  Knobs are read once per
context, so each case creates its own context."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def coeff_err(alpha, beta, ra, rb, k):
    scale = np.maximum(np.abs(ra[:k]), np.mean(np.abs(rb[1:k])))
    return (float(np.max(np.abs(alpha[:k] - ra[:k]) / scale)), float(np.max(np.abs(beta[:k] - rb[:k]) / np.abs(rb[:k]))))


@pytest.fixture(scope="module")
def problem(orc):
    nx, ny, m = 160, 144, 80
    csr = orc.lap2d(nx, ny)
    b = orc.start_vector(nx * ny)
    ref = orc.vector_lanczos(csr, b, m, reorth=1)
    return dict(nx=nx, ny=ny, m=m, b=b, ref=ref)


@pytest.mark.parametrize("env", [{}, {"LZ_NO_FOLD": "1"}, {"LZ_NO_CGS_FUSE": "1"}, {"LZ_NO_FOLD": "1", "LZ_NO_CGS_FUSE": "1"},
                                 {"LZ_CGS_ONE_CTA": "1"}, {"LZ_CGS_NO_SLICES": "1"}, {"LZ_SPMV_VARIANT": "3"}])
def test_full_reorth_paths_agree_with_oracle(lz, orc, problem, monkeypatch, env):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    torch.zeros(1, device="cuda")
    ctx = lz.Context(0)
    A = lz.Matrix.laplacian2d(ctx, problem["nx"], problem["ny"])
    alpha, beta, steps = lz.vector_lanczos(ctx, A, dev(problem["b"]), problem["m"], reorth=lz.REORTH_FULL)
    assert steps == problem["m"]
    ea, eb = coeff_err(alpha, beta, problem["ref"]["alpha"], problem["ref"]["beta"], 50)
    assert ea < 1e-10 and eb < 1e-10, (env, ea, eb)
    # beta_m is exposed for residual estimates / restarts
    bl = lz.last_coupling(ctx, 1)
    assert np.isfinite(bl[0, 0]) and bl[0, 0] > 0
    ctx.close()


def test_dgks_mode_matches_cgs2_oracle(lz, ctx, orc, problem):
    """LZ_REORTH_FULL_DGKS (second sweep only when the first removed more than 1 - 1/sqrt 2 of w) against the oracle's
    unconditional CGS2: the skipped sweeps only change w at rounding level."""
    A = lz.Matrix.laplacian2d(ctx, problem["nx"], problem["ny"])
    alpha, beta, steps = lz.vector_lanczos(ctx, A, dev(problem["b"]), problem["m"], reorth=lz.REORTH_FULL_DGKS)
    assert steps == problem["m"]
    ea, eb = coeff_err(alpha, beta, problem["ref"]["alpha"], problem["ref"]["beta"], 50)
    assert ea < 1e-10 and eb < 1e-10, (ea, eb)
    # the basis it leaves behind is orthonormal to working precision
    import ctypes as C
    rows, cols = C.c_int64(), C.c_int()
    lz.check(lz.lib().lz_vector_basis_info(ctx.h, C.byref(rows), C.byref(cols)))
    Vd = torch.empty(rows.value * cols.value, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_vector_basis_copy(ctx.h, 0, cols.value, Vd.data_ptr(), rows.value))
    ctx.sync()
    V = Vd.view(cols.value, rows.value)
    G = (V @ V.T).cpu().numpy()
    assert np.max(np.abs(G - np.eye(cols.value))) < 1e-11
    A.close()


@pytest.mark.parametrize("env", [{}, {"LZ_NO_SPMM_GRAM": "1"}, {"LZ_NO_SPMM_FUSE": "1"}, {"LZ_SPMV_VARIANT": "9"}, {"LZ_NO_XS": "1"},
                                 {"LZ_NO_XS": "1", "LZ_NO_SPMM_GRAM": "1"}, {"LZ_XS_NO_TILES": "1", "LZ_XS_FORCE": "1"},
                                 {"LZ_XS_BOX": "16,4,2"}, {"LZ_XS_BOX": "8,2,1", "LZ_XS_STAGES": "2"}])
def test_block_paths_agree_with_oracle(lz, orc, monkeypatch, env):
    """b = 16: the operand-staging SpMM with box-shaped chunks, fused DMMA subtraction and Gram epilogue (default), the same
    with a separate Gram pass, the plain SpMM + two-Gram formulation, the LDG SpMM + two-Gram formulation, the gathering
    kernel (LZ_NO_XS) with and without the Gram epilogue, the staged kernel on runs of rows, other box shapes / ring depths."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    nx, ny, nz, bw, m = 24, 20, 18, 16, 10
    n = nx * ny * nz
    csr = orc.lap3d(nx, ny, nz)
    B = orc.start_block(n, bw)
    ref = orc.block_lanczos(csr, B, m, lc=2, reorth=0)
    torch.zeros(1, device="cuda")
    ctx = lz.Context(0)
    A = lz.Matrix.laplacian3d(ctx, nx, ny, nz)
    alpha = torch.zeros(m * bw * bw, dtype=torch.float64, device="cuda")
    beta = torch.zeros((m + 1) * bw * bw, dtype=torch.float64, device="cuda")
    q = torch.zeros(m * bw, dtype=torch.float64, device="cuda")
    lz.block_lanczos(ctx, A, dev(np.ascontiguousarray(B.T).reshape(-1)), n, bw, m, alpha, beta, q, lc=2)
    assert lz.block_status(ctx, m) == m
    a = alpha.cpu().numpy().reshape(m, bw, bw).transpose(0, 2, 1)
    b = beta.cpu().numpy().reshape(m + 1, bw, bw).transpose(0, 2, 1)
    err = lambda g, w: max(np.max(np.abs(g[j] - w[j])) / np.max(np.abs(w[j])) for j in range(m))
    assert err(a, ref["alpha"]) < 1e-10 and err(b, ref["beta"]) < 1e-10, env
    ctx.close()
