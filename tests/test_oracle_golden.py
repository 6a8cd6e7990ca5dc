"""CPU suite: the oracle against the golden vectors minted from the reference's own Host code,
and the generators against independent scipy constructions."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_gold


def test_vector_lanczos_matches_reference_host_bitwise(orc, maxwell10):
    g = load_gold("maxwell_N10_vector_m100.npz")
    r = orc.vector_lanczos(maxwell10["csr"], g["b"], 100, lc=int(g["lc"]))
    assert r["steps"] == 100
    assert np.array_equal(r["alpha"], g["alpha"])
    assert np.array_equal(r["beta"], g["beta"])
    assert np.array_equal(r["q"], g["q"])


def test_block_lanczos_matches_reference_host(orc, maxwell10):
    for nc in (4, 8):
        g = load_gold("maxwell_N10_block%d_m25.npz" % nc)
        n = maxwell10["n"]
        B = g["B"].reshape(nc, n).T
        r = orc.block_lanczos(maxwell10["csr"], B, 25, lc=int(g["lc"]))
        a = g["alpha"].reshape(25, nc, nc).transpose(0, 2, 1)
        b = g["beta"].reshape(26, nc, nc).transpose(0, 2, 1)
        assert np.array_equal(r["alpha"], a)
        assert np.array_equal(r["beta"], b)
        assert np.array_equal(r["q"], g["q"])
        # polar factors are symmetric
        assert np.max(np.abs(b[1:25] - b[1:25].transpose(0, 2, 1))) < 1e-14


def test_maxwell_matrix_facts(maxwell10):
    # n = 3N(N+1)(2N+1) = 6930, 26400 true nnz, 2-4 per row, zero diagonal, symmetric
    rp, ci, va = maxwell10["csr"]
    n = maxwell10["n"]
    assert n == 6930 and len(ci) == 26400
    lens = np.diff(rp)
    assert lens.min() == 2 and lens.max() == 4
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    assert abs(A.diagonal()).max() == 0.0
    assert abs(A - A.T).max() < 1e-16
    assert abs(abs(A).max() - 8.26e-3) < 1e-5


def test_spmv_oracle_vs_scipy(orc, maxwell10):
    rp, ci, va = maxwell10["csr"]
    n = maxwell10["n"]
    x = orc.start_vector(n, 7)
    y = orc.spmv(maxwell10["csr"], x)
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    assert np.max(np.abs(y - A @ x)) < 1e-16 * 10


def test_laplacian_generators(orc):
    nx, ny, nz = 7, 5, 4
    rp, ci, va = orc.lap2d(nx, ny)
    Tx = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(nx, nx))
    Ty = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(ny, ny))
    L2 = sp.kron(sp.eye(ny), Tx) + sp.kron(Ty, sp.eye(nx))
    A = sp.csr_matrix((va, ci, rp), shape=(nx * ny, nx * ny))
    assert abs(A - L2).max() == 0 and len(va) == 5 * nx * ny - 2 * nx - 2 * ny
    assert np.all(np.diff(ci)[np.setdiff1d(np.arange(len(ci) - 1), rp[1:-1] - 1)] > 0)   # ascending columns per row
    rp, ci, va = orc.lap3d(nx, ny, nz)
    Tz = sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(nz, nz))
    L3 = (sp.kron(sp.eye(nz), sp.kron(sp.eye(ny), Tx)) + sp.kron(sp.eye(nz), sp.kron(Ty, sp.eye(nx)))
          + sp.kron(Tz, sp.eye(nx * ny)))
    A = sp.csr_matrix((va, ci, rp), shape=(nx * ny * nz, nx * ny * nz))
    assert abs(A - L3).max() == 0
    # BASELINE sizes: closed-form nnz
    assert orc.lib().orc_lap2d_nnz(4096, 4096) == 83869696
    assert orc.lib().orc_lap3d_nnz(256, 256, 256) == 117047296
    assert orc.lib().orc_lap3d_nnz(512, 512, 512) == 937951232


def test_start_vectors(orc):
    v = orc.start_vector(1000, 0x5EED)
    assert -1 <= v.min() and v.max() < 1 and abs(v.mean()) < 0.1
    # splitmix64 known answer (reference implementation of Vigna): first output for seed 0
    assert orc.lib().orc_splitmix64(0) == 0xE220A8397B1DCDAF
    V = orc.start_block(100, 4, 0x5EED)
    assert V.shape == (100, 4) and V[3, 2] == 2.0 * ((orc.lib().orc_splitmix64(0x5EED ^ 14) >> 11) * 2.0 ** -53) - 1.0


def test_sqrtm_and_assemble(orc):
    rng = np.random.default_rng(0)
    for b in (1, 4, 16, 32):
        M = rng.standard_normal((3 * b, b))
        S = M.T @ M
        R, Ri = orc.sqrtm(S)
        assert np.max(np.abs(R @ R - S)) < 1e-12 * np.abs(S).max()
        assert np.max(np.abs(R @ Ri - np.eye(b))) < 1e-10
    a = rng.standard_normal((3, 2, 2)); bt = rng.standard_normal((4, 2, 2))
    T = orc.assemble_T(a, bt)
    assert np.array_equal(T[0:2, 0:2], a[0]) and np.array_equal(T[2:4, 2:4], a[1])
    assert np.array_equal(T[0:2, 2:4], bt[1]) and np.array_equal(T[2:4, 0:2], bt[1].T)
    assert np.all(T[0:2, 4:6] == 0)


def test_reorth_keeps_coefficients(orc, maxwell10):
    # with/without full reorthogonalisation alpha/beta agree to ~1e-11 over 100 steps here
    g = load_gold("maxwell_N10_vector_m100.npz")
    r = orc.vector_lanczos(maxwell10["csr"], g["b"], 100, lc=int(g["lc"]), reorth=1, want_basis=True)
    scale = np.abs(g["beta"][1:]).max()
    assert np.max(np.abs(r["alpha"] - g["alpha"])) < 1e-9 * scale
    assert np.max(np.abs(r["beta"] - g["beta"]) / np.abs(g["beta"])) < 1e-9
    V = r["V"]
    assert np.max(np.abs(V.T @ V - np.eye(100))) < 1e-13


def test_ritz_oracle_on_laplacian(orc):
    nx = ny = 24
    csr = orc.lap2d(nx, ny)
    b = orc.start_vector(nx * ny)
    m = 120
    r = orc.vector_lanczos(csr, b, m, reorth=1)
    theta, _ = orc.ritz(r["alpha"], r["beta"], 6)
    lam = np.sort([4 - 2 * np.cos(i * np.pi / (nx + 1)) - 2 * np.cos(j * np.pi / (ny + 1))
                   for i in range(1, nx + 1) for j in range(1, ny + 1)])
    assert abs(theta[0] - lam[0]) < 1e-8 and abs(theta[-1] - lam[-1]) < 1e-8


def test_oracle_matches_reference_cuda_run(orc, maxwell10):
    """tests/golden/ref_cuda_*.npz: the reference's own CUDA drivers (sm_100 build, run on a B200 by
    tools/run_ref_cuda.sh; cuBLAS reductions, cuSOLVER syevj square roots).  The oracle and the
    Host-minted goldens agree with them to the north_star tolerance over the first 50 steps."""
    r, g = load_gold("ref_cuda_vector_N10.npz"), load_gold("maxwell_N10_vector_m100.npz")
    o = orc.vector_lanczos(maxwell10["csr"], g["b"], 100, lc=int(r["lc"]))
    scale = np.maximum(np.abs(o["alpha"][:50]), np.mean(o["beta"][1:50]))
    assert np.max(np.abs(r["alpha"][:50] - o["alpha"][:50]) / scale) < 1e-10
    assert np.max(np.abs(r["beta"][:50] - o["beta"][:50]) / o["beta"][:50]) < 1e-10
    for nc in (4, 8):
        r, g = load_gold("ref_cuda_block%d_N10.npz" % nc), load_gold("maxwell_N10_block%d_m25.npz" % nc)
        nb = 10 * nc * nc
        assert np.max(np.abs(r["alpha"][:nb] - g["alpha"][:nb])) < 1e-10 * np.abs(g["alpha"]).max()
        assert np.max(np.abs(r["beta"][:nb] - g["beta"][:nb])) < 1e-10 * np.abs(g["beta"][:nb]).max()


@pytest.mark.parametrize("N", [2, 3, 5, 10])
def test_maxwell_closed_form_matches_reference_builder(orc, N):
    """The per-row closed form that csrc/lz_maxwell.cu evaluates on the device (numpy restatement in oracle/orc.py)
    against the arrays minted from the reference's own host builder: D, its column ids, W and A = D W, exactly."""
    g = load_gold("maxwell_N%d_matrix.npz" % N)
    n = int(g["n_rows"])
    Dv, Dc, Wv, Av = orc.maxwell_closed_form(N)
    assert Dv.shape == (n, 4) and n == 3 * N * (N + 1) * (2 * N + 1)
    assert np.array_equal(Dv, g["D_data"].reshape(4, n).T) and np.array_equal(Dc, g["D_idx"].reshape(4, n).T)
    assert np.array_equal(Wv, g["W_data"])
    assert np.array_equal(Av, g["ell_data"].reshape(4, n).T) and np.array_equal(Dc, g["ell_idx"].reshape(4, n).T)
