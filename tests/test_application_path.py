"""the application path of the reference harness -- expm(T_end T) e_1 contracted with the receiver
row q (test_lanczos.cu:97-110, :266-283) and the fdtd validator (methods/fdtd.hpp).

CPU part: the oracle restatements against the goldens minted from the reference's own Host containers
(fdtd bit for bit), and the library's host-side expm / solution entry points against the oracle and scipy.
GPU part: lz_fdtd_vector / lz_fdtd_block against the oracle, and the physics-level regression the reference
prints: Lanczos solution vs fdtd solution."""
import numpy as np
import pytest
import scipy.linalg

from conftest import load_gold


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def maxwell_csr(orc):
    gm = load_gold("maxwell_N10_matrix.npz")
    return int(gm["n_rows"]), orc.ell_to_csr(int(gm["n_rows"]), int(gm["width"]), gm["ell_data"], gm["ell_idx"])


def blocks(flat, m, bw):
    return flat[:m * bw * bw].reshape(m, bw, bw).transpose(0, 2, 1)


# ------------------------------------------------------------------------------------------ CPU

def test_oracle_fdtd_vector_bit_exact_vs_reference_host(orc):
    g = load_gold("maxwell_N10_vector_m100.npz")
    n, csr = maxwell_csr(orc)
    orc.set_threads(1)
    u = orc.fdtd_vector(csr, g["b"], int(g["fdtd_steps"]), 1.0)
    assert u[int(g["lc"])] == g["fdtd_u_lc"][0]


@pytest.mark.parametrize("nc", [4, 8])
def test_oracle_fdtd_block_bit_exact_vs_reference_host(orc, nc):
    g = load_gold("maxwell_N10_block%d_m25.npz" % nc)
    n, csr = maxwell_csr(orc)
    orc.set_threads(1)
    U = orc.fdtd_block(csr, g["B"].reshape(nc, n).T, int(g["fdtd_steps"]), 1.0)
    assert np.array_equal(U[int(g["lc"])], g["fdtd_row_lc"])


def test_expm_sym_library_vs_oracle_and_scipy(lz, orc):
    rng = np.random.default_rng(7)
    for n in (1, 2, 7, 33, 100):
        A = rng.standard_normal((n, n))
        A = (A + A.T) * 0.2
        E, Eo, Es = lz.expm_sym(A), orc.expm_sym(A), scipy.linalg.expm(A)
        assert np.max(np.abs(E - Es)) < 1e-12 * np.max(np.abs(Es))
        assert np.max(np.abs(E - Eo)) < 1e-12 * np.max(np.abs(Es))
    # only the lower triangle is read (uplo = LOWER, lib_utils.hpp:556)
    A = rng.standard_normal((6, 6)); A = (A + A.T) * 0.3
    B = A.copy(); B[np.triu_indices(6, 1)] = 99.0
    assert np.array_equal(lz.expm_sym(B), lz.expm_sym(A))


def test_lanczos_solution_vector_matches_reference_fdtd(lz, orc):
    """The golden alpha/beta/q (reference Host code) through expm reproduce the reference's fdtd value: the
    regression the harness prints as 'Relative error'."""
    g = load_gold("maxwell_N10_vector_m100.npz")
    s_lib = lz.lanczos_solution(g["alpha"], g["beta"], g["q"], 1.0)
    s_orc = orc.lanczos_solution(g["alpha"], g["beta"], g["q"], 1.0)
    T = orc.assemble_T(g["alpha"], g["beta"])
    s_np = g["beta"][0] * (scipy.linalg.expm(T)[:, 0] @ g["q"])
    assert abs(s_lib - s_orc) < 1e-12 * abs(s_orc) and abs(s_lib - s_np) < 1e-12 * abs(s_np)
    assert abs(s_lib - g["fdtd_u_lc"][0]) < 1e-9 * abs(g["fdtd_u_lc"][0])


@pytest.mark.parametrize("nc", [4, 8])
def test_lanczos_solution_block_matches_reference_fdtd(lz, orc, nc):
    g = load_gold("maxwell_N10_block%d_m25.npz" % nc)
    m = 25
    s_lib = lz.lanczos_solution(g["alpha"], g["beta"], g["q"], 1.0, bw=nc)
    s_orc = orc.lanczos_solution(blocks(g["alpha"], m, nc), blocks(g["beta"], m, nc), g["q"], 1.0)
    assert np.max(np.abs(s_lib - s_orc)) < 1e-12 * np.max(np.abs(s_orc))
    # Euler with 20000 steps carries an O(dt) error: agreement to 1e-6 here, 1e-9 at the vector case's 1e5 steps
    assert np.max(np.abs(s_lib - g["fdtd_row_lc"])) < 1e-6 * np.max(np.abs(g["fdtd_row_lc"]))


# ------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["csr", "ell"])
def test_fdtd_vector_vs_oracle(lz, ctx, orc, fmt):
    import torch
    g, gm = load_gold("maxwell_N10_vector_m100.npz"), load_gold("maxwell_N10_matrix.npz")
    n, csr = maxwell_csr(orc)
    if fmt == "csr":
        A = lz.Matrix.from_csr(ctx, dev(csr[0]), dev(csr[1]), dev(csr[2]))
    else:
        A = lz.Matrix.from_ell(ctx, n, n, 4, 0, dev(gm["ell_data"]), dev(gm["ell_idx"].astype(np.int32)))
    steps = 2000
    orc.set_threads(1)
    ref = orc.fdtd_vector(csr, g["b"], steps, 1.0)
    u_out = torch.empty(n, dtype=torch.float64, device="cuda")
    val = lz.fdtd_vector(ctx, A, dev(g["b"]), steps, 1.0, lc=int(g["lc"]), u_out=u_out)
    u = u_out.cpu().numpy()
    assert np.max(np.abs(u - ref)) < 1e-12 * np.max(np.abs(ref))          # fma vs mul+add: a few ulps per step
    assert val == u[int(g["lc"])]
    A.close()


@pytest.mark.gpu
def test_fdtd_vector_reference_step_count_matches_golden(lz, ctx, orc):
    """100000 steps as in test_lanczos.cu:118, against the value the reference's own Host code produced."""
    g = load_gold("maxwell_N10_vector_m100.npz")
    n, csr = maxwell_csr(orc)
    A = lz.Matrix.from_csr(ctx, dev(csr[0]), dev(csr[1]), dev(csr[2]))
    val = lz.fdtd_vector(ctx, A, dev(g["b"]), int(g["fdtd_steps"]), 1.0, lc=int(g["lc"]))
    assert abs(val - g["fdtd_u_lc"][0]) < 1e-11 * abs(g["fdtd_u_lc"][0])
    # and the Lanczos solution computed on the device agrees with it (the harness' relative error)
    alpha, beta, steps = lz.vector_lanczos(ctx, A, dev(g["b"]), 100, lc=int(g["lc"]), q=(q := dev(np.zeros(100))))
    sol = lz.lanczos_solution(alpha, beta, q.cpu().numpy(), 1.0)
    assert abs(sol - val) < 1e-9 * abs(val)
    A.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nc", [4, 8])
def test_fdtd_block_vs_oracle_and_golden(lz, ctx, orc, nc):
    g = load_gold("maxwell_N10_block%d_m25.npz" % nc)
    n, csr = maxwell_csr(orc)
    A = lz.Matrix.from_csr(ctx, dev(csr[0]), dev(csr[1]), dev(csr[2]))
    row = lz.fdtd_block(ctx, A, dev(g["B"]), n, nc, int(g["fdtd_steps"]), 1.0, lc=int(g["lc"]))
    assert np.max(np.abs(row - g["fdtd_row_lc"])) < 1e-11 * np.max(np.abs(g["fdtd_row_lc"]))
    A.close()


@pytest.mark.gpu
def test_fdtd_on_laplacian_matches_oracle(lz, ctx, orc):
    """Generated 2-D operator (negative-definite scaling so Euler is stable): device loop vs oracle loop."""
    import torch
    nx = ny = 96
    rp, ci, va = orc.lap2d(nx, ny)
    va = -0.05 * va
    A = lz.Matrix.from_csr(ctx, dev(rp), dev(ci), dev(va))
    u0 = orc.start_vector(nx * ny)
    ref = orc.fdtd_vector((rp, ci, va), u0, 500, 1.0)
    u_out = torch.empty(nx * ny, dtype=torch.float64, device="cuda")
    lz.fdtd_vector(ctx, A, dev(u0), 500, 1.0, lc=3, u_out=u_out)
    assert np.max(np.abs(u_out.cpu().numpy() - ref)) < 1e-12 * np.max(np.abs(ref))
    A.close()
