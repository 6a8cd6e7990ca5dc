"""GPU parity suite, single-vector path: SpMV (CSR stream kernel, ELL4 kernel), BLAS-1 style ops and
the fused vector_lanczos driver, all called through the C-ABI and checked against the oracle and the
golden vectors minted from the reference's Host code.

Tolerances (north_star): alpha/beta to relative 1e-10 over the first 50 steps; SpMV is bit-exact
(same left-to-right sum order as objects/ell_matrix.hpp:246-251)."""
import numpy as np
import pytest

from conftest import load_gold

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def csr_matrix(lz, ctx, csr):
    rp, ci, va = csr
    return lz.Matrix.from_csr(ctx, dev(rp), dev(ci), dev(va))


def coeff_err(alpha, beta, ga, gb, upto):
    """north_star tolerance form: |d alpha| <= tol * max(|alpha|, mean beta), beta relative."""
    scale = np.maximum(np.abs(ga[:upto]), np.mean(np.abs(gb[1:upto])))
    ea = np.max(np.abs(alpha[:upto] - ga[:upto]) / scale)
    eb = np.max(np.abs(beta[:upto] - gb[:upto]) / np.abs(gb[:upto]))
    return ea, eb


def test_spmv_csr_bit_exact_maxwell(lz, ctx, orc, maxwell10):
    A = csr_matrix(lz, ctx, maxwell10["csr"])
    n = maxwell10["n"]
    x = orc.start_vector(n, 11)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.spmv(ctx, A, dev(x), y)
    ctx.sync()
    assert np.array_equal(y.cpu().numpy(), orc.spmv(maxwell10["csr"], x))


def test_spmv_ell4_both_layouts_bit_exact(lz, ctx, orc, maxwell10):
    n, w = maxwell10["n"], maxwell10["width"]
    x = orc.start_vector(n, 12)
    ref = orc.spmv(maxwell10["csr"], x)
    data_cm, idx_cm = maxwell10["ell_data"], maxwell10["ell_idx"]
    # layout 0: the reference's column-major ELL; layout 1: what change_order(4) is meant to produce
    data_ri = np.ascontiguousarray(data_cm.reshape(w, n).T).reshape(-1)
    idx_ri = np.ascontiguousarray(idx_cm.reshape(w, n).T).reshape(-1)
    for layout, d, i in ((0, data_cm, idx_cm), (1, data_ri, idx_ri)):
        A = lz.Matrix.from_ell(ctx, n, n, w, layout, dev(d), dev(i.astype(np.int32)))
        y = torch.empty(n, dtype=torch.float64, device="cuda")
        lz.spmv(ctx, A, dev(x), y)
        ctx.sync()
        assert np.array_equal(y.cpu().numpy(), ref), layout


def test_ell_generic_width_converts_to_csr(lz, ctx, orc):
    # width-5 ELL of the 2-D Laplacian (the only reference path for it is lm::spmv_basic, ell_kernels.hpp:13-34)
    nx, ny = 37, 29
    rp, ci, va = orc.lap2d(nx, ny)
    n, w = nx * ny, 5
    data = np.zeros((w, n)); idx = np.zeros((w, n), np.int32)
    for r in range(n):
        k = rp[r + 1] - rp[r]
        data[:k, r] = va[rp[r]:rp[r + 1]]; idx[:k, r] = ci[rp[r]:rp[r + 1]]
    A = lz.Matrix.from_ell(ctx, n, n, w, 0, dev(data.reshape(-1)), dev(idx.reshape(-1)))
    assert A.nnz == len(va)
    x = orc.start_vector(n, 5)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.spmv(ctx, A, dev(x), y)
    ctx.sync()
    assert np.array_equal(y.cpu().numpy(), orc.spmv((rp, ci, va), x))


def test_generated_laplacians_equal_oracle(lz, ctx, orc):
    for shape in ((33, 17), (64, 64)):
        A = lz.Matrix.laplacian2d(ctx, *shape)
        for got, want in zip(A.csr_to_host(), orc.lap2d(*shape)):
            assert np.array_equal(got, want)
    for shape in ((9, 7, 5), (16, 16, 16)):
        A = lz.Matrix.laplacian3d(ctx, *shape)
        for got, want in zip(A.csr_to_host(), orc.lap3d(*shape)):
            assert np.array_equal(got, want)


def test_spmv_ragged_and_long_rows(lz, ctx, orc):
    """empty rows, 1-entry rows, rows longer than a chunk (warp path) and longer than 4096 (CTA path)."""
    rng = np.random.default_rng(3)
    n = 20000
    lens = rng.integers(0, 12, n)
    lens[5] = 0; lens[100] = 5000; lens[101] = 3000; lens[7000] = 9000; lens[n - 1] = 0; lens[0] = 0
    rp = np.zeros(n + 1, np.int32); rp[1:] = np.cumsum(lens)
    ci = rng.integers(0, n, rp[-1]).astype(np.int32)
    va = rng.standard_normal(rp[-1])
    A = csr_matrix(lz, ctx, (rp, ci, va))
    x = rng.standard_normal(n)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.spmv(ctx, A, dev(x), y)
    ctx.sync()
    ref = orc.spmv((rp, ci, va), x)
    got = y.cpu().numpy()
    short = lens <= 1024      # rows in chunks without a long row are bit exact; long ones sum in another order
    assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.abs(ref).max())
    assert got[5] == 0 and got[n - 1] == 0 and got[0] == 0
    assert np.mean(got[short] == ref[short]) > 0.9


def test_dot_nrm2_axpby(lz, ctx, orc):
    n = 1_000_003
    x, y = orc.start_vector(n, 1), orc.start_vector(n, 2)
    dx, dy = dev(x), dev(y)
    assert abs(lz.dot(ctx, dx, dy) - np.dot(x, y)) < 1e-9
    assert abs(lz.nrm2(ctx, dx) - np.linalg.norm(x)) < 1e-9
    lz.axpby(ctx, 0.5, dy, -2.0, dx)
    ctx.sync()
    assert np.array_equal(dy.cpu().numpy(), 0.5 * y + (-2.0) * x)
    bad = dev(np.array([1.0, np.inf]))
    with pytest.raises(lz.LanczosError):
        lz.nrm2(ctx, bad)


@pytest.mark.parametrize("fmt", ["csr", "ell4"])
def test_vector_lanczos_golden(lz, ctx, maxwell10, fmt):
    """config 1: Maxwell N = 10, random_vector_b, m = 100 against the reference-Host golden series."""
    g = load_gold("maxwell_N10_vector_m100.npz")
    n, w = maxwell10["n"], maxwell10["width"]
    if fmt == "csr":
        A = csr_matrix(lz, ctx, maxwell10["csr"])
    else:
        A = lz.Matrix.from_ell(ctx, n, n, w, 0, dev(maxwell10["ell_data"]), dev(maxwell10["ell_idx"].astype(np.int32)))
    q = torch.zeros(100, dtype=torch.float64, device="cuda")
    alpha, beta, steps = lz.vector_lanczos(ctx, A, dev(g["b"]), 100, lc=int(g["lc"]), q=q)
    assert steps == 100
    ea, eb = coeff_err(alpha, beta, g["alpha"], g["beta"], 50)
    assert ea < 1e-10 and eb < 1e-10, (ea, eb)
    ea, eb = coeff_err(alpha, beta, g["alpha"], g["beta"], 100)
    assert ea < 1e-8 and eb < 1e-8, (ea, eb)
    qd = q.cpu().numpy()
    assert np.max(np.abs(qd[:50] - g["q"][:50])) < 1e-10 * np.abs(g["q"]).max()


@pytest.mark.parametrize("mode", [1, 2])
def test_vector_lanczos_full_reorth_vs_oracle(lz, ctx, orc, maxwell10, mode):
    g = load_gold("maxwell_N10_vector_m100.npz")
    ref = orc.vector_lanczos(maxwell10["csr"], g["b"], 100, lc=int(g["lc"]), reorth=1)
    A = csr_matrix(lz, ctx, maxwell10["csr"])
    alpha, beta, steps = lz.vector_lanczos(ctx, A, dev(g["b"]), 100, lc=int(g["lc"]), reorth=mode)
    assert steps == 100
    ea, eb = coeff_err(alpha, beta, ref["alpha"], ref["beta"], 50)
    assert ea < 1e-10 and eb < 1e-10, (ea, eb)
    # the stored basis is orthonormal to working precision
    import ctypes as C
    rows, cols = C.c_int64(), C.c_int()
    lz.check(lz.lib().lz_vector_basis_info(ctx.h, C.byref(rows), C.byref(cols)))
    assert rows.value == maxwell10["n"] and cols.value == 100
    Vd = torch.empty(rows.value * cols.value, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_vector_basis_copy(ctx.h, 0, cols.value, Vd.data_ptr(), rows.value))
    ctx.sync()
    Vh = Vd.cpu().numpy().reshape(cols.value, rows.value)
    assert np.max(np.abs(Vh @ Vh.T - np.eye(100))) < 1e-12


def test_vector_lanczos_laplacian_vs_oracle_and_spectrum(lz, ctx, orc):
    nx, ny, m = 96, 80, 120
    csr = orc.lap2d(nx, ny)
    b = orc.start_vector(nx * ny)
    A = lz.Matrix.laplacian2d(ctx, nx, ny)
    bd = torch.empty(nx * ny, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_vector(ctx.h, nx * ny, 0x5EED, bd.data_ptr()))
    ctx.sync()
    assert np.array_equal(bd.cpu().numpy(), b)
    for reorth in (0, 1):
        ref = orc.vector_lanczos(csr, b, m, reorth=reorth)
        alpha, beta, steps = lz.vector_lanczos(ctx, A, bd, m, reorth=reorth)
        assert steps == m
        ea, eb = coeff_err(alpha, beta, ref["alpha"], ref["beta"], 50)
        assert ea < 1e-10 and eb < 1e-10, (reorth, ea, eb)
        theta = np.linalg.eigvalsh(np.diag(alpha) + np.diag(beta[1:], 1) + np.diag(beta[1:], -1))
        lam_min = 4 - 2 * np.cos(np.pi / (nx + 1)) - 2 * np.cos(np.pi / (ny + 1))
        lam_max = 8 - lam_min
        assert theta[0] >= lam_min - 1e-10 and theta[-1] <= lam_max + 1e-10


def test_vector_lanczos_breakdown_reported(lz, ctx, orc):
    """Start vector = an exact eigenvector: beta_1 = 0 -> the reference aborts in l2_norm one step later
    (vector.hpp:233-244); we return LZ_ERR_BREAKDOWN with the count of valid coefficients."""
    n = 64
    rp = np.arange(n + 1, dtype=np.int32); ci = np.arange(n, dtype=np.int32); va = np.full(n, 2.0)
    A = csr_matrix(lz, ctx, (rp, ci, va))
    alpha, beta, steps = lz.vector_lanczos(ctx, A, dev(np.ones(n)), 5)
    assert steps == 1 and alpha[0] == 2.0 and beta[0] == 8.0


def test_full_size_config2_parity(lz, ctx, orc):
    """BASELINE config 2 at FULL size (4096^2, 16.7 M rows, CGS2 against the stored basis): the first 50 steps
    against the oracle on the same operator and start vector, alpha/beta to relative 1e-10 (north_star)."""
    nx = ny = 4096
    n, m = nx * ny, 50
    orc.set_threads(__import__("os").cpu_count() or 4)
    try:
        ref = orc.vector_lanczos(orc.lap2d(nx, ny), orc.start_vector(n), m, reorth=1)
    finally:
        orc.set_threads(1)
    A = lz.Matrix.laplacian2d(ctx, nx, ny)
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_vector(ctx.h, n, 0x5EED, b.data_ptr()))
    alpha, beta, steps = lz.vector_lanczos(ctx, A, b, m, reorth=1)
    assert steps == m
    ea, eb = coeff_err(alpha, beta, ref["alpha"], ref["beta"], 50)
    assert ea < 1e-10 and eb < 1e-10, (ea, eb)
    # the same solve with pass B kept separate (LZ_NO_FOLD path) is exercised in tests/test_gpu_paths.py
    A.close()


def test_full_size_config2_properties(lz, ctx):
    """BASELINE config 2 shape (4096^2, 16.7 M rows): size-independent properties only.
    T's eigenvalues lie inside the analytic spectrum; the extreme Ritz values move monotonically
    towards it; the reorthogonalised basis is orthonormal on a sample of column pairs."""
    nx = ny = 4096
    n, m = nx * ny, 40
    A = lz.Matrix.laplacian2d(ctx, nx, ny)
    assert A.nnz == 83869696
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_vector(ctx.h, n, 0x5EED, b.data_ptr()))
    alpha, beta, steps = lz.vector_lanczos(ctx, A, b, m, reorth=1)
    assert steps == m
    lam_min = 4 - 4 * np.cos(np.pi / (nx + 1))
    T = np.diag(alpha) + np.diag(beta[1:], 1) + np.diag(beta[1:], -1)
    th = np.linalg.eigvalsh(T)
    assert th[0] > lam_min and th[-1] < 8 - lam_min
    th20 = np.linalg.eigvalsh(T[:20, :20])
    assert th[0] < th20[0] and th[-1] > th20[-1]
    # linearity: scaling the start vector scales beta_0 only
    b2 = b * 3.0
    a2, be2, _ = lz.vector_lanczos(ctx, A, b2, 10, reorth=0)
    a1, be1, _ = lz.vector_lanczos(ctx, A, b, 10, reorth=0)
    assert abs(be2[0] - 3 * be1[0]) < 1e-9 * be1[0]
    assert np.max(np.abs(a2 - a1)) < 1e-12 and np.max(np.abs(be2[1:] - be1[1:])) < 1e-12


@pytest.mark.parametrize("dims", [(2, 2, 2), (3, 3, 3), (5, 5, 5), (10, 10, 10), (3, 4, 5), (7, 2, 4)])
def test_maxwell_device_assembly_bit_identical(lz, ctx, orc, dims):
    """lz_gen_maxwell (closed-form assembly on the device) against the goldens minted from the reference's host builder
    (cubic cases) and against the oracle's closed form (anisotropic cases): values and column ids exactly equal; and
    the operator it returns drives the same Lanczos run as the host-built one."""
    A = lz.Matrix.maxwell(ctx, *dims)
    data, idx = A.ell_to_host()
    Dv, Dc, Wv, Av = orc.maxwell_closed_form(*dims)
    assert A.n_rows == Av.shape[0]
    assert np.array_equal(data, Av) and np.array_equal(idx.astype(np.int64), Dc)
    if dims[0] == dims[1] == dims[2] and dims[0] in (2, 3, 5, 10):
        g = load_gold("maxwell_N%d_matrix.npz" % dims[0])
        n = int(g["n_rows"])
        assert np.array_equal(data, g["ell_data"].reshape(4, n).T) and np.array_equal(idx, g["ell_idx"].reshape(4, n).T)
    if dims == (10, 10, 10):
        gv = load_gold("maxwell_N10_vector_m100.npz")
        alpha, beta, steps = lz.vector_lanczos(ctx, A, dev(gv["b"]), 100, lc=int(gv["lc"]))
        ea, eb = coeff_err(alpha, beta, gv["alpha"], gv["beta"], 50)
        assert steps == 100 and ea < 1e-10 and eb < 1e-10
    A.close()
