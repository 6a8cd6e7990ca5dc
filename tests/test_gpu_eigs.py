"""GPU suite, SURVEY.md 8f-1 / 8f-4: Ritz vectors, thick-restart Lanczos, selective reorthogonalisation and the
checkpoint of a run -- through the C-ABI, against the oracle's numpy restatement (oracle/orc.py) and, where the
operator allows it, against the analytic spectrum of the Laplacian.

Tolerances: converged Ritz values and residual norms to 1e-8 (north_star); a resumed run bit-identical to the
uninterrupted one."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def lap2d_spectrum(nx, ny):
    return np.sort(np.array([4 - 2 * np.cos(np.pi * i / (nx + 1)) - 2 * np.cos(np.pi * j / (ny + 1))
                             for i in range(1, nx + 1) for j in range(1, ny + 1)]))


def true_residuals(lz, ctx, A, X, theta, n, k):
    """|| A x_i - theta_i x_i || column by column on the device (lz_spmv + torch)."""
    out = []
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    for i in range(k):
        xi = X[i * n:(i + 1) * n]
        lz.spmv(ctx, A, xi, y)
        ctx.sync()
        out.append(float(torch.linalg.norm(y - theta[i] * xi)))
    return np.array(out)


@pytest.mark.parametrize("which", [0, 1, 2])
def test_thick_restart_vs_oracle_and_analytic_spectrum(lz, ctx, orc, which):
    nx, ny, k, m_max = 40, 36, 6, 30
    n = nx * ny
    csr = orc.lap2d(nx, ny)
    b = orc.start_vector(n)
    A = lz.Matrix.laplacian2d(ctx, nx, ny)
    X = torch.zeros(n * k, dtype=torch.float64, device="cuda")
    theta, resid, info = lz.eigs_thick_restart(ctx, A, dev(b), k, which=which, m_max=m_max, tol=1e-10, X=X, ldx=n)
    ctx.sync()
    assert info["converged"] == k and info["basis"] == m_max and info["restarts"] >= 1
    th_o, res_o, X_o, info_o = orc.thick_restart_lanczos(csr, b, k, which, m_max, 1e-10)
    assert info_o["converged"] == k
    assert np.max(np.abs(theta - th_o)) < 1e-9
    lam = lap2d_spectrum(nx, ny)
    want = lam[:k] if which == 0 else lam[-k:] if which == 1 else np.r_[lam[:k // 2], lam[-(k - k // 2):]]
    assert np.max(np.abs(theta - want)) < 1e-8
    # Ritz vectors: orthonormal, and the TRUE residuals agree with the estimates |beta_m y_m| (both below 1e-8)
    Xh = X.cpu().numpy().reshape(k, n)
    assert np.max(np.abs(Xh @ Xh.T - np.eye(k))) < 1e-10
    tr = true_residuals(lz, ctx, A, X, theta, n, k)
    assert np.max(tr) < 1e-8 and np.max(resid) < 1e-8
    assert np.max(np.abs(tr - resid)) < 1e-8
    # up to sign the vectors are the oracle's
    for i in range(k):
        assert min(np.linalg.norm(Xh[i] - X_o[:, i]), np.linalg.norm(Xh[i] + X_o[:, i])) < 1e-6
    A.close()


def test_thick_restart_bounded_basis_3d(lz, ctx, orc):
    """config-3 style target at reduced size: 16 smallest eigenpairs of the 7-point Laplacian on 32 x 28 x 24 inside a
    48-vector basis (an unrestarted run would need several hundred stored vectors)."""
    nx, ny, nz, k, m_max = 32, 28, 24, 16, 48
    n = nx * ny * nz
    A = lz.Matrix.laplacian3d(ctx, nx, ny, nz)
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_vector(ctx.h, n, 0x5EED, b.data_ptr()))
    theta, resid, info = lz.eigs_thick_restart(ctx, A, b, k, which=0, m_max=m_max, tol=1e-9, max_restarts=400)
    assert info["converged"] == k, info
    lam = np.sort(np.array([6 - 2 * np.cos(np.pi * i / (nx + 1)) - 2 * np.cos(np.pi * j / (ny + 1)) - 2 * np.cos(np.pi * l / (nz + 1))
                            for i in range(1, 5) for j in range(1, 5) for l in range(1, 5)]))[:k]
    assert np.max(np.abs(theta - lam)) < 1e-8, np.max(np.abs(theta - lam))
    A.close()


def test_ritz_vectors_are_V_times_Y(lz, ctx, orc):
    nx, ny, m, k = 48, 40, 70, 9
    n = nx * ny
    A = lz.Matrix.laplacian2d(ctx, nx, ny)
    b = dev(orc.start_vector(n))
    alpha, beta, steps = lz.vector_lanczos(ctx, A, b, m, reorth=lz.REORTH_FULL)
    T = np.diag(alpha) + np.diag(beta[1:], 1) + np.diag(beta[1:], -1)
    w, Y = np.linalg.eigh(T)
    sel = np.r_[np.arange(4), np.arange(m - 5, m)]
    ld = n + 32
    X = torch.zeros(ld * k, dtype=torch.float64, device="cuda")
    lz.ritz_vectors(ctx, Y[:, sel], X, ld)
    Vd = torch.empty(n * m, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_vector_basis_copy(ctx.h, 0, m, Vd.data_ptr(), n))
    ctx.sync()
    V = Vd.cpu().numpy().reshape(m, n).T
    Xh = X.cpu().numpy().reshape(k, ld)[:, :n].T
    assert np.max(np.abs(Xh - V @ Y[:, sel])) < 1e-13
    assert np.all(X.cpu().numpy().reshape(k, ld)[:, n:] == 0)               # padding rows untouched
    # estimate |beta_m y_m| against the true residual for the extremal pairs
    bm = lz.last_coupling(ctx, 1)[0, 0]
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    for c, i in enumerate(sel):
        xi = X[c * ld:c * ld + n]
        lz.spmv(ctx, A, xi, y)
        ctx.sync()
        tr = float(torch.linalg.norm(y - w[i] * xi))
        assert abs(tr - abs(bm * Y[m - 1, i])) < 1e-8
    A.close()


@pytest.mark.parametrize("reorth", [0, 1, 3])
def test_checkpoint_resume_is_bit_identical(lz, orc, tmp_path, reorth):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    nx, ny, m, cut = 56, 44, 64, 27
    n = nx * ny
    torch.zeros(1, device="cuda")
    b = dev(orc.start_vector(n))
    ctx1 = lz.Context(0)
    A1 = lz.Matrix.laplacian2d(ctx1, nx, ny)
    a_ref, b_ref, steps = lz.vector_lanczos(ctx1, A1, b, m, lc=5, reorth=reorth)
    assert steps == m
    # the same run in two pieces with a save in between ...
    lz.vector_lanczos_begin(ctx1, A1, b, m, lc=5, reorth=reorth)
    a1, b1, done = lz.vector_lanczos_advance(ctx1, cut, m)
    assert done == cut and np.array_equal(a1, a_ref[:cut]) and np.array_equal(b1, b_ref[:cut])
    path = tmp_path / "run.ckpt"
    lz.checkpoint_save(ctx1, path)
    ctx1.close()
    # ... restored on a fresh context
    ctx2 = lz.Context(0)
    A2 = lz.Matrix.laplacian2d(ctx2, nx, ny)
    lz.checkpoint_load(ctx2, A2, path)
    a2, b2, done = lz.vector_lanczos_advance(ctx2, m - cut, m)
    assert done == m
    assert np.array_equal(a2, a_ref) and np.array_equal(b2, b_ref)
    # a checkpoint of another operator is refused
    A3 = lz.Matrix.laplacian2d(ctx2, nx, ny + 1)
    with pytest.raises(lz.LanczosError):
        lz.checkpoint_load(ctx2, A3, path)
    ctx2.close()


def test_selective_reorthogonalisation(lz, ctx, orc):
    """omega-recurrence mode: far fewer reorthogonalisation steps than full CGS2, semi-orthogonal basis (sqrt(eps)
    level), Ritz values to working precision, no spurious copies of converged eigenvalues."""
    nx, ny, m = 40, 36, 200
    n = nx * ny
    A = lz.Matrix.laplacian2d(ctx, nx, ny)
    b = dev(orc.start_vector(n))
    a_f, b_f, _ = lz.vector_lanczos(ctx, A, b, m, reorth=lz.REORTH_FULL)
    a_s, b_s, steps = lz.vector_lanczos(ctx, A, b, m, reorth=lz.REORTH_SELECTIVE)
    assert steps == m
    count = lz.reorth_count(ctx)
    assert 0 < count < m // 2, count
    T = lambda a, bb: np.diag(a) + np.diag(bb[1:], 1) + np.diag(bb[1:], -1)
    th_f, th_s = np.linalg.eigvalsh(T(a_f, b_f)), np.linalg.eigvalsh(T(a_s, b_s))
    # extremal Ritz values agree with the fully reorthogonalised run and with the spectrum; no ghost eigenvalues
    assert np.max(np.abs(th_s[:5] - th_f[:5])) < 1e-8 and np.max(np.abs(th_s[-5:] - th_f[-5:])) < 1e-8
    lam = lap2d_spectrum(nx, ny)
    assert abs(th_s[0] - lam[0]) < 1e-8 and abs(th_s[-1] - lam[-1]) < 1e-8
    assert np.min(np.diff(th_s[-6:])) > 1e-6 and np.min(np.diff(th_s[:6])) > 1e-6
    # semi-orthogonality of the stored basis
    Vd = torch.empty(n * m, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_vector_basis_copy(ctx.h, 0, m, Vd.data_ptr(), n))
    ctx.sync()
    V = Vd.view(m, n)
    G = (V @ V.T).cpu().numpy() - np.eye(m)
    assert np.max(np.abs(G)) < 1e-6, np.max(np.abs(G))
    # without any reorthogonalisation the same basis is far from orthogonal by then (the mode is doing something):
    # the selective run's estimates must have fired well before that level
    a_n, b_n, _ = lz.vector_lanczos(ctx, A, b, m, reorth=lz.REORTH_NONE)
    th_n = np.linalg.eigvalsh(T(a_n, b_n))
    assert abs(th_n[-1] - lam[-1]) < 1e-8          # (the extreme Ritz value itself is still fine)
    A.close()


@pytest.mark.parametrize("bw,which", [(8, 0), (16, 0), (8, 2)])
def test_block_thick_restart_degenerate_spectrum(lz, ctx, orc, bw, which):
    """config-3 style: the CUBIC 7-point Laplacian has three- and six-fold eigenvalues, which a single Lanczos vector
    cannot resolve.  Block thick restart (block CGS2 on the fp64 tensor pipe, DMMA basis compression) must return every
    copy: against the analytic spectrum with multiplicities, against the oracle's restarted run, and with true residuals."""
    nx = 12
    n, k, p = nx ** 3, 12, (6 if bw == 8 else 5)
    csr = orc.lap3d(nx, nx, nx)
    B = orc.start_block(n, bw)
    A = lz.Matrix.laplacian3d(ctx, nx, nx, nx)
    X = torch.zeros(n * k, dtype=torch.float64, device="cuda")
    theta, resid, info = lz.block_eigs_thick_restart(ctx, A, dev(np.ascontiguousarray(B.T).reshape(-1)), n, bw, k, which=which,
                                                     p_blocks=p, tol=1e-10, max_restarts=400, X=X, ldx=n)
    ctx.sync()
    assert info["converged"] == k and info["basis"] == p * bw, info
    th_o, res_o, X_o, info_o = orc.block_thick_restart_lanczos(csr, B, k, which, p, 1e-10, 400)
    assert info_o["converged"] == k
    assert np.max(np.abs(theta - th_o)) < 1e-8
    c = 2 * np.cos(np.pi * np.arange(1, nx + 1) / (nx + 1))
    lam = np.sort((6 - c[:, None, None] - c[None, :, None] - c[None, None, :]).reshape(-1))
    want = lam[:k] if which == 0 else np.r_[lam[:k // 2], lam[-(k - k // 2):]]
    assert np.max(np.abs(theta - want)) < 1e-8                 # multiplicities included (3-fold levels appear 3 times)
    Xh = X.cpu().numpy().reshape(k, n)
    assert np.max(np.abs(Xh @ Xh.T - np.eye(k))) < 1e-10
    tr = true_residuals(lz, ctx, A, X, theta, n, k)
    assert np.max(tr) < 1e-8 and np.max(resid) < 1e-8
    A.close()
