"""GPU parity suite, block path: SpMM, the tall-skinny dense products (DMMA for b in {8,16,32}, SIMT
otherwise), the b x b matrix square root and the block Lanczos driver -- through the C-ABI, against
the oracle and the golden series minted from the reference's Host containers.

Tolerances: SpMM bit-exact; dense products 1e-13 relative (different summation order); block
coefficients 1e-10 relative to the largest entry of the block over the first 10 blocks (north_star:
1e-10 over the first 50 scalar steps = 12 blocks of 4 / 6 blocks of 8)."""
import numpy as np
import pytest

from conftest import load_gold

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def cm(a):
    """(n, b) array -> device buffer in column-major storage."""
    return dev(np.ascontiguousarray(a.T).reshape(-1))


def from_cm(t, n, b, ld=None):
    ld = ld or n
    return t.cpu().numpy().reshape(b, ld)[:, :n].T


def csr_matrix(lz, ctx, csr):
    rp, ci, va = csr
    return lz.Matrix.from_csr(ctx, dev(rp), dev(ci), dev(va))


@pytest.mark.parametrize("b", [1, 4, 16])
def test_spmm_colmajor_bit_exact(lz, ctx, orc, maxwell10, b):
    n, w = maxwell10["n"], maxwell10["width"]
    X = orc.start_block(n, b, 3)
    ref = orc.spmm(maxwell10["csr"], X)
    ld = n + 5
    for A in (csr_matrix(lz, ctx, maxwell10["csr"]),
              lz.Matrix.from_ell(ctx, n, n, w, 0, dev(maxwell10["ell_data"]), dev(maxwell10["ell_idx"].astype(np.int32)))):
        Xd = torch.zeros(b * ld, dtype=torch.float64, device="cuda")
        Xd.view(b, ld)[:, :n] = dev(np.ascontiguousarray(X.T))
        Yd = torch.zeros(b * ld, dtype=torch.float64, device="cuda")
        lz.spmm(ctx, A, b, Xd, ld, Yd, ld)
        ctx.sync()
        assert np.array_equal(from_cm(Yd, n, b, ld), ref)


@pytest.mark.parametrize("b", [1, 3, 4, 8, 16, 32])
def test_dense_products(lz, ctx, orc, b):
    rng = np.random.default_rng(b)
    n, ld = 10007, 10016
    T1, T2 = rng.standard_normal((n, b)), rng.standard_normal((n, b))
    S = rng.standard_normal((b, b))
    pad = lambda a: np.concatenate([a, np.full((ld - n, b), np.nan)])     # padding must never be read
    d1, d2 = cm(pad(T1)), cm(pad(T2))
    L = lz.lib()
    R = torch.zeros(b * b, dtype=torch.float64, device="cuda")
    lz.check(L.lz_mm_tt(ctx.h, n, b, d1.data_ptr(), ld, R.data_ptr()))
    ctx.sync()
    G = R.cpu().numpy().reshape(b, b).T
    assert np.max(np.abs(G - T1.T @ T1)) < 1e-13 * np.abs(T1.T @ T1).max() * 10
    lz.check(L.lz_mm_tt2(ctx.h, n, b, d1.data_ptr(), ld, d2.data_ptr(), ld, R.data_ptr()))
    ctx.sync()
    G2 = R.cpu().numpy().reshape(b, b).T
    want = 0.5 * (T1.T @ T2) + 0.5 * (T2.T @ T1)
    assert np.max(np.abs(G2 - want)) < 1e-12 * max(1.0, np.abs(want).max())
    assert np.max(np.abs(G2 - G2.T)) < 1e-12
    # R = beta R + alpha T S, separate output and in place (R aliasing T, as mm_cublas(0,1,F1,beta0,F1))
    Sd = cm(S)
    Rd = cm(pad(T2))
    lz.check(L.lz_mm_ts(ctx.h, n, b, 1.0, -1.0, d1.data_ptr(), ld, Sd.data_ptr(), Rd.data_ptr(), ld))
    ctx.sync()
    got = from_cm(Rd, n, b, ld)
    assert np.max(np.abs(got - (T2 - T1 @ S))) < 1e-12 * np.abs(T1 @ S).max()
    inpl = cm(pad(T1))
    lz.check(L.lz_mm_ts(ctx.h, n, b, 0.0, 1.0, inpl.data_ptr(), ld, Sd.data_ptr(), inpl.data_ptr(), ld))
    ctx.sync()
    assert np.max(np.abs(from_cm(inpl, n, b, ld) - T1 @ S)) < 1e-12 * np.abs(T1 @ S).max()


@pytest.mark.parametrize("b", [1, 2, 4, 5, 16, 31, 32])
def test_sqrtm(lz, ctx, orc, b):
    rng = np.random.default_rng(100 + b)
    M = rng.standard_normal((4 * b + 3, b))
    S = M.T @ M
    Sl = np.tril(S) + np.triu(np.full((b, b), 7.0), 1)        # garbage above the diagonal: only the lower triangle is read
    Sd, Si = cm(Sl), torch.zeros(b * b, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_sqrtm(ctx.h, b, Sd.data_ptr(), Si.data_ptr()))
    ctx.sync()
    R, Ri = Sd.cpu().numpy().reshape(b, b).T, Si.cpu().numpy().reshape(b, b).T
    Ro, Rio = orc.sqrtm(S)
    assert np.max(np.abs(R - Ro)) < 1e-12 * np.abs(Ro).max()
    assert np.max(np.abs(Ri - Rio)) < 1e-10 * np.abs(Rio).max()
    assert np.max(np.abs(R @ R - S)) < 1e-12 * np.abs(S).max()


def test_copy_row_and_assemble_T(lz, ctx, orc):
    rng = np.random.default_rng(9)
    n, b, m = 50, 4, 3
    Q = rng.standard_normal((n, b))
    q = torch.zeros(3 * b, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_copy_row(ctx.h, 17, b, cm(Q).data_ptr(), n, q.data_ptr(), b))
    alpha, beta = rng.standard_normal((m, b, b)), rng.standard_normal((m, b, b))
    a_dev = dev(np.ascontiguousarray(alpha.transpose(0, 2, 1)).reshape(-1))
    b_dev = dev(np.ascontiguousarray(beta.transpose(0, 2, 1)).reshape(-1))
    T = torch.empty((m * b) ** 2, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_assemble_T(ctx.h, m, b, a_dev.data_ptr(), b_dev.data_ptr(), T.data_ptr()))
    ctx.sync()
    assert np.array_equal(q.cpu().numpy()[b:2 * b], Q[17])
    assert np.array_equal(T.cpu().numpy().reshape(m * b, m * b).T, orc.assemble_T(alpha, beta))


def run_block(lz, ctx, A, B, m, lc, reorth=0):
    n, bw = B.shape
    alpha = torch.zeros(m * bw * bw, dtype=torch.float64, device="cuda")
    beta = torch.zeros((m + 1) * bw * bw, dtype=torch.float64, device="cuda")
    q = torch.zeros(m * bw, dtype=torch.float64, device="cuda")
    lz.block_lanczos(ctx, A, cm(B), n, bw, m, alpha, beta, q, lc=lc, reorth=reorth)
    ctx.sync()
    a = alpha.cpu().numpy().reshape(m, bw, bw).transpose(0, 2, 1)
    b = beta.cpu().numpy().reshape(m + 1, bw, bw).transpose(0, 2, 1)
    return a, b, q.cpu().numpy()


def block_err(got, want, upto):
    return max(np.max(np.abs(got[j] - want[j])) / np.max(np.abs(want[j])) for j in range(upto))


@pytest.mark.parametrize("nc,fmt", [(4, "csr"), (8, "csr"), (4, "ell4"), (8, "ell4")])
def test_block_lanczos_golden(lz, ctx, maxwell10, nc, fmt):
    g = load_gold("maxwell_N10_block%d_m25.npz" % nc)
    n, w, m = maxwell10["n"], maxwell10["width"], 25
    if fmt == "csr":
        A = csr_matrix(lz, ctx, maxwell10["csr"])
    else:
        A = lz.Matrix.from_ell(ctx, n, n, w, 0, dev(maxwell10["ell_data"]), dev(maxwell10["ell_idx"].astype(np.int32)))
    B = g["B"].reshape(nc, n).T
    a, b, q = run_block(lz, ctx, A, B, m, int(g["lc"]))
    ga = g["alpha"].reshape(m, nc, nc).transpose(0, 2, 1)
    gb = g["beta"].reshape(m + 1, nc, nc).transpose(0, 2, 1)
    assert block_err(a, ga, 10) < 1e-10 and block_err(b, gb, 10) < 1e-10
    assert block_err(a, ga, m) < 1e-7 and block_err(b, gb, m) < 1e-7
    assert np.max(np.abs(q[:10 * nc] - g["q"][:10 * nc])) < 1e-10 * np.abs(g["q"]).max()
    # Ritz values of the assembled T agree much more tightly than the blocks
    from oracle import orc
    th = np.linalg.eigvalsh(orc.assemble_T(a, b))
    tg = np.linalg.eigvalsh(orc.assemble_T(ga, gb))
    assert np.max(np.abs(th - tg)) < 1e-12


@pytest.mark.parametrize("bw", [2, 16, 32])
def test_block_lanczos_laplacian_vs_oracle(lz, ctx, orc, bw):
    nx, ny, nz, m = 20, 18, 16, 8
    csr = orc.lap3d(nx, ny, nz)
    n = nx * ny * nz
    B = orc.start_block(n, bw)
    A = lz.Matrix.laplacian3d(ctx, nx, ny, nz)
    Bd = torch.empty(n * bw, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_block(ctx.h, n, bw, n, 0x5EED, Bd.data_ptr()))
    ctx.sync()
    assert np.array_equal(from_cm(Bd, n, bw), B)
    for reorth in (0, 1):
        ref = orc.block_lanczos(csr, B, m, lc=5, reorth=reorth)
        a, b, q = run_block(lz, ctx, A, B, m, 5, reorth=reorth)
        assert block_err(a, ref["alpha"], m) < 1e-10, (bw, reorth)
        assert block_err(b, ref["beta"], m) < 1e-10, (bw, reorth)
        assert np.max(np.abs(q - ref["q"])) < 1e-10
        th = np.linalg.eigvalsh(orc.assemble_T(a, b))
        assert th[0] > 0 and th[-1] < 12


def test_block_breakdown_is_reported(lz, ctx, orc):
    """A rank-deficient start block (two equal columns): B^T B is singular, the reference would carry Inf/NaN on
    silently; lz_block_status names the first failing block and returns LZ_ERR_BREAKDOWN."""
    nx, ny, nz, bw, m = 12, 10, 8, 8, 4
    n = nx * ny * nz
    A = lz.Matrix.laplacian3d(ctx, nx, ny, nz)
    B = orc.start_block(n, bw)
    good = run_block(lz, ctx, A, B, m, 0)
    assert lz.block_status(ctx, m) == m and np.all(np.isfinite(good[0]))
    B[:, 3] = B[:, 1]
    run_block(lz, ctx, A, B, m, 0)
    assert lz.block_status(ctx, m) == 0
    import ctypes as C
    done = C.c_int(-1)
    assert lz.lib().lz_block_status(ctx.h, m, C.byref(done)) == -4 and done.value == 0


@pytest.mark.parametrize("bw", [8, 16])
def test_block_recurrence_order_long_run_without_reorth(lz, ctx, orc, bw):
    """200 blocks without reorthogonalisation against the oracle, which follows the reference's order (W -= Q0 beta_j
    BEFORE alpha_j is formed, methods/block_lanczos.hpp:152-155).  bw = 16 runs the SpMM with the fused DMMA
    subtraction, bw = 8 the two-Gram formulation alpha_j = sym(G1 - G2 beta_j): both must reproduce the reference
    order to 1e-10 over the first 12 blocks (50 scalar steps at b = 4) and keep T's Ritz values to 1e-8 over the
    whole run, where round 1's pre-subtraction alpha let sym(beta_j^T Q0^T Q1) leak into alpha_j."""
    nx, ny, nz, m = 16, 14, 12, 200 if bw == 8 else 120
    csr = orc.lap3d(nx, ny, nz)
    n = nx * ny * nz
    B = orc.start_block(n, bw)
    A = lz.Matrix.laplacian3d(ctx, nx, ny, nz)
    ref = orc.block_lanczos(csr, B, m, lc=3, reorth=0)
    a, b, q = run_block(lz, ctx, A, B, m, 3, reorth=0)
    assert lz.block_status(ctx, m) == m
    assert block_err(a, ref["alpha"], 12) < 1e-10 and block_err(b, ref["beta"], 12) < 1e-10
    # alpha_j stays symmetric and T's extremal Ritz values agree over the whole run
    assert max(np.max(np.abs(a[j] - a[j].T)) for j in range(m)) < 1e-12
    th = np.linalg.eigvalsh(orc.assemble_T(a, b))
    tr = np.linalg.eigvalsh(orc.assemble_T(ref["alpha"], ref["beta"]))
    assert abs(th[0] - tr[0]) < 1e-8 and abs(th[-1] - tr[-1]) < 1e-8
    # beta_m (coupling to the unbuilt block) is exposed for residual estimates
    bl = lz.last_coupling(ctx, bw)
    assert np.all(np.isfinite(bl)) and np.max(np.abs(bl - bl.T)) < 1e-10 * np.abs(bl).max()


@pytest.mark.parametrize("reorth", [0, 1])
def test_full_size_config3_parity(lz, ctx, orc, reorth):
    """BASELINE config 3 at FULL size (256^3, b = 16): 6 blocks against the oracle on the same operator and start
    block, block coefficients to 1e-10 (north_star tolerance), with and without block CGS2."""
    nx, bw, m = 256, 16, 6
    n = nx ** 3
    orc.set_threads(__import__("os").cpu_count() or 4)
    try:
        csr = orc.lap3d(nx, nx, nx)
        B = orc.start_block(n, bw)
        ref = orc.block_lanczos(csr, B, m, lc=12345, reorth=reorth)
    finally:
        orc.set_threads(1)
    del csr
    A = lz.Matrix.laplacian3d(ctx, nx, nx, nx)
    Bd = torch.empty(n * bw, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_block(ctx.h, n, bw, n, 0x5EED, Bd.data_ptr()))
    alpha = torch.zeros(m * bw * bw, dtype=torch.float64, device="cuda")
    beta = torch.zeros((m + 1) * bw * bw, dtype=torch.float64, device="cuda")
    q = torch.zeros(m * bw, dtype=torch.float64, device="cuda")
    lz.block_lanczos(ctx, A, Bd, n, bw, m, alpha, beta, q, lc=12345, reorth=reorth)
    assert lz.block_status(ctx, m) == m
    a = alpha.cpu().numpy().reshape(m, bw, bw).transpose(0, 2, 1)
    b = beta.cpu().numpy().reshape(m + 1, bw, bw).transpose(0, 2, 1)
    assert block_err(a, ref["alpha"], m) < 1e-10, block_err(a, ref["alpha"], m)
    assert block_err(b, ref["beta"], m) < 1e-10, block_err(b, ref["beta"], m)
    assert np.max(np.abs(q.cpu().numpy() - ref["q"])) < 1e-10 * np.abs(ref["q"]).max()
    A.close()


@pytest.mark.parametrize("bw", [8, 16])
def test_block_dgks_mode_matches_cgs2_oracle(lz, ctx, orc, bw):
    """LZ_REORTH_FULL_DGKS for blocks: one block Gram-Schmidt sweep, a second one only when a column of W lost more than
    half of its squared norm (decided on the device).  Against the oracle's unconditional CGS2: blocks to 1e-10, and the
    basis it leaves is orthonormal to working precision."""
    nx, ny, nz, m = 20, 18, 16, 10
    csr = orc.lap3d(nx, ny, nz)
    n = nx * ny * nz
    B = orc.start_block(n, bw)
    A = lz.Matrix.laplacian3d(ctx, nx, ny, nz)
    ref = orc.block_lanczos(csr, B, m, lc=5, reorth=1, want_basis=True)
    a, b, q = run_block(lz, ctx, A, B, m, 5, reorth=lz.REORTH_FULL_DGKS)
    assert lz.block_status(ctx, m) == m
    assert block_err(a, ref["alpha"], m) < 1e-10 and block_err(b, ref["beta"], m) < 1e-10
    assert np.max(np.abs(q - ref["q"])) < 1e-10
    V = ref["V"]
    assert np.max(np.abs(V.T @ V - np.eye(m * bw))) < 1e-12          # (the oracle's basis: what the run must reproduce)
    A.close()


def test_full_size_config3_properties(lz, ctx):
    """BASELINE config 3 shape: 256^3 7-point Laplacian (16.7 M rows), b = 16.  Size-independent checks:
    beta blocks symmetric positive definite, alpha symmetric, Ritz values inside (0, 12), and the
    block-tridiagonal T reproduces the 3-term recurrence residual on a sampled column."""
    nx = 256
    n, bw, m = nx ** 3, 16, 6
    A = lz.Matrix.laplacian3d(ctx, nx, nx, nx)
    assert A.nnz == 117047296
    B = torch.empty(n * bw, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_block(ctx.h, n, bw, n, 0x5EED, B.data_ptr()))
    alpha = torch.zeros(m * bw * bw, dtype=torch.float64, device="cuda")
    beta = torch.zeros((m + 1) * bw * bw, dtype=torch.float64, device="cuda")
    q = torch.zeros(m * bw, dtype=torch.float64, device="cuda")
    lz.block_lanczos(ctx, A, B, n, bw, m, alpha, beta, q, lc=12345)
    ctx.sync()
    a = alpha.cpu().numpy().reshape(m, bw, bw).transpose(0, 2, 1)
    b = beta.cpu().numpy().reshape(m + 1, bw, bw).transpose(0, 2, 1)
    for j in range(m):
        assert np.max(np.abs(a[j] - a[j].T)) < 1e-12
        assert np.max(np.abs(b[j] - b[j].T)) < 1e-10 * np.abs(b[j]).max()
        assert np.linalg.eigvalsh(b[j])[0] > 0
    from oracle import orc
    th = np.linalg.eigvalsh(orc.assemble_T(a, b))
    assert th[0] > 0 and th[-1] < 12
    # beta[0]^2 = B^T B: check one entry against a direct dot product
    Bm = B.view(bw, n)
    assert abs((b[0] @ b[0])[2, 5] - float(torch.dot(Bm[2], Bm[5]))) < 1e-8 * n


def test_rmat_device_build_matches_oracle_and_block_parity(lz, ctx, orc):
    """config 4 shape at reduced scale: R-MAT graph Laplacian (power-law rows; the reference's ELL cannot
    hold it).  The device-built CSR equals the oracle's numpy construction exactly, SpMV agrees on the
    hub rows (long-row paths), and block Lanczos b = 32 matches the oracle."""
    _rmat_case(lz, ctx, orc, 12, 6, 40)


def test_rmat_scale18_parity(lz, ctx, orc):
    """config 4 at the largest scale the CPU oracle finishes quickly (2^18 rows, 8 M non-zeros, hub rows of
    tens of thousands of entries): device-built CSR identical to the oracle's, SpMV on the row-split schedule,
    block Lanczos b = 32 (6 blocks) and 40 reorthogonalised vector steps to 1e-10."""
    orc.set_threads(__import__("os").cpu_count() or 4)
    try:
        _rmat_case(lz, ctx, orc, 18, 6, 40)
    finally:
        orc.set_threads(1)


def _rmat_case(lz, ctx, orc, scale, m, mv):
    rp, ci, va = orc.rmat_laplacian(scale)
    A = lz.Matrix.rmat_laplacian(ctx, scale)
    got = A.csr_to_host()
    assert np.array_equal(got[0], rp) and np.array_equal(got[1], ci) and np.array_equal(got[2], va)
    n = 1 << scale
    lens = np.diff(rp)
    assert lens.max() > 500                       # hub rows exist
    x = orc.start_vector(n, 3)
    y = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.spmv(ctx, A, dev(x), y)
    ctx.sync()
    ref = orc.spmv((rp, ci, va), x)
    assert np.max(np.abs(y.cpu().numpy() - ref)) < 1e-11 * np.abs(ref).max()
    bw = 32
    B = orc.start_block(n, bw)
    o = orc.block_lanczos((rp, ci, va), B, m, lc=7)
    a, b, q = run_block(lz, ctx, A, B, m, 7)
    assert block_err(a, o["alpha"], m) < 1e-10 and block_err(b, o["beta"], m) < 1e-10
    # without reorthogonalisation two roundings of this recurrence part ways once the dominant Ritz value
    # has converged (~15 steps on this spectrum), so the 40-step comparison runs with full reorth
    al, be, steps = lz.vector_lanczos(ctx, A, dev(B[:, 0].copy()), mv, reorth=1)
    ov = orc.vector_lanczos((rp, ci, va), B[:, 0].copy(), mv, reorth=1)
    assert np.max(np.abs(al - ov["alpha"])) < 1e-10 * np.abs(ov["alpha"]).max()
    assert np.max(np.abs(be - ov["beta"]) / ov["beta"]) < 1e-10


@pytest.mark.parametrize("dims", [(40, 36, 24), (64, 24, 18), (96, 80)])
def test_box_chunk_staged_spmm_on_ragged_grids(lz, ctx, orc, dims):
    """Operand-staging SpMM (lz_spmm_xs.cuh) with box-shaped chunks on grids whose sides are NOT multiples of the box
    (clipped boxes, chunk padding, row map): the schedule must engage (kind 2), the block recurrence (fused subtraction +
    Gram epilogue), the reorthogonalised one (fused subtraction only) and the plain product (block fdtd) must agree with
    the oracle; with LZ_NO_XS the same coefficients come from the gathering kernel (tests/test_gpu_paths.py)."""
    bw, m = 16, 6
    if len(dims) == 3:
        A = lz.Matrix.laplacian3d(ctx, *dims); csr = orc.lap3d(*dims)
    else:
        A = lz.Matrix.laplacian2d(ctx, *dims); csr = orc.lap2d(*dims)
    n = int(np.prod(dims))
    kind, box, win = lz.spmm_schedule(ctx, A, bw)
    assert kind == 2 and box[0] >= 8 and win > 0, (kind, box, win)
    assert lz.spmm_schedule(ctx, A, 8)[0] == 0            # narrower panels keep the gathering kernel (measured slower)
    B = orc.start_block(n, bw)
    for reorth in (0, 1):
        ref = orc.block_lanczos(csr, B, m, lc=7, reorth=reorth)
        a, b, q = run_block(lz, ctx, A, B, m, 7, reorth=reorth)
        assert block_err(a, ref["alpha"], m) < 1e-10 and block_err(b, ref["beta"], m) < 1e-10, (dims, reorth)
        assert np.max(np.abs(q - ref["q"])) < 1e-10 * np.abs(ref["q"]).max()
    # plain product: 40 explicit Euler steps of U' = -0.05 A U against the oracle's loop
    rp, ci, va = csr
    As = lz.Matrix.from_csr(ctx, torch.from_numpy(rp).cuda(), torch.from_numpy(ci).cuda(), torch.from_numpy(-0.05 * va).cuda())
    assert lz.spmm_schedule(ctx, As, bw)[0] == 2
    U0 = orc.start_block(n, bw)
    want = orc.fdtd_block((rp, ci, -0.05 * va), U0, 40, 1.0)
    got = lz.fdtd_block(ctx, As, cm(U0), n, bw, 40, 1.0, lc=n // 3)
    assert np.max(np.abs(got - want[n // 3])) < 1e-12 * np.max(np.abs(want[n // 3]))
    As.close(); A.close()
