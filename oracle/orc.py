"""ctypes front-end of oracle/liboracle.so (the CPU restatement in lanczos_oracle.c).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never from the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def build():
    """(Re)build liboracle.so and, when the reference tree is present, oracle/_ref."""
    subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    i64, i32, u64, dbl = C.c_int64, C.c_int, C.c_uint64, C.c_double
    vp = C.c_void_p
    L.orc_splitmix64.restype = u64
    L.orc_splitmix64.argtypes = [u64]
    L.orc_start_vector.argtypes = [i64, u64, _f64p]
    L.orc_start_block.argtypes = [i64, i32, i64, u64, _f64p]
    L.orc_lap2d_nnz.restype = i64
    L.orc_lap2d_nnz.argtypes = [i64, i64]
    L.orc_lap2d_csr.argtypes = [i64, i64, _i32p, _i32p, _f64p]
    L.orc_lap3d_nnz.restype = i64
    L.orc_lap3d_nnz.argtypes = [i64, i64, i64]
    L.orc_lap3d_csr.argtypes = [i64, i64, i64, _i32p, _i32p, _f64p]
    L.orc_rmat_edges.argtypes = [i32, i64, u64, _i32p, _i32p]
    L.orc_ell_to_csr.restype = i64
    L.orc_ell_to_csr.argtypes = [i64, i32, _f64p, _u32p, _i32p, _i32p, _f64p]
    L.orc_csr_spmv.argtypes = [i64, _i32p, _i32p, _f64p, _f64p, _f64p]
    L.orc_csr_spmm.argtypes = [i64, _i32p, _i32p, _f64p, i32, _f64p, i64, _f64p, i64]
    L.orc_dot.restype = dbl
    L.orc_dot.argtypes = [i64, _f64p, _f64p]
    L.orc_mm_tt.argtypes = [i64, i32, _f64p, i64, _f64p]
    L.orc_mm_tt2.argtypes = [i64, i32, _f64p, i64, _f64p, i64, _f64p]
    L.orc_mm_ts.argtypes = [i64, i32, dbl, dbl, _f64p, i64, _f64p, _f64p, i64]
    L.orc_jacobi_eig.restype = i32
    L.orc_jacobi_eig.argtypes = [i32, _f64p, _f64p, _f64p]
    L.orc_sqrtm.argtypes = [i32, _f64p, _f64p]
    L.orc_vector_lanczos.restype = i32
    L.orc_vector_lanczos.argtypes = [i64, _i32p, _i32p, _f64p, _f64p, i32, i64, i32, _f64p, _f64p, _f64p, vp]
    L.orc_block_lanczos.restype = i32
    L.orc_block_lanczos.argtypes = [i64, _i32p, _i32p, _f64p, _f64p, i32, i32, i64, i32, _f64p, _f64p, _f64p, vp]
    L.orc_assemble_T.argtypes = [i32, i32, _f64p, _f64p, _f64p]
    L.orc_expm_sym.argtypes = [i32, _f64p]
    L.orc_lanczos_solution.argtypes = [i32, i32, _f64p, _f64p, _f64p, dbl, _f64p]
    L.orc_fdtd_vector.argtypes = [i64, _i32p, _i32p, _f64p, _f64p, i64, dbl]
    L.orc_fdtd_block.argtypes = [i64, _i32p, _i32p, _f64p, i32, _f64p, i64, i64, dbl]
    L.orc_set_threads.argtypes = [i32]
    L.orc_get_threads.restype = i32
    _LIB = L
    return L


# ------------------------------------------------------------------------------- generators

def set_threads(t):
    lib().orc_set_threads(int(t))


def start_vector(n, seed=0x5EED):
    v = np.empty(n, dtype=np.float64)
    lib().orc_start_vector(n, seed, v)
    return v


def start_block(n, b, seed=0x5EED):
    """column-major n x b (returned as an (n, b) Fortran-ordered array)."""
    buf = np.empty(n * b, dtype=np.float64)
    lib().orc_start_block(n, b, n, seed, buf)
    return buf.reshape(b, n).T


def lap2d(nx, ny):
    n, nnz = nx * ny, lib().orc_lap2d_nnz(nx, ny)
    rp, ci, va = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    lib().orc_lap2d_csr(nx, ny, rp, ci, va)
    return rp, ci, va


def lap3d(nx, ny, nz):
    n, nnz = nx * ny * nz, lib().orc_lap3d_nnz(nx, ny, nz)
    rp, ci, va = np.empty(n + 1, np.int32), np.empty(nnz, np.int32), np.empty(nnz, np.float64)
    lib().orc_lap3d_csr(nx, ny, nz, rp, ci, va)
    return rp, ci, va


def rmat_edges(scale, n_edges, seed=0x5EED):
    s, d = np.empty(n_edges, np.int32), np.empty(n_edges, np.int32)
    lib().orc_rmat_edges(scale, n_edges, seed, s, d)
    return s, d


def rmat_laplacian(scale, edge_factor=16, seed=0x5EED):
    """Symmetrised, de-duplicated, loop-free R-MAT graph Laplacian L = D - A as CSR (SURVEY 8d cfg 4)."""
    n = 1 << scale
    s, d = rmat_edges(scale, n * edge_factor, seed)
    keep = s != d
    s, d = s[keep].astype(np.int64), d[keep].astype(np.int64)
    key = np.unique(np.concatenate([s * n + d, d * n + s]))
    r, c = key // n, key % n
    deg = np.bincount(r, minlength=n)
    # off-diagonal -1 plus one diagonal entry per row, columns ascending
    rows = np.concatenate([r, np.arange(n)])
    cols = np.concatenate([c, np.arange(n)])
    vals = np.concatenate([-np.ones(len(r)), deg.astype(np.float64)])
    order = np.lexsort((cols, rows))
    rows, cols, vals = rows[order], cols[order], vals[order]
    rp = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(rows, minlength=n), out=rp[1:])
    return rp.astype(np.int32), cols.astype(np.int32), np.ascontiguousarray(vals)


def ell_to_csr(n, width, ell_data, ell_idx):
    rp = np.empty(n + 1, np.int32)
    ci = np.empty(n * width, np.int32)
    va = np.empty(n * width, np.float64)
    nnz = lib().orc_ell_to_csr(n, width, np.ascontiguousarray(ell_data, np.float64),
                               np.ascontiguousarray(ell_idx, np.uint32), rp, ci, va)
    return rp, ci[:nnz].copy(), va[:nnz].copy()


# ------------------------------------------------------------------------------- operators

def spmv(csr, x):
    rp, ci, va = csr
    y = np.empty(len(rp) - 1, np.float64)
    lib().orc_csr_spmv(len(rp) - 1, rp, ci, va, np.ascontiguousarray(x, np.float64), y)
    return y


def spmm(csr, X):
    """X: (n, b) array; returns (n, b) Fortran-ordered."""
    rp, ci, va = csr
    n, b = X.shape
    Xf = np.ascontiguousarray(X.T, np.float64).reshape(-1)     # column-major storage
    Y = np.empty(n * b, np.float64)
    lib().orc_csr_spmm(n, rp, ci, va, b, Xf, n, Y, n)
    return Y.reshape(b, n).T


def sqrtm(S):
    b = S.shape[0]
    s = np.ascontiguousarray(S.T, np.float64).reshape(-1).copy()
    si = np.empty_like(s)
    lib().orc_sqrtm(b, s, si)
    return s.reshape(b, b).T, si.reshape(b, b).T


def vector_lanczos(csr, b, m, lc=0, reorth=0, want_basis=False):
    rp, ci, va = csr
    n = len(rp) - 1
    alpha, beta, q = np.zeros(m), np.zeros(m), np.zeros(m)
    V = np.empty(n * m, np.float64) if want_basis else None
    done = lib().orc_vector_lanczos(n, rp, ci, va, np.ascontiguousarray(b, np.float64), m, lc, reorth,
                                    alpha, beta, q, V.ctypes.data if V is not None else None)
    out = dict(alpha=alpha, beta=beta, q=q, steps=done)
    if V is not None:
        out["V"] = V.reshape(m, n).T
    return out


def block_lanczos(csr, B, m, lc=0, reorth=0, want_basis=False):
    """B: (n, bw).  alpha: (m, bw, bw) with alpha[j][r, c]; beta: (m+1, bw, bw)."""
    rp, ci, va = csr
    n, bw = B.shape
    Bf = np.ascontiguousarray(B.T, np.float64).reshape(-1)
    alpha = np.zeros(m * bw * bw)
    beta = np.zeros((m + 1) * bw * bw)
    q = np.zeros(m * bw)
    V = np.empty(n * m * bw, np.float64) if want_basis else None
    done = lib().orc_block_lanczos(n, rp, ci, va, Bf, bw, m, lc, reorth, alpha, beta, q,
                                   V.ctypes.data if V is not None else None)
    out = dict(alpha=alpha.reshape(m, bw, bw).transpose(0, 2, 1),
               beta=beta.reshape(m + 1, bw, bw).transpose(0, 2, 1), q=q, steps=done)
    if V is not None:
        out["V"] = V.reshape(m * bw, n).T
    return out


def assemble_T(alpha, beta):
    """alpha (m,bw,bw), beta (>=m,bw,bw) as returned by block_lanczos (or 1-D for bw = 1)."""
    alpha, beta = np.asarray(alpha, np.float64), np.asarray(beta, np.float64)
    if alpha.ndim == 1:
        alpha, beta = alpha.reshape(-1, 1, 1), beta.reshape(-1, 1, 1)
    m, bw = alpha.shape[0], alpha.shape[1]
    a = np.ascontiguousarray(alpha.transpose(0, 2, 1)).reshape(-1)
    b = np.ascontiguousarray(beta[:m].transpose(0, 2, 1)).reshape(-1)
    T = np.empty((m * bw) ** 2, np.float64)
    lib().orc_assemble_T(m, bw, a, b, T)
    return T.reshape(m * bw, m * bw).T


def ritz(alpha, beta, k, beta_last=None):
    """k extremal Ritz values (k//2 smallest, k - k//2 largest) of eig(T) and residual estimates
    |beta_m * y_last| (scalar case) / ||beta_m Y_lastblock|| (block case) when beta_last is given."""
    T = assemble_T(alpha, beta)
    w, Y = np.linalg.eigh(T)
    lo = k // 2
    sel = np.r_[np.arange(lo), np.arange(len(w) - (k - lo), len(w))]
    res = None
    if beta_last is not None:
        bl = np.atleast_2d(np.asarray(beta_last, np.float64))
        bw = bl.shape[0]
        res = np.linalg.norm(bl @ Y[-bw:, sel], axis=0)
    return w[sel], res


def thick_restart_lanczos(csr, b, k, which=0, m_max=None, tol=1e-10, max_restarts=200):
    """numpy restatement of the thick-restart Lanczos in csrc/lz_eigs.cu (Wu & Simon's TRLan with full CGS2
    reorthogonalisation), same restart rule: keep the k wanted Ritz pairs plus max(1, 2(m-k)/5) of their
    neighbours, compress V <- V Y, continue from the residual vector.  which: 0 smallest, 1 largest, 2 both ends.
    TEST INFRASTRUCTURE ONLY.  Returns (theta ascending, residual estimates, X, info)."""
    n = len(csr[0]) - 1
    m = m_max or max(2 * k + 8, 32)
    V = np.zeros((n, m + 1))
    H = np.zeros((m, m))
    beta0 = np.linalg.norm(b)
    V[:, 0] = b / beta0
    j0, restarts, matvecs = 0, 0, 0
    beta_prev = 0.0
    while True:
        for j in range(j0, m):
            w = spmv(csr, V[:, j].copy())
            if j > j0:
                w -= beta_prev * V[:, j - 1]
            c1 = V[:, :j + 1].T @ w
            w -= V[:, :j + 1] @ c1
            c2 = V[:, :j + 1].T @ w
            w -= V[:, :j + 1] @ c2
            H[j, j] = c1[j]
            if j > j0:
                H[j - 1, j] = H[j, j - 1] = beta_prev
            beta_prev = np.linalg.norm(w)
            V[:, j + 1] = w / beta_prev
        matvecs += m - j0
        beta_m = beta_prev
        d, Z = np.linalg.eigh(H)
        lo = k if which == 0 else 0 if which == 1 else k // 2
        hi = k - lo
        wanted = list(range(lo)) + list(range(m - hi, m))
        res = np.abs(beta_m * Z[m - 1, :])
        nconv = int(np.sum(res[wanted] <= tol * np.max(np.abs(d))))
        if nconv == k or restarts == max_restarts:
            break
        extra = min(max(1, (m - k) * 2 // 5), m - k - 3)
        elo = extra if which == 0 else 0 if which == 1 else extra // 2
        ehi = extra - elo
        keep = list(range(lo + elo)) + list(range(m - (hi + ehi), m))
        kk = len(keep)
        V[:, :kk] = V[:, :m] @ Z[:, keep]
        V[:, kk] = V[:, m]
        H[:] = 0.0
        for c, i in enumerate(keep):
            H[c, c] = d[i]
            H[c, kk] = H[kk, c] = beta_m * Z[m - 1, i]
        j0 = kk
        restarts += 1
    X = V[:, :m] @ Z[:, wanted]
    return d[wanted], res[wanted], X, dict(converged=nconv, restarts=restarts, matvecs=matvecs, basis=m)


def block_thick_restart_lanczos(csr, B, k, which=0, p_blocks=None, tol=1e-10, max_restarts=200):
    """numpy restatement of lz_block_eigs_thick_restart (csrc/lz_block.cu): block recurrence with full block CGS2,
    symmetric square roots for the coupling blocks, the same restart rule (wanted pairs plus max(b, 2(N-k)/5) neighbours,
    rounded up to whole blocks).  TEST INFRASTRUCTURE ONLY.  Returns (theta ascending, residual estimates, X, info)."""
    n, b = B.shape
    p = p_blocks or (2 * ((k + b - 1) // b) + 4)
    N = p * b

    def sym_sqrt(G):
        w, U = np.linalg.eigh(0.5 * (G + G.T))
        w = np.abs(w)
        return (U * np.sqrt(w)) @ U.T, (U / np.sqrt(w)) @ U.T

    V = np.zeros((n, N + b))
    H = np.zeros((N, N))
    _, binv = sym_sqrt(B.T @ B)
    V[:, :b] = B @ binv
    j0, restarts, matvecs = 0, 0, 0
    while True:
        for j in range(j0, p):
            Q = V[:, j * b:(j + 1) * b]
            W = spmm(csr, np.ascontiguousarray(Q))
            G = Q.T @ W
            alpha = 0.5 * (G + G.T)
            W = W - Q @ alpha
            for _ in range(2):
                Vj = V[:, :(j + 1) * b]
                W = W - Vj @ (Vj.T @ W)
            beta, binv = sym_sqrt(W.T @ W)
            H[j * b:(j + 1) * b, j * b:(j + 1) * b] = alpha
            if j + 1 < p:
                H[(j + 1) * b:(j + 2) * b, j * b:(j + 1) * b] = beta
                H[j * b:(j + 1) * b, (j + 1) * b:(j + 2) * b] = beta.T
            V[:, (j + 1) * b:(j + 2) * b] = W @ binv
        matvecs += (p - j0) * b
        beta_p = beta
        d, Z = np.linalg.eigh(H)
        lo = k if which == 0 else 0 if which == 1 else k // 2
        hi = k - lo
        wanted = list(range(lo)) + list(range(N - hi, N))
        res = np.linalg.norm(beta_p @ Z[N - b:, :], axis=0)
        nconv = int(np.sum(res[wanted] <= tol * np.max(np.abs(d))))
        if nconv == k or restarts == max_restarts:
            break
        kk = k + max(b, (N - k) * 2 // 5)
        kk = min(((kk + b - 1) // b) * b, (p - 2) * b)
        extra = kk - k
        elo = extra if which == 0 else 0 if which == 1 else extra // 2
        ehi = extra - elo
        keep = list(range(lo + elo)) + list(range(N - (hi + ehi), N))
        V[:, :kk] = V[:, :N] @ Z[:, keep]
        V[:, kk:kk + b] = V[:, N:N + b]
        H[:] = 0.0
        H[np.arange(kk), np.arange(kk)] = d[keep]
        S = beta_p @ Z[N - b:, keep]
        H[kk:kk + b, :kk] = S
        H[:kk, kk:kk + b] = S.T
        j0 = kk // b
        restarts += 1
    X = V[:, :N] @ Z[:, wanted]
    return d[wanted], res[wanted], X, dict(converged=nconv, restarts=restarts, matvecs=matvecs, basis=N)


def _mx_axis(N):
    Np = N + 2
    h = (1.0 - 0.0) / (Np - 1)
    p = np.array([0.0 + i * h for i in range(Np)])
    h2 = ((1.0 - h) - 0.0) / (Np - 2)
    d = np.array([0.0 + i * h2 for i in range(Np - 1)]) + h / 2
    dp, dd = p[1:] - p[:-1], d[1:] - d[:-1]
    F = dict(rows=N + 1, cols=N, w=2, val=np.zeros((N + 1, 2)), col=np.zeros((N + 1, 2), np.int64))
    for r in range(N + 1):
        inv = 1.0 / dp[r]; s = 0
        if r >= 1: F["val"][r, s] = inv * -1.0; F["col"][r, s] = r - 1; s += 1
        if r < N: F["val"][r, s] = inv * 1.0; F["col"][r, s] = r
    B = dict(rows=N, cols=N + 1, w=2, val=np.zeros((N, 2)), col=np.zeros((N, 2), np.int64))
    for r in range(N):
        inv = 1.0 / dd[r]
        B["val"][r, 0] = -(inv * 1.0); B["col"][r, 0] = r
        B["val"][r, 1] = -(inv * -1.0); B["col"][r, 1] = r + 1
    ident = lambda n: dict(rows=n, cols=n, w=1, val=np.ones((n, 1)), col=np.arange(n).reshape(n, 1))
    diagv = lambda v: dict(rows=len(v), cols=len(v), w=1, val=v.reshape(-1, 1).copy(), col=np.arange(len(v)).reshape(-1, 1))
    return dict(F=F, B=B, I=ident(N), Ip=ident(N + 1), W=diagv(dp), Wh=diagv(dd))

def _mx_k3(a, b, c, sign=1.0):
    rows = a["rows"] * b["rows"] * c["rows"]; w = a["w"] * b["w"] * c["w"]
    R = np.arange(rows)
    rz, rem = R // (b["rows"] * c["rows"]), R % (b["rows"] * c["rows"])
    ry, rx = rem // c["rows"], rem % c["rows"]
    val = np.zeros((rows, w)); col = np.zeros((rows, w), np.int64)
    for s in range(w):
        ka, kb, kc = s // (b["w"] * c["w"]), (s // c["w"]) % b["w"], s % c["w"]
        val[:, s] = a["val"][rz, ka] * (b["val"][ry, kb] * c["val"][rx, kc])
        col[:, s] = a["col"][rz, ka] * (b["cols"] * c["cols"]) + (b["col"][ry, kb] * c["cols"] + c["col"][rx, kc])
    if sign != 1.0: val = val * sign
    return dict(rows=rows, cols=a["cols"] * b["cols"] * c["cols"], w=w, val=val, col=col)

def maxwell_closed_form(Nx, Ny=None, Nz=None):
    """Closed form of the reference's Matrix_A(Nx, Ny, Nz) followed by mult_diagonal (matrix_a/build_A_ell.hpp:8-255,
    objects/ell_matrix.hpp:340-361): numpy restatement of what csrc/lz_maxwell.cu evaluates per row.  Returns
    (D values (n,4), column ids (n,4), W diagonal (n,), A = D W values (n,4)).  TEST INFRASTRUCTURE ONLY."""
    x, y, z = _mx_axis(Nx), _mx_axis(Ny or Nx), _mx_axis(Nz or Nx)
    De12, De13 = _mx_k3(z["F"], y["Ip"], x["I"], -1.0), _mx_k3(z["Ip"], y["F"], x["I"])
    De21, De23 = _mx_k3(z["F"], y["I"], x["Ip"]), _mx_k3(z["Ip"], y["I"], x["F"], -1.0)
    De31, De32 = _mx_k3(z["I"], y["F"], x["Ip"], -1.0), _mx_k3(z["I"], y["Ip"], x["F"])
    Dh12, Dh13 = _mx_k3(z["B"], y["I"], x["Ip"]), _mx_k3(z["I"], y["B"], x["Ip"], -1.0)
    Dh21, Dh23 = _mx_k3(z["B"], y["Ip"], x["I"], -1.0), _mx_k3(z["I"], y["Ip"], x["B"])
    Dh31, Dh32 = _mx_k3(z["Ip"], y["B"], x["I"]), _mx_k3(z["Ip"], y["I"], x["B"], -1.0)
    def curl(b12, b13, b21, b23, b31, b32, c1, c2):
        rows = b12["rows"] + b21["rows"] + b31["rows"]
        val = np.zeros((rows, 4)); col = np.zeros((rows, 4), np.int64)
        r2, r3 = b12["rows"], b12["rows"] + b21["rows"]
        def ins(blk, r0, s0, shift):
            val[r0:r0 + blk["rows"], s0:s0 + 2] = blk["val"]; col[r0:r0 + blk["rows"], s0:s0 + 2] = blk["col"] + shift
        ins(b12, 0, 0, c1); ins(b13, 0, 2, c1 + c2)
        ins(b21, r2, 0, 0); ins(b23, r2, 2, c1 + c2)
        ins(b31, r3, 0, 0); ins(b32, r3, 2, c1)
        return val, col
    De_rows = De12["rows"] + De21["rows"] + De31["rows"]; Dh_rows = Dh12["rows"] + Dh21["rows"] + Dh31["rows"]
    Dev, Dec = curl(De12, De13, De21, De23, De31, De32, Dh12["rows"], Dh21["rows"])
    Dhv, Dhc = curl(Dh12, Dh13, Dh21, Dh23, Dh31, Dh32, De12["rows"], De21["rows"])
    Dv = np.vstack([Dhv, Dev]); Dc = np.vstack([Dhc + Dh_rows, Dec])
    Wb = [_mx_k3(z["Wh"], y["Wh"], x["W"]), _mx_k3(z["Wh"], y["W"], x["Wh"]), _mx_k3(z["W"], y["Wh"], x["Wh"]),
          _mx_k3(z["W"], y["W"], x["Wh"], -1.0), _mx_k3(z["W"], y["Wh"], x["W"], -1.0), _mx_k3(z["Wh"], y["W"], x["W"], -1.0)]
    Wv = np.concatenate([b["val"][:, 0] for b in Wb])
    Av = Dv * Wv[Dc]
    return Dv, Dc, Wv, Av

def expm_sym(T):
    T = np.array(T, np.float64)
    n = T.shape[0]
    t = np.ascontiguousarray(T.T).reshape(-1).copy()          # column-major
    lib().orc_expm_sym(n, t)
    return t.reshape(n, n).T


def lanczos_solution(alpha, beta, q, t_end=1.0):
    """alpha (m,bw,bw) / beta (>=m,bw,bw) or 1-D series; returns the bw-vector (scalar for bw = 1)."""
    alpha, beta = np.asarray(alpha, np.float64), np.asarray(beta, np.float64)
    if alpha.ndim == 1:
        alpha, beta = alpha.reshape(-1, 1, 1), beta.reshape(-1, 1, 1)
    m, bw = alpha.shape[0], alpha.shape[1]
    a = np.ascontiguousarray(alpha.transpose(0, 2, 1)).reshape(-1)
    b = np.ascontiguousarray(beta[:m].transpose(0, 2, 1)).reshape(-1)
    out = np.zeros(bw)
    lib().orc_lanczos_solution(m, bw, a, b, np.ascontiguousarray(q, np.float64), float(t_end), out)
    return out if bw > 1 else float(out[0])


def fdtd_vector(csr, u0, nsteps, t_end=1.0):
    rp, ci, va = csr
    u = np.array(u0, np.float64).copy()
    lib().orc_fdtd_vector(len(rp) - 1, rp, ci, va, u, nsteps, float(t_end))
    return u


def fdtd_block(csr, U0, nsteps, t_end=1.0):
    """U0: (n, b) array; returns (n, b)."""
    rp, ci, va = csr
    n, b = U0.shape
    U = np.ascontiguousarray(np.asarray(U0, np.float64).T).reshape(-1).copy()
    lib().orc_fdtd_block(n, rp, ci, va, b, U, n, nsteps, float(t_end))
    return U.reshape(b, n).T


# ------------------------------------------------------------------------------- _ref dumps

def read_dump(path):
    """Reader for the record container written by oracle/_ref/ref_host_dump_*."""
    out = {}
    with open(path, "rb") as f:
        raw = f.read()
    off = 0
    dts = {0: np.float64, 1: np.uint32, 2: np.int64}
    while off < len(raw):
        ln = int(np.frombuffer(raw, np.uint32, 1, off)[0]); off += 4
        name = raw[off:off + ln].decode(); off += ln
        dt = dts[raw[off]]; off += 1
        cnt = int(np.frombuffer(raw, np.uint64, 1, off)[0]); off += 8
        arr = np.frombuffer(raw, dt, cnt, off).copy(); off += cnt * arr.itemsize
        out[name] = arr if not (dt is np.int64 and cnt == 1) else int(arr[0])
    return out


def ref_dump_path(n_col=4):
    return os.path.join(_HERE, "_ref", "ref_host_dump_%d" % n_col)


def run_ref(mode, N, m, n_col=4, out=None):
    """Run the reference-derived dump tool (needs oracle/_ref built from /root/reference)."""
    import tempfile
    exe = ref_dump_path(n_col)
    if out is None:
        out = os.path.join(tempfile.mkdtemp(), "dump.bin")
    subprocess.run([exe, mode, str(N), str(m), out], check=True)
    return read_dump(out)
