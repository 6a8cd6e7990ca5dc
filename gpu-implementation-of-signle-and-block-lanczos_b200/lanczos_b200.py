"""ctypes binding of liblanczos_b200.so (include/lanczos_b200.h) for tests, bench.py and smoke().

The product is the C-ABI library and the C++ mirror under host/; this module only hands torch
device pointers to it.  There is no CPU fallback: if the CUDA library has not been built the import
of `lib()` raises, and every compute entry point fails without a B200.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblanczos_b200.so")
_LIB = None

LZ_OK = 0
REORTH_NONE, REORTH_FULL, REORTH_FULL_DGKS, REORTH_SELECTIVE = 0, 1, 2, 3
H2D, D2H, D2D = 1, 2, 3


class LanczosError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("lanczos_b200 status %d: %s" % (status, msg))
        self.status = status


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("liblanczos_b200.so is missing (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32, u64, dbl, sz = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_double, C.c_size_t
    P = C.POINTER
    sig = {
        "lz_version": (i32, []),
        "lz_last_error": (C.c_char_p, []),
        "lz_ctx_create": (i32, [i32, vp, P(vp)]),
        "lz_ctx_destroy": (i32, [vp]),
        "lz_ctx_sync": (i32, [vp]),
        "lz_ctx_device": (i32, [vp]),
        "lz_ctx_launch_count": (i64, [vp]),
        "lz_ctx_profile": (i32, [vp, i32]),
        "lz_ctx_profile_read": (i32, [vp, P(i64), P(dbl), P(dbl)]),
        "lz_malloc": (i32, [vp, sz, P(vp)]),
        "lz_free": (i32, [vp, vp]),
        "lz_memcpy": (i32, [vp, vp, vp, sz, i32]),
        "lz_memset": (i32, [vp, vp, i32, sz]),
        "lz_fill": (i32, [vp, i64, dbl, vp]),
        "lz_csr_create": (i32, [vp, i64, i64, i64, vp, vp, vp, P(vp)]),
        "lz_csr_create_host": (i32, [vp, i64, i64, i64, vp, vp, vp, P(vp)]),
        "lz_ell_create": (i32, [vp, i64, i64, i32, i32, vp, vp, P(vp)]),
        "lz_matrix_destroy": (i32, [vp]),
        "lz_matrix_info": (i32, [vp, P(i64), P(i64), P(i64)]),
        "lz_matrix_spmm_schedule": (i32, [vp, vp, i32, P(i32), P(i32), P(i32)]),
        "lz_grid_strides_host": (i32, [i64, i64, vp, vp, P(i64)]),
        "lz_box_order_host": (i32, [i64, i32, P(i64), i32, i32, i32, vp, i64, vp, P(i64)]),
        "lz_matrix_csr_view": (i32, [vp, P(vp), P(vp), P(vp)]),
        "lz_gen_maxwell": (i32, [vp, i32, i32, i32, P(vp)]),
        "lz_matrix_ell_view": (i32, [vp, P(vp), P(vp)]),
        "lz_gen_laplacian2d": (i32, [vp, i64, i64, P(vp)]),
        "lz_gen_laplacian3d": (i32, [vp, i64, i64, i64, P(vp)]),
        "lz_gen_rmat_edges": (i32, [vp, i32, i64, u64, vp, vp]),
        "lz_gen_start_vector": (i32, [vp, i64, u64, vp]),
        "lz_gen_start_block": (i32, [vp, i64, i32, i64, u64, vp]),
        "lz_spmv": (i32, [vp, vp, vp, vp]),
        "lz_spmm": (i32, [vp, vp, i32, vp, i64, vp, i64]),
        "lz_dot": (i32, [vp, i64, vp, vp, P(dbl)]),
        "lz_nrm2": (i32, [vp, i64, vp, P(dbl)]),
        "lz_axpby": (i32, [vp, i64, dbl, vp, dbl, vp]),
        "lz_mm_tt": (i32, [vp, i64, i32, vp, i64, vp]),
        "lz_mm_tt2": (i32, [vp, i64, i32, vp, i64, vp, i64, vp]),
        "lz_mm_ts": (i32, [vp, i64, i32, dbl, dbl, vp, i64, vp, vp, i64]),
        "lz_sqrtm": (i32, [vp, i32, vp, vp]),
        "lz_copy_row": (i32, [vp, i64, i32, vp, i64, vp, i64]),
        "lz_assemble_T": (i32, [vp, i32, i32, vp, vp, vp]),
        "lz_vector_lanczos": (i32, [vp, vp, vp, i32, i64, i32, vp, vp, vp, P(i32)]),
        "lz_vector_lanczos_async": (i32, [vp, vp, vp, i32, i64, i32, vp, vp, vp]),
        "lz_vector_lanczos_workspace": (i32, [vp, vp, i32, i32]),
        "lz_block_lanczos_workspace": (i32, [vp, vp, i32, i32, i32]),
        "lz_vector_lanczos_begin": (i32, [vp, vp, vp, i32, i64, i32, vp]),
        "lz_vector_lanczos_advance": (i32, [vp, i32, vp, vp, P(i32)]),
        "lz_vector_checkpoint_save": (i32, [vp, C.c_char_p]),
        "lz_vector_checkpoint_load": (i32, [vp, vp, C.c_char_p, vp]),
        "lz_vector_ritz_vectors": (i32, [vp, i32, vp, vp, i64]),
        "lz_vector_reorth_count": (i32, [vp, P(i32)]),
        "lz_eigs_thick_restart": (i32, [vp, vp, vp, i32, i32, i32, dbl, i32, vp, vp, vp, i64, vp]),
        "lz_block_eigs_thick_restart": (i32, [vp, vp, vp, i64, i32, i32, i32, i32, dbl, i32, vp, vp, vp, i64, vp]),
        "lz_vector_basis_info": (i32, [vp, P(i64), P(i32)]),
        "lz_vector_basis_copy": (i32, [vp, i32, i32, vp, i64]),
        "lz_block_lanczos": (i32, [vp, vp, vp, i64, i32, i32, i64, i32, vp, vp, vp]),
        "lz_block_status": (i32, [vp, i32, P(i32)]),
        "lz_last_coupling": (i32, [vp, i32, vp]),
        "lz_comm_status": (i32, [vp, P(i32), P(i32)]),
        "lz_ritz": (i32, [i32, i32, vp, vp, vp, i32, vp, vp]),
        "lz_expm_sym": (i32, [i32, vp]),
        "lz_lanczos_solution": (i32, [i32, i32, vp, vp, vp, dbl, vp]),
        "lz_fdtd_vector": (i32, [vp, vp, vp, i64, dbl, i64, P(dbl), vp]),
        "lz_fdtd_block": (i32, [vp, vp, vp, i64, i32, i64, dbl, i64, vp]),
        "lz_comm_unique_id": (i32, [vp]),
        "lz_comm_init": (i32, [vp, i32, i32, vp]),
        "lz_comm_destroy": (i32, [vp]),
        "lz_partition_rows": (i32, [i64, i64, i32, i32, P(i64), P(i64)]),
        "lz_gen_laplacian3d_shard": (i32, [vp, i64, i64, i64, i32, i32, P(vp)]),
        "lz_gen_laplacian2d_shard": (i32, [vp, i64, i64, i32, i32, P(vp)]),
        "lz_vector_lanczos_sharded": (i32, [vp, vp, vp, i32, i32, vp, vp]),
        "lz_csr_create_shard_host": (i32, [vp, i64, i64, vp, vp, vp, i64, i64, i64, i64, i64, i64, P(vp)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    L._signatures = sig
    _LIB = L
    return L


def check(status):
    if status != LZ_OK:
        raise LanczosError(status, lib().lz_last_error().decode(errors="replace"))


def _ptr(t):
    """device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        return t.ctypes.data
    return t.data_ptr()


class Context:
    def __init__(self, device=0, stream=None):
        import weakref
        self.h = C.c_void_p()
        check(lib().lz_ctx_create(int(device), stream, C.byref(self.h)))
        self.device = device
        self._children = weakref.WeakSet()        # operators created on this context: closed before it

    def comm_status(self):
        """(peer_mode, timed_out) of the communicator attached with lz_comm_init."""
        pm, to = C.c_int(0), C.c_int(0)
        check(lib().lz_comm_status(self.h, C.byref(pm), C.byref(to)))
        return pm.value, to.value

    def sync(self):
        check(lib().lz_ctx_sync(self.h))

    @property
    def launches(self):
        return int(lib().lz_ctx_launch_count(self.h))

    def profile(self, enable=True):
        check(lib().lz_ctx_profile(self.h, 1 if enable else 0))

    def profile_read(self):
        """{class: (launches, ms, algorithmic_bytes)} accumulated since profiling was enabled."""
        n = 10
        la, ms, by = (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
        check(lib().lz_ctx_profile_read(self.h, la, ms, by))
        names = ["spmv", "pass_b", "cgs_project", "cgs_update", "spmm", "gram", "panel", "small", "comm", "cgs_update_project"]
        return {names[i]: (int(la[i]), float(ms[i]), float(by[i])) for i in range(n)}

    def close(self):
        if self.h:
            for child in list(self._children):     # children first; the C library also tolerates the other order
                child.close()
            lib().lz_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Matrix:
    """A sparse operator resident on the device (lz_matrix)."""

    def __init__(self, ctx, handle, keep=()):
        self.ctx, self.h, self._keep = ctx, handle, keep
        ctx._children.add(self)
        nr, nc, nnz = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().lz_matrix_info(self.h, C.byref(nr), C.byref(nc), C.byref(nnz)))
        self.n_rows, self.n_cols, self.nnz = nr.value, nc.value, nnz.value

    @classmethod
    def from_csr(cls, ctx, rowptr, colidx, vals, n_cols=None):
        """torch CUDA tensors (int32, int32, float64): borrowed."""
        n = rowptr.numel() - 1
        h = C.c_void_p()
        check(lib().lz_csr_create(ctx.h, n, n_cols or n, vals.numel(), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.byref(h)))
        return cls(ctx, h, (rowptr, colidx, vals))

    @classmethod
    def from_csr_host(cls, ctx, rowptr, colidx, vals, n_cols=None):
        """numpy arrays (or pinned torch CPU tensors): copied to the device by the library."""
        n = len(rowptr) - 1
        h = C.c_void_p()
        check(lib().lz_csr_create_host(ctx.h, n, n_cols or n, len(vals), _ptr(rowptr), _ptr(colidx), _ptr(vals), C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_ell(cls, ctx, n_rows, n_cols, width, layout, data, idx):
        """torch CUDA tensors: data float64, idx int32/uint32 storage (reference Ell_matrix arrays)."""
        h = C.c_void_p()
        check(lib().lz_ell_create(ctx.h, n_rows, n_cols, width, layout, _ptr(data), _ptr(idx), C.byref(h)))
        return cls(ctx, h, (data, idx))

    @classmethod
    def maxwell(cls, ctx, nx, ny=None, nz=None):
        """The reference's Matrix_A(Nx, Ny, Nz) (A = D W), assembled on the device."""
        h = C.c_void_p()
        check(lib().lz_gen_maxwell(ctx.h, nx, ny or nx, nz or nx, C.byref(h)))
        return cls(ctx, h)

    def ell_to_host(self):
        """(data, idx) numpy copies of a width-4 ELL operator, shape (n_rows, 4) (row-interleaved device format)."""
        d, i = C.c_void_p(), C.c_void_p()
        check(lib().lz_matrix_ell_view(self.h, C.byref(d), C.byref(i)))
        data, idx = np.empty((self.n_rows, 4), np.float64), np.empty((self.n_rows, 4), np.uint32)
        check(lib().lz_memcpy(self.ctx.h, data.ctypes.data, d, data.nbytes, D2H))
        check(lib().lz_memcpy(self.ctx.h, idx.ctypes.data, i, idx.nbytes, D2H))
        return data, idx

    @classmethod
    def laplacian2d(cls, ctx, nx, ny):
        h = C.c_void_p()
        check(lib().lz_gen_laplacian2d(ctx.h, nx, ny, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def laplacian3d(cls, ctx, nx, ny, nz):
        h = C.c_void_p()
        check(lib().lz_gen_laplacian3d(ctx.h, nx, ny, nz, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def rmat_laplacian(cls, ctx, scale, edge_factor=16, seed=0x5EED):
        """Symmetrised, de-duplicated, loop-free R-MAT graph Laplacian L = D - A (config 4), built on the
        device: edges from lz_gen_rmat_edges, sort/unique/CSR with torch (input generation, not the hot path)."""
        import torch
        n, ne = 1 << scale, (1 << scale) * edge_factor
        src = torch.empty(ne, dtype=torch.int32, device="cuda")
        dst = torch.empty(ne, dtype=torch.int32, device="cuda")
        check(lib().lz_gen_rmat_edges(ctx.h, scale, ne, seed, src.data_ptr(), dst.data_ptr()))
        ctx.sync()
        s, d = src.long(), dst.long()
        del src, dst
        keep = s != d
        s, d = s[keep], d[keep]
        key = torch.unique(torch.cat([s * n + d, d * n + s]))
        del s, d, keep
        diag = torch.arange(n, device="cuda", dtype=torch.long) * (n + 1)
        key = torch.sort(torch.cat([key, diag]))[0]                 # columns ascending inside each row, diagonal included
        rows = key // n
        cols = (key - rows * n).to(torch.int32)
        del key
        counts = torch.bincount(rows, minlength=n)
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        rowptr[1:] = torch.cumsum(counts, 0)
        vals = torch.full((cols.numel(),), -1.0, dtype=torch.float64, device="cuda")
        is_diag = cols.long() == rows
        vals[is_diag] = (counts - 1).to(torch.float64)              # degree (the diagonal entry itself is not an edge)
        del rows, is_diag, counts
        return cls.from_csr(ctx, rowptr.to(torch.int32), cols, vals)

    @classmethod
    def laplacian3d_shard(cls, ctx, nx, ny, nz, world_size, rank):
        h = C.c_void_p()
        check(lib().lz_gen_laplacian3d_shard(ctx.h, nx, ny, nz, world_size, rank, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def laplacian2d_shard(cls, ctx, nx, ny, world_size, rank):
        h = C.c_void_p()
        check(lib().lz_gen_laplacian2d_shard(ctx.h, nx, ny, world_size, rank, C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def from_csr_shard_host(cls, ctx, rowptr, colidx, vals, halo_lo, halo_hi, global_rows, row_begin, bnd_lo_rows=0, bnd_hi_rows=None):
        """Row slab from host arrays (numpy / pinned torch CPU tensors), local column index space."""
        n = len(rowptr) - 1
        h = C.c_void_p()
        check(lib().lz_csr_create_shard_host(ctx.h, n, len(vals), _ptr(rowptr), _ptr(colidx), _ptr(vals), halo_lo, halo_hi,
                                             global_rows, row_begin, bnd_lo_rows, n if bnd_hi_rows is None else bnd_hi_rows, C.byref(h)))
        return cls(ctx, h)

    def csr_to_host(self):
        """(rowptr, colidx, vals) numpy copies of the device CSR arrays."""
        rp, ci, va = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib().lz_matrix_csr_view(self.h, C.byref(rp), C.byref(ci), C.byref(va)))
        out = (np.empty(self.n_rows + 1, np.int32), np.empty(self.nnz, np.int32), np.empty(self.nnz, np.float64))
        for dst, src in zip(out, (rp, ci, va)):
            if dst.nbytes:
                check(lib().lz_memcpy(self.ctx.h, dst.ctypes.data, src, dst.nbytes, D2H))
        return out

    def close(self):
        if self.h:
            lib().lz_matrix_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------ thin operator wrappers

def spmv(ctx, A, x, y):
    check(lib().lz_spmv(ctx.h, A.h, _ptr(x), _ptr(y)))


def spmm(ctx, A, b, X, ldx, Y, ldy):
    check(lib().lz_spmm(ctx.h, A.h, b, _ptr(X), ldx, _ptr(Y), ldy))


def dot(ctx, x, y):
    r = C.c_double()
    check(lib().lz_dot(ctx.h, x.numel(), _ptr(x), _ptr(y), C.byref(r)))
    return r.value


def nrm2(ctx, x):
    r = C.c_double()
    check(lib().lz_nrm2(ctx.h, x.numel(), _ptr(x), C.byref(r)))
    return r.value


def axpby(ctx, a, y, b, x):
    check(lib().lz_axpby(ctx.h, y.numel(), a, _ptr(y), b, _ptr(x)))


def vector_lanczos(ctx, A, b, m, lc=0, reorth=REORTH_NONE, q=None):
    """Returns (alpha, beta) as numpy arrays (host, like test_lanczos.cu:66-67) and steps done."""
    alpha, beta = np.zeros(m), np.zeros(m)
    steps = C.c_int(0)
    st = lib().lz_vector_lanczos(ctx.h, A.h, _ptr(b), m, lc, reorth, alpha.ctypes.data, beta.ctypes.data, _ptr(q), C.byref(steps))
    if st != LZ_OK and st != -4:
        check(st)
    return alpha, beta, steps.value


def vector_lanczos_async(ctx, A, b, m, alpha_dev, beta_dev, lc=0, reorth=REORTH_NONE, q=None):
    check(lib().lz_vector_lanczos_async(ctx.h, A.h, _ptr(b), m, lc, reorth, _ptr(alpha_dev), _ptr(beta_dev), _ptr(q)))


def vector_lanczos_begin(ctx, A, b, m_capacity, lc=0, reorth=REORTH_NONE, q=None):
    check(lib().lz_vector_lanczos_begin(ctx.h, A.h, _ptr(b), m_capacity, lc, reorth, _ptr(q)))


def vector_lanczos_advance(ctx, steps, m_capacity):
    """`steps` more steps of the run begun on ctx; returns (alpha, beta, steps_done) with all coefficients so far."""
    alpha, beta = np.zeros(m_capacity), np.zeros(m_capacity)
    done = C.c_int(0)
    st = lib().lz_vector_lanczos_advance(ctx.h, steps, alpha.ctypes.data, beta.ctypes.data, C.byref(done))
    if st != LZ_OK and st != -4:
        check(st)
    return alpha[:done.value], beta[:done.value], done.value


def checkpoint_save(ctx, path):
    check(lib().lz_vector_checkpoint_save(ctx.h, str(path).encode()))


def checkpoint_load(ctx, A, path, q=None):
    check(lib().lz_vector_checkpoint_load(ctx.h, A.h, str(path).encode(), _ptr(q)))


def ritz_vectors(ctx, Y, X, ldx):
    """X[:, :k] = V_j Y for the run on ctx; Y: (j, k) numpy array, X: device buffer (column-major, ld = ldx)."""
    Yc = np.asfortranarray(Y, dtype=np.float64)
    check(lib().lz_vector_ritz_vectors(ctx.h, Yc.shape[1], Yc.ctypes.data, _ptr(X), ldx))


def reorth_count(ctx):
    c = C.c_int(0)
    check(lib().lz_vector_reorth_count(ctx.h, C.byref(c)))
    return c.value


def eigs_thick_restart(ctx, A, b, k, which=0, m_max=None, tol=1e-10, max_restarts=200, X=None, ldx=0):
    """k extremal eigenpairs by thick-restart Lanczos; returns (theta, resid_estimates, info dict)."""
    m_max = m_max or max(2 * k + 8, 32)
    theta, resid = np.zeros(k), np.zeros(k)
    info = (C.c_int * 4)()
    check(lib().lz_eigs_thick_restart(ctx.h, A.h, _ptr(b), k, which, m_max, float(tol), max_restarts, theta.ctypes.data,
                                      resid.ctypes.data, _ptr(X), ldx, info))
    return theta, resid, dict(converged=info[0], restarts=info[1], matvecs=info[2], basis=info[3])


def block_eigs_thick_restart(ctx, A, B, ldb, bw, k, which=0, p_blocks=None, tol=1e-10, max_restarts=200, X=None, ldx=0):
    """k extremal eigenpairs by block thick-restart Lanczos; returns (theta, resid_estimates, info dict)."""
    p_blocks = p_blocks or (2 * ((k + bw - 1) // bw) + 4)
    theta, resid = np.zeros(k), np.zeros(k)
    info = (C.c_int * 4)()
    check(lib().lz_block_eigs_thick_restart(ctx.h, A.h, _ptr(B), ldb, bw, k, which, p_blocks, float(tol), max_restarts,
                                            theta.ctypes.data, resid.ctypes.data, _ptr(X), ldx, info))
    return theta, resid, dict(converged=info[0], restarts=info[1], matvecs=info[2], basis=info[3])


def block_lanczos(ctx, A, B, ldb, bw, m, alpha, beta, q, lc=0, reorth=REORTH_NONE):
    check(lib().lz_block_lanczos(ctx.h, A.h, _ptr(B), ldb, bw, m, lc, reorth, _ptr(alpha), _ptr(beta), _ptr(q)))


def spmm_schedule(ctx, A, bw):
    """(kind, box, window_rows) of the panel-product kernel the block drivers run on A at width bw:
    kind 0 gathering kernel, 1 operand-staging with runs of rows, 2 operand-staging with box-shaped chunks."""
    kind, win = C.c_int(0), C.c_int(0)
    box = (C.c_int * 3)()
    check(lib().lz_matrix_spmm_schedule(ctx.h, A.h, bw, C.byref(kind), box, C.byref(win)))
    return kind.value, tuple(box), win.value


def block_status(ctx, m):
    """blocks of the last block run that are valid (m unless a beta_j was singular / non-finite)."""
    done = C.c_int(0)
    st = lib().lz_block_status(ctx.h, m, C.byref(done))
    if st != LZ_OK and st != -4:
        check(st)
    return done.value


def last_coupling(ctx, bw=1):
    """beta_m of the last run on this context, (bw, bw) numpy array (row, col)."""
    out = np.zeros(bw * bw)
    check(lib().lz_last_coupling(ctx.h, bw, out.ctypes.data))
    return out.reshape(bw, bw).T.copy()


def ritz(alpha, beta, k, bw=1, beta_last=None):
    alpha = np.ascontiguousarray(alpha, np.float64)
    beta = np.ascontiguousarray(beta, np.float64)
    m = alpha.size // (bw * bw)
    theta, resid = np.zeros(k), np.zeros(k)
    bl = None if beta_last is None else np.ascontiguousarray(beta_last, np.float64)
    check(lib().lz_ritz(m, bw, alpha.ctypes.data, beta.ctypes.data, None if bl is None else bl.ctypes.data, k,
                        theta.ctypes.data, resid.ctypes.data))
    return theta, resid


def expm_sym(T):
    """expm of a small symmetric matrix (host; V exp(Lambda) V^T as expm_cusolver does)."""
    T = np.array(T, dtype=np.float64, order="F", copy=True)
    check(lib().lz_expm_sym(T.shape[0], T.ctypes.data))
    return T


def lanczos_solution(alpha, beta, q, t_end=1.0, bw=1):
    """q^T expm(t_end T)[:, :bw] beta_0: what the reference harness prints as the Lanczos solution."""
    alpha = np.ascontiguousarray(alpha, np.float64); beta = np.ascontiguousarray(beta, np.float64)
    q = np.ascontiguousarray(q, np.float64)
    m = alpha.size // (bw * bw)
    out = np.zeros(bw)
    check(lib().lz_lanczos_solution(m, bw, alpha.ctypes.data, beta.ctypes.data, q.ctypes.data, float(t_end), out.ctypes.data))
    return out if bw > 1 else float(out[0])


def fdtd_vector(ctx, A, u0, nsteps, t_end=1.0, lc=0, u_out=None):
    res = C.c_double(0.0)
    check(lib().lz_fdtd_vector(ctx.h, A.h, _ptr(u0), nsteps, float(t_end), lc, C.byref(res), _ptr(u_out)))
    return res.value


def fdtd_block(ctx, A, U0, ldu, bw, nsteps, t_end=1.0, lc=0):
    out = np.zeros(bw)
    check(lib().lz_fdtd_block(ctx.h, A.h, _ptr(U0), ldu, bw, nsteps, float(t_end), lc, out.ctypes.data))
    return out
