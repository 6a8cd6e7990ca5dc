"""Object lifetime across the C-ABI (round-1 verdict: smoke() segfaulted because lz_matrix_destroy
dereferenced a context that lz_ctx_destroy had already deleted).

CPU part: Context.close() closes the operators created on it first.  GPU part: the C library itself
tolerates the wrong order -- a context destroyed first orphans its operators, which can still be
destroyed (and only destroyed) afterwards; and the per-device shared-memory opt-in works for a second
context in the same process."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def test_context_close_closes_children_first(lz, monkeypatch):
    calls = []

    class FakeLib:
        def lz_ctx_destroy(self, h):
            calls.append("ctx")
            return 0

        def lz_matrix_destroy(self, h):
            calls.append("matrix")
            return 0

    ctx = lz.Context.__new__(lz.Context)
    import weakref
    ctx.h, ctx.device, ctx._children = C.c_void_p(1), 0, weakref.WeakSet()
    m = lz.Matrix.__new__(lz.Matrix)
    m.ctx, m.h, m._keep = ctx, C.c_void_p(2), ()
    ctx._children.add(m)
    monkeypatch.setattr(lz, "lib", lambda: FakeLib())
    ctx.close()
    assert calls == ["matrix", "ctx"]
    m.close()                      # already closed: no second destroy
    ctx.close()
    assert calls == ["matrix", "ctx"]


@pytest.mark.gpu
def test_ctx_destroyed_before_its_matrices(lz):
    """Raw C-ABI calls in the order that crashed round 1: lz_ctx_destroy, then lz_matrix_destroy."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    L = lz.lib()
    torch.zeros(1, device="cuda")
    h = C.c_void_p()
    assert L.lz_ctx_create(0, None, C.byref(h)) == 0
    mats = []
    for nx in (8, 16, 24):
        a = C.c_void_p()
        assert L.lz_gen_laplacian2d(h, nx, nx, C.byref(a)) == 0
        mats.append(a)
    # destroy one in the right order (exercises the unlink of a middle list element), then the context
    assert L.lz_matrix_destroy(mats[1]) == 0
    assert L.lz_ctx_destroy(h) == 0
    # orphans: info still works, compute entry points refuse, destroy is safe
    nr = C.c_int64()
    assert L.lz_matrix_info(mats[0], C.byref(nr), None, None) == 0 and nr.value == 64
    h2 = C.c_void_p()
    assert L.lz_ctx_create(0, None, C.byref(h2)) == 0
    x = torch.ones(64, dtype=torch.float64, device="cuda")
    y = torch.zeros(64, dtype=torch.float64, device="cuda")
    assert L.lz_spmv(h2, mats[0], x.data_ptr(), y.data_ptr()) != 0          # operator of a dead context
    assert L.lz_matrix_destroy(mats[0]) == 0
    assert L.lz_matrix_destroy(mats[2]) == 0
    assert L.lz_ctx_destroy(h2) == 0


@pytest.mark.gpu
def test_two_contexts_in_one_process(lz, orc):
    """Two live contexts (second device when the box has one, else the same device): the kernels that need
    more than 48 KB of dynamic shared memory must work on both (the opt-in is tracked per context)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    devs = [0, 1] if torch.cuda.device_count() > 1 else [0, 0]
    nx, ny, m = 48, 40, 30
    ref = orc.vector_lanczos(orc.lap2d(nx, ny), orc.start_vector(nx * ny), m, reorth=1)
    for d in devs:
        with torch.cuda.device(d):
            torch.zeros(1, device="cuda")
            ctx = lz.Context(d)
            A = lz.Matrix.laplacian2d(ctx, nx, ny)
            b = torch.from_numpy(orc.start_vector(nx * ny)).cuda()
            alpha, beta, steps = lz.vector_lanczos(ctx, A, b, m, reorth=lz.REORTH_FULL)
            assert steps == m
            assert np.max(np.abs(alpha - ref["alpha"])) < 1e-10 * np.abs(ref["alpha"]).max()
            ctx.close()            # closes A first
            assert not A.h
