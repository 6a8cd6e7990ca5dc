"""CPU suite: host logic of the operand-staging SpMM's box-shaped chunks (include/lanczos_b200.h: lz_grid_strides_host,
lz_box_order_host) -- stride detection from a sample of CSR rows and the box ordering of the grid's rows.  No device is
touched; the device side (windows, padded copy, the kernel) is covered by tests/test_gpu_block.py."""
import ctypes as C

import numpy as np
import pytest


def strides_of(lz, csr, first, rows):
    rp, ci, _ = csr
    sub = np.ascontiguousarray(rp[first:first + rows + 1])
    cols = np.ascontiguousarray(ci[sub[0]:sub[-1]])
    out = (C.c_int64 * 3)()
    ns = lz.lib().lz_grid_strides_host(rows, first, sub.ctypes.data, cols.ctypes.data, out)      # cols[k] = entry rowptr[first] + k
    return ns, tuple(out)


@pytest.mark.parametrize("dims", [(40, 36, 24), (256, 64, 8), (24, 20, 16)])
def test_strides_of_a_3d_stencil(lz, orc, dims):
    nx, ny, nz = dims
    csr = orc.lap3d(nx, ny, nz)
    n = nx * ny * nz
    ns, st = strides_of(lz, csr, n // 2, min(4096, n // 2))
    assert ns == 3 and st == (1, nx, nx * ny)


def test_strides_of_a_2d_stencil_and_of_operators_without_a_grid(lz, orc, maxwell10):
    csr = orc.lap2d(96, 80)
    ns, st = strides_of(lz, csr, 96 * 40, 3840)
    assert ns == 2 and st[:2] == (1, 96)
    # 1-D chain: only the unit stride
    n = 4096
    rp = np.arange(0, 3 * n + 1, 3, dtype=np.int32)
    ci = np.clip(np.repeat(np.arange(n), 3).reshape(n, 3) + np.array([-1, 0, 1]), 0, n - 1).astype(np.int32).reshape(-1)
    ns, st = strides_of(lz, (rp, ci, None), n // 2, 1024)
    assert ns == 1 and st[0] == 1
    # the reference's Maxwell operator couples field components at unrelated offsets: no nested strides
    mrp, mci, _ = maxwell10["csr"]
    nm = len(mrp) - 1
    ns, _ = strides_of(lz, (mrp, mci, None), nm // 2, min(4096, nm // 2))
    assert ns < 2
    # too small a sample
    assert strides_of(lz, csr, 0, 32)[0] == 0


@pytest.mark.parametrize("dims,box", [((40, 36, 24), (32, 2, 2)), ((64, 24, 18), (16, 4, 2)), ((96, 80), (16, 8, 1)), ((33, 7, 5), (8, 4, 4)),
                                      ((256, 256, 4), (32, 2, 2))])
def test_box_order_is_a_permutation_with_contiguous_runs(lz, dims, box):
    n = int(np.prod(dims))
    ns = len(dims)
    st = (C.c_int64 * 3)(1, dims[0], dims[0] * dims[1] if ns == 3 else 0)
    rowmap = np.empty(n, dtype=np.int32)
    crow = np.empty(n + 1, dtype=np.int32)
    nch = C.c_int64(0)
    lz.check(lz.lib().lz_box_order_host(n, ns, st, box[0], box[1], box[2], rowmap.ctypes.data, n + 1, crow.ctypes.data, C.byref(nch)))
    k = nch.value
    crow = crow[:k + 1]
    assert np.array_equal(np.sort(rowmap), np.arange(n))                       # every row exactly once
    assert crow[0] == 0 and crow[-1] == n and np.all(np.diff(crow) > 0)
    lx, ty, tz = box if ns == 3 else (box[0], box[1], 1)
    assert np.max(np.diff(crow)) <= lx * ty * tz
    nx, ny = dims[0], dims[1]
    for c in (0, k // 3, k - 1):                                                # a box is a product set of grid coordinates
        rows = rowmap[crow[c]:crow[c + 1]].astype(np.int64)
        x, y, z = rows % nx, (rows // nx) % ny if ns == 3 else rows // nx, rows // (nx * ny) if ns == 3 else 0 * rows
        assert x.max() - x.min() < lx and y.max() - y.min() < ty and z.max() - z.min() < tz
        assert len(rows) == (x.max() - x.min() + 1) * (y.max() - y.min() + 1) * (z.max() - z.min() + 1)
        runs = np.split(rows, np.where(np.diff(rows) != 1)[0] + 1)              # runs of consecutive rows along x
        assert all(len(r) == x.max() - x.min() + 1 for r in runs)
    # interior boxes of a 7-point operator: window rows per output row (the L2 -> SM traffic the schedule exists to cut)
    if ns == 3 and all(d >= 3 * b for d, b in zip(dims, box)):
        win = (lx + 2) * ty * tz + 2 * lx * tz + 2 * lx * ty
        assert win / (lx * ty * tz) < 3.6 < 5.0


def test_box_order_rejects_bad_arguments(lz):
    st = (C.c_int64 * 3)(1, 16, 100)                                            # 100 is not a multiple of 16
    rowmap = np.empty(1600, dtype=np.int32); crow = np.empty(1601, dtype=np.int32); nch = C.c_int64(0)
    assert lz.lib().lz_box_order_host(1600, 3, st, 8, 2, 2, rowmap.ctypes.data, 1601, crow.ctypes.data, C.byref(nch)) != 0
    st = (C.c_int64 * 3)(1, 16, 160)
    assert lz.lib().lz_box_order_host(1600, 3, st, 8, 2, 2, rowmap.ctypes.data, 4, crow.ctypes.data, C.byref(nch)) != 0     # chunk_cap too small
    assert lz.lib().lz_box_order_host(1600, 3, st, 8, 2, 2, rowmap.ctypes.data, 1601, crow.ctypes.data, C.byref(nch)) == 0
