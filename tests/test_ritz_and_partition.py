"""CPU suite: host-side logic of the product library that needs no GPU -- Ritz extraction
(lz_ritz) against numpy.linalg.eigh of the oracle's T, and the row partition used by the sharded
drivers, including a world_size-2 gloo run that emulates the sharded recurrence (halo exchange +
all-reduced partial sums) on the CPU with the oracle doing the local arithmetic."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, load_gold


def test_ritz_scalar_matches_eigh(lz, orc, maxwell10):
    g = load_gold("maxwell_N10_vector_m100.npz")
    a, b = g["alpha"], g["beta"]
    theta, resid = lz.ritz(a, b, 10, beta_last=np.array([0.5]))
    T = orc.assemble_T(a, b)
    w, Y = np.linalg.eigh(T)
    sel = np.r_[0:5, 95:100]
    assert np.max(np.abs(theta - w[sel])) < 1e-8 * np.abs(w).max()          # north_star: Ritz values to 1e-8
    assert np.max(np.abs(resid - 0.5 * np.abs(Y[-1, sel]))) < 1e-8
    t2, r2 = orc.ritz(a, b, 10, beta_last=np.array([[0.5]]))
    assert np.max(np.abs(theta - t2)) < 1e-12 and np.max(np.abs(resid - r2)) < 1e-10


@pytest.mark.parametrize("nc", [4, 8])
def test_ritz_block_matches_eigh(lz, orc, nc):
    g = load_gold("maxwell_N10_block%d_m25.npz" % nc)
    m = 25
    a = g["alpha"].reshape(m, nc, nc).transpose(0, 2, 1)
    b = g["beta"].reshape(m + 1, nc, nc).transpose(0, 2, 1)
    bl = b[3]
    theta, resid = lz.ritz(g["alpha"], g["beta"][:m * nc * nc], 10, bw=nc, beta_last=np.ascontiguousarray(bl.T).reshape(-1))
    t2, r2 = orc.ritz(a, b, 10, beta_last=bl)
    assert np.max(np.abs(theta - t2)) < 1e-8 * np.abs(t2).max()
    assert np.max(np.abs(resid - r2)) < 1e-8 * max(1.0, np.abs(r2).max())


def test_ritz_rejects_bad_arguments(lz):
    with pytest.raises(lz.LanczosError):
        lz.ritz(np.ones(3), np.ones(3), 5)            # k > dim(T)


def test_partition_rows(lz):
    L = lz.lib()
    lo, hi = C.c_int64(), C.c_int64()
    n, plane = 512 ** 3, 512 ** 2
    covered = 0
    for w in (1, 2, 4, 8):
        prev = 0
        for r in range(w):
            lz.check(L.lz_partition_rows(n, plane, w, r, C.byref(lo), C.byref(hi)))
            assert lo.value == prev and lo.value % plane == 0 and hi.value > lo.value
            prev = hi.value
        assert prev == n
    # ragged: 10 planes over 4 ranks -> 2,3,2,3
    sizes = []
    for r in range(4):
        lz.check(L.lz_partition_rows(10 * 7, 7, 4, r, C.byref(lo), C.byref(hi)))
        sizes.append((hi.value - lo.value) // 7)
    assert sum(sizes) == 10 and max(sizes) - min(sizes) <= 1
    assert L.lz_partition_rows(10, 3, 2, 0, C.byref(lo), C.byref(hi)) == -1      # not a multiple of the granule
    assert L.lz_partition_rows(6, 3, 4, 0, C.byref(lo), C.byref(hi)) == -1       # fewer granules than ranks


GLOO_WORKER = r'''
import os, sys, ctypes as C
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "gpu-implementation-of-signle-and-block-lanczos_b200"))
import lanczos_b200 as lz
from oracle import orc
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
orc.set_threads(1)
nx, ny, nz, m = 6, 5, 8, 12
sxy, n = nx * ny, nx * ny * nz
rp, ci, va = orc.lap3d(nx, ny, nz)
lo, hi = C.c_int64(), C.c_int64()
lz.check(lz.lib().lz_partition_rows(n, sxy, world, rank, C.byref(lo), C.byref(hi)))
r0, r1 = lo.value, hi.value
hlo = sxy if rank > 0 else 0
hhi = sxy if rank < world - 1 else 0
# local slab with the product's column convention: [lower halo | local | upper halo], shift = r0 - hlo
s, e = rp[r0], rp[r1]
lrp = (rp[r0:r1 + 1] - s).astype(np.int32)
lci = (ci[s:e] - (r0 - hlo)).astype(np.int32)
assert lci.min() >= 0 and lci.max() < hlo + (r1 - r0) + hhi
lva = va[s:e].copy()
nl = r1 - r0
def allsum(x):
    t = torch.tensor([x], dtype=torch.float64); dist.all_reduce(t); return float(t[0])
def halo(u):
    buf = np.zeros(hlo + nl + hhi); buf[hlo:hlo + nl] = u
    reqs = []
    if rank > 0:
        reqs.append(dist.isend(torch.from_numpy(u[:sxy].copy()), rank - 1))
        lo_t = torch.zeros(sxy, dtype=torch.float64); reqs.append(dist.irecv(lo_t, rank - 1))
    if rank < world - 1:
        reqs.append(dist.isend(torch.from_numpy(u[nl - sxy:].copy()), rank + 1))
        hi_t = torch.zeros(sxy, dtype=torch.float64); reqs.append(dist.irecv(hi_t, rank + 1))
    for q in reqs: q.wait()
    if rank > 0: buf[:hlo] = lo_t.numpy()
    if rank < world - 1: buf[hlo + nl:] = hi_t.numpy()
    return buf
def local_spmv(xbuf):
    y = np.zeros(nl)
    for i in range(nl):
        y[i] = np.dot(lva[lrp[i]:lrp[i + 1]], xbuf[lci[lrp[i]:lrp[i + 1]]])
    return y
b = orc.start_vector(n)[r0:r1]
alpha, beta = np.zeros(m), np.zeros(m)
beta[0] = np.sqrt(allsum(np.dot(b, b)))
q0 = b / beta[0]
w = local_spmv(halo(q0)); alpha[0] = allsum(np.dot(w, q0)); w -= alpha[0] * q0
for j in range(1, m):
    beta[j] = np.sqrt(allsum(np.dot(w, w)))
    q1 = w / beta[j]
    w = local_spmv(halo(q1)) - beta[j] * q0
    alpha[j] = allsum(np.dot(w, q1)); w -= alpha[j] * q1
    q0 = q1
ref = orc.vector_lanczos((rp, ci, va), orc.start_vector(n), m)
assert np.max(np.abs(alpha - ref["alpha"])) < 1e-12 * np.abs(ref["alpha"]).max(), "alpha"
assert np.max(np.abs(beta - ref["beta"]) / ref["beta"]) < 1e-12, "beta"
print("rank %d ok" % rank)
dist.destroy_process_group()
'''


def test_sharded_recurrence_gloo_world2(tmp_path):
    """N > 1 host logic on CPU: lz_partition_rows + the halo / column-shift convention of the sharded
    operator + all-reduced alpha/beta reproduce the global recurrence (gloo, world_size 2 and 3)."""
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    for world, port in ((2, 29611), (3, 29612)):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), ROOT]
        env = dict(os.environ, OMP_NUM_THREADS="1")
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        assert r.stdout.count("ok") == world


GLOO_BLOCK_WORKER = GLOO_WORKER.split("b = orc.start_vector(n)[r0:r1]")[0] + r'''
# ---- block recurrence (methods/block_lanczos.hpp:105-166) over the row slabs: panels are nl x bw, the halo
# exchange moves bw-wide boundary planes, every b x b Gram block is all-reduced (what lz_block_lanczos does
# on a context with a communicator)
bw, mb = 3, 6
def allsum_mat(M):
    t = torch.from_numpy(np.ascontiguousarray(M)); dist.all_reduce(t); return t.numpy()
def halo_panel(U):
    return np.stack([halo(np.ascontiguousarray(U[:, c])) for c in range(bw)], axis=1)
def local_spmm(Xbuf):
    return np.stack([local_spmv(np.ascontiguousarray(Xbuf[:, c])) for c in range(bw)], axis=1)
Bfull = orc.start_block(n, bw)
B = Bfull[r0:r1]
alpha, beta = np.zeros((mb, bw, bw)), np.zeros((mb + 1, bw, bw))
S, Sinv = orc.sqrtm(allsum_mat(B.T @ B)); beta[0] = S
Q0 = B @ Sinv
W = local_spmm(halo_panel(Q0))
G = allsum_mat(W.T @ Q0); alpha[0] = 0.5 * (G + G.T)
W = W - Q0 @ alpha[0]
for j in range(1, mb):
    S, Sinv = orc.sqrtm(allsum_mat(W.T @ W)); beta[j] = S
    Q1 = W @ Sinv
    W = local_spmm(halo_panel(Q1)) - Q0 @ beta[j]
    G = allsum_mat(W.T @ Q1); alpha[j] = 0.5 * (G + G.T)
    W = W - Q1 @ alpha[j]
    Q0 = Q1
ref = orc.block_lanczos((rp, ci, va), Bfull, mb)
assert np.max(np.abs(alpha - ref["alpha"])) < 1e-10 * np.abs(ref["alpha"]).max(), "block alpha"
assert np.max(np.abs(beta[:mb] - ref["beta"][:mb])) < 1e-10 * np.abs(ref["beta"][:mb]).max(), "block beta"
print("rank %d ok" % rank)
dist.destroy_process_group()
'''


def test_sharded_block_recurrence_gloo_world2(tmp_path):
    """Block path, N > 1 host logic on CPU: slab partition, bw-wide halo planes and all-reduced b x b Gram
    blocks reproduce the global block recurrence of the oracle (gloo, world_size 2)."""
    script = tmp_path / "worker_block.py"
    script.write_text(GLOO_BLOCK_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29613", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
