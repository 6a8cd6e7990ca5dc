"""GPU suite, N > 1: the sharded single-vector driver (row slabs, NCCL halo exchange + all-reduce)
against the single-GPU driver and the oracle.  Needs >= 2 GPUs (gpurun --gpus 2); skipped otherwise."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

WORKER = r'''
import os, sys, ctypes as C
import numpy as np, torch, torch.distributed as dist
ROOT = sys.argv[1]
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gpu-implementation-of-signle-and-block-lanczos_b200"))
import lanczos_b200 as lz
from oracle import orc
lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr); torch.zeros(1, device="cuda")
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rank, world = dist.get_rank(), dist.get_world_size()
orc.set_threads(4)

def make_ctx(env):
    # the knobs are read once per context: LZ_COMM=1 NCCL collectives, LZ_NO_OVERLAP halo exchange on the compute stream
    for k in ("LZ_COMM", "LZ_NO_OVERLAP", "LZ_NO_FOLD"):
        os.environ.pop(k, None)
    os.environ.update(env)
    ctx = lz.Context(lr)
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_ubyte * 128)(); lz.check(lz.lib().lz_comm_unique_id(buf)); ident = torch.tensor(list(buf), dtype=torch.uint8)
    ident = ident.cuda(); dist.broadcast(ident, 0)
    lz.check(lz.lib().lz_comm_init(ctx.h, world, rank, bytes(ident.cpu().tolist())))
    return ctx

results = {}
for tag, env in (("peer+overlap", {}), ("nccl+overlap", {"LZ_COMM": "1"}), ("peer, no overlap", {"LZ_NO_OVERLAP": "1"}),
                 ("peer, pass B kept", {"LZ_NO_FOLD": "1"})):
    ctx = make_ctx(env)
    for kind, dims, m in (("lap3d", (24, 20, 16), 40), ("lap2d", (64, 48), 40), ("lap3d", (40, 36, 24), 60)):
        n = int(np.prod(dims))
        if kind == "lap3d":
            A = lz.Matrix.laplacian3d_shard(ctx, *dims, world, rank); csr = orc.lap3d(*dims); gran = dims[0] * dims[1]
        else:
            A = lz.Matrix.laplacian2d_shard(ctx, *dims, world, rank); csr = orc.lap2d(*dims); gran = dims[0]
        lo, hi = C.c_int64(), C.c_int64()
        lz.check(lz.lib().lz_partition_rows(n, gran, world, rank, C.byref(lo), C.byref(hi)))
        bfull = orc.start_vector(n)
        b = torch.from_numpy(bfull[lo.value:hi.value].copy()).cuda()
        for reorth in (0, 1, 2, 3):       # none, CGS2, DGKS, selective (oracle: CGS2 for every reorthogonalising mode)
            al = torch.zeros(m, dtype=torch.float64, device="cuda"); be = torch.zeros(m, dtype=torch.float64, device="cuda")
            for rep in range(2):      # twice: the second solve reuses the arena and continues the sequence numbers
                lz.check(lz.lib().lz_vector_lanczos_sharded(ctx.h, A.h, b.data_ptr(), m, reorth, al.data_ptr(), be.data_ptr()))
            ctx.sync()
            ref = orc.vector_lanczos(csr, bfull, m, reorth=min(reorth, 1))
            a, bb = al.cpu().numpy(), be.cpu().numpy()
            scale = np.maximum(np.abs(ref["alpha"]), np.mean(ref["beta"][1:]))
            k = min(m, 50)
            ea = np.max(np.abs(a - ref["alpha"])[:k] / scale[:k]); eb = np.max((np.abs(bb - ref["beta"]) / ref["beta"])[:k])
            tol = 1e-10 if reorth < 3 else 1e-8       # selective: semi-orthogonal basis, coefficients agree to O(eps ||A||) in theory
            assert ea < tol and eb < tol, (tag, kind, dims, reorth, ea, eb)
            # every rank holds the same coefficients bit for bit
            t = al.clone(); dist.broadcast(t, 0); assert torch.equal(t, al), (tag, kind, reorth)
            results[(tag, kind, dims, reorth)] = (ea, eb)
            if reorth == 1:
                al_cgs2, be_cgs2 = al.clone(), be.clone()
        # the same slab through the host-array entry point (lz_csr_create_shard_host): identical coefficients
        rp, ci, va = A.csr_to_host()
        hlo = gran if rank > 0 else 0; hhi = gran if rank < world - 1 else 0
        A2 = lz.Matrix.from_csr_shard_host(ctx, rp, ci, va, hlo, hhi, n, lo.value, hlo, A.n_rows - hhi)
        al2 = torch.zeros(m, dtype=torch.float64, device="cuda"); be2 = torch.zeros(m, dtype=torch.float64, device="cuda")
        lz.check(lz.lib().lz_vector_lanczos_sharded(ctx.h, A2.h, b.data_ptr(), m, 1, al2.data_ptr(), be2.data_ptr()))
        ctx.sync()
        assert torch.equal(al2, al_cgs2) and torch.equal(be2, be_cgs2), (tag, "shard_host")
        A2.close(); A.close()
    # block path, row-sharded: panels carry halo rows, every Gram matrix is all-reduced
    for bw, dims, m in ((8, (24, 20, 16), 10), (16, (24, 20, 16), 8)):
        n = int(np.prod(dims)); gran = dims[0] * dims[1]
        A = lz.Matrix.laplacian3d_shard(ctx, *dims, world, rank); csr = orc.lap3d(*dims)
        lo, hi = C.c_int64(), C.c_int64()
        lz.check(lz.lib().lz_partition_rows(n, gran, world, rank, C.byref(lo), C.byref(hi)))
        Bfull = orc.start_block(n, bw)
        nl = hi.value - lo.value
        Bl = torch.from_numpy(np.ascontiguousarray(Bfull[lo.value:hi.value].T).reshape(-1)).cuda()      # local rows, column-major
        for reorth in (0, 1):
            al = torch.zeros(m * bw * bw, dtype=torch.float64, device="cuda"); be = torch.zeros((m + 1) * bw * bw, dtype=torch.float64, device="cuda")
            lz.check(lz.lib().lz_block_lanczos(ctx.h, A.h, Bl.data_ptr(), nl, bw, m, -1, reorth, al.data_ptr(), be.data_ptr(), None))
            ctx.sync()
            ref = orc.block_lanczos(csr, Bfull, m, reorth=reorth)
            a = al.cpu().numpy().reshape(m, bw, bw).transpose(0, 2, 1); b = be.cpu().numpy().reshape(m + 1, bw, bw).transpose(0, 2, 1)
            ea = max(np.max(np.abs(a[j] - ref["alpha"][j])) / np.max(np.abs(ref["alpha"][j])) for j in range(m))
            eb = max(np.max(np.abs(b[j] - ref["beta"][j])) / np.max(np.abs(ref["beta"][j])) for j in range(m))
            assert ea < 1e-10 and eb < 1e-10, (tag, "block", bw, reorth, ea, eb)
            t = al.clone(); dist.broadcast(t, 0); assert torch.equal(t, al), (tag, "block", bw, reorth)
        A.close()
    peer, timed_out = ctx.comm_status()
    assert not timed_out
    if rank == 0:
        print("mode %-18s peer_memory=%d  worst alpha/beta rel err vs oracle %.2e / %.2e" % (
            tag, peer, max(v[0] for v in results.values()), max(v[1] for v in results.values())), flush=True)
    ctx.close()
print("rank %d ok" % rank)
dist.destroy_process_group()
'''


def test_sharded_vector_lanczos_matches_oracle(tmp_path):
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else 4
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29641", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "test_gpu_multi_world%d.log" % world), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-6000:])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count(" ok") == world
