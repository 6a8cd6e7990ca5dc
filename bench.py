#!/usr/bin/env python
"""bench.py -- the contract benchmark.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg3|cfg5]

Workload (BASELINE.json configs[1]): single-vector Lanczos, m = 300 steps with full
reorthogonalisation (CGS2 against the stored basis), 2-D 5-point Laplacian 4096^2 (16.7 M rows,
83.9 M nnz, fp64).  One "step" of this bench = one whole 300-iteration solve; `value` = Lanczos
iterations per second (300 * K / time) with the operator and the start vector resident in HBM.
For N > 1 the same solve is row-sharded over N GPUs (strong scaling): halo exchange of one grid
line per neighbour before each SpMV, packed all-reduces of alpha / beta^2 / the CGS coefficients.

`e2e` drives the same solve through the C-ABI the reference-facing C++ mirror binds
(lz_csr_create_host + lz_vector_lanczos) from pinned HOST buffers: CSR arrays and the start vector
go host->device and alpha/beta come back device->host inside the timed region, every step.

--impl reference times the CPU restatement of the reference's recurrence (oracle/, OpenMP, all host
threads) on a bounded sample of the same workload; the reference repo has no CSR, no
reorthogonalisation and no CPU driver of its own, so `kind` is "port".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gpu-implementation-of-signle-and-block-lanczos_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

WORKLOADS = {
    # name: (generator, dims, m, reorth)
    "cfg2": dict(desc="single-vector Lanczos, 300 steps, full reorth (CGS2), 2-D 5-pt Laplacian 4096^2",
                 kind="lap2d", dims=(4096, 4096), m=300, reorth=1),
    "cfg3v": dict(desc="single-vector Lanczos, 300 steps, no reorth, 3-D 7-pt Laplacian 256^3",
                  kind="lap3d", dims=(256, 256, 256), m=300, reorth=0),
    "cfg5": dict(desc="single-vector Lanczos, 100 steps, no reorth, 3-D 7-pt Laplacian 512^3",
                 kind="lap3d", dims=(512, 512, 512), m=100, reorth=0),
    "small": dict(desc="single-vector Lanczos, 50 steps, full reorth, 2-D 5-pt Laplacian 512^2 (smoke-size)",
                  kind="lap2d", dims=(512, 512), m=50, reorth=1),
}
CPU_SAMPLE_ITERS = 30


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except (ValueError, IndexError):
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------- CPU arm

def cpu_sample(wl, iters):
    """The oracle's recurrence (OpenMP, all host threads) on the first `iters` iterations of the workload."""
    import numpy as np
    from oracle import orc
    threads = os.cpu_count() or 1
    orc.set_threads(threads)
    csr = orc.lap2d(*wl["dims"]) if wl["kind"] == "lap2d" else orc.lap3d(*wl["dims"])
    n = len(csr[0]) - 1
    b = orc.start_vector(n)
    t0 = time.perf_counter()
    r = orc.vector_lanczos(csr, b, iters, reorth=wl["reorth"])
    dt = time.perf_counter() - t0
    assert r["steps"] == iters and np.all(np.isfinite(r["alpha"]))
    return iters / dt, dt, threads


def run_reference(args, wl, rank):
    if rank != 0:
        return
    iters = min(CPU_SAMPLE_ITERS, wl["m"])
    for _ in range(max(0, min(args.warmup, 1))):          # one untimed pass is enough to page everything in
        cpu_sample(wl, min(4, iters))
    vals, times, threads = [], [], 1
    for _ in range(args.steps):
        v, dt, threads = cpu_sample(wl, iters)
        vals.append(v); times.append(dt)
    value = iters * len(times) / sum(times)
    sample = "first %d of %d iterations of %s (per-iteration cost grows with the basis, so this flatters the CPU)" % (
        iters, wl["m"], args.workload)
    line = {"impl": "reference", "metric": "lanczos_iterations_per_s", "value": value, "unit": "iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"]},
            "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------- GPU arm

def run_gpu(args, wl, rank, world):
    import numpy as np
    import torch
    import lanczos_b200 as lz

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    torch.zeros(1, device="cuda")
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = lz.Context(local_rank)
    m, reorth = wl["m"], wl["reorth"]
    peak, peak_kind = peaks()

    if world > 1:
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (lz.C.c_ubyte * 128)()
            lz.check(lz.lib().lz_comm_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.cuda()
        dist.broadcast(ident, 0)
        idb = bytes(ident.cpu().tolist())
        lz.check(lz.lib().lz_comm_init(ctx.h, world, rank, idb))
        A = (lz.Matrix.laplacian2d_shard(ctx, *wl["dims"], world, rank) if wl["kind"] == "lap2d"
             else lz.Matrix.laplacian3d_shard(ctx, *wl["dims"], world, rank))
    else:
        A = lz.Matrix.laplacian2d(ctx, *wl["dims"]) if wl["kind"] == "lap2d" else lz.Matrix.laplacian3d(ctx, *wl["dims"])
    n_local = A.n_rows
    n_global = int(np.prod(wl["dims"]))
    b = torch.empty(n_local, dtype=torch.float64, device="cuda")
    if world > 1:
        lo, hi = lz.C.c_int64(), lz.C.c_int64()
        granule = n_global // wl["dims"][-1]
        lz.check(lz.lib().lz_partition_rows(n_global, granule, world, rank, lz.C.byref(lo), lz.C.byref(hi)))
        full = torch.empty(n_global, dtype=torch.float64, device="cuda")
        lz.check(lz.lib().lz_gen_start_vector(ctx.h, n_global, 0x5EED, full.data_ptr()))
        ctx.sync()
        b.copy_(full[lo.value:hi.value])
        del full
    else:
        lz.check(lz.lib().lz_gen_start_vector(ctx.h, n_local, 0x5EED, b.data_ptr()))
    alpha = torch.zeros(m, dtype=torch.float64, device="cuda")
    beta = torch.zeros(m, dtype=torch.float64, device="cuda")

    def solve():
        if world > 1:
            lz.check(lz.lib().lz_vector_lanczos_sharded(ctx.h, A.h, b.data_ptr(), m, reorth, alpha.data_ptr(), beta.data_ptr()))
        else:
            lz.vector_lanczos_async(ctx, A, b, m, alpha, beta, reorth=reorth)

    def barrier():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        solve()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        solve()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile(False)
    launches = ctx.launches - launches0
    clocks = sampler.stop() if rank == 0 else None
    if dist:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(lt)
        launches = int(lt.item())
    a_host = alpha.cpu().numpy()
    assert np.all(np.isfinite(a_host)), "non-finite alpha: the solve broke down"
    value = m * args.steps / (ms_total * 1e-3)

    # ---- end-to-end through the C-ABI from pinned host buffers (single GPU path) --------------
    e2e = None
    if world == 1 and not args.no_e2e:
        rp, ci, va = A.csr_to_host()
        pin = lambda a: torch.from_numpy(a).pin_memory()
        rp_h, ci_h, va_h, b_h = pin(rp), pin(ci), pin(va), b.cpu().pin_memory()
        h2d = rp_h.numel() * 4 + ci_h.numel() * 4 + va_h.numel() * 8 + b_h.numel() * 8
        bd = torch.empty_like(b)
        a_out, b_out = np.zeros(m), np.zeros(m)

        def e2e_step():
            A2 = lz.Matrix.from_csr_host(ctx, rp_h.numpy(), ci_h.numpy(), va_h.numpy())
            lz.check(lz.lib().lz_memcpy(ctx.h, bd.data_ptr(), b_h.data_ptr(), b_h.numel() * 8, lz.H2D))
            steps = lz.C.c_int(0)
            lz.check(lz.lib().lz_vector_lanczos(ctx.h, A2.h, bd.data_ptr(), m, 0, reorth, a_out.ctypes.data,
                                                b_out.ctypes.data, None, lz.C.byref(steps)))
            A2.close()
        e2e_step()                                           # warm-up
        torch.cuda.synchronize()
        reps = max(1, min(args.steps, 2))
        t0 = time.perf_counter()
        for _ in range(reps):
            e2e_step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        assert np.allclose(a_out, a_host, rtol=1e-9, atol=1e-12)
        e2e = {"value": m / dt, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(16 * m),
               "ms_per_step": dt * 1e3}

    def teardown():
        # identical order on every rank: library communicator first, then torch's
        A.close()
        ctx.close()
        if dist:
            dist.barrier()
            dist.destroy_process_group()

    if rank != 0:
        teardown()
        return
    # ---- roofline of the dominant kernel (live CUDA-event times from the timed region) -------
    dom = max((k for k in prof if k != "comm"), key=lambda k: prof[k][1])
    la, kms, kby = prof[dom]
    achieved = kby / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(dom)
        if tj:
            traffic = tj["dram_bytes_per_algorithmic_byte"] * kby / max(la, 1)
    share = {k: round(v[1] / ms_total, 4) for k, v in prof.items() if v[0]}
    line = {"metric": "lanczos_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "rows": n_global, "m": m,
                       "reorth": ["none", "cgs2", "cgs-dgks"][reorth], "partition": "row blocks x%d" % world,
                       "l2": "inputs (>=1 GB matrix, 134 MB vectors, basis up to 40 GB) exceed the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "launches": la, "kernel_ms": kms, "algorithmic_bytes_per_launch": kby / max(la, 1),
                         "share_of_step": share},
            "gpu_launches": launches, "clocks": clocks}
    if e2e:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        v, dt, threads = cpu_sample(wl, min(CPU_SAMPLE_ITERS, m))
        line["cpu_baseline"] = {"value": v, "unit": "iterations/s", "cores": threads, "kind": "port",
                                "sample": "first %d of %d iterations of the same solve, oracle/lanczos_oracle.c with OpenMP (%.1f s)"
                                          % (min(CPU_SAMPLE_ITERS, m), m, dt)}
    print(json.dumps(line), flush=True)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_gpu(args, wl, rank, world)


if __name__ == "__main__":
    main()
