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
    """The oracle's recurrence (OpenMP, all host threads) on the first `iters` iterations of the workload.
    Returns (iterations/s, seconds, threads, alpha, beta)."""
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
    return iters / dt, dt, threads, np.array(r["alpha"][:iters]), np.array(r["beta"][:iters])


def coeff_err(alpha, beta, ra, rb):
    """north_star tolerance form: |d alpha| relative to max(|alpha|, mean beta), beta relative."""
    import numpy as np
    k = min(len(ra), len(alpha))
    scale = np.maximum(np.abs(ra[:k]), np.mean(np.abs(rb[1:k])) if k > 1 else 1.0)
    ea = float(np.max(np.abs(alpha[:k] - ra[:k]) / scale))
    eb = float(np.max(np.abs(beta[:k] - rb[:k]) / np.abs(rb[:k])))
    return ea, eb


def run_reference(args, wl, rank):
    if rank != 0:
        return
    iters = min(CPU_SAMPLE_ITERS, wl["m"])
    for _ in range(max(0, min(args.warmup, 1))):          # one untimed pass is enough to page everything in
        cpu_sample(wl, min(4, iters))
    vals, times, threads = [], [], 1
    for _ in range(args.steps):
        v, dt, threads = cpu_sample(wl, iters)[:3]
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


# ----------------------------------------------------------------------------------- extra legs (N = 1)

def _events():
    import torch
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def fp64_dgemm_peak():
    """cuBLAS DGEMM 8192^3 through torch, best of 5: the fp64 dense peak the reorthogonalisation GEMMs are held against."""
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    c = torch.empty_like(a)
    torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = _events()
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b, c
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def cfg5_leg(ctx, lz, world, rank, dist, barrier):
    """BASELINE configs[4] (512^3, 100 steps, no reorth) at THIS world size, outside the headline timed region:
    1 warm-up + 2 timed solves, device-timed, max over ranks."""
    import numpy as np
    import torch
    wl = WORKLOADS["cfg5"]
    m = wl["m"]
    n_global = int(np.prod(wl["dims"]))
    granule = n_global // wl["dims"][-1]
    A5 = lz.Matrix.laplacian3d_shard(ctx, *wl["dims"], world, rank) if world > 1 else lz.Matrix.laplacian3d(ctx, *wl["dims"])
    b5 = torch.empty(A5.n_rows, dtype=torch.float64, device="cuda")
    if world > 1:
        lo, hi = lz.C.c_int64(), lz.C.c_int64()
        lz.check(lz.lib().lz_partition_rows(n_global, granule, world, rank, lz.C.byref(lo), lz.C.byref(hi)))
        full = torch.empty(n_global, dtype=torch.float64, device="cuda")
        lz.check(lz.lib().lz_gen_start_vector(ctx.h, n_global, 0x5EED, full.data_ptr()))
        ctx.sync()
        b5.copy_(full[lo.value:hi.value])
        del full
    else:
        lz.check(lz.lib().lz_gen_start_vector(ctx.h, A5.n_rows, 0x5EED, b5.data_ptr()))
    al = torch.zeros(m, dtype=torch.float64, device="cuda")
    be = torch.zeros(m, dtype=torch.float64, device="cuda")

    def solve5():
        if world > 1:
            lz.check(lz.lib().lz_vector_lanczos_sharded(ctx.h, A5.h, b5.data_ptr(), m, 0, al.data_ptr(), be.data_ptr()))
        else:
            lz.vector_lanczos_async(ctx, A5, b5, m, al, be, reorth=0)
    solve5()
    barrier()
    reps = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        solve5()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if dist:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    a5 = al.cpu().numpy()
    assert np.all(np.isfinite(a5)) and abs(a5[0] - 6.0) < 6.0, "cfg5 leg: bad alpha"
    A5.close()
    del b5
    return {"workload": "cfg5: " + wl["desc"], "n_gpus": world, "iterations_per_s": m * reps / (ms * 1e-3), "ms_per_solve": ms / reps,
            "alpha_0": float(a5[0]), "beta_last": float(be.cpu().numpy()[m - 1])}


def extra_legs(ctx, lz, peak, A2=None, b2=None, m2=300):
    """north_star targets and the reference-CUDA baseline, measured in the same process right after the headline
    (outside its timed region): (a) fused single-vector step on 256^3 as a fraction of HBM peak; (b) block b = 16 step
    on 256^3 without and with block CGS2, the reorthogonalisation GEMMs in TFLOP/s against the fp64 DGEMM peak measured
    here and against min(peak, AI * HBM); (c) the reference's own CUDA build (sm_100) and this library's C++ mirror on
    the reference's matrix (Maxwell N = 160)."""
    import re
    import torch
    out = {}
    if A2 is not None:
        # the headline solve under the other reorthogonalisation modes (same operator, same start vector): what the
        # always-two-sweeps CGS2 of the headline costs against a conditional second sweep and against the omega recurrence
        al = torch.zeros(m2, dtype=torch.float64, device="cuda"); be = torch.zeros_like(al)
        for mode, name in ((lz.REORTH_FULL_DGKS, "cgs_dgks"), (lz.REORTH_SELECTIVE, "selective")):
            lz.vector_lanczos_async(ctx, A2, b2, m2, al, be, reorth=mode); ctx.sync()
            e0, e1 = _events()
            e0.record(); lz.vector_lanczos_async(ctx, A2, b2, m2, al, be, reorth=mode); e1.record(); torch.cuda.synchronize()
            rec = {"it_per_s": m2 / (e0.elapsed_time(e1) * 1e-3), "ms_per_solve": e0.elapsed_time(e1)}
            if mode == lz.REORTH_SELECTIVE:
                rec["steps_that_reorthogonalised"] = lz.reorth_count(ctx)
            out["cfg2_" + name] = rec
        out["cfg2_modes_note"] = ("same 300-step solve as the headline (CGS2, two sweeps every step) with LZ_REORTH_FULL_DGKS / "
                                  "LZ_REORTH_SELECTIVE; reported as algorithmic headroom, not as the headline")
    A = lz.Matrix.laplacian3d(ctx, 256, 256, 256)
    n, nnz = A.n_rows, A.nnz
    # (a) fused vector step
    m = 100
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_vector(ctx.h, n, 0x5EED, b.data_ptr()))
    al = torch.zeros(m, dtype=torch.float64, device="cuda"); be = torch.zeros_like(al)
    lz.vector_lanczos_async(ctx, A, b, m, al, be); ctx.sync()
    e0, e1 = _events()
    e0.record()
    for _ in range(3):
        lz.vector_lanczos_async(ctx, A, b, m, al, be)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3 / m
    byt = 12.0 * nnz + 52.0 * n
    out["cfg3_vector_step"] = {"workload": "3-D 7-pt Laplacian 256^3, single vector, no reorth, 100 steps",
                               "ms_per_step": ms, "it_per_s": 1e3 / ms, "algorithmic_bytes": byt,
                               "gbs": byt / ms / 1e6, "frac": byt / ms / 1e6 / peak, "target_frac": 0.80}
    del b, al, be
    # (b) block step, b = 16
    bw = 16
    peak64 = fp64_dgemm_peak()
    B = torch.empty(n * bw, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_block(ctx.h, n, bw, n, 0x5EED, B.data_ptr()))
    for reorth, mb in ((0, 12), (1, 12), (2, 12)):
        alb = torch.zeros(mb * bw * bw, dtype=torch.float64, device="cuda")
        beb = torch.zeros((mb + 1) * bw * bw, dtype=torch.float64, device="cuda")
        run = lambda: lz.block_lanczos(ctx, A, B, n, bw, mb, alb, beb, None, lc=-1, reorth=reorth)
        run(); ctx.sync()
        ctx.profile(True)
        e0, e1 = _events()
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        prof = ctx.profile_read(); ctx.profile(False)
        assert lz.block_status(ctx, mb) == mb and bool(torch.isfinite(alb).all())
        ms = e0.elapsed_time(e1) / mb
        rec = {"workload": "3-D 7-pt Laplacian 256^3, block b=16, %d blocks, %s" % (mb, ["no reorth", "block CGS2", "block CGS + conditional second sweep (DGKS)"][reorth]),
               "ms_per_step": ms, "it_per_s": 1e3 / ms,
               "classes_ms_per_step": {k: round(v[1] / mb, 4) for k, v in prof.items() if v[0]}}
        if not reorth:
            # shipped formulation: normalise (2 panel passes) + SpMM with fused subtraction and Gram (matrix + gather + Q_{j-1}
            # + W = 3) + one-term panel update with the next Gram fused (3): 8 panel passes (round 1 moved 10)
            byt = 12.0 * nnz + 4.0 * n + 8 * 8.0 * n * bw
            rec.update({"algorithmic_bytes": byt, "frac": byt / ms / 1e6 / peak, "spmm_ms": prof["spmm"][1] / max(prof["spmm"][0], 1),
                        "spmm_frac": prof["spmm"][2] / max(prof["spmm"][1], 1e-9) / 1e6 / peak})
            out["cfg3_block_step"] = rec
        elif reorth == 2:
            out["cfg3_block_reorth_dgks"] = rec
        else:
            # two sweeps per step against J = j + 1 stored blocks: flops of one product = 2 n b (J b)
            flops = sum(2 * 2.0 * n * bw * (j + 1) * bw for j in range(mb))
            ai_roof = min(peak64, (bw / 4.0) * peak / 1e3)           # AI = b/4 flop per basis byte
            for cls, name in (("cgs_project", "projection"), ("cgs_update", "update")):
                tf = flops / (prof[cls][1] * 1e-3) / 1e12
                rec[name + "_tflops"] = tf
                rec[name + "_frac_of_fp64_peak"] = tf / peak64
                rec[name + "_frac_of_min_peak_AIxHBM"] = tf / ai_roof
            rec.update({"fp64_dgemm_peak_tflops": peak64, "min_peak_AIxHBM_tflops": ai_roof, "target_frac_of_fp64_peak": 0.60})
            out["cfg3_block_reorth"] = rec
        del alb, beb
    del B
    A.close()
    # (c) the reference's CUDA build and this library's mirror harness on the reference's own operator
    ref = {}
    rdir = os.path.join(ROOT, "oracle", "_ref")
    hdir = os.path.join(PKG, "host")

    def run_cmd(cmd, pat):
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            mo = re.search(pat, r.stdout + r.stderr)
            return float(mo.group(1)) if mo else None
        except (OSError, subprocess.TimeoutExpired):
            return None

    N = 160
    tmp = "/tmp/lz_ref_dump.bin"
    for tag, nc, mode, mm in (("vector", 4, "vector", 50), ("block4", 4, "block", 10), ("block8", 8, "block", 10), ("block16", 16, "block", 10)):
        exe = os.path.join(rdir, "ref_cuda_dump_%d" % nc)
        if os.path.exists(exe):
            ref["reference_cuda_%s_it_s" % tag] = run_cmd([exe, mode, str(N), str(mm), tmp], r"\(([0-9.]+) iterations/s\)")
        hexe = os.path.join(hdir, "test_lanczos")
        if os.path.exists(hexe):
            cmd = [hexe, "-N", str(N), "-m", str(mm)] + (["--vector"] if mode == "vector" else ["--block", str(nc)])
            ref["lanczos_b200_%s_it_s" % tag] = run_cmd(cmd, r"iterations/s:\s*([0-9.]+)")
    ref["operator"] = "reference Matrix_A(160,160,160): Maxwell curl operator, n = 24 806 880, ELL width 4; cold single calls"
    out["ref_cuda_baseline"] = ref
    return out


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

    lo_row, hi_row = 0, int(np.prod(wl["dims"]))
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
    granule = n_global // wl["dims"][-1]
    b = torch.empty(n_local, dtype=torch.float64, device="cuda")
    if world > 1:
        lo, hi = lz.C.c_int64(), lz.C.c_int64()
        lz.check(lz.lib().lz_partition_rows(n_global, granule, world, rank, lz.C.byref(lo), lz.C.byref(hi)))
        lo_row, hi_row = lo.value, hi.value
        full = torch.empty(n_global, dtype=torch.float64, device="cuda")
        lz.check(lz.lib().lz_gen_start_vector(ctx.h, n_global, 0x5EED, full.data_ptr()))
        ctx.sync()
        b.copy_(full[lo_row:hi_row])
        del full
    else:
        lz.check(lz.lib().lz_gen_start_vector(ctx.h, n_local, 0x5EED, b.data_ptr()))
    alpha = torch.zeros(m, dtype=torch.float64, device="cuda")
    beta = torch.zeros(m, dtype=torch.float64, device="cuda")

    def solve(mat=None, rhs=None):
        mat, rhs = mat or A, b if rhs is None else rhs
        if world > 1:
            lz.check(lz.lib().lz_vector_lanczos_sharded(ctx.h, mat.h, rhs.data_ptr(), m, reorth, alpha.data_ptr(), beta.data_ptr()))
        else:
            lz.vector_lanczos_async(ctx, mat, rhs, m, alpha, beta, reorth=reorth)

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
    a_host, b_host = alpha.cpu().numpy(), beta.cpu().numpy()
    assert np.all(np.isfinite(a_host)), "non-finite alpha: the solve broke down"
    value = m * args.steps / (ms_total * 1e-3)
    peer_mode = None
    if world > 1:
        peer_mode, timed_out = ctx.comm_status()
        assert not timed_out, "a peer-memory wait timed out"

    # ---- parity: the first CPU_SAMPLE_ITERS coefficients of THIS solve against the CPU oracle (every rank checks) ----
    parity, cpu = None, None
    if not args.no_cpu:
        k = min(CPU_SAMPLE_ITERS, m)
        ref = torch.zeros(2 * k + 3, dtype=torch.float64)
        if rank == 0:
            v, dt, threads, ra, rb = cpu_sample(wl, k)
            cpu = (v, dt, threads)
            ref = torch.from_numpy(np.concatenate([ra, rb, [v, dt, threads]]))
        if dist:
            ref = ref.cuda()
            dist.broadcast(ref, 0)
            ref = ref.cpu()
        ref = ref.numpy()
        ea, eb = coeff_err(a_host, b_host, ref[:k], ref[k:2 * k])
        assert ea < 1e-10 and eb < 1e-10, "rank %d: alpha/beta differ from the oracle: %.3e / %.3e" % (rank, ea, eb)
        if dist:
            t = torch.tensor([ea, eb], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ea, eb = float(t[0]), float(t[1])
        parity = {"against": "oracle/lanczos_oracle.c (CPU restatement), first %d alpha/beta of the timed solve, every rank" % k,
                  "max_rel_alpha": ea, "max_rel_beta": eb, "tolerance": 1e-10}

    # ---- end-to-end through the C-ABI from pinned host buffers -----------------------------------
    # every step: the rank's CSR slab and start vector go host -> device (lz_csr_create_host / lz_csr_create_shard_host
    # + lz_memcpy), the solve runs, alpha/beta come back device -> host
    e2e = None
    if not args.no_e2e:
        rp, ci, va = A.csr_to_host()
        pin = lambda a: torch.from_numpy(a).pin_memory()
        rp_h, ci_h, va_h, b_h = pin(rp), pin(ci), pin(va), b.cpu().pin_memory()
        h2d = rp_h.numel() * 4 + ci_h.numel() * 4 + va_h.numel() * 8 + b_h.numel() * 8
        bd = torch.empty_like(b)
        a_out, b_out = np.zeros(m), np.zeros(m)
        hlo = granule if (world > 1 and rank > 0) else 0
        hhi = granule if (world > 1 and rank < world - 1) else 0

        create_ms = [0.0]

        def e2e_step():
            tc = time.perf_counter()
            if world > 1:
                A2 = lz.Matrix.from_csr_shard_host(ctx, rp_h.numpy(), ci_h.numpy(), va_h.numpy(), hlo, hhi, n_global, lo_row,
                                                   hlo, n_local - hhi)
            else:
                A2 = lz.Matrix.from_csr_host(ctx, rp_h.numpy(), ci_h.numpy(), va_h.numpy())
            lz.check(lz.lib().lz_memcpy(ctx.h, bd.data_ptr(), b_h.data_ptr(), b_h.numel() * 8, lz.H2D))
            create_ms[0] = (time.perf_counter() - tc) * 1e3       # upload + chunk schedules (the create call synchronises)
            if world > 1:
                solve(A2, bd)
                a_out[:] = alpha.cpu().numpy(); b_out[:] = beta.cpu().numpy()
            else:
                steps = lz.C.c_int(0)
                lz.check(lz.lib().lz_vector_lanczos(ctx.h, A2.h, bd.data_ptr(), m, 0, reorth, a_out.ctypes.data,
                                                    b_out.ctypes.data, None, lz.C.byref(steps)))
            A2.close()
        e2e_step()                                           # warm-up
        barrier()
        reps = max(1, min(args.steps, 3))
        times = []
        for _ in range(reps):                                # median of up to 3 steps: host-side hiccups (page pinning, allocator) are not the metric
            t0 = time.perf_counter()
            e2e_step()
            barrier()
            times.append(time.perf_counter() - t0)
        dt = sorted(times)[len(times) // 2]
        assert np.allclose(a_out, a_host, rtol=1e-9, atol=1e-12)
        if dist:
            t = torch.tensor([dt, float(h2d)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
            t2 = torch.tensor([float(h2d)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t2)
            h2d = int(t2.item())
        e2e = {"value": m / dt, "unit": "iterations/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(16 * m * world),
               "ms_per_step": dt * 1e3, "steps_ms": [round(t * 1e3, 1) for t in times], "upload_and_schedule_ms": round(create_ms[0], 1)}
        del rp_h, ci_h, va_h, b_h, bd

    extra = None
    if world == 1 and not args.no_extra and args.workload == "cfg2":
        extra = extra_legs(ctx, lz, peak, A, b, m)
    if not args.no_extra and args.workload == "cfg2":
        c5 = cfg5_leg(ctx, lz, world, rank, dist, barrier)          # every rank runs it (sharded solve)
        extra = dict(extra or {}, cfg5=c5)

    def teardown():
        # identical order on every rank: operators, library communicator + context, then torch's
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
    # shares of the PROFILED time (sum of the per-class event times); the event ring is drained when it fills, so
    # every launch of the timed region is counted
    prof_ms = sum(v[1] for v in prof.values())
    share = {k: round(v[1] / prof_ms, 4) for k, v in prof.items() if v[0]}
    line = {"metric": "lanczos_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + wl["desc"], "rows": n_global, "m": m,
                       "reorth": ["none", "cgs2", "cgs-dgks"][reorth], "partition": "row blocks x%d" % world,
                       "l2": "inputs (>=1 GB matrix, 134 MB vectors, basis up to 40 GB) exceed the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_kind": peak_kind,
                         "launches": la, "kernel_ms": kms, "algorithmic_bytes_per_launch": kby / max(la, 1),
                         "share_of_step": share, "profiled_ms": prof_ms, "timed_ms": ms_total,
                         "launches_by_class": {k: v[0] for k, v in prof.items() if v[0]}},
            "gpu_launches": launches, "clocks": clocks}
    if world > 1:
        line["comm"] = {"mode": "peer-memory (CUDA IPC over NVLink): one-shot all-reduces + pushed halos" if peer_mode else "nccl",
                        "share_of_profiled_time": share.get("comm", 0.0),
                        "note": "stand-alone all-reduce kernels only; scalar all-reduces run inside the last CTA of the "
                                "producing kernel and the halo exchange overlaps the interior SpMV on a side stream"}
    if parity:
        line["parity"] = parity
    if e2e:
        line["e2e"] = e2e
    if cpu:
        v, dt, threads = cpu
        line["cpu_baseline"] = {"value": v, "unit": "iterations/s", "cores": int(threads), "kind": "port",
                                "sample": "first %d of %d iterations of the same solve, oracle/lanczos_oracle.c with OpenMP (%.1f s)"
                                          % (min(CPU_SAMPLE_ITERS, m), m, dt)}
    if extra:
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    teardown()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-extra", action="store_true", help="skip the north_star / reference-CUDA extra legs")
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
