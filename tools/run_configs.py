"""Runs the BASELINE.json configs that are not the bench headline (3: block b=16 on 256^3, 4: block b=32
on R-MAT scale 24, single-vector no-reorth steps) and records iterations/s, per-class times and
roofline fractions in gpurun_out/configs.json.  Developer/record tool, single GPU."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-implementation-of-signle-and-block-lanczos_b200"))
import numpy as np
import torch
import lanczos_b200 as lz

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timed(ctx, fn, reps):
    fn(); ctx.sync()
    ctx.profile(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    prof = ctx.profile_read(); ctx.profile(False)
    return e0.elapsed_time(e1) / reps, {k: dict(launches=v[0] // reps, ms=v[1] / reps, gbs=(v[2] / v[1] / 1e6 if v[1] else 0)) for k, v in prof.items() if v[0]}


def vector_case(ctx, A, name, m, out):
    n, nnz = A.n_rows, A.nnz
    b = torch.empty(n, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_vector(ctx.h, n, 0x5EED, b.data_ptr()))
    al = torch.zeros(m, dtype=torch.float64, device="cuda"); be = torch.zeros_like(al)
    ms, prof = timed(ctx, lambda: lz.vector_lanczos_async(ctx, A, b, m, al, be), 3)
    byt = 12.0 * nnz + 52.0 * n
    out[name] = dict(ms_per_iter=ms / m, it_per_s=m / ms * 1e3, step_gbs=byt / (ms / m) / 1e6, step_frac=byt / (ms / m) / 1e6 / PEAK, classes=prof)


def block_case(ctx, A, name, bw, m, out, reorth=0):
    n, nnz = A.n_rows, A.nnz
    B = torch.empty(n * bw, dtype=torch.float64, device="cuda")
    lz.check(lz.lib().lz_gen_start_block(ctx.h, n, bw, n, 0x5EED, B.data_ptr()))
    al = torch.zeros(m * bw * bw, dtype=torch.float64, device="cuda"); be = torch.zeros((m + 1) * bw * bw, dtype=torch.float64, device="cuda")
    q = torch.zeros(m * bw, dtype=torch.float64, device="cuda")
    ms, prof = timed(ctx, lambda: lz.block_lanczos(ctx, A, B, n, bw, m, al, be, q, reorth=reorth), 2)
    assert torch.isfinite(al).all()
    for k in ("cgs_project", "cgs_update"):       # reorth GEMMs: flops = 2 n (J b) b per launch = bytes/8 * 2b / ... report TFLOP/s
        if k in prof:
            # bytes recorded = 8 n b (J + extra): flops = 2 * (bytes/8) * b  (each basis element does b MACs)
            prof[k]["tflops"] = 2.0 * (prof[k]["gbs"] * 1e9 / 8.0) * bw / 1e12
    byt = 12.0 * nnz + 4.0 * n + 10 * 8.0 * n * bw
    out[name] = dict(ms_per_iter=ms / m, it_per_s=m / ms * 1e3, step_gbs=byt / (ms / m) / 1e6, step_frac=byt / (ms / m) / 1e6 / PEAK, classes=prof)


def main():
    which = sys.argv[1:] or ["cfg3", "cfg3v", "cfg2v", "cfg4"]
    torch.cuda.set_device(0); torch.zeros(1, device="cuda")
    ctx = lz.Context(0)
    out = {}
    if "cfg2v" in which:
        A = lz.Matrix.laplacian2d(ctx, 4096, 4096); vector_case(ctx, A, "lap2d_4096_vector_noreorth", 100, out); A.close()
    if "cfg3v" in which or "cfg3" in which or "cfg3r" in which:
        A = lz.Matrix.laplacian3d(ctx, 256, 256, 256)
        if "cfg3v" in which: vector_case(ctx, A, "lap3d_256_vector_noreorth", 100, out)
        if "cfg3" in which: block_case(ctx, A, "lap3d_256_block16", 16, 12, out)
        if "cfg3r" in which: block_case(ctx, A, "lap3d_256_block16_reorth", 16, 24, out, reorth=1)
        A.close()
    if "cfg4" in which:
        scale = int(os.environ.get("RMAT_SCALE", "24"))
        A = lz.Matrix.rmat_laplacian(ctx, scale)
        rp = A.csr_to_host()[0]; lens = np.diff(rp)
        out["rmat_rows"] = dict(scale=scale, n=A.n_rows, nnz=A.nnz, max_row=int(lens.max()), p99=float(np.percentile(lens, 99)), mean=float(lens.mean()))
        del rp, lens
        vector_case(ctx, A, "rmat%d_vector_noreorth" % scale, 50, out)
        block_case(ctx, A, "rmat%d_block32" % scale, 32, 6, out)
        A.close()
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
