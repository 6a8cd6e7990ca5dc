"""Developer micro-benchmarks (CUDA events on the library's stream). Not the contract bench."""
import os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gpu-implementation-of-signle-and-block-lanczos_b200"))
import numpy as np
import torch
import lanczos_b200 as lz

PEAK = 6435.1

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def main():
    which = sys.argv[1:] or ["spmv", "lanczos", "reorth"]
    torch.cuda.set_device(0); torch.zeros(1, device="cuda")
    ctx = lz.Context(0)
    out = {}
    for name, mk in (("lap2d_4096", lambda: lz.Matrix.laplacian2d(ctx, 4096, 4096)),
                     ("lap3d_256", lambda: lz.Matrix.laplacian3d(ctx, 256, 256, 256))):
        A = mk(); n, nnz = A.n_rows, A.nnz
        x = torch.empty(n, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
        lz.check(lz.lib().lz_gen_start_vector(ctx.h, n, 0x5EED, x.data_ptr()))
        if "spmv" in which:
            ms = timeit(lambda: lz.spmv(ctx, A, x, y))
            byt = 12 * nnz + 20 * n + 4
            out[name + "_spmv"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
        if "lanczos" in which:
            m = 50
            al = torch.empty(m, dtype=torch.float64, device="cuda"); be = torch.empty_like(al)
            ms = timeit(lambda: lz.vector_lanczos_async(ctx, A, x, m, al, be), reps=3, warm=1) / m
            byt = 12 * nnz + 52 * n
            out[name + "_step_noreorth"] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
        if "reorth" in which and name == "lap2d_4096":
            for mode in (1, 2):
                m = 100
                al = torch.empty(m, dtype=torch.float64, device="cuda"); be = torch.empty_like(al)
                ms = timeit(lambda: lz.vector_lanczos_async(ctx, A, x, m, al, be, reorth=mode), reps=2, warm=1)
                reads = (4 if mode == 1 else 2) * 8 * n * (m * (m + 1) / 2)
                out[name + "_reorth%d_m100" % mode] = dict(ms_total=ms, it_per_s=m / ms * 1e3, basis_gbs=reads / ms / 1e6)
        A.close()
    if "block" in which:
        for bw in [int(x) for x in os.environ.get("LZ_BLOCK_WIDTHS", "8,16,32").split(",")]:
            A = lz.Matrix.laplacian3d(ctx, 256, 256, 256); n, nnz = A.n_rows, A.nnz
            m = 8
            B = torch.empty(n * bw, dtype=torch.float64, device="cuda")
            lz.check(lz.lib().lz_gen_start_block(ctx.h, n, bw, n, 0x5EED, B.data_ptr()))
            al = torch.zeros(m * bw * bw, dtype=torch.float64, device="cuda"); be = torch.zeros((m + 1) * bw * bw, dtype=torch.float64, device="cuda")
            q = torch.zeros(m * bw, dtype=torch.float64, device="cuda")
            run = lambda: lz.block_lanczos(ctx, A, B, n, bw, m, al, be, q)
            run(); ctx.sync()
            ctx.profile(True)
            ms = timeit(run, reps=2, warm=0)
            prof = ctx.profile_read(); ctx.profile(False)
            byt = 12 * nnz + 4 * n + 8 * 8 * n * bw      # matrix + 8 panel passes
            out["lap3d_256_block%d" % bw] = dict(ms_per_iter=ms / m, it_per_s=m / ms * 1e3, gbs=byt / (ms / m) / 1e6, frac=byt / (ms / m) / 1e6 / PEAK,
                                                 classes={k: dict(n=v[0], ms=round(v[1] / 2 / m, 4), gbs=round(v[2] / max(v[1], 1e-9) / 1e6, 1)) for k, v in prof.items() if v[0]})
            A.close(); del B
    if "maxwell" in which:
        # the reference's own operator (Maxwell N = 160, ELL4 + CSR shadow), warm: block b = 4, 8, 16 and the vector path
        A = lz.Matrix.maxwell(ctx, 160); n = A.n_rows
        for bw in (4, 8, 16):
            m = 8
            B = torch.empty(n * bw, dtype=torch.float64, device="cuda")
            lz.check(lz.lib().lz_gen_start_block(ctx.h, n, bw, n, 0x5EED, B.data_ptr()))
            al = torch.zeros(m * bw * bw, dtype=torch.float64, device="cuda"); be = torch.zeros((m + 1) * bw * bw, dtype=torch.float64, device="cuda")
            run = lambda: lz.block_lanczos(ctx, A, B, n, bw, m, al, be, None, lc=-1)
            run(); ctx.sync()
            ctx.profile(True)
            ms = timeit(run, reps=2, warm=0)
            prof = ctx.profile_read(); ctx.profile(False)
            out["maxwell160_block%d" % bw] = dict(ms_per_iter=ms / m, it_per_s=m / ms * 1e3,
                                                 classes={k: dict(n=v[0], ms=round(v[1] / 2 / m, 4), gbs=round(v[2] / max(v[1], 1e-9) / 1e6, 1)) for k, v in prof.items() if v[0]})
            del B
        x = torch.empty(n, dtype=torch.float64, device="cuda")
        lz.check(lz.lib().lz_gen_start_vector(ctx.h, n, 0x5EED, x.data_ptr()))
        m = 50
        al = torch.empty(m, dtype=torch.float64, device="cuda"); be = torch.empty_like(al)
        ms = timeit(lambda: lz.vector_lanczos_async(ctx, A, x, m, al, be), reps=3, warm=1) / m
        out["maxwell160_vector"] = dict(ms_per_iter=ms, it_per_s=1e3 / ms)
        A.close()
    if "spmmcm" in which:
        # the drop-in lz_spmm (column-major, reference layout) on 256^3, b = 16
        A = lz.Matrix.laplacian3d(ctx, 256, 256, 256); n, nnz = A.n_rows, A.nnz
        for bw in (4, 16):
            X = torch.empty(n * bw, dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
            lz.check(lz.lib().lz_gen_start_block(ctx.h, n, bw, n, 0x5EED, X.data_ptr()))
            ms = timeit(lambda: lz.spmm(ctx, A, bw, X, n, Y, n), reps=5, warm=2)
            byt = 12 * nnz + 4 * n + 16 * n * bw
            out["lap3d_256_lz_spmm_b%d" % bw] = dict(ms=ms, gbs=byt / ms / 1e6, frac=byt / ms / 1e6 / PEAK)
            del X, Y
        A.close()
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "devbench.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
