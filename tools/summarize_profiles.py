"""Turns the ncu output of tools/profile_round.sh (gpurun_out/<round>_*) into the tracked evidence under
profiles/: launch-list shares, raw metric pages of the full captures, DRAM traffic per algorithmic byte
(profiles/roofline_traffic.json, read by bench.py) and SASS excerpts.  Runs in the dev container (no GPU)."""
import csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
N = 4096 * 4096
NNZ = 5 * N - 4 * 4096

def launches():
    src = os.path.join(GO, R + "_bench_launches.csv")
    if not os.path.exists(src):
        return None
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]; ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    data = [(re.sub(r"[<(].*", "", r[ik]), float(r[iv].replace(",", ""))) for r in rows[1:] if r[iv]]
    per_solve = int(os.environ.get("LZ_LAUNCHES_PER_SOLVE", "2107" if R == "r01" else "1866"))
    half = data[-per_solve:]        # the timed solve (r01: 300 steps x 7 launches + 7; r02: 63 x 8 + 237 x 6 + 10 with pass B folded into CGS2)
    tot = sum(v for _, v in half); by = {}
    for k, v in half:
        by.setdefault(k, [0, 0.0]); by[k][0] += 1; by[k][1] += v
    shutil.copy(src, os.path.join(PR, R + "_bench_launches.csv"))
    return {k: dict(launches=c, ms=v / 1e6, share=v / tot) for k, (c, v) in sorted(by.items(), key=lambda kv: -kv[1][1])}, len(data)

def full(kernel):
    rep = os.path.join(GO, "%s_full_%s.ncu-rep" % (R, kernel))
    pre = os.path.join(GO, "%s_full_%s.raw.csv" % (R, kernel))          # exported on the GPU box (gpurun brings back <= 64 MiB)
    if os.path.exists(pre):
        out = open(pre).read()
    elif os.path.exists(rep):
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        return None
    open(os.path.join(PR, "%s_full_%s.raw.csv" % (R, kernel)), "w").write(out)
    rows = list(csv.reader(out.splitlines())); hdr, units, val = rows[0], rows[1], rows[2]
    def get(name):
        i = hdr.index(name); v = float(val[i].replace(",", "")); u = units[i].split("/")[0]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "ns": 1e-9}.get(u, 1)
    smem = get("launch__shared_mem_per_block_dynamic")
    d = dict(dram_bytes=get("dram__bytes_read.sum") + get("dram__bytes_write.sum"), seconds=get("gpu__time_duration.sum"),
             dyn_smem=smem,
             regs=get("launch__registers_per_thread"), grid=get("launch__grid_size"), block=get("launch__block_size"))
    if kernel == "k_cgs_update":
        K = int(round(smem / 8)); alg = 8.0 * N * (K + 2)
    elif kernel == "k_cgs_project":
        K = int(round(smem / 64)); alg = 8.0 * N * (K + 1)
    elif kernel == "k_cgs_update_project":
        # cgs_fused_smem(K, NB) (lz_vector.cu): NB tiles of K x 32, NB w slices, 256 partials, K coefficients, barriers;
        # NB is 1 or 2 depending on the shape the host picked for this K
        K = next(k for k in range(1, 1024) for nb in (1, 2) if 8 * (nb * k * 32 + nb * 32 + 256 + ((k + 1) & ~1)) + 16 == int(smem)); alg = 8.0 * N * (K + 2)
    else:
        K = None; alg = 12.0 * NNZ + 36.0 * N
    d.update(captured_K=K, algorithmic_bytes=alg, dram_bytes_per_algorithmic_byte=d["dram_bytes"] / alg,
             dram_gbs_under_ncu=d["dram_bytes"] / d["seconds"] / 1e9)
    return d

def sass(kernel_regex, name):
    lib = os.path.join(ROOT, "gpu-implementation-of-signle-and-block-lanczos_b200", "liblanczos_b200.so")
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = out.split("Function : ")
    for b in blocks[1:]:
        if re.match(kernel_regex, b):
            ops = {}
            for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", b, re.M):
                ops[m.group(1)] = ops.get(m.group(1), 0) + 1
            keep = {k: v for k, v in ops.items() if re.match(r"UBLKCP|SYNCS|DMMA|LDG\.E\.(ENL2\.)?(128|256|64)|LDG.*256|STG.*(128|256)|LDS|DFMA|UTMA", k)}
            open(os.path.join(PR, "%s_sass_%s.txt" % (R, name)), "w").write(
                "Function : " + b.split("\n")[0] + "\nopcode histogram (selected): " + json.dumps(keep, sort_keys=True) + "\n\n" + "\n".join(b.split("\n")[:400]))
            return keep
    return None

def block_kernels():
    """round 2: the block-path captures (256^3, b = 16) -> one table: time, DRAM bytes, hit rates, pipe utilisation"""
    names = ["k_spmm_ws_plain", "k_spmm_ws_fused", "k_spmm_ws_fused_gram", "k_spmm_xs", "rmat_k_spmm_ws", "rmat_k_csr_spmv_ws", "k_gram_dmma", "k_gram2_dmma", "k_panel_dmma", "k_panel2_dmma",
             "k_block_project_w", "k_block_update_w", "k_csr_spmv_ws_3d"]
    cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
            ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sector_hit_rate.pct", "l1_hit_pct"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
            ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm_read"),
            ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1_pct"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2_pct"),
            ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_active_pct"),
            ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64_pipe_active_pct"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_scoreboard"),
            ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
    out = {}
    for nme in names:
        src = os.path.join(GO, "%s_full_%s.raw.csv" % (R, nme))
        if not os.path.exists(src):
            continue
        shutil.copy(src, os.path.join(PR, "%s_full_%s.raw.csv" % (R, nme)))
        rows = list(csv.reader(open(src)))
        if len(rows) < 3:
            continue
        hdr, units, val = rows[0], rows[1], rows[2]
        rec = {"kernel": val[hdr.index("Kernel Name")][:90]}
        for key, short in cols:
            if key in hdr:
                i = hdr.index(key)
                rec[short] = "%s %s" % (val[i], units[i]) if units[i] not in ("", "%") else float(val[i].replace(",", ""))
        out[nme] = rec
    return out


def main():
    res = {}
    l = launches()
    if l:
        res["launch_shares"], res["launch_rows"] = l
    traffic = {}
    for kern, cls in (("k_cgs_update_project", "cgs_update_project"), ("k_cgs_update", "cgs_update"), ("k_cgs_project", "cgs_project"), ("k_csr_spmv_ws", "spmv")):
        f = full(kern)
        if f:
            traffic[cls] = f
    if traffic:
        traffic["_source"] = "ncu --set full captures profiles/%s_full_*.raw.csv (dram__bytes_read.sum + dram__bytes_write.sum), bench.py cfg2, launch 250 of each kernel" % R
        json.dump(traffic, open(os.path.join(PR, "roofline_traffic.json"), "w"), indent=1)
    res["traffic"] = traffic
    res["block_kernels"] = block_kernels()
    res["sass"] = {n: sass(rx, n) for rx, n in ((r"_Z20k_cgs_update_project", "k_cgs_update_project"), (r"_Z12k_cgs_update", "k_cgs_update"),
                                               (r"_Z13k_cgs_project", "k_cgs_project"), (r"_Z13k_csr_spmv_wsILi1ELi3ELi5", "k_csr_spmv_ws"),
                                               (r"_Z9k_spmm_wsILi16ELi12ELi2ELi2048ELi2ELb0", "k_spmm_ws16"), (r"_Z9k_spmm_wsILi16ELi11ELi2ELi2048ELi2ELb1ELb1", "k_spmm_ws16_fused_gram"),
                                               (r"_Z9k_spmm_xsILi16ELi15ELb1ELb1", "k_spmm_xs16_fused_gram"), (r"_Z11k_gram_dmmaILi16", "k_gram_dmma16"),
                                               (r"_Z17k_block_project_wILi16", "k_block_project_w16"), (r"_Z16k_block_update_wILi16", "k_block_update_w16"),
                                               (r"_Z13k_panel2_dmmaILi16ELb1", "k_panel2_dmma16"), (r"_Z14k_basis_rotateILi8", "k_basis_rotate8"),
                                               (r"_Z16k_peer_allreduce", "k_peer_allreduce"))}
    for f in os.listdir(GO):        # source pages of the kept reports, run logs
        if f.startswith(R + "_full_") and f.endswith(".source.csv"):
            shutil.copy(os.path.join(GO, f), os.path.join(PR, f))
    json.dump(res, open(os.path.join(PR, R + "_summary.json"), "w"), indent=1)
    for f in (R + "_bench_plain.log",):
        if os.path.exists(os.path.join(GO, f)):
            shutil.copy(os.path.join(GO, f), os.path.join(PR, f))
    print(json.dumps(res, indent=1))

if __name__ == "__main__":
    main()
