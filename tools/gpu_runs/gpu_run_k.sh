#!/bin/bash
# ncu of the R-MAT scale-24 SpMM (b = 32) and SpMV: L2 hit rate, DRAM bytes, what the gather rate is bound by
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for k in k_spmm_ws k_csr_spmv_ws; do
  timeout 900 ncu --set full --clock-control none -k "regex:$k" -s 3 -c 1 -f -o gpurun_out/r02_full_rmat_$k python tools/run_configs.py cfg4 > gpurun_out/r02_ncu_rmat_$k.log 2>&1
  ncu -i gpurun_out/r02_full_rmat_$k.ncu-rep --page raw --csv > gpurun_out/r02_full_rmat_$k.raw.csv 2>/dev/null
  rm -f gpurun_out/r02_full_rmat_$k.ncu-rep
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/r02_full_rmat_$k.raw.csv")))
h,u,v=rows[0],rows[1],rows[2]
for key in ("gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sector_hit_rate.pct","lts__t_sectors_srcunit_tex_op_read.sum","lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum","lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum","l1tex__t_sector_hit_rate.pct","sm__warps_active.avg.pct_of_peak_sustained_active","smsp__issue_active.avg.pct_of_peak_sustained_active","lts__throughput.avg.pct_of_peak_sustained_elapsed","dram__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__throughput.avg.pct_of_peak_sustained_elapsed","smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","lts__t_requests_srcunit_tex_op_read.sum"):
    if key in h: print("$k", key, v[h.index(key)], u[h.index(key)])
PY
done 2>&1 | tee gpurun_out/k_rmat_ncu.log
