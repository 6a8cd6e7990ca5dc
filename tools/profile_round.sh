#!/bin/bash
# Evidence for profiles/: launch list of the bench command, full captures of the top kernels,
# fp64 GEMM peak.  Run under gpurun on ONE GPU; every ncu step follows a plain run of the same command.
# Afterwards (in the dev container): python tools/summarize_profiles.py <round>
R=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 300 $CMD > gpurun_out/${R}_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${R}_bench_plain.log; exit 1; }
tail -1 gpurun_out/${R}_bench_plain.log | cut -c1-400
# launch list of the timed solve (the warm-up solve's 2107 launches are skipped: same kernels, same sizes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2107 -c 2400 --csv --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/${R}_bench_launches.csv)"
# full captures: one late launch (step 250 of the warm-up solve) of each dominant kernel
for k in k_cgs_update_project k_cgs_update k_cgs_project k_csr_spmv_ws; do
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^${k}\$" -s 250 -c 1 -f -o gpurun_out/${R}_full_$k $CMD > gpurun_out/${R}_ncu_$k.log 2>&1
  echo "$k: $(tail -1 gpurun_out/${R}_ncu_$k.log)"
done
