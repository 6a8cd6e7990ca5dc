#!/bin/bash
# Evidence for profiles/ (round 2): launch list of the bench command, full captures of the dominant kernels of the
# vector path (bench command) and of the block path (tools/run_configs.py), SASS comes from the .so in the dev container.
# Run under gpurun on ONE GPU; every ncu step follows a plain run of the same command.
R=${1:-r02}
mkdir -p gpurun_out
KEEP="k_spmm_ws_fused_gram k_cgs_update_project k_block_project_w"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-extra"
timeout 300 $CMD > gpurun_out/${R}_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${R}_bench_plain.log; exit 1; }
tail -1 gpurun_out/${R}_bench_plain.log | cut -c1-300
# launch list of the whole command (warm-up solve + timed solve: 2 x (300 steps x 6 launches + 4) plus the generators)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4200 --csv --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/${R}_bench_launches.csv)"
full() {  # name, kernel regex, launches to skip, command...
  local name=$1 k=$2 s=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$k" -s $s -c 1 -f -o gpurun_out/${R}_full_$name "$@" > gpurun_out/${R}_ncu_$name.log 2>&1
  echo "$name: $(grep -E "==PROF==|Error" gpurun_out/${R}_ncu_$name.log | tail -1)"
  # gpurun brings back at most 64 MiB: export the pages that are read here and keep only the reports listed in KEEP
  ncu -i gpurun_out/${R}_full_$name.ncu-rep --page raw --csv > gpurun_out/${R}_full_$name.raw.csv 2>/dev/null
  case " $KEEP " in *" $name "*) ncu -i gpurun_out/${R}_full_$name.ncu-rep --page source --csv > gpurun_out/${R}_full_$name.source.csv 2>/dev/null ;;
                    *) rm -f gpurun_out/${R}_full_$name.ncu-rep ;; esac
}
for k in k_cgs_update_project k_cgs_update k_cgs_project k_csr_spmv_ws; do
  full $k "^${k}\$" 250 $CMD
done
# block path, 256^3 b = 16.  Shipped: SpMM with the fused subtraction AND Gram epilogue + one panel pass;
# LZ_NO_SPMM_GRAM=1: fused subtraction + separate Gram; LZ_NO_SPMM_FUSE=1: plain SpMM + two-Gram pass + two-term panel pass
python tools/run_configs.py cfg3 > gpurun_out/${R}_cfg3_plain.log 2>&1
full k_spmm_ws_fused_gram "k_spmm_ws" 6 python tools/run_configs.py cfg3
full k_panel_dmma "k_panel_dmma" 13 python tools/run_configs.py cfg3
LZ_NO_SPMM_GRAM=1 python tools/run_configs.py cfg3 > gpurun_out/${R}_cfg3_nogram_plain.log 2>&1
LZ_NO_SPMM_GRAM=1 full k_spmm_ws_fused "k_spmm_ws" 6 python tools/run_configs.py cfg3
LZ_NO_SPMM_GRAM=1 full k_gram_dmma "k_gram_dmma" 6 python tools/run_configs.py cfg3
LZ_NO_SPMM_FUSE=1 python tools/run_configs.py cfg3 > gpurun_out/${R}_cfg3_nofuse_plain.log 2>&1
LZ_NO_SPMM_FUSE=1 full k_spmm_ws_plain "k_spmm_ws" 6 python tools/run_configs.py cfg3
LZ_NO_SPMM_FUSE=1 full k_panel2_dmma "k_panel2_dmma" 6 python tools/run_configs.py cfg3
LZ_NO_SPMM_FUSE=1 full k_gram2_dmma "k_gram2_dmma" 6 python tools/run_configs.py cfg3
python tools/run_configs.py cfg3r > gpurun_out/${R}_cfg3r_plain.log 2>&1
full k_block_project_w "k_block_project_w" 30 python tools/run_configs.py cfg3r
full k_block_update_w "k_block_update_w" 30 python tools/run_configs.py cfg3r
# 3-D fused pass A (north_star: 256^3 step >= 80 % of HBM)
full k_csr_spmv_ws_3d "k_csr_spmv_ws" 50 python tools/run_configs.py cfg3v
ls -la gpurun_out/${R}_full_* | awk '{print $5, $9}'; du -sh gpurun_out
