#!/bin/bash
# run AG: ncu --set full of the shipped operand-staging SpMM (final build), cfg3 block step
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 200 python tools/run_configs.py cfg3 > gpurun_out/ag_plain.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_spmm_xs" -s 6 -c 1 -f -o gpurun_out/r02_full_k_spmm_xs python tools/run_configs.py cfg3 > gpurun_out/r02_ncu_k_spmm_xs.log 2>&1
ncu -i gpurun_out/r02_full_k_spmm_xs.ncu-rep --page raw --csv > gpurun_out/r02_full_k_spmm_xs.raw.csv 2>/dev/null
ncu -i gpurun_out/r02_full_k_spmm_xs.ncu-rep --page source --csv > gpurun_out/r02_full_k_spmm_xs.source.csv 2>/dev/null
rm -f gpurun_out/r02_full_k_spmm_xs.ncu-rep
tail -2 gpurun_out/ag_plain.log | cut -c1-300
