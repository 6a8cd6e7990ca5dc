#!/bin/bash
# run O: ring depth / chunk size of the operand-staging SpMM, one ncu capture of it
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("$1", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
}
{
for env in "LZ_XS_TILE=384" "LZ_XS_TILE=256" "LZ_XS_TILE=192" "LZ_XS_TILE=256 LZ_XS_STAGES=3" "LZ_XS_TILE=256 LZ_XS_STAGES=4"; do
  env $env timeout 300 python tools/run_configs.py cfg3 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "cfg3 $env"
done
} 2>&1 | tee gpurun_out/o_sweeps.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_spmm_xs" -s 6 -c 1 -f -o gpurun_out/r02_full_k_spmm_xs python tools/run_configs.py cfg3 > gpurun_out/r02_ncu_k_spmm_xs.log 2>&1
ncu -i gpurun_out/r02_full_k_spmm_xs.ncu-rep --page raw --csv > gpurun_out/r02_full_k_spmm_xs.raw.csv 2>/dev/null
ncu -i gpurun_out/r02_full_k_spmm_xs.ncu-rep --page source --csv > gpurun_out/r02_full_k_spmm_xs.source.csv 2>/dev/null
rm -f gpurun_out/r02_full_k_spmm_xs.ncu-rep
ls -la gpurun_out/r02_full_k_spmm_xs*
