#!/bin/bash
# run V: operand-staging SpMM across widths / operators, against the gathering kernel; ncu capture
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
for env in "LZ_NO_XS=1" "LZ_X=1"; do
  env $env LZ_BLOCK_WIDTHS=4,8,16,32 timeout 300 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("widths $env", {k:(round(v["it_per_s"],1), {c:x["ms"] for c,x in v["classes"].items()}) for k,v in d.items()})
PY
  env $env timeout 300 python tools/devbench.py maxwell > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("maxwell $env", {k:(round(v["it_per_s"],1), {c:x["ms"] for c,x in v.get("classes",{}).items()}) for k,v in d.items()})
PY
  env $env timeout 300 python tools/run_configs.py cfg3r > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("cfg3r $env", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
done
} 2>&1 | tee gpurun_out/v_sweeps.log
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:k_spmm_xs" -s 6 -c 1 -f -o gpurun_out/r02_full_k_spmm_xs python tools/run_configs.py cfg3 > gpurun_out/r02_ncu_k_spmm_xs.log 2>&1
ncu -i gpurun_out/r02_full_k_spmm_xs.ncu-rep --page raw --csv > gpurun_out/r02_full_k_spmm_xs.raw.csv 2>/dev/null
ncu -i gpurun_out/r02_full_k_spmm_xs.ncu-rep --page source --csv > gpurun_out/r02_full_k_spmm_xs.source.csv 2>/dev/null
rm -f gpurun_out/r02_full_k_spmm_xs.ncu-rep
