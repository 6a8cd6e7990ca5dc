#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python tools/devbench.py maxwell 2>&1 | tail -60 > gpurun_out/e_maxwell.log
python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
for k,v in d.items(): print(k, round(v["it_per_s"],1), v.get("classes"))
PY
for r in 1 0; do
  for sl in 8 32; do
  LZ_REORDER=$r LZ_SPMM_SLICE=$sl timeout 400 python tools/run_configs.py cfg4 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("LZ_REORDER=$r LZ_SPMM_SLICE=$sl", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
  done
done 2>&1 | tee gpurun_out/e_rmat.log
bash tools/profile_round2.sh r02 2>&1 | tail -30
