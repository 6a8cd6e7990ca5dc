#!/bin/bash
# run P: slim entry loop of the operand-staging SpMM: parity, copy-only floor, widths, Maxwell
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("$1", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
}
( timeout 600 python -m pytest tests/test_gpu_block.py tests/test_application_path.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/p_pytest.log 2>&1
tail -4 gpurun_out/p_pytest.log
{
for env in "LZ_NO_XS=" "LZ_SPMM_HINT=32" "LZ_NO_XS=1" "LZ_XS_TILE=384"; do
  env $env timeout 300 python tools/run_configs.py cfg3 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "cfg3 $env"
done
for env in "LZ_NO_XS=1" "LZ_NO_XS="; do
  env $env LZ_BLOCK_WIDTHS=4,8,32 timeout 300 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
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
done
} 2>&1 | tee gpurun_out/p_sweeps.log
