#!/bin/bash
# run N: operand-staging SpMM (lz_spmm_xs.cuh): parity first, then timing against the gathering kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("$1", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
}
( timeout 600 python -m pytest tests/test_gpu_block.py tests/test_application_path.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/n_pytest.log 2>&1
tail -6 gpurun_out/n_pytest.log
{
for env in "LZ_NO_XS=1" "LZ_NO_XS=" "LZ_XS_STAGES=2" "LZ_SPMM_HINT=1"; do
  env $env timeout 300 python tools/run_configs.py cfg3 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "cfg3 $env"
done
env LZ_NO_XS=1 timeout 300 python tools/run_configs.py cfg3r > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "cfg3r LZ_NO_XS=1"
timeout 300 python tools/run_configs.py cfg3r > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "cfg3r xs"
} 2>&1 | tee gpurun_out/n_sweeps.log
