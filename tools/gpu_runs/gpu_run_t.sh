#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("$1", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
}
{
for env in "LZ_X=1" "LZ_SPMM_HINT=32" "LZ_XS_BOX=16,4,2" "LZ_XS_BOX=64,1,2" "LZ_XS_BOX=32,1,4"; do
  env $env timeout 300 python tools/run_configs.py cfg3 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "cfg3 $env"
done
} 2>&1 | tee gpurun_out/t_sweeps.log
( timeout 600 python -m pytest tests/test_gpu_block.py -m gpu -x -q 2>&1 | tail -5 ) > gpurun_out/t_pytest.log 2>&1
tail -3 gpurun_out/t_pytest.log
