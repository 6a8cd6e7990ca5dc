#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("$1", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
}
( timeout 600 python -m pytest tests/test_gpu_block.py tests/test_application_path.py -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s_pytest.log 2>&1
tail -4 gpurun_out/s_pytest.log
{
for env in "LZ_X=1" "LZ_SPMM_HINT=32" "LZ_XS_NO_TILES=1" "LZ_XS_NO_TILES=1 LZ_SPMM_HINT=32" "LZ_NO_XS=1" "LZ_XS_BOX=32,2,2" "LZ_XS_BOX=16,2,2" "LZ_XS_BOX=16,2,2 LZ_SPMM_HINT=32" "LZ_XS_BOX=8,4,4" "LZ_XS_BOX=16,4,1"; do
  env $env timeout 300 python tools/run_configs.py cfg3 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "cfg3 $env"
done
} 2>&1 | tee gpurun_out/s_sweeps.log
