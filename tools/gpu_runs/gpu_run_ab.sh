#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python tools/run_configs.py cfg4 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
python - <<PY | tee gpurun_out/ab_rmat.log
import json
d=json.load(open("gpurun_out/configs.json"))
print("rmat", {k:(round(v.get("it_per_s",0),2), {c:(x["launches"], round(x["ms"]/max(x["launches"],1),3)) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
( timeout 600 python -m pytest tests/test_gpu_block.py tests/test_gpu_single.py -m gpu -x -q -k "rmat or split or long" 2>&1 | tail -4 )
