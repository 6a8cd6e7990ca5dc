#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for L in 256 64 32 16; do
  LZ_SPLIT_L=$L timeout 400 python tools/run_configs.py cfg4 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("LZ_SPLIT_L=$L", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
done 2>&1 | tee gpurun_out/l_rmat_split.log
( timeout 600 python -m pytest tests/test_gpu_block.py -m gpu -q -x -k "rmat" 2>&1 | tail -3 )
( LZ_SPLIT_L=32 timeout 600 python -m pytest tests/test_gpu_block.py tests/test_gpu_single.py -m gpu -q -x -k "rmat or split or long" 2>&1 | tail -3 )
