#!/bin/bash
for v in ${VARIANTS:-0 3 20}; do   # 0 default, 3 coarse schedule everywhere, 20 fine schedule everywhere
  LZ_SPMV_VARIANT=$v timeout 200 python tools/run_configs.py cfg3v cfg2v > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("variant $v", {k:(round(v["ms_per_iter"],4), round(v["step_frac"],3), round(v["classes"]["spmv"]["ms"]/v["classes"]["spmv"]["launches"],4)) for k,v in d.items()})
PY
done
