#!/bin/bash
# plain SpMV and fused Lanczos step per kernel variant (dev-time knob LZ_SPMV_VARIANT)
for v in ${VARIANTS:-0 3 20}; do
  LZ_SPMV_VARIANT=$v timeout 200 python tools/devbench.py spmv lanczos > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("variant $v", {k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d.items()})
PY
done
