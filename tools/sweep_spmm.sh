#!/bin/bash
# SpMM kernel sweep: chunk-map run length (LZ_SPMM_RUN; 0 = one contiguous range per CTA) x kernel shape (LZ_SPMM_SHAPE)
for sh in ${SHAPES:-0}; do   # only shape 0 is compiled in (see profiles/r01_spmv_variants.md for the others)
for r in ${RUNS:-0 1 2 4 8 16}; do
  LZ_SPMM_SHAPE=$sh LZ_SPMM_RUN=$r timeout 200 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("shape $sh run $r", {k:(round(v["it_per_s"],1), v["classes"]["spmm"]) for k,v in d.items()})
PY
done
done
