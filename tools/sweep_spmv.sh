#!/bin/bash
# dev-time sweep of the TMA SpMV variants (env knobs read in lz_ctx_create)
for tile in ${TILES:-1536 3072}; do
  for v in ${VARIANTS:-0 1 2 3}; do
    LZ_SPMV_TILE=$tile LZ_SPMV_VARIANT=$v python tools/devbench.py ${WHAT:-spmv} > /tmp/o.json 2>/tmp/e.txt || tail -3 /tmp/e.txt
    python - <<PY
import json
d=json.load(open('/tmp/o.json'))
print("tile=$tile v=$v", {k: round(v.get('frac', v.get('it_per_s',0)),3) for k,v in d.items()})
PY
  done
done
