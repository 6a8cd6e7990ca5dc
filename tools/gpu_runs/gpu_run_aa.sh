#!/bin/bash
# run AA: R-MAT with the faster ordered combine, then the full verification as the driver runs it
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python tools/run_configs.py cfg4 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
python - <<PY | tee gpurun_out/aa_rmat.log
import json
d=json.load(open("gpurun_out/configs.json"))
print("rmat", {k:(round(v.get("it_per_s",0),2), {c:(x["launches"], round(x["ms"]/max(x["launches"],1),3)) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
( timeout 1800 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -30 ) > gpurun_out/aa_pytest.log 2>&1
tail -12 gpurun_out/aa_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ; echo "smoke rc=$?" ) > gpurun_out/aa_smoke.log 2>&1
tail -3 gpurun_out/aa_smoke.log
( timeout 900 python bench.py --steps 3 --warmup 3 ; echo "bench rc=$?" ) > gpurun_out/aa_bench.log 2>&1
tail -c 400 gpurun_out/aa_bench.log
