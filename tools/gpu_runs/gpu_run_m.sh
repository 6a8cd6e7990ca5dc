#!/bin/bash
# run M: SpMM split of its own (L=32) on R-MAT, trip rotation A/B on the 3-D block step, then full verification
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("$1", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
}
{
timeout 400 python tools/run_configs.py cfg4 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "rmat default(256/32)"
LZ_SPLIT_L_MM=48 timeout 400 python tools/run_configs.py cfg4 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "rmat mm=48"
LZ_SPLIT_L_MM=24 timeout 400 python tools/run_configs.py cfg4 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "rmat mm=24"
for h in 0 16 0 16; do
LZ_SPMM_HINT=$h timeout 400 python tools/run_configs.py cfg3 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log; show "cfg3 hint=$h"
done
} 2>&1 | tee gpurun_out/m_sweeps.log
( timeout 1800 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40 ) > gpurun_out/m_pytest.log 2>&1
tail -18 gpurun_out/m_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ; echo "smoke rc=$?" ) > gpurun_out/m_smoke.log 2>&1
tail -3 gpurun_out/m_smoke.log
( timeout 900 python bench.py --steps 3 --warmup 3 ; echo "bench rc=$?" ) > gpurun_out/m_bench.log 2>&1
tail -c 600 gpurun_out/m_bench.log
