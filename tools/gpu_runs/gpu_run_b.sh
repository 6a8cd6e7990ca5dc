#!/bin/bash
# round-2 GPU call B (1 GPU): new eigensolver tests + path tests, SpMM L2-policy sweep, drop-in SpMM timing
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gpu_eigs.py tests/test_gpu_paths.py tests/test_lifetime.py tests/test_gpu_block.py -m gpu -q -x --deselect tests/test_gpu_block.py::test_full_size_config3_parity 2>&1 | tail -40 ) > gpurun_out/b_pytest.log 2>&1
tail -15 gpurun_out/b_pytest.log
for h in 0 4 6 7 15 8 14 11; do
  LZ_SPMM_HINT=$h LZ_BLOCK_WIDTHS=16 timeout 200 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("hint $h", {k:(round(v["it_per_s"],1), v["classes"]["spmm"]["ms"], v["classes"].get("gram",{}).get("ms"), v["classes"]["panel"]["ms"]) for k,v in d.items()})
PY
done 2>&1 | tee gpurun_out/b_spmm_hint_sweep.log
for h in 0 2 3 8 10 11; do
  LZ_SPMM_HINT=$h LZ_BLOCK_WIDTHS=8,32 timeout 300 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("hint $h", {k:(round(v["it_per_s"],1), v["classes"]["spmm"]["ms"], v["classes"].get("gram",{}).get("ms"), v["classes"]["panel"]["ms"]) for k,v in d.items()})
PY
done 2>&1 | tee gpurun_out/b_spmm_hint_sweep_8_32.log
timeout 200 python tools/devbench.py spmmcm 2>&1 | tail -5 | tee gpurun_out/b_spmm_cm.log
for r in 1 0; do
  LZ_REORDER=$r timeout 400 python tools/run_configs.py cfg4 > /tmp/o.log 2>&1 || tail -5 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/configs.json"))
print("LZ_REORDER=$r", {k:(round(v.get("it_per_s",0),2), {c:round(x["ms"]/max(x["launches"],1),3) for c,x in v.get("classes",{}).items()}) for k,v in d.items() if "classes" in v})
PY
done 2>&1 | tee gpurun_out/b_rmat.log
