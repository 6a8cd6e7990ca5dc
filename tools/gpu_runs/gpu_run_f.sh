#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gpu_paths.py tests/test_gpu_block.py tests/test_gpu_eigs.py -m gpu -q -x 2>&1 | tail -30 ) > gpurun_out/f_pytest.log 2>&1
tail -8 gpurun_out/f_pytest.log
show() {
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("$1", {k:(round(v["it_per_s"],1), {c:x["ms"] for c,x in v["classes"].items()}) for k,v in d.items()})
PY
}
for cfg in "LZ_DUMMY=1" "LZ_NO_SPMM_GRAM=1" "LZ_NO_SPMM_FUSE=1"; do
  env $cfg LZ_BLOCK_WIDTHS=16 timeout 200 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  show "$cfg"
done 2>&1 | tee gpurun_out/f_block16_variants.log
timeout 300 python tools/devbench.py maxwell > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
show maxwell | tee -a gpurun_out/f_block16_variants.log
( timeout 900 python bench.py --steps 3 --warmup 3 ; echo "bench rc=$?" ) > gpurun_out/f_bench.log 2>&1
tail -c 2500 gpurun_out/f_bench.log
bash tools/profile_round2.sh r02 2>&1 | tail -30
