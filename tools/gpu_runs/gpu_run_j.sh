#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gpu_single.py tests/test_gpu_paths.py tests/test_gpu_block.py tests/test_gpu_eigs.py tests/test_application_path.py -m gpu -q -x \
    --deselect tests/test_gpu_block.py::test_full_size_config3_parity 2>&1 | tail -15 ) > gpurun_out/j_pytest.log 2>&1
tail -6 gpurun_out/j_pytest.log
for cfg in "LZ_DUMMY=1" "LZ_NO_TRANSPOSE=1"; do
  env $cfg timeout 300 python tools/devbench.py spmv lanczos > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("$cfg", {k:(round(v["ms"],4), round(v["frac"],3)) for k,v in d.items() if "ms" in v})
PY
done 2>&1 | tee gpurun_out/j_spmv_transpose.log
show() {
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("$1", {k:(round(v["it_per_s"],1), {c:x["ms"] for c,x in v["classes"].items()}) for k,v in d.items()})
PY
}
for cfg in "LZ_SPMM_SHAPE=0" "LZ_SPMM_SHAPE=1" "LZ_SPMM_SHAPE=2" "LZ_SPMM_SHAPE=0 LZ_NO_SPMM_FUSE=1" "LZ_SPMM_SHAPE=1 LZ_NO_SPMM_FUSE=1" "LZ_SPMM_SHAPE=2 LZ_NO_SPMM_FUSE=1"; do
  env $cfg LZ_BLOCK_WIDTHS=16 timeout 200 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  show "$cfg"
done 2>&1 | tee gpurun_out/j_spmm_shapes.log
( timeout 600 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --workload cfg5 ) 2>&1 | grep "^{" | cut -c1-160
( LZ_NO_TRANSPOSE=1 timeout 600 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --workload cfg5 ) 2>&1 | grep "^{" | cut -c1-160
