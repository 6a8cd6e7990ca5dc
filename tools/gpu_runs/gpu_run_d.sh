#!/bin/bash
# round-2 GPU call D (1 GPU): eigensolver / path / Maxwell / block tests, block-step variants, mirror harness on Maxwell N=160
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1200 python -m pytest tests/test_gpu_eigs.py tests/test_gpu_paths.py tests/test_lifetime.py tests/test_gpu_block.py tests/test_gpu_single.py tests/test_host_mirror.py -m gpu -q \
    --deselect tests/test_gpu_block.py::test_full_size_config3_parity --deselect tests/test_gpu_single.py::test_full_size_config2_parity 2>&1 | tail -40 ) > gpurun_out/d_pytest.log 2>&1
tail -12 gpurun_out/d_pytest.log
show() {
  python - <<PY
import json
d=json.load(open("gpurun_out/devbench.json"))
print("$1", {k:(round(v["it_per_s"],1), {c:x["ms"] for c,x in v["classes"].items()}) for k,v in d.items()})
PY
}
for cfg in "LZ_DUMMY=1" "LZ_PANEL_PAD=2064" "LZ_PANEL_PAD=33040" "LZ_NO_SPMM_FUSE=1" "LZ_NO_SPMM_FUSE=1 LZ_PANEL_PAD=2064" "LZ_NO_SPMM_FUSE=1 LZ_SPMM_HINT=2"; do
  env $cfg LZ_BLOCK_WIDTHS=16 timeout 200 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
  show "$cfg"
done 2>&1 | tee gpurun_out/d_block16_variants.log
LZ_BLOCK_WIDTHS=4,8,32 timeout 300 python tools/devbench.py block > /tmp/o.log 2>&1 || tail -3 /tmp/o.log
show "b=4,8,32" | tee -a gpurun_out/d_block16_variants.log
H=gpu-implementation-of-signle-and-block-lanczos_b200/host
{ for nc in 4 8 16; do echo "maxwell N=160 block $nc:"; timeout 600 $H/test_lanczos -N 160 -m 10 --block $nc | grep -E "iterations"; done
  echo "maxwell_dev N=160 vector:"; timeout 600 $H/test_lanczos -N 160 -m 50 --vector --matrix maxwell_dev | grep -E "iterations|elapsed"
  echo "maxwell (host assembly) N=160 vector:"; ( time timeout 600 $H/test_lanczos -N 160 -m 50 --vector | grep -E "iterations|elapsed" ) 2>&1 | grep -E "iterations|elapsed|real"
  echo "maxwell_dev N=160 vector wall:"; ( time timeout 600 $H/test_lanczos -N 160 -m 50 --vector --matrix maxwell_dev > /dev/null ) 2>&1 | grep real
} 2>&1 | tee gpurun_out/d_mirror_maxwell.log
