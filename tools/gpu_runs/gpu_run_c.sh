#!/bin/bash
# round-2 GPU call C (N GPUs, default 2): sharded parity across the comm modes, then the bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/c_pytest_n$N.log 2>&1
tail -12 gpurun_out/c_pytest_n$N.log
cat gpurun_out/test_gpu_multi_world*.log 2>/dev/null | grep -E "mode|ok" | head -20
run() {  # label, env..., -- bench args
  local label=$1; shift
  ( env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $N $BARGS 2>&1 | grep -E "^\{|Error|error|assert" | tail -3 ) > gpurun_out/c_bench_${label}_n$N.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/c_bench_${label}_n$N.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("${label} N=$N:", round(d["value"],1), "it/s", "ms/step", round(d["ms_per_step"],1), "comm", d.get("comm",{}).get("mode"), d.get("comm",{}).get("share_of_profiled_time"),
          "e2e", d.get("e2e",{}).get("value"), "parity", d.get("parity",{}).get("max_rel_alpha"), "share", d["roofline"]["share_of_step"])
except Exception as e:
    print("${label} N=$N: FAILED", e); print(open("gpurun_out/c_bench_${label}_n$N.log").read()[-1500:])
PY
}
BARGS="--steps 2 --warmup 1" run cfg2_peer LZ_DUMMY=1
BARGS="--steps 2 --warmup 1 --no-e2e --no-cpu" run cfg2_nccl LZ_COMM=1
BARGS="--steps 2 --warmup 1 --no-e2e --no-cpu" run cfg2_peer_nooverlap LZ_NO_OVERLAP=1
BARGS="--steps 3 --warmup 1 --no-e2e --no-cpu --workload cfg5" run cfg5_peer LZ_DUMMY=1
BARGS="--steps 3 --warmup 1 --no-e2e --no-cpu --workload cfg5" run cfg5_nccl LZ_COMM=1
BARGS="--steps 3 --warmup 1 --no-e2e --no-cpu --workload cfg5" run cfg5_peer_nooverlap LZ_NO_OVERLAP=1
