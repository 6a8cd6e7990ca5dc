#!/bin/bash
# round-2 GPU call G (8 GPUs): the bench exactly as the driver launches it at N = 8, plus config 5
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {
  local label=$1; shift
  ( env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $N $BARGS 2>&1 | grep -E "^\{|Error|error|assert|Traceback" | tail -5 ) > gpurun_out/g_bench_${label}_n$N.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/g_bench_${label}_n$N.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("${label} N=$N:", round(d["value"],1), "it/s", "ms/step", round(d["ms_per_step"],1), "comm", d.get("comm",{}).get("mode","")[:12], d.get("comm",{}).get("share_of_profiled_time"),
          "e2e", d.get("e2e",{}).get("value"), "parity", d.get("parity",{}).get("max_rel_alpha"), "share", d["roofline"]["share_of_step"])
except Exception as e:
    print("${label} N=$N: FAILED", e); print(open("gpurun_out/g_bench_${label}_n$N.log").read()[-2500:])
PY
}
BARGS="--steps 3 --warmup 3" run cfg2_peer LZ_DUMMY=1
BARGS="--steps 3 --warmup 1 --no-e2e --no-cpu --workload cfg5" run cfg5_peer LZ_DUMMY=1
BARGS="--steps 3 --warmup 1 --no-e2e --no-cpu" run cfg2_nccl LZ_COMM=1
