#!/bin/bash
# run W (2 GPUs): sharded parity in every comm / reorth mode, then the default bench line at 2 ranks (with its cfg5 leg)
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -30 ) > gpurun_out/w_pytest_n$N.log 2>&1
tail -8 gpurun_out/w_pytest_n$N.log
cat gpurun_out/test_gpu_multi_world*.log 2>/dev/null | grep -E "mode|ok" | head -40
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 2 --warmup 3 2>&1 | grep -v "^W1018\|OMP_NUM_THREADS\|^\*\*\*" | tail -60 ) > gpurun_out/w_bench_n$N.log 2>&1
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/w_bench_n$N.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("N=$N:", round(d["value"],1), "it/s", "comm", d.get("comm",{}).get("mode"), "e2e", d.get("e2e",{}).get("value"), "parity", d.get("parity",{}).get("max_rel_alpha"), "cfg5", d.get("extra",{}).get("cfg5"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/w_bench_n$N.log").read()[-1500:])
PY
