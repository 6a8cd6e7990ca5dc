#!/bin/bash
# run Y (N GPUs): the default bench line at N ranks, with its cfg5 leg
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 bench.py --gpus $N --steps 3 --warmup 3 2>&1 | grep -v "^W1018\|OMP_NUM_THREADS\|^\*\*\*" | tail -60 ) > gpurun_out/y_bench_n$N.log 2>&1
python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/y_bench_n$N.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("N=$N:", round(d["value"],1), "it/s", "comm", d.get("comm",{}).get("mode"), "e2e", d.get("e2e"), "parity", d.get("parity",{}).get("max_rel_alpha"), "cfg5", d.get("extra",{}).get("cfg5"))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/y_bench_n$N.log").read()[-2500:])
PY
