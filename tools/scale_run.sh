#!/bin/bash
# strong-scaling record: workload $1 at N = 8,4,2,1 on one box (gpurun --gpus 8)
W=${1:-cfg5}
mkdir -p gpurun_out
: > gpurun_out/scale_$W.jsonl
P=29800
for N in ${NS:-8 4 2 1}; do
  P=$((P+1))
  if [ "$N" = "1" ]; then
    timeout 240 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | grep '^{' >> gpurun_out/scale_$W.jsonl
  else
    timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --workload $W --steps 3 --warmup 3 2>&1 | grep '^{' >> gpurun_out/scale_$W.jsonl
  fi
  echo "N=$N rc=$?"
done
python - <<PY
import json
rows=[json.loads(l) for l in open("gpurun_out/scale_$W.jsonl")]
base=[r for r in rows if r["n_gpus"]==1]
for r in rows:
    print(r["n_gpus"], "GPUs:", round(r["value"],1), "it/s", round(r["ms_per_step"],2), "ms/solve", "speedup", round(r["value"]/base[0]["value"],2) if base else None, r["roofline"]["share_of_step"])
PY
