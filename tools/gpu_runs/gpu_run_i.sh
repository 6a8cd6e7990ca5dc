#!/bin/bash
# N = 8 variants of the headline (A/B of the small-shard choices)
N=${1:-8}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() {
  local label=$1; shift
  ( env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 \
      bench.py --gpus $N --steps 2 --warmup 1 --no-e2e --no-cpu 2>&1 | grep -E "^\{|Error|error|assert|Traceback" | tail -5 ) > gpurun_out/i_bench_${label}_n$N.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/i_bench_${label}_n$N.log") if x.startswith("{")][-1]; d=json.loads(l); r=d["roofline"]
    per={k: round(v*r["profiled_ms"]*1e3/(d["steps"]*300),1) for k,v in r["share_of_step"].items()}
    print("${label} N=$N:", round(d["value"],1), "it/s", "us/iter", per, "unprofiled", round((r["timed_ms"]-r["profiled_ms"])*1e3/(d["steps"]*300),1))
except Exception as e:
    print("${label} N=$N: FAILED", e); print(open("gpurun_out/i_bench_${label}_n$N.log").read()[-1500:])
PY
}
run default LZ_DUMMY=1
run rpt8 LZ_CGS_RPT=8
run updmult2 LZ_CGS_UPD_MULT=2
run nooverlap LZ_NO_OVERLAP=1
run fusemin1 LZ_CGS_FUSE_MIN_K=1
