#!/bin/bash
# full verification as the driver runs it: the whole -m gpu suite, smoke(), the bench line; plus config 5 on one GPU
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1800 python -m pytest tests -m gpu -x -q --durations=12 2>&1 | tail -40 ) > gpurun_out/x_pytest.log 2>&1
tail -18 gpurun_out/x_pytest.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ; echo "smoke rc=$?" ) > gpurun_out/x_smoke.log 2>&1
tail -3 gpurun_out/x_smoke.log
( timeout 900 python bench.py --steps 3 --warmup 3 ; echo "bench rc=$?" ) > gpurun_out/x_bench.log 2>&1
tail -c 600 gpurun_out/x_bench.log
( timeout 600 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu --workload cfg5 ; echo "bench rc=$?" ) > gpurun_out/x_bencx_cfg5.log 2>&1
grep "^{" gpurun_out/x_bencx_cfg5.log | cut -c1-200
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ; echo "ref rc=$?" ) > gpurun_out/x_bencx_ref.log 2>&1
tail -c 400 gpurun_out/x_bencx_ref.log
