#!/bin/bash
# round-2 GPU call A (1 GPU): the whole -m gpu suite, smoke(), the bench line with its extra legs.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 2>&1 | tail -60 ) > gpurun_out/a_pytest.log 2>&1
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ; echo "smoke rc=$?" ) > gpurun_out/a_smoke.log 2>&1
( timeout 900 python bench.py --steps 3 --warmup 3 ; echo "bench rc=$?" ) > gpurun_out/a_bench.log 2>&1
( timeout 300 python tools/devbench.py block ; echo "devbench rc=$?" ) > gpurun_out/a_devbench.log 2>&1
tail -5 gpurun_out/a_pytest.log; tail -3 gpurun_out/a_smoke.log; tail -c 3000 gpurun_out/a_bench.log
