#!/bin/bash
# Evidence for profiles/: launch list of the bench command, full captures of the top kernels,
# fp64 GEMM peak.  Run under gpurun on ONE GPU; every ncu step follows a plain run of the same command.
R=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
timeout 300 $CMD > gpurun_out/${R}_bench_plain.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/${R}_bench_plain.log; exit 1; }
tail -1 gpurun_out/${R}_bench_plain.log | cut -c1-400
# launch list of the timed solve (launches 2407.. of the process: warm-up solve first)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2405 -c 2410 --csv --log-file gpurun_out/${R}_bench_launches.csv $CMD > gpurun_out/${R}_ncu_launches.log 2>&1
echo "launch list rows: $(wc -l < gpurun_out/${R}_bench_launches.csv)"
# full captures: one late launch of each dominant kernel
for k in k_cgs_update k_cgs_project k_csr_spmv_ws; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 250 -c 1 -o gpurun_out/${R}_full_$k $CMD > gpurun_out/${R}_ncu_$k.log 2>&1
  echo "$k: $(tail -1 gpurun_out/${R}_ncu_$k.log)"
done
# fp64 dense peak (denominator of the tensor roofline): cuBLAS DGEMM 8192^3, burst and sustained
timeout 200 python - <<'PY' | tee gpurun_out/${R}_fp64_peak.json
import torch, json, time
n=8192
a=torch.randn(n,n,dtype=torch.float64,device="cuda"); b=torch.randn(n,n,dtype=torch.float64,device="cuda")
for _ in range(2): torch.matmul(a,b)
torch.cuda.synchronize()
best=1e9
for _ in range(5):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a,b); e1.record(); torch.cuda.synchronize(); best=min(best,e0.elapsed_time(e1))
t0=time.time(); cnt=0
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
while time.time()-t0<4.0:
    torch.matmul(a,b); cnt+=1
    if cnt%8==0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
print(json.dumps({"fp64_tflops_burst": 2*n**3/best/1e9, "fp64_tflops_sustained": 2*n**3*cnt/e0.elapsed_time(e1)/1e9, "how": "torch.matmul fp64 8192^3 (cuBLAS DGEMM), best of 5 / 4 s loop"}))
PY
