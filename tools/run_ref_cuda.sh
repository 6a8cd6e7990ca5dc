#!/bin/bash
# On the GPU box: the reference's own CUDA build (sm_100) beside the new library on the reference's
# own matrix family (Maxwell, ELL width 4).  Writes gpurun_out/ref_cuda_*.{bin,log}.
mkdir -p gpurun_out
R=oracle/_ref
H=gpu-implementation-of-signle-and-block-lanczos_b200/host
N=${1:-160}
{
echo "== reference CUDA (sm_100), Maxwell N=$N"
timeout 600 $R/ref_cuda_dump_4 vector $N 50 gpurun_out/ref_cuda_vector_N$N.bin
for nc in 4 8 16; do timeout 600 $R/ref_cuda_dump_$nc block $N 10 gpurun_out/ref_cuda_block${nc}_N$N.bin; done
echo "== lanczos_b200 harness, same operator and sizes"
timeout 600 $H/test_lanczos -N $N -m 50 --vector | grep -E "elapsed|iterations"
for nc in 4 8 16; do timeout 600 $H/test_lanczos -N $N -m 10 --block $nc | grep -E "elapsed|iterations"; done
echo "== small parity dumps (N=10)"
timeout 120 $R/ref_cuda_dump_4 vector 10 100 gpurun_out/ref_cuda_vector_N10.bin
timeout 120 $R/ref_cuda_dump_4 block 10 25 gpurun_out/ref_cuda_block4_N10.bin
timeout 120 $R/ref_cuda_dump_8 block 10 25 gpurun_out/ref_cuda_block8_N10.bin
} 2>&1 | tee gpurun_out/ref_cuda_baseline.log
