#!/bin/bash
# Builds oracle/_ref/ref_cuda_dump_{4,8,16}: the reference's own CUDA drivers for sm_100 (baseline +
# golden source on the GPU box).  TEST / BASELINE INFRASTRUCTURE ONLY.
# Works on a throw-away copy of the reference under $TMP; the only edit is the stub of the dead
# legacy-texture block of kernels/spmv_spmm.hpp (removed from CUDA 12).  No reference source enters the repo.
set -e
REF=${REF:-/root/reference/source}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -d "$REF" ] || { echo "reference tree absent: keeping prebuilt oracle/_ref"; exit 0; }
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
cp -r "$REF" "$TMP/ref"
python3 - "$TMP/ref/kernels/spmv_spmm.hpp" <<'PY'
import sys
p = sys.argv[1]
lines = open(p).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("texture<float> texf"))
end = next(i for i, l in enumerate(lines) if l.strip() == "namespace ell{")
stub = ["namespace text{",
        "template <bool cache, typename T> __device__ T fetch_from_texture1D(const int i, T* x) { return x[i]; }",
        "template <bool cache, typename T> __device__ T fetch_from_texture2D(const int i, T* x, const int, const int) { return x[i]; }",
        "};", ""]
open(p, "w").write("\n".join(lines[:start] + stub + lines[end:]))
PY
mkdir -p "$HERE/_ref"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
CCBIN=$([ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)
for nc in 4 8 16; do
  $NVCC -arch=sm_100 -O3 -w -ccbin $CCBIN -DN_COL=$nc -I"$TMP/ref" -o "$HERE/_ref/ref_cuda_dump_$nc" "$HERE/ref_cuda_dump.cu" -lcublas -lcusolver
done
ls -la "$HERE/_ref"
