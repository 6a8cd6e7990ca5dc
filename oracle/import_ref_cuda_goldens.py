"""Turn the dumps written on the GPU box by tools/run_ref_cuda.sh (gpurun_out/ref_cuda_*_N10.bin: the
REFERENCE's own CUDA drivers, sm_100 build, B200) into committed fixtures tests/golden/ref_cuda_*.npz.
TEST INFRASTRUCTURE ONLY.   python oracle/import_ref_cuda_goldens.py [gpurun_out]"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import orc  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "gpurun_out")
gold = os.path.join(os.path.dirname(HERE), "tests", "golden")
for name in ("vector_N10", "block4_N10", "block8_N10"):
    d = orc.read_dump(os.path.join(src, "ref_cuda_%s.bin" % name))
    np.savez_compressed(os.path.join(gold, "ref_cuda_%s.npz" % name), **{k: v for k, v in d.items() if k != "elapsed_s"})
    print(name, {k: (v.shape if hasattr(v, "shape") else v) for k, v in d.items()})
