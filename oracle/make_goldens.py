"""Mint tests/golden/*.npz from the reference's own Host code (oracle/_ref/ref_host_dump_*).

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference to build oracle/_ref):

    python oracle/make_goldens.py

The fixtures pin (a) the Maxwell operator built by matrix_a/build_A_ell.hpp (D, W and A = D*W in the
reference's column-major ELL layout) at N = 2, 3, 5, 10, (b) glibc-rand() start vectors exactly as
test_lanczos.cu draws them (one rand() consumed for lc first), and (c) the alpha/beta/q series of
methods/vector_lanczos.hpp:20-66 (m = 100) and methods/block_lanczos.hpp:105-166 (N_COL = 4 and 8,
m = 25) executed over the reference's Host containers, and (d) the harness' fdtd validator
(methods/fdtd.hpp) run over the same containers: u[lc] after 100000 Euler steps (vector) and row lc of U
after 20000 steps (block).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import orc  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    orc.build()
    os.makedirs(GOLD, exist_ok=True)
    for N in (2, 3, 5, 10):
        d = orc.run_ref("matrix", N, 1)
        np.savez_compressed(os.path.join(GOLD, "maxwell_N%d_matrix.npz" % N),
                            N=N, n_rows=d["n_rows"], n_cols=d["n_cols"], width=d["width"],
                            D_data=d["D_data"], D_idx=d["D_idx"], W_data=d["W_data"], W_idx=d["W_idx"],
                            W_width=d["W_width"], ell_data=d["ell_data"], ell_idx=d["ell_idx"])
    d = orc.run_ref("vector", 10, 100)
    np.savez_compressed(os.path.join(GOLD, "maxwell_N10_vector_m100.npz"),
                        N=10, m=100, lc=d["lc"], b=d["b"], alpha=d["alpha"], beta=d["beta"], q=d["q"],
                        fdtd_steps=d["fdtd_steps"], fdtd_u_lc=d["fdtd_u_lc"])
    for nc in (4, 8):
        d = orc.run_ref("block", 10, 25, n_col=nc)
        np.savez_compressed(os.path.join(GOLD, "maxwell_N10_block%d_m25.npz" % nc),
                            N=10, m=25, lc=d["lc"], n_col=nc, B=d["B"], alpha=d["alpha"], beta=d["beta"], q=d["q"],
                            fdtd_steps=d["fdtd_steps"], fdtd_row_lc=d["fdtd_row_lc"])
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
