"""The C++ mirror of the reference surface (host/): the problem generator against the goldens minted
by the reference's own builder (CPU, bit-exact), and the harness test_lanczos driving the library
through the reference's driver signatures (GPU)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_gold

HOST = os.path.join(ROOT, "gpu-implementation-of-signle-and-block-lanczos_b200", "host")


@pytest.fixture(scope="module")
def host_built():
    subprocess.run(["make", "-s", "-C", HOST, "all"], check=True)
    return HOST


@pytest.mark.parametrize("N", [2, 3, 5, 10])
def test_matrix_a_bit_identical_to_reference_builder(host_built, orc, tmp_path, N):
    out = str(tmp_path / "ma.bin")
    subprocess.run([os.path.join(host_built, "dump_matrix_a"), str(N), out], check=True)
    d, g = orc.read_dump(out), load_gold("maxwell_N%d_matrix.npz" % N)
    assert d["n_rows"] == int(g["n_rows"]) == 3 * N * (N + 1) * (2 * N + 1)
    assert d["width"] == 4
    for key in ("D_data", "D_idx", "W_data", "W_idx", "ell_data", "ell_idx"):
        assert np.array_equal(d[key], g[key]), key
    if N == 10:
        gv = load_gold("maxwell_N10_vector_m100.npz")
        assert d["lc"] == int(gv["lc"]) and np.array_equal(d["b"], gv["b"])    # glibc rand() stream, harness order


def run_harness(host, args, tmp_path, exe="test_lanczos"):
    out = str(tmp_path / "run.bin")
    r = subprocess.run([os.path.join(host, exe)] + args + ["--dump", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout, out


@pytest.mark.gpu
@pytest.mark.parametrize("fmt", ["ell", "csr"])
def test_harness_vector_matches_golden(host_built, orc, tmp_path, fmt):
    stdout, out = run_harness(host_built, ["-N", "10", "-m", "100", "--vector", "--format", fmt, "--k", "6"], tmp_path)
    d, g = orc.read_dump(out), load_gold("maxwell_N10_vector_m100.npz")
    assert np.array_equal(d["b"], g["b"]) and d["lc"] == int(g["lc"])
    scale = np.maximum(np.abs(g["alpha"][:50]), np.mean(np.abs(g["beta"][1:50])))
    assert np.max(np.abs(d["alpha"][:50] - g["alpha"][:50]) / scale) < 1e-10
    assert np.max(np.abs(d["beta"][:50] - g["beta"][:50]) / np.abs(g["beta"][:50])) < 1e-10
    assert np.max(np.abs(d["q"][:50] - g["q"][:50])) < 1e-10 * np.abs(g["q"]).max()
    # Ritz values printed by the harness = eig(T) of the golden coefficients, to 1e-8 (north_star)
    w = np.linalg.eigvalsh(orc.assemble_T(g["alpha"], g["beta"]))
    assert np.max(np.abs(d["theta"] - np.r_[w[:3], w[-3:]])) < 1e-8 * np.abs(w).max()
    assert "elapsed time" in stdout and "iterations/s" in stdout


@pytest.mark.gpu
def test_harness_block_matches_golden(host_built, orc, tmp_path):
    stdout, out = run_harness(host_built, ["-N", "10", "-m", "25", "--block", "4"], tmp_path)
    d, g = orc.read_dump(out), load_gold("maxwell_N10_block4_m25.npz")
    assert np.array_equal(d["B"], g["B"])
    nb = 10 * 16
    assert np.max(np.abs(d["alpha"][:nb] - g["alpha"][:nb])) < 1e-10 * np.abs(g["alpha"]).max()
    assert np.max(np.abs(d["beta"][:nb] - g["beta"][:nb])) < 1e-10 * np.abs(g["beta"][:nb]).max()
    m, bw = 25, 4
    T = d["T"].reshape(m * bw, m * bw).T
    a = d["alpha"].reshape(m, bw, bw).transpose(0, 2, 1)
    b = d["beta"].reshape(m + 1, bw, bw).transpose(0, 2, 1)
    assert np.array_equal(T, orc.assemble_T(a, b))                 # Assemble_T through the mirror == oracle layout
    w = np.linalg.eigvalsh(T)
    assert np.max(np.abs(d["theta"] - np.r_[w[:2], w[-2:]])) < 1e-8 * np.abs(w).max()


@pytest.mark.gpu
def test_harness_laplacian_full_reorth(host_built, orc, tmp_path):
    stdout, out = run_harness(host_built, ["--matrix", "lap2d", "-N", "200", "-m", "80", "--vector", "--reorth", "full"], tmp_path)
    d = orc.read_dump(out)
    ref = orc.vector_lanczos(orc.lap2d(200, 200), orc.start_vector(200 * 200), 80, reorth=1)
    assert np.max(np.abs(d["alpha"][:50] - ref["alpha"][:50])) < 1e-10 * np.abs(ref["alpha"]).max()
    assert np.max(np.abs(d["beta"][:50] - ref["beta"][:50]) / ref["beta"][:50]) < 1e-10


@pytest.mark.gpu
def test_harness_vector_application_path(host_built, orc, tmp_path):
    """The post-processing of test_lanczos.cu:97-123 through the mirror: solution from expm(T) and q, the fdtd
    validator with the reference's 100000 steps, and their relative error."""
    stdout, out = run_harness(host_built, ["-N", "10", "-m", "100", "--vector", "--fdtd", "100000"], tmp_path)
    d, g = orc.read_dump(out), load_gold("maxwell_N10_vector_m100.npz")
    assert abs(d["fdtd"][0] - g["fdtd_u_lc"][0]) < 1e-11 * abs(g["fdtd_u_lc"][0])
    assert abs(d["solution"][0] - orc.lanczos_solution(g["alpha"], g["beta"], g["q"], 1.0)) < 1e-10 * abs(d["solution"][0])
    assert abs(d["solution"][0] - d["fdtd"][0]) < 1e-9 * abs(d["fdtd"][0])
    assert "Relative error for vector lanczos is" in stdout and "Solution from fdtd" in stdout


@pytest.mark.gpu
def test_harness_block_application_path(host_built, orc, tmp_path):
    stdout, out = run_harness(host_built, ["-N", "10", "-m", "25", "--block", "4", "--fdtd", "20000"], tmp_path)
    d, g = orc.read_dump(out), load_gold("maxwell_N10_block4_m25.npz")
    assert np.max(np.abs(d["fdtd"] - g["fdtd_row_lc"])) < 1e-11 * np.max(np.abs(g["fdtd_row_lc"]))
    a = g["alpha"].reshape(25, 4, 4).transpose(0, 2, 1)
    b = g["beta"].reshape(26, 4, 4).transpose(0, 2, 1)
    ref = orc.lanczos_solution(a, b, g["q"], 1.0)
    assert np.max(np.abs(d["solution"] - ref)) < 1e-9 * np.max(np.abs(ref))
    assert "Relative error for block lanczos is" in stdout


def test_host_space_fdtd_and_expm_bit_identical_to_reference_host(host_built, orc, tmp_path):
    """Host-space branches of the mirror (no GPU): fdtd_vector / ftdt_block over the mirror's Host containers
    reproduce the reference's own Host run bit for bit (goldens), expm_cusolver matches scipy."""
    import scipy.linalg
    out = str(tmp_path / "hs.bin")
    g = load_gold("maxwell_N10_vector_m100.npz")
    subprocess.run([os.path.join(host_built, "host_space_check"), "vector", "10", str(int(g["fdtd_steps"])), out], check=True)
    d = orc.read_dump(out)
    assert d["lc"] == int(g["lc"]) and d["fdtd"][0] == g["fdtd_u_lc"][0]
    Ti, To = d["expm_in"].reshape(12, 12).T, d["expm_out"].reshape(12, 12).T
    assert np.max(np.abs(To - scipy.linalg.expm(Ti))) < 1e-12 * np.max(np.abs(To))
    gb = load_gold("maxwell_N10_block4_m25.npz")
    subprocess.run([os.path.join(host_built, "host_space_check"), "block", "10", str(int(gb["fdtd_steps"])), out], check=True)
    d = orc.read_dump(out)
    assert np.array_equal(d["fdtd"], gb["fdtd_row_lc"])


def test_matrix_market_reader_matches_scipy(host_built, orc, tmp_path):
    """coordinate files, general and symmetric, pattern and duplicate entries; the CSR arrays and
    a Host-space Csr_matrix::spmv against scipy."""
    import scipy.io
    import scipy.sparse as sp
    A = sp.random(40, 40, density=0.1, random_state=5, format="coo"); A = sp.coo_matrix(A + A.T)
    B = sp.random(30, 50, density=0.15, random_state=6, format="coo")
    scipy.io.mmwrite(str(tmp_path / "sym.mtx"), A, symmetry="symmetric")
    scipy.io.mmwrite(str(tmp_path / "gen.mtx"), B, symmetry="general")
    (tmp_path / "pat.mtx").write_text("%%MatrixMarket matrix coordinate pattern general\n% comment\n3 4 5\n1 1\n3 4\n2 2\n1 1\n3 1\n")
    P = sp.csr_matrix(np.array([[2.0, 0, 0, 0], [0, 1.0, 0, 0], [1.0, 0, 0, 1.0]]))      # duplicate (1,1) summed
    out = str(tmp_path / "m.bin")
    for name, M in (("sym.mtx", sp.csr_matrix(A)), ("gen.mtx", sp.csr_matrix(B)), ("pat.mtx", P)):
        subprocess.run([os.path.join(host_built, "host_space_check"), "mtx", str(tmp_path / name), "0", out], check=True)
        d = orc.read_dump(out)
        M.sort_indices()
        assert tuple(d["dims"]) == (M.shape[0], M.shape[1], M.nnz)
        assert np.array_equal(d["row_ptr"].astype(np.int64), M.indptr) and np.array_equal(d["col_idx"].astype(np.int64), M.indices)
        assert np.allclose(d["data"], M.data, rtol=1e-15, atol=0)
        x = 1.0 + 0.25 * np.arange(M.shape[1])
        assert np.allclose(d["y"], M @ x, rtol=1e-14, atol=1e-14)


@pytest.mark.gpu
def test_harness_matrix_market_operator(host_built, orc, tmp_path):
    """--matrix mtx: a symmetric MatrixMarket file through the harness vs the oracle on the same CSR."""
    import scipy.io
    import scipy.sparse as sp
    rp, ci, va = orc.lap2d(40, 30)
    n = len(rp) - 1
    A = sp.csr_matrix((va, ci, rp), shape=(n, n))
    scipy.io.mmwrite(str(tmp_path / "lap.mtx"), sp.coo_matrix(sp.tril(A)), symmetry="symmetric")
    stdout, out = run_harness(host_built, ["--matrix", "mtx", "--file", str(tmp_path / "lap.mtx"), "-m", "40", "--vector", "--reorth", "full"], tmp_path)
    d = orc.read_dump(out)
    ref = orc.vector_lanczos((rp, ci, va), orc.start_vector(n), 40, reorth=1)
    assert np.max(np.abs(d["alpha"] - ref["alpha"])) < 1e-10 * np.abs(ref["alpha"]).max()
    assert np.max(np.abs(d["beta"] - ref["beta"]) / ref["beta"]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["dgks", "selective"])
def test_harness_reorth_modes_and_residuals(host_built, orc, tmp_path, mode):
    """--reorth dgks / selective through the mirror's driver signatures, and the residual estimates the harness now prints
    next to the Ritz values (|beta_m y_m| from lz_last_coupling + lz_ritz)."""
    stdout, out = run_harness(host_built, ["--matrix", "lap2d", "-N", "120", "-m", "150", "--vector", "--reorth", mode, "--k", "6"], tmp_path)
    d = orc.read_dump(out)
    ref = orc.vector_lanczos(orc.lap2d(120, 120), orc.start_vector(120 * 120), 150, reorth=1)
    tol = 1e-10 if mode == "dgks" else 1e-8
    assert np.max(np.abs(d["alpha"][:50] - ref["alpha"][:50])) < tol * np.abs(ref["alpha"]).max()
    T = orc.assemble_T(ref["alpha"], ref["beta"])
    w, Y = np.linalg.eigh(T)
    sel = np.r_[np.arange(3), np.arange(147, 150)]
    assert np.max(np.abs(d["theta"] - w[sel])) < 1e-8 * np.abs(w).max()
    assert "residual estimates" in stdout and d["resid"].shape == (6,) and np.all(np.isfinite(d["resid"])) and np.all(d["resid"] >= 0)
    # |beta_m y_m| with beta_m from the oracle's run continued by one step
    ref2 = orc.vector_lanczos(orc.lap2d(120, 120), orc.start_vector(120 * 120), 151, reorth=1)
    want = np.abs(ref2["beta"][150] * Y[149, sel])
    assert np.max(np.abs(d["resid"] - want)) < 1e-8


@pytest.mark.gpu
def test_harness_device_assembled_maxwell_equals_host_assembled(host_built, orc, tmp_path):
    """--matrix maxwell_dev (lz_gen_maxwell) drives the same run as the host-assembled operator: the operator arrays are
    bit-identical, so with the same deterministic start vector the coefficients are identical too."""
    g = load_gold("maxwell_N10_matrix.npz")
    n, w = int(g["n_rows"]), int(g["width"])
    csr = orc.ell_to_csr(n, w, g["ell_data"], g["ell_idx"])
    stdout, out = run_harness(host_built, ["--matrix", "maxwell_dev", "-N", "10", "-m", "60", "--vector"], tmp_path)
    d = orc.read_dump(out)
    ref = orc.vector_lanczos(csr, orc.start_vector(n), 60, lc=int(d["lc"]) if "lc" in d else 0)
    scale = np.maximum(np.abs(ref["alpha"][:50]), np.mean(np.abs(ref["beta"][1:50])))
    assert np.max(np.abs(d["alpha"][:50] - ref["alpha"][:50]) / scale) < 1e-10
    assert np.max(np.abs(d["beta"][:50] - ref["beta"][:50]) / np.abs(ref["beta"][:50])) < 1e-10
