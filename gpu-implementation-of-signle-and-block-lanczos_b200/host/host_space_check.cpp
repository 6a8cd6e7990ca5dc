// host_space_check.cpp -- runs the Host-space branches of the mirror (the reference's own loops over the
// Host containers) for the application path: fdtd_vector / ftdt_block (methods/fdtd.hpp) on the Maxwell
// operator with the rand()-drawn right-hand sides of the harness, and expm_cusolver on a small symmetric
// matrix.  Host-only: needs no GPU.  The CPU test suite compares the output with the goldens minted from
// the reference's own Host code (bit for bit).
// usage: host_space_check <vector|block> N steps out.bin
//        host_space_check mtx <file.mtx> 0 out.bin     (MatrixMarket reader + Host-space Csr_matrix::spmv)
#ifndef N_COL
#define N_COL 4
#endif
#include <cstring>

#include "utils/common.hpp"
#include "utils/lib_utils.hpp"
#include "methods/fdtd.hpp"
#include "matrix_a/build_A_ell.hpp"
#include "matrix_a/matrix_market.hpp"

static FILE *g_out;
static void put(const char *name, int dtype, uint64_t count, const void *data, size_t elt)
{
    uint32_t len = (uint32_t)std::strlen(name);
    uint8_t dt = (uint8_t)dtype;
    std::fwrite(&len, 4, 1, g_out); std::fwrite(name, 1, len, g_out); std::fwrite(&dt, 1, 1, g_out);
    std::fwrite(&count, 8, 1, g_out); std::fwrite(data, elt, count, g_out);
}

int main(int argc, char **argv)
{
    if (argc < 5) return 2;
    const std::string mode = argv[1];
    g_out = std::fopen(argv[4], "wb");
    if (!g_out) return 3;
    if (mode == "mtx") {
        Csr_matrix<double> M = read_matrix_market<double>(argv[2]);
        const Csr_matrix<double> &Mc = M;
        int64_t dims[3] = {(int64_t)M.n_rows(), (int64_t)M.n_cols(), (int64_t)M.nnz()};
        put("dims", 2, 3, dims, 8);
        put("row_ptr", 1, M.n_rows() + 1, Mc.row_ptr(), 4);
        put("col_idx", 1, M.nnz(), Mc.col_idx(), 4);
        put("data", 0, M.nnz(), Mc.data(), 8);
        Vector<double> x((unsigned int)M.n_cols(), MemorySpace::Host), y((unsigned int)M.n_rows(), MemorySpace::Host);
        for (unsigned int i = 0; i < M.n_cols(); ++i) x(i) = 1.0 + 0.25 * i;
        M.spmv(x, y);
        put("y", 0, M.n_rows(), y.data(), 8);
        std::fclose(g_out);
        return 0;
    }
    const unsigned int N = (unsigned int)std::atoi(argv[2]), steps = (unsigned int)std::atoi(argv[3]);
    const unsigned int lc = 1 + (rand() % 100);                       // test_lanczos.cu:326
    auto info = Matrix_A<double>(N, N, N);
    Ell_matrix<double> A = info.first, W = info.second;
    A.mult_diagonal(W);                                               // :43, :191 (column-major ELL, Host)
    const unsigned int n = (unsigned int)A.n_rows();
    int64_t v = lc;
    put("lc", 2, 1, &v, 8);
    if (mode == "vector") {
        Vector<double> b = random_vector_b<double>(n);
        double r = fdtd_vector(A, b, steps, 1.0, lc);
        put("fdtd", 0, 1, &r, 8);
        // expm_cusolver on a Host-space symmetric matrix
        const unsigned int m = 12;
        Dense_matrix<double> T(m, m, MemorySpace::Host);
        for (unsigned int j = 0; j < m; ++j)
            for (unsigned int i = 0; i < m; ++i) T(i + j * m) = 0.3 * std::cos(0.7 * (i + 1) * (j + 1)) + (i == j ? 0.1 * i : 0.0);
        for (unsigned int j = 0; j < m; ++j)
            for (unsigned int i = 0; i < j; ++i) T(i + j * m) = T(j + i * m);
        put("expm_in", 0, m * m, T.data(), 8);
        expm_cusolver(T);
        put("expm_out", 0, m * m, T.data(), 8);
    } else {
        Dense_matrix<double> B = random_matrix_B<double>(n);
        Vector<double> r = ftdt_block<double>(A, B, steps, 1.0, lc);
        put("fdtd", 0, N_COL, r.data(), 8);
    }
    std::fclose(g_out);
    return 0;
}
