// dump_matrix_a.cpp -- writes Matrix_A<double>(N,N,N) (D, W and A = D*W) in the record container of
// tests' reference-side dump tool so the CPU test suite can compare the mirror's builder with the golden
// arrays minted by the reference's own builder.  Host-only: needs no GPU.
#include <cstring>

#include "matrix_a/build_A_ell.hpp"

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
    if (argc < 3) return 2;
    const unsigned int N = (unsigned int)std::atoi(argv[1]);
    g_out = std::fopen(argv[2], "wb");
    if (!g_out) return 3;
    auto info = Matrix_A<double>(N, N, N);
    Ell_matrix<double> D = info.first, W = info.second;
    int64_t v;
    put("D_data", 0, D.size(), const_cast<const Ell_matrix<double> &>(D).data(), 8);
    put("D_idx", 1, D.size(), const_cast<const Ell_matrix<double> &>(D).idx(), 4);
    put("W_data", 0, W.size(), const_cast<const Ell_matrix<double> &>(W).data(), 8);
    put("W_idx", 1, W.size(), const_cast<const Ell_matrix<double> &>(W).idx(), 4);
    D.mult_diagonal(W);
    v = (int64_t)D.n_rows(); put("n_rows", 2, 1, &v, 8);
    v = (int64_t)D.n_cols(); put("n_cols", 2, 1, &v, 8);
    v = (int64_t)D.width(); put("width", 2, 1, &v, 8);
    put("ell_data", 0, D.size(), const_cast<const Ell_matrix<double> &>(D).data(), 8);
    put("ell_idx", 1, D.size(), const_cast<const Ell_matrix<double> &>(D).idx(), 4);
    // the rand()-drawn right-hand sides in harness order (one rand() for lc first)
    const unsigned int lc = 1 + (rand() % 100);
    v = lc; put("lc", 2, 1, &v, 8);
    Vector<double> b = random_vector_b<double>((unsigned int)D.n_rows());
    put("b", 0, b.size(), b.data(), 8);
    std::fclose(g_out);
    return 0;
}
