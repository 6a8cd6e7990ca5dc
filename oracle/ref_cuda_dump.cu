/*
 * ref_cuda_dump.cu -- runs the REFERENCE's own CUDA drivers and dumps what they produce.
 * TEST / BASELINE INFRASTRUCTURE ONLY (never part of the product).
 *
 * Compiled by oracle/build_ref_cuda.sh against a build-time copy of /root/reference/source in which
 * exactly one thing is changed: the dead legacy-texture block of kernels/spmv_spmm.hpp (lines 7-98,
 * texture<> references and cudaBindTexture* calls that CUDA >= 12 no longer has) is replaced by
 * pass-through fetch_from_texture{1D,2D} templates.  Everything else -- containers, kernels,
 * cuBLAS/cuSOLVER wrappers, drivers -- is the reference's code, compiled for sm_100.
 *
 * The set-up follows test_lanczos.cu with the one correction SURVEY.md appendix A-1 calls for:
 * change_order(4) runs on the DEVICE copy (the Host branch only moves ELL column 0).  Drivers:
 * vector_lanczos<double> (methods/vector_lanczos.hpp:8-67, the correct non-BLAS variant) and
 * block_lanczos_blas<double> (methods/block_lanczos.hpp:88-167).
 *
 * usage: ref_cuda_dump <vector|block> N m out.bin     (record container of ref_host_dump.cpp)
 */
#ifndef N_COL
#define N_COL 4
#endif
#define USE_BLAS true

#include <cstring>

#include "utils/common.hpp"
#include "utils/lib_utils.hpp"
#include "methods/vector_lanczos.hpp"
#include "methods/block_lanczos.hpp"
#include "matrix_a/build_A_ell.hpp"

static FILE *g_out = nullptr;
static void put(const char *name, int dtype, uint64_t count, const void *data, size_t elt)
{
    uint32_t len = (uint32_t)std::strlen(name);
    uint8_t dt = (uint8_t)dtype;
    std::fwrite(&len, 4, 1, g_out); std::fwrite(name, 1, len, g_out); std::fwrite(&dt, 1, 1, g_out);
    std::fwrite(&count, 8, 1, g_out); std::fwrite(data, elt, count, g_out);
}
static void put_i64(const char *name, int64_t v) { put(name, 2, 1, &v, 8); }
static void put_f64s(const char *name, double v) { put(name, 0, 1, &v, 8); }

int main(int argc, char **argv)
{
    if (argc < 5) { std::fprintf(stderr, "usage: %s <vector|block> N m out.bin\n", argv[0]); return 2; }
    const std::string mode = argv[1];
    const unsigned int N = (unsigned int)std::atoi(argv[2]), m = (unsigned int)std::atoi(argv[3]);
    g_out = std::fopen(argv[4], "wb");
    if (!g_out) return 3;
    typedef double T;
    const MemorySpace mem_cuda = MemorySpace::CUDA;
    const unsigned int lc = 1 + (rand() % 100);
    auto info = Matrix_A<T>(N, N, N);
    Ell_matrix<T> D_host = info.first;
    Ell_matrix<T> W_host = info.second;
    const unsigned int n = (unsigned int)D_host.n_rows();
    D_host.mult_diagonal(W_host);
    Ell_matrix<T> A = D_host.copy_to_device();
    A.change_order(4);                                  /* device branch: lm::change_major */
    put_i64("N", N); put_i64("m", m); put_i64("lc", lc); put_i64("n_rows", n); put_i64("n_col", N_COL);
    steady_clock time = steady_clock();
    if (mode == "vector") {
        Vector<T> b_host = random_vector_b<T>(n);
        Vector<T> b = b_host.copy_to_device();
        Vector<T> q(m, mem_cuda), q0(b), q1(b), w(b);
        std::vector<T> alpha(m), beta(m);
        cudaDeviceSynchronize();
        time.start();
        vector_lanczos<T>(A, b, m, lc, q, alpha.data(), beta.data(), q0, q1, w);
        cudaDeviceSynchronize();
        time.end();
        Vector<T> qh = q.copy_to_host();
        put("alpha", 0, m, alpha.data(), 8); put("beta", 0, m, beta.data(), 8); put("q", 0, m, qh.data(), 8);
    } else {
        Dense_matrix<T> B_host = random_matrix_B<T>(n);
        Dense_matrix<T> B = B_host.copy_to_device();
        Vector<T> q(m * N_COL, mem_cuda);
        Dense_matrix<T> Q0(B), Q1(B), W(B);
        Dense_matrix<T> *alpha = new Dense_matrix<T>[m];
        Dense_matrix<T> *beta = new Dense_matrix<T>[m + 1];
        for (unsigned int i = 0; i < m; ++i) {
            alpha[i] = Dense_matrix<T>(N_COL, N_COL, mem_cuda);
            beta[i] = Dense_matrix<T>(N_COL, N_COL, mem_cuda);
        }
        beta[m] = Dense_matrix<T>(N_COL, N_COL, mem_cuda);
        cublasHandle_t cublasH;
        CUBLAS_CHECK(cublasCreate(&cublasH));
        cusolver_args<T> args = cusolver_args<T>();
        Vector<T> eigen_val(N_COL, mem_cuda);
        initiate_cusolver(args, beta[0], eigen_val);
        cudaDeviceSynchronize();
        time.start();
        block_lanczos_blas<T>(A, B, m, lc, q, alpha, beta, Q0, Q1, W, args, eigen_val, cublasH, 0, 0);
        cudaDeviceSynchronize();
        time.end();
        std::vector<T> a((size_t)m * N_COL * N_COL), bt((size_t)(m + 1) * N_COL * N_COL);
        for (unsigned int i = 0; i < m; ++i) {
            Dense_matrix<T> h = alpha[i].copy_to_host();
            std::memcpy(&a[(size_t)i * N_COL * N_COL], h.data(), sizeof(T) * N_COL * N_COL);
        }
        for (unsigned int i = 0; i <= m; ++i) {
            Dense_matrix<T> h = beta[i].copy_to_host();
            std::memcpy(&bt[(size_t)i * N_COL * N_COL], h.data(), sizeof(T) * N_COL * N_COL);
        }
        Vector<T> qh = q.copy_to_host();
        put("alpha", 0, a.size(), a.data(), 8); put("beta", 0, bt.size(), bt.data(), 8); put("q", 0, qh.size(), qh.data(), 8);
    }
    put_f64s("elapsed_s", time.duration());
    std::printf("%s N=%u n=%u m=%u N_COL=%d elapsed %.6f s  (%.3f iterations/s)\n", mode.c_str(), N, n, m, N_COL,
                time.duration(), m / time.duration());
    std::fclose(g_out);
    return 0;
}
