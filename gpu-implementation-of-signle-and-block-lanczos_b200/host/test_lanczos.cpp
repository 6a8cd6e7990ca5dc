// test_lanczos.cpp -- the harness, in the shape of the reference's test_lanczos.cu (:20-362):
// build A = D*W (Maxwell, Matrix_A) or a synthetic CSR operator, preallocate every buffer, time the
// driver between two device synchronisations, then post-process (assemble T, Ritz values).
// Plain C++ linked against liblanczos_b200.so; N_COL and USE_BLAS keep their compile-time meaning,
// and the new options are real run-time flags:
//   -N <grid points per dim>   -m <iterations>            (reference flags, :338-345)
//   --matrix maxwell|maxwell_dev|lap2d|lap3d|mtx (--file <MatrixMarket file>)   --block <b>|--vector   --reorth none|full|dgks|selective   --k <ritz pairs>
//   --fdtd <steps> (run the fdtd validator and print the relative error, test_lanczos.cu:115-121, :287-299)   -T <T_end>
//   --format ell|csr   --dump <file>  (alpha/beta/q in the parity tests' record format, for the parity tests)
#ifndef N_COL
#define N_COL 4
#endif
#define USE_BLAS true

#include <cstring>

#include "utils/common.hpp"
#include "utils/lib_utils.hpp"
#include "methods/vector_lanczos.hpp"
#include "methods/block_lanczos.hpp"
#include "methods/fdtd.hpp"
#include "matrix_a/build_A_ell.hpp"
#include "matrix_a/matrix_market.hpp"
#include "objects/tridiagonal_matrix.hpp"

struct Options {
    unsigned int N = 10, m = 5, k = 4, block = N_COL, fdtd_steps = 0;     // fdtd_steps = 0: skip the validator
    bool use_block = true, csr = false;
    std::string matrix = "maxwell", dump, file;
    double T_end = 1;
};

static FILE *g_dump = nullptr;
static void put(const char *name, int dtype, uint64_t count, const void *data, size_t elt)
{
    if (!g_dump) return;
    uint32_t len = (uint32_t)std::strlen(name);
    uint8_t dt = (uint8_t)dtype;
    std::fwrite(&len, 4, 1, g_dump); std::fwrite(name, 1, len, g_dump); std::fwrite(&dt, 1, 1, g_dump);
    std::fwrite(&count, 8, 1, g_dump); std::fwrite(data, elt, count, g_dump);
}
static void put_i64(const char *name, int64_t v) { put(name, 2, 1, &v, 8); }

// synthetic operators of BASELINE.json configs 2/3/5, generated on the device
template <typename type_t>
struct DeviceOperator {
    lz_matrix *op = nullptr;
    std::size_t rows = 0;
    lz_matrix *device_operator() const { return op; }
    std::size_t n_rows() const { return rows; }
    // device-only operator: the Host branches of the drivers are never taken with it
    void spmv(Vector<type_t> &, Vector<type_t> &) { std::cout << "implement later" << std::endl; std::abort(); }
    void spmm(Dense_matrix<type_t> &, Dense_matrix<type_t> &) { std::cout << "implement later" << std::endl; std::abort(); }
};

template <typename type_t, typename Matrix>
void run_vector(Matrix &A, Vector<type_t> &b, const Options &o, unsigned int lc)
{
    const unsigned int m = o.m;
    steady_clock time;
    Vector<type_t> q(m, MemorySpace::CUDA), q0(b), q1(b), w(b);      // test_lanczos.cu:56-63
    std::vector<type_t> alpha(m), beta(m);
    cublasHandle_t cublasH;
    CUBLAS_CHECK(cublasCreate(&cublasH));
    // every buffer exists before the clock starts (the reference preallocates q0, q1, w the same way, :56-63)
    AssertCuda(lz_vector_lanczos_workspace(lanczos_context(), A.device_operator(), (int)m, lzb::reorth_mode()));
    cudaDeviceSynchronize_();
    time.start();
#ifdef USE_BLAS
    vector_lanczos_blas<type_t>(A, b, m, lc, q, alpha.data(), beta.data(), q0, q1, w, cublasH);
#else
    vector_lanczos<type_t>(A, b, m, lc, q, alpha.data(), beta.data(), q0, q1, w);
#endif
    cudaDeviceSynchronize_();
    time.end();
    std::cout << "elapsed time: " << std::setw(11) << time.duration() << std::endl;
    std::cout << "iterations/s: " << m / time.duration() << std::endl;
    const unsigned int k = std::min(o.k, m);
    std::vector<double> theta(k), resid(k);
    double beta_m = 0.0;                                              // coupling to the unbuilt q_{m+1}: residual estimates
    AssertCuda(lz_last_coupling(lanczos_context(), 1, &beta_m));
    AssertCuda(lz_ritz((int)m, 1, alpha.data(), beta.data(), &beta_m, (int)k, theta.data(), resid.data()));
    std::cout << "Ritz values (" << k << " extremal): ";
    for (double t : theta) std::cout << std::setprecision(14) << t << " ";
    std::cout << std::endl << "residual estimates |beta_m y_m|: ";
    for (double r : resid) std::cout << std::setprecision(4) << r << " ";
    std::cout << std::endl;
    put("resid", 0, k, resid.data(), 8);
    Vector<type_t> qh = q.copy_to_host();
    put("alpha", 0, m, alpha.data(), 8); put("beta", 0, m, beta.data(), 8); put("q", 0, m, qh.data(), 8);
    put("theta", 0, k, theta.data(), 8);
    // the approximate solution from T and q (test_lanczos.cu:97-110; the reference leaves T zero there,
    // SURVEY appendix A-5 -- here T is assembled from alpha/beta)
    Dense_matrix<type_t> T(m, m, MemorySpace::Host);
    for (unsigned int i = 0; i < m; ++i) {
        T(i + i * m) = alpha[i];
        if (i + 1 < m) { T((i + 1) + i * m) = beta[i + 1]; T(i + (i + 1) * m) = beta[i + 1]; }
    }
    T.mult_scalar(o.T_end);
    expm_cusolver(T);
    Vector<type_t> e1(m, MemorySpace::Host);
    copy_column_to_vector<type_t>(T, e1, 0);
    type_t solution = e1.dot(qh);
    solution = beta[0] * solution;
    std::cout << "The solution for vector_lanczos " << std::endl << std::setprecision(14) << solution << std::endl;
    put("solution", 0, 1, &solution, 8);
    if (o.fdtd_steps) {                                                // :115-121
        type_t fdtd_solution = fdtd_vector(A, b, o.fdtd_steps, o.T_end, lc);
        std::cout << "Solution from fdtd " << fdtd_solution << std::endl;
        std::cout << "Relative error for vector lanczos is " << std::abs(solution - fdtd_solution) / std::abs(fdtd_solution) << std::endl;
        put("fdtd", 0, 1, &fdtd_solution, 8);
    }
    CUBLAS_CHECK(cublasDestroy(cublasH));
}

template <typename type_t, typename Matrix>
void run_block(Matrix &A, Dense_matrix<type_t> &B, const Options &o, unsigned int lc)
{
    const unsigned int m = o.m, bw = (unsigned int)B.n_cols();
    const MemorySpace mem_cuda = MemorySpace::CUDA;
    steady_clock time;
    Vector<type_t> q(m * bw, mem_cuda);                               // test_lanczos.cu:203-223
    Dense_matrix<type_t> Q0(B), Q1(B), W(B);
    Dense_matrix<type_t> *alpha = new Dense_matrix<type_t>[m];
    Dense_matrix<type_t> *beta = new Dense_matrix<type_t>[m + 1];
    for (unsigned int i = 0; i < m; ++i) {
        alpha[i] = Dense_matrix<type_t>(bw, bw, mem_cuda);
        beta[i] = Dense_matrix<type_t>(bw, bw, mem_cuda);
    }
    beta[m] = Dense_matrix<type_t>(bw, bw, mem_cuda);
    cublasHandle_t cublasH;
    CUBLAS_CHECK(cublasCreate(&cublasH));
    std::cout << " start Lanczos " << std::endl;
    cusolver_args<type_t> args = cusolver_args<type_t>();
    Vector<type_t> eigen_val(bw, mem_cuda);
    initiate_cusolver(args, beta[0], eigen_val);
    // every buffer exists before the clock starts (the reference preallocates Q0, Q1, W and the cusolver workspace, :203-233)
    AssertCuda(lz_block_lanczos_workspace(lanczos_context(), A.device_operator(), (int)bw, (int)m,
                                          lzb::reorth_mode() == LZ_REORTH_NONE ? LZ_REORTH_NONE : LZ_REORTH_FULL));
    cudaDeviceSynchronize_();
    time.start();
#ifdef USE_BLAS
    block_lanczos_blas<type_t>(A, B, m, lc, q, alpha, beta, Q0, Q1, W, args, eigen_val, cublasH, 0, 0);
#else
    block_lanczos(A, B, m, lc, q, alpha, beta, Q0, Q1, W, 0, 0);
#endif
    cudaDeviceSynchronize_();
    time.end();
    std::cout << " end Lanczos " << std::endl;
    std::cout << "elapsed time: " << std::setw(11) << time.duration() << std::endl;
    std::cout << "iterations/s: " << m / time.duration() << std::endl;
    // T and its extremal Ritz values (the syevd(T) of expm_cusolver, lib_utils.hpp:542-590)
    Dense_matrix<type_t> T = Assemble_T(m, alpha, beta);
    const std::size_t bb = (std::size_t)bw * bw;
    std::vector<double> a(m * bb), bt((m + 1) * bb);
    for (unsigned int i = 0; i < m; ++i) lzb::dcopy(&a[i * bb], alpha[i].data(), bb * 8, LZ_D2H);
    for (unsigned int i = 0; i <= m; ++i) lzb::dcopy(&bt[i * bb], beta[i].data(), bb * 8, LZ_D2H);
    const unsigned int k = std::min<unsigned int>(o.k, m * bw);
    std::vector<double> theta(k), resid(k);
    std::vector<double> beta_m(bb);                                   // coupling to the unbuilt block: residual estimates
    AssertCuda(lz_last_coupling(lanczos_context(), (int)bw, beta_m.data()));
    AssertCuda(lz_ritz((int)m, (int)bw, a.data(), bt.data(), beta_m.data(), (int)k, theta.data(), resid.data()));
    std::cout << "Ritz values (" << k << " extremal): ";
    for (double t : theta) std::cout << std::setprecision(14) << t << " ";
    std::cout << std::endl << "residual estimates ||beta_m Y_m||: ";
    for (double r : resid) std::cout << std::setprecision(4) << r << " ";
    std::cout << std::endl;
    put("resid", 0, k, resid.data(), 8);
    Vector<type_t> qh = q.copy_to_host();
    Dense_matrix<type_t> Th = T.copy_to_host();
    put("alpha", 0, a.size(), a.data(), 8); put("beta", 0, bt.size(), bt.data(), 8); put("q", 0, qh.size(), qh.data(), 8);
    put("theta", 0, k, theta.data(), 8); put("T", 0, Th.size(), Th.data(), 8);
    // the approximate solution from T and q (test_lanczos.cu:266-283)
    T.mult_scalar(o.T_end);
    expm_cusolver(T);
    Dense_matrix<type_t> F1(m * bw, bw, mem_cuda);
    copy_columns_to_matrix<type_t>(T, F1, bw);
    {
        Dense_matrix<type_t> F0(F1);
        mm_cublas(0, 1., F0, beta[0], F1, cublasH);                  // F1 = expm(T)[:, 0:b] * sqrtm(B^T B)
    }
    Vector<type_t> solution(bw, mem_cuda);
    vm_cublas(F1, q, solution, cublasH);
    std::cout << "Solution for block lanczos";
    solution.print();
    Vector<type_t> sh = solution.copy_to_host();
    put("solution", 0, bw, sh.data(), 8);
    if (o.fdtd_steps) {                                                // :287-299
        std::cout << " start fdtd " << std::endl;
        Vector<type_t> fdtd_solution = ftdt_block<type_t>(A, B, o.fdtd_steps, o.T_end, lc);
        std::cout << "Solution from fdtd ";
        fdtd_solution.print();
        Vector<type_t> fh = fdtd_solution.copy_to_host();
        put("fdtd", 0, bw, fh.data(), 8);
        solution.sadd(1, -1, fdtd_solution);
        type_t relative_error = solution.l2_norm() / fdtd_solution.l2_norm();
        std::cout << "Relative error for block lanczos is " << relative_error << std::endl;
    }
    CUBLAS_CHECK(cublasDestroy(cublasH));
    delete[] alpha;
    delete[] beta;
}

template <typename type_t>
void test_Lanczos(const Options &o, unsigned int lc)
{
    if (o.matrix == "maxwell") {
        auto info = Matrix_A<type_t>(o.N, o.N, o.N);                   // test_lanczos.cu:29-49, 142-195
        Ell_matrix<type_t> D_host = info.first;
        Ell_matrix<type_t> W_host = info.second;
        const unsigned int n_rows = (unsigned int)D_host.n_rows();
        std::cout << " the size of the problem is " << std::endl;
        print(n_rows);
        put_i64("n_rows", n_rows); put_i64("lc", lc);
        D_host.mult_diagonal(W_host);                                  // A = D*W, symmetric
        if (o.use_block) {
            Dense_matrix<type_t> B_host = random_matrix_B<type_t>(n_rows, o.block);
            put("B", 0, B_host.size(), B_host.data(), 8);
            Dense_matrix<type_t> B = B_host.copy_to_device();
            if (o.csr) { Csr_matrix<type_t> A = Csr_matrix<type_t>(D_host).copy_to_device(); run_block<type_t>(A, B, o, lc); }
            else { D_host.change_order(4); Ell_matrix<type_t> A = D_host.copy_to_device(); run_block<type_t>(A, B, o, lc); }
        } else {
            Vector<type_t> b_host = random_vector_b<type_t>(n_rows);
            put("b", 0, b_host.size(), b_host.data(), 8);
            Vector<type_t> b = b_host.copy_to_device();
            if (o.csr) { Csr_matrix<type_t> A = Csr_matrix<type_t>(D_host).copy_to_device(); run_vector<type_t>(A, b, o, lc); }
            else { D_host.change_order(4); Ell_matrix<type_t> A = D_host.copy_to_device(); run_vector<type_t>(A, b, o, lc); }
        }
        return;
    }
    if (o.matrix == "mtx") {                                           // symmetric operator from a MatrixMarket file
        Csr_matrix<type_t> A = read_matrix_market<type_t>(o.file).copy_to_device();
        if (A.n_rows() != A.n_cols()) { std::cout << "the operator must be square" << std::endl; std::abort(); }
        std::cout << " the size of the problem is " << std::endl;
        print(A.n_rows());
        lc = lc % A.n_rows();
        if (o.use_block) {
            Dense_matrix<type_t> B(A.n_rows(), o.block, MemorySpace::CUDA);
            AssertCuda(lz_gen_start_block(lanczos_context(), (int64_t)A.n_rows(), (int)o.block, (int64_t)A.n_rows(), 0x5EED, B.data()));
            run_block<type_t>(A, B, o, lc);
        } else {
            Vector<type_t> b(A.n_rows(), MemorySpace::CUDA);
            AssertCuda(lz_gen_start_vector(lanczos_context(), (int64_t)A.n_rows(), 0x5EED, b.data()));
            run_vector<type_t>(A, b, o, lc);
        }
        return;
    }
    DeviceOperator<type_t> A;
    if (o.matrix == "lap2d") { AssertCuda(lz_gen_laplacian2d(lanczos_context(), o.N, o.N, &A.op)); A.rows = (std::size_t)o.N * o.N; }
    else if (o.matrix == "lap3d") { AssertCuda(lz_gen_laplacian3d(lanczos_context(), o.N, o.N, o.N, &A.op)); A.rows = (std::size_t)o.N * o.N * o.N; }
    else if (o.matrix == "maxwell_dev") {      // the reference's operator A = D*W assembled on the device (bit-identical to Matrix_A + mult_diagonal)
        AssertCuda(lz_gen_maxwell(lanczos_context(), (int)o.N, (int)o.N, (int)o.N, &A.op));
        int64_t nr = 0;
        AssertCuda(lz_matrix_info(A.op, &nr, nullptr, nullptr));
        A.rows = (std::size_t)nr;
    }
    else { std::cout << "unknown --matrix " << o.matrix << std::endl; std::abort(); }
    std::cout << " the size of the problem is " << std::endl;
    print(A.rows);
    lc = lc % A.rows;
    if (o.use_block) {
        Dense_matrix<type_t> B(A.rows, o.block, MemorySpace::CUDA);
        AssertCuda(lz_gen_start_block(lanczos_context(), (int64_t)A.rows, (int)o.block, (int64_t)A.rows, 0x5EED, B.data()));
        run_block<type_t>(A, B, o, lc);
    } else {
        Vector<type_t> b(A.rows, MemorySpace::CUDA);
        AssertCuda(lz_gen_start_vector(lanczos_context(), (int64_t)A.rows, 0x5EED, b.data()));
        run_vector<type_t>(A, b, o, lc);
    }
    lz_matrix_destroy(A.op);
}

int main(int argc, char **argv)
{
    Options o;
    // location of interest: one rand() before the right-hand side draws its own (test_lanczos.cu:326)
    unsigned int lc = 1 + (rand() % 100);
    for (int l = 1; l < argc; ++l) {
        const std::string opt = argv[l];
        auto next = [&]() -> std::string {
            if (l + 1 >= argc) { std::cout << "Error, option " << opt << " expects a value" << std::endl; std::abort(); }
            return argv[++l];
        };
        if (opt == "-N") o.N = (unsigned int)std::stod(next());
        else if (opt == "-m") o.m = (unsigned int)std::stod(next());
        else if (opt == "--k") o.k = (unsigned int)std::stod(next());
        else if (opt == "--fdtd") o.fdtd_steps = (unsigned int)std::stod(next());
        else if (opt == "-T") o.T_end = std::stod(next());
        else if (opt == "--matrix") o.matrix = next();
        else if (opt == "--file") o.file = next();
        else if (opt == "--block") { o.use_block = true; o.block = (unsigned int)std::stod(next()); }
        else if (opt == "--vector") o.use_block = false;
        else if (opt == "--format") o.csr = next() == "csr";
        else if (opt == "--dump") o.dump = next();
        else if (opt == "--reorth") {
            const std::string r = next();
            lzb::reorth_mode() = r == "full" ? LZ_REORTH_FULL : r == "dgks" ? LZ_REORTH_FULL_DGKS : r == "selective" ? LZ_REORTH_SELECTIVE : LZ_REORTH_NONE;
        } else if (opt == "-blas") next();                               // advertised by the reference, ignored there too
        else { std::cout << "Error, unknown option " << opt << std::endl; std::abort(); }
    }
    if (!o.dump.empty()) g_dump = std::fopen(o.dump.c_str(), "wb");
    put_i64("N", o.N); put_i64("m", o.m);
    test_Lanczos<double>(o, lc);
    if (g_dump) std::fclose(g_dump);
    return 0;
}
