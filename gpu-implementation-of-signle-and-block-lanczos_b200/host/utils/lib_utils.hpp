// utils/lib_utils.hpp -- the reference's cuBLAS / cuSOLVER wrapper names (utils/lib_utils.hpp:11-805)
// bound to the new kernels.  cublasHandle_t / cusolver_args are kept as inert types so
// block_lanczos_blas's signature and test_lanczos.cu's setup code compile unchanged; no NVIDIA
// library is called.  The own-kernel names of the reference (mm_tt, mm_tt2, mm_ts,
// my_sqrtm_cusolver -- kernels/mm_tt.hpp:158, mm_tt2.hpp:185, mm_ts.hpp:223-273,
// my_sqrtm_cusolver.hpp:366) are provided too, without their tuning arguments' restrictions.
#ifndef lzb_lib_utils_hpp
#define lzb_lib_utils_hpp

#include "../objects/ell_matrix.hpp"

typedef void *cublasHandle_t;
inline int cublasCreate(cublasHandle_t *h) { *h = nullptr; lanczos_context(); return LZ_OK; }
inline int cublasDestroy(cublasHandle_t) { return LZ_OK; }

template <typename type_t>
struct cusolver_args {                      // lib_utils.hpp:11-24; nothing to configure any more
    cusolver_args() : lwork(0), syevj_work(nullptr), syevj_info(nullptr), syevj_params(nullptr), cusolverH(nullptr) {}
    int lwork;
    type_t *syevj_work;
    int *syevj_info;
    void *syevj_params;
    void *cusolverH;
};
inline int cusolverDnDestroySyevjInfo(void *) { return LZ_OK; }
inline int cusolverDnDestroy(void *) { return LZ_OK; }
template <typename type_t>
void initiate_cusolver(cusolver_args<type_t> &, Dense_matrix<type_t> &, Vector<type_t> &) {}   // :747-805

// result = my_scalar*result + other_scalar*other1*other2          (mm_cublas :28-75, mm_ts/mm_ts2)
inline void mm_cublas(const double my_scalar, const double other_scalar, Dense_matrix<double> &other1, Dense_matrix<double> &other2,
                      Dense_matrix<double> &result, cublasHandle_t = nullptr)
{
    CUBLAS_CHECK(lz_mm_ts(lanczos_context(), (int64_t)other1.n_rows(), (int)other2.n_rows(), my_scalar, other_scalar, other1.data(),
                          (int64_t)other1.n_rows(), other2.data(), result.data(), (int64_t)result.n_rows()));
}
inline void mm_ts(const unsigned int, const unsigned int, Dense_matrix<double> &T, Dense_matrix<double> &S, Dense_matrix<double> &R)
{
    mm_cublas(0., 1., T, S, R);
}
inline void mm_ts(const unsigned int, const unsigned int, const double my_scalar, const double other_scalar, Dense_matrix<double> &T,
                  Dense_matrix<double> &S, Dense_matrix<double> &R)
{
    mm_cublas(my_scalar, other_scalar, T, S, R);     // the reference ignores the scalars here (appendix A-4); we honour them
}
// result = T^T T                                                   (mm_tt_cublas :80-123, mm_tt)
inline void mm_tt_cublas(Dense_matrix<double> &T, Dense_matrix<double> &result, cublasHandle_t = nullptr)
{
    CUBLAS_CHECK(lz_mm_tt(lanczos_context(), (int64_t)T.n_rows(), (int)T.n_cols(), T.data(), (int64_t)T.n_rows(), result.data()));
}
inline void mm_tt(const unsigned int, const unsigned int, Dense_matrix<double> &T, Dense_matrix<double> &result) { mm_tt_cublas(T, result); }
// result = 0.5 (T1^T T2 + T2^T T1)                                 (mm_tt2_cublas :126-202, mm_tt2)
inline void mm_tt2_cublas(Dense_matrix<double> &T1, Dense_matrix<double> &T2, Dense_matrix<double> &result, cublasHandle_t = nullptr)
{
    CUBLAS_CHECK(lz_mm_tt2(lanczos_context(), (int64_t)T1.n_rows(), (int)T1.n_cols(), T1.data(), (int64_t)T1.n_rows(), T2.data(),
                           (int64_t)T2.n_rows(), result.data()));
}
inline void mm_tt2(const unsigned int, const unsigned int, Dense_matrix<double> &T1, Dense_matrix<double> &T2, Dense_matrix<double> &result)
{
    mm_tt2_cublas(T1, T2, result);
}
// result = mat^T vec                                                (vm_cublas :388-429); small: host arithmetic
inline void vm_cublas(Dense_matrix<double> &mat, Vector<double> &vec, Vector<double> &result, cublasHandle_t = nullptr)
{
    const bool dev = mat.memory_space() == MemorySpace::CUDA;
    Dense_matrix<double> m = dev ? mat.copy_to_host() : mat;
    Vector<double> v = vec.memory_space() == MemorySpace::CUDA ? vec.copy_to_host() : vec;
    Vector<double> r(mat.n_cols(), MemorySpace::Host);
    for (std::size_t c = 0; c < m.n_cols(); ++c) {
        double s = 0;
        for (std::size_t i = 0; i < m.n_rows(); ++i) s += m(i + c * m.n_rows()) * v(i);
        r(c) = s;
    }
    result = result.memory_space() == MemorySpace::CUDA ? r.copy_to_device() : r;
}
// level-1 wrappers (:431-538): y += a*x, dot, nrm2, scal
inline void vec_add_cublas(const double other_scalar, Vector<double> &vec1, Vector<double> &vec2, cublasHandle_t = nullptr)
{
    vec2.add(other_scalar, vec1);
}
inline void dot_cublas(Vector<double> &vec1, Vector<double> &vec2, double *result, cublasHandle_t = nullptr) { *result = vec1.dot(vec2); }
inline void l2_norm_cublas(Vector<double> &vec, double *result, cublasHandle_t = nullptr)
{
    CUBLAS_CHECK(lz_nrm2(lanczos_context(), (int64_t)vec.size(), vec.data(), result));
}
inline void mult_scalar_cublas(Vector<double> &vec, const double scalar, cublasHandle_t = nullptr) { vec.mult_scalar(scalar); }

// T <- expm(T) = V exp(Lambda) V^T for the small symmetric projected matrix    (expm_cusolver :542-590 +
// Dense_matrix::custom_mult; host arithmetic inside the library, lz_expm_sym)
inline void expm_cusolver(Dense_matrix<double> &T)
{
    const bool dev = T.memory_space() == MemorySpace::CUDA;
    Dense_matrix<double> h = dev ? T.copy_to_host() : T;
    CUSOLVER_CHECK(lz_expm_sym((int)h.n_rows(), h.data()));
    T = dev ? h.copy_to_device() : h;
}

// beta <- beta^{1/2}, beta_inv <- beta^{-1/2}      (sqrtm_cusolver :696-745, my_sqrtm_cusolver.hpp:366-376)
inline void sqrtm_cusolver(Vector<double> &, Dense_matrix<double> &beta, Dense_matrix<double> &beta_inv, cusolver_args<double> &)
{
    CUSOLVER_CHECK(lz_sqrtm(lanczos_context(), (int)beta.n_cols(), beta.data(), beta_inv.data()));
}
template <typename type_t>
void my_sqrtm_cusolver(Dense_matrix<type_t> &A, Dense_matrix<type_t> &A_inv)
{
    CUSOLVER_CHECK(lz_sqrtm(lanczos_context(), (int)A.n_cols(), reinterpret_cast<double *>(A.data()), reinterpret_cast<double *>(A_inv.data())));
}

#endif
