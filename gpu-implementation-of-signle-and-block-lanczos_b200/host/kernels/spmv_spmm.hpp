// kernels/spmv_spmm.hpp -- the operator boundary: spmv(A, in, out) and spmm(A, in, out), same
// signatures as the reference (kernels/spmv_spmm.hpp:209-333), plus the Csr_matrix overloads.
// Forward to lz_spmv / lz_spmm (the hand-written sm_100a kernels); a Host-space matrix prints
// "implement later" and aborts, where the reference printed it and silently did nothing (:229-232).
#ifndef lzb_spmv_spmm_hpp
#define lzb_spmv_spmm_hpp

#include "../objects/csr_matrix.hpp"

template <typename Number>
void spmv(Ell_matrix<Number> &A, Vector<Number> &in, Vector<Number> &out)
{
    AssertCuda(lz_spmv(lanczos_context(), A.device_operator(), reinterpret_cast<const double *>(in.data()),
                       reinterpret_cast<double *>(out.data())));
}
template <typename Number>
void spmv(Csr_matrix<Number> &A, Vector<Number> &in, Vector<Number> &out)
{
    AssertCuda(lz_spmv(lanczos_context(), A.device_operator(), reinterpret_cast<const double *>(in.data()),
                       reinterpret_cast<double *>(out.data())));
}
template <typename Number>
void spmm(Ell_matrix<Number> &A, Dense_matrix<Number> &in, Dense_matrix<Number> &out)
{
    AssertCuda(lz_spmm(lanczos_context(), A.device_operator(), (int)in.n_cols(), reinterpret_cast<const double *>(in.data()),
                       (int64_t)in.n_rows(), reinterpret_cast<double *>(out.data()), (int64_t)out.n_rows()));
}
template <typename Number>
void spmm(Csr_matrix<Number> &A, Dense_matrix<Number> &in, Dense_matrix<Number> &out)
{
    AssertCuda(lz_spmm(lanczos_context(), A.device_operator(), (int)in.n_cols(), reinterpret_cast<const double *>(in.data()),
                       (int64_t)in.n_rows(), reinterpret_cast<double *>(out.data()), (int64_t)out.n_rows()));
}

#endif
