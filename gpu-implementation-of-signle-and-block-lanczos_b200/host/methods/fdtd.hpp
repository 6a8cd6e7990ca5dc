// methods/fdtd.hpp -- the harness' validator with the reference's names and signatures
// (methods/fdtd.hpp:6-56): Nsteps explicit Euler steps u <- u + dt * A u from u0 (dt = T_end / Nsteps),
// returning u(lc) (vector) or row lc of U (block).  A CUDA-space operator runs the whole loop inside the
// library (lz_fdtd_vector: one fused SpMV pass per step instead of spmv + Vector::add; lz_fdtd_block);
// a Host-space operator runs the reference's own loop over the Host containers.
#ifndef lzb_fdtd_hpp
#define lzb_fdtd_hpp

#include "../kernels/spmv_spmm.hpp"
#include "copy_functions.hpp"

template <typename type_t, typename Matrix>
type_t fdtd_vector(Matrix &A, Vector<type_t> &u0, const unsigned int Nsteps, const double T_end, const unsigned int lc)
{
    if (u0.memory_space() == MemorySpace::CUDA) {
        lzb::require_device_type<type_t>();
        double result = 0;
        AssertCuda(lz_fdtd_vector(lanczos_context(), A.device_operator(), reinterpret_cast<const double *>(u0.data()), (int64_t)Nsteps,
                                  T_end, (int64_t)lc, &result, nullptr));
        return (type_t)result;
    }
    type_t dt = (type_t)(T_end / Nsteps);
    Vector<type_t> dudt(u0);
    Vector<type_t> u(u0);
    for (unsigned int i = 0; i < Nsteps; ++i) {
        A.spmv(u, dudt);
        u.add(dt, dudt);
    }
    return u(lc);
}

template <typename type_t, typename Matrix>
Vector<type_t> ftdt_block(Matrix &A, Dense_matrix<type_t> &U0, const unsigned int Nsteps, const double T_end, unsigned int lc)
{
    if (U0.memory_space() == MemorySpace::CUDA) {
        lzb::require_device_type<type_t>();
        Vector<type_t> host((unsigned int)U0.n_cols(), MemorySpace::Host);
        AssertCuda(lz_fdtd_block(lanczos_context(), A.device_operator(), reinterpret_cast<const double *>(U0.data()), (int64_t)U0.n_rows(),
                                 (int)U0.n_cols(), (int64_t)Nsteps, T_end, (int64_t)lc, reinterpret_cast<double *>(host.data())));
        return host.copy_to_device();
    }
    type_t dt = (type_t)(T_end / Nsteps);
    Dense_matrix<type_t> dUdT(U0);
    Dense_matrix<type_t> U(U0);
    for (unsigned int i = 0; i < Nsteps; ++i) {
        A.spmm(U, dUdT);
        U.sadd(1, dt, dUdT);
    }
    Vector<type_t> result((unsigned int)U0.n_cols(), U0.memory_space());
    copy_row_to_vector(lc, 0, U, result);
    return result;
}

#endif
