// objects/tridiagonal_matrix.hpp -- Assemble_T: dense (block-)tridiagonal T from the Lanczos
// coefficients (reference objects/tridiagonal_matrix.hpp:90-205).  Block version: blocks
// alpha[0..m), beta[1..m) (beta[b] is the coupling between block b-1 and b; upper block as is, lower
// block transposed -- the CUDA branch semantics, all m blocks).  Scalar version: correct tridiagonal
// (the reference's scalar Assemble_T copies the sub-diagonal into the diagonal).
#ifndef lzb_tridiagonal_matrix_hpp
#define lzb_tridiagonal_matrix_hpp

#include "dense_matrix.hpp"

template <typename Number>
Dense_matrix<Number> Assemble_T(const unsigned int n_blocks, Dense_matrix<Number> *diag_blocks, Dense_matrix<Number> *subdiag_blocks)
{
    const unsigned int bd = diag_blocks[0].n_rows();
    const MemorySpace mem = diag_blocks[0].memory_space();
    const std::size_t N = (std::size_t)n_blocks * bd, bb = (std::size_t)bd * bd;
    Dense_matrix<Number> T(N, N, mem);
    if (mem == MemorySpace::CUDA) {
        // gather the separately allocated blocks into contiguous device arrays for the C-ABI
        Dense_matrix<Number> a(bb, n_blocks, mem), b(bb, n_blocks, mem);
        for (unsigned int k = 0; k < n_blocks; ++k) {
            lzb::dcopy(a.data() + k * bb, diag_blocks[k].data(), bb * sizeof(Number), LZ_D2D);
            if (k >= 1) lzb::dcopy(b.data() + k * bb, subdiag_blocks[k].data(), bb * sizeof(Number), LZ_D2D);
        }
        AssertCuda(lz_assemble_T(lanczos_context(), (int)n_blocks, (int)bd, reinterpret_cast<const double *>(a.data()),
                                 reinterpret_cast<const double *>(b.data()), reinterpret_cast<double *>(T.data())));
        return T;
    }
    for (unsigned int k = 0; k < n_blocks; ++k)
        for (std::size_t i = 0; i < bb; ++i) {
            const std::size_t r = i % bd, c = i / bd;
            T((k * bd + r) + (k * bd + c) * N) = diag_blocks[k](i);
            if (k >= 1) {
                T(((k - 1) * bd + r) + (k * bd + c) * N) = subdiag_blocks[k](i);
                T((k * bd + c) + ((k - 1) * bd + r) * N) = subdiag_blocks[k](i);
            }
        }
    return T;
}

// scalar series (host arrays, as test_lanczos.cu:66-67 keeps them): T(j,j) = alpha[j], T(j,j+1) = T(j+1,j) = beta[j+1]
template <typename Number>
Dense_matrix<Number> Assemble_T(const unsigned int n_entries, const Number *alpha, const Number *beta, MemorySpace mem = MemorySpace::Host)
{
    Dense_matrix<Number> T(n_entries, n_entries, MemorySpace::Host);
    for (unsigned int j = 0; j < n_entries; ++j) {
        T(j + (std::size_t)j * n_entries) = alpha[j];
        if (j + 1 < n_entries) {
            T(j + (std::size_t)(j + 1) * n_entries) = beta[j + 1];
            T((j + 1) + (std::size_t)j * n_entries) = beta[j + 1];
        }
    }
    return mem == MemorySpace::CUDA ? T.copy_to_device() : T;
}

#endif
