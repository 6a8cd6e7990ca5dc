// methods/block_lanczos.hpp -- block_lanczos / block_lanczos_blas<T> with the reference's signatures
// (methods/block_lanczos.hpp:13-24, :88-103).  alpha[0..m) and beta[0..m] are arrays of separately
// allocated device Dense_matrix blocks, as the harness builds them (test_lanczos.cu:215-223); the
// library works on contiguous block arrays, so the wrapper stages them and copies the blocks back.
// beta[0] = (B^T B)^{1/2}, beta[m] = the last inverse square root (the reference's scratch slot).
#ifndef lzb_block_lanczos_hpp
#define lzb_block_lanczos_hpp

#include "vector_lanczos.hpp"

template <typename type_t, typename Matrix>
void block_lanczos_blas(Matrix &A, Dense_matrix<type_t> &B, const unsigned int m, unsigned int lc, Vector<type_t> &q,
                        Dense_matrix<type_t> *alpha, Dense_matrix<type_t> *beta, Dense_matrix<type_t> & /*Q0*/, Dense_matrix<type_t> & /*Q1*/,
                        Dense_matrix<type_t> & /*W*/, cusolver_args<type_t> & /*args*/, Vector<type_t> /*eigen_val*/, cublasHandle_t /*cublasH*/,
                        const unsigned int /*n_blocks*/, const unsigned int /*n_loads*/)
{
    lzb::require_device_type<type_t>();
    const std::size_t bw = B.n_cols(), bb = bw * bw;
    Dense_matrix<type_t> a(bb, m, MemorySpace::CUDA), bt(bb, m + 1, MemorySpace::CUDA);
    const int rm = lzb::reorth_mode();     // the block driver knows none / full (CGS2) / DGKS; the vector-only selective mode maps to DGKS
    const int mode = rm == LZ_REORTH_NONE ? LZ_REORTH_NONE : (rm == LZ_REORTH_FULL || !(bw == 8 || bw == 16 || bw == 32)) ? LZ_REORTH_FULL : LZ_REORTH_FULL_DGKS;
    AssertCuda(lz_block_lanczos(lanczos_context(), A.device_operator(), reinterpret_cast<const double *>(B.data()), (int64_t)B.n_rows(), (int)bw,
                                (int)m, lc, mode, reinterpret_cast<double *>(a.data()), reinterpret_cast<double *>(bt.data()),
                                reinterpret_cast<double *>(q.data())));
    // the reference never looks at the eigen-solver's info (utils/lib_utils.hpp:650-745) and carries Inf/NaN on;
    // here a singular W^T W is reported (the coefficients before the named block are valid)
    int blocks_done = 0;
    if (lz_block_status(lanczos_context(), (int)m, &blocks_done) != LZ_OK)
        std::fprintf(stderr, "block_lanczos: %s (%d of %u blocks valid)\n", lz_last_error(), blocks_done, m);
    for (unsigned int j = 0; j < m; ++j) lzb::dcopy(alpha[j].data(), a.data() + j * bb, bb * sizeof(type_t), LZ_D2D);
    for (unsigned int j = 0; j <= m; ++j) lzb::dcopy(beta[j].data(), bt.data() + j * bb, bb * sizeof(type_t), LZ_D2D);
}

// the reference's own-kernel variant (float only there); same driver here
template <typename type_t, typename Matrix>
void block_lanczos(Matrix &A, Dense_matrix<type_t> &B, const unsigned int m, unsigned int lc, Vector<type_t> &q, Dense_matrix<type_t> *alpha,
                   Dense_matrix<type_t> *beta, Dense_matrix<type_t> &Q0, Dense_matrix<type_t> &Q1, Dense_matrix<type_t> &W,
                   const unsigned int n_blocks, const unsigned int n_loads)
{
    cusolver_args<type_t> args;
    Vector<type_t> ev(B.n_cols(), MemorySpace::Host);
    block_lanczos_blas<type_t>(A, B, m, lc, q, alpha, beta, Q0, Q1, W, args, ev, nullptr, n_blocks, n_loads);
}

#endif
