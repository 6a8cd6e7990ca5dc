/*
 * lanczos_oracle.h -- CPU restatement of the reference's Lanczos hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the reported CPU baseline.
 *
 * Every function cites the reference lines (relative to /root/reference/source/)
 * it restates.  Parity status: PINNED -- the restatement is checked bit-for-bit
 * against the reference's own Host containers (oracle/_ref/ref_host_dump, built from
 * the reference headers where they lie) and against the committed fixtures in
 * tests/golden/ that the same binary minted (see oracle/make_goldens.py).
 */
#ifndef LANCZOS_ORACLE_H
#define LANCZOS_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* --- generators ---------------------------------- */
uint64_t orc_splitmix64(uint64_t x);
/* v[i] = 2*u(splitmix64(seed ^ i)) - 1,  u = (x >> 11) * 2^-53 */
void orc_start_vector(int64_t n, uint64_t seed, double *v);
/* column-major n x b block with leading dimension ld: V[i + c*ld] = 2*u(splitmix64(seed ^ (i*b+c))) - 1 */
void orc_start_block(int64_t n, int b, int64_t ld, uint64_t seed, double *V);
int64_t orc_lap2d_nnz(int64_t nx, int64_t ny);
void orc_lap2d_csr(int64_t nx, int64_t ny, int32_t *rowptr, int32_t *colidx, double *vals);
int64_t orc_lap3d_nnz(int64_t nx, int64_t ny, int64_t nz);
void orc_lap3d_csr(int64_t nx, int64_t ny, int64_t nz, int32_t *rowptr, int32_t *colidx, double *vals);
/* R-MAT edge list (directed, before symmetrisation): edge e -> (src[e], dst[e]) */
void orc_rmat_edges(int scale, int64_t n_edges, uint64_t seed, int32_t *src, int32_t *dst);

/* --- format helpers ------------------------------------------------------------------ */
/* column-major ELL (objects/ell_matrix.hpp:14-21: data[r + k*n]) -> CSR keeping ELL column
 * order inside each row and dropping explicit zeros.  Returns nnz. rowptr has n+1 entries. */
int64_t orc_ell_to_csr(int64_t n, int width, const double *ell_data, const uint32_t *ell_idx,
                       int32_t *rowptr, int32_t *colidx, double *vals);

/* --- operators ----------------------------------------------------------------------- */
/* y = A x, sequential per-row sum in storage order (objects/ell_matrix.hpp:246-251) */
void orc_csr_spmv(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
                  const double *x, double *y);
/* Y = A X, X/Y column-major with leading dimension ld (objects/ell_matrix.hpp:287-300) */
void orc_csr_spmm(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
                  int b, const double *X, int64_t ldx, double *Y, int64_t ldy);
double orc_dot(int64_t n, const double *x, const double *y);   /* objects/vector.hpp:268-274 */

/* --- dense helpers (utils/lib_utils.hpp semantics) ------------------------------------ */
/* R = T^T T  (b x b, column-major)            utils/lib_utils.hpp:80-123 */
void orc_mm_tt(int64_t n, int b, const double *T, int64_t ld, double *R);
/* R = 0.5 (T1^T T2 + T2^T T1)                 utils/lib_utils.hpp:126-202 */
void orc_mm_tt2(int64_t n, int b, const double *T1, int64_t ld1, const double *T2, int64_t ld2, double *R);
/* R = beta*R + alpha * T S   (T n x b, S b x b) utils/lib_utils.hpp:28-75 */
void orc_mm_ts(int64_t n, int b, double beta, double alpha, const double *T, int64_t ldt,
               const double *S, double *R, int64_t ldr);
/* symmetric eigen-decomposition by cyclic Jacobi: A (n x n col-major, destroyed) -> w ascending,
 * V columns.  Returns sweeps used. */
int orc_jacobi_eig(int n, double *A, double *w, double *V);
/* S <- (S)^{1/2}, Sinv <- S^{-1/2} via V sqrt(|L|) V^T     utils/lib_utils.hpp:650-745 */
void orc_sqrtm(int b, double *S, double *Sinv);

/* --- drivers ------------------------------------------------------------------------- */
/* methods/vector_lanczos.hpp:20-66.  reorth: 0 none, 1 full CGS2 against the stored basis
 * (extension; basis is internal).  alpha[m], beta[m] (beta[0] = ||b||), q[m] = row lc of basis.
 * Vout (optional, n*m column-major) receives the basis.  Returns steps completed
 * (< m on breakdown: beta_j == 0 or non-finite). */
int orc_vector_lanczos(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
                       const double *b, int m, int64_t lc, int reorth,
                       double *alpha, double *beta, double *q, double *Vout);

/* methods/block_lanczos.hpp:105-166 (the BLAS variant's semantics).  B: n x bw column-major
 * (ld = n).  alpha: m blocks bw*bw; beta: (m+1) blocks (beta[0] = (B^T B)^{1/2}, beta[m] = scratch
 * inverse, as in the reference); q: m*bw.  reorth: 0 none, 1 full block-CGS2 (extension).
 * Vout optional (n * m*bw, column-major). */
int orc_block_lanczos(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
                      const double *B, int bw, int m, int64_t lc, int reorth,
                      double *alpha, double *beta, double *q, double *Vout);

/* objects/tridiagonal_matrix.hpp:90-127 (CUDA branch semantics: all m blocks).
 * T: (m*bw)^2 column-major, zero-filled by the callee. */
void orc_assemble_T(int m, int bw, const double *alpha, const double *beta, double *T);

/* expm of a small symmetric matrix as the reference forms it: syevd then out[r,c] = sum_i V[r,i] exp(w_i) V[c,i]
 * (expm_cusolver, utils/lib_utils.hpp:542-590; custom_mult, kernels/dense_kernels.hpp:53-78).  In place. */
void orc_expm_sym(int n, double *T);
/* the harness post-processing (test_lanczos.cu:100-110, :270-283): T = t_end * Assemble_T(alpha, beta),
 * F1 = expm(T)[:, 0:bw] * beta_0, solution = F1^T q.  beta holds beta_0 .. beta_{m-1}. */
void orc_lanczos_solution(int m, int bw, const double *alpha, const double *beta, const double *q, double t_end,
                          double *solution);
/* methods/fdtd.hpp:6-31: nsteps times { dudt = A u; u += dt * dudt }, dt = t_end / nsteps; u in place */
void orc_fdtd_vector(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals, double *u,
                     int64_t nsteps, double t_end);
/* methods/fdtd.hpp:33-56, U column-major n x b with leading dimension ld, in place */
void orc_fdtd_block(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals, int b, double *U,
                    int64_t ld, int64_t nsteps, double t_end);

/* thread control for the baseline timing (1 => strictly sequential, reference summation order) */
void orc_set_threads(int t);
int orc_get_threads(void);

#ifdef __cplusplus
}
#endif
#endif
