/*
 * lanczos_b200.h -- C-ABI of the B200-native (sm_100a) single-vector / block Lanczos path.
 *
 * Drop-in boundary for the hot path of ibrohimmn1994/GPU-implementation-of-signle-and-block-Lanczos:
 * every entry point names the reference interface (file:line under source/) it replaces.  The
 * reference has no FFI of its own (one nvcc translation unit of headers); the C++ mirror under
 * gpu-implementation-of-signle-and-block-lanczos_b200/host/ keeps the reference's class and function
 * names and forwards raw pointers here.  See INTEGRATION.md for the binding a maintainer adds.
 *
 * Conventions
 *   - plain C types only; all array arguments are DEVICE pointers unless the name ends in _host;
 *   - buffers passed in are BORROWED (never freed, never reallocated); scratch belongs to the
 *     lz_ctx and is (re)sized outside timed regions by the *_workspace / first call;
 *   - dense blocks use the reference layout: column-major, leading dimension ld (dense_matrix.hpp:9);
 *   - every call returns LZ_OK (0) or a negative status; lz_last_error() gives the message
 *     (reference policy: AssertCuda aborts, CUBLAS_CHECK throws -- utils/common.hpp:83-112; the C++
 *     mirror maps non-zero statuses back onto abort/throw);
 *   - calls enqueue on the context's stream and return without synchronising unless they hand a
 *     result back in host memory; one host thread per context;
 *   - there is NO CPU fallback: without a CUDA device every compute entry point fails with
 *     LZ_ERR_CUDA.
 */
#ifndef LANCZOS_B200_H
#define LANCZOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LZ_OK 0
#define LZ_ERR_INVALID (-1)   /* bad argument                                   */
#define LZ_ERR_CUDA (-2)      /* CUDA runtime / launch failure (AssertCuda)     */
#define LZ_ERR_ALLOC (-3)     /* device allocation failed                       */
#define LZ_ERR_BREAKDOWN (-4) /* non-finite norm (vector.hpp:233-244 aborts)    */
#define LZ_ERR_COMM (-5)      /* NCCL / rendezvous failure                      */
#define LZ_ERR_UNSUPPORTED (-6)

/* reorthogonalisation modes (extension: the reference has none) */
#define LZ_REORTH_NONE 0
#define LZ_REORTH_FULL 1      /* classical Gram-Schmidt, always two sweeps (CGS2)             */
#define LZ_REORTH_FULL_DGKS 2 /* second sweep only when the first one removed > 1 - 1/sqrt2   */
#define LZ_REORTH_SELECTIVE 3 /* partial reorthogonalisation: the omega recurrence (device-side) decides per step
                                whether w is reorthogonalised against the stored basis (CGS2) -- semi-orthogonal
                                basis, Ritz values to working precision, a fraction of the basis traffic           */

/* lz_memcpy kinds */
#define LZ_H2D 1
#define LZ_D2H 2
#define LZ_D2D 3

typedef struct lz_ctx lz_ctx;       /* device, stream, scratch, (optional) communicator */
typedef struct lz_matrix lz_matrix; /* a sparse operator resident on the device         */

/* ---- library / context ---------------------------------------------------------------- */
int lz_version(void);
const char *lz_last_error(void);
/* stream: a cudaStream_t cast to void* (NULL = the legacy default stream the reference uses,
 * test_lanczos.cu:74-91) */
int lz_ctx_create(int device, void *stream, lz_ctx **out);
int lz_ctx_destroy(lz_ctx *ctx);
int lz_ctx_sync(lz_ctx *ctx);                       /* cudaDeviceSynchronize role, test_lanczos.cu:239 */
int lz_ctx_device(const lz_ctx *ctx);
/* number of kernels this context has launched (bench.py's gpu_launches) */
int64_t lz_ctx_launch_count(const lz_ctx *ctx);

/* per-kernel-class timing with CUDA events on the context's stream (measurement only; bench.py).
 * classes: 0 spmv(+fused), 1 pass B, 2 CGS project, 3 CGS update, 4 spmm, 5 gram, 6 panel, 7 small, 8 comm,
 * 9 fused CGS update+project; the three output arrays have LZ_PROFILE_CLASSES entries */
#define LZ_PROFILE_CLASSES 10
int lz_ctx_profile(lz_ctx *ctx, int enable);
int lz_ctx_profile_read(lz_ctx *ctx, int64_t *launches, double *ms, double *bytes);

/* raw device memory for the C++ containers (objects/vector.hpp:24-39 cudaMalloc/cudaFree) */
int lz_malloc(lz_ctx *ctx, size_t bytes, void **dptr);
int lz_free(lz_ctx *ctx, void *dptr);
int lz_memcpy(lz_ctx *ctx, void *dst, const void *src, size_t bytes, int kind);  /* synchronous */
int lz_memset(lz_ctx *ctx, void *dptr, int value, size_t bytes);
int lz_fill(lz_ctx *ctx, int64_t n, double value, double *x);   /* v::set_entries, vector_kernels.hpp:11-20 */

/* ---- sparse operators ------------------------------------------------------------------ */
/* CSR (new container in the reference's style).  Arrays are borrowed device
 * pointers: rowptr[n_rows+1], colidx[nnz] (int32), vals[nnz] (fp64).  Builds the row-block
 * schedule (short rows streamed through shared memory, long rows split).
 * The arrays must stay alive AND UNCHANGED while the operator exists: the operator keeps derived, re-ordered copies of
 * them (length-binned virtual rows of power-law operators; the chunk-ordered copy of the operand-staging SpMM, built by
 * the first panel product), some of them lazily -- re-create the operator after changing a value.  One operator is used
 * from one host thread at a time (the lazily built schedules are not locked). */
int lz_csr_create(lz_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz, const int32_t *rowptr,
                  const int32_t *colidx, const double *vals, lz_matrix **out);
/* same from HOST arrays: the library owns the device copy (Csr_matrix::copy_to_device role) */
int lz_csr_create_host(lz_ctx *ctx, int64_t n_rows, int64_t n_cols, int64_t nnz,
                       const int32_t *rowptr_host, const int32_t *colidx_host,
                       const double *vals_host, lz_matrix **out);
/* ELLPACK as the reference stores it (objects/ell_matrix.hpp:14-21): data[T], idx[unsigned].
 * layout 0 = column-major data[r + k*n_rows]; layout 1 = row-interleaved data[width*r + k]
 * (what change_order(4) is meant to produce, ell_matrix.hpp:362-403).  Borrowed device arrays;
 * width 4 row-interleaved runs the dedicated kernel that replaces ell::SpMV / ell::SpMM
 * (kernels/spmv_spmm.hpp:105-199); anything else is converted to CSR on the device. */
int lz_ell_create(lz_ctx *ctx, int64_t n_rows, int64_t n_cols, int width, int layout,
                  const double *data, const uint32_t *idx, lz_matrix **out);
int lz_matrix_destroy(lz_matrix *A);
int lz_matrix_info(const lz_matrix *A, int64_t *n_rows, int64_t *n_cols, int64_t *nnz);
/* Host logic of the box-shaped chunks of the operand-staging SpMM (new; pure host code, no device needed -- exported so
 * that it can be tested and reused).  lz_grid_strides_host: nested strides 1 | nx | nx*ny of a structured-grid operator
 * from a SAMPLE of its rows (host CSR arrays of rows [first_row, first_row + sample_rows): rowptr_host has
 * sample_rows + 1 entries, colidx_host[k] is the column of entry rowptr_host[0] + k); returns the number of strides found (0..3)
 * or a negative status.  lz_box_order_host: the rows of that grid in box order (boxes of lx x ty x tz points, clipped at
 * the faces): rowmap_host[n_rows], chunk_row_host[*n_chunks + 1] (chunk_cap entries available). */
int lz_grid_strides_host(int64_t sample_rows, int64_t first_row, const int32_t *rowptr_host, const int32_t *colidx_host,
                         int64_t strides[3]);
int lz_box_order_host(int64_t n_rows, int n_strides, const int64_t strides[3], int lx, int ty, int tz, int32_t *rowmap_host,
                      int64_t chunk_cap, int32_t *chunk_row_host, int64_t *n_chunks);
/* Which panel-product kernel lz_block_lanczos / lz_fdtd_block run on A at width bw (new; the reference has one SpMM,
 * kernels/spmv_spmm.hpp:137-199).  Builds the lazily built schedules if needed.  kind: 0 gathering kernel, 1 operand-
 * staging kernel with chunks of consecutive rows, 2 operand-staging kernel with box-shaped chunks (box[3] = rows along the
 * unit stride, runs along the 2nd and 3rd stride; window_rows = largest number of X rows a chunk stages). */
int lz_matrix_spmm_schedule(lz_ctx *ctx, const lz_matrix *A, int bw, int *kind, int box[3], int *window_rows);
/* borrowed views of the CSR arrays held by A (NULL for a native ELL4 operator) */
int lz_matrix_csr_view(const lz_matrix *A, const int32_t **rowptr, const int32_t **colidx,
                       const double **vals);

/* synthetic operators generated on the device (BASELINE.json configs 2-5) */
int lz_gen_laplacian2d(lz_ctx *ctx, int64_t nx, int64_t ny, lz_matrix **out);
int lz_gen_laplacian3d(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, lz_matrix **out);
/* The reference's own test operator, assembled on the device: A = D W of Matrix_A(Nx, Ny, Nz)
 * (matrix_a/build_A_ell.hpp:8-255 followed by Ell_matrix::mult_diagonal, test_lanczos.cu:29-49) -- the 3-D Maxwell curl
 * operator on a staggered grid, n = 3 N (N+1)(2N+1) rows for Nx = Ny = Nz = N, width-4 ELL.  Every entry comes from
 * the closed form of the Kronecker products in the reference's operation order: the arrays are bit-identical to the
 * host builder's.  The result is a native row-interleaved ELL4 operator (lz_matrix_ell_view shows its arrays). */
int lz_gen_maxwell(lz_ctx *ctx, int Nx, int Ny, int Nz, lz_matrix **out);
/* borrowed views of the row-interleaved ELL arrays of a width-4 ELL operator (NULL for CSR operators) */
int lz_matrix_ell_view(const lz_matrix *A, const double **data, const uint32_t **idx);
/* directed R-MAT edge list, (a,b,c,d) = (0.57,0.19,0.19,0.05), counter-based RNG (config 4); the caller
 * symmetrises / de-duplicates / adds the Laplacian diagonal (tools/rmat.py) and hands the CSR to lz_csr_create */
int lz_gen_rmat_edges(lz_ctx *ctx, int scale, int64_t n_edges, uint64_t seed, int32_t *src, int32_t *dst);
/* v[i] = 2*u(splitmix64(seed ^ i)) - 1 ; block: V[i + c*ld] = 2*u(splitmix64(seed ^ (i*b+c))) - 1 */
int lz_gen_start_vector(lz_ctx *ctx, int64_t n, uint64_t seed, double *v);
int lz_gen_start_block(lz_ctx *ctx, int64_t n, int b, int64_t ld, uint64_t seed, double *V);

/* y = A x          replaces spmv(Ell_matrix&,Vector&,Vector&)  kernels/spmv_spmm.hpp:209-260
 *                  and Ell_matrix::spmv                        objects/ell_matrix.hpp:228-245 */
int lz_spmv(lz_ctx *ctx, const lz_matrix *A, const double *x, double *y);
/* Y = A X (b columns, column-major)   replaces spmm(...)       kernels/spmv_spmm.hpp:262-333
 *                  and Ell_matrix::spmm                        objects/ell_matrix.hpp:267-286 */
int lz_spmm(lz_ctx *ctx, const lz_matrix *A, int b, const double *X, int64_t ldx, double *Y, int64_t ldy);

/* ---- vector reductions / updates (kernels/vector_kernels.hpp, utils/lib_utils.hpp:431-538) ---- */
/* result to HOST (synchronises, as Vector::dot does: objects/vector.hpp:249-275) */
int lz_dot(lz_ctx *ctx, int64_t n, const double *x, const double *y, double *result_host);
/* l2 norm; LZ_ERR_BREAKDOWN when not finite (objects/vector.hpp:233-244) */
int lz_nrm2(lz_ctx *ctx, int64_t n, const double *x, double *result_host);
/* y = a*y + b*x     Vector::sadd / v::vector_update   vector_kernels.hpp:22-33 */
int lz_axpby(lz_ctx *ctx, int64_t n, double a, double *y, double b, const double *x);

/* ---- tall-skinny dense products (kernels/mm_tt.hpp, mm_tt2.hpp, mm_ts.hpp; lib_utils.hpp:28-202) ---- */
/* R = T^T T            (b x b, column-major, device)   mm_tt / mm_tt_cublas */
int lz_mm_tt(lz_ctx *ctx, int64_t n, int b, const double *T, int64_t ld, double *R);
/* R = 0.5 (T1^T T2 + T2^T T1)                          mm_tt2 / mm_tt2_cublas */
int lz_mm_tt2(lz_ctx *ctx, int64_t n, int b, const double *T1, int64_t ld1, const double *T2,
              int64_t ld2, double *R);
/* R = beta R + alpha T S   (T,R: n x b; S: b x b device) mm_ts / mm_cublas; R may alias T */
int lz_mm_ts(lz_ctx *ctx, int64_t n, int b, double beta, double alpha, const double *T, int64_t ldt,
             const double *S, double *R, int64_t ldr);
/* S <- S^{1/2}, Sinv <- S^{-1/2} (SPD b x b, reads the lower triangle, abs of eigenvalues)
 *                     my_sqrtm_cusolver / sqrtm_cusolver  kernels/my_sqrtm_cusolver.hpp:366-376,
 *                                                         utils/lib_utils.hpp:650-745 */
int lz_sqrtm(lz_ctx *ctx, int b, double *S, double *Sinv);
/* q[off + c] = Q[lc + c*ld], c < b     copy_row_to_vector  methods/copy_functions.hpp:31-46 */
int lz_copy_row(lz_ctx *ctx, int64_t lc, int b, const double *Q, int64_t ld, double *q, int64_t off);
/* dense block-tridiagonal T ((m*b)^2, column-major, device) from alpha[m], beta[1..m-1] blocks
 * stored back to back            Assemble_T  objects/tridiagonal_matrix.hpp:90-127 */
int lz_assemble_T(lz_ctx *ctx, int m, int b, const double *alpha, const double *beta, double *T);

/* ---- drivers ---------------------------------------------------------------------------- */
/* vector_lanczos<double>  methods/vector_lanczos.hpp:8-67.
 *   b      : start vector (device, n), not modified
 *   m      : steps;  lc: receiver row (copy_vector_element, copy_functions.hpp:116-133)
 *   reorth : LZ_REORTH_*  (full modes keep an n x m basis slab inside the context)
 *   alpha_host[m], beta_host[m] (beta[0] = ||b||): HOST arrays as in test_lanczos.cu:66-67
 *   q      : device, m entries (row lc of the Krylov basis)
 *   steps_done: number of valid coefficients (< m after a breakdown; then LZ_ERR_BREAKDOWN)
 * One fused SpMV pass (w = A q_j - beta_j q_{j-1}, alpha_j in the epilogue) and one fused update
 * pass (w -= alpha_j q_j, ||w||^2 in the epilogue) per step; scalars stay on the device. */
int lz_vector_lanczos(lz_ctx *ctx, const lz_matrix *A, const double *b, int m, int64_t lc, int reorth,
                      double *alpha_host, double *beta_host, double *q, int *steps_done);
/* same, but coefficients stay in DEVICE arrays and nothing synchronises (bench / graph use) */
int lz_vector_lanczos_async(lz_ctx *ctx, const lz_matrix *A, const double *b, int m, int64_t lc,
                            int reorth, double *alpha_dev, double *beta_dev, double *q);
/* Pre-size what the drivers would allocate on their first call for this operator and these sizes (work vectors / panels,
 * basis slab, reduction scratch), so that a timed first call -- the reference harness times exactly one cold call,
 * test_lanczos.cu:74-91 -- measures the iteration and not cudaMalloc.  Optional; unsharded contexts. */
int lz_vector_lanczos_workspace(lz_ctx *ctx, const lz_matrix *A, int m, int reorth);
int lz_block_lanczos_workspace(lz_ctx *ctx, const lz_matrix *A, int bw, int m, int reorth);
/* The same run in pieces (extension).  begin: sets up a run of at most m_capacity steps (work vectors,
 * basis slab for the reorthogonalising modes, beta_0 = ||b||).  advance: `steps` more Lanczos steps, then ALL
 * coefficients so far to the host (alpha_host / beta_host: m_capacity entries, or NULL) and the number of valid steps
 * (LZ_ERR_BREAKDOWN as in lz_vector_lanczos).  lz_vector_lanczos = begin + advance(m). */
int lz_vector_lanczos_begin(lz_ctx *ctx, const lz_matrix *A, const double *b, int m_capacity, int64_t lc, int reorth,
                            double *q);
int lz_vector_lanczos_advance(lz_ctx *ctx, int steps, double *alpha_host, double *beta_host, int *steps_done);
/* Save the current run -- (q_{j-1}, q_j, alpha, beta, 1/beta, j, the basis columns, the omega rows of a selective run)
 * -- to a file, and recreate it on a context (this or another one, another process) for the same operator; the
 * continued run reproduces the uninterrupted one bit for bit. */
int lz_vector_checkpoint_save(lz_ctx *ctx, const char *path);
int lz_vector_checkpoint_load(lz_ctx *ctx, const lz_matrix *A, const char *path, double *q);
/* Ritz vectors X[:, 0..k) = V_j Y of the current run (full / selective reorthogonalisation keeps V): Y_host is the
 * j x k column-major matrix of eigenvectors of T (j = steps done), X device column-major with leading dimension ldx.
 * A tall-skinny fp64 tensor-core product (mma.m8n8k4) over the row-tiled basis. */
int lz_vector_ritz_vectors(lz_ctx *ctx, int k, const double *Y_host, double *X, int64_t ldx);
/* number of steps of the current LZ_REORTH_SELECTIVE run that reorthogonalised (synchronises) */
int lz_vector_reorth_count(lz_ctx *ctx, int *count);
/* Thick-restart Lanczos (extension): k extremal eigenpairs inside a basis of m_max vectors.
 *   which: 0 smallest, 1 largest, 2 both ends (k/2 smallest + k - k/2 largest, as lz_ritz);  tol: a pair is converged
 *   when its residual estimate |beta_m y_m| <= tol * max|theta|;  at most max_restarts compressions of the basis.
 *   theta_host[k] ascending, resid_host[k] (or NULL) the estimates, X (or NULL): device n x k column-major Ritz vectors.
 *   info4 (or NULL): converged pairs, restarts, operator applications, basis size. */
int lz_eigs_thick_restart(lz_ctx *ctx, const lz_matrix *A, const double *b, int k, int which, int m_max, double tol,
                          int max_restarts, double *theta_host, double *resid_host, double *X, int64_t ldx, int *info4);
/* Block thick-restart Lanczos (BASELINE config 3: block size 16, k = 64): the same for multiple / clustered eigenvalues,
 * which a single Lanczos vector cannot resolve.  B: n x bw start block (column-major, ldb), bw in {8, 16, 32}; the basis
 * holds p_blocks blocks (p_blocks >= ceil(k / bw) + 3).  Block recurrence with block CGS2 against all stored blocks
 * (fp64 tensor-core projection / update), host eigensolve of the (p_blocks bw)^2 projected matrix, DMMA compression of
 * the basis to the kept Ritz vectors.  Outputs as lz_eigs_thick_restart; residual estimate ||beta_p Y_p||. */
int lz_block_eigs_thick_restart(lz_ctx *ctx, const lz_matrix *A, const double *B, int64_t ldb, int bw, int k, int which,
                                int p_blocks, double tol, int max_restarts, double *theta_host, double *resid_host,
                                double *X, int64_t ldx, int *info4);
/* Krylov basis kept by the last full-reorth run (stored row-tiled inside the context):
 * copy columns j0 .. j0+ncols-1 into dst (device, column-major, leading dimension ldd >= rows) */
int lz_vector_basis_info(lz_ctx *ctx, int64_t *rows, int *cols);
int lz_vector_basis_copy(lz_ctx *ctx, int j0, int ncols, double *dst, int64_t ldd);

/* block_lanczos_blas<double>  methods/block_lanczos.hpp:88-167 (and block_lanczos :13-80).
 *   B      : n x bw start block, column-major, leading dimension ldb (device), not modified
 *   alpha  : device, m blocks of bw*bw (column-major each), alpha[j] at alpha + j*bw*bw
 *   beta   : device, (m+1) blocks; beta[0] = (B^T B)^{1/2}, beta[m] = last inverse square root
 *            (scratch slot exactly as the reference uses it, block_lanczos.hpp:111,142)
 *   q      : device, m*bw entries (row lc of every Q_j)
 * bw in {1..32}.  With a communicator attached (lz_comm_init) A is this rank's row slab, B its local rows
 * (lc = -1, q = NULL): halo rows of the panels are exchanged before every SpMM and every b x b Gram matrix is
 * all-reduced, so alpha/beta are identical on every rank. */
int lz_block_lanczos(lz_ctx *ctx, const lz_matrix *A, const double *B, int64_t ldb, int bw, int m,
                     int64_t lc, int reorth, double *alpha, double *beta, double *q);

/* Outcome of the last lz_block_lanczos run on this context (synchronises).  *blocks_done = m when every
 * W^T W was positive definite to working precision; otherwise the index j of the first singular / non-finite
 * beta_j -- alpha[0..j) and beta[0..j] are valid -- and the call returns LZ_ERR_BREAKDOWN.  The reference never
 * looks (cusolver info is ignored, utils/lib_utils.hpp:650-745); the vector driver reports the same through
 * steps_done. */
int lz_block_status(lz_ctx *ctx, int m, int *blocks_done);
/* beta_m of the last run on this context -- the coupling to the next, unbuilt Lanczos vector / block -- copied to
 * HOST (bw*bw doubles, column-major; bw = 1 after a vector run).  Input of lz_ritz's residual estimate. */
int lz_last_coupling(lz_ctx *ctx, int bw, double *beta_last_host);

/* ---- Ritz extraction (the syevd(T) inside expm_cusolver, lib_utils.hpp:542-590) ---- */
/* alpha_host/beta_host: m blocks of bw*bw each (bw = 1: the scalar series; beta[0] is ignored, the
 * coupling blocks are beta[1..m-1]); beta_last_host: the bw*bw block beta_m coupling to the next
 * (unbuilt) block, or NULL.  k extremal Ritz values: k/2 smallest then k-k/2 largest, ascending;
 * resid[i] = || beta_last * Y[last block, i] ||  (0 when beta_last is NULL). Host computation on
 * the tiny projected matrix. */
int lz_ritz(int m, int bw, const double *alpha_host, const double *beta_host,
            const double *beta_last_host, int k, double *theta_host, double *resid_host);

/* ---- application path of the harness ---- */
/* T <- expm(T) for a small symmetric matrix (n x n, column-major, host; only the lower triangle is read):
 * V exp(Lambda) V^T, the construction of expm_cusolver + custom_mult (utils/lib_utils.hpp:542-590,
 * kernels/dense_kernels.hpp:53-78).  Host arithmetic inside the library, n <= 2048. */
int lz_expm_sym(int n, double *T_host);
/* solution[bw] = q^T expm(t_end * T)[:, 0:bw] beta_0 with T assembled from alpha/beta exactly as
 * Assemble_T does (test_lanczos.cu:100-110 for bw = 1, :270-283 for blocks).  alpha_host: m blocks,
 * beta_host: beta_0 .. beta_{m-1} (at least m blocks), q_host: the m*bw receiver-row entries. */
int lz_lanczos_solution(int m, int bw, const double *alpha_host, const double *beta_host, const double *q_host,
                        double t_end, double *solution_host);
/* fdtd validator (methods/fdtd.hpp:6-31): nsteps explicit Euler steps u <- u + (t_end/nsteps) A u from u0
 * (device, n, not modified), one fused SpMV pass per step; *result_host = u[lc] (lc = -1 and NULL to skip),
 * u_out (device, n) receives the final vector when not NULL. */
int lz_fdtd_vector(lz_ctx *ctx, const lz_matrix *A, const double *u0, int64_t nsteps, double t_end, int64_t lc,
                   double *result_host, double *u_out);
/* block variant (ftdt_block, methods/fdtd.hpp:33-56): U0 column-major n x bw with leading dimension ldu;
 * result_host[bw] = row lc of the final panel. */
int lz_fdtd_block(lz_ctx *ctx, const lz_matrix *A, const double *U0, int64_t ldu, int bw, int64_t nsteps,
                  double t_end, int64_t lc, double *result_host);

/* ---- multi-GPU: one process per GPU, rows of A and of every Krylov vector sharded ---- */
/* opaque 128-byte NCCL id minted by rank 0 and broadcast by the launcher (torch.distributed) */
int lz_comm_unique_id(void *id128_host);
int lz_comm_init(lz_ctx *ctx, int world_size, int rank, const void *id128_host);
int lz_comm_destroy(lz_ctx *ctx);
/* *peer_mode = 1 when small reductions and halo exchanges run over CUDA-IPC peer memory (NVLink stores + flags)
 * instead of NCCL; *timed_out = 1 when a peer-memory wait gave up because a rank disappeared (every result since
 * is invalid).  Synchronises.  LZ_COMM=1 in the environment forces NCCL, LZ_COMM=2 makes peer memory mandatory. */
int lz_comm_status(lz_ctx *ctx, int *peer_mode, int *timed_out);
/* contiguous row-block partition (host logic): rows are dealt in whole granules (a grid plane /
 * line for the stencil operators, 1 for anything else); rank r owns rows [begin,end) */
int lz_partition_rows(int64_t n_rows, int64_t granule, int world_size, int rank, int64_t *begin, int64_t *end);
/* local slab of the 7-/5-point Laplacian: rows [begin,end) = lz_partition_rows(n, plane, ...) with
 * column ids shifted to the index space [lower halo | local | upper halo]; halo = one xy-plane
 * (3-D) / one x-line (2-D) per existing neighbour */
int lz_gen_laplacian3d_shard(lz_ctx *ctx, int64_t nx, int64_t ny, int64_t nz, int world_size,
                             int rank, lz_matrix **out);
int lz_gen_laplacian2d_shard(lz_ctx *ctx, int64_t nx, int64_t ny, int world_size, int rank,
                             lz_matrix **out);
/* row slab of ANY square operator from HOST CSR arrays (the sharded twin of lz_csr_create_host): rows
 * [row_begin, row_begin + n_local) of a global_rows x global_rows operator, column ids already shifted into the local
 * index space [lower halo | local | upper halo] with halo_lo / halo_hi entries owned by ranks r-1 / r+1 (contiguous
 * row-block partitions of banded operators).  Rows [0, bnd_lo_rows) are those that touch the lower halo, rows
 * [bnd_hi_rows, n_local) the upper one: the chunks in between run while the halo exchange is in flight. */
int lz_csr_create_shard_host(lz_ctx *ctx, int64_t n_local, int64_t nnz, const int32_t *rowptr_host,
                             const int32_t *colidx_host, const double *vals_host, int64_t halo_lo, int64_t halo_hi,
                             int64_t global_rows, int64_t row_begin, int64_t bnd_lo_rows, int64_t bnd_hi_rows,
                             lz_matrix **out);
/* sharded single-vector Lanczos: b_local holds this rank's rows; halo exchange with the two
 * neighbouring ranks before every SpMV, packed all-reduce of the alpha / beta^2 partials.
 * alpha/beta (device, m) are identical on every rank. */
int lz_vector_lanczos_sharded(lz_ctx *ctx, const lz_matrix *A_local, const double *b_local, int m,
                              int reorth, double *alpha_dev, double *beta_dev);

#ifdef __cplusplus
}
#endif
#endif /* LANCZOS_B200_H */
