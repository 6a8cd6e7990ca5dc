// lz_block.cu -- block path: SpMM kernels and the block Lanczos driver.
//
// Replaces ell::SpMM (kernels/spmv_spmm.hpp:137-199), lm::spmm_basic (kernels/ell_kernels.hpp:37-61)
// and block_lanczos / block_lanczos_blas (methods/block_lanczos.hpp:13-167).
//
// Internal layout of every n x bw block is ROW-MAJOR (bw contiguous doubles per row): one gathered
// row of the dense block is a single 64..256-byte segment, where the reference's column-major
// layout costs bw separate sectors per non-zero.  The C-ABI keeps the reference's
// column-major layout; the driver converts once on the way in.
#include <math.h>

#include "lz_dense_host.cuh"

#define SPMM_THREADS 256

// ---------------------------------------------------------------------------------------------
// Row-major SpMM, CSR:  W[i,:] = sum_k A[i,k] X[col_k,:]   (- Q0[i,:] B  when FUSE_SUB)
// A group of LW = BW/2 lanes owns a row; each lane carries two adjacent columns (128-bit loads of
// the gathered row, 128-bit stores).  Rows are dealt to groups in contiguous slabs so neighbouring
// rows -- which share most of their gathered rows for banded operators -- sit in the same warp/CTA
// and hit L1.  The B (bw x bw, column-major) operand of the fused subtraction lives in shared memory.
// ---------------------------------------------------------------------------------------------
template <int BW, bool FUSE_SUB>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_rm(int64_t n_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
          const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ W,
          const double *__restrict__ Q0, const double *__restrict__ Bm)
{
    constexpr int LW = BW / 2;                       // lanes per row
    constexpr int RPW = 32 / LW;                     // rows per warp at a time
    constexpr int BWP = BW + 1;                      // padded column stride: conflict-free across the group's lanes
    __shared__ double bs[FUSE_SUB ? BW * BWP : 1];
    if (FUSE_SUB) {
        for (int e = threadIdx.x; e < BW * BW; e += SPMM_THREADS) bs[(e % BW) + (e / BW) * BWP] = Bm[e];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, sub = lane / LW, l = lane % LW;
    const int64_t warp = (int64_t)blockIdx.x * (SPMM_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (SPMM_THREADS / 32);
    for (int64_t r0 = warp * RPW; r0 < n_rows; r0 += n_warps * RPW) {
        const int64_t r = r0 + sub;
        const bool valid = r < n_rows;               // no early exit: the group shuffles below need every lane
        const int s = valid ? rowptr[r] : 0, e = valid ? rowptr[r + 1] : 0;
        double a0 = 0.0, a1 = 0.0;
        int k = s;
        for (; k + 4 <= e; k += 4) {                 // four gathers in flight per lane
            int c[4]; double v[4]; double2 x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { c[u] = colidx[k + u]; v[u] = vals[k + u]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) x[u] = __ldg(reinterpret_cast<const double2 *>(X + (int64_t)c[u] * BW) + l);
#pragma unroll
            for (int u = 0; u < 4; ++u) { a0 = fma(v[u], x[u].x, a0); a1 = fma(v[u], x[u].y, a1); }
        }
        for (; k < e; ++k) {
            const double v = vals[k];
            const double2 x = __ldg(reinterpret_cast<const double2 *>(X + (int64_t)colidx[k] * BW) + l);
            a0 = fma(v, x.x, a0); a1 = fma(v, x.y, a1);
        }
        if (FUSE_SUB) {
            // W[i, 2l..2l+1] -= sum_p Q0[i,p] B[p, 2l..2l+1]; Q0 row broadcast through the group's lanes
            const double2 q = valid ? __ldg(reinterpret_cast<const double2 *>(Q0 + r * BW) + l) : make_double2(0.0, 0.0);
#pragma unroll
            for (int p = 0; p < LW; ++p) {
                const double q0 = __shfl_sync(0xffffffffu, q.x, sub * LW + p);
                const double q1 = __shfl_sync(0xffffffffu, q.y, sub * LW + p);
                a0 = fma(-q0, bs[(2 * p) + (2 * l) * BWP], a0);
                a1 = fma(-q0, bs[(2 * p) + (2 * l + 1) * BWP], a1);
                a0 = fma(-q1, bs[(2 * p + 1) + (2 * l) * BWP], a0);
                a1 = fma(-q1, bs[(2 * p + 1) + (2 * l + 1) * BWP], a1);
            }
        }
        if (valid) *(reinterpret_cast<double2 *>(W + r * BW) + l) = make_double2(a0, a1);
    }
}

// any bw <= 32 (odd widths, bw = 1): one lane per column
template <bool FUSE_SUB>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_rm_any(int bw, int64_t n_rows, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
              const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ W,
              const double *__restrict__ Q0, const double *__restrict__ Bm)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (SPMM_THREADS / 32) + (threadIdx.x >> 5);
    const int64_t n_warps = (int64_t)gridDim.x * (SPMM_THREADS / 32);
    for (int64_t r = warp; r < n_rows; r += n_warps) {
        if (lane >= bw) continue;
        double a = 0.0;
        for (int k = rowptr[r]; k < rowptr[r + 1]; ++k) a = fma(vals[k], __ldg(X + (int64_t)colidx[k] * bw + lane), a);
        if (FUSE_SUB)
            for (int p = 0; p < bw; ++p) a = fma(-Q0[r * bw + p], Bm[p + lane * bw], a);
        W[r * bw + lane] = a;
    }
}

// width-4 row-interleaved ELL (the reference's device format) as the operator of the row-major SpMM
template <int BW, bool FUSE_SUB>
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_rm_ell4(int64_t n_rows, const double *__restrict__ data, const uint32_t *__restrict__ idx,
               const double *__restrict__ X, double *__restrict__ W, const double *__restrict__ Q0,
               const double *__restrict__ Bm)
{
    const int64_t r = ((int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x) / BW;
    const int c = threadIdx.x % BW;                  // BW divides 256 for BW in {1,2,4,8,16,32}
    if (r >= n_rows) return;
    double a = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) a = fma(data[4 * r + k], __ldg(X + (int64_t)idx[4 * r + k] * BW + c), a);
    if (FUSE_SUB)
        for (int p = 0; p < BW; ++p) a = fma(-Q0[r * BW + p], Bm[p + c * BW], a);
    W[r * BW + c] = a;
}

// column-major SpMM of the C-ABI (reference layout): one thread per row keeps its row of A in
// registers and walks the b columns; gathers are coalesced across rows for banded operators.
__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_cm(int64_t n_rows, int b, const int32_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
          const double *__restrict__ vals, const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy)
{
    const int64_t r = (int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x;
    if (r >= n_rows) return;
    const int s = rowptr[r], e = rowptr[r + 1];
    if (e - s <= 8) {
        int c[8]; double v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = (s + k < e) ? colidx[s + k] : 0; v[k] = (s + k < e) ? vals[s + k] : 0.0; }
        for (int col = 0; col < b; ++col) {
            const double *xc = X + (int64_t)col * ldx;
            double t = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) if (s + k < e) t = __dadd_rn(t, __dmul_rn(v[k], __ldg(xc + c[k])));
            Y[r + (int64_t)col * ldy] = t;
        }
    } else {
        for (int col = 0; col < b; ++col) {
            const double *xc = X + (int64_t)col * ldx;
            double t = 0.0;
            for (int k = s; k < e; ++k) t = __dadd_rn(t, __dmul_rn(vals[k], __ldg(xc + colidx[k])));
            Y[r + (int64_t)col * ldy] = t;
        }
    }
}

__global__ void __launch_bounds__(SPMM_THREADS)
k_spmm_cm_ell4(int64_t n_rows, int b, const double *__restrict__ data, const uint32_t *__restrict__ idx,
               const double *__restrict__ X, int64_t ldx, double *__restrict__ Y, int64_t ldy)
{
    const int64_t r = (int64_t)blockIdx.x * SPMM_THREADS + threadIdx.x;
    if (r >= n_rows) return;
    double v0, v1, v2, v3;
    lz_ld256(data + 4 * r, v0, v1, v2, v3);
    const uint4 c = *reinterpret_cast<const uint4 *>(idx + 4 * r);
    for (int col = 0; col < b; ++col) {
        const double *xc = X + (int64_t)col * ldx;
        double t = __dmul_rn(v0, __ldg(xc + c.x));
        t = __dadd_rn(t, __dmul_rn(v1, __ldg(xc + c.y)));
        t = __dadd_rn(t, __dmul_rn(v2, __ldg(xc + c.z)));
        t = __dadd_rn(t, __dmul_rn(v3, __ldg(xc + c.w)));
        Y[r + (int64_t)col * ldy] = t;
    }
}

// column-major (ld) <-> row-major (bw) through a shared-memory tile
__global__ void __launch_bounds__(256) k_cm_to_rm(int64_t n, int bw, const double *__restrict__ src, int64_t ld, double *__restrict__ dst)
{
    __shared__ double t[32][33];
    const int64_t base = (int64_t)blockIdx.x * 32;
    for (int e = threadIdx.x; e < 32 * bw; e += 256) {
        const int r = e % 32, c = e / 32;
        t[r][c] = (base + r < n) ? src[base + r + (int64_t)c * ld] : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 32 * bw; e += 256) {
        const int r = e / bw, c = e % bw;
        if (base + r < n) dst[(base + r) * bw + c] = t[r][c];
    }
}

static int spmm_rm(lz_ctx *ctx, const lz_matrix *A, int bw, const double *X, double *W, const double *Q0, const double *Bm)
{
    const int64_t n = A->n_rows;
    const bool fuse = Q0 != nullptr;
    lz_prof_begin(ctx, LZ_K_SPMM, 12.0 * (double)A->nnz + 4.0 * (double)n + 16.0 * (double)n * bw + (fuse ? 8.0 * (double)n * bw : 0.0));
    if (A->format == LZ_FMT_ELL4) {
        const unsigned grid = (unsigned)((n * bw + SPMM_THREADS - 1) / SPMM_THREADS);
#define ELL_CASE(B)                                                                                               \
    if (fuse) k_spmm_rm_ell4<B, true><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n, A->ell_data, A->ell_idx, X, W, Q0, Bm); \
    else k_spmm_rm_ell4<B, false><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n, A->ell_data, A->ell_idx, X, W, Q0, Bm)
        if (bw == 1) { ELL_CASE(1); } else if (bw == 2) { ELL_CASE(2); } else if (bw == 4) { ELL_CASE(4); }
        else if (bw == 8) { ELL_CASE(8); } else if (bw == 16) { ELL_CASE(16); } else if (bw == 32) { ELL_CASE(32); }
        else { lz_set_error("block width %d is not supported on an ELL4 operator (use 1,2,4,8,16,32)", bw); return LZ_ERR_UNSUPPORTED; }
#undef ELL_CASE
    } else {
        int64_t want = (n + 63) / 64;
        int64_t cap = (int64_t)ctx->sm_count * 16;
        const unsigned grid = (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
#define CSR_CASE(B)                                                                                                     \
    if (fuse) k_spmm_rm<B, true><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n, A->rowptr, A->colidx, A->vals, X, W, Q0, Bm); \
    else k_spmm_rm<B, false><<<grid, SPMM_THREADS, 0, ctx->stream>>>(n, A->rowptr, A->colidx, A->vals, X, W, Q0, Bm)
        if (bw == 4) { CSR_CASE(4); } else if (bw == 8) { CSR_CASE(8); } else if (bw == 16) { CSR_CASE(16); }
        else if (bw == 32) { CSR_CASE(32); }
        else {
            if (fuse) k_spmm_rm_any<true><<<grid, SPMM_THREADS, 0, ctx->stream>>>(bw, n, A->rowptr, A->colidx, A->vals, X, W, Q0, Bm);
            else k_spmm_rm_any<false><<<grid, SPMM_THREADS, 0, ctx->stream>>>(bw, n, A->rowptr, A->colidx, A->vals, X, W, Q0, Bm);
        }
#undef CSR_CASE
    }
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return LZ_OK;
}

// basis block j (row-major) -> stored basis; block CGS sweep over the stored blocks
__global__ void k_flag_init(int *flags) { flags[0] = 0x7fffffff; flags[2] = 0x7fffffff; }

// C = V_j^T W for one stored block, then W -= V_j C   (classical block Gram-Schmidt, one block at a
// time would be modified GS; we compute all projections first to stay classical)
static int block_cgs_sweep(lz_ctx *ctx, int64_t n, int bw, int nblocks, const double *V, double *W, double *C)
{
    const size_t pan = (size_t)n * bw, bb = (size_t)bw * bw;
    for (int j = 0; j < nblocks; ++j) LZ_TRY(lz_gram(ctx, n, bw, true, V + pan * j, 0, W, 0, C + bb * j, 0));
    for (int j = 0; j < nblocks; ++j) LZ_TRY(lz_panel(ctx, n, bw, true, V + pan * j, 0, C + bb * j, 1.0, -1.0, W, 0, nullptr));
    return LZ_OK;
}

extern "C" {

int lz_spmm(lz_ctx *ctx, const lz_matrix *A, int b, const double *X, int64_t ldx, double *Y, int64_t ldy)
{
    LZ_CHECK(ctx && A && X && Y && b >= 1, LZ_ERR_INVALID, "lz_spmm: bad arguments");
    LZ_CHECK(ldx >= A->n_cols && ldy >= A->n_rows, LZ_ERR_INVALID, "lz_spmm: leading dimensions too small");
    LZ_CHECK(X != Y, LZ_ERR_INVALID, "lz_spmm: X and Y must not alias");
    const unsigned grid = (unsigned)((A->n_rows + SPMM_THREADS - 1) / SPMM_THREADS);
    lz_prof_begin(ctx, LZ_K_SPMM, 12.0 * (double)A->nnz + 4.0 * (double)A->n_rows + 16.0 * (double)A->n_rows * b);
    if (A->format == LZ_FMT_ELL4) k_spmm_cm_ell4<<<grid, SPMM_THREADS, 0, ctx->stream>>>(A->n_rows, b, A->ell_data, A->ell_idx, X, ldx, Y, ldy);
    else k_spmm_cm<<<grid, SPMM_THREADS, 0, ctx->stream>>>(A->n_rows, b, A->rowptr, A->colidx, A->vals, X, ldx, Y, ldy);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return LZ_OK;
}

int lz_block_lanczos(lz_ctx *ctx, const lz_matrix *A, const double *B, int64_t ldb, int bw, int m, int64_t lc,
                     int reorth, double *alpha, double *beta, double *q)
{
    LZ_CHECK(ctx && A && B && alpha && beta && m >= 1, LZ_ERR_INVALID, "lz_block_lanczos: bad arguments");
    LZ_CHECK(bw >= 1 && bw <= 32, LZ_ERR_INVALID, "lz_block_lanczos: block width %d outside 1..32", bw);
    const int64_t n = A->n_rows;
    LZ_CHECK(A->n_cols == n && ldb >= n, LZ_ERR_INVALID, "lz_block_lanczos: operator must be square and ldb >= n");
    LZ_CHECK(lc >= 0 && lc < n, LZ_ERR_INVALID, "lz_block_lanczos: lc out of range");
    LZ_CHECK(reorth == LZ_REORTH_NONE || reorth == LZ_REORTH_FULL, LZ_ERR_INVALID, "lz_block_lanczos: reorth mode %d", reorth);
    LZ_CUDA(cudaSetDevice(ctx->device));
    const size_t pan = (size_t)n * bw, bb = (size_t)bw * bw;
    size_t work_bytes = sizeof(double) * (3 * pan + (reorth ? bb * m : 0) + 64);
    void *work;
    LZ_TRY(lz_ctx_workspace(ctx, work_bytes, &work));
    double *Q0 = (double *)work, *Q1 = Q0 + pan, *W = Q1 + pan, *C = W + pan;
    double *V = nullptr;
    if (reorth) LZ_TRY(lz_ctx_basis(ctx, (int64_t)pan, m, &V));     // block j at V + j*pan, row-major
    double *binv = beta + bb * m;                                         // beta[m]: scratch inverse (block_lanczos.hpp:111)
    int *flag = ctx->flags + 2;
    k_flag_init<<<1, 1, 0, ctx->stream>>>(ctx->flags);
    LZ_LAUNCH_CHECK(ctx);

    // W <- B in row-major; beta[0] = (B^T B)^{1/2}; Q0 = B beta[0]^{-1}                     (:106-114)
    k_cm_to_rm<<<(unsigned)((n + 31) / 32), 256, 0, ctx->stream>>>(n, bw, B, ldb, W);
    LZ_LAUNCH_CHECK(ctx);
    LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, W, 0, beta, 0));
    LZ_TRY(lz_sqrtm_launch(ctx, bw, beta, binv, flag));
    LZ_TRY(lz_panel(ctx, n, bw, true, W, 0, binv, 0.0, 1.0, Q0, 0, nullptr));
    if (q) LZ_TRY(lz_copy_row_launch(ctx, lc, bw, true, Q0, 0, q, 0));                        // :117
    if (V) LZ_CUDA(cudaMemcpyAsync(V, Q0, sizeof(double) * pan, cudaMemcpyDeviceToDevice, ctx->stream));
    LZ_TRY(spmm_rm(ctx, A, bw, Q0, W, nullptr, nullptr));                                     // :121
    LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, Q0, 0, alpha, 1));                                 // :124
    double *G = ctx->scalars + 4096;                                                          // W^T W of the updated W
    LZ_TRY(lz_panel(ctx, n, bw, true, Q0, 0, alpha, 1.0, -1.0, W, 0, reorth ? nullptr : G));  // :128 (+ :137 fused)
    if (reorth) {
        LZ_TRY(block_cgs_sweep(ctx, n, bw, 1, V, W, C));
        LZ_TRY(block_cgs_sweep(ctx, n, bw, 1, V, W, C));
        LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, W, 0, G, 0));
    }
    for (int j = 1; j < m; ++j) {                                                             // :132-166
        double *bj = beta + bb * j, *aj = alpha + bb * j;
        LZ_CUDA(cudaMemcpyAsync(bj, G, sizeof(double) * bb, cudaMemcpyDeviceToDevice, ctx->stream));   // :137
        LZ_TRY(lz_sqrtm_launch(ctx, bw, bj, binv, flag));                                     // :142
        LZ_TRY(lz_panel(ctx, n, bw, true, W, 0, binv, 0.0, 1.0, Q1, 0, nullptr));             // :145
        LZ_TRY(spmm_rm(ctx, A, bw, Q1, W, Q0, bj));                                           // :149 + :152 fused
        LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, Q1, 0, aj, 1));                                // :155
        LZ_TRY(lz_panel(ctx, n, bw, true, Q1, 0, aj, 1.0, -1.0, W, 0, reorth ? nullptr : G)); // :159 (+ next :137)
        double *t = Q0; Q0 = Q1; Q1 = t;                                                      // :162 (no copy)
        if (q) LZ_TRY(lz_copy_row_launch(ctx, lc, bw, true, Q0, 0, q, (int64_t)j * bw));      // :165
        if (V) {
            LZ_CUDA(cudaMemcpyAsync(V + pan * j, Q0, sizeof(double) * pan, cudaMemcpyDeviceToDevice, ctx->stream));
            LZ_TRY(block_cgs_sweep(ctx, n, bw, j + 1, V, W, C));
            LZ_TRY(block_cgs_sweep(ctx, n, bw, j + 1, V, W, C));
            LZ_TRY(lz_gram(ctx, n, bw, true, W, 0, W, 0, G, 0));
        }
    }
    return LZ_OK;
}

}  // extern "C"
