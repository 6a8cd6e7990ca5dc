// lz_dense.cu -- host launchers and C-ABI entry points of the tall-skinny dense products, the
// b x b matrix square root, and the small helpers of the block path (row extraction, T assembly).
#include <math.h>

#include "lz_dense_host.cuh"

// ---------------------------------------------------------------------------------------------
// launch helpers (shared with lz_block.cu through lz_dense_host.cuh)
// ---------------------------------------------------------------------------------------------
static inline int dense_grid(const lz_ctx *ctx, int64_t n)
{
    int64_t slabs = (n + 31) / 32;
    int64_t want = (slabs + LZ_DENSE_WARPS - 1) / LZ_DENSE_WARPS;
    int64_t cap = (int64_t)ctx->sm_count * 4;
    return (int)(want < 1 ? 1 : (want < cap ? want : cap));
}

// W^T W of a row-major panel, executed only when *run_flag != 0 (G keeps its value otherwise); bw in {8, 16, 32}
int lz_gram_if(lz_ctx *ctx, int64_t n, int bw, const double *W, double *G, const int *run_flag)
{
    const int grid = dense_grid(ctx, n);
    void *w;
    LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)grid * bw * bw, &w));
    lz_prof_begin(ctx, LZ_K_GRAM, 8.0 * (double)n * bw);
    if (bw == 8) k_gram_dmma<8, true, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, W, 0, W, 0, (double *)w, run_flag);
    else if (bw == 16) k_gram_dmma<16, true, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, W, 0, W, 0, (double *)w, run_flag);
    else if (bw == 32) k_gram_dmma<32, true, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, W, 0, W, 0, (double *)w, run_flag);
    else { lz_set_error("lz_gram_if: block width %d", bw); return LZ_ERR_UNSUPPORTED; }
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    k_gram_reduce<<<(bw * bw + 7) / 8, 256, 0, ctx->stream>>>(bw, grid, (const double *)w, bw * bw, G, 0, run_flag);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

template <bool RMX, bool RMY>
static int gram_launch(lz_ctx *ctx, int64_t n, int bw, const double *X, int64_t ldx, const double *Y, int64_t ldy,
                       double *G, int mode, double *gpart, int grid)
{
    lz_prof_begin(ctx, LZ_K_GRAM, 8.0 * (double)n * bw * (X == Y ? 1.0 : 2.0));
    if (bw == 8) k_gram_dmma<8, RMX, RMY><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, X, ldx, Y, ldy, gpart);
    else if (bw == 16) k_gram_dmma<16, RMX, RMY><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, X, ldx, Y, ldy, gpart);
    else if (bw == 32) k_gram_dmma<32, RMX, RMY><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, X, ldx, Y, ldy, gpart);
    else k_gram_simt<RMX, RMY><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, bw, X, ldx, Y, ldy, gpart);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    k_gram_reduce<<<(bw * bw + 7) / 8, 256, 0, ctx->stream>>>(bw, grid, gpart, bw * bw, G, mode);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

enum { SB_S8A = 8192, SB_S8B = 8192 + 64 };      // 8 x 8 block-diagonal copies of 4 x 4 factors (ctx->scalars)

int lz_gram(lz_ctx *ctx, int64_t n, int bw, bool rm, const double *X, int64_t ldx, const double *Y, int64_t ldy,
            double *G, int mode)
{
    const int grid = dense_grid(ctx, n);
    void *w;
    if (rm && bw == 4 && n % 2 == 0 && (uintptr_t)X % 16 == 0 && (uintptr_t)Y % 16 == 0) {      // width 4 on the tensor pipe (pairs of rows)
        LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)grid * 64, &w));
        lz_prof_begin(ctx, LZ_K_GRAM, 8.0 * (double)n * bw * (X == Y ? 1.0 : 2.0));
        k_gram_dmma<8, true, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n / 2, X, 0, Y, 0, (double *)w);
        LZ_LAUNCH_CHECK(ctx);
        lz_prof_end(ctx);
        k_gram_reduce_fold4<<<2, 256, 0, ctx->stream>>>(grid, (const double *)w, G, mode);
        LZ_LAUNCH_CHECK(ctx);
        return LZ_OK;
    }
    LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)grid * bw * bw, &w));
    if (rm) return gram_launch<true, true>(ctx, n, bw, X, ldx, Y, ldy, G, mode, (double *)w, grid);
    return gram_launch<false, false>(ctx, n, bw, X, ldx, Y, ldy, G, mode, (double *)w, grid);
}

template <bool RM>
static int panel_launch(lz_ctx *ctx, int64_t n, int bw, const double *T, int64_t ldt, const double *S, double beta,
                        double alpha, double *R, int64_t ldr, double *G, double *gpart, int grid)
{
    const bool gram = G != nullptr;
    lz_prof_begin(ctx, LZ_K_PANEL, 8.0 * (double)n * bw * (beta != 0.0 ? 3.0 : 2.0));
#define LZ_PANEL_CASE(B)                                                                                            \
    if (gram) k_panel_dmma<B, RM, RM, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, T, ldt, S, beta, alpha, R, ldr, gpart); \
    else k_panel_dmma<B, RM, RM, false><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, T, ldt, S, beta, alpha, R, ldr, gpart)
    if (bw == 8) { LZ_PANEL_CASE(8); }
    else if (bw == 16) { LZ_PANEL_CASE(16); }
    else if (bw == 32) { LZ_PANEL_CASE(32); }
    else {
        k_panel_simt<RM, RM><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, bw, T, ldt, S, beta, alpha, R, ldr);
    }
#undef LZ_PANEL_CASE
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    if (gram) {
        if (bw == 8 || bw == 16 || bw == 32) {
            k_gram_reduce<<<(bw * bw + 7) / 8, 256, 0, ctx->stream>>>(bw, grid, gpart, bw * bw, G, 0);
            LZ_LAUNCH_CHECK(ctx);
        } else {
            return gram_launch<RM, RM>(ctx, n, bw, R, ldr, R, ldr, G, 0, gpart, grid);
        }
    }
    return LZ_OK;
}

int lz_panel(lz_ctx *ctx, int64_t n, int bw, bool rm, const double *T, int64_t ldt, const double *S, double beta,
             double alpha, double *R, int64_t ldr, double *G_opt)
{
    const int grid = dense_grid(ctx, n);
    void *w = nullptr;
    if (rm && bw == 4 && n % 2 == 0 && (uintptr_t)T % 16 == 0 && (uintptr_t)R % 16 == 0) {      // width 4 on the tensor pipe (pairs of rows)
        double *S8 = ctx->scalars + SB_S8A;
        k_blockdiag2<<<1, 64, 0, ctx->stream>>>(S, S8);
        LZ_LAUNCH_CHECK(ctx);
        if (G_opt) LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)grid * 64, &w));
        lz_prof_begin(ctx, LZ_K_PANEL, 8.0 * (double)n * bw * (beta != 0.0 ? 3.0 : 2.0));
        if (G_opt) k_panel_dmma<8, true, true, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n / 2, T, 0, S8, beta, alpha, R, 0, (double *)w);
        else k_panel_dmma<8, true, true, false><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n / 2, T, 0, S8, beta, alpha, R, 0, nullptr);
        LZ_LAUNCH_CHECK(ctx);
        lz_prof_end(ctx);
        if (G_opt) {
            k_gram_reduce_fold4<<<2, 256, 0, ctx->stream>>>(grid, (const double *)w, G_opt, 0);
            LZ_LAUNCH_CHECK(ctx);
        }
        return LZ_OK;
    }
    if (G_opt) LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)grid * bw * bw, &w));
    if (rm) return panel_launch<true>(ctx, n, bw, T, ldt, S, beta, alpha, R, ldr, G_opt, (double *)w, grid);
    return panel_launch<false>(ctx, n, bw, T, ldt, S, beta, alpha, R, ldr, G_opt, (double *)w, grid);
}

// W -= T1 S1 + T2 S2 (row-major), G_opt receives W_new^T W_new
int lz_panel2(lz_ctx *ctx, int64_t n, int bw, const double *T1, const double *S1, const double *T2, const double *S2, double *W, double *G_opt)
{
    const int grid = dense_grid(ctx, n);
    void *w = nullptr;
    if (bw == 4 && n % 2 == 0 && (uintptr_t)T1 % 16 == 0 && (uintptr_t)T2 % 16 == 0 && (uintptr_t)W % 16 == 0) {   // width 4: pairs of rows
        double *S8a = ctx->scalars + SB_S8A, *S8b = ctx->scalars + SB_S8B;
        k_blockdiag2<<<1, 64, 0, ctx->stream>>>(S1, S8a);
        LZ_LAUNCH_CHECK(ctx);
        k_blockdiag2<<<1, 64, 0, ctx->stream>>>(S2, S8b);
        LZ_LAUNCH_CHECK(ctx);
        if (G_opt) LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)grid * 64, &w));
        lz_prof_begin(ctx, LZ_K_PANEL, 8.0 * (double)n * bw * 4.0);
        if (G_opt) k_panel2_dmma<8, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n / 2, T1, S8a, T2, S8b, W, (double *)w);
        else k_panel2_dmma<8, false><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n / 2, T1, S8a, T2, S8b, W, nullptr);
        LZ_LAUNCH_CHECK(ctx);
        lz_prof_end(ctx);
        if (G_opt) {
            k_gram_reduce_fold4<<<2, 256, 0, ctx->stream>>>(grid, (const double *)w, G_opt, 0);
            LZ_LAUNCH_CHECK(ctx);
        }
        return LZ_OK;
    }
    if (!(bw == 8 || bw == 16 || bw == 32)) {
        LZ_TRY(lz_panel(ctx, n, bw, true, T1, 0, S1, 1.0, -1.0, W, 0, nullptr));
        return lz_panel(ctx, n, bw, true, T2, 0, S2, 1.0, -1.0, W, 0, G_opt);
    }
    if (G_opt) LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)grid * bw * bw, &w));
    lz_prof_begin(ctx, LZ_K_PANEL, 8.0 * (double)n * bw * 4.0);
#define LZ_P2(B)                                                                                                        \
    if (G_opt) k_panel2_dmma<B, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, T1, S1, T2, S2, W, (double *)w);   \
    else k_panel2_dmma<B, false><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, T1, S1, T2, S2, W, (double *)w)
    if (bw == 8) { LZ_P2(8); } else if (bw == 16) { LZ_P2(16); } else { LZ_P2(32); }
#undef LZ_P2
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    if (G_opt) {
        k_gram_reduce<<<(bw * bw + 7) / 8, 256, 0, ctx->stream>>>(bw, grid, (const double *)w, bw * bw, G_opt, 0);
        LZ_LAUNCH_CHECK(ctx);
    }
    return LZ_OK;
}

// G1 = X^T Y1, G2 = X^T Y2 (row-major panels, one read of X); then alpha = sym(G1 - G2 Bm)
int lz_gram2(lz_ctx *ctx, int64_t n, int bw, const double *X, const double *Y1, const double *Y2, double *G1, double *G2)
{
    if (!(bw == 8 || bw == 16 || bw == 32)) {
        LZ_TRY(lz_gram(ctx, n, bw, true, X, 0, Y1, 0, G1, 0));
        return lz_gram(ctx, n, bw, true, X, 0, Y2, 0, G2, 0);
    }
    const int grid = dense_grid(ctx, n);
    void *w;
    LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)grid * 2 * bw * bw, &w));
    lz_prof_begin(ctx, LZ_K_GRAM, 8.0 * (double)n * bw * 3.0);
    if (bw == 8) k_gram2_dmma<8, false><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, X, Y1, Y2, (double *)w);
    else if (bw == 16) k_gram2_dmma<16, false><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, X, Y1, Y2, (double *)w);
    else k_gram2_dmma<32, true><<<grid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, X, Y1, Y2, (double *)w);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    k_gram_reduce<<<(bw * bw + 7) / 8, 256, 0, ctx->stream>>>(bw, grid, (const double *)w, 2 * bw * bw, G1, 0);
    LZ_LAUNCH_CHECK(ctx);
    k_gram_reduce<<<(bw * bw + 7) / 8, 256, 0, ctx->stream>>>(bw, grid, (const double *)w + bw * bw, 2 * bw * bw, G2, 0);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

int lz_alpha_from_grams(lz_ctx *ctx, int bw, const double *G1, const double *G2, const double *Bm, double *alpha)
{
    k_alpha_from_grams<<<1, 256, 0, ctx->stream>>>(bw, G1, G2, Bm, alpha);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

// one classical block Gram-Schmidt sweep of W (row-major n x bw) against J stored row-major blocks
template <int BW, int JB>
static int block_cgs_launch(lz_ctx *ctx, int64_t n, int J, const double *V, int64_t pan, double *W, double *C, bool sharded, const int *run_flag)
{
    const int gx = ctx->sm_count * 2, batches = (J + JB - 1) / JB;
    const size_t bb = (size_t)BW * BW;
    void *w;
    LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * ((size_t)batches * gx * JB * bb + (size_t)J * bb), &w));
    double *gpart = (double *)w, *Cf = gpart + (size_t)batches * gx * JB * bb;
    const int ugrid = ctx->sm_count * 4;
    lz_prof_begin(ctx, LZ_K_PROJECT, 8.0 * (double)n * BW * (J + batches));
    if constexpr (BW >= 16) k_block_project_w<BW, JB><<<dim3(gx, batches), LZ_DENSE_THREADS, 0, ctx->stream>>>(n, J, V, pan, W, gpart, run_flag);
    else k_block_project<BW, JB><<<dim3(gx, batches), LZ_DENSE_THREADS, 0, ctx->stream>>>(n, J, V, pan, W, gpart, run_flag);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    if constexpr (BW >= 16) k_block_project_reduce_w<BW, JB><<<J, 1024, 0, ctx->stream>>>(J, gx, gpart, C, run_flag);
    else k_block_project_reduce<JB><<<J, 1024, 0, ctx->stream>>>(BW, J, gx, gpart, C, run_flag);
    LZ_LAUNCH_CHECK(ctx);
    if (sharded) LZ_TRY(lz_comm_allreduce_sum(ctx, C, (size_t)J * bb));      // the coefficients of all ranks' row slabs add up
    if constexpr (BW >= 16) {        // fragment-ordered negative for the update kernel (local permutation)
        k_block_coef_frag<BW><<<J, 256, 0, ctx->stream>>>(J, C, Cf, run_flag);
        LZ_LAUNCH_CHECK(ctx);
    }
    lz_prof_begin(ctx, LZ_K_UPDATE, 8.0 * (double)n * BW * (J + 2));
    if constexpr (BW >= 16) k_block_update_w<BW><<<ugrid < 1 ? 1 : ugrid, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, J, V, pan, Cf, W, run_flag);
    else k_block_update<BW><<<dense_grid(ctx, n), LZ_DENSE_THREADS, 0, ctx->stream>>>(n, J, V, pan, C, W, run_flag);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return LZ_OK;
}

// out = sum_j V_j Yc_j  for J stored row-major blocks and J coefficient blocks Yc (b x b column-major, back to back):
// the update kernel of the block CGS with W = 0 and C = -Yc (thick restart: one output block of V Y)
template <int BW>
static int block_combine_launch(lz_ctx *ctx, int64_t n, int J, const double *V, int64_t pan, double *C /* holds -Yc */, double *out)
{
    const size_t bb = (size_t)BW * BW;
    LZ_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)n * BW, ctx->stream));
    lz_prof_begin(ctx, LZ_K_UPDATE, 8.0 * (double)n * BW * (J + 2));
    if constexpr (BW >= 16) {
        void *w;
        LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)J * bb, &w));
        k_block_coef_frag<BW><<<J, 256, 0, ctx->stream>>>(J, C, (double *)w);
        LZ_LAUNCH_CHECK(ctx);
        k_block_update_w<BW><<<ctx->sm_count * 4, LZ_DENSE_THREADS, 0, ctx->stream>>>(n, J, V, pan, (const double *)w, out);
    } else {
        k_block_update<BW><<<dense_grid(ctx, n), LZ_DENSE_THREADS, 0, ctx->stream>>>(n, J, V, pan, C, out);
    }
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return LZ_OK;
}

int lz_block_combine(lz_ctx *ctx, int64_t n, int bw, int J, const double *V, int64_t pan, double *negY, double *out)
{
    if (bw == 8) return block_combine_launch<8>(ctx, n, J, V, pan, negY, out);
    if (bw == 16) return block_combine_launch<16>(ctx, n, J, V, pan, negY, out);
    if (bw == 32) return block_combine_launch<32>(ctx, n, J, V, pan, negY, out);
    lz_set_error("lz_block_combine: block width %d (use 8, 16 or 32)", bw);
    return LZ_ERR_UNSUPPORTED;
}

int lz_block_dgks_test(lz_ctx *ctx, int bw, const double *G_before, const double *G_after, int *flag)
{
    k_block_dgks_test<<<1, 32, 0, ctx->stream>>>(bw, G_before, G_after, flag);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

int lz_block_cgs(lz_ctx *ctx, int64_t n, int bw, int J, const double *V, int64_t pan, double *W, double *C, bool sharded, const int *run_flag)
{
    if (bw == 8) return block_cgs_launch<8, 8>(ctx, n, J, V, pan, W, C, sharded, run_flag);
    if (bw == 16) return block_cgs_launch<16, 4>(ctx, n, J, V, pan, W, C, sharded, run_flag);
    if (bw == 32) return block_cgs_launch<32, 2>(ctx, n, J, V, pan, W, C, sharded, run_flag);
    LZ_CHECK(run_flag == nullptr, LZ_ERR_UNSUPPORTED, "conditional block sweeps need a block width of 8, 16 or 32");
    // generic widths: block-by-block products (SIMT)
    const size_t bb = (size_t)bw * bw;
    for (int j = 0; j < J; ++j) LZ_TRY(lz_gram(ctx, n, bw, true, V + pan * j, 0, W, 0, C + bb * j, 0));
    if (sharded) LZ_TRY(lz_comm_allreduce_sum(ctx, C, (size_t)J * bb));
    for (int j = 0; j < J; ++j) LZ_TRY(lz_panel(ctx, n, bw, true, V + pan * j, 0, C + bb * j, 1.0, -1.0, W, 0, nullptr));
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// b x b symmetric eigen-decomposition + matrix square root, one CTA.
// Parallel cyclic Jacobi (round-robin pairing: b/2 disjoint rotations per round, b-1 rounds per
// sweep), then S = V sqrt|L| V^T, Sinv = V |L|^{-1/2} V^T  -- the semantics of
// cusolverDnDsyevjBatched + custom_mult2 (utils/lib_utils.hpp:650-745) and of
// sqrtm::My_sqrtm_cusolver (kernels/my_sqrtm_cusolver.hpp:174-361).  Reads the lower triangle.
// ---------------------------------------------------------------------------------------------
#define SQ_LD 33
__global__ void __launch_bounds__(1024) k_sqrtm(int b, double *__restrict__ S, double *__restrict__ Sinv, int *flags, int jidx)
{
    __shared__ double A[32 * SQ_LD], V[32 * SQ_LD], cs[32], sn[32], lam[32];
    __shared__ int pp[32], qq[32];
    __shared__ double off2, dia2;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int r = tid % b, c = tid / b;           // valid when tid < b*b
    const bool act = tid < b * b;
    if (act) {
        A[r + c * SQ_LD] = (r >= c) ? S[r + c * b] : S[c + r * b];
        V[r + c * SQ_LD] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();
    const int m = (b + 1) & ~1;                   // players in the round-robin tournament (pad odd b)
    bool last = false;                            // Jacobi converges quadratically: one sweep after off/diag < 1e-9 is at rounding level
    for (int sweep = 0; sweep < 30 && b > 1 && !last; ++sweep) {
        if (tid == 0) { off2 = 0.0; dia2 = 0.0; }
        __syncthreads();
        {
            const double v = act ? A[r + c * SQ_LD] : 0.0;
            double o = (act && r != c) ? v * v : 0.0, d = (act && r == c) ? v * v : 0.0;
            o = lz_warp_sum(o); d = lz_warp_sum(d);
            if ((tid & 31) == 0) { atomicAdd(&off2, o); atomicAdd(&dia2, d); }
        }
        __syncthreads();
        if (off2 <= 1e-30 * dia2) break;
        last = off2 <= 1e-18 * dia2;
        for (int round = 0; round < m - 1; ++round) {
            // pairing of round `round`: player 0 fixed, the others rotate
            if (tid < m / 2) {
                int a0 = (tid == 0) ? 0 : 1 + (tid - 1 + round) % (m - 1);
                int a1 = 1 + (m - 1 - tid - 1 + round) % (m - 1);
                int p = min(a0, a1), q = max(a0, a1);
                double cc = 1.0, ss = 0.0;
                if (q < b) {
                    const double apq = A[p + q * SQ_LD];
                    if (apq != 0.0) {
                        const double tau = (A[q + q * SQ_LD] - A[p + p * SQ_LD]) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        cc = 1.0 / sqrt(1.0 + t * t);
                        ss = t * cc;
                    }
                } else { p = q = -1; }
                pp[tid] = p; qq[tid] = q; cs[tid] = cc; sn[tid] = ss;
            }
            __syncthreads();
            // columns: A <- A J, V <- V J   (thread (pair, k) handles row k of the two columns)
            for (int e = tid; e < (m / 2) * b; e += nthr) {
                const int pr = e / b, k = e % b, p = pp[pr], q = qq[pr];
                if (p >= 0) {
                    const double cc = cs[pr], ss = sn[pr];
                    const double akp = A[k + p * SQ_LD], akq = A[k + q * SQ_LD];
                    A[k + p * SQ_LD] = cc * akp - ss * akq;
                    A[k + q * SQ_LD] = ss * akp + cc * akq;
                    const double vkp = V[k + p * SQ_LD], vkq = V[k + q * SQ_LD];
                    V[k + p * SQ_LD] = cc * vkp - ss * vkq;
                    V[k + q * SQ_LD] = ss * vkp + cc * vkq;
                }
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int e = tid; e < (m / 2) * b; e += nthr) {
                const int pr = e / b, k = e % b, p = pp[pr], q = qq[pr];
                if (p >= 0) {
                    const double cc = cs[pr], ss = sn[pr];
                    const double apk = A[p + k * SQ_LD], aqk = A[q + k * SQ_LD];
                    A[p + k * SQ_LD] = cc * apk - ss * aqk;
                    A[q + k * SQ_LD] = ss * apk + cc * aqk;
                }
            }
            __syncthreads();
        }
    }
    if (tid < b) lam[tid] = fabs(A[tid + tid * SQ_LD]);
    __syncthreads();
    if (tid < b) {
        // singular (to working precision) or non-finite block: remember the FIRST failing block index.
        // W^T W of a rank-deficient W has eigenvalues at rounding level relative to the largest one.
        double lmax = 0.0;
        for (int i = 0; i < b; ++i) lmax = fmax(lmax, lam[i]);
        if (!(lam[tid] > (double)b * 2.220446049250313e-16 * lmax) || !isfinite(lam[tid])) atomicMin(flags, jidx);
    }
    __syncthreads();
    if (act) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = 0; i < b; ++i) {
            const double rt = sqrt(lam[i]);
            s1 += V[r + i * SQ_LD] * rt * V[c + i * SQ_LD];
            s2 += V[r + i * SQ_LD] * 1.0 / rt * V[c + i * SQ_LD];
        }
        S[r + c * b] = s1;
        Sinv[r + c * b] = s2;
    }
}

int lz_sqrtm_launch(lz_ctx *ctx, int b, double *S, double *Sinv, int *flag, int jidx)
{
    lz_prof_begin(ctx, LZ_K_SMALL, 0.0);
    int threads = ((b * b + 31) / 32) * 32;
    k_sqrtm<<<1, threads, 0, ctx->stream>>>(b, S, Sinv, flag, jidx);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return LZ_OK;
}

// q[off + c] = Q[lc, c]
template <bool RM>
__global__ void k_copy_row(int64_t lc, int b, const double *__restrict__ Q, int64_t ld, double *__restrict__ q, int64_t off)
{
    const LzLay<RM> l{ld, b};
    if (threadIdx.x < b) q[off + threadIdx.x] = Q[l.at(lc, threadIdx.x)];
}

int lz_copy_row_launch(lz_ctx *ctx, int64_t lc, int b, bool rm, const double *Q, int64_t ld, double *q, int64_t off)
{
    if (rm) k_copy_row<true><<<1, 32, 0, ctx->stream>>>(lc, b, Q, ld, q, off);
    else k_copy_row<false><<<1, 32, 0, ctx->stream>>>(lc, b, Q, ld, q, off);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

// dense block-tridiagonal T from alpha[0..m), beta[1..m)   (objects/tridiagonal_matrix.hpp:13-54,90-127)
__global__ void k_assemble_T(int m, int b, const double *__restrict__ alpha, const double *__restrict__ beta, double *__restrict__ T)
{
    const int N = m * b;
    const int64_t total = (int64_t)m * b * b;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int blk = (int)(e / (b * b)), i = (int)(e % (b * b)), r = i % b, c = i / b;
        T[(blk * b + r) + (int64_t)(blk * b + c) * N] = alpha[e];
        if (blk >= 1) {
            const double v = beta[e];
            T[((blk - 1) * b + r) + (int64_t)(blk * b + c) * N] = v;
            T[(blk * b + c) + (int64_t)((blk - 1) * b + r) * N] = v;
        }
    }
}

extern "C" {

int lz_mm_tt(lz_ctx *ctx, int64_t n, int b, const double *T, int64_t ld, double *R)
{
    LZ_CHECK(ctx && T && R && n > 0 && b >= 1 && b <= 32 && ld >= n, LZ_ERR_INVALID, "lz_mm_tt: bad arguments (b must be 1..32)");
    return lz_gram(ctx, n, b, false, T, ld, T, ld, R, 0);
}

int lz_mm_tt2(lz_ctx *ctx, int64_t n, int b, const double *T1, int64_t ld1, const double *T2, int64_t ld2, double *R)
{
    LZ_CHECK(ctx && T1 && T2 && R && n > 0 && b >= 1 && b <= 32 && ld1 >= n && ld2 >= n, LZ_ERR_INVALID, "lz_mm_tt2: bad arguments");
    return lz_gram(ctx, n, b, false, T1, ld1, T2, ld2, R, 1);
}

int lz_mm_ts(lz_ctx *ctx, int64_t n, int b, double beta, double alpha, const double *T, int64_t ldt, const double *S,
             double *R, int64_t ldr)
{
    LZ_CHECK(ctx && T && S && R && n > 0 && b >= 1 && b <= 32 && ldt >= n && ldr >= n, LZ_ERR_INVALID, "lz_mm_ts: bad arguments");
    return lz_panel(ctx, n, b, false, T, ldt, S, beta, alpha, R, ldr, nullptr);
}

int lz_sqrtm(lz_ctx *ctx, int b, double *S, double *Sinv)
{
    LZ_CHECK(ctx && S && Sinv && b >= 1 && b <= 32, LZ_ERR_INVALID, "lz_sqrtm: bad arguments (b must be 1..32)");
    return lz_sqrtm_launch(ctx, b, S, Sinv, ctx->flags + 2, 0);
}

int lz_copy_row(lz_ctx *ctx, int64_t lc, int b, const double *Q, int64_t ld, double *q, int64_t off)
{
    LZ_CHECK(ctx && Q && q && b >= 1 && b <= 32 && lc >= 0 && lc < ld, LZ_ERR_INVALID, "lz_copy_row: bad arguments");
    return lz_copy_row_launch(ctx, lc, b, false, Q, ld, q, off);
}

int lz_assemble_T(lz_ctx *ctx, int m, int b, const double *alpha, const double *beta, double *T)
{
    LZ_CHECK(ctx && alpha && beta && T && m >= 1 && b >= 1, LZ_ERR_INVALID, "lz_assemble_T: bad arguments");
    const size_t N = (size_t)m * b;
    LZ_CUDA(cudaMemsetAsync(T, 0, sizeof(double) * N * N, ctx->stream));
    k_assemble_T<<<64, 256, 0, ctx->stream>>>(m, b, alpha, beta, T);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

}  // extern "C"
