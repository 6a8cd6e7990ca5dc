// lz_dense_host.cuh -- host-side launchers of the dense block kernels shared by lz_dense.cu and lz_block.cu
#pragma once
#include "lz_dense.cuh"

// G = X^T Y (mode 0) or 0.5 (X^T Y) + 0.5 (X^T Y)^T (mode 1); rm: panels row-major (ld ignored) or column-major
int lz_gram(lz_ctx *ctx, int64_t n, int bw, bool rm, const double *X, int64_t ldx, const double *Y, int64_t ldy,
            double *G, int mode);
// R = beta R + alpha T S ; G_opt (device bw*bw) additionally receives R_new^T R_new
int lz_panel(lz_ctx *ctx, int64_t n, int bw, bool rm, const double *T, int64_t ldt, const double *S, double beta,
             double alpha, double *R, int64_t ldr, double *G_opt);
int lz_sqrtm_launch(lz_ctx *ctx, int b, double *S, double *Sinv, int *flag, int jidx = 0);
int lz_copy_row_launch(lz_ctx *ctx, int64_t lc, int b, bool rm, const double *Q, int64_t ld, double *q, int64_t off);
// one classical block Gram-Schmidt sweep: C_j = V_j^T W (j < J), W -= sum_j V_j C_j; V_j row-major, `pan` apart
// run_flag (device int, optional): the sweep's kernels return at once when it is 0 (conditional second sweep)
int lz_block_cgs(lz_ctx *ctx, int64_t n, int bw, int J, const double *V, int64_t pan, double *W, double *C, bool sharded = false,
                 const int *run_flag = nullptr);
// flag <- 1 iff some column of W lost more than half of its squared norm (diagonals of W^T W before / after a sweep)
int lz_block_dgks_test(lz_ctx *ctx, int bw, const double *G_before, const double *G_after, int *flag);
// W -= T1 S1 + T2 S2 (row-major panels), G_opt (device bw*bw) receives W_new^T W_new
int lz_panel2(lz_ctx *ctx, int64_t n, int bw, const double *T1, const double *S1, const double *T2, const double *S2, double *W, double *G_opt);
// G1 = X^T Y1 and G2 = X^T Y2 from one read of X (row-major); alpha = sym(G1 - G2 Bm)
int lz_gram2(lz_ctx *ctx, int64_t n, int bw, const double *X, const double *Y1, const double *Y2, double *G1, double *G2);
int lz_alpha_from_grams(lz_ctx *ctx, int bw, const double *G1, const double *G2, const double *Bm, double *alpha);
// out (row-major n x bw) = sum_j V_j Yc_j ; negY holds the J coefficient blocks negated (b x b column-major each)
int lz_block_combine(lz_ctx *ctx, int64_t n, int bw, int J, const double *V, int64_t pan, double *negY, double *out);
// G = W^T W only when *run_flag != 0 (row-major W, bw in {8,16,32})
int lz_gram_if(lz_ctx *ctx, int64_t n, int bw, const double *W, double *G, const int *run_flag);
