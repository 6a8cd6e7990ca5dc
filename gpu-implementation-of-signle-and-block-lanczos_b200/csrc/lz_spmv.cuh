// lz_spmv.cuh -- fp64 SpMV kernels for sm_100a with the Lanczos "pass A" fused into the epilogue.
//
// Replaces ell::SpMV (kernels/spmv_spmm.hpp:105-135), lm::spmv_basic (kernels/ell_kernels.hpp:13-34)
// and, in fused mode, the spmv + Vector::add + Vector::dot sequence of
// methods/vector_lanczos.hpp:51-57 (one HBM pass instead of three).
//
// CSR kernel ("stream" scheduling): one CTA per row-aligned chunk of ~LZ_SPMV_TILE non-zeros.
//   phase 1: the CTA walks its contiguous slice of vals/colidx with 128-bit/64-bit coalesced loads,
//            gathers x and parks val*x products in shared memory (conflict-free 16-byte stores);
//   phase 2: one thread per row adds its products left to right (the reference Host order,
//            objects/ell_matrix.hpp:246-251) and runs the epilogue.
//   chunks that contain a long row (more products than shared memory holds) fall back to
//   warp-per-row / CTA-per-row walks of global memory.
// ELL4 kernel: one thread per row, one 256-bit load of the four values and one 128-bit load of the
// four column indices (the reference's row-interleaved width-4 layout).
#pragma once
#include "lz_common.cuh"

enum { LZ_EPI_PLAIN = 0, LZ_EPI_LANCZOS = 1 };

// Arguments of the fused Lanczos epilogue (pass A of step j):
//   q_j[i]  = x_own[i] * invb[j]                 (lazy normalisation: the same product the reference
//                                                 materialises with mult_scalar, vector_lanczos.hpp:48)
//   w[i]    = sum_k A[i,k] q_j[col_k]  -  beta[j] * (u_prev[i] * invb[j-1])       (:51,:54)
//   alpha_j = sum_i w[i] q_j[i]                                                   (:57)
struct LzPassA {
    const double *x_own;    // x restricted to the locally owned rows (x + halo_lo for shards)
    const double *u_prev;   // unnormalised q_{j-1}
    const double *invb;     // invb[j] = 1 / beta_j
    const double *beta;     // beta[j]
    double *alpha_out;      // &alpha[j] (written by the last CTA) or NULL when the caller all-reduces
    double *alpha_partial;  // where the last CTA leaves the local sum (always)
    double *vcol;           // basis column j to fill with q_j, or NULL
    double *qout;           // &q[j]: receives q_j[lc] (copy_vector_element, copy_functions.hpp:116-133)
    int64_t lc;
    int j;
    int first;              // step 0: no q_{-1} term
    double *partials;
    unsigned int *ticket;
};

template <int MODE>
struct LzRowEpi {
    double sx, sprev, beta;
    const LzPassA &a;
    __device__ __forceinline__ LzRowEpi(const LzPassA &args) : a(args)
    {
        sx = 1.0; sprev = 0.0; beta = 0.0;
        if (MODE == LZ_EPI_LANCZOS) {
            sx = a.invb[a.j];
            if (!a.first) { sprev = a.invb[a.j - 1]; beta = a.beta[a.j]; }
        }
    }
    // scale applied to every gathered x value
    __device__ __forceinline__ double xs(double xv) const { return MODE == LZ_EPI_LANCZOS ? __dmul_rn(xv, sx) : xv; }
    // finish row i whose raw sum is t; returns the row's contribution to alpha
    __device__ __forceinline__ double finish(int64_t i, double t, double *__restrict__ y) const
    {
        if (MODE == LZ_EPI_PLAIN) { y[i] = t; return 0.0; }
        const double qi = __dmul_rn(a.x_own[i], sx);
        double w = t;
        if (!a.first) w = __dadd_rn(t, __dmul_rn(-beta, __dmul_rn(a.u_prev[i], sprev)));
        y[i] = w;
        if (a.vcol) a.vcol[i] = qi;
        if (i == a.lc && a.qout) *a.qout = qi;
        return __dmul_rn(w, qi);
    }
};

template <int MODE>
__device__ __forceinline__ void lz_spmv_finalize(double acc, const LzPassA &a, double *red)
{
    if (MODE != LZ_EPI_LANCZOS) return;
    acc = lz_block_sum<LZ_SPMV_THREADS>(acc, red);
    double total;
    if (lz_grid_sum<LZ_SPMV_THREADS, 1>(&acc, a.partials, a.ticket, red, &total)) {
        if (threadIdx.x == 0) {
            *a.alpha_partial = total;
            if (a.alpha_out) *a.alpha_out = total;
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(LZ_SPMV_THREADS)
k_csr_spmv(const int32_t *__restrict__ chunk_row, const int32_t *__restrict__ rowptr,
           const int32_t *__restrict__ colidx, const double *__restrict__ vals,
           const double *__restrict__ x, double *__restrict__ y, const LzPassA args)
{
    __shared__ __align__(16) double prod[LZ_SPMV_CAP + 2];
    __shared__ double red[32];
    const int tid = threadIdx.x;
    const int r0 = chunk_row[blockIdx.x], r1 = chunk_row[blockIdx.x + 1];
    const LzRowEpi<MODE> epi(args);
    double acc = 0.0;
    if (r1 > r0) {
        const int p0 = rowptr[r0], p1 = rowptr[r1];
        const int a0 = p0 & ~1;                       // products live at prod[p - a0]
        if (p1 - a0 <= LZ_SPMV_CAP) {
            // ---- phase 1: coalesced product generation --------------------------------------
            const int pe = p1 & ~1;                   // pairs cover [a0, pe)
            for (int p = a0 + 2 * tid; p < pe; p += 2 * LZ_SPMV_THREADS) {
                const double2 v = *reinterpret_cast<const double2 *>(vals + p);
                const int2 c = *reinterpret_cast<const int2 *>(colidx + p);
                double2 pr;
                // the first pair may start one entry before p0 (belongs to the previous chunk):
                // its column is valid memory, its product is never read
                pr.x = __dmul_rn(v.x, epi.xs(__ldg(x + c.x)));
                pr.y = __dmul_rn(v.y, epi.xs(__ldg(x + c.y)));
                *reinterpret_cast<double2 *>(prod + (p - a0)) = pr;
            }
            if (tid == 0 && pe < p1) prod[pe - a0] = __dmul_rn(vals[pe], epi.xs(__ldg(x + colidx[pe])));
            __syncthreads();
            // ---- phase 2: one thread per row, left-to-right sum -------------------------------
            for (int r = r0 + tid; r < r1; r += LZ_SPMV_THREADS) {
                const int s = rowptr[r] - a0, e = rowptr[r + 1] - a0;
                double t = 0.0;
                for (int k = s; k < e; ++k) t = __dadd_rn(t, prod[k]);
                acc += epi.finish(r, t, y);
            }
        } else {
            // ---- long-row fallback: warp per row, then CTA per very long row -------------------
            const int lane = tid & 31, warp = tid >> 5;
            for (int r = r0 + warp; r < r1; r += LZ_SPMV_THREADS / 32) {
                const int s = rowptr[r], e = rowptr[r + 1];
                if (e - s > 4096) continue;
                double t = 0.0;
                for (int k = s + lane; k < e; k += 32) t += vals[k] * epi.xs(__ldg(x + colidx[k]));
                t = lz_warp_sum(t);
                if (lane == 0) acc += epi.finish(r, t, y);
            }
            for (int r = r0; r < r1; ++r) {
                const int s = rowptr[r], e = rowptr[r + 1];
                if (e - s <= 4096) continue;
                double t = 0.0;
                for (int k = s + tid; k < e; k += LZ_SPMV_THREADS) t += vals[k] * epi.xs(__ldg(x + colidx[k]));
                t = lz_block_sum<LZ_SPMV_THREADS>(t, red);
                if (tid == 0) acc += epi.finish(r, t, y);
            }
        }
    }
    lz_spmv_finalize<MODE>(acc, args, red);
}

// width-4 row-interleaved ELL (data[4r+k], idx[4r+k]); zero padding entries carry idx 0
template <int MODE>
__global__ void __launch_bounds__(LZ_SPMV_THREADS)
k_ell4_spmv(int64_t n_rows, const double *__restrict__ data, const uint32_t *__restrict__ idx,
            const double *__restrict__ x, double *__restrict__ y, const LzPassA args)
{
    __shared__ double red[32];
    const LzRowEpi<MODE> epi(args);
    double acc = 0.0;
    const int64_t r = (int64_t)blockIdx.x * LZ_SPMV_THREADS + threadIdx.x;
    if (r < n_rows) {
        double v0, v1, v2, v3;
        lz_ld256(data + 4 * r, v0, v1, v2, v3);
        const uint4 c = *reinterpret_cast<const uint4 *>(idx + 4 * r);
        double t = __dmul_rn(v0, epi.xs(__ldg(x + c.x)));
        t = __dadd_rn(t, __dmul_rn(v1, epi.xs(__ldg(x + c.y))));
        t = __dadd_rn(t, __dmul_rn(v2, epi.xs(__ldg(x + c.z))));
        t = __dadd_rn(t, __dmul_rn(v3, epi.xs(__ldg(x + c.w))));
        acc = epi.finish(r, t, y);
    }
    lz_spmv_finalize<MODE>(acc, args, red);
}

// host-side launcher shared by lz_spmv() and the drivers
template <int MODE>
static inline int lz_launch_spmv(lz_ctx *ctx, const lz_matrix *A, const double *x, double *y, const LzPassA &args)
{
    if (A->format == LZ_FMT_ELL4) {
        const unsigned grid = (unsigned)((A->n_rows + LZ_SPMV_THREADS - 1) / LZ_SPMV_THREADS);
        k_ell4_spmv<MODE><<<grid, LZ_SPMV_THREADS, 0, ctx->stream>>>(A->n_rows, A->ell_data, A->ell_idx, x, y, args);
    } else {
        k_csr_spmv<MODE><<<A->n_chunks, LZ_SPMV_THREADS, 0, ctx->stream>>>(A->chunk_row, A->rowptr, A->colidx, A->vals, x, y, args);
    }
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}
