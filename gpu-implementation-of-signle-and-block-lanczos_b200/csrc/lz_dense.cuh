// lz_dense.cuh -- tall-skinny dense contractions of block Lanczos on the fp64 tensor pipe.
//
// Replaces the reference's mm_tt / mm_tt2 / mm_ts SIMT kernels (kernels/mm_tt.hpp, mm_tt2.hpp,
// mm_ts.hpp -- float only, and mm_tt/mm_tt2 multiply .z*.x), their cuBLAS
// stand-ins (utils/lib_utils.hpp:28-202) and the non-compiling wmma sketches under
// tensor_core_unfinished_work/.  tcgen05.mma has no fp64 kind (ptxas rejects .kind::f64), so the
// native fp64 tensor instruction on sm_100a is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4); operands go
// global -> registers directly in fragment order (every 32-byte sector is used once), accumulators
// stay in registers, and partial results are combined in a fixed order (deterministic).
//
// Panels are n x BW.  Layout is a template parameter: row-major (BW contiguous, the drivers'
// internal layout) or column-major with leading dimension ld (the reference's Dense_matrix layout,
// objects/dense_matrix.hpp:9).  BW in {8,16,32} runs on DMMA; any 1 <= BW <= 32 has a SIMT path.
#pragma once
#include "lz_common.cuh"

template <bool RM>
struct LzLay {
    int64_t ld;   // column-major leading dimension (ignored for row-major)
    int bw;
    __device__ __forceinline__ int64_t at(int64_t i, int c) const { return RM ? i * bw + c : (int64_t)c * ld + i; }
};

__device__ __forceinline__ void lz_dmma(double &d0, double &d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

#define LZ_DENSE_THREADS 256
#define LZ_DENSE_WARPS (LZ_DENSE_THREADS / 32)

// ---------------------------------------------------------------------------------------------
// G_partial[cta] = X^T Y over the CTA's rows  (BW x BW, column-major G[p + q*BW]).
// Each warp walks 32-row slabs; per 4-row group it loads BW/8 A fragments (X^T) and BW/8 B
// fragments (Y) -- one double per thread each -- and issues (BW/8)^2 DMMAs.
// ---------------------------------------------------------------------------------------------
template <int BW, bool RMX, bool RMY>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_gram_dmma(int64_t n, const double *__restrict__ X, int64_t ldx, const double *__restrict__ Y, int64_t ldy,
            double *__restrict__ gpart, const int *__restrict__ run_flag = nullptr)
{
    if (run_flag && *run_flag == 0) return;
    constexpr int T = BW / 8;
    const LzLay<RMX> lx{ldx, BW};
    const LzLay<RMY> ly{ldy, BW};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    double acc[T][T][2];
#pragma unroll
    for (int a = 0; a < T; ++a)
#pragma unroll
        for (int b = 0; b < T; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    const int64_t n_slabs = (n + 31) / 32;
    const int64_t wglobal = (int64_t)blockIdx.x * LZ_DENSE_WARPS + warp, wtotal = (int64_t)gridDim.x * LZ_DENSE_WARPS;
    for (int64_t slab = wglobal; slab < n_slabs; slab += wtotal) {
        const int64_t base = slab * 32;
        if (base + 32 <= n) {
#pragma unroll 2
            for (int g = 0; g < 8; ++g) {
                const int64_t i = base + g * 4 + kk;
                double xa[T], yb[T];
#pragma unroll
                for (int t = 0; t < T; ++t) { xa[t] = __ldg(X + lx.at(i, t * 8 + mm)); yb[t] = __ldg(Y + ly.at(i, t * 8 + mm)); }
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) lz_dmma(acc[a][b][0], acc[a][b][1], xa[a], yb[b]);
            }
        } else {
            for (int g = 0; g < 8; ++g) {
                const int64_t i = base + g * 4 + kk;
                double xa[T], yb[T];
#pragma unroll
                for (int t = 0; t < T; ++t) {
                    xa[t] = i < n ? X[lx.at(i, t * 8 + mm)] : 0.0;
                    yb[t] = i < n ? Y[ly.at(i, t * 8 + mm)] : 0.0;
                }
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) lz_dmma(acc[a][b][0], acc[a][b][1], xa[a], yb[b]);
            }
        }
    }
    // combine the warps of the CTA in a fixed order (warp 0 first), then publish the CTA partial
    __shared__ double sm[BW * BW];
    for (int w = 0; w < LZ_DENSE_WARPS; ++w) {
        if (warp == w) {
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) {
                    const int p = a * 8 + mm, q = b * 8 + 2 * kk;
                    if (w == 0) { sm[p + q * BW] = acc[a][b][0]; sm[p + (q + 1) * BW] = acc[a][b][1]; }
                    else { sm[p + q * BW] += acc[a][b][0]; sm[p + (q + 1) * BW] += acc[a][b][1]; }
                }
        }
        __syncthreads();
    }
    for (int e = threadIdx.x; e < BW * BW; e += LZ_DENSE_THREADS) gpart[(size_t)blockIdx.x * BW * BW + e] = sm[e];
}

// generic SIMT Gram for any 1 <= bw <= 32: thread e owns output (p,q) = (e % bw, e / bw);
// rows are staged through shared memory 32 at a time.
template <bool RMX, bool RMY>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_gram_simt(int64_t n, int bw, const double *__restrict__ X, int64_t ldx, const double *__restrict__ Y, int64_t ldy,
            double *__restrict__ gpart)
{
    const LzLay<RMX> lx{ldx, bw};
    const LzLay<RMY> ly{ldy, bw};
    __shared__ double xs[32][33], ys[32][33];
    double acc[4] = {0, 0, 0, 0};                       // outputs e = tid + k*256 (bw*bw <= 1024)
    const int64_t n_slabs = (n + 31) / 32;
    for (int64_t slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
        const int64_t base = slab * 32;
        __syncthreads();
        for (int e = threadIdx.x; e < 32 * bw; e += LZ_DENSE_THREADS) {
            const int r = RMX ? e / bw : e % 32, c = RMX ? e % bw : e / 32;
            const int64_t i = base + r;
            xs[r][c] = i < n ? X[lx.at(i, c)] : 0.0;
            ys[r][c] = i < n ? Y[ly.at(i, c)] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int e = threadIdx.x + k * LZ_DENSE_THREADS;
            if (e < bw * bw) {
                const int p = e % bw, q = e / bw;
                double s = acc[k];
                for (int r = 0; r < 32; ++r) s = fma(xs[r][p], ys[r][q], s);
                acc[k] = s;
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int e = threadIdx.x + k * LZ_DENSE_THREADS;
        if (e < bw * bw) gpart[(size_t)blockIdx.x * bw * bw + e] = acc[k];
    }
}

// G = sum over CTA partials (fixed order).  mode 0: G as is; 1: 0.5*G + 0.5*G^T (mm_tt2, lib_utils.hpp:165-202).
// One warp per output entry: the lanes stride over the partials, then a fixed shuffle tree -- the serial
// per-thread walk over ~600 partials this replaces cost more than the Gram kernel's own tail.
// gstride: doubles between consecutive partials (bw*bw, or 2*bw*bw for the two-Gram kernel).
static __global__ void __launch_bounds__(256)
k_gram_reduce(int bw, int n_parts, const double *__restrict__ gpart, int gstride, double *__restrict__ G, int mode,
              const int *__restrict__ run_flag = nullptr)
{
    if (run_flag && *run_flag == 0) return;
    const int lane = threadIdx.x & 31, e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= bw * bw) return;
    const int p = e % bw, q = e / bw, et = q + p * bw;
    double s = 0.0, st = 0.0;
    for (int i = lane; i < n_parts; i += 32) {
        s += gpart[(size_t)i * gstride + e];
        if (mode == 1) st += gpart[(size_t)i * gstride + et];
    }
    s = lz_warp_sum(s);
    if (mode == 1) st = lz_warp_sum(st);
    if (lane == 0) G[e] = mode == 0 ? s : 0.5 * s + 0.5 * st;
}

// Width-4 panels on the tensor pipe.  mma.m8n8k4 wants 8 columns, so a row-major n x 4 panel (n even) is read as an
// (n/2) x 8 panel whose row i is [row 2i | row 2i+1]:  X8^T Y8 = [[Xe^T Ye, Xe^T Yo], [Xo^T Ye, Xo^T Yo]] and the 4 x 4
// Gram matrix is the sum of the two diagonal blocks;  T8 blockdiag(S, S) = [Te S | To S] is the panel product itself.
static __global__ void k_blockdiag2(const double *__restrict__ S, double *__restrict__ S8)
{
    const int e = threadIdx.x;
    if (e < 64) { const int r = e % 8, c = e / 8; S8[e] = (r / 4 == c / 4) ? S[(r % 4) + (c % 4) * 4] : 0.0; }
}
static __global__ void __launch_bounds__(256)
k_gram_reduce_fold4(int n_parts, const double *__restrict__ gpart /* [parts][64] */, double *__restrict__ G, int mode)
{
    const int lane = threadIdx.x & 31, e = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (e >= 16) return;
    const int p = e % 4, q = e / 4;
    double s = 0.0, st = 0.0;
    for (int i = lane; i < n_parts; i += 32) {
        const double *g = gpart + (size_t)i * 64;
        s += g[p + q * 8] + g[(p + 4) + (q + 4) * 8];
        if (mode == 1) st += g[q + p * 8] + g[(q + 4) + (p + 4) * 8];
    }
    s = lz_warp_sum(s);
    if (mode == 1) st = lz_warp_sum(st);
    if (lane == 0) G[e] = mode == 0 ? s : 0.5 * s + 0.5 * st;
}

// two Gram matrices from one read of X:  G1_partial = X^T Y1,  G2_partial = X^T Y2   (row-major panels).
// Used for the reference order of the block recurrence when the SpMM cannot subtract Q_{j-1} beta_j itself:
// alpha_j = sym(Q_j^T (A Q_j - Q_{j-1} beta_j)) = sym(G1 - G2 beta_j) with G1 = Q_j^T (A Q_j), G2 = Q_j^T Q_{j-1}.
template <int BW, bool SPLIT>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_gram2_dmma(int64_t n, const double *__restrict__ X, const double *__restrict__ Y1, const double *__restrict__ Y2,
             double *__restrict__ gpart /* [cta][2][BW*BW] */)
{
    // SPLIT (BW = 32, 32 accumulator doubles per product): warps 0-3 form X^T Y1 and warps 4-7 X^T Y2 over the SAME
    // slabs (X is fetched by both halves, the second time from L1); otherwise every warp forms both products.
    constexpr int T = BW / 8, NP = SPLIT ? 1 : 2, WPP = SPLIT ? LZ_DENSE_WARPS / 2 : LZ_DENSE_WARPS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    const int half = SPLIT ? warp / WPP : 0, wsub = SPLIT ? warp % WPP : warp;
    double acc[NP][T][T][2];
#pragma unroll
    for (int h = 0; h < NP; ++h)
#pragma unroll
        for (int a = 0; a < T; ++a)
#pragma unroll
            for (int b = 0; b < T; ++b) acc[h][a][b][0] = acc[h][a][b][1] = 0.0;
    const double *Ya = SPLIT ? (half ? Y2 : Y1) : Y1;
    const int64_t n_slabs = (n + 31) / 32;
    const int64_t wglobal = (int64_t)blockIdx.x * WPP + wsub, wtotal = (int64_t)gridDim.x * WPP;
    for (int64_t slab = wglobal; slab < n_slabs; slab += wtotal) {
#pragma unroll 2
        for (int g = 0; g < 8; ++g) {
            const int64_t i = slab * 32 + g * 4 + kk;
            const bool ok = i < n;
            double xa[T], y1[T], y2[SPLIT ? 1 : T];
#pragma unroll
            for (int t = 0; t < T; ++t) {
                xa[t] = ok ? __ldg(X + i * BW + t * 8 + mm) : 0.0;
                y1[t] = ok ? __ldg(Ya + i * BW + t * 8 + mm) : 0.0;
                if (!SPLIT) y2[t] = ok ? __ldg(Y2 + i * BW + t * 8 + mm) : 0.0;
            }
#pragma unroll
            for (int a = 0; a < T; ++a)
#pragma unroll
                for (int b = 0; b < T; ++b) {
                    lz_dmma(acc[0][a][b][0], acc[0][a][b][1], xa[a], y1[b]);
                    if (!SPLIT) lz_dmma(acc[NP - 1][a][b][0], acc[NP - 1][a][b][1], xa[a], y2[b]);
                }
        }
    }
    __shared__ double sm[BW * BW];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        for (int w = 0; w < WPP; ++w) {
            const int wv = SPLIT ? h * WPP + w : w;        // the warp whose accumulators are added in this turn
            if (warp == wv) {
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) {
                        const int p = a * 8 + mm, q = b * 8 + 2 * kk;
                        const double v0 = acc[SPLIT ? 0 : h][a][b][0], v1 = acc[SPLIT ? 0 : h][a][b][1];
                        if (w == 0) { sm[p + q * BW] = v0; sm[p + (q + 1) * BW] = v1; }
                        else { sm[p + q * BW] += v0; sm[p + (q + 1) * BW] += v1; }
                    }
            }
            __syncthreads();
        }
        for (int e = threadIdx.x; e < BW * BW; e += LZ_DENSE_THREADS) gpart[((size_t)blockIdx.x * 2 + h) * BW * BW + e] = sm[e];
        __syncthreads();
    }
}

// block DGKS test: a second Gram-Schmidt sweep is needed iff the first one removed more than half of the squared norm
// of SOME column of W ("twice is enough", per column: diag(W'^T W') < 0.5 diag(W^T W)).  flag <- 1 / 0.
static __global__ void k_block_dgks_test(int bw, const double *__restrict__ G_before, const double *__restrict__ G_after, int *flag)
{
    __shared__ int need;
    if (threadIdx.x == 0) need = 0;
    __syncthreads();
    if (threadIdx.x < bw) {
        const double b = G_before[threadIdx.x + threadIdx.x * bw], a = G_after[threadIdx.x + threadIdx.x * bw];
        if (!(a >= 0.5 * b)) atomicOr(&need, 1);          // also true for NaN: when in doubt, sweep again
    }
    __syncthreads();
    if (threadIdx.x == 0) *flag = need;
}

// alpha = sym(G1 - G2 B): one CTA, b x b column-major matrices
static __global__ void __launch_bounds__(256)
k_alpha_from_grams(int bw, const double *__restrict__ G1, const double *__restrict__ G2, const double *__restrict__ Bm, double *__restrict__ alpha)
{
    __shared__ double t[32 * 32];
    for (int e = threadIdx.x; e < bw * bw; e += blockDim.x) {
        const int p = e % bw, q = e / bw;
        double s = G1[e];
        for (int k = 0; k < bw; ++k) s = fma(-G2[p + k * bw], Bm[k + q * bw], s);
        t[e] = s;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < bw * bw; e += blockDim.x) {
        const int p = e % bw, q = e / bw;
        alpha[e] = 0.5 * t[p + q * bw] + 0.5 * t[q + p * bw];
    }
}

// ---------------------------------------------------------------------------------------------
// R = beta * R + alpha * T S     (T, R: n x BW panels; S: BW x BW column-major, device).
// Optionally (GRAM) also G_partial = R_new^T R_new, fused in the same pass through a per-warp
// shared-memory transpose of the freshly written 8-row group.
// R may alias T (rows are independent and every group reads T before it writes R).
// ---------------------------------------------------------------------------------------------
template <int BW, bool RMT, bool RMR, bool GRAM>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_panel_dmma(int64_t n, const double *T_, int64_t ldt, const double *__restrict__ S, double beta, double alpha,
             double *R_, int64_t ldr, double *__restrict__ gpart)
{
    constexpr int NT = BW / 8, KT = BW / 4;
    const LzLay<RMT> lt{ldt, BW};
    const LzLay<RMR> lr{ldr, BW};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    // B fragments of alpha*S: B[k][nn] = S[kt*4 + kk, nt*8 + mm]
    double sb[KT][NT];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) sb[kt][nt] = alpha * S[(kt * 4 + kk) + (nt * 8 + mm) * BW];
    double gacc[GRAM ? NT : 1][GRAM ? NT : 1][2];
    if (GRAM) {
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) gacc[a][b][0] = gacc[a][b][1] = 0.0;
    }
    __shared__ double stage[GRAM ? LZ_DENSE_WARPS : 1][GRAM ? 8 * (BW + 1) : 1];

    const int64_t n_slabs = (n + 31) / 32;
    const int64_t wglobal = (int64_t)blockIdx.x * LZ_DENSE_WARPS + warp, wtotal = (int64_t)gridDim.x * LZ_DENSE_WARPS;
    for (int64_t slab = wglobal; slab < n_slabs; slab += wtotal) {
#pragma unroll 2
        for (int g = 0; g < 4; ++g) {
            const int64_t i = slab * 32 + g * 8 + mm;      // this thread's row in A / C fragments
            const bool ok = i < n;
            double ta[KT];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) ta[kt] = ok ? T_[lt.at(i, kt * 4 + kk)] : 0.0;
            double d[NT][2];
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                d[nt][0] = d[nt][1] = 0.0;
                if (beta != 0.0 && ok) {
                    d[nt][0] = beta * R_[lr.at(i, nt * 8 + 2 * kk)];
                    d[nt][1] = beta * R_[lr.at(i, nt * 8 + 2 * kk + 1)];
                }
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) lz_dmma(d[nt][0], d[nt][1], ta[kt], sb[kt][nt]);
            if (ok) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    if (RMR) {
                        *reinterpret_cast<double2 *>(R_ + lr.at(i, nt * 8 + 2 * kk)) = make_double2(d[nt][0], d[nt][1]);
                    } else {
                        R_[lr.at(i, nt * 8 + 2 * kk)] = d[nt][0];
                        R_[lr.at(i, nt * 8 + 2 * kk + 1)] = d[nt][1];
                    }
                }
            }
            if (GRAM) {
                // 8 new rows -> shared (row mm, cols) -> Gram fragments: two k-steps of 4 rows
                double *st = stage[warp];
                __syncwarp();
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    st[mm * (BW + 1) + nt * 8 + 2 * kk] = ok ? d[nt][0] : 0.0;
                    st[mm * (BW + 1) + nt * 8 + 2 * kk + 1] = ok ? d[nt][1] : 0.0;
                }
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double fr[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) fr[t] = st[(h * 4 + kk) * (BW + 1) + t * 8 + mm];
#pragma unroll
                    for (int a = 0; a < NT; ++a)
#pragma unroll
                        for (int b = 0; b < NT; ++b) lz_dmma(gacc[a][b][0], gacc[a][b][1], fr[a], fr[b]);
                }
            }
        }
    }
    if (GRAM) {
        __shared__ double sm[BW * BW];
        for (int w = 0; w < LZ_DENSE_WARPS; ++w) {
            if (warp == w) {
#pragma unroll
                for (int a = 0; a < NT; ++a)
#pragma unroll
                    for (int b = 0; b < NT; ++b) {
                        const int p = a * 8 + mm, q = b * 8 + 2 * kk;
                        if (w == 0) { sm[p + q * BW] = gacc[a][b][0]; sm[p + (q + 1) * BW] = gacc[a][b][1]; }
                        else { sm[p + q * BW] += gacc[a][b][0]; sm[p + (q + 1) * BW] += gacc[a][b][1]; }
                    }
            }
            __syncthreads();
        }
        for (int e = threadIdx.x; e < BW * BW; e += LZ_DENSE_THREADS) gpart[(size_t)blockIdx.x * BW * BW + e] = sm[e];
    }
}

// generic SIMT panel update for any bw: one thread per row, S in shared memory
template <bool RMT, bool RMR>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_panel_simt(int64_t n, int bw, const double *T_, int64_t ldt, const double *__restrict__ S, double beta, double alpha,
             double *R_, int64_t ldr)
{
    const LzLay<RMT> lt{ldt, bw};
    const LzLay<RMR> lr{ldr, bw};
    __shared__ double ss[32 * 32];
    for (int e = threadIdx.x; e < bw * bw; e += LZ_DENSE_THREADS) ss[e] = S[e];
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * LZ_DENSE_THREADS;
    for (int64_t i = (int64_t)blockIdx.x * LZ_DENSE_THREADS + threadIdx.x; i < n; i += stride) {
        double row[32];
        for (int k = 0; k < bw; ++k) row[k] = T_[lt.at(i, k)];
        for (int j = 0; j < bw; ++j) {
            double s = 0.0;
            for (int k = 0; k < bw; ++k) s = fma(row[k], ss[k + j * bw], s);
            const double r = beta != 0.0 ? beta * R_[lr.at(i, j)] : 0.0;
            R_[lr.at(i, j)] = r + alpha * s;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Block reorthogonalisation against J stored row-major blocks V_0..V_{J-1} (each n x BW, `pan`
// doubles apart).  Two kernels per classical Gram-Schmidt sweep, each streaming the basis once:
//   k_block_project : C_j = V_j^T W for a batch of JB stored blocks per CTA row (blockIdx.y); the W
//                     fragments of a 4-row group are loaded once and reused for the JB blocks, so W
//                     is re-read J/JB times instead of J times.
//   k_block_update  : W -= sum_j V_j C_j ; the W tile IS the DMMA accumulator, so W is read and
//                     written exactly once per sweep; the C_j fragments are reused over four 8-row groups.
// flops 4 n (J BW) BW per sweep over 16 n J BW bytes: AI -> BW/4 flop/B.
// ---------------------------------------------------------------------------------------------
template <int BW, int JB>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_block_project(int64_t n, int J, const double *__restrict__ V, int64_t pan, const double *__restrict__ W,
                double *__restrict__ gpart /* [gridDim.y][gridDim.x][JB][BW*BW] */, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;          // conditional second sweep (block DGKS)
    constexpr int T = BW / 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    const int j0 = blockIdx.y * JB;
    double acc[JB][T][T][2];
#pragma unroll
    for (int jb = 0; jb < JB; ++jb)
#pragma unroll
        for (int a = 0; a < T; ++a)
#pragma unroll
            for (int b = 0; b < T; ++b) acc[jb][a][b][0] = acc[jb][a][b][1] = 0.0;
    const int64_t n_slabs = (n + 31) / 32;
    const int64_t wglobal = (int64_t)blockIdx.x * LZ_DENSE_WARPS + warp, wtotal = (int64_t)gridDim.x * LZ_DENSE_WARPS;
    for (int64_t slab = wglobal; slab < n_slabs; slab += wtotal) {
#pragma unroll 2
        for (int g = 0; g < 8; ++g) {
            const int64_t i = slab * 32 + g * 4 + kk;
            const bool ok = i < n;
            double yb[T];
#pragma unroll
            for (int t = 0; t < T; ++t) yb[t] = ok ? __ldg(W + i * BW + t * 8 + mm) : 0.0;
#pragma unroll
            for (int jb = 0; jb < JB; ++jb) {
                if (j0 + jb < J) {
                    const double *Vj = V + (int64_t)(j0 + jb) * pan;
                    double xa[T];
#pragma unroll
                    for (int t = 0; t < T; ++t) xa[t] = ok ? __ldcs(Vj + i * BW + t * 8 + mm) : 0.0;
#pragma unroll
                    for (int a = 0; a < T; ++a)
#pragma unroll
                        for (int b = 0; b < T; ++b) lz_dmma(acc[jb][a][b][0], acc[jb][a][b][1], xa[a], yb[b]);
                }
            }
        }
    }
    __shared__ double sm[BW * BW];
#pragma unroll
    for (int jb = 0; jb < JB; ++jb) {
        for (int w = 0; w < LZ_DENSE_WARPS; ++w) {
            if (warp == w) {
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) {
                        const int p = a * 8 + mm, q = b * 8 + 2 * kk;
                        if (w == 0) { sm[p + q * BW] = acc[jb][a][b][0]; sm[p + (q + 1) * BW] = acc[jb][a][b][1]; }
                        else { sm[p + q * BW] += acc[jb][a][b][0]; sm[p + (q + 1) * BW] += acc[jb][a][b][1]; }
                    }
            }
            __syncthreads();
        }
        double *out = gpart + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * JB + jb) * BW * BW;
        for (int e = threadIdx.x; e < BW * BW; e += LZ_DENSE_THREADS) out[e] = sm[e];
        __syncthreads();
    }
}

// C[j] = sum over CTAs of the partials of stored block j (fixed order): one warp per entry, lanes stride over the partials
template <int JB>
static __global__ void __launch_bounds__(1024)
k_block_project_reduce(int bw, int J, int n_parts, const double *__restrict__ gpart, double *__restrict__ C, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;
    const int j = blockIdx.x, batch = j / JB, jb = j % JB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = warp; e < bw * bw; e += nw) {
        double s = 0.0;
        for (int p = lane; p < n_parts; p += 32) s += gpart[(((size_t)batch * n_parts + p) * JB + jb) * bw * bw + e];
        s = lz_warp_sum(s);
        if (lane == 0) C[(size_t)j * bw * bw + e] = s;
    }
}

template <int BW>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_block_update(int64_t n, int J, const double *__restrict__ V, int64_t pan, const double *__restrict__ C, double *__restrict__ W,
               const int *__restrict__ run_flag = nullptr)
{
    if (run_flag && *run_flag == 0) return;
    constexpr int NT = BW / 8, KT = BW / 4, G = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    const int64_t n_slabs = (n + 31) / 32;
    const int64_t wglobal = (int64_t)blockIdx.x * LZ_DENSE_WARPS + warp, wtotal = (int64_t)gridDim.x * LZ_DENSE_WARPS;
    for (int64_t slab = wglobal; slab < n_slabs; slab += wtotal) {
        double d[G][NT][2];
        int64_t row[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            row[g] = slab * 32 + g * 8 + mm;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double2 v = row[g] < n ? *reinterpret_cast<const double2 *>(W + row[g] * BW + nt * 8 + 2 * kk) : make_double2(0.0, 0.0);
                d[g][nt][0] = v.x; d[g][nt][1] = v.y;
            }
        }
        for (int j = 0; j < J; ++j) {
            const double *Vj = V + (int64_t)j * pan, *Cj = C + (size_t)j * BW * BW;
            double sb[KT][NT];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) sb[kt][nt] = -__ldg(Cj + (kt * 4 + kk) + (nt * 8 + mm) * BW);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                double ta[KT];
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) ta[kt] = row[g] < n ? __ldcs(Vj + row[g] * BW + kt * 4 + kk) : 0.0;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int kt = 0; kt < KT; ++kt) lz_dmma(d[g][nt][0], d[g][nt][1], ta[kt], sb[kt][nt]);
            }
        }
#pragma unroll
        for (int g = 0; g < G; ++g)
            if (row[g] < n) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    *reinterpret_cast<double2 *>(W + row[g] * BW + nt * 8 + 2 * kk) = make_double2(d[g][nt][0], d[g][nt][1]);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Wide-load versions for BW in {16, 32}: fragment-ordered 8-byte loads cost one L1 wavefront per 64
// bytes (four half-lines per instruction), and the L1 data pipe -- not HBM -- bounds the kernel.
// The MMA only needs a consistent labelling of rows/columns inside a tile, so the tiles are built
// from PERMUTED columns and every load becomes a full-line vector load with no shuffle:
//   project: lane (kk,mm) loads 16 bytes  X[i0+kk, 16v+2mm .. +1]  -> tile 2v+h owns columns 16v+2mm+h
//   update : lane (kk,mm) loads 32 bytes  V[i0+mm, 16c+4kk .. +3]  -> k-tile 4c+e owns columns 16c+4kk+e
// and the C_j operand of the update is pre-arranged in fragment order (Cf) by the reduce kernel.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 lz_ld128_stream(const double *p)
{
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

template <int BW, int JB>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_block_project_w(int64_t n, int J, const double *__restrict__ V, int64_t pan, const double *__restrict__ W,
                  double *__restrict__ gpart /* [gridDim.y][gridDim.x][JB][BW*BW] */, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;          // conditional second sweep (block DGKS)
    constexpr int T = BW / 8, NV = BW / 16;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    const int j0 = blockIdx.y * JB;
    double acc[JB][T][T][2];
#pragma unroll
    for (int jb = 0; jb < JB; ++jb)
#pragma unroll
        for (int a = 0; a < T; ++a)
#pragma unroll
            for (int b = 0; b < T; ++b) acc[jb][a][b][0] = acc[jb][a][b][1] = 0.0;
    const int64_t n_slabs = (n + 31) / 32;
    const int64_t wglobal = (int64_t)blockIdx.x * LZ_DENSE_WARPS + warp, wtotal = (int64_t)gridDim.x * LZ_DENSE_WARPS;
    for (int64_t slab = wglobal; slab < n_slabs; slab += wtotal) {
#pragma unroll 2
        for (int g = 0; g < 8; ++g) {
            const int64_t i = slab * 32 + g * 4 + kk;
            const bool ok = i < n;
            double yb[T];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const double2 t = ok ? __ldg(reinterpret_cast<const double2 *>(W + i * BW + 16 * v + 2 * mm)) : make_double2(0.0, 0.0);
                yb[2 * v] = t.x; yb[2 * v + 1] = t.y;
            }
            // all loads of the group first, then the MMAs: a warp waits once per group, not once per block
            double xa[JB][T];
#pragma unroll
            for (int jb = 0; jb < JB; ++jb) {
                const bool live = ok && (j0 + jb < J);
                const double *Vj = V + (int64_t)(j0 + jb) * pan;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const double2 t = live ? lz_ld128_stream(Vj + i * BW + 16 * v + 2 * mm) : make_double2(0.0, 0.0);
                    xa[jb][2 * v] = t.x; xa[jb][2 * v + 1] = t.y;
                }
            }
#pragma unroll
            for (int jb = 0; jb < JB; ++jb)
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b) lz_dmma(acc[jb][a][b][0], acc[jb][a][b][1], xa[jb][a], yb[b]);
        }
    }
    __shared__ double sm[BW * BW];
#pragma unroll
    for (int jb = 0; jb < JB; ++jb) {
        for (int w = 0; w < LZ_DENSE_WARPS; ++w) {
            if (warp == w) {
#pragma unroll
                for (int a = 0; a < T; ++a)
#pragma unroll
                    for (int b = 0; b < T; ++b)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int p = 16 * (a / 2) + 2 * mm + (a % 2);
                            const int q = 16 * (b / 2) + 2 * (2 * kk + e) + (b % 2);
                            if (w == 0) sm[p + q * BW] = acc[jb][a][b][e];
                            else sm[p + q * BW] += acc[jb][a][b][e];
                        }
            }
            __syncthreads();
        }
        double *out = gpart + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * JB + jb) * BW * BW;
        for (int e = threadIdx.x; e < BW * BW; e += LZ_DENSE_THREADS) out[e] = sm[e];
        __syncthreads();
    }
}

// C[j] = sum of partials (fixed order, one warp per entry)
template <int BW, int JB>
static __global__ void __launch_bounds__(1024)
k_block_project_reduce_w(int J, int n_parts, const double *__restrict__ gpart, double *__restrict__ C, const int *__restrict__ run_flag)
{
    if (run_flag && *run_flag == 0) return;
    const int j = blockIdx.x, batch = j / JB, jb = j % JB;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int e = warp; e < BW * BW; e += nw) {
        double s = 0.0;
        for (int p = lane; p < n_parts; p += 32) s += gpart[(((size_t)batch * n_parts + p) * JB + jb) * BW * BW + e];
        s = lz_warp_sum(s);
        if (lane == 0) C[(size_t)j * BW * BW + e] = s;
    }
}

// Cf[j][kt][nt][lane] = -C_j[16c+4kk+e, 8nt+mm] with kt = 4c+e: the update kernel's B fragments.  A separate (local)
// step so that sharded runs all-reduce C once and permute afterwards.
template <int BW>
static __global__ void __launch_bounds__(256)
k_block_coef_frag(int J, const double *__restrict__ C, double *__restrict__ Cf, const int *__restrict__ run_flag = nullptr)
{
    if (run_flag && *run_flag == 0) return;
    constexpr int KT = BW / 4, NT = BW / 8;
    const int j = blockIdx.x;
    for (int e = threadIdx.x; e < KT * NT * 32; e += blockDim.x) {
        const int lane = e % 32, nt = (e / 32) % NT, kt = e / (32 * NT);
        const int kk = lane & 3, mm = lane >> 2, c = kt / 4, el = kt % 4;
        Cf[(size_t)j * BW * BW + e] = -C[(size_t)j * BW * BW + (16 * c + 4 * kk + el) + (nt * 8 + mm) * BW];
    }
}

template <int BW>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_block_update_w(int64_t n, int J, const double *__restrict__ V, int64_t pan, const double *__restrict__ Cf, double *__restrict__ W,
                 const int *__restrict__ run_flag = nullptr)
{
    if (run_flag && *run_flag == 0) return;
    constexpr int NT = BW / 8, KT = BW / 4, NC = BW / 16, G = BW == 16 ? 4 : 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    const int64_t n_slabs = (n + 8 * G - 1) / (8 * G);
    const int64_t wglobal = (int64_t)blockIdx.x * LZ_DENSE_WARPS + warp, wtotal = (int64_t)gridDim.x * LZ_DENSE_WARPS;
    for (int64_t slab = wglobal; slab < n_slabs; slab += wtotal) {
        double d[G][NT][2];
        int64_t row[G];
#pragma unroll
        for (int g = 0; g < G; ++g) {
            row[g] = slab * (8 * G) + g * 8 + mm;
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double2 v = row[g] < n ? *reinterpret_cast<const double2 *>(W + row[g] * BW + nt * 8 + 2 * kk) : make_double2(0.0, 0.0);
                d[g][nt][0] = v.x; d[g][nt][1] = v.y;
            }
        }
        // operands of stored block j: the V rows of the slab (one 256-bit load per 16 columns) and
        // the fragment-ordered C_j; block j+1 is fetched while block j is multiplied
        double ta[2][G][KT];
        auto fetch = [&](int buf, int j) {
            const double *Vj = V + (int64_t)j * pan;
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    if (row[g] < n) lz_ld256_stream(Vj + row[g] * BW + 16 * c + 4 * kk, ta[buf][g][4 * c], ta[buf][g][4 * c + 1], ta[buf][g][4 * c + 2], ta[buf][g][4 * c + 3]);
                    else ta[buf][g][4 * c] = ta[buf][g][4 * c + 1] = ta[buf][g][4 * c + 2] = ta[buf][g][4 * c + 3] = 0.0;
                }
        };
        auto multiply = [&](int buf, int j) {
            const double *Cj = Cf + (size_t)j * BW * BW;      // a few KB per block, L1-resident across the CTA's warps
            double sb[KT][NT];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) sb[kt][nt] = __ldg(Cj + (kt * NT + nt) * 32 + lane);
#pragma unroll
            for (int g = 0; g < G; ++g)
#pragma unroll
                for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) lz_dmma(d[g][nt][0], d[g][nt][1], ta[buf][g][kt], sb[kt][nt]);
        };
        fetch(0, 0);
        int j = 0;
        for (; j + 2 <= J; j += 2) {
            fetch(1, j + 1);
            multiply(0, j);
            if (j + 2 < J) fetch(0, j + 2);
            multiply(1, j + 1);
        }
        if (j < J) multiply(0, j);
#pragma unroll
        for (int g = 0; g < G; ++g)
            if (row[g] < n) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    *reinterpret_cast<double2 *>(W + row[g] * BW + nt * 8 + 2 * kk) = make_double2(d[g][nt][0], d[g][nt][1]);
            }
    }
}

// ---------------------------------------------------------------------------------------------
// W -= T1 S1 + T2 S2  (row-major panels, S column-major BW x BW), optionally G_partial = W_new^T W_new.
// One pass for the two subtractions of a block Lanczos step (W -= Q_{j-1} beta_j, W -= Q_j alpha_j)
// with the next step's Gram fused: W is read and written once.
// ---------------------------------------------------------------------------------------------
template <int BW, bool GRAM>
__global__ void __launch_bounds__(LZ_DENSE_THREADS)
k_panel2_dmma(int64_t n, const double *__restrict__ T1, const double *__restrict__ S1, const double *__restrict__ T2,
              const double *__restrict__ S2, double *__restrict__ R_, double *__restrict__ gpart)
{
    constexpr int NT = BW / 8, KT = BW / 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    double s1[KT][NT], s2[KT][NT];
#pragma unroll
    for (int kt = 0; kt < KT; ++kt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            s1[kt][nt] = -S1[(kt * 4 + kk) + (nt * 8 + mm) * BW];
            s2[kt][nt] = -S2[(kt * 4 + kk) + (nt * 8 + mm) * BW];
        }
    double gacc[GRAM ? NT : 1][GRAM ? NT : 1][2];
    if (GRAM) {
#pragma unroll
        for (int a = 0; a < NT; ++a)
#pragma unroll
            for (int b = 0; b < NT; ++b) gacc[a][b][0] = gacc[a][b][1] = 0.0;
    }
    __shared__ double stage[GRAM ? LZ_DENSE_WARPS : 1][GRAM ? 8 * (BW + 1) : 1];
    const int64_t n_slabs = (n + 31) / 32;
    const int64_t wglobal = (int64_t)blockIdx.x * LZ_DENSE_WARPS + warp, wtotal = (int64_t)gridDim.x * LZ_DENSE_WARPS;
    for (int64_t slab = wglobal; slab < n_slabs; slab += wtotal) {
#pragma unroll 2
        for (int g = 0; g < 4; ++g) {
            const int64_t i = slab * 32 + g * 8 + mm;
            const bool ok = i < n;
            double ta[KT], tb[KT], d[NT][2];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
                ta[kt] = ok ? __ldg(T1 + i * BW + kt * 4 + kk) : 0.0;
                tb[kt] = ok ? __ldg(T2 + i * BW + kt * 4 + kk) : 0.0;
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                const double2 v = ok ? *reinterpret_cast<const double2 *>(R_ + i * BW + nt * 8 + 2 * kk) : make_double2(0.0, 0.0);
                d[nt][0] = v.x; d[nt][1] = v.y;
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    lz_dmma(d[nt][0], d[nt][1], ta[kt], s1[kt][nt]);
                    lz_dmma(d[nt][0], d[nt][1], tb[kt], s2[kt][nt]);
                }
            if (ok) {
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    *reinterpret_cast<double2 *>(R_ + i * BW + nt * 8 + 2 * kk) = make_double2(d[nt][0], d[nt][1]);
            }
            if (GRAM) {
                double *st = stage[warp];
                __syncwarp();
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    st[mm * (BW + 1) + nt * 8 + 2 * kk] = ok ? d[nt][0] : 0.0;
                    st[mm * (BW + 1) + nt * 8 + 2 * kk + 1] = ok ? d[nt][1] : 0.0;
                }
                __syncwarp();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    double fr[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) fr[t] = st[(h * 4 + kk) * (BW + 1) + t * 8 + mm];
#pragma unroll
                    for (int a = 0; a < NT; ++a)
#pragma unroll
                        for (int b = 0; b < NT; ++b) lz_dmma(gacc[a][b][0], gacc[a][b][1], fr[a], fr[b]);
                }
            }
        }
    }
    if (GRAM) {
        __shared__ double sm[BW * BW];
        for (int w = 0; w < LZ_DENSE_WARPS; ++w) {
            if (warp == w) {
#pragma unroll
                for (int a = 0; a < NT; ++a)
#pragma unroll
                    for (int b = 0; b < NT; ++b) {
                        const int p = a * 8 + mm, q = b * 8 + 2 * kk;
                        if (w == 0) { sm[p + q * BW] = gacc[a][b][0]; sm[p + (q + 1) * BW] = gacc[a][b][1]; }
                        else { sm[p + q * BW] += gacc[a][b][0]; sm[p + (q + 1) * BW] += gacc[a][b][1]; }
                    }
            }
            __syncthreads();
        }
        for (int e = threadIdx.x; e < BW * BW; e += LZ_DENSE_THREADS) gpart[(size_t)blockIdx.x * BW * BW + e] = sm[e];
    }
}
