// lz_maxwell.cu -- the reference's test operator assembled ON THE DEVICE.
//
// Matrix_A(Nx, Ny, Nz) of matrix_a/build_A_ell.hpp:8-255 is the 3-D Maxwell curl operator on a staggered (Yee)
// grid, D = [0 Dh; De 0] in ELL of width 4, made symmetric by the diagonal metric W: A = D W
// (Ell_matrix::mult_diagonal, objects/ell_matrix.hpp:340-361).  The reference -- and the C++ mirror's Host builder --
// form it with host loops over Kronecker products (seconds at N = 160, n = 24.8 M).  Every entry has a closed form:
//   block K3(a, b, c) = a (x) (b (x) c) with exactly one 1-D difference factor (F or B, two entries per row) and two
//   identities:  value = a(rz,ka) * (b(ry,kb) * c(rx,kc)),  column = a.col * (b.cols c.cols) + b.col * c.cols + c.col,
//   negated blocks multiply the product by -1, and A's entry is that value times W[column], W being the same kind of
//   triple product of the 1-D cell sizes.
// One thread per row evaluates its four slots from the six 1-D tables (computed on the host with the reference's own
// Linspace / Diff arithmetic, a few KB) in the reference's operation order, so the arrays are BIT-IDENTICAL to the
// host builder's (tests/golden/maxwell_N*_matrix.npz).  Output: the device format of the reference, row-interleaved
// width-4 ELL, plus the CSR shadow the block path uses.
#include <vector>

#include "lz_common.cuh"

enum { FT_F = 0, FT_B = 1, FT_I = 2, FT_IP = 3 };       // 1-D factors: forward / backward difference, identities N and N+1
enum { MT_W = 0, MT_WH = 1 };                           // 1-D metrics: primal (N+1 cells) / dual (N cells)

struct MxAxis {
    int N;
    const double *Fv, *Bv, *dp, *dd;    // F: (N+1) x 2 values, B: N x 2 values, cell sizes
    const int *Fc, *Bc;                 // column ids of the difference entries
};
struct MxBlock { int t[3]; double sign; int64_t shift; };          // factor types for (z, y, x), sign, column shift
struct MxParams {
    MxAxis ax[3];                       // z, y, x
    MxBlock blk[12];                    // Dh12 Dh13 Dh21 Dh23 Dh31 Dh32 De12 De13 De21 De23 De31 De32
    int64_t row_end[6];                 // cumulative row ends of the six row blocks (E1 E2 E3 H1 H2 H3)
    int wt[6][3];                       // metric types (z, y, x) of the six diagonal blocks of W
    double wsign[6];
    int64_t n_rows;
};

__device__ __forceinline__ int mx_rows(const MxAxis &a, int t) { return (t == FT_F || t == FT_IP) ? a.N + 1 : a.N; }
__device__ __forceinline__ int mx_cols(const MxAxis &a, int t) { return (t == FT_B || t == FT_IP) ? a.N + 1 : a.N; }
__device__ __forceinline__ int mx_width(int t) { return (t == FT_F || t == FT_B) ? 2 : 1; }
__device__ __forceinline__ void mx_factor(const MxAxis &a, int t, int r, int k, double &v, int &c)
{
    if (t == FT_F) { v = a.Fv[2 * r + k]; c = a.Fc[2 * r + k]; }
    else if (t == FT_B) { v = a.Bv[2 * r + k]; c = a.Bc[2 * r + k]; }
    else { v = 1.0; c = r; }
}
__device__ __forceinline__ int mx_msize(const MxAxis &a, int t) { return t == MT_W ? a.N + 1 : a.N; }
__device__ __forceinline__ double mx_metric(const MxAxis &a, int t, int i) { return t == MT_W ? a.dp[i] : a.dd[i]; }

// W[col]: find the diagonal block, split the index into (z, y, x), multiply outer * (middle * inner), apply the sign
__device__ double mx_w(const MxParams &P, int64_t col)
{
    int b = 0;
    int64_t base = 0;
    while (b < 5 && col >= P.row_end[b]) { base = P.row_end[b]; ++b; }
    const int64_t i = col - base;
    const int sy = mx_msize(P.ax[1], P.wt[b][1]), sx = mx_msize(P.ax[2], P.wt[b][2]);
    const int iz = (int)(i / ((int64_t)sy * sx)), rem = (int)(i % ((int64_t)sy * sx)), iy = rem / sx, ix = rem % sx;
    double v = __dmul_rn(mx_metric(P.ax[0], P.wt[b][0], iz), __dmul_rn(mx_metric(P.ax[1], P.wt[b][1], iy), mx_metric(P.ax[2], P.wt[b][2], ix)));
    if (P.wsign[b] != 1.0) v = __dmul_rn(v, P.wsign[b]);
    return v;
}

__global__ void __launch_bounds__(256)
k_maxwell(const MxParams P, double *__restrict__ data, uint32_t *__restrict__ idx)
{
    const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (row >= P.n_rows) return;
    int rb = 0;
    int64_t base = 0;
    while (rb < 5 && row >= P.row_end[rb]) { base = P.row_end[rb]; ++rb; }
    const int64_t r = row - base;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const MxBlock &B = P.blk[2 * rb + half];
        const int ry_n = mx_rows(P.ax[1], B.t[1]), rx_n = mx_rows(P.ax[2], B.t[2]);
        const int rz = (int)(r / ((int64_t)ry_n * rx_n)), rem = (int)(r % ((int64_t)ry_n * rx_n)), ry = rem / rx_n, rx = rem % rx_n;
        const int wy = mx_width(B.t[1]), wx = mx_width(B.t[2]);
        const int64_t cy_n = mx_cols(P.ax[1], B.t[1]), cx_n = mx_cols(P.ax[2], B.t[2]);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int ka = s / (wy * wx), kb = (s / wx) % wy, kc = s % wx;
            double va, vb, vc;
            int ca, cb, cc;
            mx_factor(P.ax[0], B.t[0], rz, ka, va, ca);
            mx_factor(P.ax[1], B.t[1], ry, kb, vb, cb);
            mx_factor(P.ax[2], B.t[2], rx, kc, vc, cc);
            double v = __dmul_rn(va, __dmul_rn(vb, vc));                 // ell_kron(a, ell_kron(b, c))
            if (B.sign != 1.0) v = __dmul_rn(v, B.sign);                 // mult_scalar(-1.)
            const int64_t col = (int64_t)ca * (cy_n * cx_n) + ((int64_t)cb * cx_n + cc) + B.shift;
            data[4 * row + 2 * half + s] = __dmul_rn(v, mx_w(P, col));   // mult_diagonal(W)
            idx[4 * row + 2 * half + s] = (uint32_t)col;
        }
    }
}

// the reference's 1-D grids (build_ell_utils.hpp: Linspace, Diff; build_A_ell.hpp: the two difference factors)
static void host_axis(int N, std::vector<double> &Fv, std::vector<int> &Fc, std::vector<double> &Bv, std::vector<int> &Bc,
                      std::vector<double> &dp, std::vector<double> &dd)
{
    const double lo = 0., hi = 1.;
    const unsigned int Np = (unsigned int)N + 2;
    const double h = (hi - lo) / (Np - 1);
    std::vector<double> p(Np), d(Np - 1);
    for (unsigned int i = 0; i < Np; ++i) p[i] = lo + i * h;
    const double h2 = ((hi - h) - lo) / ((Np - 1) - 1);
    for (unsigned int i = 0; i < Np - 1; ++i) d[i] = lo + i * h2;
    for (unsigned int i = 0; i < Np - 1; ++i) d[i] = d[i] + h / 2;
    dp.resize(N + 1); dd.resize(N);
    for (int i = 0; i < N + 1; ++i) dp[i] = p[i + 1] - p[i];
    for (int i = 0; i < N; ++i) dd[i] = d[i + 1] - d[i];
    Fv.assign(2 * (N + 1), 0.0); Fc.assign(2 * (N + 1), 0);
    for (int r = 0; r <= N; ++r) {
        const double inv = 1. / dp[r];
        int slot = 0;
        if (r >= 1) { Fv[2 * r + slot] = inv * -1.; Fc[2 * r + slot] = r - 1; ++slot; }
        if (r < N) { Fv[2 * r + slot] = inv * 1.; Fc[2 * r + slot] = r; }
    }
    Bv.assign(2 * N, 0.0); Bc.assign(2 * N, 0);
    for (int r = 0; r < N; ++r) {
        const double inv = 1. / dd[r];
        Bv[2 * r] = 0. * (inv * 1.) + -1. * (inv * 1.);   Bc[2 * r] = r;
        Bv[2 * r + 1] = 0. * (inv * -1.) + -1. * (inv * -1.); Bc[2 * r + 1] = r + 1;
    }
}

extern "C" int lz_gen_maxwell(lz_ctx *ctx, int Nx, int Ny, int Nz, lz_matrix **out)
{
    LZ_CHECK(ctx && out && Nx >= 1 && Ny >= 1 && Nz >= 1, LZ_ERR_INVALID, "lz_gen_maxwell: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    const int Ns[3] = {Nz, Ny, Nx};
    MxParams P;
    memset(&P, 0, sizeof(P));
    // 1-D tables -> one device buffer
    std::vector<double> dbuf;
    std::vector<int> ibuf;
    size_t doff[3][4], ioff[3][2];
    for (int a = 0; a < 3; ++a) {
        std::vector<double> Fv, Bv, dp, dd;
        std::vector<int> Fc, Bc;
        host_axis(Ns[a], Fv, Fc, Bv, Bc, dp, dd);
        doff[a][0] = dbuf.size(); dbuf.insert(dbuf.end(), Fv.begin(), Fv.end());
        doff[a][1] = dbuf.size(); dbuf.insert(dbuf.end(), Bv.begin(), Bv.end());
        doff[a][2] = dbuf.size(); dbuf.insert(dbuf.end(), dp.begin(), dp.end());
        doff[a][3] = dbuf.size(); dbuf.insert(dbuf.end(), dd.begin(), dd.end());
        ioff[a][0] = ibuf.size(); ibuf.insert(ibuf.end(), Fc.begin(), Fc.end());
        ioff[a][1] = ibuf.size(); ibuf.insert(ibuf.end(), Bc.begin(), Bc.end());
    }
    double *dtab;
    int *itab;
    LZ_CUDA(cudaMalloc(&dtab, sizeof(double) * dbuf.size()));
    LZ_CUDA(cudaMalloc(&itab, sizeof(int) * ibuf.size()));
    LZ_CUDA(cudaMemcpyAsync(dtab, dbuf.data(), sizeof(double) * dbuf.size(), cudaMemcpyHostToDevice, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(itab, ibuf.data(), sizeof(int) * ibuf.size(), cudaMemcpyHostToDevice, ctx->stream));
    for (int a = 0; a < 3; ++a)
        P.ax[a] = MxAxis{Ns[a], dtab + doff[a][0], dtab + doff[a][1], dtab + doff[a][2], dtab + doff[a][3], itab + ioff[a][0], itab + ioff[a][1]};
    auto rows = [&](int a, int t) -> int64_t { return (t == FT_F || t == FT_IP) ? Ns[a] + 1 : Ns[a]; };
    auto k3rows = [&](int tz, int ty, int tx) -> int64_t { return rows(0, tz) * rows(1, ty) * rows(2, tx); };
    // build_A_ell.hpp: the twelve curl blocks (factor types for z, y, x and the sign)
    const int T[12][3] = {
        {FT_B, FT_I, FT_IP}, {FT_I, FT_B, FT_IP},     // Dh_12,  -Dh_13
        {FT_B, FT_IP, FT_I}, {FT_I, FT_IP, FT_B},     // -Dh_21, Dh_23
        {FT_IP, FT_B, FT_I}, {FT_IP, FT_I, FT_B},     // Dh_31,  -Dh_32
        {FT_F, FT_IP, FT_I}, {FT_IP, FT_F, FT_I},     // -De_12, De_13
        {FT_F, FT_I, FT_IP}, {FT_IP, FT_I, FT_F},     // De_21,  -De_23
        {FT_I, FT_F, FT_IP}, {FT_I, FT_IP, FT_F}};    // -De_31, De_32
    const double S[12] = {1, -1, -1, 1, 1, -1, -1, 1, 1, -1, -1, 1};
    int64_t br[12];
    for (int b = 0; b < 12; ++b) br[b] = k3rows(T[b][0], T[b][1], T[b][2]);
    const int64_t Dh_rows = br[0] + br[2] + br[4], De_rows = br[6] + br[8] + br[10];
    // column shifts inside a curl (insert(C, b12, 0, 0, c1) ...) with c1 / c2 the row counts of the OTHER curl's first two
    // blocks, plus Dh's shift by Dh_rows in D = [0 Dh; De 0]
    const int64_t c1h = br[6], c2h = br[8], c1e = br[0], c2e = br[2];
    const int64_t sh[12] = {c1h + Dh_rows, c1h + c2h + Dh_rows, Dh_rows, c1h + c2h + Dh_rows, Dh_rows, c1h + Dh_rows,
                            c1e, c1e + c2e, 0, c1e + c2e, 0, c1e};
    for (int b = 0; b < 12; ++b) { P.blk[b].t[0] = T[b][0]; P.blk[b].t[1] = T[b][1]; P.blk[b].t[2] = T[b][2]; P.blk[b].sign = S[b]; P.blk[b].shift = sh[b]; }
    int64_t acc = 0;
    for (int rbk = 0; rbk < 6; ++rbk) { acc += br[2 * rbk]; P.row_end[rbk] = acc; }
    const int WT[6][3] = {{MT_WH, MT_WH, MT_W}, {MT_WH, MT_W, MT_WH}, {MT_W, MT_WH, MT_WH},
                          {MT_W, MT_W, MT_WH}, {MT_W, MT_WH, MT_W}, {MT_WH, MT_W, MT_W}};
    for (int b = 0; b < 6; ++b) { for (int a = 0; a < 3; ++a) P.wt[b][a] = WT[b][a]; P.wsign[b] = b < 3 ? 1.0 : -1.0; }
    const int64_t n = Dh_rows + De_rows;
    P.n_rows = n;
    LZ_CHECK(n < 2147483647LL / 4, LZ_ERR_UNSUPPORTED, "lz_gen_maxwell: %lld rows exceed the int32 index space", (long long)n);
    // the metric blocks must tile the columns exactly like the row blocks (D is square): checked here once
    for (int b = 0; b < 6; ++b) {
        int64_t sz = 1;
        for (int a = 0; a < 3; ++a) sz *= (WT[b][a] == MT_W ? Ns[a] + 1 : Ns[a]);
        LZ_CHECK(sz == br[2 * b], LZ_ERR_INVALID, "lz_gen_maxwell: internal block size mismatch");
    }
    lz_matrix *A = lz_new_matrix(ctx, LZ_FMT_ELL4, n, n, n * 4);
    A->owns = 1;
    double *od;
    uint32_t *oi;
    LZ_CUDA(cudaMalloc(&od, sizeof(double) * (size_t)n * 4));
    LZ_CUDA(cudaMalloc(&oi, sizeof(uint32_t) * (size_t)n * 4));
    A->ell_data = od; A->ell_idx = oi;
    A->max_row_nnz = 4;
    k_maxwell<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(P, od, oi);
    ctx->launches++;
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    cudaFree(dtab);
    cudaFree(itab);
    if (e != cudaSuccess) { lz_set_error("lz_gen_maxwell: %s", cudaGetErrorString(e)); lz_matrix_destroy(A); return LZ_ERR_CUDA; }
    int st = lz_ell4_build_shadow(ctx, A);
    if (st != LZ_OK) { lz_matrix_destroy(A); return st; }
    *out = A;
    return LZ_OK;
}
