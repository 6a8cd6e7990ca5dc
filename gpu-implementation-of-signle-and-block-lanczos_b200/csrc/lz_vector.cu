// lz_vector.cu -- single-vector path: SpMV entry point, BLAS-1 style reductions/updates, the
// classical Gram-Schmidt sweeps against the stored Krylov basis, and the vector_lanczos driver.
//
// Reference being replaced: kernels/vector_kernels.hpp (do_dot, vector_update), objects/vector.hpp
// (dot, l2_norm, sadd), utils/lib_utils.hpp:431-538 (cuBLAS level-1 wrappers) and
// methods/vector_lanczos.hpp:8-67.  Everything here keeps its scalars on the device: no
// cudaMalloc, no D2H copy and no host synchronisation inside the iteration loop.
#include <math.h>
#include <stdlib.h>

#include "lz_spmv.cuh"

#define VT 256                 // threads of the streaming kernels
#define V_ROWS_PER_THREAD 8    // two 256-bit loads per thread per column

// indices into ctx->scalars used by the drivers
enum { S_NRM2 = 0, S_NRM2_BEFORE = 1, S_ALPHA_LOCAL = 2, S_TMP = 3, S_BETA_LAST = 4 };
// indices into ctx->flags
enum { F_BREAKDOWN = 0, F_SECOND_SWEEP = 1, F_FORCE = 4, F_REORTH_COUNT = 5 };
#define SB_VECTOR_LIMIT 4096   // the vector drivers' share of the context's scalar bank (the block driver owns the rest)
// tickets
enum { T_SPMV = 0, T_DOT = 1, T_UPD = 2, T_PROJ = 3 };

// ---------------------------------------------------------------------------------------------
// dot / nrm2 / axpby
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(VT) k_dot(int64_t n, const double *__restrict__ x, const double *__restrict__ y,
                                            double *partials, unsigned int *ticket, double *out)
{
    __shared__ double red[32];
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * VT;
    for (int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x; i < n; i += stride) acc += x[i] * y[i];
    acc = lz_block_sum<VT>(acc, red);
    double total;
    if (lz_grid_sum<VT, 1>(&acc, partials, ticket, red, &total) && threadIdx.x == 0) *out = total;
}

__global__ void __launch_bounds__(VT) k_axpby(int64_t n, double a, double *__restrict__ y, double b, const double *__restrict__ x)
{
    const int64_t i = (int64_t)blockIdx.x * VT + threadIdx.x;
    if (i < n) y[i] = __dadd_rn(__dmul_rn(a, y[i]), __dmul_rn(b, x[i]));   // vector_kernels.hpp:22-33
}

static inline unsigned stream_grid(const lz_ctx *ctx, int64_t n, int per_cta)
{
    int64_t want = (n + per_cta - 1) / per_cta;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    return (unsigned)(want < 1 ? 1 : (want < cap ? want : cap));
}

static int dot_async(lz_ctx *ctx, int64_t n, const double *x, const double *y, double *out_dev)
{
    k_dot<<<stream_grid(ctx, n, VT * 4), VT, 0, ctx->stream>>>(n, x, y, ctx->partials, ctx->tickets + T_DOT, out_dev);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// Lanczos pass B:  w -= alpha_j * q_j  (q_j = u_cur * invb[j]);  nrm2 = ||w||^2 in the epilogue
// (methods/vector_lanczos.hpp:60 and :44 of the next trip, fused).
// The last CTA finalises beta_{j+1} = sqrt(nrm2), invb[j+1] and the breakdown flag.
// ---------------------------------------------------------------------------------------------
struct LzFinal {
    double *beta;        // beta[j+1] <- sqrt(total)            (NULL: skip)
    double *invb;        // invb[j+1] <- 1 / beta[j+1]
    double *nrm2_out;    // raw total
    int *flags;
    int jn;              // j + 1
    int finalize;        // 0: only publish the raw local total (NCCL-mode sharded runs all-reduce it afterwards)
    const LzPeerDesc *pd;        // peer-memory mode: the last CTA all-reduces the total over the ranks, then finalises
    unsigned long long seq;
};

// returns the total it finalised with (all-reduced over the ranks in peer mode)
__device__ __forceinline__ double lz_finalize_beta(const LzFinal &f, double total)
{
    if (f.pd) lz_peer_sum_thread<1>(f.pd, f.seq, &total);
    *f.nrm2_out = total;
    if (!f.finalize) return total;
    const double b = sqrt(total);
    if (f.beta) f.beta[f.jn] = b;
    f.invb[f.jn] = 1.0 / b;
    if (!(isfinite(total)) || total == 0.0) atomicMin(f.flags + F_BREAKDOWN, f.jn);   // vector.hpp:233-244
    return total;
}

template <bool VEC>
__global__ void __launch_bounds__(VT)
k_pass_b(int64_t n, double *__restrict__ w, const double *__restrict__ u_cur, const double *__restrict__ alpha,
         const double *__restrict__ invb, int j, double *partials, unsigned int *ticket, const LzFinal fin)
{
    __shared__ double red[32];
    const double na = -alpha[j], sx = invb[j];
    double acc = 0.0;
    const int64_t stride = (int64_t)gridDim.x * VT, t0 = (int64_t)blockIdx.x * VT + threadIdx.x;
    if (VEC) {   // both pointers 32-byte aligned: 256-bit loads/stores
        const int64_t n4 = n >> 2;
        for (int64_t g = t0; g < n4; g += stride) {
            double a[4], u[4];
            lz_ld256(w + 4 * g, a[0], a[1], a[2], a[3]);
            lz_ld256(u_cur + 4 * g, u[0], u[1], u[2], u[3]);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                a[t] = __dadd_rn(a[t], __dmul_rn(na, __dmul_rn(u[t], sx)));
                acc = fma(a[t], a[t], acc);
            }
            lz_st256(w + 4 * g, a[0], a[1], a[2], a[3]);
        }
        for (int64_t i = (n4 << 2) + t0; i < n; i += stride) {
            const double v = __dadd_rn(w[i], __dmul_rn(na, __dmul_rn(u_cur[i], sx)));
            w[i] = v;
            acc = fma(v, v, acc);
        }
    } else {
        for (int64_t i = t0; i < n; i += stride) {
            const double v = __dadd_rn(w[i], __dmul_rn(na, __dmul_rn(u_cur[i], sx)));
            w[i] = v;
            acc = fma(v, v, acc);
        }
    }
    acc = lz_block_sum<VT>(acc, red);
    double total;
    if (lz_grid_sum<VT, 1>(&acc, partials, ticket, red, &total) && threadIdx.x == 0) lz_finalize_beta(fin, total);
}

static int launch_pass_b(lz_ctx *ctx, int64_t n, double *w, const double *u_cur, const double *alpha, const double *invb,
                         int j, const LzFinal &fin)
{
    const bool vec = ((uintptr_t)w % 32 == 0) && ((uintptr_t)u_cur % 32 == 0);
    const unsigned grid = stream_grid(ctx, n, VT * 8);
    if (vec) k_pass_b<true><<<grid, VT, 0, ctx->stream>>>(n, w, u_cur, alpha, invb, j, ctx->partials, ctx->tickets + T_DOT, fin);
    else k_pass_b<false><<<grid, VT, 0, ctx->stream>>>(n, w, u_cur, alpha, invb, j, ctx->partials, ctx->tickets + T_DOT, fin);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// Classical Gram-Schmidt against the stored basis V (column j at V + j*ld, ld % 4 == 0).
//   k_cgs_project : c[k] = V[:,k] . w      for k < K   (deterministic two-stage reduction)
//   k_cgs_update  : w -= V[:, :K] c ;  ||w||^2 in the epilogue
// Both stream V exactly once with 256-bit loads; each thread keeps 8 rows of w in registers and
// walks the K columns, so w itself is read once per sweep.
// ---------------------------------------------------------------------------------------------
#define CGS_TILE (VT * V_ROWS_PER_THREAD)   // rows per CTA trip (8 rows per thread)
#define CGS_UNROLL 4
// Work granularity: a CTA takes whole tiles round-robin, so with T tiles on G CTAs the kernel lasts ceil(T / G) tile
// times -- at 2 M rows per GPU (config 2 on 8 GPUs) that is 1024 tiles on 296 CTAs = 4 rounds for 3.46 rounds of work,
// the 15 % that capped the round-1 scaling at 8 GPUs.  RPT = 4 (one 256-bit load per column and thread, twice the
// column unroll to keep the same bytes in flight) halves the tile when there are fewer than ~8 tiles per CTA.
template <int RPT> struct CgsShape { static constexpr int HALVES = RPT / 4, UNROLL = RPT == 8 ? 4 : 6, TILE = VT * RPT; };

template <int RPT>
__global__ void __launch_bounds__(VT)
k_cgs_project(int64_t n, int K, const double *__restrict__ V, int64_t ts, int64_t cs, const double *__restrict__ w,
              double *__restrict__ cpart /* gridDim.x * K */, const int *__restrict__ flags, int need_flag)
{
    constexpr int H = CgsShape<RPT>::HALVES, U = CgsShape<RPT>::UNROLL, TILE = CgsShape<RPT>::TILE;
    extern __shared__ double csm[];           // [VT/32][K] per-warp partial coefficients
    if (need_flag && flags[F_SECOND_SWEEP] == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *mine = csm + (size_t)warp * K;
    for (int k = lane; k < K; k += 32) mine[k] = 0.0;
    __syncwarp();
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE + (int64_t)warp * (32 * RPT) + lane * 4;
        // rows base..base+3 (and base+128..base+131 when RPT = 8) of this warp's slice
        int64_t r[H]; const double *p[H]; double wv[H][4]; bool full = true;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            r[h] = base + 128 * h;
            // element (i, k) of the basis lives at V[(i >> 5) * ts + k * cs + (i & 31)]  (see lz_ctx_basis)
            p[h] = V + (r[h] >> 5) * ts + (r[h] & 31);
            const bool f = r[h] + 3 < n;
            full = full && f;
#pragma unroll
            for (int t = 0; t < 4; ++t) wv[h][t] = 0.0;
            if (f) lz_ld256(w + r[h], wv[h][0], wv[h][1], wv[h][2], wv[h][3]);
            else for (int t = 0; t < 4; ++t) if (r[h] + t < n) wv[h][t] = w[r[h] + t];
        }
        int k = 0;
        if (full) {
            for (; k + U <= K; k += U) {
                double v[U][H][4];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int h = 0; h < H; ++h) lz_ld256_stream(p[h] + (int64_t)(k + u) * cs, v[u][h][0], v[u][h][1], v[u][h][2], v[u][h][3]);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    double sacc = v[u][0][0] * wv[0][0];
                    sacc = fma(v[u][0][1], wv[0][1], sacc); sacc = fma(v[u][0][2], wv[0][2], sacc); sacc = fma(v[u][0][3], wv[0][3], sacc);
#pragma unroll
                    for (int h = 1; h < H; ++h)
#pragma unroll
                        for (int t = 0; t < 4; ++t) sacc = fma(v[u][h][t], wv[h][t], sacc);
                    sacc = lz_warp_sum(sacc);
                    if (lane == 0) mine[k + u] += sacc;
                }
            }
        }
        for (; k < K; ++k) {
            const int64_t ko = (int64_t)k * cs;
            double sacc = 0.0;
            for (int h = 0; h < H; ++h)
                for (int t = 0; t < 4; ++t)
                    if (r[h] + t < n) sacc = fma(p[h][ko + t], wv[h][t], sacc);
            sacc = lz_warp_sum(sacc);
            if (lane == 0) mine[k] += sacc;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += VT) {
        double sacc = 0.0;
#pragma unroll
        for (int wv2 = 0; wv2 < VT / 32; ++wv2) sacc += csm[(size_t)wv2 * K + k];
        cpart[(size_t)blockIdx.x * K + k] = sacc;
    }
}

// ---------------------------------------------------------------------------------------------
// Fused sweep-1 update + sweep-2 projection of CGS2:  w' = w - V c1 ;  c2_partial = V^T w'.
// The two separate kernels stream the basis twice; here a CTA takes 32-row tiles -- one tile of the
// row-tiled basis is K*256 contiguous bytes, i.e. ONE bulk async copy -- into a double-buffered
// shared-memory tile (the next tile lands while this one is used), finishes w' for the tile's rows
// and immediately accumulates the second projection from the same buffer.  CGS2 then reads the basis
// three times per step instead of four.
// Phase 1: thread = (row, column group); phase 2: thread = column with a lane-skewed row order (32
// lanes on 32 banks although their columns are 32 doubles apart), register accumulators over all tiles.
// (A chunked ring with a producer warp and a deeper tile ring were measured slower: profiles/.)
// ---------------------------------------------------------------------------------------------
#define CF_R 32                 // rows per tile (= the basis layout's tile height)
#define CF_THREADS 256
#define CF_MAXC 2               // columns per thread: K <= CF_MAXC * CF_THREADS

// threads sharing one column in phase 2 (each takes CF_R / S rows); S * K <= CF_THREADS
__host__ __device__ __forceinline__ int cgs_fused_slices(int K) { return K <= 32 ? 8 : K <= 64 ? 4 : K <= 128 ? 2 : 1; }

__global__ void __launch_bounds__(CF_THREADS)
k_cgs_update_project(int64_t n, int K, const double *__restrict__ V, int64_t ts, double *__restrict__ w,
                     const double *__restrict__ c1, double *__restrict__ cpart2 /* gridDim.x * S * K */, const int S, const int NB)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // NB = 2: double-buffered tiles, one CTA per SM.  NB = 1: one tile per CTA and several CTAs per SM, so the
    // copy of one CTA overlaps the phases of another and the barriers / serial section of one CTA no longer
    // idle the SM (host picks: cgs_fused_shape)
    double *vt = reinterpret_cast<double *>(smem_raw);                          // [NB][K][CF_R]
    double *wt = vt + (size_t)NB * K * CF_R;                                    // [NB][CF_R]
    double *red = wt + NB * CF_R;                                                // [CF_THREADS/CF_R][CF_R]
    double *c1s = red + CF_THREADS;                                             // [K]
    uint64_t *full = reinterpret_cast<uint64_t *>(c1s + ((K + 1) & ~1));        // [2]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < K; k += CF_THREADS) c1s[k] = c1[k];
    if (tid == 0) {
        lz_mbar_init(&full[0], 1); lz_mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t n_tiles = (n + CF_R - 1) / CF_R;
    auto issue = [&](int64_t tile, int buf) {        // lane 0 of warp 0: the tile's K columns are one contiguous block
        if (lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            lz_mbar_expect_tx(&full[buf], (uint32_t)((K + 1) * CF_R * 8));
            lz_bulk_g2s(vt + (size_t)buf * K * CF_R, V + tile * ts, (uint32_t)K * CF_R * 8, &full[buf]);
            lz_bulk_g2s(wt + buf * CF_R, w + tile * CF_R, CF_R * 8, &full[buf]);
        }
    };
    double acc[CF_MAXC];
#pragma unroll
    for (int u = 0; u < CF_MAXC; ++u) acc[u] = 0.0;
    const int r = tid % CF_R, g = tid / CF_R;          // phase-1 role: row r, column group g of CF_THREADS/CF_R
    // phase-2 role: column k0, row slice `part` of S.  With few columns one warp would carry all of phase 2
    // while seven wait at the barrier (ncu at K = 21: 60 % barrier stalls, 0.8 us per tile whatever K), so for
    // K <= 128 the 32 rows of a column are shared by S = 2, 4, 8 threads in different warps.
    const int Kp = CF_THREADS / S, RP = CF_R / S;
    const int part = tid / Kp, k0 = tid - part * Kp, i0 = part * RP;
    if (NB == 2 && warp == 0 && (int64_t)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = NB == 2 ? (it & 1) : 0;
        if (warp == 0) {
            if (NB == 1) issue(tile, 0);
            else if (tile + gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1);
        }
        lz_mbar_wait(&full[buf], NB == 2 ? ((it >> 1) & 1) : (it & 1));
        const double *tv = vt + (size_t)buf * K * CF_R;
        double *tw = wt + buf * CF_R;
        const int rows_valid = (int)min((int64_t)CF_R, n - tile * CF_R);
        // phase 1: partial V c1 over this group's columns, then the row's new value
        double s = 0.0;
        for (int k = g; k < K; k += CF_THREADS / CF_R) s = fma(c1s[k], tv[(size_t)k * CF_R + r], s);
        red[g * CF_R + r] = s;
        __syncthreads();
        if (tid < CF_R) {
            double t = tw[tid];
#pragma unroll
            for (int gg = 0; gg < CF_THREADS / CF_R; ++gg) t -= red[gg * CF_R + tid];
            if (tid >= rows_valid) t = 0.0;            // rows past n: whatever the copy brought in is ignored
            tw[tid] = t;
            if (tid < rows_valid) w[tile * CF_R + tid] = t;
        }
        __syncthreads();
        // phase 2: this thread's columns dotted with the new rows; lane-skewed row order keeps the
        // 32 lanes of a warp on 32 different banks although their columns are CF_R doubles apart
        if (S == 1) {
#pragma unroll
            for (int u = 0; u < CF_MAXC; ++u) {
                const int k = tid + u * CF_THREADS;
                if (k < K) {
                    const double *col = tv + (size_t)k * CF_R;
                    double a = acc[u];
#pragma unroll 8
                    for (int i = 0; i < CF_R; ++i) {
                        const int rr = (i + lane) & (CF_R - 1);
                        if (rr < rows_valid) a = fma(col[rr], tw[rr], a);
                    }
                    acc[u] = a;
                }
            }
        } else if (k0 < K) {
            const double *col = tv + (size_t)k0 * CF_R;
            double a = acc[0];
#pragma unroll 4
            for (int i = i0; i < i0 + RP; ++i) {
                const int rr = (i + lane) & (CF_R - 1);
                if (rr < rows_valid) a = fma(col[rr], tw[rr], a);
            }
            acc[0] = a;
        }
        __syncthreads();                               // the buffer may be refilled from here on
    }
#pragma unroll
    for (int u = 0; u < CF_MAXC; ++u) {
        const int k = k0 + u * CF_THREADS;
        if (k < K && (u == 0 || S == 1)) cpart2[((size_t)blockIdx.x * S + part) * K + k] = acc[u];
    }
}

static inline size_t cgs_fused_smem(int K, int NB)
{
    return sizeof(double) * ((size_t)NB * K * CF_R + (size_t)NB * CF_R + CF_THREADS + ((K + 1) & ~1)) + 16;
}

// CTAs per SM and tile buffers per CTA for the fused kernel: as many CTAs as shared memory allows (up to 4)
// with double buffering for small K, two single-buffer CTAs for the rest
static inline void cgs_fused_shape(int K, int order, int *ctas, int *nb)
{
    const size_t budget = 214 * 1024;
    static const int cand[3][6][2] = {
        {{4, 2}, {2, 2}, {2, 1}, {1, 2}, {1, 2}, {1, 2}},
        {{4, 2}, {4, 1}, {3, 1}, {2, 1}, {1, 2}, {1, 2}},
        {{4, 2}, {3, 2}, {2, 2}, {3, 1}, {2, 1}, {1, 2}}};
    for (int i = 0; i < 6; ++i) {
        const int c = cand[order][i][0], b = cand[order][i][1];
        if (c * (cgs_fused_smem(K, b) + 1024) <= budget) { *ctas = c; *nb = b; return; }
    }
    *ctas = 1; *nb = 2;
}

// c[k] = sum over CTAs of cpart[cta][k]  (fixed order); optionally all K in one small launch
__global__ void __launch_bounds__(VT)
k_cgs_reduce(int K, int n_parts, const double *__restrict__ cpart, double *__restrict__ c,
             const int *__restrict__ flags, int need_flag, double *__restrict__ alpha_out /* alpha_j = c[K-1] (folded pass B) or NULL */)
{
    if (need_flag && flags[F_SECOND_SWEEP] == 0) return;
    __shared__ double red[32];
    const int k = blockIdx.x;
    double s = 0.0;
    for (int p = threadIdx.x; p < n_parts; p += VT) s += cpart[(size_t)p * K + k];
    s = lz_block_sum<VT>(s, red);
    if (threadIdx.x == 0) {
        c[k] = s;
        if (alpha_out && k == K - 1) *alpha_out = s;
    }
}

template <int RPT>
__global__ void __launch_bounds__(VT, 3)
k_cgs_update(int64_t n, int K, const double *__restrict__ V, int64_t ts, int64_t cs, double *__restrict__ w,
             const double *__restrict__ c, double *partials, unsigned int *ticket, const LzFinal fin,
             int *flags, int need_flag, int dgks_test, const double *nrm2_before, int skip_share)
{
    constexpr int H = CgsShape<RPT>::HALVES, U = CgsShape<RPT>::UNROLL, TILE = CgsShape<RPT>::TILE;
    extern __shared__ double csm[];           // c[K]
    __shared__ double red[32];
    if (need_flag && flags[F_SECOND_SWEEP] == 0) {
        // skipped sweep (selective reorthogonalisation).  Sharded runs issue their collectives unconditionally, so
        // the skipping kernel still takes part: peer mode adds a zero, NCCL mode re-publishes the norm pass B left
        // (rank 0 contributes it, the others zero) so that the all-reduce + finalisation that follows changes nothing
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            if (fin.pd) { double z = 0.0; lz_peer_sum_thread<1>(fin.pd, fin.seq, &z); }
            else if (skip_share >= 0) {
                // bit 0: this is rank 0 (it carries the value, the others zero); bit 1: the norm to keep is the one the previous
                // sweep left in nrm2_out (DGKS), otherwise the one pass B left (selective)
                const double keep = (skip_share & 2) ? *fin.nrm2_out : *nrm2_before;
                *fin.nrm2_out = (skip_share & 1) ? keep : 0.0;
            }
        }
        return;
    }
    for (int k = threadIdx.x; k < K; k += VT) csm[k] = c[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc = 0.0;
    const int64_t n_tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t base = tile * TILE + (int64_t)warp * (32 * RPT) + lane * 4;
        int64_t r[H]; const double *p[H]; double wv[H][4]; bool fl[H]; bool full = true;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            r[h] = base + 128 * h;
            // element (i, k) of the basis lives at V[(i >> 5) * ts + k * cs + (i & 31)]  (see lz_ctx_basis)
            p[h] = V + (r[h] >> 5) * ts + (r[h] & 31);
            fl[h] = r[h] + 3 < n;
            full = full && fl[h];
#pragma unroll
            for (int t = 0; t < 4; ++t) wv[h][t] = 0.0;
            if (fl[h]) lz_ld256(w + r[h], wv[h][0], wv[h][1], wv[h][2], wv[h][3]);
            else for (int t = 0; t < 4; ++t) if (r[h] + t < n) wv[h][t] = w[r[h] + t];
        }
        int k = 0;
        if (full) {
            for (; k + U <= K; k += U) {
                double v[U][H][4];
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int h = 0; h < H; ++h) lz_ld256_stream(p[h] + (int64_t)(k + u) * cs, v[u][h][0], v[u][h][1], v[u][h][2], v[u][h][3]);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double ck = -csm[k + u];
#pragma unroll
                    for (int h = 0; h < H; ++h)
#pragma unroll
                        for (int t = 0; t < 4; ++t) wv[h][t] = fma(ck, v[u][h][t], wv[h][t]);
                }
            }
        }
        for (; k < K; ++k) {
            const int64_t ko = (int64_t)k * cs;
            const double ck = -csm[k];
            for (int h = 0; h < H; ++h)
                for (int t = 0; t < 4; ++t)
                    if (r[h] + t < n) wv[h][t] = fma(ck, p[h][ko + t], wv[h][t]);
        }
#pragma unroll
        for (int h = 0; h < H; ++h) {
            if (fl[h]) lz_st256(w + r[h], wv[h][0], wv[h][1], wv[h][2], wv[h][3]);
            else for (int t = 0; t < 4; ++t) if (r[h] + t < n) w[r[h] + t] = wv[h][t];
#pragma unroll
            for (int t = 0; t < 4; ++t) acc = fma(wv[h][t], wv[h][t], acc);
        }
    }
    acc = lz_block_sum<VT>(acc, red);
    double total;
    if (lz_grid_sum<VT, 1>(&acc, partials, ticket, red, &total) && threadIdx.x == 0) {
        total = lz_finalize_beta(fin, total);
        // DGKS: a second sweep is needed only if this one removed a large part of w.  (NCCL-mode sharded runs hold only
        // the local share here: their test runs in the epilogue of the all-reduce that follows, k_ar_epilogue.)
        if (dgks_test && (fin.finalize || !fin.pd)) flags[F_SECOND_SWEEP] = (total < 0.5 * (*nrm2_before)) ? 1 : 0;
    }
}

// beta0 = ||b||: finalises beta[0], invb[0]
__global__ void k_finalize_first(const double *nrm2, double *beta, double *invb, int *flags)
{
    const double t = *nrm2, b = sqrt(t);
    beta[0] = b;
    invb[0] = 1.0 / b;
    flags[F_BREAKDOWN] = (!isfinite(t) || t == 0.0) ? 0 : 0x7fffffff;
    flags[F_SECOND_SWEEP] = 1;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static inline int64_t round_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// Sharded runs: a kernel whose last CTA finalises a norm either all-reduces it in place over peer memory (the
// sequence number is drawn HERE, immediately before the launch, so every rank issues its collectives in the same
// order) or -- NCCL mode -- only publishes the local total, and finish_norm() adds the all-reduce + finalisation.
static LzFinal arm_final(lz_ctx *ctx, const LzFinal &fin, bool sharded, bool want_norm)
{
    LzFinal f = fin;
    f.pd = nullptr; f.seq = 0;
    if (!want_norm) { f.finalize = 0; f.beta = nullptr; return f; }
    if (sharded) {
        if (lz_comm_peer(ctx)) { f.pd = lz_comm_desc(ctx); f.seq = lz_comm_next_seq(ctx); f.finalize = 1; }
        else f.finalize = 0;
    }
    return f;
}

static int finish_norm(lz_ctx *ctx, const LzFinal &f, bool sharded, bool want_norm, int dgks_test = 0)
{
    if (!sharded || !want_norm || lz_comm_peer(ctx)) return LZ_OK;
    LzArEpi e;
    memset(&e, 0, sizeof(e));
    e.beta = f.beta; e.invb = f.invb; e.flags = f.flags + F_BREAKDOWN; e.jn = f.jn;
    if (dgks_test) { e.dgks_flag = f.flags + F_SECOND_SWEEP; e.dgks_before = ctx->scalars + S_NRM2_BEFORE; }
    return lz_comm_allreduce_sum(ctx, f.nrm2_out, 1, &e);
}

// rows per thread of the streaming CGS kernels: 8, or 4 when that leaves fewer than ~8 tiles per CTA (see CgsShape)
static inline int cgs_rows_per_thread(const lz_ctx *ctx, int64_t n, unsigned max_grid)
{
    if (ctx->knobs.cgs_rpt == 4 || ctx->knobs.cgs_rpt == 8) return ctx->knobs.cgs_rpt;
    const int64_t tiles8 = (n + VT * 8 - 1) / (VT * 8);
    return tiles8 < (int64_t)max_grid * 8 ? 4 : 8;
}

static int launch_cgs_project(lz_ctx *ctx, const LzCgs &g, int64_t n, int K, const double *w, int need_flag)
{
    const size_t smem = sizeof(double) * (VT / 32) * (size_t)K;
    lz_prof_begin(ctx, LZ_K_PROJECT, 8.0 * (double)n * (K + 1));
    if (g.rpt == 8) k_cgs_project<8><<<g.grid, VT, smem, ctx->stream>>>(n, K, g.V, g.ts, g.cs, w, g.cpart, ctx->flags, need_flag);
    else k_cgs_project<4><<<g.grid, VT, smem, ctx->stream>>>(n, K, g.V, g.ts, g.cs, w, g.cpart, ctx->flags, need_flag);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return LZ_OK;
}

static int launch_cgs_update(lz_ctx *ctx, const LzCgs &g, int64_t n, int K, double *w, const LzFinal &fin, int need_flag, int dgks_test,
                             bool sharded, bool want_norm)
{
    // (a variant that gave each warp one tile and four adjacent columns per load -- fully contiguous 1 KB
    // reads of the tiled slab -- measured 4 % slower than this generic kernel: profiles/r01_cgs_fusion.md)
    const int mult = ctx->knobs.cgs_upd_mult;      // default 3 CTAs/SM: 5.9 -> 6.6 TB/s on the row-tiled basis
    const unsigned cap = (unsigned)(ctx->sm_count * mult);
    const int rpt = cgs_rows_per_thread(ctx, n, cap);
    const unsigned want = stream_grid(ctx, n, VT * rpt);
    const unsigned grid = want < cap ? want : cap;
    const LzFinal f = arm_final(ctx, fin, sharded, want_norm);
    const int share = (sharded && want_norm && !f.pd) ? ((lz_comm_rank(ctx) == 0 ? 1 : 0) | (ctx->vrun && ctx->vrun->reorth == LZ_REORTH_FULL_DGKS ? 2 : 0)) : -1;
    lz_prof_begin(ctx, LZ_K_UPDATE, 8.0 * (double)n * (K + 2));
    if (rpt == 8)
        k_cgs_update<8><<<grid, VT, sizeof(double) * (size_t)K, ctx->stream>>>(
            n, K, g.V, g.ts, g.cs, w, g.c, ctx->partials, ctx->tickets + T_UPD, f, ctx->flags, need_flag, dgks_test, ctx->scalars + S_NRM2_BEFORE, share);
    else
        k_cgs_update<4><<<grid, VT, sizeof(double) * (size_t)K, ctx->stream>>>(
            n, K, g.V, g.ts, g.cs, w, g.c, ctx->partials, ctx->tickets + T_UPD, f, ctx->flags, need_flag, dgks_test, ctx->scalars + S_NRM2_BEFORE, share);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    return finish_norm(ctx, f, sharded, want_norm, dgks_test);
}

// partial coefficients -> c (fixed order), all-reduced over the ranks when sharded; alpha_out (optional) receives
// c[K-1]: with pass B folded into the first sweep, c1[j] = q_j . w IS alpha_j
static int cgs_reduce(lz_ctx *ctx, const LzCgs &g, int K, int n_parts, int need_flag, bool sharded, double *alpha_out)
{
    k_cgs_reduce<<<K, VT, 0, ctx->stream>>>(K, n_parts, g.cpart, g.c, ctx->flags, need_flag, sharded ? nullptr : alpha_out);
    LZ_LAUNCH_CHECK(ctx);
    if (!sharded) return LZ_OK;
    LzArEpi e;
    memset(&e, 0, sizeof(e));
    e.copy_dst = alpha_out; e.copy_idx = K - 1;
    return lz_comm_allreduce_sum(ctx, g.c, (size_t)K, alpha_out ? &e : nullptr);
}

// one CGS sweep of w against the first K basis columns; the update's epilogue finalises beta[jn] when want_norm
static int cgs_sweep(lz_ctx *ctx, const LzCgs &g, int64_t n, int K, double *w, const LzFinal &fin,
                     int need_flag, int dgks_test, bool sharded, bool want_norm, double *alpha_out)
{
    LZ_TRY(launch_cgs_project(ctx, g, n, K, w, need_flag));
    LZ_TRY(cgs_reduce(ctx, g, K, (int)g.grid, need_flag, sharded, alpha_out));
    return launch_cgs_update(ctx, g, n, K, w, fin, need_flag, dgks_test, sharded, want_norm);
}

// CGS2 with the middle two basis streams fused: project, [update + project], update.
// Falls back to two plain sweeps when the tile does not fit in shared memory (large K) or the
// operands are not 16-byte aligned for the bulk copies.
static int cgs2_fused(lz_ctx *ctx, const LzCgs &g, int64_t n, int K, double *w, const LzFinal &fin, bool sharded, double *alpha_out)
{
    int ctas = 1, NB = 2;
    if (!ctx->knobs.cgs_one_cta) cgs_fused_shape(K, ctx->knobs.cgs_shape_order, &ctas, &NB);
    const size_t smem = cgs_fused_smem(K, NB);
    const bool ok = smem <= 220 * 1024 && K <= CF_MAXC * CF_THREADS && g.cs == CF_R && ((uintptr_t)w % 16 == 0) && !ctx->knobs.no_cgs_fuse &&
                    K >= ctx->knobs.cgs_fuse_min_k;     // below ~64 columns the fused kernel sits on its per-tile floor (0.79 us per 32 rows,
                                                        // profiles/r01_cgs_fusion.md) and two plain sweeps stream less time than it takes
    if (!ok) {
        LZ_TRY(cgs_sweep(ctx, g, n, K, w, fin, 0, 0, sharded, false, alpha_out));
        return cgs_sweep(ctx, g, n, K, w, fin, 0, 0, sharded, true, nullptr);
    }
    LZ_TRY(lz_func_smem_optin(ctx, (const void *)k_cgs_update_project, 220 * 1024, true));
    // sweep 1 projection
    LZ_TRY(launch_cgs_project(ctx, g, n, K, w, 0));
    LZ_TRY(cgs_reduce(ctx, g, K, (int)g.grid, 0, sharded, alpha_out));
    // sweep 1 update + sweep 2 projection, one basis stream
    const unsigned fgrid = (unsigned)(ctx->sm_count * ctas);
    lz_prof_begin(ctx, LZ_K_UPDPROJ, 8.0 * (double)n * (K + 2));
    const int S = ctx->knobs.cgs_no_slices ? 1 : cgs_fused_slices(K);
    k_cgs_update_project<<<fgrid, CF_THREADS, smem, ctx->stream>>>(n, K, g.V, g.ts, w, g.c, g.cpart, S, NB);
    LZ_LAUNCH_CHECK(ctx);
    lz_prof_end(ctx);
    LZ_TRY(cgs_reduce(ctx, g, K, (int)fgrid * S, 0, sharded, nullptr));
    // sweep 2 update (+ ||w||^2, beta finalisation)
    return launch_cgs_update(ctx, g, n, K, w, fin, 0, 0, sharded, true);
}

__global__ void k_copy_scalar(const double *src, double *dst) { *dst = *src; }

// ---------------------------------------------------------------------------------------------
// Selective (partial) reorthogonalisation: Simon's omega recurrence estimates  omega_{j+1,k} ~ q_{j+1} . q_k  from the
// scalars alone; w is reorthogonalised against the stored basis only when an estimate passes sqrt(eps), and once more
// on the following step.  The recurrence runs in one CTA on the device and leaves its decision in a flag the (then
// conditional) CGS kernels read, so the iteration still never synchronises with the host.
//   beta_{j+1} omega_{j+1,k} = beta_{k+1} omega_{j,k+1} + (alpha_k - alpha_j) omega_{j,k} + beta_k omega_{j,k-1}
//                              - beta_j omega_{j-1,k} + theta,   |theta| = eps (beta_{k+1} + beta_{j+1})
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_omega_step(int j, double psi, const double *__restrict__ alpha, const double *__restrict__ beta,
             const double *__restrict__ om_old, const double *__restrict__ om_cur, double *__restrict__ om_new, int *flags)
{
    __shared__ double red[32];
    __shared__ int need_s;
    const double eps = 2.220446049250313e-16;
    const double bj1 = beta[j + 1], aj = alpha[j], bj = beta[j];
    double mx = 0.0;
    for (int k = threadIdx.x; k < j; k += 256) {
        // om_cur[k] = omega_{j,k} (om_cur[j] = 1), om_old[k] = omega_{j-1,k} (om_old[j-1] = 1)
        double t = beta[k + 1] * om_cur[k + 1] + (alpha[k] - aj) * om_cur[k] - bj * om_old[k];
        if (k > 0) t += beta[k] * om_cur[k - 1];
        const double th = eps * (beta[k + 1] + bj1);
        t = (t + (t >= 0.0 ? th : -th)) / bj1;
        om_new[k] = t;
        mx = fmax(mx, fabs(t));
    }
    // block max through the sum helper's shared array
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_down_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m2 = 0.0;
        for (int wv = 0; wv < 8; ++wv) m2 = fmax(m2, red[wv]);
        const int forced = flags[F_FORCE];
        const int need = (m2 > 1.4901161193847656e-08 /* sqrt(eps) */ || forced || !isfinite(m2)) ? 1 : 0;
        flags[F_SECOND_SWEEP] = need;
        flags[F_FORCE] = (need && !forced) ? 1 : 0;          // a triggered step is followed by exactly one forced step
        if (need) flags[F_REORTH_COUNT] += 1;
        need_s = need;
        om_new[j] = psi;
        om_new[j + 1] = 1.0;
    }
    __syncthreads();
    if (need_s)
        for (int k = threadIdx.x; k <= j; k += 256) om_new[k] = eps;     // q_{j+1} is about to be made orthogonal to working precision
}

__global__ void k_omega_init(int len, double *om_old, double *om_cur, double *om_new, int *flags)
{
    for (int k = threadIdx.x; k < len; k += blockDim.x) { om_old[k] = 0.0; om_cur[k] = 0.0; om_new[k] = 0.0; }
    __syncthreads();
    if (threadIdx.x == 0) { om_cur[0] = 1.0; flags[F_FORCE] = 0; flags[F_REORTH_COUNT] = 0; flags[F_SECOND_SWEEP] = 0; }
}

// omega rows live behind alpha in the scalar bank: three arrays of m + 2
static inline double *omega_row(const LzVecRun &R, int which) { return R.alpha + R.m + (size_t)which * (R.m + 2); }

static int omega_init(lz_ctx *ctx, LzVecRun &R)
{
    R.om[0] = omega_row(R, 0); R.om[1] = omega_row(R, 1); R.om[2] = omega_row(R, 2);
    k_omega_init<<<1, 256, 0, ctx->stream>>>(R.m + 2, R.om[0], R.om[1], R.om[2], ctx->flags);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

static int omega_step(lz_ctx *ctx, LzVecRun &R, int j)
{
    const double n_glob = (double)(R.A->global_rows > 0 ? R.A->global_rows : R.n);
    const double psi = 2.220446049250313e-16 * sqrt(n_glob);       // q_{j+1} . q_j right after the three-term step
    k_omega_step<<<1, 256, 0, ctx->stream>>>(j, psi, R.alpha, R.beta, R.om[0], R.om[1], R.om[2], ctx->flags);
    LZ_LAUNCH_CHECK(ctx);
    double *t = R.om[0]; R.om[0] = R.om[1]; R.om[1] = R.om[2]; R.om[2] = t;      // (old, cur, new) <- (cur, new, old)
    return LZ_OK;
}

// The single-vector driver.  Device arrays: alpha[m], beta[m+1], invb[m+1].
// With a communicator attached to the context (lz_comm_init) the operator is this rank's row slab,
// b holds the local rows, every gather source carries [lower halo | local | upper halo] and every
// reduction is completed over the ranks before it is consumed; without one the same code runs the
// single-GPU path with no collective at all.
//
// Step j (full reorthogonalisation, the default "fold"): pass A  w = A q_j - beta_j q_{j-1} (+ basis column j),
// then CGS2 against columns 0..j.  The first sweep's coefficient c1[j] = q_j . w is alpha_j, so the reference's
// separate  w -= alpha_j q_j  pass (vector_lanczos.hpp:60) and its reduction are part of the sweep: three basis
// streams and -- sharded -- three small collectives per step (c1, c2, ||w||^2).
// Sharded, overlapped: the boundary planes of q_{j+1} leave for the neighbours on a side stream as soon as the last
// update of step j retires, the interior chunks of the next pass A (no halo column) run meanwhile, the two
// boundary chunk ranges follow once the halo has landed.
//
// The run is a persistent object of the context (LzVecRun): vec_begin sets it up, vec_steps advances it, so a solve
// can be continued, checkpointed (lz_vector_checkpoint_*) and thick-restarted (lz_eigs_thick_restart).
int lz_vec_setup(lz_ctx *ctx, const lz_matrix *A, int m, int64_t lc, int reorth, double *q)
{
    const int64_t n = A->n_rows;
    const bool sharded = ctx->comm != nullptr && lz_comm_world(ctx) > 1;
    const int64_t hlo = A->halo_lo, hhi = A->halo_hi;
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_vector_lanczos: the operator belongs to another (or a destroyed) context");
    LZ_CHECK(A->n_cols == n + hlo + hhi, LZ_ERR_INVALID, "lz_vector_lanczos: operator must be square (%lld x %lld)", (long long)n, (long long)A->n_cols);
    LZ_CHECK(sharded || (hlo == 0 && hhi == 0), LZ_ERR_INVALID, "lz_vector_lanczos: a sharded operator needs lz_comm_init");
    LZ_CHECK(lc >= -1 && lc < n, LZ_ERR_INVALID, "lz_vector_lanczos: lc %lld out of range", (long long)lc);
    LZ_CHECK(reorth >= LZ_REORTH_NONE && reorth <= LZ_REORTH_SELECTIVE, LZ_ERR_INVALID, "lz_vector_lanczos: reorth mode %d", reorth);
    LZ_CHECK(3 * m + 3 + 16 + (reorth == LZ_REORTH_SELECTIVE ? 3 * (m + 2) : 0) <= SB_VECTOR_LIMIT, LZ_ERR_UNSUPPORTED,
             "lz_vector_lanczos: m = %d exceeds the scalar bank", m);
    // three rotating work vectors (the reference's q0, q1, w: test_lanczos.cu:57-59), each laid out
    // [pad | lower halo | local rows | upper halo] with the local part 32-byte aligned.  Sharded: the layout is
    // made IDENTICAL on every rank (largest halo / span), so a neighbour's halo region is found by offset.
    int64_t off = round_up(hlo, 4), span = off + n + hhi, n_below = -1;
    if (sharded) {
        const int64_t mine[4] = {hlo, n, hhi, 0};
        int64_t all[4 * LZ_MAX_RANKS];
        LZ_TRY(lz_comm_gather4(ctx, mine, all));
        const int world = lz_comm_world(ctx), rank = lz_comm_rank(ctx);
        int64_t max_hlo = 0;
        for (int r = 0; r < world; ++r) max_hlo = all[4 * r] > max_hlo ? all[4 * r] : max_hlo;
        off = round_up(max_hlo, 4);
        span = 0;
        for (int r = 0; r < world; ++r) { const int64_t sp = off + all[4 * r + 1] + all[4 * r + 2]; span = sp > span ? sp : span; }
        n_below = rank > 0 ? all[4 * (rank - 1) + 1] : 0;
    }
    const int64_t stride = round_up(span, 4);
    LzCgs g = {nullptr, 0, 0, nullptr, nullptr, 0, 8};
    const int cgs_rpt = cgs_rows_per_thread(ctx, n, (unsigned)(ctx->sm_count * 2));
    const unsigned cgs_grid = stream_grid(ctx, n, VT * cgs_rpt) < (unsigned)(ctx->sm_count * 2)
                                  ? stream_grid(ctx, n, VT * cgs_rpt) : (unsigned)(ctx->sm_count * 2);
    size_t work_bytes = sizeof(double) * (size_t)stride * 3;
    // projection partials: one row of m per CTA; the fused kernel writes S * K <= CF_THREADS entries per CTA
    const size_t cpart_len = std::max((size_t)ctx->sm_count * 2 * m, (size_t)ctx->sm_count * 4 * CF_THREADS);
    if (reorth) work_bytes += sizeof(double) * (cpart_len + m + 8);
    void *work;
    if (sharded) LZ_TRY(lz_comm_arena(ctx, work_bytes, (size_t)m + 8, &work));
    else LZ_TRY(lz_ctx_workspace(ctx, work_bytes, &work));
    if (reorth) {
        LZ_CHECK(sizeof(double) * (VT / 32) * (size_t)m <= 200 * 1024, LZ_ERR_UNSUPPORTED,
                 "lz_vector_lanczos: m = %d too large for the projection kernel's shared memory", m);
        g.cpart = (double *)work + 3 * stride;
        g.c = g.cpart + cpart_len;
        g.grid = cgs_grid;
        g.rpt = cgs_rpt;
        LZ_TRY(lz_ctx_basis(ctx, n, m, &g.V));
        g.ts = ctx->basis_ts; g.cs = ctx->basis_cs;
        LZ_TRY(lz_func_smem_optin(ctx, (const void *)k_cgs_project<8>, 200 * 1024));
        LZ_TRY(lz_func_smem_optin(ctx, (const void *)k_cgs_project<4>, 200 * 1024));
    }
    if (!ctx->vrun) ctx->vrun = new LzVecRun();
    LzVecRun &R = *ctx->vrun;
    memset(&R, 0, sizeof(R));
    R.A = A; R.n = n; R.hlo = hlo; R.hhi = hhi; R.n_below = n_below; R.lc = lc; R.stride = stride;
    R.m = m; R.reorth = reorth; R.sharded = sharded;
    R.fold = reorth == LZ_REORTH_FULL && !ctx->knobs.no_fold;
    // interior chunks exist and the operator knows them: overlap the halo exchange with them
    R.overlap = sharded && !ctx->knobs.no_overlap && A->has_split && A->bnd_hi > A->bnd_lo && A->format == LZ_FMT_CSR && A->tma_ok && !A->vrowptr;
    R.u_prev = (double *)work + off; R.u_cur = R.u_prev + stride; R.w = R.u_cur + stride;
    R.g = g;
    R.beta = ctx->scalars + 16; R.invb = R.beta + (m + 1); R.alpha = R.invb + (m + 1);
    R.q = q;
    R.j = 0; R.first_next = 1;
    return LZ_OK;
}

// u_cur <- b ; beta[0] = ||b||  (vector_lanczos.hpp:21)
int lz_vec_start(lz_ctx *ctx, const double *b)
{
    LzVecRun &R = *ctx->vrun;
    double *sc = ctx->scalars;
    LZ_CUDA(cudaMemcpyAsync(R.u_cur, b, sizeof(double) * R.n, cudaMemcpyDeviceToDevice, ctx->stream));
    LZ_TRY(dot_async(ctx, R.n, b, b, sc + S_NRM2));
    if (R.sharded) LZ_TRY(lz_comm_allreduce_sum(ctx, sc + S_NRM2, 1));
    k_finalize_first<<<1, 1, 0, ctx->stream>>>(sc + S_NRM2, R.beta, R.invb, ctx->flags);
    LZ_LAUNCH_CHECK(ctx);
    if (R.reorth == LZ_REORTH_SELECTIVE) LZ_TRY(omega_init(ctx, R));
    R.j = 0; R.first_next = 1;
    return LZ_OK;
}

// steps [R.j, j_end)
int lz_vec_steps(lz_ctx *ctx, int j_end)
{
    LzVecRun &R = *ctx->vrun;
    LZ_CHECK(j_end <= R.m, LZ_ERR_INVALID, "lz_vector_lanczos: step %d beyond the capacity %d of this run", j_end, R.m);
    const lz_matrix *A = R.A;
    const int64_t n = R.n, hlo = R.hlo, hhi = R.hhi;
    const bool sharded = R.sharded, fold = R.fold, overlap = R.overlap;
    const bool peer = sharded && lz_comm_peer(ctx);
    const int reorth = R.reorth;
    const LzCgs &g = R.g;
    double *alpha = R.alpha, *beta = R.beta, *invb = R.invb, *sc = ctx->scalars;
    for (int j = R.j; j < j_end; ++j) {
        double *u_prev = R.u_prev, *u_cur = R.u_cur, *w = R.w;
        // neighbours' boundary planes of q_j (unnormalised, like everything in the rotating buffers)
        if (sharded) LZ_TRY(lz_comm_halo_exchange(ctx, u_cur, n, hlo, hhi, R.n_below, overlap));
        LzPassA pa;
        memset(&pa, 0, sizeof(pa));
        pa.x_own = u_cur; pa.u_prev = u_prev; pa.invb = invb; pa.beta = beta;
        pa.alpha_out = (sharded || fold) ? nullptr : alpha + j; pa.alpha_partial = sc + S_ALPHA_LOCAL;
        pa.vcol = reorth ? g.V + (size_t)j * g.cs : nullptr;
        pa.vts = g.ts;
        pa.qout = (R.q && R.lc >= 0) ? R.q + j : nullptr;
        pa.lc = R.lc; pa.j = j; pa.first = R.first_next;
        pa.partials = ctx->partials; pa.ticket = ctx->tickets + T_SPMV;
        lz_prof_begin(ctx, LZ_K_SPMV, 12.0 * (double)A->nnz + 28.0 * (double)n + (reorth ? 8.0 * (double)n : 0.0));
        if (overlap) {
            pa.alpha_hold = 1;
            LZ_TRY(lz_spmv_any<LZ_EPI_LANCZOS>(ctx, A, u_cur - hlo, w, pa, 1));       // interior chunks: no halo column
            LZ_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_halo, 0));
            pa.alpha_hold = 0; pa.alpha_accum = 1;
            if (peer && !fold) { pa.pd = lz_comm_desc(ctx); pa.seq = lz_comm_next_seq(ctx); pa.alpha_out = alpha + j; }
            LZ_TRY(lz_spmv_any<LZ_EPI_LANCZOS>(ctx, A, u_cur - hlo, w, pa, 2));       // the two boundary chunk ranges
        } else {
            if (peer && !fold) { pa.pd = lz_comm_desc(ctx); pa.seq = lz_comm_next_seq(ctx); pa.alpha_out = alpha + j; }
            LZ_TRY(lz_spmv_any<LZ_EPI_LANCZOS>(ctx, A, u_cur - hlo, w, pa));          // :51,:54,:57
        }
        lz_prof_end(ctx);
        if (sharded && !peer && !fold) {
            LzArEpi e;
            memset(&e, 0, sizeof(e));
            e.copy_dst = alpha + j; e.copy_idx = 0;
            LZ_TRY(lz_comm_allreduce_sum(ctx, sc + S_ALPHA_LOCAL, 1, &e));
        }
        const LzFinal fin = {beta, invb, sc + (reorth ? S_NRM2_BEFORE : S_NRM2), ctx->flags, j + 1, 1, nullptr, 0};
        if (!fold) {
            const LzFinal fb = arm_final(ctx, fin, sharded, true);
            lz_prof_begin(ctx, LZ_K_PASSB, 24.0 * (double)n);
            LZ_TRY(launch_pass_b(ctx, n, w, u_cur, alpha, invb, j, fb));              // :60,:44
            lz_prof_end(ctx);
            LZ_TRY(finish_norm(ctx, fb, sharded, true));
        }
        if (reorth) {
            const LzFinal f2 = {beta, invb, sc + S_NRM2, ctx->flags, j + 1, 1, nullptr, 0};
            if (reorth == LZ_REORTH_FULL) {
                LZ_TRY(cgs2_fused(ctx, g, n, j + 1, w, f2, sharded, fold ? alpha + j : nullptr));
            } else if (reorth == LZ_REORTH_FULL_DGKS) {
                LZ_TRY(cgs_sweep(ctx, g, n, j + 1, w, f2, 0, 1, sharded, true, nullptr));
                LZ_TRY(cgs_sweep(ctx, g, n, j + 1, w, f2, 1, 0, sharded, true, nullptr));
            } else {
                // selective: the omega recurrence decides on the device whether this step reorthogonalises; both
                // sweeps are then conditional on the flag it leaves (sharded runs keep their collectives unconditional)
                LZ_TRY(omega_step(ctx, R, j));
                LZ_TRY(cgs_sweep(ctx, g, n, j + 1, w, f2, 1, 0, sharded, false, nullptr));
                LZ_TRY(cgs_sweep(ctx, g, n, j + 1, w, f2, 1, 0, sharded, true, nullptr));
            }
        }
        R.u_prev = u_cur; R.u_cur = w; R.w = u_prev;                                  // :62 (pointer rotation, no copy)
        R.j = j + 1; R.first_next = 0;
    }
    k_copy_scalar<<<1, 1, 0, ctx->stream>>>(beta + R.j, sc + S_BETA_LAST);
    LZ_LAUNCH_CHECK(ctx);
    ctx->last_coupling_slot = S_BETA_LAST;
    return LZ_OK;
}

static int vector_lanczos_core(lz_ctx *ctx, const lz_matrix *A, const double *b, int m, int64_t lc, int reorth, double *q)
{
    LZ_TRY(lz_vec_setup(ctx, A, m, lc, reorth, q));
    LZ_TRY(lz_vec_start(ctx, b));
    return lz_vec_steps(ctx, m);
}

// columns j0 .. j0+ncols-1 of the stored basis -> column-major dst (leading dimension ldd)
__global__ void k_basis_copy(int64_t n, int j0, int ncols, const double *__restrict__ V, int64_t ts, int64_t cs, double *__restrict__ dst, int64_t ldd)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n * ncols) return;
    const int64_t i = e % n;
    const int c = (int)(e / n);
    dst[i + (int64_t)c * ldd] = V[(i >> 5) * ts + (int64_t)(j0 + c) * cs + (i & 31)];
}

extern "C" {

int lz_spmv(lz_ctx *ctx, const lz_matrix *A, const double *x, double *y)
{
    LZ_CHECK(ctx && A && x && y, LZ_ERR_INVALID, "lz_spmv: NULL argument");
    LZ_CHECK(x != y, LZ_ERR_INVALID, "lz_spmv: x and y must not alias");
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_spmv: the operator belongs to another (or a destroyed) context");
    LzPassA none;
    memset(&none, 0, sizeof(none));
    return lz_spmv_any<LZ_EPI_PLAIN>(ctx, A, x, y, none);
}

// The fdtd validator of the harness (methods/fdtd.hpp:6-31): nsteps explicit Euler steps
// u <- u + dt * A u, dt = t_end / nsteps, then u[lc].  The reference runs spmv + Vector::add per step
// (two kernels, three vector passes); here one fused SpMV pass per step between two work vectors.
int lz_fdtd_vector(lz_ctx *ctx, const lz_matrix *A, const double *u0, int64_t nsteps, double t_end, int64_t lc,
                   double *result_host, double *u_out)
{
    LZ_CHECK(ctx && A && u0 && nsteps >= 1, LZ_ERR_INVALID, "lz_fdtd_vector: bad arguments");
    const int64_t n = A->n_rows;
    LZ_CHECK(A->ctx == ctx, LZ_ERR_INVALID, "lz_fdtd_vector: the operator belongs to another (or a destroyed) context");
    LZ_CHECK(A->n_cols == n && A->halo_lo == 0 && A->halo_hi == 0, LZ_ERR_INVALID, "lz_fdtd_vector: operator must be square and unsharded");
    LZ_CHECK(lc >= -1 && lc < n && (lc >= 0) == (result_host != nullptr), LZ_ERR_INVALID, "lz_fdtd_vector: lc / result mismatch");
    LZ_CUDA(cudaSetDevice(ctx->device));
    const int64_t stride = round_up(n, 4);
    void *work;
    LZ_TRY(lz_ctx_workspace(ctx, sizeof(double) * (size_t)stride * 2, &work));
    double *a = (double *)work, *b = a + stride;
    LZ_CUDA(cudaMemcpyAsync(a, u0, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    LzPassA pa;
    memset(&pa, 0, sizeof(pa));
    pa.dt = t_end / (double)nsteps;
    for (int64_t i = 0; i < nsteps; ++i) {
        pa.x_own = a;
        LZ_TRY(lz_spmv_any<LZ_EPI_EULER>(ctx, A, a, b, pa));
        std::swap(a, b);
    }
    if (u_out) LZ_CUDA(cudaMemcpyAsync(u_out, a, sizeof(double) * n, cudaMemcpyDeviceToDevice, ctx->stream));
    if (result_host) LZ_CUDA(cudaMemcpyAsync(result_host, a + lc, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

int lz_dot(lz_ctx *ctx, int64_t n, const double *x, const double *y, double *result_host)
{
    LZ_CHECK(ctx && x && y && result_host && n > 0, LZ_ERR_INVALID, "lz_dot: bad arguments");
    LZ_TRY(dot_async(ctx, n, x, y, ctx->scalars + S_TMP));
    LZ_CUDA(cudaMemcpyAsync(result_host, ctx->scalars + S_TMP, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

int lz_nrm2(lz_ctx *ctx, int64_t n, const double *x, double *result_host)
{
    double s = 0.0;
    LZ_TRY(lz_dot(ctx, n, x, x, &s));
    if (!isfinite(s)) {
        lz_set_error("lz_nrm2: the norm is not finite");     // vector.hpp:239-241 aborts here
        return LZ_ERR_BREAKDOWN;
    }
    *result_host = sqrt(s);
    return LZ_OK;
}

int lz_axpby(lz_ctx *ctx, int64_t n, double a, double *y, double b, const double *x)
{
    LZ_CHECK(ctx && x && y && n > 0, LZ_ERR_INVALID, "lz_axpby: bad arguments");
    k_axpby<<<(unsigned)((n + VT - 1) / VT), VT, 0, ctx->stream>>>(n, a, y, b, x);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

int lz_vector_lanczos_async(lz_ctx *ctx, const lz_matrix *A, const double *b, int m, int64_t lc, int reorth,
                            double *alpha_dev, double *beta_dev, double *q)
{
    LZ_CHECK(ctx && A && b && alpha_dev && beta_dev && m >= 1, LZ_ERR_INVALID, "lz_vector_lanczos_async: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_TRY(vector_lanczos_core(ctx, A, b, m, lc, reorth, q));
    const LzVecRun &R = *ctx->vrun;
    LZ_CUDA(cudaMemcpyAsync(alpha_dev, R.alpha, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(beta_dev, R.beta, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
    return LZ_OK;
}

}  // extern "C"

// coefficients of the steps done so far to the host + the breakdown report (synchronises)
int lz_vec_report(lz_ctx *ctx, double *alpha_host, double *beta_host, int *steps_done)
{
    const LzVecRun &R = *ctx->vrun;
    int flag = 0;
    if (alpha_host && R.j) LZ_CUDA(cudaMemcpyAsync(alpha_host, R.alpha, sizeof(double) * R.j, cudaMemcpyDeviceToHost, ctx->stream));
    if (beta_host && R.j) LZ_CUDA(cudaMemcpyAsync(beta_host, R.beta, sizeof(double) * R.j, cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(&flag, ctx->flags + F_BREAKDOWN, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    // flag = first j whose beta_j is zero / non-finite: coefficients alpha[0..j-1], beta[0..j-1] are valid
    const int done = (flag < R.j) ? flag : R.j;
    if (steps_done) *steps_done = done;
    if (done < R.j) {
        lz_set_error("lz_vector_lanczos: breakdown, beta[%d] is zero or not finite", done);
        return LZ_ERR_BREAKDOWN;
    }
    return LZ_OK;
}

extern "C" {

int lz_vector_lanczos(lz_ctx *ctx, const lz_matrix *A, const double *b, int m, int64_t lc, int reorth,
                      double *alpha_host, double *beta_host, double *q, int *steps_done)
{
    LZ_CHECK(ctx && A && b && alpha_host && beta_host && m >= 1, LZ_ERR_INVALID, "lz_vector_lanczos: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_TRY(vector_lanczos_core(ctx, A, b, m, lc, reorth, q));
    return lz_vec_report(ctx, alpha_host, beta_host, steps_done);
}

// The same run in pieces: begin sets up a run of at most m_capacity steps (basis slab, work vectors, beta_0), advance
// performs `steps` more and reports all coefficients so far.  Between two advances the state can be saved
// (lz_vector_checkpoint_save) and, in another process / on another context, restored (lz_vector_checkpoint_load).
int lz_vector_lanczos_begin(lz_ctx *ctx, const lz_matrix *A, const double *b, int m_capacity, int64_t lc, int reorth, double *q)
{
    LZ_CHECK(ctx && A && b && m_capacity >= 1, LZ_ERR_INVALID, "lz_vector_lanczos_begin: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_TRY(lz_vec_setup(ctx, A, m_capacity, lc, reorth, q));
    return lz_vec_start(ctx, b);
}

// everything lz_vector_lanczos would allocate on its first call (work vectors, basis slab, scratch), ahead of a timed call
int lz_vector_lanczos_workspace(lz_ctx *ctx, const lz_matrix *A, int m, int reorth)
{
    LZ_CHECK(ctx && A && m >= 1, LZ_ERR_INVALID, "lz_vector_lanczos_workspace: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_TRY(lz_vec_setup(ctx, A, m, -1, reorth, nullptr));
    if (A->vrowptr) { void *p; LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * ((size_t)A->n_virtual + 64), &p)); }
    ctx->vrun->A = nullptr;       // sizes only: no run has begun
    return LZ_OK;
}

int lz_vector_lanczos_advance(lz_ctx *ctx, int steps, double *alpha_host, double *beta_host, int *steps_done)
{
    LZ_CHECK(ctx && ctx->vrun && ctx->vrun->A, LZ_ERR_INVALID, "lz_vector_lanczos_advance: no run has been begun on this context");
    LZ_CHECK(steps >= 0 && ctx->vrun->j + steps <= ctx->vrun->m, LZ_ERR_INVALID, "lz_vector_lanczos_advance: %d more steps exceed the capacity %d",
             steps, ctx->vrun->m);
    LZ_CHECK(ctx->vrun->A->ctx == ctx, LZ_ERR_INVALID, "lz_vector_lanczos_advance: the run's operator has been orphaned");
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_TRY(lz_vec_steps(ctx, ctx->vrun->j + steps));
    return lz_vec_report(ctx, alpha_host, beta_host, steps_done);
}

int lz_vector_lanczos_sharded(lz_ctx *ctx, const lz_matrix *A_local, const double *b_local, int m, int reorth,
                              double *alpha_dev, double *beta_dev)
{
    LZ_CHECK(ctx && A_local && b_local && alpha_dev && beta_dev && m >= 1, LZ_ERR_INVALID, "lz_vector_lanczos_sharded: bad arguments");
    LZ_CHECK(ctx->comm, LZ_ERR_COMM, "lz_vector_lanczos_sharded: call lz_comm_init first");
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_TRY(vector_lanczos_core(ctx, A_local, b_local, m, -1, reorth, nullptr));
    const LzVecRun &R = *ctx->vrun;
    LZ_CUDA(cudaMemcpyAsync(alpha_dev, R.alpha, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
    LZ_CUDA(cudaMemcpyAsync(beta_dev, R.beta, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
    return LZ_OK;
}

int lz_vector_basis_copy(lz_ctx *ctx, int j0, int ncols, double *dst, int64_t ldd)
{
    LZ_CHECK(ctx && ctx->basis && dst, LZ_ERR_INVALID, "lz_vector_basis_copy: no basis has been built on this context");
    LZ_CHECK(j0 >= 0 && ncols >= 1 && j0 + ncols <= ctx->basis_cols && ldd >= ctx->basis_rows, LZ_ERR_INVALID, "lz_vector_basis_copy: bad range");
    const int64_t total = ctx->basis_rows * ncols;
    k_basis_copy<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ctx->basis_rows, j0, ncols, ctx->basis, ctx->basis_ts, ctx->basis_cs, dst, ldd);
    LZ_LAUNCH_CHECK(ctx);
    return LZ_OK;
}

int lz_vector_basis_info(lz_ctx *ctx, int64_t *rows, int *cols)
{
    LZ_CHECK(ctx && ctx->basis, LZ_ERR_INVALID, "lz_vector_basis_info: no basis has been built on this context");
    if (rows) *rows = ctx->basis_rows;
    if (cols) *cols = ctx->basis_cols;
    return LZ_OK;
}

}  // extern "C"
