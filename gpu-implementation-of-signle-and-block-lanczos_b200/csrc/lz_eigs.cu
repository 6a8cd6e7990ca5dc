// lz_eigs.cu -- what turns the Lanczos recurrence into a bounded-memory eigensolver:
//   * Ritz vectors X = V Y as a tall-skinny DMMA product over the row-tiled Krylov basis,
//   * thick-restart Lanczos (Wu & Simon): after m_max steps the basis is compressed IN PLACE to the wanted Ritz
//     vectors (converged ones stay locked in front), the residual vector becomes the next Lanczos vector and the
//     recurrence continues -- k extremal eigenpairs inside a fixed basis budget,
//   * save / restore of a run (q_{j-1}, q_j, alpha, beta, j, basis) between two lz_vector_lanczos_advance calls.
// The reference has none of this: it diagonalises T once inside expm_cusolver (utils/lib_utils.hpp:542-590,
// objects/tridiagonal_matrix.hpp:90-127) and keeps no basis.  Host arithmetic here is confined to the small projected
// matrix (at most m_max x m_max); everything of length n stays on the device.
#include <math.h>
#include <stdio.h>

#include <algorithm>
#include <vector>

#include "lz_dense.cuh"

int lz_sym_eig_full(int N, std::vector<double> &A, std::vector<double> &d, std::vector<double> &Zt);   // lz_ritz.cu

// ---------------------------------------------------------------------------------------------
// out[:, col0 .. col0 + 8 NT) = V[:, 0..m) * Y[:, col0 ..)        (V row-tiled: element (i,k) at V[(i>>5)*ts + k*32 + (i&31)])
// One warp owns 8 rows of a 32-row tile and ALL 8*NT output columns of the pass: the accumulators are the C fragments
// of NT mma.m8n8k4 tiles, the A fragments come straight from the tile (4 k x 8 rows = four 64-byte segments per load),
// the B fragments from a 32-row chunk of Y staged in shared memory.  Because a warp reads every input column of its
// rows before it writes any output column of the same rows, the product may overwrite V itself (thick restart).
// ---------------------------------------------------------------------------------------------
#define ROT_THREADS 128
#define ROT_KC 32

template <int NT>
__global__ void __launch_bounds__(ROT_THREADS)
k_basis_rotate(int64_t n_rows, int m, int kcols, int col0, const double *V, int64_t ts, const double *__restrict__ Y, int ldy,
               double *out, int64_t ts_out, int64_t cs_out)
{
    extern __shared__ double ys[];                      // [ROT_KC][8 * NT + 4]
    constexpr int LDS = 8 * NT + 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kk = lane & 3, mm = lane >> 2;
    const int64_t n_tiles = (n_rows + 31) / 32;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        double acc[NT][2];
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = 0.0;
        const double *vt = V + tile * ts + 8 * warp + mm;
        for (int k0 = 0; k0 < m; k0 += ROT_KC) {
            __syncthreads();
            for (int e = threadIdx.x; e < ROT_KC * 8 * NT; e += ROT_THREADS) {
                const int kr = e % ROT_KC, c = e / ROT_KC;          // consecutive threads walk down a column of Y
                const int k = k0 + kr, col = col0 + c;
                ys[kr * LDS + c] = (k < m && col < kcols) ? Y[k + (size_t)col * ldy] : 0.0;
            }
            __syncthreads();
#pragma unroll 2
            for (int kt = 0; kt < ROT_KC / 4; ++kt) {
                const int k = k0 + kt * 4 + kk;
                const double a = k < m ? vt[(int64_t)k * 32] : 0.0;
#pragma unroll
                for (int t = 0; t < NT; ++t) lz_dmma(acc[t][0], acc[t][1], a, ys[(kt * 4 + kk) * LDS + t * 8 + mm]);
            }
        }
        const int64_t row = tile * 32 + 8 * warp + mm;
#pragma unroll
        for (int t = 0; t < NT; ++t)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = col0 + t * 8 + 2 * kk + e;
                if (col < kcols && row < n_rows) out[(row >> 5) * ts_out + (int64_t)col * cs_out + (row & 31)] = acc[t][e];
            }
    }
}

// out (layout ts_out / cs_out) = V[:, 0..m) Y[:, 0..kcols);  Y device, column-major m x kcols.  in_place: out == V
static int basis_rotate(lz_ctx *ctx, int64_t n, int m, int kcols, const double *V, int64_t ts, const double *Y, double *out,
                        int64_t ts_out, int64_t cs_out, bool in_place)
{
    const int64_t n_tiles = (n + 31) / 32;
    int grid = (int)std::min<int64_t>(n_tiles, (int64_t)ctx->sm_count * 8);
    if (grid < 1) grid = 1;
    lz_prof_begin(ctx, LZ_K_PANEL, 8.0 * (double)n * (m + kcols));
#define ROT_LAUNCH(NTV, c0)                                                                                               \
    do {                                                                                                                  \
        const size_t smem = sizeof(double) * ROT_KC * (8 * NTV + 4);                                                      \
        LZ_TRY(lz_func_smem_optin(ctx, (const void *)k_basis_rotate<NTV>, (int)smem));                                    \
        k_basis_rotate<NTV><<<grid, ROT_THREADS, smem, ctx->stream>>>(n, m, kcols, c0, V, ts, Y, m, out, ts_out, cs_out); \
        LZ_LAUNCH_CHECK(ctx);                                                                                             \
    } while (0)
    if (in_place) {
        // one pass must cover every output column (a second pass would read columns the first one overwrote)
        if (kcols <= 64) ROT_LAUNCH(8, 0);
        else if (kcols <= 128) ROT_LAUNCH(16, 0);
        else if (kcols <= 192) ROT_LAUNCH(24, 0);
        else if (kcols <= 256) ROT_LAUNCH(32, 0);
        else { lz_set_error("basis_rotate: %d columns kept in place (at most 256)", kcols); return LZ_ERR_UNSUPPORTED; }
    } else {
        for (int c0 = 0; c0 < kcols; c0 += 64) ROT_LAUNCH(8, c0);
    }
#undef ROT_LAUNCH
    lz_prof_end(ctx);
    return LZ_OK;
}

// basis column `col` <-> a plain vector of n doubles
__global__ void k_basis_col_get(int64_t n, int col, const double *__restrict__ V, int64_t ts, int64_t cs, double *__restrict__ x)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = V[(i >> 5) * ts + (int64_t)col * cs + (i & 31)];
}
__global__ void k_basis_col_set(int64_t n, int col, double *__restrict__ V, int64_t ts, int64_t cs, const double *__restrict__ x)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) V[(i >> 5) * ts + (int64_t)col * cs + (i & 31)] = x[i];
}
__global__ void k_set_restart_scalars(double *beta, double *invb, int from, int to)
{
    beta[to] = beta[from];
    invb[to] = invb[from];
}

extern "C" {

// X[:, 0..k) = V_j Y   with V_j the first j = steps-done columns of the stored basis of the current run (full or
// selective reorthogonalisation) and Y_host a j x k column-major matrix of host doubles (eigenvectors of T, as
// lz_ritz / the caller's own solver produce them).  X: device, column-major, leading dimension ldx >= rows.
int lz_vector_ritz_vectors(lz_ctx *ctx, int k, const double *Y_host, double *X, int64_t ldx)
{
    LZ_CHECK(ctx && ctx->vrun && ctx->vrun->A && Y_host && X && k >= 1, LZ_ERR_INVALID, "lz_vector_ritz_vectors: bad arguments / no run");
    const LzVecRun &R = *ctx->vrun;
    LZ_CHECK(R.reorth != LZ_REORTH_NONE && R.g.V && R.j >= 1, LZ_ERR_INVALID, "lz_vector_ritz_vectors: the run keeps no basis (reorth = none)");
    LZ_CHECK(ldx >= R.n, LZ_ERR_INVALID, "lz_vector_ritz_vectors: ldx too small");
    LZ_CUDA(cudaSetDevice(ctx->device));
    void *yd;
    LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)R.j * k, &yd));
    LZ_CUDA(cudaMemcpyAsync(yd, Y_host, sizeof(double) * (size_t)R.j * k, cudaMemcpyHostToDevice, ctx->stream));
    // column-major X with leading dimension ldx is the tiled formula with ts = 32, cs = ldx
    return basis_rotate(ctx, R.n, R.j, k, R.g.V, R.g.ts, (const double *)yd, X, 32, ldx, false);
}

// ---------------------------------------------------------------------------------------------
// checkpoint of a run between two advances
// ---------------------------------------------------------------------------------------------
struct LzCkptHeader {
    char magic[8];
    int64_t n, global_rows, row_begin, lc;
    int32_t m, j, reorth, first_next, flag_breakdown, flag_force, flag_count, reserved;
};

int lz_vector_checkpoint_save(lz_ctx *ctx, const char *path)
{
    LZ_CHECK(ctx && path && ctx->vrun && ctx->vrun->A, LZ_ERR_INVALID, "lz_vector_checkpoint_save: no run on this context");
    const LzVecRun &R = *ctx->vrun;
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    FILE *f = fopen(path, "wb");
    LZ_CHECK(f, LZ_ERR_INVALID, "lz_vector_checkpoint_save: cannot open %s", path);
    LzCkptHeader h;
    memset(&h, 0, sizeof(h));
    memcpy(h.magic, "LZVCKPT1", 8);
    h.n = R.n; h.global_rows = R.A->global_rows; h.row_begin = R.A->row_begin; h.lc = R.lc;
    h.m = R.m; h.j = R.j; h.reorth = R.reorth; h.first_next = R.first_next;
    int flags[8];
    cudaMemcpy(flags, ctx->flags, sizeof(flags), cudaMemcpyDeviceToHost);
    h.flag_breakdown = flags[0]; h.flag_force = flags[4]; h.flag_count = flags[5];
    bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
    // scalar series: alpha[m], beta[m+1], invb[m+1] (+ the three omega rows of a selective run)
    const size_t n_sc = (size_t)3 * R.m + 2 + (R.reorth == LZ_REORTH_SELECTIVE ? 3 * (size_t)(R.m + 2) : 0);
    std::vector<double> sc(n_sc), vec((size_t)R.n);
    cudaMemcpy(sc.data(), R.beta, sizeof(double) * n_sc, cudaMemcpyDeviceToHost);      // beta, invb, alpha, omega are contiguous
    ok = ok && fwrite(sc.data(), sizeof(double), n_sc, f) == n_sc;
    int om_order[3] = {0, 1, 2};
    if (R.reorth == LZ_REORTH_SELECTIVE)
        for (int i = 0; i < 3; ++i) om_order[i] = (int)((R.om[i] - (R.alpha + R.m)) / (R.m + 2));
    ok = ok && fwrite(om_order, sizeof(int), 3, f) == 3;
    for (const double *v : {R.u_prev, R.u_cur}) {
        cudaMemcpy(vec.data(), v, sizeof(double) * R.n, cudaMemcpyDeviceToHost);
        ok = ok && fwrite(vec.data(), sizeof(double), (size_t)R.n, f) == (size_t)R.n;
    }
    if (R.reorth != LZ_REORTH_NONE) {
        void *tmp;
        LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * (size_t)R.n, &tmp));
        for (int c = 0; c < R.j && ok; ++c) {
            k_basis_col_get<<<(unsigned)((R.n + 255) / 256), 256, 0, ctx->stream>>>(R.n, c, R.g.V, R.g.ts, R.g.cs, (double *)tmp);
            cudaMemcpyAsync(vec.data(), tmp, sizeof(double) * R.n, cudaMemcpyDeviceToHost, ctx->stream);
            cudaStreamSynchronize(ctx->stream);
            ok = fwrite(vec.data(), sizeof(double), (size_t)R.n, f) == (size_t)R.n;
        }
    }
    ok = (fclose(f) == 0) && ok && cudaGetLastError() == cudaSuccess;
    LZ_CHECK(ok, LZ_ERR_INVALID, "lz_vector_checkpoint_save: write to %s failed", path);
    return LZ_OK;
}

// Recreates the saved run on this context for the operator A (which must be the operator the run was made with:
// rows and position are checked).  q (optional, device, m entries) resumes the receiver-row series.
int lz_vector_checkpoint_load(lz_ctx *ctx, const lz_matrix *A, const char *path, double *q)
{
    LZ_CHECK(ctx && A && path, LZ_ERR_INVALID, "lz_vector_checkpoint_load: bad arguments");
    LZ_CUDA(cudaSetDevice(ctx->device));
    FILE *f = fopen(path, "rb");
    LZ_CHECK(f, LZ_ERR_INVALID, "lz_vector_checkpoint_load: cannot open %s", path);
    LzCkptHeader h;
    bool ok = fread(&h, sizeof(h), 1, f) == 1 && memcmp(h.magic, "LZVCKPT1", 8) == 0;
    if (!ok || h.n != A->n_rows || h.global_rows != A->global_rows || h.row_begin != A->row_begin || h.m < 1 || h.j < 0 || h.j > h.m) {
        fclose(f);
        lz_set_error("lz_vector_checkpoint_load: %s is not a checkpoint of this operator", path);
        return LZ_ERR_INVALID;
    }
    int st = lz_vec_setup(ctx, A, h.m, h.lc, h.reorth, q);
    if (st != LZ_OK) { fclose(f); return st; }
    LzVecRun &R = *ctx->vrun;
    const size_t n_sc = (size_t)3 * R.m + 2 + (R.reorth == LZ_REORTH_SELECTIVE ? 3 * (size_t)(R.m + 2) : 0);
    std::vector<double> sc(n_sc), vec((size_t)R.n);
    ok = fread(sc.data(), sizeof(double), n_sc, f) == n_sc;
    int om_order[3] = {0, 1, 2};
    ok = ok && fread(om_order, sizeof(int), 3, f) == 3;
    if (ok) cudaMemcpy(R.beta, sc.data(), sizeof(double) * n_sc, cudaMemcpyHostToDevice);
    if (R.reorth == LZ_REORTH_SELECTIVE)
        for (int i = 0; i < 3; ++i) R.om[i] = R.alpha + R.m + (size_t)om_order[i] * (R.m + 2);
    for (double *v : {R.u_prev, R.u_cur}) {
        ok = ok && fread(vec.data(), sizeof(double), (size_t)R.n, f) == (size_t)R.n;
        if (ok) cudaMemcpy(v, vec.data(), sizeof(double) * R.n, cudaMemcpyHostToDevice);
    }
    if (R.reorth != LZ_REORTH_NONE) {
        void *tmp;
        st = lz_ctx_scratch(ctx, sizeof(double) * (size_t)R.n, &tmp);
        if (st != LZ_OK) { fclose(f); return st; }
        for (int c = 0; c < h.j && ok; ++c) {
            ok = fread(vec.data(), sizeof(double), (size_t)R.n, f) == (size_t)R.n;
            cudaMemcpyAsync(tmp, vec.data(), sizeof(double) * R.n, cudaMemcpyHostToDevice, ctx->stream);
            k_basis_col_set<<<(unsigned)((R.n + 255) / 256), 256, 0, ctx->stream>>>(R.n, c, R.g.V, R.g.ts, R.g.cs, (const double *)tmp);
            cudaStreamSynchronize(ctx->stream);
        }
    }
    fclose(f);
    int flags[8];
    cudaMemcpy(flags, ctx->flags, sizeof(flags), cudaMemcpyDeviceToHost);
    flags[0] = h.flag_breakdown; flags[1] = 1; flags[4] = h.flag_force; flags[5] = h.flag_count;
    cudaMemcpy(ctx->flags, flags, sizeof(flags), cudaMemcpyHostToDevice);
    R.j = h.j; R.first_next = h.first_next;
    LZ_CHECK(ok && cudaGetLastError() == cudaSuccess, LZ_ERR_INVALID, "lz_vector_checkpoint_load: %s is truncated", path);
    return LZ_OK;
}

// steps of the current selective run that reorthogonalised (synchronises)
int lz_vector_reorth_count(lz_ctx *ctx, int *count)
{
    LZ_CHECK(ctx && count, LZ_ERR_INVALID, "lz_vector_reorth_count: bad arguments");
    LZ_CUDA(cudaMemcpyAsync(count, ctx->flags + 5, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// Thick-restart Lanczos: k extremal eigenpairs of the symmetric operator A inside a basis of m_max vectors.
//   which: 0 smallest, 1 largest, 2 both ends (k/2 smallest, k - k/2 largest -- the convention of lz_ritz)
//   tol  : a pair counts as converged when |beta_m y_m| <= tol * max_i |theta_i|
// Every step is the fused pass A + CGS2 of the full-reorthogonalisation driver, whose first-sweep coefficients ARE
// the projected matrix: after a restart the coupling row s_i = beta_m y_{m,i} to the kept Ritz vectors is removed by
// the same sweep that delivers alpha, so the restarted recurrence needs no special kernel.
// ---------------------------------------------------------------------------------------------
int lz_eigs_thick_restart(lz_ctx *ctx, const lz_matrix *A, const double *b, int k, int which, int m_max, double tol,
                          int max_restarts, double *theta_host, double *resid_host, double *X, int64_t ldx, int *info4)
{
    LZ_CHECK(ctx && A && b && theta_host && k >= 1 && which >= 0 && which <= 2 && tol > 0.0 && max_restarts >= 0, LZ_ERR_INVALID,
             "lz_eigs_thick_restart: bad arguments");
    LZ_CHECK(m_max >= k + 4 && m_max <= 1024, LZ_ERR_INVALID, "lz_eigs_thick_restart: basis budget m_max = %d must be in [k + 4, 1024]", m_max);
    LZ_CHECK(!X || ldx >= A->n_rows, LZ_ERR_INVALID, "lz_eigs_thick_restart: ldx too small");
    LZ_CHECK(!ctx->knobs.no_fold, LZ_ERR_UNSUPPORTED, "lz_eigs_thick_restart: not available with LZ_NO_FOLD");
    LZ_CUDA(cudaSetDevice(ctx->device));
    LZ_TRY(lz_vec_setup(ctx, A, m_max, -1, LZ_REORTH_FULL, nullptr));
    LZ_TRY(lz_vec_start(ctx, b));
    LzVecRun &R = *ctx->vrun;
    const int m = m_max;
    std::vector<double> H((size_t)m * m, 0.0), alpha(m), beta(m + 1), d, Zt, Hw;
    std::vector<int> order(m), keep;
    int j0 = 0, restarts = 0, matvecs = 0, nconv = 0;
    std::vector<int> wanted;
    double beta_m = 0.0;
    for (;;) {
        LZ_TRY(lz_vec_steps(ctx, m));
        matvecs += m - j0;
        int flag = 0;
        LZ_CUDA(cudaMemcpyAsync(alpha.data(), R.alpha, sizeof(double) * m, cudaMemcpyDeviceToHost, ctx->stream));
        LZ_CUDA(cudaMemcpyAsync(beta.data(), R.beta, sizeof(double) * (m + 1), cudaMemcpyDeviceToHost, ctx->stream));
        LZ_CUDA(cudaMemcpyAsync(&flag, ctx->flags, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));
        if (flag <= m) {
            lz_set_error("lz_eigs_thick_restart: breakdown at step %d (invariant subspace reached or non-finite data)", flag);
            return LZ_ERR_BREAKDOWN;
        }
        for (int j = j0; j < m; ++j) {
            H[j + (size_t)j * m] = alpha[j];
            if (j > j0) H[(j - 1) + (size_t)j * m] = H[j + (size_t)(j - 1) * m] = beta[j];
        }
        beta_m = beta[m];
        Hw = H;
        LZ_TRY(lz_sym_eig_full(m, Hw, d, Zt));            // Zt[r * m + i] = component r of eigenvector i
        for (int i = 0; i < m; ++i) order[i] = i;
        std::sort(order.begin(), order.end(), [&](int x, int y) { return d[x] < d[y]; });
        double scale = 0.0;
        for (int i = 0; i < m; ++i) scale = std::max(scale, fabs(d[i]));
        const int lo = which == 0 ? k : which == 1 ? 0 : k / 2, hi = k - lo;       // wanted from the low / high end
        wanted.clear();
        for (int t = 0; t < lo; ++t) wanted.push_back(order[t]);
        for (int t = 0; t < hi; ++t) wanted.push_back(order[m - hi + t]);
        nconv = 0;
        for (int idx : wanted)
            if (fabs(beta_m * Zt[(size_t)(m - 1) * m + idx]) <= tol * scale) ++nconv;
        if (nconv == k || restarts == max_restarts) break;
        // keep the wanted pairs plus a share of their neighbours (same ends): faster convergence than keeping k alone
        int extra = std::max(1, (m - k) * 2 / 5);
        extra = std::min(extra, m - k - 3);
        const int elo = which == 0 ? extra : which == 1 ? 0 : extra / 2, ehi = extra - elo;
        keep.clear();
        for (int t = 0; t < lo + elo; ++t) keep.push_back(order[t]);
        for (int t = 0; t < hi + ehi; ++t) keep.push_back(order[m - (hi + ehi) + t]);
        const int kk = (int)keep.size();
        // V[:, 0..kk) <- V Y_keep (in place), then q_m becomes Lanczos vector number kk
        std::vector<double> Y((size_t)m * kk);
        for (int c = 0; c < kk; ++c)
            for (int r = 0; r < m; ++r) Y[r + (size_t)c * m] = Zt[(size_t)r * m + keep[c]];
        void *yd;
        LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * Y.size(), &yd));
        LZ_CUDA(cudaMemcpyAsync(yd, Y.data(), sizeof(double) * Y.size(), cudaMemcpyHostToDevice, ctx->stream));
        LZ_TRY(basis_rotate(ctx, R.n, m, kk, R.g.V, R.g.ts, (const double *)yd, R.g.V, R.g.ts, R.g.cs, true));
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));            // Y (host vector) and the scratch are reused next round
        k_set_restart_scalars<<<1, 1, 0, ctx->stream>>>(R.beta, R.invb, m, kk);
        LZ_LAUNCH_CHECK(ctx);
        std::fill(H.begin(), H.end(), 0.0);
        for (int c = 0; c < kk; ++c) {
            H[c + (size_t)c * m] = d[keep[c]];
            const double s = beta_m * Zt[(size_t)(m - 1) * m + keep[c]];
            H[c + (size_t)kk * m] = H[kk + (size_t)c * m] = s;
        }
        j0 = kk;
        R.j = kk; R.first_next = 1;
        ++restarts;
    }
    // report the wanted pairs in ascending order
    std::sort(wanted.begin(), wanted.end(), [&](int x, int y) { return d[x] < d[y]; });
    for (int t = 0; t < k; ++t) {
        theta_host[t] = d[wanted[t]];
        if (resid_host) resid_host[t] = fabs(beta_m * Zt[(size_t)(m - 1) * m + wanted[t]]);
    }
    if (X) {
        std::vector<double> Y((size_t)m * k);
        for (int c = 0; c < k; ++c)
            for (int r = 0; r < m; ++r) Y[r + (size_t)c * m] = Zt[(size_t)r * m + wanted[c]];
        void *yd;
        LZ_TRY(lz_ctx_scratch(ctx, sizeof(double) * Y.size(), &yd));
        LZ_CUDA(cudaMemcpyAsync(yd, Y.data(), sizeof(double) * Y.size(), cudaMemcpyHostToDevice, ctx->stream));
        LZ_TRY(basis_rotate(ctx, R.n, m, k, R.g.V, R.g.ts, (const double *)yd, X, 32, ldx, false));
        LZ_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if (info4) { info4[0] = nconv; info4[1] = restarts; info4[2] = matvecs; info4[3] = m; }
    return LZ_OK;
}

}  // extern "C"
