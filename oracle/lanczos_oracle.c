/*
 * lanczos_oracle.c -- CPU restatement of the reference's Lanczos hot path (plain C + OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY (see lanczos_oracle.h).  Never linked into, imported by, or
 * executed from the product path; the product has no CPU fallback.
 *
 * Citations are relative to /root/reference/source/.  With orc_set_threads(1) every
 * reduction runs strictly left-to-right, which is the order of the reference's Host
 * container loops; oracle/_ref/ref_host_dump (the reference's own headers) pins that.
 * Build with -ffp-contract=off so that a*x + b*y is two roundings, as in the reference
 * Host loops compiled without FMA contraction.
 */
#include "lanczos_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 1;

void orc_set_threads(int t)
{
    if (t < 1) t = 1;
    g_threads = t;
#ifdef _OPENMP
    omp_set_num_threads(t);
#endif
}
int orc_get_threads(void) { return g_threads; }

/* ------------------------------------------------------------------ generators -------- */

uint64_t orc_splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

static inline double u01(uint64_t x) { return (double)(x >> 11) * (1.0 / 9007199254740992.0); }

void orc_start_vector(int64_t n, uint64_t seed, double *v)
{
#pragma omp parallel for if (g_threads > 1)
    for (int64_t i = 0; i < n; ++i) v[i] = 2.0 * u01(orc_splitmix64(seed ^ (uint64_t)i)) - 1.0;
}

void orc_start_block(int64_t n, int b, int64_t ld, uint64_t seed, double *V)
{
#pragma omp parallel for if (g_threads > 1)
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < b; ++c)
            V[i + (int64_t)c * ld] = 2.0 * u01(orc_splitmix64(seed ^ (uint64_t)(i * b + c))) - 1.0;
}

int64_t orc_lap2d_nnz(int64_t nx, int64_t ny) { return 5 * nx * ny - 2 * nx - 2 * ny; }

/* 5-point Dirichlet Laplacian, row = y*nx + x, columns ascending */
void orc_lap2d_csr(int64_t nx, int64_t ny, int32_t *rowptr, int32_t *colidx, double *vals)
{
    int64_t p = 0;
    for (int64_t y = 0; y < ny; ++y)
        for (int64_t x = 0; x < nx; ++x) {
            int64_t i = y * nx + x;
            rowptr[i] = (int32_t)p;
            if (y > 0)      { colidx[p] = (int32_t)(i - nx); vals[p++] = -1.0; }
            if (x > 0)      { colidx[p] = (int32_t)(i - 1);  vals[p++] = -1.0; }
            colidx[p] = (int32_t)i; vals[p++] = 4.0;
            if (x < nx - 1) { colidx[p] = (int32_t)(i + 1);  vals[p++] = -1.0; }
            if (y < ny - 1) { colidx[p] = (int32_t)(i + nx); vals[p++] = -1.0; }
        }
    rowptr[nx * ny] = (int32_t)p;
}

int64_t orc_lap3d_nnz(int64_t nx, int64_t ny, int64_t nz)
{
    return 7 * nx * ny * nz - 2 * (nx * ny + ny * nz + nx * nz);
}

void orc_lap3d_csr(int64_t nx, int64_t ny, int64_t nz, int32_t *rowptr, int32_t *colidx, double *vals)
{
    int64_t p = 0, sxy = nx * ny;
    for (int64_t z = 0; z < nz; ++z)
        for (int64_t y = 0; y < ny; ++y)
            for (int64_t x = 0; x < nx; ++x) {
                int64_t i = z * sxy + y * nx + x;
                rowptr[i] = (int32_t)p;
                if (z > 0)      { colidx[p] = (int32_t)(i - sxy); vals[p++] = -1.0; }
                if (y > 0)      { colidx[p] = (int32_t)(i - nx);  vals[p++] = -1.0; }
                if (x > 0)      { colidx[p] = (int32_t)(i - 1);   vals[p++] = -1.0; }
                colidx[p] = (int32_t)i; vals[p++] = 6.0;
                if (x < nx - 1) { colidx[p] = (int32_t)(i + 1);   vals[p++] = -1.0; }
                if (y < ny - 1) { colidx[p] = (int32_t)(i + nx);  vals[p++] = -1.0; }
                if (z < nz - 1) { colidx[p] = (int32_t)(i + sxy); vals[p++] = -1.0; }
            }
    rowptr[nx * ny * nz] = (int32_t)p;
}

/* R-MAT (a,b,c,d) = (0.57,0.19,0.19,0.05); one counter-based draw per (edge, level). */
void orc_rmat_edges(int scale, int64_t n_edges, uint64_t seed, int32_t *src, int32_t *dst)
{
#pragma omp parallel for if (g_threads > 1)
    for (int64_t e = 0; e < n_edges; ++e) {
        uint32_t r = 0, c = 0;
        for (int l = 0; l < scale; ++l) {
            double u = u01(orc_splitmix64(seed ^ ((uint64_t)e * 64ULL + (uint64_t)l)));
            int q = (u < 0.57) ? 0 : (u < 0.76) ? 1 : (u < 0.95) ? 2 : 3;
            r = (r << 1) | (uint32_t)(q >> 1);
            c = (c << 1) | (uint32_t)(q & 1);
        }
        src[e] = (int32_t)r;
        dst[e] = (int32_t)c;
    }
}

/* ------------------------------------------------------------------ formats ----------- */

int64_t orc_ell_to_csr(int64_t n, int width, const double *ell_data, const uint32_t *ell_idx,
                       int32_t *rowptr, int32_t *colidx, double *vals)
{
    int64_t p = 0;
    for (int64_t r = 0; r < n; ++r) {
        rowptr[r] = (int32_t)p;
        for (int k = 0; k < width; ++k) {
            double v = ell_data[r + (int64_t)k * n];
            if (v != 0.0) { colidx[p] = (int32_t)ell_idx[r + (int64_t)k * n]; vals[p++] = v; }
        }
    }
    rowptr[n] = (int32_t)p;
    return p;
}

/* ------------------------------------------------------------------ operators --------- */

void orc_csr_spmv(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
                  const double *x, double *y)
{
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t i = 0; i < n; ++i) {
        double s = 0.0;
        for (int32_t p = rowptr[i]; p < rowptr[i + 1]; ++p) s += vals[p] * x[colidx[p]];
        y[i] = s;
    }
}

void orc_csr_spmm(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
                  int b, const double *X, int64_t ldx, double *Y, int64_t ldy)
{
    for (int c = 0; c < b; ++c) orc_csr_spmv(n, rowptr, colidx, vals, X + (int64_t)c * ldx, Y + (int64_t)c * ldy);
}

double orc_dot(int64_t n, const double *x, const double *y)
{
    if (g_threads > 1) {
        double s = 0.0;
#pragma omp parallel for reduction(+ : s) schedule(static)
        for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
        return s;
    }
    double s = 0.0;
    for (int64_t i = 0; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* y = a*y + b*x : objects/vector.hpp:228-231 */
static void axpby(int64_t n, double a, double *y, double b, const double *x)
{
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t i = 0; i < n; ++i) y[i] = a * y[i] + b * x[i];
}

/* ------------------------------------------------------------------ dense helpers ----- */

void orc_mm_tt(int64_t n, int b, const double *T, int64_t ld, double *R)
{
    for (int j = 0; j < b; ++j)
        for (int i = 0; i < b; ++i) R[i + j * b] = orc_dot(n, T + (int64_t)i * ld, T + (int64_t)j * ld);
}

void orc_mm_tt2(int64_t n, int b, const double *T1, int64_t ld1, const double *T2, int64_t ld2, double *R)
{
    for (int j = 0; j < b; ++j)
        for (int i = 0; i < b; ++i) {
            double a = orc_dot(n, T1 + (int64_t)i * ld1, T2 + (int64_t)j * ld2);
            double c = orc_dot(n, T2 + (int64_t)i * ld2, T1 + (int64_t)j * ld1);
            R[i + j * b] = 0.5 * a + 0.5 * c;
        }
}

void orc_mm_ts(int64_t n, int b, double beta, double alpha, const double *T, int64_t ldt,
               const double *S, double *R, int64_t ldr)
{
    /* R may alias T (the reference calls mm_cublas(0,1,F1,beta0,F1)), so go row by row */
#pragma omp parallel for schedule(static) if (g_threads > 1)
    for (int64_t i = 0; i < n; ++i) {
        double row[64], out[64];
        for (int k = 0; k < b; ++k) row[k] = T[i + (int64_t)k * ldt];
        for (int j = 0; j < b; ++j) {
            double s = 0.0;
            for (int k = 0; k < b; ++k) s += row[k] * S[k + j * b];
            out[j] = s;
        }
        for (int j = 0; j < b; ++j) {
            double r = (beta == 0.0) ? 0.0 : beta * R[i + (int64_t)j * ldr];
            R[i + (int64_t)j * ldr] = r + alpha * out[j];
        }
    }
}

int orc_jacobi_eig(int n, double *A, double *w, double *V)
{
    for (int i = 0; i < n * n; ++i) V[i] = 0.0;
    for (int i = 0; i < n; ++i) V[i + i * n] = 1.0;
    int sweep;
    for (sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) {
                if (i != j) off += A[i + j * n] * A[i + j * n];
                else diag += A[i + j * n] * A[i + j * n];
            }
        if (off <= 1e-60 * diag || off == 0.0) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double apq = A[p + q * n];
                if (apq == 0.0) continue;
                double app = A[p + p * n], aqq = A[q + q * n];
                double tau = (aqq - app) / (2.0 * apq);
                double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                double c = 1.0 / sqrt(1.0 + t * t), s = t * c;
                for (int k = 0; k < n; ++k) {
                    double akp = A[k + p * n], akq = A[k + q * n];
                    A[k + p * n] = c * akp - s * akq;
                    A[k + q * n] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    double apk = A[p + k * n], aqk = A[q + k * n];
                    A[p + k * n] = c * apk - s * aqk;
                    A[q + k * n] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    double vkp = V[k + p * n], vkq = V[k + q * n];
                    V[k + p * n] = c * vkp - s * vkq;
                    V[k + q * n] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; ++i) w[i] = A[i + i * n];
    /* sort ascending (syevj returns sorted eigenvalues: utils/lib_utils.hpp:772) */
    for (int i = 0; i < n - 1; ++i) {
        int k = i;
        for (int j = i + 1; j < n; ++j) if (w[j] < w[k]) k = j;
        if (k != i) {
            double t = w[i]; w[i] = w[k]; w[k] = t;
            for (int r = 0; r < n; ++r) { double u = V[r + i * n]; V[r + i * n] = V[r + k * n]; V[r + k * n] = u; }
        }
    }
    return sweep;
}

void orc_sqrtm(int b, double *S, double *Sinv)
{
    double *A = (double *)malloc(sizeof(double) * b * b);
    double *V = (double *)malloc(sizeof(double) * b * b);
    double *w = (double *)malloc(sizeof(double) * b);
    /* syevj reads the lower triangle only (CUBLAS_FILL_MODE_LOWER, lib_utils.hpp:703) */
    for (int j = 0; j < b; ++j)
        for (int i = 0; i < b; ++i) A[i + j * b] = (i >= j) ? S[i + j * b] : S[j + i * b];
    orc_jacobi_eig(b, A, w, V);
    for (int col = 0; col < b; ++col)
        for (int row = 0; row < b; ++row) {
            double s1 = 0.0, s2 = 0.0;
            for (int i = 0; i < b; ++i) {          /* custom_mult2, lib_utils.hpp:673-686 */
                double rt = sqrt(fabs(w[i]));
                s1 += V[row + i * b] * rt * V[col + i * b];
                s2 += V[row + i * b] * 1.0 / rt * V[col + i * b];
            }
            S[row + col * b] = s1;
            Sinv[row + col * b] = s2;
        }
    free(A); free(V); free(w);
}

/* ------------------------------------------------------------------ drivers ----------- */

/* one classical Gram-Schmidt sweep of w against the first k columns of V (n x k, ld = n) */
static void cgs_sweep(int64_t n, int k, const double *V, double *w, double *c)
{
    for (int j = 0; j < k; ++j) c[j] = orc_dot(n, V + (int64_t)j * n, w);
    for (int j = 0; j < k; ++j) axpby(n, 1.0, w, -c[j], V + (int64_t)j * n);
}

int orc_vector_lanczos(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
                       const double *b, int m, int64_t lc, int reorth,
                       double *alpha, double *beta, double *q, double *Vout)
{
    double *q0 = (double *)malloc(sizeof(double) * n);
    double *q1 = (double *)malloc(sizeof(double) * n);
    double *w = (double *)malloc(sizeof(double) * n);
    double *V = Vout, *c = NULL;
    int ownV = 0;
    if (reorth && !V) { V = (double *)malloc(sizeof(double) * n * (size_t)m); ownV = 1; }
    if (reorth) c = (double *)malloc(sizeof(double) * m);
    memcpy(q0, b, sizeof(double) * n);

    int done = 0;
    beta[0] = sqrt(orc_dot(n, b, b));                         /* :21 */
    axpby(n, 0.0, q0, 1.0 / beta[0], q0);                     /* :24 */
    q[0] = q0[lc];                                            /* :27 */
    if (V) memcpy(V, q0, sizeof(double) * n);
    orc_csr_spmv(n, rowptr, colidx, vals, q0, w);             /* :30 */
    alpha[0] = orc_dot(n, w, q0);                             /* :33 */
    axpby(n, 1.0, w, -alpha[0], q0);                          /* :36 */
    if (reorth) { cgs_sweep(n, 1, V, w, c); cgs_sweep(n, 1, V, w, c); }
    done = 1;
    for (int j = 1; j < m; ++j) {                             /* :39-66 */
        double nrm2 = orc_dot(n, w, w);
        if (!isfinite(nrm2) || nrm2 == 0.0) break;            /* vector.hpp:233-244 aborts */
        beta[j] = sqrt(nrm2);                                 /* :44 */
        memcpy(q1, w, sizeof(double) * n);                    /* :47 */
        axpby(n, 0.0, q1, 1.0 / beta[j], q1);                 /* :48 */
        orc_csr_spmv(n, rowptr, colidx, vals, q1, w);         /* :51 */
        axpby(n, 1.0, w, -beta[j], q0);                       /* :54 */
        alpha[j] = orc_dot(n, w, q1);                         /* :57 */
        axpby(n, 1.0, w, -alpha[j], q1);                      /* :60 */
        memcpy(q0, q1, sizeof(double) * n);                   /* :62 */
        q[j] = q0[lc];                                        /* :65 */
        if (V) memcpy(V + (int64_t)j * n, q0, sizeof(double) * n);
        if (reorth) { cgs_sweep(n, j + 1, V, w, c); cgs_sweep(n, j + 1, V, w, c); }
        done = j + 1;
    }
    free(q0); free(q1); free(w); free(c);
    if (ownV) free(V);
    return done;
}

/* block CGS sweep: W -= Vk (Vk^T W), Vk = first kc columns of V (n x kc, ld n) */
static void block_cgs_sweep(int64_t n, int kc, int bw, const double *V, double *W, double *C)
{
    for (int j = 0; j < bw; ++j)
        for (int i = 0; i < kc; ++i) C[i + (int64_t)j * kc] = orc_dot(n, V + (int64_t)i * n, W + (int64_t)j * n);
    for (int j = 0; j < bw; ++j)
        for (int i = 0; i < kc; ++i) axpby(n, 1.0, W + (int64_t)j * n, -C[i + (int64_t)j * kc], V + (int64_t)i * n);
}

int orc_block_lanczos(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
                      const double *B, int bw, int m, int64_t lc, int reorth,
                      double *alpha, double *beta, double *q, double *Vout)
{
    const size_t pan = (size_t)n * bw, bb = (size_t)bw * bw;
    double *Q0 = (double *)malloc(sizeof(double) * pan);
    double *Q1 = (double *)malloc(sizeof(double) * pan);
    double *W = (double *)malloc(sizeof(double) * pan);
    double *V = Vout, *C = NULL;
    int ownV = 0;
    if (reorth && !V) { V = (double *)malloc(sizeof(double) * pan * m); ownV = 1; }
    if (reorth) C = (double *)malloc(sizeof(double) * bb * m);
    double *binv = beta + bb * m;                                            /* beta[m] scratch */

    orc_mm_tt(n, bw, B, n, beta);                                            /* :106 */
    orc_sqrtm(bw, beta, binv);                                               /* :111 */
    orc_mm_ts(n, bw, 0.0, 1.0, B, n, binv, Q0, n);                           /* :114 */
    for (int c = 0; c < bw; ++c) q[c] = Q0[lc + (int64_t)c * n];             /* :117 */
    if (V) memcpy(V, Q0, sizeof(double) * pan);
    orc_csr_spmm(n, rowptr, colidx, vals, bw, Q0, n, W, n);                  /* :121 */
    orc_mm_tt2(n, bw, W, n, Q0, n, alpha);                                   /* :124 */
    orc_mm_ts(n, bw, 1.0, -1.0, Q0, n, alpha, W, n);                         /* :128 */
    if (reorth) { block_cgs_sweep(n, bw, bw, V, W, C); block_cgs_sweep(n, bw, bw, V, W, C); }
    int done = 1;
    for (int j = 1; j < m; ++j) {                                            /* :132-166 */
        double *bj = beta + bb * j, *aj = alpha + bb * j;
        orc_mm_tt(n, bw, W, n, bj);                                          /* :137 */
        orc_sqrtm(bw, bj, binv);                                             /* :142 */
        orc_mm_ts(n, bw, 0.0, 1.0, W, n, binv, Q1, n);                       /* :145 */
        orc_csr_spmm(n, rowptr, colidx, vals, bw, Q1, n, W, n);              /* :149 */
        orc_mm_ts(n, bw, 1.0, -1.0, Q0, n, bj, W, n);                        /* :152 */
        orc_mm_tt2(n, bw, W, n, Q1, n, aj);                                  /* :155 */
        orc_mm_ts(n, bw, 1.0, -1.0, Q1, n, aj, W, n);                        /* :159 */
        memcpy(Q0, Q1, sizeof(double) * pan);                                /* :162 */
        for (int c = 0; c < bw; ++c) q[(size_t)j * bw + c] = Q0[lc + (int64_t)c * n];   /* :165 */
        if (V) memcpy(V + pan * j, Q0, sizeof(double) * pan);
        if (reorth) {
            block_cgs_sweep(n, (j + 1) * bw, bw, V, W, C);
            block_cgs_sweep(n, (j + 1) * bw, bw, V, W, C);
        }
        done = j + 1;
    }
    free(Q0); free(Q1); free(W); free(C);
    if (ownV) free(V);
    return done;
}

void orc_assemble_T(int m, int bw, const double *alpha, const double *beta, double *T)
{
    const int N = m * bw;
    memset(T, 0, sizeof(double) * (size_t)N * N);
    for (int blk = 0; blk < m; ++blk)
        for (int i = 0; i < bw * bw; ++i) {
            int r = i % bw, c = i / bw;
            T[(blk * bw + r) + (size_t)(blk * bw + c) * N] = alpha[(size_t)blk * bw * bw + i];
            if (blk >= 1) {
                double v = beta[(size_t)blk * bw * bw + i];
                T[((blk - 1) * bw + r) + (size_t)(blk * bw + c) * N] = v;     /* upper block (b-1,b) */
                T[(blk * bw + c) + (size_t)((blk - 1) * bw + r) * N] = v;     /* mirrored transpose */
            }
        }
}

/* ---- application path -------------------------------------------------------- */

void orc_expm_sym(int n, double *T)
{
    double *A = (double *)malloc(sizeof(double) * n * n), *V = (double *)malloc(sizeof(double) * n * n);
    double *w = (double *)malloc(sizeof(double) * n);
    /* syevd with uplo = LOWER reads the lower triangle only (lib_utils.hpp:556) */
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i) A[i + j * n] = i >= j ? T[i + j * n] : T[j + i * n];
    orc_jacobi_eig(n, A, w, V);
    for (int col = 0; col < n; ++col)
        for (int row = 0; row < n; ++row) {
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += V[row + i * n] * exp(w[i]) * V[col + i * n];   /* dense_kernels.hpp:64-69 */
            T[row + col * n] = s;
        }
    free(A); free(V); free(w);
}

void orc_lanczos_solution(int m, int bw, const double *alpha, const double *beta, const double *q, double t_end,
                          double *solution)
{
    const int N = m * bw;
    double *T = (double *)calloc((size_t)N * N, sizeof(double));
    orc_assemble_T(m, bw, alpha, beta, T);                                  /* test_lanczos.cu:270 */
    for (int i = 0; i < N * N; ++i) T[i] *= t_end;                          /* :271 */
    orc_expm_sym(N, T);                                                     /* :272 */
    for (int c = 0; c < bw; ++c) {                                          /* :275-283 */
        double s = 0.0;
        for (int i = 0; i < N; ++i) {
            double f = 0.0;
            for (int k = 0; k < bw; ++k) f += T[i + (size_t)k * N] * beta[k + c * bw];
            s += q[i] * f;
        }
        solution[c] = s;
    }
    free(T);
}

void orc_fdtd_vector(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals, double *u,
                     int64_t nsteps, double t_end)
{
    const double dt = t_end / (double)nsteps;
    double *dudt = (double *)malloc(sizeof(double) * n);
    for (int64_t s = 0; s < nsteps; ++s) {
        orc_csr_spmv(n, rowptr, colidx, vals, u, dudt);                     /* fdtd.hpp:21 */
        for (int64_t i = 0; i < n; ++i) u[i] = u[i] + dt * dudt[i];         /* :22 */
    }
    free(dudt);
}

void orc_fdtd_block(int64_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals, int b, double *U,
                    int64_t ld, int64_t nsteps, double t_end)
{
    const double dt = t_end / (double)nsteps;
    double *dU = (double *)malloc(sizeof(double) * ld * b);
    for (int64_t s = 0; s < nsteps; ++s) {
        orc_csr_spmm(n, rowptr, colidx, vals, b, U, ld, dU, ld);            /* fdtd.hpp:47 */
        for (int c = 0; c < b; ++c)
            for (int64_t i = 0; i < n; ++i) U[i + c * ld] = U[i + c * ld] + dt * dU[i + c * ld];   /* :48 */
    }
    free(dU);
}
