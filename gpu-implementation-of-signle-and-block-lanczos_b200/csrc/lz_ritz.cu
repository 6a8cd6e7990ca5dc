// lz_ritz.cu -- Ritz values and residual estimates from the projected matrix T.
//
// The reference only diagonalises T inside expm_cusolver (cusolverDnDsyevd, utils/lib_utils.hpp:
// 542-590, T assembled by objects/tridiagonal_matrix.hpp:90-127).  T is tiny ((m*bw)^2), so this is
// host arithmetic inside the library: Householder tridiagonalisation (block case) followed by the
// implicit-shift QL iteration, carrying only the rows of the eigenvector matrix that the residual
// estimate needs.  Residual of Ritz pair i:  || beta_m * Y[last block rows, i] ||.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "lz_common.cuh"

namespace {

// Symmetric tridiagonal QL with implicit shifts.  d[0..n): diagonal, e[0..n-1): sub-diagonal
// (e[i] couples i and i+1).  Z is an (nz x n) row bundle of the accumulated transformation
// (column j of the full eigenvector matrix restricted to nz selected rows), updated in place.
// Returns false if an eigenvalue fails to converge.
bool tridiag_ql(int n, std::vector<double> &d, std::vector<double> &e, int nz, std::vector<double> &Z)
{
    e.resize(n, 0.0);
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0;
        while (true) {
            int m = l;
            for (; m < n - 1; ++m) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) <= 2.3e-16 * dd) break;
            }
            if (m == l) break;
            if (++iter > 200) return false;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = hypot(g, 1.0);
            g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
            double s = 1.0, c = 1.0, p = 0.0;
            int i = m - 1;
            for (; i >= l; --i) {
                double f = s * e[i];
                const double b = c * e[i];
                r = hypot(f, g);
                e[i + 1] = r;
                if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                s = f / r;
                c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
                for (int k = 0; k < nz; ++k) {
                    double *z = &Z[(size_t)k * n];
                    f = z[i + 1];
                    z[i + 1] = s * z[i] + c * f;
                    z[i] = c * z[i] - s * f;
                }
            }
            if (r == 0.0 && i >= l) continue;
            d[l] -= p;
            e[l] = g;
            e[m] = 0.0;
        }
    }
    return true;
}

// Householder reduction of the dense symmetric A (n x n, column-major, destroyed) to tridiagonal
// form.  On return d/e hold the tridiagonal and Q (n x n, column-major) the orthogonal factor with
// A = Q T Q^T.
void householder_tridiag(int n, std::vector<double> &A, std::vector<double> &d, std::vector<double> &e, std::vector<double> &Q)
{
    Q.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) Q[i + (size_t)i * n] = 1.0;
    std::vector<double> v(n), p(n), w(n);
    for (int k = 0; k < n - 2; ++k) {
        // annihilate A[k+2.., k]
        double alpha = 0.0;
        for (int i = k + 1; i < n; ++i) alpha += A[i + (size_t)k * n] * A[i + (size_t)k * n];
        alpha = sqrt(alpha);
        if (alpha == 0.0) continue;
        if (A[k + 1 + (size_t)k * n] > 0.0) alpha = -alpha;
        for (int i = 0; i < n; ++i) v[i] = 0.0;
        v[k + 1] = A[k + 1 + (size_t)k * n] - alpha;
        for (int i = k + 2; i < n; ++i) v[i] = A[i + (size_t)k * n];
        double vn = 0.0;
        for (int i = k + 1; i < n; ++i) vn += v[i] * v[i];
        if (vn == 0.0) continue;
        const double beta = 2.0 / vn;
        // p = beta A v ; w = p - (beta/2)(p.v) v ; A -= v w^T + w v^T
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = k + 1; j < n; ++j) s += A[i + (size_t)j * n] * v[j];
            p[i] = beta * s;
        }
        double pv = 0.0;
        for (int i = k + 1; i < n; ++i) pv += p[i] * v[i];
        for (int i = 0; i < n; ++i) w[i] = p[i] - 0.5 * beta * pv * v[i];
        for (int j = 0; j < n; ++j)
            for (int i = 0; i < n; ++i) A[i + (size_t)j * n] -= v[i] * w[j] + w[i] * v[j];
        // Q <- Q (I - beta v v^T)
        for (int i = 0; i < n; ++i) {
            double s = 0.0;
            for (int j = k + 1; j < n; ++j) s += Q[i + (size_t)j * n] * v[j];
            s *= beta;
            for (int j = k + 1; j < n; ++j) Q[i + (size_t)j * n] -= s * v[j];
        }
    }
    d.resize(n);
    e.assign(n, 0.0);
    for (int i = 0; i < n; ++i) d[i] = A[i + (size_t)i * n];
    for (int i = 0; i + 1 < n; ++i) e[i] = A[i + 1 + (size_t)i * n];
}

}  // namespace

extern "C" int lz_ritz(int m, int bw, const double *alpha, const double *beta, const double *beta_last, int k,
                       double *theta, double *resid)
{
    LZ_CHECK(m >= 1 && bw >= 1 && alpha && beta && theta && k >= 1, LZ_ERR_INVALID, "lz_ritz: bad arguments");
    const int N = m * bw;
    LZ_CHECK(k <= N, LZ_ERR_INVALID, "lz_ritz: k = %d exceeds the dimension %d of T", k, N);
    LZ_CHECK(N <= 4096, LZ_ERR_UNSUPPORTED, "lz_ritz: T of dimension %d is too large for the host eigensolver", N);
    std::vector<double> d, e, Z;
    const int nz = bw;                      // rows of the eigenvector matrix we carry: the last block
    if (bw == 1) {
        d.assign(alpha, alpha + m);
        e.assign(m, 0.0);
        for (int i = 0; i + 1 < m; ++i) e[i] = beta[i + 1];
        Z.assign((size_t)N, 0.0);
        Z[N - 1] = 1.0;                     // e_last^T * I
    } else {
        // dense T exactly as Assemble_T lays it out (tridiagonal_matrix.hpp:13-54)
        std::vector<double> T((size_t)N * N, 0.0), Q;
        for (int blk = 0; blk < m; ++blk)
            for (int i = 0; i < bw * bw; ++i) {
                const int r = i % bw, c = i / bw;
                T[(blk * bw + r) + (size_t)(blk * bw + c) * N] = alpha[(size_t)blk * bw * bw + i];
                if (blk >= 1) {
                    const double v = beta[(size_t)blk * bw * bw + i];
                    T[((blk - 1) * bw + r) + (size_t)(blk * bw + c) * N] = v;
                    T[(blk * bw + c) + (size_t)((blk - 1) * bw + r) * N] = v;
                }
            }
        // syevd reads one triangle; symmetrise the diagonal blocks the same way (lower wins)
        for (int j = 0; j < N; ++j)
            for (int i = j + 1; i < N; ++i) T[j + (size_t)i * N] = T[i + (size_t)j * N];
        householder_tridiag(N, T, d, e, Q);
        Z.assign((size_t)nz * N, 0.0);
        for (int r = 0; r < nz; ++r)
            for (int j = 0; j < N; ++j) Z[(size_t)r * N + j] = Q[(N - bw + r) + (size_t)j * N];
    }
    if (!tridiag_ql(N, d, e, nz, Z)) {
        lz_set_error("lz_ritz: QL iteration did not converge");
        return LZ_ERR_BREAKDOWN;
    }
    std::vector<int> order(N);
    for (int i = 0; i < N; ++i) order[i] = i;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return d[a] < d[b]; });
    const int lo = k / 2;
    for (int t = 0; t < k; ++t) {
        const int idx = t < lo ? order[t] : order[N - (k - t)];
        theta[t] = d[idx];
        if (resid) {
            double r2 = 0.0;
            if (beta_last) {
                for (int rr = 0; rr < bw; ++rr) {      // (beta_last * y_last)[rr]
                    double s = 0.0;
                    for (int cc = 0; cc < bw; ++cc) s += beta_last[rr + cc * bw] * Z[(size_t)cc * N + idx];
                    r2 += s * s;
                }
            }
            resid[t] = sqrt(r2);
        }
    }
    return LZ_OK;
}

// ---------------------------------------------------------------------------------------------
// expm of a small symmetric matrix.  The reference forms expm(T) as
// V exp(Lambda) V^T from cusolverDnDsyevd (expm_cusolver, utils/lib_utils.hpp:542-590, and the
// custom_mult kernel, kernels/dense_kernels.hpp:53-78: out[r,c] = sum_i V[r,i] exp(w_i) V[c,i]).
// Same construction here, on the host: Householder tridiagonalisation, implicit QL with the full
// eigenvector matrix, then the triple product.  Only the lower triangle of T is read (uplo = LOWER).
// ---------------------------------------------------------------------------------------------
int lz_sym_eig_full(int N, std::vector<double> &A /* col-major, destroyed */, std::vector<double> &d, std::vector<double> &Zt)
{
    for (int j = 0; j < N; ++j)
        for (int i = j + 1; i < N; ++i) A[j + (size_t)i * N] = A[i + (size_t)j * N];
    std::vector<double> e, Q;
    if (N == 1) { d.assign(1, A[0]); Zt.assign(1, 1.0); return LZ_OK; }
    householder_tridiag(N, A, d, e, Q);
    // tridiag_ql carries rows of the accumulated transformation: row r of Zt = row r of Q, so that after
    // the iteration Zt[r * N + i] = component r of eigenvector i
    Zt.assign((size_t)N * N, 0.0);
    for (int r = 0; r < N; ++r)
        for (int j = 0; j < N; ++j) Zt[(size_t)r * N + j] = Q[r + (size_t)j * N];
    if (!tridiag_ql(N, d, e, N, Zt)) {
        lz_set_error("symmetric eigensolver: QL iteration did not converge");
        return LZ_ERR_BREAKDOWN;
    }
    return LZ_OK;
}

// same for a plain array (col-major N x N, destroyed)
int lz_sym_eig_full_c(int N, double *A, std::vector<double> &d, std::vector<double> &Zt)
{
    std::vector<double> a(A, A + (size_t)N * N);
    return lz_sym_eig_full(N, a, d, Zt);
}

extern "C" int lz_expm_sym(int n, double *T_host)
{
    LZ_CHECK(n >= 1 && T_host, LZ_ERR_INVALID, "lz_expm_sym: bad arguments");
    LZ_CHECK(n <= 2048, LZ_ERR_UNSUPPORTED, "lz_expm_sym: dimension %d is too large for the host eigensolver", n);
    std::vector<double> A(T_host, T_host + (size_t)n * n), d, Zt;
    LZ_TRY(lz_sym_eig_full(n, A, d, Zt));
    std::vector<double> ex(n);
    for (int i = 0; i < n; ++i) ex[i] = exp(d[i]);
    for (int c = 0; c < n; ++c)
        for (int r = 0; r < n; ++r) {
            double s = 0.0;
            const double *zr = &Zt[(size_t)r * n], *zc = &Zt[(size_t)c * n];
            for (int i = 0; i < n; ++i) s += zr[i] * ex[i] * zc[i];
            T_host[r + (size_t)c * n] = s;
        }
    return LZ_OK;
}

// solution = q^T expm(t_end T)[:, 0:bw] beta_0   (the harness post-processing, test_lanczos.cu:100-110 and
// :270-283): bw = 1 gives the scalar beta_0 * sum_j expm(t_end T)[j,0] q[j]
extern "C" int lz_lanczos_solution(int m, int bw, const double *alpha, const double *beta, const double *q, double t_end,
                                   double *solution)
{
    LZ_CHECK(m >= 1 && bw >= 1 && alpha && beta && q && solution, LZ_ERR_INVALID, "lz_lanczos_solution: bad arguments");
    const int N = m * bw;
    LZ_CHECK(N <= 2048, LZ_ERR_UNSUPPORTED, "lz_lanczos_solution: T of dimension %d is too large", N);
    std::vector<double> T((size_t)N * N, 0.0);
    if (bw == 1) {
        for (int i = 0; i < m; ++i) T[i + (size_t)i * N] = t_end * alpha[i];
        for (int i = 0; i + 1 < m; ++i) T[(i + 1) + (size_t)i * N] = T[i + (size_t)(i + 1) * N] = t_end * beta[i + 1];
    } else {
        for (int blk = 0; blk < m; ++blk)
            for (int i = 0; i < bw * bw; ++i) {
                const int r = i % bw, c = i / bw;
                T[(blk * bw + r) + (size_t)(blk * bw + c) * N] = t_end * alpha[(size_t)blk * bw * bw + i];
                if (blk >= 1) {
                    const double v = t_end * beta[(size_t)blk * bw * bw + i];
                    T[((blk - 1) * bw + r) + (size_t)(blk * bw + c) * N] = v;
                    T[(blk * bw + c) + (size_t)((blk - 1) * bw + r) * N] = v;
                }
            }
    }
    LZ_TRY(lz_expm_sym(N, T.data()));
    // F1 = E[:, 0:bw] * beta_0 ; solution = q^T F1
    for (int c = 0; c < bw; ++c) {
        double s = 0.0;
        for (int i = 0; i < N; ++i) {
            double f = 0.0;
            for (int k = 0; k < bw; ++k) f += T[i + (size_t)k * N] * (bw == 1 ? beta[0] : beta[k + (size_t)c * bw]);
            s += q[i] * f;
        }
        solution[c] = s;
    }
    return LZ_OK;
}
