/*
 * ref_host_dump.cpp -- golden-vector minting tool.  TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the REFERENCE's own Host code (objects/*.hpp containers, matrix_a/ builder, the
 * copy helpers) from where it lies under /root/reference/source with g++ -DDISABLE_CUDA and
 * runs the three-term recurrences of methods/vector_lanczos.hpp:20-66 and
 * methods/block_lanczos.hpp:105-166 over those containers.  The driver files themselves
 * cannot be included under g++ (they pull in kernels/spmv_spmm.hpp, which is unguarded CUDA),
 * so the ~30 driver lines are re-issued here call for call against the reference's Host
 * branches: Vector::l2_norm/dot/mult_scalar/add/operator=, Ell_matrix::spmv/spmm,
 * Dense_matrix::mm/tra.  The b x b matrix square root (cuSOLVER syevjBatched + custom_mult2 in
 * the reference, utils/lib_utils.hpp:650-745) has no Host branch; it is taken from
 * lanczos_oracle.c (orc_sqrtm) and therefore only pinned mathematically.
 *
 * Inputs follow test_lanczos.cu exactly: main() consumes one rand() for lc (:326) before
 * random_vector_b / random_matrix_B draw theirs (glibc rand(), never seeded), the matrix is
 * D.mult_diagonal(W) (:43,:191) and is left in column-major ELL order (the Host change_order(4)
 * is defective -- SURVEY.md appendix A-1 -- and Ell_matrix::spmv expects column-major anyway).
 *
 * usage: ref_host_dump <vector|block|matrix> N m out.bin
 * Output container: repeated records {u32 name_len, name, u8 dtype(0=f64,1=u32,2=i64), u64 count, data}.
 */
#ifndef N_COL
#define N_COL 4
#endif
#ifndef DISABLE_CUDA
#define DISABLE_CUDA
#endif

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <string>
#include <tuple>
#include <vector>

#include "utils/common.hpp"
#include "objects/ell_matrix.hpp"
#include "methods/copy_functions.hpp"
#include "matrix_a/build_A_ell.hpp"

#include "lanczos_oracle.h"

static FILE *g_out = nullptr;

static void put(const char *name, int dtype, uint64_t count, const void *data, size_t elt)
{
    uint32_t len = (uint32_t)std::strlen(name);
    uint8_t dt = (uint8_t)dtype;
    std::fwrite(&len, 4, 1, g_out);
    std::fwrite(name, 1, len, g_out);
    std::fwrite(&dt, 1, 1, g_out);
    std::fwrite(&count, 8, 1, g_out);
    std::fwrite(data, elt, count, g_out);
}
static void put_f64(const char *name, uint64_t count, const double *d) { put(name, 0, count, d, 8); }
static void put_u32(const char *name, uint64_t count, const unsigned int *d) { put(name, 1, count, d, 4); }
static void put_i64(const char *name, int64_t v) { put(name, 2, 1, &v, 8); }

typedef double T;

static void dump_matrix(Ell_matrix<T> &A)
{
    put_i64("n_rows", (int64_t)A.n_rows());
    put_i64("n_cols", (int64_t)A.n_cols());
    put_i64("width", (int64_t)A.width());
    put_f64("ell_data", A.size(), A.data());
    put_u32("ell_idx", A.size(), A.idx());
}

int main(int argc, char **argv)
{
    if (argc < 5) {
        std::fprintf(stderr, "usage: %s <vector|block|matrix> N m out.bin\n", argv[0]);
        return 2;
    }
    const std::string mode = argv[1];
    const unsigned int N = (unsigned int)std::atoi(argv[2]);
    const unsigned int m = (unsigned int)std::atoi(argv[3]);
    g_out = std::fopen(argv[4], "wb");
    if (!g_out) return 3;

    const unsigned int lc = 1 + (rand() % 100);          /* test_lanczos.cu:326 */

    auto info = Matrix_A<T>(N, N, N);                    /* test_lanczos.cu:29,142 */
    Ell_matrix<T> A = info.first;
    Ell_matrix<T> Wd = info.second;
    const unsigned int n = (unsigned int)A.n_rows();
    put_i64("N", N);
    put_i64("m", m);
    put_i64("lc", lc);
    put_i64("n_col", N_COL);
    if (mode == "matrix") {
        /* D and W separately, then A = D*W */
        put_f64("D_data", A.size(), A.data());
        put_u32("D_idx", A.size(), A.idx());
        put_f64("W_data", Wd.size(), Wd.data());
        put_u32("W_idx", Wd.size(), Wd.idx());
        put_i64("W_width", (int64_t)Wd.width());
        A.mult_diagonal(Wd);
        dump_matrix(A);
        std::fclose(g_out);
        return 0;
    }
    A.mult_diagonal(Wd);                                 /* test_lanczos.cu:43,191 */
    dump_matrix(A);

    if (mode == "vector") {
        Vector<T> b = random_vector_b<T>(n);             /* the non-degenerate start vector */
        put_f64("b", n, b.data());
        Vector<T> q(m, MemorySpace::Host), q0(b), q1(b), w(b);
        std::vector<T> alpha(m), beta(m);
        /* methods/vector_lanczos.hpp:20-66, call for call */
        beta[0] = b.l2_norm();
        q0.mult_scalar(1. / beta[0]);
        copy_vector_element<T>(q0, lc, q, 0);
        A.spmv(q0, w);
        alpha[0] = w.dot(q0);
        w.add(-alpha[0], q0);
        unsigned int j = 0;
        while (j < m - 1) {
            ++j;
            beta[j] = w.l2_norm();
            q1 = w;
            q1.mult_scalar(1. / beta[j]);
            A.spmv(q1, w);
            w.add(-beta[j], q0);
            alpha[j] = w.dot(q1);
            w.add(-alpha[j], q1);
            q0 = q1;
            copy_vector_element<T>(q0, lc, q, j);
        }
        put_f64("alpha", m, alpha.data());
        put_f64("beta", m, beta.data());
        put_f64("q", m, q.data());
        {   /* methods/fdtd.hpp:6-31 call for call (spmv(A,u,dudt) -> the Host member), T_end = 1,
               Nsteps = 100000 as in test_lanczos.cu:118 */
            const unsigned int Nsteps = 100000;
            const double T_end = 1;
            T dt = T_end / Nsteps;
            Vector<T> dudt(b);
            Vector<T> u(b);
            for (unsigned int i = 0; i < Nsteps; ++i) {
                A.spmv(u, dudt);
                u.add(dt, dudt);
            }
            T result = u(lc);
            put_i64("fdtd_steps", Nsteps);
            put_f64("fdtd_u_lc", 1, &result);
        }
    } else if (mode == "block") {
        Dense_matrix<T> B = random_matrix_B<T>(n);       /* test_lanczos.cu:154 (unpadded) */
        put_f64("B", (uint64_t)n * N_COL, B.data());
        const MemorySpace host = MemorySpace::Host;
        Vector<T> q(m * N_COL, host);
        Dense_matrix<T> Q0(B), Q1(B), W(B);
        std::vector<Dense_matrix<T>> alpha(m), beta(m + 1);
        for (unsigned int i = 0; i < m; ++i) {
            alpha[i] = Dense_matrix<T>(N_COL, N_COL, host);
            beta[i] = Dense_matrix<T>(N_COL, N_COL, host);
        }
        beta[m] = Dense_matrix<T>(N_COL, N_COL, host);

        auto gram = [&](Dense_matrix<T> &X, Dense_matrix<T> &Y, Dense_matrix<T> &R) {
            Dense_matrix<T> Xt(X);                       /* gemm(OP_T, OP_N): R = X^T Y */
            Xt.tra();
            R.mm(0., 1., Xt, Y);
        };
        auto sym_gram = [&](Dense_matrix<T> &X, Dense_matrix<T> &Y, Dense_matrix<T> &R) {
            Dense_matrix<T> Xt(X), Yt(Y);                /* lib_utils.hpp:165-202: R = X^T Y, then */
            Xt.tra();                                    /* R = 0.5 R + 0.5 Y^T X                  */
            Yt.tra();
            R.mm(0., 1., Xt, Y);
            R.mm(0.5, 0.5, Yt, X);
        };
        auto row_lc = [&](Dense_matrix<T> &Q, unsigned int off) {
            for (unsigned int c = 0; c < N_COL; ++c) q(off + c) = Q(lc + c * Q.n_rows());
        };

        gram(B, B, beta[0]);                                               /* :106 */
        orc_sqrtm(N_COL, beta[0].data(), beta[m].data());                  /* :111 */
        Q0.mm(0., 1., B, beta[m]);                                         /* :114 */
        row_lc(Q0, 0);                                                     /* :117 */
        A.spmm(Q0, W);                                                     /* :121 */
        sym_gram(W, Q0, alpha[0]);                                         /* :124 */
        W.mm(1., -1., Q0, alpha[0]);                                       /* :128 */
        unsigned int j = 0;
        while (j < m - 1) {
            ++j;
            gram(W, W, beta[j]);                                           /* :137 */
            orc_sqrtm(N_COL, beta[j].data(), beta[m].data());              /* :142 */
            Q1.mm(0., 1., W, beta[m]);                                     /* :145 */
            A.spmm(Q1, W);                                                 /* :149 */
            W.mm(1., -1., Q0, beta[j]);                                    /* :152 */
            sym_gram(W, Q1, alpha[j]);                                     /* :155 */
            W.mm(1., -1., Q1, alpha[j]);                                   /* :159 */
            Q0 = Q1;                                                       /* :162 */
            row_lc(Q0, j * N_COL);                                         /* :165 */
        }
        std::vector<T> a(m * N_COL * N_COL), bt((m + 1) * N_COL * N_COL);
        for (unsigned int i = 0; i < m; ++i)
            std::memcpy(&a[i * N_COL * N_COL], alpha[i].data(), sizeof(T) * N_COL * N_COL);
        for (unsigned int i = 0; i <= m; ++i)
            std::memcpy(&bt[i * N_COL * N_COL], beta[i].data(), sizeof(T) * N_COL * N_COL);
        put_f64("alpha", a.size(), a.data());
        put_f64("beta", bt.size(), bt.data());
        put_f64("q", m * N_COL, q.data());
        {   /* ftdt_block, methods/fdtd.hpp:33-56, call for call; 20000 steps keep the minting run short
               (the harness default of 10^6, test_lanczos.cu:336, is the same loop) */
            const unsigned int Nsteps = 20000;
            const double T_end = 1;
            T dt = T_end / Nsteps;
            Dense_matrix<T> dUdT(B);
            Dense_matrix<T> U(B);
            for (unsigned int i = 0; i < Nsteps; ++i) {
                A.spmm(U, dUdT);
                U.sadd(1, dt, dUdT);
            }
            std::vector<T> row(N_COL);
            for (unsigned int c = 0; c < N_COL; ++c) row[c] = U(lc + c * U.n_rows());
            put_i64("fdtd_steps", Nsteps);
            put_f64("fdtd_row_lc", N_COL, row.data());
        }
    }
    std::fclose(g_out);
    return 0;
}
