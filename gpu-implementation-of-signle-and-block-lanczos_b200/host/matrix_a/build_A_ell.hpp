// matrix_a/build_A_ell.hpp -- Matrix_A<Number>(Nx, Ny, Nz): the reference's test operator
// (matrix_a/build_A_ell.hpp:8-255): the 3-D Maxwell curl operator on a staggered (Yee) grid,
// D = [0 Dh; De 0] in column-major ELL of width 4, and the diagonal metric W such that A = D*W is
// symmetric (Ell_matrix::mult_diagonal).  Returns the pair (D, W) like the reference.
//
// The 1-D difference factors are written straight into ELL form (the reference forms dense
// (N+1) x N matrices, multiplies them and compresses); every stored value is produced by the same
// floating-point operations in the same order, so the arrays are bit-identical -- tests/golden/
// maxwell_N*_matrix.npz, minted by the reference's own builder, pins that.
#ifndef lzb_build_A_ell_hpp
#define lzb_build_A_ell_hpp

#include "build_ell_utils.hpp"

namespace lzb {
// primal -> dual difference: (N+1) x N, row r holds -1/dp[r] at column r-1 and +1/dp[r] at column r
template <typename Number>
Ell_matrix<Number> forward_factor(unsigned int N, const Vector<Number> &dp)
{
    Ell_matrix<Number> X(N + 1, 2 * (N + 1), N, MemorySpace::Host);
    for (unsigned int r = 0; r <= N; ++r) {
        const Number inv = 1. / dp(r);
        unsigned int slot = 0;
        if (r >= 1) { X(r + slot * (N + 1)) = inv * Number(-1.); X[r + slot * (N + 1)] = r - 1; ++slot; }
        if (r < N)  { X(r + slot * (N + 1)) = inv * Number(1.);  X[r + slot * (N + 1)] = r; }
    }
    return X;
}
// dual -> primal difference with the sign flipped: N x (N+1), row r holds -1/dd[r] at r and +1/dd[r] at r+1
template <typename Number>
Ell_matrix<Number> backward_factor(unsigned int N, const Vector<Number> &dd)
{
    Ell_matrix<Number> X(N, 2 * N, N + 1, MemorySpace::Host);
    for (unsigned int r = 0; r < N; ++r) {
        const Number inv = 1. / dd(r);
        X(r) = Number(0.) * (inv * Number(1.)) + Number(-1.) * (inv * Number(1.));      X[r] = r;
        X(r + N) = Number(0.) * (inv * Number(-1.)) + Number(-1.) * (inv * Number(-1.)); X[r + N] = r + 1;
    }
    return X;
}
template <typename Number>
Ell_matrix<Number> diagonal_of(const Vector<Number> &d)
{
    Ell_matrix<Number> W(d.size(), d.size(), d.size(), MemorySpace::Host);
    for (std::size_t i = 0; i < d.size(); ++i) { W(i) = d(i); W[i] = (unsigned int)i; }
    return W;
}
template <typename Number>
struct Axis {
    Ell_matrix<Number> F, B, I, Ip, W, Wh;      // forward / backward factor, identities N and N+1, metrics
};
template <typename Number>
Axis<Number> make_axis(unsigned int N)
{
    const Number lo = 0., hi = 1.;
    const unsigned int Np = N + 2;
    const Number h = (hi - lo) / (Np - 1);
    Vector<Number> p = Linspace<Number>(lo, hi, Np);
    Vector<Number> d = Linspace<Number>(lo, hi - h, Np - 1);
    d.add_scalar(h / 2);
    Vector<Number> dp = Diff<Number>(p), dd = Diff<Number>(d);
    return Axis<Number>{forward_factor<Number>(N, dp), backward_factor<Number>(N, dd), diag<Number>(N, 1.), diag<Number>(N + 1, 1.),
                        diagonal_of<Number>(dp), diagonal_of<Number>(dd)};
}
}  // namespace lzb

template <typename Number>
std::pair<Ell_matrix<Number>, Ell_matrix<Number>> Matrix_A(const unsigned int Nx, const unsigned int Ny, const unsigned int Nz)
{
    const MemorySpace mem = MemorySpace::Host;
    lzb::Axis<Number> x = lzb::make_axis<Number>(Nx), y = lzb::make_axis<Number>(Ny), z = lzb::make_axis<Number>(Nz);
    auto neg = [](Ell_matrix<Number> M) { M.mult_scalar(-1.); return M; };
    // three-factor Kronecker products  K3(a, b, c) = a (x) (b (x) c); exactly one factor is a difference
    auto K3 = [](Ell_matrix<Number> &a, Ell_matrix<Number> &b, Ell_matrix<Number> &c) {
        Ell_matrix<Number> inner = ell_kron(b, c, true, true);
        return ell_kron(a, inner, true, true);
    };
    // curl blocks acting on E (rows: H components) ...
    Ell_matrix<Number> De_12 = neg(K3(z.F, y.Ip, x.I)), De_13 = K3(z.Ip, y.F, x.I);
    Ell_matrix<Number> De_21 = K3(z.F, y.I, x.Ip),      De_23 = neg(K3(z.Ip, y.I, x.F));
    Ell_matrix<Number> De_31 = neg(K3(z.I, y.F, x.Ip)), De_32 = K3(z.I, y.Ip, x.F);
    // ... and on H (rows: E components)
    Ell_matrix<Number> Dh_12 = K3(z.B, y.I, x.Ip),      Dh_13 = neg(K3(z.I, y.B, x.Ip));
    Ell_matrix<Number> Dh_21 = neg(K3(z.B, y.Ip, x.I)), Dh_23 = K3(z.I, y.Ip, x.B);
    Ell_matrix<Number> Dh_31 = K3(z.Ip, y.B, x.I),      Dh_32 = neg(K3(z.Ip, y.I, x.B));

    auto curl = [&](Ell_matrix<Number> &b12, Ell_matrix<Number> &b13, Ell_matrix<Number> &b21, Ell_matrix<Number> &b23,
                    Ell_matrix<Number> &b31, Ell_matrix<Number> &b32, unsigned int c1, unsigned int c2, std::size_t n_cols) {
        const std::size_t rows = b12.n_rows() + b21.n_rows() + b31.n_rows();
        const std::size_t size = b12.size() + b13.size() + b21.size() + b23.size() + b31.size() + b32.size();
        Ell_matrix<Number> C(rows, size, n_cols, mem);
        const unsigned int w = (unsigned int)b12.width(), r2 = (unsigned int)b12.n_rows(), r3 = r2 + (unsigned int)b21.n_rows();
        insert(C, b12, 0, 0, c1);   insert(C, b13, 0, w, c1 + c2);
        insert(C, b21, r2, 0, 0);   insert(C, b23, r2, w, c1 + c2);
        insert(C, b31, r3, 0, 0);   insert(C, b32, r3, w, c1);
        return C;
    };
    const std::size_t De_rows = De_12.n_rows() + De_21.n_rows() + De_31.n_rows();
    const std::size_t Dh_rows = Dh_12.n_rows() + Dh_21.n_rows() + Dh_31.n_rows();
    Ell_matrix<Number> De = curl(De_12, De_13, De_21, De_23, De_31, De_32, (unsigned int)Dh_12.n_rows(), (unsigned int)Dh_21.n_rows(), Dh_rows);
    Ell_matrix<Number> Dh = curl(Dh_12, Dh_13, Dh_21, Dh_23, Dh_31, Dh_32, (unsigned int)De_12.n_rows(), (unsigned int)De_21.n_rows(), De_rows);

    const std::size_t D_rows = De_rows + Dh_rows;
    Ell_matrix<Number> D(D_rows, De.size() + Dh.size(), D_rows, mem);
    insert(D, Dh, 0, 0, (unsigned int)Dh_rows);            // D = [0 Dh; De 0]
    insert(D, De, (unsigned int)Dh.n_rows(), 0, 0);

    // metric W = diag(We, -Wh): products of the 1-D cell sizes, outer * (middle * inner)
    Ell_matrix<Number> We_11 = K3(z.Wh, y.Wh, x.W), We_22 = K3(z.Wh, y.W, x.Wh), We_33 = K3(z.W, y.Wh, x.Wh);
    Ell_matrix<Number> Wh_11 = neg(K3(z.W, y.W, x.Wh)), Wh_22 = neg(K3(z.W, y.Wh, x.W)), Wh_33 = neg(K3(z.Wh, y.W, x.W));
    const std::size_t We_rows = We_11.n_rows() + We_22.n_rows() + We_33.n_rows();
    const std::size_t Wh_rows = Wh_11.n_rows() + Wh_22.n_rows() + Wh_33.n_rows();
    Ell_matrix<Number> W(We_rows + Wh_rows, We_rows + Wh_rows, We_rows + Wh_rows, mem);
    unsigned int at = 0;
    for (Ell_matrix<Number> *blk : {&We_11, &We_22, &We_33, &Wh_11, &Wh_22, &Wh_33}) {
        insert(W, *blk, at, 0, at);
        at += (unsigned int)blk->n_rows();
    }
    return std::make_pair(D, W);
}

#endif
