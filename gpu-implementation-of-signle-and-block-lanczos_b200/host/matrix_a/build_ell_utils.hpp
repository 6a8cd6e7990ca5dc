// matrix_a/build_ell_utils.hpp -- host helpers of the problem generator, same names and meaning as
// the reference's matrix_a/build_ell_utils.hpp:8-280: Kronecker products on ELL operands
// (Ikron, kronI), block placement (insert), 1-D grids (Linspace, Diff), small builders (diag,
// bidiagonal, dense_to_ell, diag_inv) and the right-hand sides (gaussian_* / random_*).
// Both Kronecker flavours are one routine here (ell_kron); results are bit-identical to the
// reference's because every entry is a single product of the same two operands in the same order.
#ifndef lzb_build_ell_utils_hpp
#define lzb_build_ell_utils_hpp

#include <tuple>

#include "../methods/copy_functions.hpp"
#include "../objects/ell_matrix.hpp"

#ifndef N_COL
#define N_COL 4
#endif

// kron(A, B) for column-major ELL operands.  Row (ra*Brows + rb), slot (ka*Bwidth + kb) holds
// A(ra,ka)*B(rb,kb) at column A.col*Bcols + B.col.  use_a / use_b select which factors contribute
// their values (the reference's Ikron ignores the identity's values, kronI multiplies by them).
template <typename Number>
Ell_matrix<Number> ell_kron(const Ell_matrix<Number> &A, const Ell_matrix<Number> &B, bool use_a, bool use_b)
{
    const std::size_t ar = A.n_rows(), br = B.n_rows(), aw = A.width(), bw = B.width(), bc = B.n_cols();
    const std::size_t rows = ar * br, width = aw * bw;
    Ell_matrix<Number> out(rows, rows * width, A.n_cols() * bc, A.memory_space());
    for (std::size_t ka = 0; ka < aw; ++ka)
        for (std::size_t kb = 0; kb < bw; ++kb)
            for (std::size_t ra = 0; ra < ar; ++ra)
                for (std::size_t rb = 0; rb < br; ++rb) {
                    const std::size_t sa = ra + ka * ar, sb = rb + kb * br;
                    const std::size_t dst = (ra * br + rb) + (ka * bw + kb) * rows;
                    const Number va = A(sa), vb = B(sb);
                    out(dst) = use_a && use_b ? va * vb : (use_a ? va : vb);
                    out[dst] = (unsigned int)(A[sa] * bc + B[sb]);
                }
    return out;
}
// I (x) X : the identity contributes its size only                   (reference :8-33)
template <typename Number>
Ell_matrix<Number> Ikron(Ell_matrix<Number> &I, Ell_matrix<Number> &X) { return ell_kron(I, X, false, true); }
// X (x) I : values X(i)*I(j)                                          (reference :37-58)
template <typename Number>
Ell_matrix<Number> kronI(Ell_matrix<Number> &X, Ell_matrix<Number> &I) { return ell_kron(X, I, true, true); }

// place block d into D at row r_loc, ELL slot c_loc, shifting its column ids by c_shift (:61-81)
template <typename Number>
void insert(Ell_matrix<Number> &D, Ell_matrix<Number> &d, const unsigned int r_loc, const unsigned int c_loc, const unsigned int c_shift)
{
    const std::size_t Dr = D.n_rows(), dr = d.n_rows();
    for (std::size_t k = 0; k < d.width(); ++k)
        for (std::size_t r = 0; r < dr; ++r) {
            const std::size_t dst = (r_loc + r) + (c_loc + k) * Dr;
            D(dst) = d(r + k * dr);
            D[dst] = d[r + k * dr] + c_shift;
        }
}

template <typename Number>
Vector<Number> Diff(const Vector<Number> &other)
{
    Vector<Number> diff(other.size() - 1, MemorySpace::Host);
    for (std::size_t i = 0; i + 1 < other.size(); ++i) diff(i) = other(i + 1) - other(i);
    return diff;
}
template <typename Number>
Vector<Number> Linspace(const Number x_l, const Number x_r, const unsigned int N)
{
    const Number h = (x_r - x_l) / (N - 1);
    Vector<Number> grid(N, MemorySpace::Host);
    for (unsigned int i = 0; i < N; ++i) grid(i) = x_l + i * h;
    return grid;
}
template <typename Number>
Ell_matrix<Number> diag(const unsigned int N, const Number scalar)
{
    Ell_matrix<Number> I(N, N, N, MemorySpace::Host);
    for (unsigned int i = 0; i < N; ++i) { I(i) = scalar; I[i] = i; }
    return I;
}
// N x (N+1) with `diagonal` on (i,i) and `upper_diagonal` on (i,i+1)
template <typename Number>
Dense_matrix<Number> bidiagonal(const unsigned int N, Number diagonal, Number upper_diagonal)
{
    Dense_matrix<Number> result(N, N + 1, MemorySpace::Host);
    for (unsigned int i = 0; i < N; ++i) { result(i, i) = diagonal; result(i, i + 1) = upper_diagonal; }
    return result;
}
// non-zeros of each dense row, left to right, into `width` ELL slots
template <typename Number>
Ell_matrix<Number> dense_to_ell(const Dense_matrix<Number> &dense, unsigned int width)
{
    const std::size_t n_rows = dense.n_rows(), n_cols = dense.n_cols();
    Ell_matrix<Number> ell(n_rows, width * n_rows, n_cols, dense.memory_space());
    for (std::size_t i = 0; i < n_rows; ++i) {
        std::size_t slot = 0;
        for (std::size_t j = 0; j < n_cols; ++j)
            if (dense(i, j) != 0) { ell(i + slot * n_rows) = dense(i, j); ell[i + slot * n_rows] = (unsigned int)j; ++slot; }
    }
    return ell;
}
template <typename Number>
Dense_matrix<Number> diag_inv(Dense_matrix<Number> &diag_mat)
{
    Dense_matrix<Number> result(diag_mat);
    for (std::size_t i = 0; i < diag_mat.n_rows(); ++i) result(i, i) = 1. / diag_mat(i, i);
    return result;
}

// ---- right-hand sides -------------------------------------------------------------------------
template <typename Number>
std::tuple<Vector<Number>, Vector<Number>, Vector<Number>> grid_3D(Vector<Number> &x, Vector<Number> &y, Vector<Number> &z)
{
    const std::size_t xs = x.size(), ys = y.size(), zs = z.size(), size = xs * ys * zs;
    Vector<Number> X(size, x.memory_space()), Y(size, x.memory_space()), Z(size, x.memory_space());
    for (std::size_t i = 0; i < size; ++i) {
        X(i) = x(i % xs);
        Y(i) = y((i / xs) % ys);
        Z(i) = z((i / (xs * ys)) % zs);
    }
    return std::make_tuple(X, Y, Z);
}
// exp(-|p - shift|^2) on the grid points; entries past the grid stay zero (one field component only)
template <typename Number>
Vector<Number> gaussian_3D(const unsigned int N, Vector<Number> &x, Vector<Number> &y, Vector<Number> &z, Number shift)
{
    auto g = grid_3D<Number>(x, y, z);
    Vector<Number> &X = std::get<0>(g), &Y = std::get<1>(g), &Z = std::get<2>(g);
    Vector<Number> result(N, x.memory_space());
    for (std::size_t i = 0; i < X.size(); ++i)
        result(i) = std::exp(-std::pow(X(i) - shift, 2) - std::pow(Y(i) - shift, 2) - std::pow(Z(i) - shift, 2));
    return result;
}
namespace lzb {
template <typename Number>
void gaussian_axes(unsigned int N, Vector<Number> &xp, Vector<Number> &yp, Vector<Number> &zd)
{
    const Number lo = 0., hi = 1., h = (hi - lo) / (N + 1);
    xp = Linspace<Number>(lo + h, hi - h, N);
    yp = Linspace<Number>(lo + h, hi - h, N);
    zd = Linspace<Number>(lo + h / 2, hi - h / 2, N + 1);
}
}  // namespace lzb
template <typename Number>
Vector<Number> gaussian_vector_b(const unsigned int N, const unsigned int n_rows)
{
    Vector<Number> xp(1, MemorySpace::Host), yp(1, MemorySpace::Host), zd(1, MemorySpace::Host);
    lzb::gaussian_axes<Number>(N, xp, yp, zd);
    return gaussian_3D<Number>(n_rows, xp, yp, zd, 0.5);
}
template <typename Number>
Vector<Number> random_vector_b(const unsigned int n_rows)
{
    Vector<Number> b_host(n_rows, MemorySpace::Host);
    for (unsigned int i = 0; i < n_rows; ++i) b_host(i) = ((Number)rand() / (RAND_MAX)) + 1;
    return b_host;
}
template <typename Number>
Dense_matrix<Number> gaussian_matrix_B(const unsigned int N, const unsigned int n_rows, const unsigned int n_col = N_COL)
{
    Vector<Number> xp(1, MemorySpace::Host), yp(1, MemorySpace::Host), zd(1, MemorySpace::Host);
    lzb::gaussian_axes<Number>(N, xp, yp, zd);
    Dense_matrix<Number> B_host(n_rows, n_col, MemorySpace::Host);
    for (unsigned int i = 0; i < n_col; ++i) {
        Vector<Number> b_host = gaussian_3D<Number>(n_rows, xp, yp, zd, 0.1 * (i + 1));
        copy_vector_to_column(b_host, B_host, i);
    }
    return B_host;
}
template <typename Number>
Dense_matrix<Number> random_matrix_B(const unsigned int n_rows, const unsigned int n_col = N_COL)
{
    Dense_matrix<Number> B_host(n_rows, n_col, MemorySpace::Host);
    for (std::size_t i = 0; i < (std::size_t)n_rows * n_col; ++i) B_host(i) = ((Number)rand() / (RAND_MAX)) + 1;
    return B_host;
}

#endif
