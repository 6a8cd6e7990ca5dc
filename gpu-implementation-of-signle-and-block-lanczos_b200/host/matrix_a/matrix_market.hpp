// matrix_a/matrix_market.hpp -- MatrixMarket coordinate reader for the new Csr_matrix (matrix
// ingest; the reference only builds its Maxwell operator in code, matrix_a/build_A_ell.hpp).
// Supports "matrix coordinate real|integer|pattern general|symmetric": entries are sorted by (row, column),
// duplicates are summed, the lower/upper mirror of a symmetric file is materialised, pattern entries get 1.
// Returns a Host-space matrix; copy_to_device() hands it to the library like any other Csr_matrix.
#ifndef lzb_matrix_market_hpp
#define lzb_matrix_market_hpp

#include <algorithm>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../objects/csr_matrix.hpp"

template <typename Number>
Csr_matrix<Number> read_matrix_market(const std::string &path)
{
    std::ifstream in(path);
    if (!in) throw std::runtime_error("read_matrix_market: cannot open " + path);
    std::string line;
    if (!std::getline(in, line)) throw std::runtime_error("read_matrix_market: empty file " + path);
    std::istringstream hdr(line);
    std::string banner, object, format, field, symmetry;
    hdr >> banner >> object >> format >> field >> symmetry;
    for (std::string *s : {&object, &format, &field, &symmetry})
        std::transform(s->begin(), s->end(), s->begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (banner != "%%MatrixMarket" || object != "matrix" || format != "coordinate")
        throw std::runtime_error("read_matrix_market: only 'matrix coordinate' files are supported");
    const bool pattern = field == "pattern";
    if (!pattern && field != "real" && field != "integer") throw std::runtime_error("read_matrix_market: unsupported field " + field);
    const bool symmetric = symmetry == "symmetric";
    if (!symmetric && symmetry != "general") throw std::runtime_error("read_matrix_market: unsupported symmetry " + symmetry);
    while (std::getline(in, line))
        if (!line.empty() && line[0] != '%') break;
    std::size_t n_rows = 0, n_cols = 0, n_entries = 0;
    {
        std::istringstream sz(line);
        if (!(sz >> n_rows >> n_cols >> n_entries)) throw std::runtime_error("read_matrix_market: bad size line");
    }
    std::vector<std::tuple<int32_t, int32_t, Number>> e;
    e.reserve(symmetric ? 2 * n_entries : n_entries);
    for (std::size_t k = 0; k < n_entries; ++k) {
        std::size_t r, c;
        double v = 1.0;
        if (!(in >> r >> c)) throw std::runtime_error("read_matrix_market: truncated file");
        if (!pattern && !(in >> v)) throw std::runtime_error("read_matrix_market: truncated file");
        if (r < 1 || r > n_rows || c < 1 || c > n_cols) throw std::runtime_error("read_matrix_market: index out of range");
        e.emplace_back((int32_t)(r - 1), (int32_t)(c - 1), (Number)v);
        if (symmetric && r != c) e.emplace_back((int32_t)(c - 1), (int32_t)(r - 1), (Number)v);
    }
    std::sort(e.begin(), e.end(), [](const auto &a, const auto &b) {
        return std::get<0>(a) != std::get<0>(b) ? std::get<0>(a) < std::get<0>(b) : std::get<1>(a) < std::get<1>(b);
    });
    std::size_t nnz = 0;                                  // after summing duplicates
    for (std::size_t k = 0; k < e.size(); ++k)
        if (k == 0 || std::get<0>(e[k]) != std::get<0>(e[k - 1]) || std::get<1>(e[k]) != std::get<1>(e[k - 1])) ++nnz;
    Csr_matrix<Number> A(n_rows, n_cols, nnz, MemorySpace::Host);
    int32_t *rp = A.row_ptr(), *ci = A.col_idx();
    Number *va = A.data();
    for (std::size_t r = 0; r <= n_rows; ++r) rp[r] = 0;
    std::size_t p = 0;
    for (std::size_t k = 0; k < e.size(); ++k) {
        if (k > 0 && std::get<0>(e[k]) == std::get<0>(e[k - 1]) && std::get<1>(e[k]) == std::get<1>(e[k - 1])) {
            va[p - 1] += std::get<2>(e[k]);
            continue;
        }
        ci[p] = std::get<1>(e[k]);
        va[p] = std::get<2>(e[k]);
        rp[std::get<0>(e[k]) + 1]++;
        ++p;
    }
    for (std::size_t r = 0; r < n_rows; ++r) rp[r + 1] += rp[r];
    return A;
}

#endif
