// objects/csr_matrix.hpp -- Csr_matrix<Number>: NEW container in the style of Ell_matrix (the
// reference has no CSR): n_rows, n_cols, nnz, row_ptr / col_idx (int32) / data,
// MemorySpace, copy_to_device/host, memory_consumption, spmv/spmm members, and a converter from the
// reference's ELL so Matrix_A feeds the same path.
#ifndef lzb_csr_matrix_hpp
#define lzb_csr_matrix_hpp

#include "ell_matrix.hpp"

template <typename Number>
class Csr_matrix {
    std::size_t _n_rows, _n_cols, _nnz;
    int32_t *_row_ptr, *_col_idx;
    Number *_data;
    MemorySpace _memory_space;
    mutable lz_matrix *_op;

    void release()
    {
        drop_operator();
        if (_memory_space == MemorySpace::CUDA) { lzb::dfree(_row_ptr); lzb::dfree(_col_idx); lzb::dfree(_data); }
        else { delete[] _row_ptr; delete[] _col_idx; delete[] _data; }
        _row_ptr = _col_idx = nullptr; _data = nullptr;
    }
    void set_size(std::size_t n_rows, std::size_t n_cols, std::size_t nnz)
    {
        release();
        if (_memory_space == MemorySpace::CUDA) {
            lzb::require_device_type<Number>();
            _row_ptr = static_cast<int32_t *>(lzb::dmalloc((n_rows + 1) * sizeof(int32_t)));
            _col_idx = static_cast<int32_t *>(lzb::dmalloc((nnz + 8) * sizeof(int32_t)));
            _data = static_cast<Number *>(lzb::dmalloc((nnz + 8) * sizeof(Number)));
        } else {
            _row_ptr = new int32_t[n_rows + 1]();
            _col_idx = new int32_t[nnz + 8]();
            _data = new Number[nnz + 8]();
        }
        _n_rows = n_rows; _n_cols = n_cols; _nnz = nnz;
    }

public:
    Csr_matrix(std::size_t n_rows, std::size_t n_cols, std::size_t nnz, MemorySpace memory_space)
        : _n_rows(0), _n_cols(0), _nnz(0), _row_ptr(nullptr), _col_idx(nullptr), _data(nullptr), _memory_space(memory_space), _op(nullptr)
    {
        set_size(n_rows, n_cols, nnz);
    }
    // from the reference's ELL (either layout, Host): explicit zeros dropped, ELL column order kept
    explicit Csr_matrix(const Ell_matrix<Number> &ell)
        : _n_rows(0), _n_cols(0), _nnz(0), _row_ptr(nullptr), _col_idx(nullptr), _data(nullptr), _memory_space(MemorySpace::Host), _op(nullptr)
    {
        const Ell_matrix<Number> h = ell.memory_space() == MemorySpace::CUDA ? ell.copy_to_host() : ell;
        const std::size_t n = h.n_rows(), w = h.width();
        std::size_t nnz = 0;
        for (std::size_t i = 0; i < h.size(); ++i) nnz += (h(i) != Number(0));
        set_size(n, h.n_cols(), nnz);
        std::size_t p = 0;
        for (std::size_t r = 0; r < n; ++r) {
            _row_ptr[r] = (int32_t)p;
            for (std::size_t k = 0; k < w; ++k) {
                const std::size_t s = h.layout() == 0 ? r + k * n : r * w + k;
                if (h(s) != Number(0)) { _col_idx[p] = (int32_t)h[s]; _data[p] = h(s); ++p; }
            }
        }
        _row_ptr[n] = (int32_t)p;
    }
    Csr_matrix(const Csr_matrix &o)
        : _n_rows(0), _n_cols(0), _nnz(0), _row_ptr(nullptr), _col_idx(nullptr), _data(nullptr), _memory_space(o._memory_space), _op(nullptr)
    {
        set_size(o._n_rows, o._n_cols, o._nnz);
        copy_arrays(o, _memory_space == MemorySpace::CUDA ? LZ_D2D : 0);
    }
    Csr_matrix &operator=(const Csr_matrix &o)
    {
        if (this == &o) return *this;
        release();
        _memory_space = o._memory_space;
        set_size(o._n_rows, o._n_cols, o._nnz);
        copy_arrays(o, _memory_space == MemorySpace::CUDA ? LZ_D2D : 0);
        return *this;
    }
    ~Csr_matrix() { release(); }

    std::size_t n_rows() const { return _n_rows; }
    std::size_t n_cols() const { return _n_cols; }
    std::size_t nnz() const { return _nnz; }
    int32_t *row_ptr() { drop_operator(); return _row_ptr; }
    int32_t *col_idx() { drop_operator(); return _col_idx; }
    Number *data() { drop_operator(); return _data; }
    const int32_t *row_ptr() const { return _row_ptr; }
    const int32_t *col_idx() const { return _col_idx; }
    const Number *data() const { return _data; }
    MemorySpace memory_space() const { return _memory_space; }
    std::size_t memory_consumption() const { return _nnz * (sizeof(Number) + sizeof(int32_t)) + (_n_rows + 1) * sizeof(int32_t); }

    void drop_operator() const
    {
        if (_op) { lz_matrix_destroy(_op); _op = nullptr; }
    }
    lz_matrix *device_operator() const
    {
        if (_memory_space != MemorySpace::CUDA) { std::cout << "implement later" << std::endl; std::abort(); }
        if (!_op)
            AssertCuda(lz_csr_create(lanczos_context(), (int64_t)_n_rows, (int64_t)_n_cols, (int64_t)_nnz, _row_ptr, _col_idx,
                                     reinterpret_cast<const double *>(_data), &_op));
        return _op;
    }
    void spmv(Vector<Number> &vec, Vector<Number> &result) const
    {
        if (_memory_space == MemorySpace::CUDA) {
            AssertCuda(lz_spmv(lanczos_context(), device_operator(), reinterpret_cast<const double *>(vec.data()),
                               reinterpret_cast<double *>(result.data())));
            return;
        }
        for (std::size_t r = 0; r < _n_rows; ++r) {
            Number s = 0;
            for (int32_t p = _row_ptr[r]; p < _row_ptr[r + 1]; ++p) s += _data[p] * vec(_col_idx[p]);
            result(r) = s;
        }
    }
    void spmm(const Dense_matrix<Number> &mat, Dense_matrix<Number> &result) const
    {
        if (_memory_space == MemorySpace::CUDA) {
            AssertCuda(lz_spmm(lanczos_context(), device_operator(), (int)mat.n_cols(), reinterpret_cast<const double *>(mat.data()),
                               (int64_t)mat.n_rows(), reinterpret_cast<double *>(result.data()), (int64_t)result.n_rows()));
            return;
        }
        for (std::size_t c = 0; c < mat.n_cols(); ++c)
            for (std::size_t r = 0; r < _n_rows; ++r) {
                Number s = 0;
                for (int32_t p = _row_ptr[r]; p < _row_ptr[r + 1]; ++p) s += _data[p] * mat(_col_idx[p] + c * mat.n_rows());
                result(r + c * result.n_rows()) = s;
            }
    }
    const Csr_matrix copy_to_device() const
    {
        if (_memory_space == MemorySpace::CUDA) { std::cout << "You are already in the device" << std::endl; return *this; }
        Csr_matrix<Number> other(_n_rows, _n_cols, _nnz, MemorySpace::CUDA);
        other.copy_arrays(*this, LZ_H2D);
        return other;
    }
    const Csr_matrix copy_to_host() const
    {
        if (_memory_space == MemorySpace::Host) { std::cout << "You are already in the host" << std::endl; return *this; }
        Csr_matrix<Number> other(_n_rows, _n_cols, _nnz, MemorySpace::Host);
        other.copy_arrays(*this, LZ_D2H);
        return other;
    }

private:
    void copy_arrays(const Csr_matrix &o, int kind)
    {
        if (kind) {
            lzb::dcopy(_row_ptr, o._row_ptr, (_n_rows + 1) * sizeof(int32_t), kind);
            lzb::dcopy(_col_idx, o._col_idx, _nnz * sizeof(int32_t), kind);
            lzb::dcopy(_data, o._data, _nnz * sizeof(Number), kind);
        } else {
            for (std::size_t i = 0; i <= _n_rows; ++i) _row_ptr[i] = o._row_ptr[i];
            for (std::size_t i = 0; i < _nnz; ++i) { _col_idx[i] = o._col_idx[i]; _data[i] = o._data[i]; }
        }
    }
};

#endif
