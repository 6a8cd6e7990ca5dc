// objects/ell_matrix.hpp -- Ell_matrix<Number>: the reference's ELLPACK container
// (objects/ell_matrix.hpp:10-544): values Number*, indices unsigned*, n_rows/n_cols/size/width,
// column-major data[r + k*n_rows] until change_order(4) interleaves rows (data[4r + k]).
// operator() indexes values, operator[] indexes column ids, exactly as in the reference (:100-115).
// The device operator handle (lz_matrix) is created lazily by spmv/spmm and dropped on any change.
#ifndef lzb_ell_matrix_hpp
#define lzb_ell_matrix_hpp

#include "dense_matrix.hpp"

template <typename Number>
class Ell_matrix {
    std::size_t _n_rows, _n_cols, _size, _width;
    Number *_data;
    unsigned int *_idx;
    MemorySpace _memory_space;
    int _layout;                 // 0: column-major (reference default), 1: row-interleaved by `width`
    mutable lz_matrix *_op;      // cached device operator

    void release()
    {
        drop_operator();
        if (_memory_space == MemorySpace::CUDA) { lzb::dfree(_data); lzb::dfree(_idx); }
        else { delete[] _data; delete[] _idx; }
        _data = nullptr; _idx = nullptr;
    }
    void set_size(std::size_t n_rows, std::size_t size, std::size_t n_cols)
    {
        release();
        if (_memory_space == MemorySpace::CUDA) {
            lzb::require_device_type<Number>();
            _data = static_cast<Number *>(lzb::dmalloc(size * sizeof(Number)));
            _idx = static_cast<unsigned int *>(lzb::dmalloc(size * sizeof(unsigned int)));
        } else {
            _data = new Number[size];
            _idx = new unsigned int[size];
        }
        _n_rows = n_rows; _n_cols = n_cols; _size = size; _width = n_rows ? size / n_rows : 0;
    }
    void copy_from(const Ell_matrix &o)
    {
        const int kind = _memory_space == MemorySpace::CUDA ? LZ_D2D : 0;
        if (kind) {
            lzb::dcopy(_data, o._data, _size * sizeof(Number), kind);
            lzb::dcopy(_idx, o._idx, _size * sizeof(unsigned int), kind);
        } else {
            for (std::size_t i = 0; i < _size; ++i) { _data[i] = o._data[i]; _idx[i] = o._idx[i]; }
        }
        _layout = o._layout;
    }

public:
    static const int block_size = Vector<Number>::block_size;
    Ell_matrix(const std::size_t n_rows, const std::size_t size, const std::size_t n_cols, const MemorySpace memory_space)
        : _n_rows(0), _n_cols(0), _size(0), _width(0), _data(nullptr), _idx(nullptr), _memory_space(memory_space), _layout(0), _op(nullptr)
    {
        set_size(n_rows, size, n_cols);
        if (_memory_space == MemorySpace::CUDA) {
            AssertCuda(lz_memset(lanczos_context(), _data, 0, _size * sizeof(Number)));
            AssertCuda(lz_memset(lanczos_context(), _idx, 0, _size * sizeof(unsigned int)));
        } else {
            for (std::size_t i = 0; i < _size; ++i) { _data[i] = 0; _idx[i] = 0; }
        }
    }
    Ell_matrix(const Ell_matrix &other)
        : _n_rows(0), _n_cols(0), _size(0), _width(0), _data(nullptr), _idx(nullptr), _memory_space(other._memory_space), _layout(0), _op(nullptr)
    {
        set_size(other._n_rows, other._size, other._n_cols);
        copy_from(other);
    }
    ~Ell_matrix() { release(); }
    Ell_matrix &operator=(const Ell_matrix &other)
    {
        if (this == &other) return *this;
        release();
        _memory_space = other._memory_space;
        set_size(other._n_rows, other._size, other._n_cols);
        copy_from(other);
        return *this;
    }
    // () values, [] column ids -- a write through either invalidates the cached operator
    const Number &operator()(const std::size_t index) const { return _data[index]; }
    Number &operator()(const std::size_t index) { drop_operator(); return _data[index]; }
    const unsigned int &operator[](const std::size_t index) const { return _idx[index]; }
    unsigned int &operator[](const std::size_t index) { drop_operator(); return _idx[index]; }

    Number *data() { drop_operator(); return _data; }
    const Number *data() const { return _data; }
    unsigned int *idx() { drop_operator(); return _idx; }
    const unsigned int *idx() const { return _idx; }
    std::size_t n_rows() const { return _n_rows; }
    std::size_t n_cols() const { return _n_cols; }
    std::size_t size() const { return _size; }
    std::size_t width() const { return _width; }
    int layout() const { return _layout; }
    MemorySpace memory_space() const { return _memory_space; }
    std::size_t memory_consumption() const { return _size * (sizeof(Number) + sizeof(unsigned int)); }

    void drop_operator() const
    {
        if (_op) { lz_matrix_destroy(_op); _op = nullptr; }
    }
    // device operator over this matrix's arrays (borrowed); any width, either layout
    lz_matrix *device_operator() const
    {
        if (_memory_space != MemorySpace::CUDA) { std::cout << "implement later" << std::endl; std::abort(); }  // spmv_spmm.hpp:229-232
        if (!_op)
            AssertCuda(lz_ell_create(lanczos_context(), (int64_t)_n_rows, (int64_t)_n_cols, (int)_width, _layout,
                                     reinterpret_cast<const double *>(_data), _idx, &_op));
        return _op;
    }

    // generic-width SpMV / SpMM (ell_matrix.hpp:228-301); the Host branches are the reference loops
    void spmv(Vector<Number> &vec, Vector<Number> &result) const
    {
        if (_memory_space == MemorySpace::CUDA) {
            AssertCuda(lz_spmv(lanczos_context(), device_operator(), reinterpret_cast<const double *>(vec.data()),
                               reinterpret_cast<double *>(result.data())));
            return;
        }
        result = 0.;
        if (_layout == 0) {
            for (std::size_t i = 0; i < _size; ++i) result(i % _n_rows) += _data[i] * vec(_idx[i]);
        } else {
            for (std::size_t i = 0; i < _size; ++i) result(i / _width) += _data[i] * vec(_idx[i]);
        }
    }
    void spmm(const Dense_matrix<Number> &mat, Dense_matrix<Number> &result) const
    {
        if (_memory_space == MemorySpace::CUDA) {
            AssertCuda(lz_spmm(lanczos_context(), device_operator(), (int)mat.n_cols(), reinterpret_cast<const double *>(mat.data()),
                               (int64_t)mat.n_rows(), reinterpret_cast<double *>(result.data()), (int64_t)result.n_rows()));
            return;
        }
        result = 0;
        for (std::size_t c = 0; c < mat.n_cols(); ++c)
            for (std::size_t i = 0; i < _size; ++i) {
                const std::size_t r = _layout == 0 ? i % _n_rows : i / _width;
                result(r + c * result.n_rows()) += _data[i] * mat(_idx[i] + c * mat.n_rows());
            }
    }
    void mult_scalar(Number scalar)
    {
        host_only("mult_scalar");
        for (std::size_t i = 0; i < _size; ++i) _data[i] = _data[i] * scalar;
        drop_operator();
    }
    void diag_inv()
    {
        host_only("diag_inv");
        for (std::size_t i = 0; i < _size; ++i) _data[i] = 1. / _data[i];
    }
    void diag_sqrt()
    {
        host_only("diag_sqrt");
        for (std::size_t i = 0; i < _size; ++i) _data[i] = std::sqrt(_data[i]);
    }
    // this = this * diag  (data[i] *= diag.data[idx[i]], ell_matrix.hpp:340-361)
    void mult_diagonal(const Ell_matrix &diag)
    {
        host_only("mult_diagonal");
        for (std::size_t i = 0; i < _size; ++i) _data[i] = _data[i] * diag._data[_idx[i]];
        drop_operator();
    }
    // column-major -> row-interleaved by `stride` columns (the INTENDED effect of change_order,
    // i.e. what the reference's CUDA branch lm::change_major does, ell_kernels.hpp:99-121; its Host
    // branch only moves ELL column 0 -- SURVEY appendix A-1 -- and is not reproduced)
    void change_order(const unsigned int stride)
    {
        if (_layout == 1 || stride != _width) {
            if (_layout == 1) return;
            std::cout << " change_order: stride must equal the ELL width " << std::endl;
            std::abort();
        }
        const bool dev = _memory_space == MemorySpace::CUDA;
        Ell_matrix h = dev ? copy_to_host() : *this;
        Ell_matrix r(_n_rows, _size, _n_cols, MemorySpace::Host);
        for (std::size_t row = 0; row < _n_rows; ++row)
            for (std::size_t k = 0; k < _width; ++k) {
                r._data[row * _width + k] = h._data[row + k * _n_rows];
                r._idx[row * _width + k] = h._idx[row + k * _n_rows];
            }
        r._layout = 1;
        *this = dev ? r.copy_to_device() : r;
    }
    const Ell_matrix copy_to_device() const
    {
        if (_memory_space == MemorySpace::CUDA) { std::cout << "You are already in the device" << std::endl; return *this; }
        Ell_matrix<Number> other(_n_rows, _size, _n_cols, MemorySpace::CUDA);
        lzb::dcopy(other._data, _data, _size * sizeof(Number), LZ_H2D);
        lzb::dcopy(other._idx, _idx, _size * sizeof(unsigned int), LZ_H2D);
        other._layout = _layout;
        return other;
    }
    const Ell_matrix copy_to_host() const
    {
        if (_memory_space == MemorySpace::Host) { std::cout << "You are already in the host" << std::endl; return *this; }
        Ell_matrix<Number> other(_n_rows, _size, _n_cols, MemorySpace::Host);
        lzb::dcopy(other._data, _data, _size * sizeof(Number), LZ_D2H);
        lzb::dcopy(other._idx, _idx, _size * sizeof(unsigned int), LZ_D2H);
        other._layout = _layout;
        return other;
    }
    void print() const
    {
        if (_memory_space == MemorySpace::CUDA) { copy_to_host().print(); return; }
        for (std::size_t i = 0; i < _size; ++i) std::cout << _data[i] << " (" << _idx[i] << ") ";
        std::cout << std::endl;
    }

private:
    void host_only(const char *what) const
    {
        if (_memory_space == MemorySpace::CUDA) { std::cout << " not implemented: Ell_matrix::" << what << " on the device" << std::endl; std::abort(); }
    }
};

#endif
