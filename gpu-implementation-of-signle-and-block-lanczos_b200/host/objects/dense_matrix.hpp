// objects/dense_matrix.hpp -- Dense_matrix<Number>: owning COLUMN-MAJOR matrix, ld = n_rows
// (reference objects/dense_matrix.hpp:9-506).  Same public surface: ctor from sizes / from a Vector
// (diagonal), operator()(row,col) and (index), data()/size()/n_rows()/n_cols(), mm/tra (host only, as
// in the reference), sadd, mult_scalar, padding, copy_to_device/host.
#ifndef lzb_dense_matrix_hpp
#define lzb_dense_matrix_hpp

#include "vector.hpp"

template <typename Number>
class Dense_matrix {
    std::size_t _n_rows, _n_cols, _size;
    Number *_data;
    MemorySpace _memory_space;

    void release()
    {
        if (_memory_space == MemorySpace::CUDA) lzb::dfree(_data);
        else delete[] _data;
        _data = nullptr;
    }
    void set_size(std::size_t n_rows, std::size_t n_cols)
    {
        release();
        _size = n_rows * n_cols;
        if (_memory_space == MemorySpace::CUDA) {
            lzb::require_device_type<Number>();
            _data = static_cast<Number *>(lzb::dmalloc(_size * sizeof(Number)));
        } else {
            _data = new Number[_size];
        }
        _n_rows = n_rows;
        _n_cols = n_cols;
    }
    void fill(Number v)
    {
        if (_memory_space == MemorySpace::CUDA) {
            if (_size) AssertCuda(lz_fill(lanczos_context(), (int64_t)_size, (double)v, reinterpret_cast<double *>(_data)));
        } else {
            for (std::size_t i = 0; i < _size; ++i) _data[i] = v;
        }
    }
    void copy_from(const Dense_matrix &other)
    {
        if (_memory_space == MemorySpace::CUDA) lzb::dcopy(_data, other._data, _size * sizeof(Number), LZ_D2D);
        else for (std::size_t i = 0; i < _size; ++i) _data[i] = other._data[i];
    }

public:
    static const int block_size = Vector<Number>::block_size;
    Dense_matrix() : _n_rows(0), _n_cols(0), _size(0), _data(nullptr), _memory_space(MemorySpace::Host) {}
    Dense_matrix(const std::size_t n_rows, const std::size_t n_cols, const MemorySpace memory_space)
        : _n_rows(0), _n_cols(0), _size(0), _data(nullptr), _memory_space(memory_space)
    {
        set_size(n_rows, n_cols);
        fill(Number(0));
    }
    // diagonal matrix from a vector (dense_matrix.hpp:106-133)
    Dense_matrix(const Vector<Number> &vec) : _n_rows(0), _n_cols(0), _size(0), _data(nullptr), _memory_space(vec.memory_space())
    {
        set_size(vec.size(), vec.size());
        fill(Number(0));
        if (_memory_space == MemorySpace::CUDA) {
            Vector<Number> h = vec.copy_to_host();
            Dense_matrix<Number> tmp(h);
            lzb::dcopy(_data, tmp._data, _size * sizeof(Number), LZ_H2D);
        } else {
            for (std::size_t i = 0; i < _n_rows; ++i) _data[i + i * _n_rows] = vec(i);
        }
    }
    Dense_matrix(const Dense_matrix &other) : _n_rows(0), _n_cols(0), _size(0), _data(nullptr), _memory_space(other._memory_space)
    {
        set_size(other._n_rows, other._n_cols);
        copy_from(other);
    }
    ~Dense_matrix() { release(); }
    Dense_matrix &operator=(const Dense_matrix &other)
    {
        if (this == &other) return *this;
        if (_memory_space != other._memory_space || _n_rows != other._n_rows || _n_cols != other._n_cols) {
            release();
            _memory_space = other._memory_space;
            set_size(other._n_rows, other._n_cols);
        }
        copy_from(other);
        return *this;
    }
    Dense_matrix &operator=(const Number scalar) { fill(scalar); return *this; }

    const Number &operator()(const std::size_t row, const std::size_t col) const { return _data[row + _n_rows * col]; }
    Number &operator()(const std::size_t row, const std::size_t col) { return _data[row + _n_rows * col]; }
    const Number &operator()(const std::size_t index) const { return _data[index]; }
    Number &operator()(const std::size_t index) { return _data[index]; }
    Number *data() { return _data; }
    const Number *data() const { return _data; }
    std::size_t size() const { return _size; }
    std::size_t n_rows() const { return _n_rows; }
    std::size_t n_cols() const { return _n_cols; }
    MemorySpace memory_space() const { return _memory_space; }
    std::size_t memory_consumption() const { return _size * sizeof(Number); }

    // this = my_scalar*this + other_scalar*other1*other2.  Host: the reference's triple loop
    // (dense_matrix.hpp:299-323); CUDA: tall-skinny panel product when other2 is square and small.
    void mm(const Number my_scalar, const Number other_scalar, Dense_matrix &other1, Dense_matrix &other2)
    {
        if (_memory_space == MemorySpace::CUDA) {
            if (other2._n_rows != other2._n_cols || other2._n_rows > 32 || other1._n_cols != other2._n_rows) {
                std::cout << " Dense_matrix::mm on the device supports (n x b)*(b x b), b <= 32 " << std::endl;
                std::abort();
            }
            AssertCuda(lz_mm_ts(lanczos_context(), (int64_t)_n_rows, (int)other2._n_rows, (double)my_scalar, (double)other_scalar,
                                reinterpret_cast<const double *>(other1._data), (int64_t)other1._n_rows,
                                reinterpret_cast<const double *>(other2._data), reinterpret_cast<double *>(_data), (int64_t)_n_rows));
            return;
        }
        Dense_matrix temp(*this);
        for (std::size_t col = 0; col < _n_cols; ++col)
            for (std::size_t row = 0; row < _n_rows; ++row) {
                Number sum = 0;
                for (std::size_t el = 0; el < other1._n_cols; ++el)
                    sum += other1._data[row + _n_rows * el] * other2._data[el + other1._n_cols * col];
                temp._data[row + _n_rows * col] = my_scalar * _data[row + _n_rows * col] + other_scalar * sum;
            }
        for (std::size_t el = 0; el < _size; ++el) _data[el] = temp._data[el];
    }
    void tra()
    {
        if (_memory_space == MemorySpace::CUDA) {
            Dense_matrix h = copy_to_host();
            h.tra();
            *this = h.copy_to_device();
            return;
        }
        Dense_matrix T(_n_cols, _n_rows, _memory_space);
        for (std::size_t col = 0; col < _n_cols; ++col)
            for (std::size_t row = 0; row < _n_rows; ++row) T._data[col + _n_cols * row] = _data[row + _n_rows * col];
        *this = T;
    }
    // this = my_scalar*this + other_scalar*other     (dm::mm_add, dense_kernels.hpp:36-48)
    void sadd(const Number my_scalar, const Number other_scalar, const Dense_matrix &other)
    {
        if (_size != other._size) { std::cout << "The matrices have different sizes" << std::endl; std::abort(); }
        if (_memory_space == MemorySpace::CUDA) {
            AssertCuda(lz_axpby(lanczos_context(), (int64_t)_size, (double)my_scalar, reinterpret_cast<double *>(_data),
                                (double)other_scalar, reinterpret_cast<const double *>(other._data)));
        } else {
            for (std::size_t i = 0; i < _size; ++i) _data[i] = my_scalar * _data[i] + other_scalar * other._data[i];
        }
    }
    void mult_scalar(const Number scalar) { sadd(0., scalar, *this); }
    // pad the rows up to a multiple of `pads` with zeros (dense_matrix.hpp:437-467); the new kernels
    // need no padding, the call is kept for source compatibility with test_lanczos.cu:174-187
    void padding(const unsigned int pads)
    {
        const std::size_t new_rows = ((_n_rows + pads - 1) / pads) * pads;
        if (new_rows == _n_rows) return;
        const bool dev = _memory_space == MemorySpace::CUDA;
        Dense_matrix h = dev ? copy_to_host() : *this;
        Dense_matrix r(new_rows, _n_cols, MemorySpace::Host);
        for (std::size_t c = 0; c < _n_cols; ++c)
            for (std::size_t i = 0; i < _n_rows; ++i) r._data[i + c * new_rows] = h._data[i + c * _n_rows];
        *this = dev ? r.copy_to_device() : r;
    }
    const Dense_matrix copy_to_device() const
    {
        if (_memory_space == MemorySpace::CUDA) { std::cout << "You are already in the device" << std::endl; return *this; }
        Dense_matrix<Number> other(_n_rows, _n_cols, MemorySpace::CUDA);
        lzb::dcopy(other._data, _data, _size * sizeof(Number), LZ_H2D);
        return other;
    }
    const Dense_matrix copy_to_host() const
    {
        if (_memory_space == MemorySpace::Host) { std::cout << "You are already in the host" << std::endl; return *this; }
        Dense_matrix<Number> other(_n_rows, _n_cols, MemorySpace::Host);
        lzb::dcopy(other._data, _data, _size * sizeof(Number), LZ_D2H);
        return other;
    }
    void print() const
    {
        if (_memory_space == MemorySpace::CUDA) { copy_to_host().print(); return; }
        for (std::size_t r = 0; r < _n_rows; ++r) {
            for (std::size_t c = 0; c < _n_cols; ++c) std::cout << std::setprecision(10) << _data[r + c * _n_rows] << " ";
            std::cout << std::endl;
        }
    }
};

#endif
