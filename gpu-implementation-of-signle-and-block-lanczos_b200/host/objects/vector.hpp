// objects/vector.hpp -- Vector<Number>, the reference's owning 1-D container (objects/vector.hpp:17-376)
// with the same public surface: MemorySpace tag, sadd/add/mult_scalar/add_scalar, dot/l2_norm,
// copy_to_device/copy_to_host, data()/size()/memory_consumption(), operator=(scalar).
// Host branches are plain loops (same expressions as the reference, :228-231, :268-274); CUDA
// branches forward to the C-ABI.  Differences that matter for performance, not for results:
// operator= reuses the allocation when sizes match, dot() allocates nothing.
#ifndef lzb_vector_hpp
#define lzb_vector_hpp

#include "../utils/common.hpp"

enum class MemorySpace { Host, CUDA };   // objects/vector.hpp:11-15

namespace lzb {
template <typename Number>
struct device_ok { static const bool value = false; };
template <>
struct device_ok<double> { static const bool value = true; };
template <typename Number>
inline void require_device_type()
{
    if (!device_ok<Number>::value) {
        std::cerr << "lanczos_b200: the device path is fp64 only (Number = double)" << std::endl;
        std::abort();
    }
}
inline void *dmalloc(std::size_t bytes)
{
    void *p = nullptr;
    AssertCuda(lz_malloc(lanczos_context(), bytes, &p));
    return p;
}
inline void dfree(void *p)
{
    if (p) AssertCuda(lz_free(lanczos_context(), p));
}
inline void dcopy(void *dst, const void *src, std::size_t bytes, int kind)
{
    AssertCuda(lz_memcpy(lanczos_context(), dst, src, bytes, kind));
}
}  // namespace lzb

template <typename Number>
class Vector {
    Number *_data;
    std::size_t _size;
    MemorySpace _memory_space;

    void release()
    {
        if (_memory_space == MemorySpace::CUDA) lzb::dfree(_data);
        else delete[] _data;
        _data = nullptr;
    }
    void set_size(std::size_t size)
    {
        release();
        if (_memory_space == MemorySpace::CUDA) {
            lzb::require_device_type<Number>();
            _data = static_cast<Number *>(lzb::dmalloc(size * sizeof(Number)));
        } else {
            _data = new Number[size];
        }
        _size = size;
    }
    void assert_size(const Vector &other) const
    {
        if (_size != other._size) {
            std::cout << "The vectors have different sizes" << std::endl;
            std::abort();
        }
    }

public:
    static const int block_size = 256;

    Vector(std::vector<Number> &array, const MemorySpace memory_space) : _data(nullptr), _size(0), _memory_space(memory_space)
    {
        set_size(array.size());
        if (_memory_space == MemorySpace::CUDA) lzb::dcopy(_data, array.data(), _size * sizeof(Number), LZ_H2D);
        else for (std::size_t i = 0; i < _size; ++i) _data[i] = array[i];
    }
    Vector(const std::size_t size, const MemorySpace memory_space) : Vector(size, Number(0), memory_space) {}
    Vector(std::size_t size, const Number scalar, const MemorySpace memory_space) : _data(nullptr), _size(0), _memory_space(memory_space)
    {
        set_size(size);
        *this = scalar;
    }
    Vector(const Vector &other) : _data(nullptr), _size(0), _memory_space(other._memory_space)
    {
        set_size(other._size);
        copy_from(other);
    }
    ~Vector() { release(); }

    Vector &operator=(const Vector &other)
    {
        if (this == &other) return *this;
        if (_memory_space != other._memory_space || _size != other._size) {
            release();
            _memory_space = other._memory_space;
            set_size(other._size);
        }
        copy_from(other);
        return *this;
    }
    Vector &operator=(const Number scalar)
    {
        if (_memory_space == MemorySpace::CUDA) {
            if (_size) AssertCuda(lz_fill(lanczos_context(), (int64_t)_size, (double)scalar, reinterpret_cast<double *>(_data)));
        } else {
            for (std::size_t i = 0; i < _size; ++i) _data[i] = scalar;
        }
        return *this;
    }
    const Number &operator()(const std::size_t index) const { return _data[index]; }
    Number &operator()(const std::size_t index) { return _data[index]; }
    MemorySpace memory_space() const { return _memory_space; }

    void add_scalar(const Number scalar)
    {
        Vector<Number> vec(_size, scalar, _memory_space);
        sadd(1., 1., vec);
    }
    void mult_scalar(const Number scalar) { sadd(0., scalar, *this); }
    void add(const Number vec_scalar, const Vector &vec) { sadd(1., vec_scalar, vec); }
    // this = my_scalar*this + vec_scalar*vec          (v::vector_update, vector_kernels.hpp:22-33)
    void sadd(const Number my_scalar, const Number vec_scalar, const Vector &vec)
    {
        assert_size(vec);
        if (_memory_space == MemorySpace::CUDA) {
            AssertCuda(lz_axpby(lanczos_context(), (int64_t)_size, (double)my_scalar, reinterpret_cast<double *>(_data),
                                (double)vec_scalar, reinterpret_cast<const double *>(vec._data)));
        } else {
            for (std::size_t i = 0; i < _size; ++i) _data[i] = my_scalar * _data[i] + vec_scalar * vec._data[i];
        }
    }
    Number l2_norm() const
    {
        const Number norm_squared = norm_square();
        if (std::isfinite(norm_squared)) return std::sqrt(norm_squared);
        std::cout << "The norm is not finite" << std::endl;     // vector.hpp:239-241
        std::abort();
        return 0;
    }
    Number norm_square() const { return dot(*this); }
    Number dot(const Vector &other) const
    {
        assert_size(other);
        if (_memory_space == MemorySpace::CUDA) {
            double r = 0.0;
            AssertCuda(lz_dot(lanczos_context(), (int64_t)_size, reinterpret_cast<const double *>(_data),
                              reinterpret_cast<const double *>(other._data), &r));
            return (Number)r;
        }
        Number sum = 0;
        for (std::size_t i = 0; i < _size; ++i) sum += _data[i] * other._data[i];
        return sum;
    }
    const Vector copy_to_device() const
    {
        if (_memory_space == MemorySpace::CUDA) {
            std::cout << "You are already in the device" << std::endl;
            return *this;
        }
        Vector<Number> other(_size, MemorySpace::CUDA);
        lzb::dcopy(other._data, _data, _size * sizeof(Number), LZ_H2D);
        return other;
    }
    const Vector copy_to_host() const
    {
        if (_memory_space == MemorySpace::Host) {
            std::cout << "You are already in the host" << std::endl;
            return *this;
        }
        Vector<Number> other(_size, MemorySpace::Host);
        lzb::dcopy(other._data, _data, _size * sizeof(Number), LZ_D2H);
        return other;
    }
    Number *data() { return _data; }
    const Number *data() const { return _data; }
    std::size_t size() const { return _size; }
    std::size_t memory_consumption() const { return _size * sizeof(Number); }
    void print() const
    {
        if (_memory_space == MemorySpace::CUDA) { copy_to_host().print(); return; }
        for (std::size_t i = 0; i < _size; ++i) std::cout << std::setprecision(12) << _data[i] << " ";
        std::cout << std::endl;
    }

private:
    void copy_from(const Vector &other)
    {
        if (_memory_space == MemorySpace::CUDA) lzb::dcopy(_data, other._data, _size * sizeof(Number), LZ_D2D);
        else for (std::size_t i = 0; i < _size; ++i) _data[i] = other._data[i];
    }
};

#endif
