// Minimal stand-in for <boost/multi_array.hpp> -- TEST INFRASTRUCTURE ONLY.
//
// Boost is not installed in this image.  The reference's hot-path sources use a
// small subset of boost::multi_array (dense row-major storage, nested operator[]
// views, origin()/begin()/end() on views, element-wise view assignment).  This
// header provides exactly that subset so that the reference sources under
// /root/reference can be compiled *in place* into oracle/_ref/ (see
// oracle/Makefile).  It is our own code, written from the documented semantics
// of the library; it is never linked into the product library.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cstddef>
#include <vector>

namespace boost {

namespace ma_detail {
template <std::size_t N>
struct extent_gen {
    std::array<std::size_t, N> e;
    extent_gen<N + 1> operator[](std::size_t n) const {
        extent_gen<N + 1> r;
        for (std::size_t i = 0; i < N; i++) r.e[i] = e[i];
        r.e[N] = n;
        return r;
    }
};

// view onto an N-dimensional dense row-major block
template <typename T, std::size_t N>
class view {
public:
    view(T* base, const std::size_t* shape, const std::size_t* strides)
        : _base(base), _shape(shape), _strides(strides) {}
    view<T, N - 1> operator[](std::size_t i) const {
        return view<T, N - 1>(_base + i * _strides[0], _shape + 1, _strides + 1);
    }
    T* origin() const { return _base; }
    const std::size_t* shape() const { return _shape; }
    std::size_t size() const { return _shape[0]; }
    std::size_t num_elements() const {
        std::size_t n = 1;
        for (std::size_t i = 0; i < N; i++) n *= _shape[i];
        return n;
    }
    template <typename Other>
    const view& operator=(const Other& other) const {
        assert(other.num_elements() == num_elements());
        std::copy_n(other.origin(), num_elements(), _base);
        return *this;
    }
    const view& operator=(const view& other) const {
        std::copy_n(other.origin(), num_elements(), _base);
        return *this;
    }
    view(const view&) = default;
private:
    T* _base;
    const std::size_t* _shape;
    const std::size_t* _strides;
};

template <typename T>
class view<T, 1> {
public:
    typedef T* iterator;
    view(T* base, const std::size_t* shape, const std::size_t* strides)
        : _base(base), _shape(shape), _strides(strides) {}
    T& operator[](std::size_t i) const { return _base[i]; }
    T* origin() const { return _base; }
    T* begin() const { return _base; }
    T* end() const { return _base + _shape[0]; }
    const std::size_t* shape() const { return _shape; }
    std::size_t size() const { return _shape[0]; }
    std::size_t num_elements() const { return _shape[0]; }
    template <typename Other>
    const view& operator=(const Other& other) const {
        assert(other.num_elements() == num_elements());
        std::copy_n(other.origin(), num_elements(), _base);
        return *this;
    }
    const view& operator=(const view& other) const {
        std::copy_n(other.origin(), num_elements(), _base);
        return *this;
    }
    view(const view&) = default;
private:
    T* _base;
    const std::size_t* _shape;
    const std::size_t* _strides;
};
}  // namespace ma_detail

static const ma_detail::extent_gen<0> extents = ma_detail::extent_gen<0>();

template <typename T, std::size_t N>
class multi_array {
public:
    typedef T element;
    typedef T value_type;
    multi_array() { _shape.fill(0); _strides.fill(0); }
    explicit multi_array(const ma_detail::extent_gen<N>& ext) {
        _shape = ext.e;
        _init();
    }
    multi_array(const multi_array&) = default;
    multi_array(multi_array&&) = default;
    multi_array& operator=(multi_array&&) = default;
    multi_array& operator=(const multi_array& other) {
        // boost requires equal shapes for assignment; an empty target adopts the shape
        _shape = other._shape;
        _strides = other._strides;
        _store = other._store;
        return *this;
    }
    template <typename V>
    multi_array(const ma_detail::view<V, N>& v) {
        for (std::size_t i = 0; i < N; i++) _shape[i] = v.shape()[i];
        _init();
        std::copy_n(v.origin(), num_elements(), _store.data());
    }

    T* data() { return _store.data(); }
    const T* data() const { return _store.data(); }
    T* origin() { return _store.data(); }
    const T* origin() const { return _store.data(); }
    std::size_t num_elements() const { return _store.size(); }
    const std::size_t* shape() const { return _shape.data(); }
    std::size_t size() const { return _shape[0]; }

    template <std::size_t M = N>
    typename std::enable_if<(M > 1), ma_detail::view<T, N - 1>>::type
    operator[](std::size_t i) {
        return ma_detail::view<T, N - 1>(_store.data() + i * _strides[0],
                                         _shape.data() + 1, _strides.data() + 1);
    }
    template <std::size_t M = N>
    typename std::enable_if<(M > 1), ma_detail::view<const T, N - 1>>::type
    operator[](std::size_t i) const {
        return ma_detail::view<const T, N - 1>(_store.data() + i * _strides[0],
                                               _shape.data() + 1, _strides.data() + 1);
    }
    template <std::size_t M = N>
    typename std::enable_if<(M == 1), T&>::type operator[](std::size_t i) {
        return _store[i];
    }
    template <std::size_t M = N>
    typename std::enable_if<(M == 1), const T&>::type operator[](std::size_t i) const {
        return _store[i];
    }
    T* begin() { return _store.data(); }
    T* end() { return _store.data() + (N == 1 ? _store.size() : 0); }
    const T* begin() const { return _store.data(); }
    const T* end() const { return _store.data() + (N == 1 ? _store.size() : 0); }

private:
    void _init() {
        std::size_t n = 1;
        for (std::size_t i = N; i-- > 0;) {
            _strides[i] = n;
            n *= _shape[i];
        }
        _store.assign(n, T());
    }
    std::array<std::size_t, N> _shape;
    std::array<std::size_t, N> _strides;
    std::vector<T> _store;
};

}  // namespace boost
