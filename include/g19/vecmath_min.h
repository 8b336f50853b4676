// g19/vecmath_min.h -- the few GLM names the interface headers use, for builds
// where the reference's vendored GLM (3rd_party/glm 0.9.8.2) is not on the
// include path. A drop-in build of the reference's main.cpp uses the real GLM;
// these headers pick it up automatically when <glm/glm.hpp> exists.
#pragma once
#include <cmath>
#include <cstddef>

namespace glm {
template <typename T> struct tvec3 {
    union { T x, r; };
    union { T y, g; };
    union { T z, b; };
    tvec3() : x(0), y(0), z(0) {}
    tvec3(T a, T b_, T c) : x(a), y(b_), z(c) {}
    explicit tvec3(T s) : x(s), y(s), z(s) {}
    template <typename U> tvec3(const tvec3<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)) {}
    T& operator[](std::size_t i) { return i == 0 ? x : (i == 1 ? y : z); }
    const T& operator[](std::size_t i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename T> struct tvec2 {
    T x, y;
    tvec2() : x(0), y(0) {}
    tvec2(T a, T b) : x(a), y(b) {}
};
typedef tvec3<double> dvec3;
typedef tvec3<float> vec3;
typedef tvec2<double> dvec2;
template <typename T> tvec3<T> operator+(tvec3<T> a, tvec3<T> b) { return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> tvec3<T> operator-(tvec3<T> a, tvec3<T> b) { return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> tvec3<T> operator-(tvec3<T> a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <typename T> tvec3<T> operator*(tvec3<T> a, T s) { return tvec3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> tvec3<T> operator*(T s, tvec3<T> a) { return tvec3<T>(s * a.x, s * a.y, s * a.z); }
template <typename T> T dot(tvec3<T> a, tvec3<T> b) { T tx = a.x * b.x, ty = a.y * b.y, tz = a.z * b.z; return tx + ty + tz; }
template <typename T> tvec3<T> cross(tvec3<T> a, tvec3<T> b) {
    return tvec3<T>(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
template <typename T> tvec3<T> normalize(tvec3<T> v) { return v * (T(1) / std::sqrt(dot(v, v))); }
template <typename T> T length(tvec3<T> v) { return std::sqrt(dot(v, v)); }
} // namespace glm
