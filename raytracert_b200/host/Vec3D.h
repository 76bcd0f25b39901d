// Vec3D<T> / Vec3Df -- host-side 3-vector with the API surface of the reference's Vec3D.h
// (reference: CG_Project/Vec3D.h:56-273, typedefs :291-293), written from scratch.
//
// Arithmetic contract (matters for bit-parity of everything computed on the host, e.g. face normals):
//   * dotProduct is evaluated left to right, a0*b0 + a1*b1 + a2*b2            (Vec3D.h:192-194)
//   * crossProduct component i is a[j]*b[k] - a[k]*b[j]                          (Vec3D.h:185-191)
//   * normalize(): len = (T)sqrt(dot); if len == 0 leave untouched; inv = 1.0f/len; three multiplies
//     (Vec3D.h:142-151) -- a reciprocal-multiply, not a divide, and never rsqrt.
//   * operator/ divides each component by the scalar (Vec3D.h:36-38).
// Host translation units are built with -ffp-contract=off so none of this is fused.
#pragma once
#include <cmath>
#include <cstddef>
#include <iosfwd>
#include <ostream>
#include <istream>

template <typename T>
class Vec3D {
public:
    T p[3];

    Vec3D() : p{T(), T(), T()} {}
    Vec3D(T x, T y, T z) : p{x, y, z} {}
    explicit Vec3D(const T* src) : p{src[0], src[1], src[2]} {}

    T& operator[](int i) { return p[i]; }
    const T& operator[](int i) const { return p[i]; }
    T* pointer() { return p; }
    const T* pointer() const { return p; }

    Vec3D& init(T x, T y, T z) { p[0] = x; p[1] = y; p[2] = z; return *this; }

    Vec3D& operator+=(const Vec3D& o) { for (int i = 0; i < 3; ++i) p[i] += o.p[i]; return *this; }
    Vec3D& operator-=(const Vec3D& o) { for (int i = 0; i < 3; ++i) p[i] -= o.p[i]; return *this; }
    Vec3D& operator*=(const Vec3D& o) { for (int i = 0; i < 3; ++i) p[i] *= o.p[i]; return *this; }
    Vec3D& operator/=(const Vec3D& o) { for (int i = 0; i < 3; ++i) p[i] /= o.p[i]; return *this; }
    Vec3D& operator*=(T s) { for (int i = 0; i < 3; ++i) p[i] *= s; return *this; }
    Vec3D& operator/=(T s) { for (int i = 0; i < 3; ++i) p[i] /= s; return *this; }

    static T dotProduct(const Vec3D& a, const Vec3D& b) { return a.p[0] * b.p[0] + a.p[1] * b.p[1] + a.p[2] * b.p[2]; }
    static Vec3D crossProduct(const Vec3D& a, const Vec3D& b) {
        return Vec3D(a.p[1] * b.p[2] - a.p[2] * b.p[1],
                     a.p[2] * b.p[0] - a.p[0] * b.p[2],
                     a.p[0] * b.p[1] - a.p[1] * b.p[0]);
    }
    T getSquaredLength() const { return dotProduct(*this, *this); }
    T getLength() const { return (T)std::sqrt((double)getSquaredLength()); }
    // Returns the length before normalisation; a zero vector is left as it is.
    T normalize() {
        T len = getLength();
        if (len == 0.0f) return 0;
        T inv = 1.0f / len;
        p[0] *= inv; p[1] *= inv; p[2] *= inv;
        return len;
    }
    static T squaredDistance(const Vec3D& a, const Vec3D& b) { Vec3D d(a.p[0] - b.p[0], a.p[1] - b.p[1], a.p[2] - b.p[2]); return d.getSquaredLength(); }
    static T distance(const Vec3D& a, const Vec3D& b) { Vec3D d(a.p[0] - b.p[0], a.p[1] - b.p[1], a.p[2] - b.p[2]); return d.getLength(); }
    static Vec3D segment(const Vec3D& a, const Vec3D& b) { return Vec3D(b.p[0] - a.p[0], b.p[1] - a.p[1], b.p[2] - a.p[2]); }
    void fromTo(const Vec3D& a, const Vec3D& b) { *this = segment(a, b); }
    static Vec3D interpolate(const Vec3D& u, const Vec3D& v, T alpha) { return u * (1.0f - alpha) + v * alpha; }
    static Vec3D projectOntoVector(const Vec3D& v1, const Vec3D& v2) { return v2 * dotProduct(v1, v2); }
};

template <typename T> inline Vec3D<T> operator+(const Vec3D<T>& a, const Vec3D<T>& b) { return Vec3D<T>(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
template <typename T> inline Vec3D<T> operator-(const Vec3D<T>& a, const Vec3D<T>& b) { return Vec3D<T>(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
template <typename T> inline Vec3D<T> operator-(const Vec3D<T>& a) { return Vec3D<T>(-a[0], -a[1], -a[2]); }
template <typename T> inline Vec3D<T> operator*(const Vec3D<T>& a, const Vec3D<T>& b) { return Vec3D<T>(a[0] * b[0], a[1] * b[1], a[2] * b[2]); }
template <typename T> inline Vec3D<T> operator*(const Vec3D<T>& a, float s) { return Vec3D<T>(a[0] * s, a[1] * s, a[2] * s); }
template <typename T> inline Vec3D<T> operator*(float s, const Vec3D<T>& a) { return Vec3D<T>(a[0] * s, a[1] * s, a[2] * s); }
template <typename T> inline Vec3D<T> operator/(const Vec3D<T>& a, float s) { return Vec3D<T>(a[0] / s, a[1] / s, a[2] / s); }
template <typename T> inline bool operator==(const Vec3D<T>& a, const Vec3D<T>& b) { return a[0] == b[0] && a[1] == b[1] && a[2] == b[2]; }
template <typename T> inline bool operator!=(const Vec3D<T>& a, const Vec3D<T>& b) { return !(a == b); }
template <typename T> inline std::ostream& operator<<(std::ostream& os, const Vec3D<T>& v) { return os << v[0] << " " << v[1] << " " << v[2]; }
template <typename T> inline std::istream& operator>>(std::istream& is, Vec3D<T>& v) { return is >> v[0] >> v[1] >> v[2]; }

typedef Vec3D<float> Vec3Df;
typedef Vec3D<double> Vec3Dd;
typedef Vec3D<int> Vec3Di;
