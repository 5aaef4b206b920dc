// Stand-in for Embree's common/math/vec.h (see oracle/ref_shim/README.md): Vec3fa / Vec3f / Vec2f with plain float arithmetic.
#pragma once
#include "math.h"
namespace embree {
struct Vec2f { float x, y; Vec2f() {} Vec2f(float a) : x(a), y(a) {} Vec2f(float a, float b) : x(a), y(b) {} };
struct Vec3f { float x, y, z; Vec3f() {} Vec3f(float a) : x(a), y(a), z(a) {} Vec3f(float a, float b, float c) : x(a), y(b), z(c) {} };
struct alignas(16) Vec3fa {
    float x, y, z; union { float w; int a; unsigned u; };
    Vec3fa() {} Vec3fa(float s) : x(s), y(s), z(s), w(0) {} Vec3fa(float a_, float b_, float c_) : x(a_), y(b_), z(c_), w(0) {}
    Vec3fa(float a_, float b_, float c_, float d_) : x(a_), y(b_), z(c_), w(d_) {}
    Vec3fa(const Vec3f& v) : x(v.x), y(v.y), z(v.z), w(0) {}
    Vec3fa(ZeroTy) : x(0), y(0), z(0), w(0) {} Vec3fa(OneTy) : x(1), y(1), z(1), w(0) {}
    float& operator[](size_t i) { return (&x)[i]; } const float& operator[](size_t i) const { return (&x)[i]; }
};
typedef Vec3fa Vec3ff;
inline Vec3fa operator+(const Vec3fa& a) { return a; }
inline Vec3fa operator-(const Vec3fa& a) { return Vec3fa(-a.x, -a.y, -a.z); }
inline Vec3fa operator+(const Vec3fa& a, const Vec3fa& b) { return Vec3fa(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3fa operator-(const Vec3fa& a, const Vec3fa& b) { return Vec3fa(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3fa operator*(const Vec3fa& a, const Vec3fa& b) { return Vec3fa(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3fa operator*(const Vec3fa& a, float b) { return Vec3fa(a.x * b, a.y * b, a.z * b); }
inline Vec3fa operator*(float a, const Vec3fa& b) { return Vec3fa(a * b.x, a * b.y, a * b.z); }
inline Vec3fa operator/(const Vec3fa& a, const Vec3fa& b) { return Vec3fa(a.x / b.x, a.y / b.y, a.z / b.z); }
inline Vec3fa operator/(const Vec3fa& a, float b) { return Vec3fa(a.x / b, a.y / b, a.z / b); }
inline Vec3fa operator/(float a, const Vec3fa& b) { return Vec3fa(a / b.x, a / b.y, a / b.z); }
inline Vec3fa& operator+=(Vec3fa& a, const Vec3fa& b) { return a = a + b; }
inline Vec3fa& operator-=(Vec3fa& a, const Vec3fa& b) { return a = a - b; }
inline Vec3fa& operator*=(Vec3fa& a, const Vec3fa& b) { return a = a * b; }
inline Vec3fa& operator*=(Vec3fa& a, float b) { return a = a * b; }
inline Vec3fa& operator/=(Vec3fa& a, float b) { return a = a / b; }
inline bool operator==(const Vec3fa& a, const Vec3fa& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
inline bool operator!=(const Vec3fa& a, const Vec3fa& b) { return !(a == b); }
inline float dot(const Vec3fa& a, const Vec3fa& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3fa cross(const Vec3fa& a, const Vec3fa& b) { return Vec3fa(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float sqr_length(const Vec3fa& a) { return dot(a, a); }
inline float length(const Vec3fa& a) { return ::sqrtf(dot(a, a)); }
inline Vec3fa normalize(const Vec3fa& a) { return a / length(a); }
inline float distance(const Vec3fa& a, const Vec3fa& b) { return length(a - b); }
inline Vec3fa min(const Vec3fa& a, const Vec3fa& b) { return Vec3fa(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline Vec3fa max(const Vec3fa& a, const Vec3fa& b) { return Vec3fa(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline Vec3fa abs(const Vec3fa& a) { return Vec3fa(::fabsf(a.x), ::fabsf(a.y), ::fabsf(a.z)); }
inline Vec3fa neg(const Vec3fa& a) { return -a; }
inline float reduce_add(const Vec3fa& a) { return a.x + a.y + a.z; }
inline float reduce_max(const Vec3fa& a) { return max(a.x, max(a.y, a.z)); }
inline float reduce_min(const Vec3fa& a) { return min(a.x, min(a.y, a.z)); }
}  // namespace embree
