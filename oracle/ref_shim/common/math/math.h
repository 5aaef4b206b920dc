// Stand-in for Embree's common/math/math.h (see oracle/ref_shim/README.md): scalar helpers in namespace embree.
#pragma once
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <cstddef>
#include <limits>
#include <algorithm>
namespace embree {
struct ZeroTy { operator float() const { return 0.0f; } operator double() const { return 0.0; } operator int() const { return 0; } };
struct OneTy { operator float() const { return 1.0f; } operator double() const { return 1.0; } operator int() const { return 1; } };
struct PosInfTy { operator float() const { return std::numeric_limits<float>::infinity(); } operator double() const { return std::numeric_limits<double>::infinity(); } };
struct NegInfTy { operator float() const { return -std::numeric_limits<float>::infinity(); } operator double() const { return -std::numeric_limits<double>::infinity(); } };
static const ZeroTy zero = ZeroTy(); static const OneTy one = OneTy(); static const PosInfTy inf = PosInfTy(); static const PosInfTy pos_inf = PosInfTy(); static const NegInfTy neg_inf = NegInfTy();
inline float sqrt(float x) { return ::sqrtf(x); } inline double sqrt(double x) { return ::sqrt(x); }
inline float floor(float x) { return ::floorf(x); } inline double floor(double x) { return ::floor(x); }
inline float ceil(float x) { return ::ceilf(x); } inline double ceil(double x) { return ::ceil(x); }
inline float exp(float x) { return ::expf(x); } inline double exp(double x) { return ::exp(x); }
inline float log(float x) { return ::logf(x); } inline double log(double x) { return ::log(x); }
inline float pow(float x, float y) { return ::powf(x, y); } inline double pow(double x, double y) { return ::pow(x, y); }
inline float pow(float x, int y) { return ::powf(x, (float)y); } inline double pow(double x, int y) { return ::pow(x, (double)y); }
inline float abs(float x) { return ::fabsf(x); } inline double abs(double x) { return ::fabs(x); } inline int abs(int x) { return x < 0 ? -x : x; }
inline float sqr(float x) { return x * x; } inline double sqr(double x) { return x * x; }
inline float rcp(float x) { return 1.0f / x; } inline double rcp(double x) { return 1.0 / x; }
inline float rsqrt(float x) { return 1.0f / ::sqrtf(x); }
inline float sin(float x) { return ::sinf(x); } inline double sin(double x) { return ::sin(x); }
inline float cos(float x) { return ::cosf(x); } inline double cos(double x) { return ::cos(x); }
inline float tan(float x) { return ::tanf(x); } inline double tan(double x) { return ::tan(x); }
inline float acos(float x) { return ::acosf(x); } inline double acos(double x) { return ::acos(x); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
inline float min(float a, float b) { return a < b ? a : b; } inline double min(double a, double b) { return a < b ? a : b; } inline int min(int a, int b) { return a < b ? a : b; }
inline float max(float a, float b) { return a < b ? b : a; } inline double max(double a, double b) { return a < b ? b : a; } inline int max(int a, int b) { return a < b ? b : a; }
inline double min(double a, float b) { return min(a, (double)b); } inline double min(float a, double b) { return min((double)a, b); }
inline double max(double a, float b) { return max(a, (double)b); } inline double max(float a, double b) { return max((double)a, b); }
inline float clamp(float x, float lo = 0.0f, float hi = 1.0f) { return max(lo, min(x, hi)); }
}  // namespace embree
