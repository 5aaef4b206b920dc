// Stand-in for Embree's tutorial header common/core/ray.h (see oracle/ref_shim/README.md): a Ray whose memory layout is RTCRayHit's
// (the reference casts Ray* to RTCRayHit*).
#pragma once
#include <embree3/rtcore.h>
#include "../math/vec.h"
namespace embree {
struct alignas(16) Ray {
    Vec3fa org;            // org.w = tnear
    Vec3fa dir;            // dir.w = time
    float tfar; unsigned int mask, id, flags;
    Vec3f Ng; float u, v; unsigned int primID, geomID, instID;
    Ray() {}
    Ray(const Vec3fa& o, const Vec3fa& d, float tnear_ = 0.0f, float tfar_ = std::numeric_limits<float>::infinity(), float time_ = 0.0f, unsigned mask_ = (unsigned)-1,
        unsigned geomID_ = RTC_INVALID_GEOMETRY_ID, unsigned primID_ = RTC_INVALID_GEOMETRY_ID)
        : org(o), dir(d), tfar(tfar_), mask(mask_), id(0), flags(0), Ng(0, 0, 0), u(0), v(0), primID(primID_), geomID(geomID_), instID(RTC_INVALID_GEOMETRY_ID) { org.w = tnear_; dir.w = time_; }
    float& tnear() { return org.w; } float& time() { return dir.w; }
};
static_assert(sizeof(Ray) >= sizeof(RTCRayHit), "Ray must cover RTCRayHit");
static_assert(offsetof(RTCRayHit, hit) == 48, "RTCRayHit layout");
inline void init_Ray(Ray& ray, const Vec3fa& org, const Vec3fa& dir, float tnear, float tfar, float time = 0.0f, unsigned mask = (unsigned)-1,
                     unsigned geomID = RTC_INVALID_GEOMETRY_ID, unsigned primID = RTC_INVALID_GEOMETRY_ID) {
    ray.org = org; ray.org.w = tnear; ray.dir = dir; ray.dir.w = time; ray.tfar = tfar; ray.mask = mask; ray.id = 0; ray.flags = 0;
    ray.Ng = Vec3f(0, 0, 0); ray.u = 0; ray.v = 0; ray.primID = primID; ray.geomID = geomID; ray.instID = RTC_INVALID_GEOMETRY_ID;
}
inline RTCRayHit* RTCRayHit_(Ray& r) { return (RTCRayHit*)&r; }
}  // namespace embree
