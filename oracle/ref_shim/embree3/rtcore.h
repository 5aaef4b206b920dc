// Stand-in for <embree3/rtcore.h> (see oracle/ref_shim/README.md).  Only what the reference renderer calls:
// device/scene/geometry life cycle for ONE triangle mesh and rtcIntersect1M (closest hit).  Written from scratch.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include <limits>

#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)
enum RTCGeometryType { RTC_GEOMETRY_TYPE_TRIANGLE = 0 };
enum RTCBufferType { RTC_BUFFER_TYPE_INDEX = 0, RTC_BUFFER_TYPE_VERTEX = 1 };
enum RTCFormat { RTC_FORMAT_UINT3 = 0x5003, RTC_FORMAT_FLOAT3 = 0x9003 };
enum RTCBuildQuality { RTC_BUILD_QUALITY_LOW = 0, RTC_BUILD_QUALITY_MEDIUM = 1, RTC_BUILD_QUALITY_HIGH = 2 };
enum RTCIntersectContextFlags { RTC_INTERSECT_CONTEXT_FLAG_NONE = 0, RTC_INTERSECT_CONTEXT_FLAG_INCOHERENT = 0, RTC_INTERSECT_CONTEXT_FLAG_COHERENT = 1 };

struct RTCRay { float org_x, org_y, org_z, tnear, dir_x, dir_y, dir_z, time, tfar; unsigned int mask, id, flags; };
struct RTCHit { float Ng_x, Ng_y, Ng_z, u, v; unsigned int primID, geomID, instID[1]; };
struct RTCRayHit { RTCRay ray; RTCHit hit; };
struct RTCIntersectContext { RTCIntersectContextFlags flags; void* filter; unsigned int instID[1]; };
inline void rtcInitIntersectContext(RTCIntersectContext* c) { c->flags = RTC_INTERSECT_CONTEXT_FLAG_NONE; c->filter = nullptr; c->instID[0] = RTC_INVALID_GEOMETRY_ID; }

namespace refshim {
struct Geometry {
    int refs = 1; std::vector<unsigned char> vbuf, ibuf; size_t vstride = 0, istride = 0, nv = 0, nt = 0;
};
struct Node { float lo[3], hi[3]; int left, right, first, count; };
struct Scene {
    Geometry* geom = nullptr; std::vector<Node> nodes; std::vector<int> order;
    const float* vert(size_t i) const { return (const float*)(geom->vbuf.data() + i * geom->vstride); }
    const int* tri(size_t i) const { return (const int*)(geom->ibuf.data() + i * geom->istride); }
    void bounds(int t, float lo[3], float hi[3]) const {
        const int* ix = tri(t);
        for (int a = 0; a < 3; ++a) { lo[a] = std::numeric_limits<float>::max(); hi[a] = -lo[a]; }
        for (int k = 0; k < 3; ++k) { const float* p = vert(ix[k]); for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    }
    int build(int first, int count) {
        Node n; for (int a = 0; a < 3; ++a) { n.lo[a] = std::numeric_limits<float>::max(); n.hi[a] = -n.lo[a]; }
        std::vector<float> cen(count);
        for (int i = 0; i < count; ++i) { float lo[3], hi[3]; bounds(order[first + i], lo, hi); for (int a = 0; a < 3; ++a) { n.lo[a] = std::min(n.lo[a], lo[a]); n.hi[a] = std::max(n.hi[a], hi[a]); } }
        n.left = n.right = -1; n.first = first; n.count = count;
        int id = (int)nodes.size(); nodes.push_back(n);
        if (count > 4) {
            int ax = 0; for (int a = 1; a < 3; ++a) if (n.hi[a] - n.lo[a] > n.hi[ax] - n.lo[ax]) ax = a;
            int mid = count / 2;
            std::nth_element(order.begin() + first, order.begin() + first + mid, order.begin() + first + count, [&](int x, int y) {
                float lx[3], hx[3], ly[3], hy[3]; bounds(x, lx, hx); bounds(y, ly, hy); return lx[ax] + hx[ax] < ly[ax] + hy[ax]; });
            int l = build(first, mid), r = build(first + mid, count - mid);
            nodes[id].left = l; nodes[id].right = r; nodes[id].count = 0;
        }
        return id;
    }
    void commit() { nodes.clear(); order.resize(geom ? geom->nt : 0); for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i; if (!order.empty()) build(0, (int)order.size()); }
    // closest hit; double-precision Moeller-Trumbore, u/v are the weights of the 2nd/3rd vertex as in Embree
    void test(int t, const double o[3], const double d[3], double tnear, double& best, RTCRayHit& rh) const {
        const int* ix = tri(t); const float *a = vert(ix[0]), *b = vert(ix[1]), *c = vert(ix[2]);
        double e1[3], e2[3], p[3], s[3], q[3];
        for (int k = 0; k < 3; ++k) { e1[k] = (double)b[k] - a[k]; e2[k] = (double)c[k] - a[k]; s[k] = o[k] - a[k]; }
        p[0] = d[1] * e2[2] - d[2] * e2[1]; p[1] = d[2] * e2[0] - d[0] * e2[2]; p[2] = d[0] * e2[1] - d[1] * e2[0];
        double det = e1[0] * p[0] + e1[1] * p[1] + e1[2] * p[2];
        if (det == 0.0) return;
        double inv = 1.0 / det, u = (s[0] * p[0] + s[1] * p[1] + s[2] * p[2]) * inv;
        if (u < 0.0 || u > 1.0) return;
        q[0] = s[1] * e1[2] - s[2] * e1[1]; q[1] = s[2] * e1[0] - s[0] * e1[2]; q[2] = s[0] * e1[1] - s[1] * e1[0];
        double v = (d[0] * q[0] + d[1] * q[1] + d[2] * q[2]) * inv;
        if (v < 0.0 || u + v > 1.0) return;
        double tt = (e2[0] * q[0] + e2[1] * q[1] + e2[2] * q[2]) * inv;
        if (tt < tnear || tt >= best) return;
        best = tt; rh.ray.tfar = (float)tt; rh.hit.u = (float)u; rh.hit.v = (float)v; rh.hit.primID = (unsigned)t; rh.hit.geomID = 0;
        rh.hit.Ng_x = (float)(e1[1] * e2[2] - e1[2] * e2[1]); rh.hit.Ng_y = (float)(e1[2] * e2[0] - e1[0] * e2[2]); rh.hit.Ng_z = (float)(e1[0] * e2[1] - e1[1] * e2[0]);
    }
    void intersect(RTCRayHit& rh) const {
        if (nodes.empty()) return;
        const double o[3] = {rh.ray.org_x, rh.ray.org_y, rh.ray.org_z}, d[3] = {rh.ray.dir_x, rh.ray.dir_y, rh.ray.dir_z};
        double best = rh.ray.tfar, tnear = rh.ray.tnear;
        int stack[128], sp = 0; stack[sp++] = 0;
        while (sp) {
            const Node& n = nodes[stack[--sp]];
            double t0 = tnear, t1 = best; bool miss = false;
            for (int a = 0; a < 3 && !miss; ++a) {
                double lo = (double)n.lo[a] - 1e-6, hi = (double)n.hi[a] + 1e-6;
                if (d[a] == 0.0) { if (o[a] < lo || o[a] > hi) miss = true; continue; }
                double ta = (lo - o[a]) / d[a], tb = (hi - o[a]) / d[a]; if (ta > tb) std::swap(ta, tb);
                t0 = std::max(t0, ta); t1 = std::min(t1, tb); if (t0 > t1) miss = true;
            }
            if (miss) continue;
            if (n.left < 0) { for (int i = 0; i < n.count; ++i) test(order[n.first + i], o, d, tnear, best, rh); }
            else { stack[sp++] = n.left; stack[sp++] = n.right; }
        }
    }
};
struct Device { int refs = 1; };
}  // namespace refshim

typedef refshim::Device* RTCDevice;
typedef refshim::Scene* RTCScene;
typedef refshim::Geometry* RTCGeometry;

inline RTCDevice rtcNewDevice(const char*) { return new refshim::Device(); }
inline void rtcReleaseDevice(RTCDevice d) { delete d; }
inline RTCScene rtcNewScene(RTCDevice) { return new refshim::Scene(); }
inline void rtcReleaseScene(RTCScene s) { if (s) { if (s->geom && --s->geom->refs == 0) delete s->geom; delete s; } }
inline RTCGeometry rtcNewGeometry(RTCDevice, RTCGeometryType) { return new refshim::Geometry(); }
inline void* rtcSetNewGeometryBuffer(RTCGeometry g, RTCBufferType type, unsigned, RTCFormat, size_t stride, size_t count) {
    if (type == RTC_BUFFER_TYPE_VERTEX) { g->vbuf.assign(stride * count + 16, 0); g->vstride = stride; g->nv = count; return g->vbuf.data(); }
    g->ibuf.assign(stride * count + 16, 0); g->istride = stride; g->nt = count; return g->ibuf.data();
}
inline void rtcCommitGeometry(RTCGeometry) {}
inline unsigned rtcAttachGeometry(RTCScene s, RTCGeometry g) { s->geom = g; ++g->refs; return 0; }
inline void rtcReleaseGeometry(RTCGeometry g) { if (--g->refs == 0) delete g; }
inline void rtcSetSceneBuildQuality(RTCScene, RTCBuildQuality) {}
inline void rtcCommitScene(RTCScene s) { s->commit(); }
__attribute__((noinline)) inline void rtcIntersect1(RTCScene s, RTCIntersectContext*, RTCRayHit* rh) { s->intersect(*rh); }
__attribute__((noinline)) inline void rtcIntersect1M(RTCScene s, RTCIntersectContext*, RTCRayHit* rh, unsigned int M, size_t byteStride) {
    for (unsigned int i = 0; i < M; ++i) s->intersect(*(RTCRayHit*)((char*)rh + i * byteStride));
}
