// nlos_core.cuh — per-thread building blocks of the B200 transient renderer.
//
// Everything here is a pure function of its arguments (no threadIdx, no shared memory), so the same
// code can be instantiated inside the sm_100a kernels and, for logic checks only, inside a host test
// harness (tests/emul).  The library is compiled with -fmad=false: an FMA happens exactly where fmaf()
// is written, '/' and sqrtf() are IEEE — that is what makes per-sample visibility and bin indices
// bit-identical to the CPU oracle (DESIGN.md "Arithmetic contract").
//
// Reference formulas (relative to /root/reference/transient_rendering_cython/):
//   sampling / visibility / binning   smoothed_transient/transient_and_gradient.cpp:184-232
//   gradient terms                    smoothed_transient/transient_and_gradient.cpp:944-1001
//   GGX                               ggx/ggx_confocal.cpp:13-231, ggx/transient_and_gradient.cpp:756-780
//   bits -> [0,1)                     smoothed_transient/rng_sse.h:33-42
#pragma once
#include <cstdint>
#include <cmath>
#include <vector_types.h>

#if defined(__CUDACC__)
#define NLOS_HD __host__ __device__ __forceinline__
#else
#define NLOS_HD inline
#endif

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

namespace nlos {

// ------------------------------------------------------------------ tiny vector algebra (pinned op order)
struct f3 { float x, y, z; };
NLOS_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
NLOS_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
NLOS_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
NLOS_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
NLOS_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
NLOS_HD f3 operator*(float s, f3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
NLOS_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
NLOS_HD float dot3(f3 a, f3 b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }
NLOS_HD f3 cross3(f3 a, f3 b) {
  return mk3(fmaf(a.y, b.z, -(a.z * b.y)), fmaf(a.z, b.x, -(a.x * b.z)), fmaf(a.x, b.y, -(a.y * b.x)));
}
NLOS_HD float len3(f3 a) { return sqrtf(dot3(a, a)); }
NLOS_HD f3 blend3(float u, f3 a, float v, f3 b, float w, f3 c) {
  return mk3(fmaf(w, c.x, fmaf(v, b.x, u * a.x)), fmaf(w, c.y, fmaf(v, b.y, u * a.y)), fmaf(w, c.z, fmaf(v, b.z, u * a.z)));
}
NLOS_HD float blend1(float u, float a, float v, float b, float w, float c) { return fmaf(w, c, fmaf(v, b, u * a)); }
NLOS_HD f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }

NLOS_HD int f2i(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_int(f);
#else
  union { float f; int i; } c; c.f = f; return c.i;
#endif
}
NLOS_HD float i2f(int i) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(i);
#else
  union { float f; int i; } c; c.i = i; return c.f;
#endif
}

// ------------------------------------------------------------------ counter-based RNG: Philox4x32-10
NLOS_HD void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
NLOS_HD float bits_to_unit(uint32_t x) { return i2f((int)((x >> 9) | 0x3f800000u)) - 1.0f; }

// the two uniforms of sample k of (global source index src, original triangle index tri)
NLOS_HD void sample_ST(uint64_t seed, int64_t src, int tri, int k, float& S, float& T) {
  uint32_t o[4];
  philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)tri, (uint32_t)src, (uint32_t)(k >> 1), (uint32_t)((uint64_t)src >> 32), o);
  if (k & 1) { S = bits_to_unit(o[2]); T = bits_to_unit(o[3]); } else { S = bits_to_unit(o[0]); T = bits_to_unit(o[1]); }
}

// ------------------------------------------------------------------ triangle records (device layout, 64 B each)
// TraceTri (4 x float4): [v0.xyz | prim] [e1.xyz | -] [e2.xyz | -] [Ng.xyz | -]   e1=v0-v1, e2=v2-v0, Ng=cross(e2,e1)
// ShadeTri (4 x float4): [v1.xyz | A] [v2.xyz | nf.x] [v3.xyz | nf.y] [nf.z | i1 | i2 | i3]
struct TriRec { f3 v0, e1, e2, Ng; };
NLOS_HD TriRec make_tri(f3 a, f3 b, f3 c) { TriRec t; t.v0 = a; t.e1 = a - b; t.e2 = c - a; t.Ng = cross3(t.e2, t.e1); return t; }

// Moeller-Trumbore in Embree's edge form; division only on a valid hit.  p = (1-u-v) v0 + u v1 + v v2.
NLOS_HD bool isect(const TriRec& tr, f3 o, f3 d, float& t, float& u, float& v) {
  const f3 C = tr.v0 - o;
  const f3 R = cross3(C, d);
  const float den = dot3(tr.Ng, d);
  const float absDen = fabsf(den);
  const float sgn = den < 0.0f ? -1.0f : 1.0f;
  const float U = dot3(R, tr.e2) * sgn;
  const float V = dot3(R, tr.e1) * sgn;
  const float T = dot3(tr.Ng, C) * sgn;
  if (!(den != 0.0f) || !(U >= 0.0f) || !(V >= 0.0f) || !(U + V <= absDen) || !(T > 0.0f)) return false;
  t = T / absDen; u = U / absDen; v = V / absDen;
  return true;
}

// ------------------------------------------------------------------ BVH node (64 B): two child boxes + links
// a = (lo0.x lo0.y lo0.z hi0.x)  b = (hi0.y hi0.z lo1.x lo1.y)  c = (lo1.z hi1.x hi1.y hi1.z)
// d = (ref0, ref1, -, -): ref >= 0 is an internal node index, ref < 0 encodes a leaf run ~((first << 3) | (count-1))
struct BvhNode { float4 a, b, c; int4 d; };
NLOS_HD int leaf_ref(int first, int count) { return ~((first << 3) | (count - 1)); }
NLOS_HD int leaf_first(int ref) { return (~ref) >> 3; }
NLOS_HD int leaf_count(int ref) { return ((~ref) & 7) + 1; }

#ifndef NLOS_LEAFMAX
#define NLOS_LEAFMAX 4
#endif
constexpr int kLeafMax = NLOS_LEAFMAX;      // a subtree with <= kLeafMax (<= 8) triangles is tested linearly
constexpr int kStack = 64;       // >= depth of a 62-bit-key LBVH

NLOS_HD float safe_rcp(float x) {
  const float tiny = 1e-20f;
  if (fabsf(x) < tiny) x = (f2i(x) < 0) ? -tiny : tiny;
  return 1.0f / x;
}

struct Ray { f3 o, d, id, oid; };
NLOS_HD Ray make_ray(f3 o, f3 d) {
  Ray r; r.o = o; r.d = d; r.id = mk3(safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z));
  r.oid = mk3(o.x * r.id.x, o.y * r.id.y, o.z * r.id.z); return r;
}
// slab test against a (padded) box, clipped to [0, tlim]; returns entry distance
NLOS_HD bool slab(const Ray& r, float lox, float loy, float loz, float hix, float hiy, float hiz, float tlim, float& tn) {
  const float tx1 = fmaf(lox, r.id.x, -r.oid.x), tx2 = fmaf(hix, r.id.x, -r.oid.x);
  const float ty1 = fmaf(loy, r.id.y, -r.oid.y), ty2 = fmaf(hiy, r.id.y, -r.oid.y);
  const float tz1 = fmaf(loz, r.id.z, -r.oid.z), tz2 = fmaf(hiz, r.id.z, -r.oid.z);
  const float tmn = fmaxf(fmaxf(fminf(tx1, tx2), fminf(ty1, ty2)), fmaxf(fminf(tz1, tz2), 0.0f));
  const float tmx = fminf(fminf(fmaxf(tx1, tx2), fmaxf(ty1, ty2)), fminf(fmaxf(tz1, tz2), tlim));
  tn = tmn;
  return tmn <= tmx;
}

#if defined(__CUDA_ARCH__)
#define NLOS_LDG4(p) __ldg(p)
#else
#define NLOS_LDG4(p) (*(p))
#endif

// 32-byte read-only load (one LDG.E.256 on sm_100a: half the L1 wavefronts of two LDG.128); p must be 32-byte aligned
NLOS_HD void ld256(const float4* __restrict__ p, float4& a, float4& b) {
#if defined(__CUDA_ARCH__)
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p));
#else
  a = p[0]; b = p[1];
#endif
}

// does triangle j (sorted index) beat (t_self, prim_self) in the lexicographic nearest-hit order?
NLOS_HD bool tri_occludes(const float4* __restrict__ ttris, int j, const Ray& r, float t_self, int prim_self) {
  float4 q0, q1, q2, q3;
  ld256(ttris + 4 * (size_t)j, q0, q1);
  ld256(ttris + 4 * (size_t)j + 2, q2, q3);
  const int prim = f2i(q0.w);
  if (prim == prim_self) return false;
  TriRec tr; tr.v0 = xyz(q0); tr.e1 = xyz(q1); tr.e2 = xyz(q2); tr.Ng = xyz(q3);
  float t, u, v;
  if (!isect(tr, r.o, r.d, t, u, v)) return false;
  return t < t_self || (t == t_self && prim < prim_self);
}

// Same predicate as tri_occludes() with one combined validity test and the IEEE division only when the outcome is not already
// decided by a 4-ulp bracket around T = t_self * |den| (the division result is what the contract compares, so the
// undecided sliver still divides; logic only, no arithmetic of isect() is changed).
NLOS_HD bool tri_occludes_od(const float4* __restrict__ ttris, int j, f3 ro, f3 rd, float t_self, int prim_self) {
  float4 q0, q1, q2, q3;
  ld256(ttris + 4 * (size_t)j, q0, q1);
  ld256(ttris + 4 * (size_t)j + 2, q2, q3);
  const int prim = f2i(q0.w);
  const f3 C = xyz(q0) - ro;
  const f3 R = cross3(C, rd);
  const f3 Ng = xyz(q3);
  const float den = dot3(Ng, rd);
  const float absDen = fabsf(den);
  const float sgn = den < 0.0f ? -1.0f : 1.0f;
  const float U = dot3(R, xyz(q2)) * sgn;
  const float V = dot3(R, xyz(q1)) * sgn;
  const float T = dot3(Ng, C) * sgn;
  const bool valid = (den != 0.0f) & (U >= 0.0f) & (V >= 0.0f) & (U + V <= absDen) & (T > 0.0f) & (prim != prim_self);
  if (!valid) return false;
  const float ref = absDen * t_self;
  if (T < ref * 0.99999952f) return true;          // t = T/|den| is certainly < t_self
  if (T > ref * 1.00000048f) return false;         // certainly > t_self
  const float t = T / absDen;
  return t < t_self || (t == t_self && prim < prim_self);
}
NLOS_HD bool tri_occludes_fast(const float4* __restrict__ ttris, int j, const Ray& r, float t_self, int prim_self) {
  return tri_occludes_od(ttris, j, r.o, r.d, t_self, prim_self);
}

// Any-hit query equivalent to "nearest hit (min t, ties -> lowest prim) is NOT prim_self".
// num_internal = F-1 Karras nodes (root = 0); when the whole mesh is one leaf run, root_count = F.
NLOS_HD bool occluded(const BvhNode* __restrict__ nodes, const float4* __restrict__ ttris, int root_count,
                      const Ray& r, float t_self, int prim_self, uint32_t* n_box = nullptr, uint32_t* n_tri = nullptr) {
  if (root_count > 0) {
    for (int j = 0; j < root_count; ++j) { if (n_tri) ++*n_tri; if (tri_occludes(ttris, j, r, t_self, prim_self)) return true; }
    return false;
  }
  const float tlim = t_self * 1.000001f;        // a tie on t must still be reachable
  int stack[kStack]; int sp = 0; int node = 0;
  while (true) {
    const float4 a = NLOS_LDG4(&nodes[node].a), b = NLOS_LDG4(&nodes[node].b), c = NLOS_LDG4(&nodes[node].c);
    const int4 d = *reinterpret_cast<const int4*>(&nodes[node].d);
    float t0, t1;
    bool h0 = slab(r, a.x, a.y, a.z, a.w, b.x, b.y, tlim, t0);
    bool h1 = slab(r, b.z, b.w, c.x, c.y, c.z, c.w, tlim, t1);
    if (n_box) *n_box += 2;
    if (h0 && d.x < 0) { const int f0 = leaf_first(d.x), c0 = leaf_count(d.x); for (int j = 0; j < c0; ++j) { if (n_tri) ++*n_tri; if (tri_occludes(ttris, f0 + j, r, t_self, prim_self)) return true; } h0 = false; }
    if (h1 && d.y < 0) { const int f1 = leaf_first(d.y), c1 = leaf_count(d.y); for (int j = 0; j < c1; ++j) { if (n_tri) ++*n_tri; if (tri_occludes(ttris, f1 + j, r, t_self, prim_self)) return true; } h1 = false; }
    if (h0 && h1) {
      const bool first0 = t0 <= t1;
      stack[sp++] = first0 ? d.y : d.x; node = first0 ? d.x : d.y;
    } else if (h0) node = d.x;
    else if (h1) node = d.y;
    else { if (sp == 0) return false; node = stack[--sp]; }
  }
}

// Nearest hit over the whole mesh: lexicographic minimum of (t, primitive index) over all float-valid intersections —
// the restatement of rtcIntersect1 used by the embree_intersector API (embree_intersector/c_embree_intersector.cpp:20-46).
struct HitRec { int prim; float t, u, v; };
NLOS_HD void hit_run(const float4* __restrict__ ttris, int first, int cnt, const Ray& r, HitRec& best) {
  for (int j = 0; j < cnt; ++j) {
    float4 q0, q1, q2, q3;
    ld256(ttris + 4 * (size_t)(first + j), q0, q1);
    ld256(ttris + 4 * (size_t)(first + j) + 2, q2, q3);
    TriRec tr; tr.v0 = xyz(q0); tr.e1 = xyz(q1); tr.e2 = xyz(q2); tr.Ng = xyz(q3);
    float t, u, v;
    if (!isect(tr, r.o, r.d, t, u, v)) continue;
    const int prim = f2i(q0.w);
    if (t < best.t || (t == best.t && prim < best.prim)) { best.prim = prim; best.t = t; best.u = u; best.v = v; }
  }
}
NLOS_HD HitRec nearest_hit(const BvhNode* __restrict__ nodes, const float4* __restrict__ ttris, int root_count, const Ray& r) {
  HitRec best; best.prim = -1; best.t = 3.0e38f; best.u = best.v = 0.f;
  if (root_count > 0) { hit_run(ttris, 0, root_count, r, best); return best; }
  int stack[kStack]; int sp = 0; int node = 0;
  while (true) {
    const float4 a = NLOS_LDG4(&nodes[node].a), b = NLOS_LDG4(&nodes[node].b), c = NLOS_LDG4(&nodes[node].c);
    const int4 d = *reinterpret_cast<const int4*>(&nodes[node].d);
    const float tlim = best.t * 1.000001f;
    float t0, t1;
    bool h0 = slab(r, a.x, a.y, a.z, a.w, b.x, b.y, tlim, t0);
    bool h1 = slab(r, b.z, b.w, c.x, c.y, c.z, c.w, tlim, t1);
    if (h0 && d.x < 0) { hit_run(ttris, leaf_first(d.x), leaf_count(d.x), r, best); h0 = false; }
    if (h1 && d.y < 0) { hit_run(ttris, leaf_first(d.y), leaf_count(d.y), r, best); h1 = false; }
    if (h0 && h1) { const bool first0 = t0 <= t1; stack[sp++] = first0 ? d.y : d.x; node = first0 ? d.x : d.y; }
    else if (h0) node = d.x;
    else if (h1) node = d.y;
    else {
      // nodes pushed earlier may have become farther than the best hit found since: re-test on pop is not needed for
      // correctness (their triangles simply lose), only costs time
      if (sp == 0) return best;
      node = stack[--sp];
    }
  }
}

// "while-while" variant used by the hot forward kernel: every lane first walks internal nodes until it holds a leaf
// run (or is done), then all lanes test their leaf triangles together — the warp executes box tests with box tests and
// triangle tests with triangle tests instead of interleaving them per lane.  Same answer as occluded().
constexpr int kSentinel = 0x7fffffff;

NLOS_HD bool occluded_ww(const BvhNode* __restrict__ nodes, const float4* __restrict__ ttris, int root_count,
                         const Ray& r, float t_self, int prim_self) {
  const float tlim = t_self * 1.000001f;
  int stack[kStack]; int sp = 0;
  int cur = root_count > 0 ? leaf_ref(0, root_count) : 0;
  while (cur != kSentinel) {
    while ((unsigned)cur < (unsigned)kSentinel) {                    // internal node
      const float4 a = NLOS_LDG4(&nodes[cur].a), b = NLOS_LDG4(&nodes[cur].b), c = NLOS_LDG4(&nodes[cur].c);
      const int4 d = *reinterpret_cast<const int4*>(&nodes[cur].d);
      float t0, t1;
      const bool h0 = slab(r, a.x, a.y, a.z, a.w, b.x, b.y, tlim, t0);
      const bool h1 = slab(r, b.z, b.w, c.x, c.y, c.z, c.w, tlim, t1);
      const int r0 = d.x, r1 = d.y;
      if (h0 && h1) {
        const bool first0 = t0 <= t1;
        stack[sp++] = first0 ? r1 : r0; cur = first0 ? r0 : r1;
      } else if (h0) cur = r0;
      else if (h1) cur = r1;
      else cur = sp ? stack[--sp] : kSentinel;
    }
    while (cur < 0) {                                                // leaf run of 1..4 triangles
      const int first = leaf_first(cur), cnt = leaf_count(cur);
      for (int j = 0; j < cnt; ++j) if (tri_occludes(ttris, first + j, r, t_self, prim_self)) return true;
      cur = sp ? stack[--sp] : kSentinel;
    }
  }
  return false;
}

// ------------------------------------------------------------------ per-source perspective grid (forward pass, DESIGN.md "K1g")
// Every ray of one wall point o leaves the same origin, and (outside the first-generation mode) a traced ray has n_o.d > 0, so
// seen from o the scene is a 2-D picture:  u = (x-o).a / (x-o).n,  v = (x-o).b / (x-o).n  with (a, b, n) an orthonormal frame
// around the wall normal.  A ray is ONE point of that picture; a triangle can only be hit by rays whose point lies inside its
// projected bounding rectangle (padded for float round-off).  The rectangles are binned into a G x G grid per source; a ray
// scans the list of its cell with a packed integer pre-check and runs the exact Moeller-Trumbore test (tri_occludes_fast) on
// the survivors — the visibility ANSWER is still the brute-force definition, the grid only decides which triangles are tried.
//
// Coordinates are quantised to Q = G * 128 steps per axis (7 sub-cell bits).  An entry holds the triangle's rectangle clipped to
// the cell, in sub-cell units, as four guarded bytes  E = [0x80-lo_u | 0x80+hi_u | 0x80-lo_v | 0x80+hi_v];  the ray holds
// R = su - (su<<8) + (sv<<16) - (sv<<24) (mod 2^32).  Every byte lane of E + R stays within [1,255] (no carry between lanes), and
// bit 7 of the four lanes says su >= lo_u, su <= hi_u, sv >= lo_v, sv <= hi_v: the whole 2-D containment test is one add, one
// and, one compare.
struct PGridFrame {
  f3 o, a, b, n;            // origin, projection frame (n = unit wall normal)
  float u0, v0, su, sv;     // uq = clamp(floor((u - u0) * su), 0, Q-1)
  int G;                    // cells per axis (0: no grid for this source -> per-ray BVH query)
  float qmax;               // Q - 1 as float
};
constexpr int kPgSub = 7;                      // sub-cell bits
constexpr unsigned kPgMask = 0x80808080u;
constexpr float kPgPad = 1.0e-5f;              // relative padding of projected coordinates (~30x the float error of the projection)

NLOS_HD float pg_rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
NLOS_HD void pg_make_axes(f3 n_raw, f3& a, f3& b, f3& n, bool& ok) {
  const float l = len3(n_raw);
  ok = l > 0.0f && l < 3.0e38f;
  n = ok ? n_raw * pg_rcp(l) : mk3(0.f, 0.f, 1.f);
  const float ax = fabsf(n.x), ay = fabsf(n.y), az = fabsf(n.z);
  f3 e = (ax <= ay && ax <= az) ? mk3(1.f, 0.f, 0.f) : (ay <= az ? mk3(0.f, 1.f, 0.f) : mk3(0.f, 0.f, 1.f));
  const float k = dot3(e, n);
  a = e - n * k; a = a * pg_rcp(len3(a));
  b = cross3(n, a);
}
// projection of a scene point: picture coordinates (u,v), depth z along n (grid mode needs z >= zmin > 0 for every vertex) and
// m = relative round-off margin of this projection (the padding of a coordinate c is m * (1 + |c|))
NLOS_HD void pg_project(const PGridFrame& g, f3 x, float& u, float& v, float& m, float& z) {
  const f3 rel = x - g.o;
  z = dot3(rel, g.n);
  const float rz = pg_rcp(z);
  u = dot3(rel, g.a) * rz; v = dot3(rel, g.b) * rz;
  m = kPgPad * ((fabsf(rel.x) + fabsf(rel.y) + fabsf(rel.z)) * rz);
}
NLOS_HD int pg_quant(float x, float x0, float s, float qmax) {   // monotone in x
  float t = (x - x0) * s;
  t = fminf(fmaxf(t, 0.0f), qmax);
  return (int)t;
}
// quantised rectangle [a0,a1] x [b0,b1] of a triangle from the picture coordinates of its vertices and the source-wide padding
// (pad_u, pad_v >= the padding of every single vertex)
NLOS_HD void pg_tri_rect(const PGridFrame& g, float u1, float v1, float u2, float v2, float u3, float v3, float pad_u, float pad_v,
                         int& a0, int& a1, int& b0, int& b1) {
  const float ulo = fminf(u1, fminf(u2, u3)) - pad_u, uhi = fmaxf(u1, fmaxf(u2, u3)) + pad_u;
  const float vlo = fminf(v1, fminf(v2, v3)) - pad_v, vhi = fmaxf(v1, fmaxf(v2, v3)) + pad_v;
  a0 = pg_quant(ulo, g.u0, g.su, g.qmax); a1 = pg_quant(uhi, g.u0, g.su, g.qmax);
  b0 = pg_quant(vlo, g.v0, g.sv, g.qmax); b1 = pg_quant(vhi, g.v0, g.sv, g.qmax);
}
// entry word of a rectangle clipped to cell (cx, cy)
NLOS_HD unsigned pg_entry(int uq0, int uq1, int vq0, int vq1, int cx, int cy) {
  const int bu = cx << kPgSub, bv = cy << kPgSub;
  int lu = uq0 - bu, hu = uq1 - bu, lv = vq0 - bv, hv = vq1 - bv;
  lu = lu < 0 ? 0 : lu; hu = hu > 127 ? 127 : hu; lv = lv < 0 ? 0 : lv; hv = hv > 127 ? 127 : hv;
  return (unsigned)(0x80 - lu) | ((unsigned)(0x80 + hu) << 8) | ((unsigned)(0x80 - lv) << 16) | ((unsigned)(0x80 + hv) << 24);
}
// quantised picture point of a ray direction d (n.d > 0): its cell (cx, cy) and its pre-check word
NLOS_HD void pg_ray(const PGridFrame& g, f3 d, int& cx, int& cy, unsigned& R) {
  const float rz = pg_rcp(dot3(d, g.n));
  const int uq = pg_quant(dot3(d, g.a) * rz, g.u0, g.su, g.qmax), vq = pg_quant(dot3(d, g.b) * rz, g.v0, g.sv, g.qmax);
  cx = uq >> kPgSub; cy = vq >> kPgSub;
  const unsigned su = (unsigned)(uq & 127), sv = (unsigned)(vq & 127);
  R = su - (su << 8) + (sv << 16) - (sv << 24);
}
NLOS_HD bool pg_precheck(unsigned E, unsigned R) { return ((E + R) & kPgMask) == kPgMask; }
// frame of one source: axes only (G = 0: no grid yet); pg_set_rect() then fixes the quantisation from the picture rectangle
// [U0,U1] x [V0,V1] that contains every projected vertex (with its padding)
NLOS_HD void pg_init_frame(f3 o, f3 n_raw, PGridFrame& g, bool& ok) {
  pg_make_axes(n_raw, g.a, g.b, g.n, ok);
  g.o = o; g.G = 0; g.u0 = g.v0 = 0.f; g.su = g.sv = 1.f; g.qmax = 0.f;
}
NLOS_HD void pg_set_rect(PGridFrame& g, float U0, float U1, float V0, float V1, int G) {
  g.G = 0;
  if (G <= 0 || !(U1 >= U0) || !(V1 >= V0)) return;
  const float eu = 1.0e-3f * (U1 - U0) + 1.0e-6f, ev = 1.0e-3f * (V1 - V0) + 1.0e-6f;
  U0 -= eu; U1 += eu; V0 -= ev; V1 += ev;
  if (!(U1 - U0 < 3.0e30f) || !(V1 - V0 < 3.0e30f)) return;
  const float Q = (float)(G << kPgSub);
  g.u0 = U0; g.v0 = V0; g.su = Q / (U1 - U0); g.sv = Q / (V1 - V0); g.qmax = Q - 1.0f; g.G = G;
}
// halve the resolution (entry budget exceeded): same rectangle, coarser cells
NLOS_HD void pg_coarsen(PGridFrame& g) {
  const int G = g.G > 1 ? g.G >> 1 : 1;
  const float k = (float)G / (float)g.G;
  g.su *= k; g.sv *= k; g.G = G; g.qmax = (float)(G << kPgSub) - 1.0f;
}
NLOS_HD float pg_zmin(const float* blo, const float* bhi) { return 1.0e-3f * ((bhi[0] - blo[0]) + (bhi[1] - blo[1]) + (bhi[2] - blo[2])) + 1.0e-30f; }

// ------------------------------------------------------------------ shared perspective grid of a GROUP of wall points (DESIGN.md "K1s")
// The picture of the section above belongs to ONE projection centre.  Seen from a centre c, a ray that leaves another point
// o' = c + delta is no longer a point of the picture, but it is a straight line of the projective space (u, v, w), w = 1/Z:
//     u(w) = u_inf + s_u * w,   u_inf = d.a / d.n,   s_u = delta.a - (delta.n) u_inf      (v alike)
// (for wall points of one plane delta.n = 0 and the slope is just the offset of the wall point from the centre).  So the triangles
// are binned ONCE per group of neighbouring wall points into a 3-D grid (G x G picture cells x K slices of w); a ray visits the
// slices from the nearest one to the slice of its own hit and, in each, tests the list of the cell that the slice-centre point of its
// line falls into.  The line moves by at most S * dw / 2 inside a slice (S = slope bound of the group, dw = slice thickness): the
// rectangles are expanded by that much, and rays whose slope exceeds S take the per-ray BVH query instead.
// Every entry carries, besides the guarded rectangle bytes of the section above, three EDGE words and a fine depth index j (one of 32
// sub-slices of its slice: the middle of the triangle's own w range).  An edge word holds one picture edge of the triangle in the
// sub-cell units of the entry's cell (shifted by kGgBias) as signed bytes [a | b | c_hi | c_lo]: a x + b y + 255 c_hi + c_lo >= 0
// inside, rounded OUTWARDS by the quantisation slack, the padding and the movement of a ray across the triangle's own w range.  A
// candidate that passed the rectangle bytes is tested at the point of the ray's line at sub-slice j: one dp4a per edge against the
// ray word [x | y | 255 | 1].  Only candidates that pass all three edges get the exact Moeller-Trumbore test — little more than the
// ray's own triangle.  As before, the grid only SELECTS candidates: the answer is the exact float test's.
struct GGFrame {
  f3 o, a, b, n;            // projection centre (centre of the group), frame (n = unit wall normal of the group's first member)
  float u0, v0, su, sv;     // uq = clamp(floor((u - u0) * su), 0, Q-1)
  int G;                    // cells per axis (0: no grid for this group -> per-ray BVH query)
  float qmax;               // Q - 1 as float
  int K; float kmax;        // slices of w
  float w1, sw, dw;         // slice of w: clamp(floor((w1 - w) * sw), 0, K-1), slice 0 = nearest; dw = 1 / sw
  float Su, Sv;             // slope bounds |du/dw|, |dv/dw| of the rays that may use the grid
  float pad_u, pad_v;       // round-off padding of picture coordinates
  float eu, ev;             // expansion of a triangle's picture rectangle (padding + movement of a ray inside a slice)
  float pad_w;              // round-off padding of w
};
constexpr int kGgFine = 32;                    // sub-slices per slice (5 bits beside the 27-bit triangle index)
constexpr int kGgBias = 16;                    // shift of the sub-cell coordinates in the edge words: points up to 16 units outside the cell stay representable
constexpr int kGgSpan = 128 + 2 * kGgBias;     // biased coordinates lie in [0, kGgSpan)
constexpr float kGgNorm = 100.0f;              // |a|, |b| <= kGgNorm: |a x + b y| <= 2 * 100 * 160 = 32000 < 255 * 127
NLOS_HD int dp4a_s8_u8(unsigned a, unsigned b) {        // sum of the four signed bytes of a times the four unsigned bytes of b
#if defined(__CUDA_ARCH__)
  int d; asm("dp4a.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(0)); return d;
#else
  int d = 0;
  for (int i = 0; i < 4; ++i) d += (int)(signed char)((a >> (8 * i)) & 0xffu) * (int)((b >> (8 * i)) & 0xffu);
  return d;
#endif
}
// projection of a scene point: picture coordinates, w = 1/Z, Z, relative round-off margin (as pg_project)
NLOS_HD void gg_project(f3 o, f3 a, f3 b, f3 n, f3 x, float& u, float& v, float& w, float& z, float& m) {
  const f3 rel = x - o;
  z = dot3(rel, n);
  w = pg_rcp(z);
  u = dot3(rel, a) * w; v = dot3(rel, b) * w;
  m = kPgPad * ((fabsf(rel.x) + fabsf(rel.y) + fabsf(rel.z)) * w);
}
NLOS_HD int gg_slice(const GGFrame& g, float w) { float t = (g.w1 - w) * g.sw; t = fminf(fmaxf(t, 0.0f), g.kmax); return (int)t; }   // monotone (decreasing) in w
NLOS_HD int gg_fine(const GGFrame& g, float w, int k) {          // sub-slice of w inside slice k (clamped: w may lie in another slice)
  float t = ((g.w1 - w) * g.sw - (float)k) * (float)kGgFine; t = fminf(fmaxf(t, 0.0f), (float)(kGgFine - 1)); return (int)t;
}
// one edge of a projected triangle, in the biased sub-cell units of a cell: (xa,ya)->(xb,yb) with (xc,yc) on the inner side.
// ex, ey: how far (sub-cell units) the true picture point may be from the real-valued point the tested integer point was cut from.
// Returns the edge word; 'never' is set when no point of the cell can satisfy the edge (the triangle misses the cell).
NLOS_HD unsigned gg_edge_word(float xa, float ya, float xb, float yb, float xc, float yc, float ex, float ey, bool& never) {
  float A = ya - yb, B = xb - xa;
  float C = -fmaf(A, xa, B * ya);
  const float side = fmaf(A, xc, fmaf(B, yc, C));
  if (side < 0.0f) { A = -A; B = -B; C = -C; }
  const float mx = fmaxf(fabsf(A), fabsf(B));
  if (!(mx > 1.0e-20f) || !(mx < 1.0e20f)) return 0u;                     // degenerate in the picture: edge always passes (a = b = c = 0)
  const float s = kGgNorm / mx;
  const float fa = rintf(A * s), fb = rintf(B * s);
  // slack: rounding of a, b (<= 0.5 each) over |x|, |y| <= span + e; distance of the tested integer point from the true one; float evaluation of C
  const float slack = 0.5f * (((float)kGgSpan + 1.0f + ex) + ((float)kGgSpan + 1.0f + ey)) + fabsf(fa) * (1.0f + ex) + fabsf(fb) * (1.0f + ey) +
                      1.0e-5f * s * (fabsf(A * xa) + fabsf(B * ya)) + 2.0f;
  if (!(fabsf(side) * s > slack)) return 0u;                               // thinner than the slack (edge-on): orientation not certain, edge always passes
  const float cf = C * s + slack;
  if (!(cf < 32000.0f)) return 0u;                                         // (nearly) every point of the cell satisfies it: always passes
  if (cf < -32100.0f) { never = true; return 0u; }
  const int c = (int)ceilf(cf) + 1;
  const int chi = (c >= 0 ? c + 127 : c - 127) / 255;
  const int clo = c - 255 * chi;                                           // in [-127, 127]
  return ((unsigned)(int)fa & 0xffu) | (((unsigned)(int)fb & 0xffu) << 8) | (((unsigned)chi & 0xffu) << 16) | (((unsigned)clo & 0xffu) << 24);
}
NLOS_HD bool gg_edges_pass(unsigned e0, unsigned e1, unsigned e2, unsigned rayword) {
  return (dp4a_s8_u8(e0, rayword) | dp4a_s8_u8(e1, rayword) | dp4a_s8_u8(e2, rayword)) >= 0;
}
// quantisation, slices, slope bounds and expansions of a group's frame from the extent of its projected vertices
// ([U0,U1] x [V0,V1] x [W0,W1], mm = largest round-off margin) and of its members (largest |delta.a|, |delta.b|, |delta.n|)
NLOS_HD void gg_finish_frame(GGFrame& g, float U0, float U1, float V0, float V1, float W0, float W1, float mm,
                             float max_da, float max_db, float max_dn, int G, int K) {
  g.G = 0; g.K = K; g.kmax = (float)(K - 1);
  if (G <= 0 || K <= 0 || !(U1 >= U0) || !(V1 >= V0) || !(W1 >= W0) || !(W0 > 0.0f) || !(mm < 1.0f)) return;
  const float ua = fmaxf(fabsf(U0), fabsf(U1)), va = fmaxf(fabsf(V0), fabsf(V1));
  g.pad_u = mm * (1.0f + ua); g.pad_v = mm * (1.0f + va);
  g.pad_w = mm * W1;
  g.dw = fmaxf((W1 - W0) + 2.0f * g.pad_w, 1.0e-30f) / (float)K;
  g.sw = 1.0f / g.dw; g.w1 = W1 + g.pad_w;
  g.Su = (max_da + max_dn * (2.0f * ua + 1.0f)) * 1.001f + 1.0e-12f;
  g.Sv = (max_db + max_dn * (2.0f * va + 1.0f)) * 1.001f + 1.0e-12f;
  g.eu = g.pad_u + g.Su * (0.5f * g.dw + 2.0f * g.pad_w); g.ev = g.pad_v + g.Sv * (0.5f * g.dw + 2.0f * g.pad_w);
  U0 -= g.eu; U1 += g.eu; V0 -= g.ev; V1 += g.ev;
  const float mu = 1.0e-3f * (U1 - U0) + 1.0e-6f, mv = 1.0e-3f * (V1 - V0) + 1.0e-6f;
  U0 -= mu; U1 += mu; V0 -= mv; V1 += mv;
  if (!(U1 - U0 < 3.0e30f) || !(V1 - V0 < 3.0e30f)) return;
  const float Q = (float)(G << kPgSub);
  g.u0 = U0; g.v0 = V0; g.su = Q / (U1 - U0); g.sv = Q / (V1 - V0); g.qmax = Q - 1.0f; g.G = G;
}
NLOS_HD void gg_coarsen(GGFrame& g) {        // halve the picture resolution (entry budget exceeded): same rectangle, coarser cells
  const int G = g.G > 1 ? g.G >> 1 : 1;
  const float k = (float)G / (float)g.G;
  g.su *= k; g.sv *= k; g.G = G; g.qmax = (float)(G << kPgSub) - 1.0f;
}
// quantised, expanded rectangle and slice range of a triangle from its projected vertices (u, v, w); wlo / whi: its padded w range
NLOS_HD void gg_tri_box(const GGFrame& g, float u1, float v1, float w1, float u2, float v2, float w2, float u3, float v3, float w3,
                        int& a0, int& a1, int& b0, int& b1, int& k0, int& k1, float& wlo, float& whi) {
  const float ulo = fminf(u1, fminf(u2, u3)) - g.eu, uhi = fmaxf(u1, fmaxf(u2, u3)) + g.eu;
  const float vlo = fminf(v1, fminf(v2, v3)) - g.ev, vhi = fmaxf(v1, fmaxf(v2, v3)) + g.ev;
  a0 = pg_quant(ulo, g.u0, g.su, g.qmax); a1 = pg_quant(uhi, g.u0, g.su, g.qmax);
  b0 = pg_quant(vlo, g.v0, g.sv, g.qmax); b1 = pg_quant(vhi, g.v0, g.sv, g.qmax);
  whi = fmaxf(w1, fmaxf(w2, w3)) + g.pad_w; wlo = fminf(w1, fminf(w2, w3)) - g.pad_w;
  k0 = gg_slice(g, whi); k1 = gg_slice(g, wlo);
}
// fine depth index of a triangle's entry in slice k (middle sub-slice of its w range clipped to the slice) and the half width, in
// sub-slices, of the w interval that index stands for
NLOS_HD int gg_tri_fine(const GGFrame& g, float wlo, float whi, int k0, int k1, int k, float& half) {
  const int ja = k > k0 ? 0 : gg_fine(g, whi, k), jb = k < k1 ? kGgFine - 1 : gg_fine(g, wlo, k);
  const int jm = (ja + jb) >> 1;
  half = 0.5f * (float)(jb - ja + 1) + 1.5f;            // the range itself, the middle's rounding, one sub-slice for the float assignment of ja / jb
  return jm;
}
// the three edge words of a triangle in cell (cx, cy); half = largest gg_tri_fine() half width over the triangle's slices;
// false: the triangle provably misses the cell
NLOS_HD bool gg_tri_edges(const GGFrame& g, float u1, float v1, float u2, float v2, float u3, float v3, int cx, int cy, float half,
                          unsigned& e0, unsigned& e1, unsigned& e2) {
  const float bx = (float)((cx << kPgSub) - kGgBias), by = (float)((cy << kPgSub) - kGgBias);
  const float x1 = (u1 - g.u0) * g.su - bx, y1 = (v1 - g.v0) * g.sv - by;
  const float x2 = (u2 - g.u0) * g.su - bx, y2 = (v2 - g.v0) * g.sv - by;
  const float x3 = (u3 - g.u0) * g.su - bx, y3 = (v3 - g.v0) * g.sv - by;
  const float hw = half * g.dw * (1.0f / (float)kGgFine) + 2.0f * g.pad_w;
  const float ex = (g.pad_u + g.Su * hw) * g.su, ey = (g.pad_v + g.Sv * hw) * g.sv;
  bool never = false;
  e0 = gg_edge_word(x1, y1, x2, y2, x3, y3, ex, ey, never);
  e1 = gg_edge_word(x2, y2, x3, y3, x1, y1, ex, ey, never);
  e2 = gg_edge_word(x3, y3, x1, y1, x2, y2, ex, ey, never);
  return !never;
}
// the line of a ray in the group's projective space; false: the ray cannot use the grid (points away from the picture's half space,
// or moves faster than the group's slope bound)
struct GGRay { float ui, vi, su, sv; int kr; };
NLOS_HD bool gg_ray_setup(const GGFrame& g, float da, float db, float dn /* delta in the frame */, f3 d, float ts, GGRay& r) {
  const float zn = dot3(d, g.n);
  if (!(zn > 0.0f)) return false;
  const float rz = pg_rcp(zn);
  r.ui = dot3(d, g.a) * rz; r.vi = dot3(d, g.b) * rz;
  r.su = da - dn * r.ui; r.sv = db - dn * r.vi;
  if (!(fabsf(r.su) <= g.Su) || !(fabsf(r.sv) <= g.Sv)) return false;
  const float zh = fmaf(ts, zn, dn);
  if (!(zh > 0.0f)) return false;
  r.kr = gg_slice(g, pg_rcp(zh) - g.pad_w);
  return true;
}
// the ray's walk through the slices: quantised picture coordinates (float, in sub-cell units of the whole picture) of the point at the
// centre of slice k are  fu = au + k bu,  fv = av + k bv
struct GGWalk { float au, bu, av, bv; };
NLOS_HD GGWalk gg_ray_walk(const GGFrame& g, const GGRay& r) {
  GGWalk w;
  const float w0 = g.w1 - 0.5f * g.dw;
  w.au = (fmaf(r.su, w0, r.ui) - g.u0) * g.su; w.bu = -(r.su * g.dw) * g.su;
  w.av = (fmaf(r.sv, w0, r.vi) - g.v0) * g.sv; w.bv = -(r.sv * g.dw) * g.sv;
  return w;
}
// quantised point of the walk in slice k (kf = (float)k): uq, vq in [0, Q-1]; the cell is (uq >> 7, vq >> 7)
NLOS_HD void gg_walk_point(const GGFrame& g, const GGWalk& w, float kf, int& uq, int& vq) {
  const float fu = fminf(fmaxf(fmaf(kf, w.bu, w.au), 0.0f), g.qmax), fv = fminf(fmaxf(fmaf(kf, w.bv, w.av), 0.0f), g.qmax);
  uq = (int)fu; vq = (int)fv;
}
NLOS_HD unsigned gg_rect_word(int uq, int vq) {                  // rectangle check word of a quantised point (see pg_ray)
  const unsigned su = (unsigned)(uq & 127), sv = (unsigned)(vq & 127);
  return su - (su << 8) + (sv << 16) - (sv << 24);
}
// biased sub-cell coordinates of the walk's point at sub-slice 0 of slice k, and their change per sub-slice (edge check)
struct GGFinePoint { float x0, y0, dx, dy; };
NLOS_HD GGFinePoint gg_walk_fine(const GGWalk& w, float kf, int uq, int vq) {
  GGFinePoint q;
  q.dx = w.bu * (1.0f / (float)kGgFine); q.dy = w.bv * (1.0f / (float)kGgFine);
  const float half = 0.5f * (float)(kGgFine - 1);
  q.x0 = (fmaf(kf, w.bu, w.au) - (float)((uq & ~127) - kGgBias)) - half * q.dx;
  q.y0 = (fmaf(kf, w.bv, w.av) - (float)((vq & ~127) - kGgBias)) - half * q.dy;
  return q;
}
// edge check word of the ray at sub-slice j; false: the point left the representable range (the candidate then goes to the exact test)
NLOS_HD bool gg_ray_word(const GGFinePoint& q, int j, unsigned& rayword) {
  const int x = (int)floorf(fmaf((float)j, q.dx, q.x0)), y = (int)floorf(fmaf((float)j, q.dy, q.y0));
  rayword = (unsigned)x | ((unsigned)y << 8) | 0x01ff0000u;
  return (unsigned)x < (unsigned)kGgSpan && (unsigned)y < (unsigned)kGgSpan;
}

// ------------------------------------------------------------------ LBVH construction helpers (Karras 2012)
NLOS_HD uint32_t expand_bits10(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u; return v;
}
NLOS_HD uint32_t morton30(float x, float y, float z) {   // x,y,z in [0,1]
  const float s = 1024.0f;
  const uint32_t xi = (uint32_t)fminf(fmaxf(x * s, 0.0f), 1023.0f);
  const uint32_t yi = (uint32_t)fminf(fmaxf(y * s, 0.0f), 1023.0f);
  const uint32_t zi = (uint32_t)fminf(fmaxf(z * s, 0.0f), 1023.0f);
  return (expand_bits10(xi) << 2) | (expand_bits10(yi) << 1) | expand_bits10(zi);
}
NLOS_HD int clz64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}
// keys are unique (low 32 bits = triangle index), so delta is a plain common-prefix length
NLOS_HD int lbvh_delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  return clz64(keys[i] ^ keys[j]);
}
// internal node i covers sorted leaves [first,last]; children split at 'split' | 'split+1'
NLOS_HD void lbvh_range(const uint64_t* __restrict__ keys, int n, int i, int& first, int& last, int& split) {
  const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lbvh_delta(keys, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1) if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  const int j = i + l * d;
  const int dnode = lbvh_delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t <= 1) break;
  }
  split = i + s * d + (d < 0 ? -1 : 0);
  first = i < j ? i : j; last = i < j ? j : i;
}

// ------------------------------------------------------------------ GGX (confocal: wi = wo = h = w), x = n.w
NLOS_HD float ggx_D(float a, float nw) {
  if (nw <= 0) return 0.0f;
  const float nw2 = nw * nw;
  const float be = (1.0f - nw2) / (a * a) / nw2;
  const float root = (1.0f + be) * nw2;
  float result = 1.0f / ((float)M_PI * a * a * root * root);
  if (result * nw < 1e-20f) result = 0;
  return result;
}
NLOS_HD float ggx_G1(float a, float nw) {
  if (nw <= 0) return 0.0f;
  if (nw >= 1.0f) return 1.0f;
  const float root = a * a + (1.0f - a * a) * nw * nw;
  return 2.0f / (nw + sqrtf(root));
}
NLOS_HD float ggx_eval(float a, float nw) {
  if (nw <= 0) return 0.0f;
  const float Dv = ggx_D(a, nw); if (Dv == 0) return 0.0f;
  const float g = ggx_G1(a, nw);
  return Dv * (g * g) / 4.0f;
}
NLOS_HD float ggx_eval_adiff(float a, float nw) {
  if (nw <= 0) return 0.0f;
  const float Dv = ggx_D(a, nw); if (Dv == 0) return 0.0f;
  const float g1 = ggx_G1(a, nw), Gv = g1 * g1;
  const float nw2 = nw * nw, a2 = a * a;
  const float val = a2 * nw2 - nw2 + 1;
  const float Dp = -(2.0f * a * (a2 * nw2 + nw2 - 1)) / ((float)M_PI * val * val * val);
  float g1p = 0.0f;
  if (nw < 1.0f) { const float vv = sqrtf(a2 - nw2 * (a2 - 1)); const float root = nw + vv; g1p = 2.0f * a * (nw2 - 1.0f) / (vv * root * root); }
  const float Gp = 2.0f * g1p * g1;
  return (Dp * Gv + Gp * Dv) / 4.0f;
}
// d f / d (n.w); df/dn = S*w, df/dw = S*n
NLOS_HD float ggx_eval_xdiff(float a, float nw) {
  if (nw <= 0) return 0.0f;
  const float Dv = ggx_D(a, nw); if (Dv == 0) return 0.0f;
  const float g1 = ggx_G1(a, nw), Gv = g1 * g1;
  const float nw2 = nw * nw, a2 = a * a;
  const float root = (a2 - 1.0f) * nw2 + 1.0f;
  const float Dp = -(4.0f * a2 * nw * (a2 - 1.0f)) / ((float)M_PI * root * root * root);
  float g1p = 0.0f;
  if (nw < 1.0f) { const float temp = sqrtf(a2 - nw2 * (a2 - 1.0f)); const float rt = nw + temp; g1p = -2.0f * (1.0f - (nw * (a2 - 1.0f)) / temp) / rt / rt; }
  const float Gp = 2.0f * g1p * g1;
  return (Dp * Gv + Gp * Dv) / 4.0f;
}

// ------------------------------------------------------------------ one stratified sample
struct ShadeTri { f3 v1, v2, v3, nf; float A; int i1, i2, i3; };

struct SampleGeom { f3 d; float u, v, w, r, t; };

// generate the sample point / ray direction (TG.cpp:184-195) and run the self intersection that yields
// the hit barycentrics the reference reads back from Embree (TG.cpp:208-212).
NLOS_HD bool sample_self_hit_st(float S, float T, f3 o, const ShadeTri& st, const TriRec& tr, SampleGeom& g) {
  const float sqrtT = sqrtf(T);
  const float u = 1 - sqrtT, v = (1 - S) * sqrtT, w = S * sqrtT;
  const f3 point = blend3(u, st.v1, v, st.v2, w, st.v3);
  const f3 q = point - o;
  const float inv = 1.0f / len3(q);
  g.d = q * inv;
  float t, hu, hv;
  if (!isect(tr, o, g.d, t, hu, hv)) return false;
  g.t = t; g.v = hu; g.w = hv; g.u = 1.0f - hu - hv;
  const f3 pt = blend3(g.u, st.v1, g.v, st.v2, g.w, st.v3);
  g.r = len3(pt - o);
  return true;
}
NLOS_HD bool sample_self_hit(uint64_t seed, int64_t src_global, int prim, int k, f3 o, const ShadeTri& st, const TriRec& tr, SampleGeom& g) {
  float S, T; sample_ST(seed, src_global, prim, k, S, T);
  return sample_self_hit_st(S, T, o, st, tr, g);
}

// tap -> coarse-bin grouping shared by the smoothed forward splat and the gradient (DESIGN.md "K-tap
// restructuring"): tap i of a sample in fine bin m0 lands in coarse bin floor((m0 + i - 2rs)/r).
// For coarse bin b the taps are i in [ilo, ihi) with:
NLOS_HD void tap_span(int64_t m0, int b, int r, int half /*2rs*/, int K, int& ilo, int& ihi) {
  int64_t lo = (int64_t)b * r - m0 + half, hi = lo + r;
  ilo = lo < 0 ? 0 : (lo > K ? K : (int)lo);
  ihi = hi < 0 ? 0 : (hi > K ? K : (int)hi);
}
// 32-bit versions for the kernels (fine-bin indices are range-checked to +-2^30 first): the 64-bit ones above cost ~50 SASS
// instructions per division (two per visible sample in the gradient kernel)
NLOS_HD int floordiv32(int a, int b /* > 0 */) { const int q = a / b; return (a - q * b < 0) ? q - 1 : q; }
NLOS_HD void tap_span32(int m0, int b, int r, int half /*2rs*/, int K, int& ilo, int& ihi) {
  const int lo = b * r - m0 + half, hi = lo + r;
  ilo = lo < 0 ? 0 : (lo > K ? K : lo);
  ihi = hi < 0 ? 0 : (hi > K ? K : hi);
}
NLOS_HD int64_t floordiv(int64_t a, int64_t b) { int64_t q = a / b; return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q; }

}  // namespace nlos
