// lbvh.cu — K0/K0a: per-call scene setup on the device.
//
// Replaces the reference's Embree scene construction (rtcNewGeometry .. rtcCommitScene with
// RTC_BUILD_QUALITY_HIGH, smoothed_transient/stratifiedStreamedGradientRenderer.cpp:473-511) and the
// per-(source,triangle) task preamble that recomputes v1, edges, face normal and area for every task
// (smoothed_transient/transient_and_gradient.cpp:146-159).
//
// Pipeline (all on cx.stream, no host sync):
//   k_scene_bounds : centroid bounds + max|coordinate|            (V,F reads; a few atomics per warp)
//   k_morton_keys  : key = morton30(centroid) << 32 | triangle     (F)
//   cub radix sort : 62-bit keys                                   (F)
//   k_tri_records  : TraceTri / ShadeTri / padded leaf boxes in Morton order   (F, 3 vertex gathers each)
//   k_karras       : Karras-2012 topology, one thread per internal node        (F-1)
//   k_refit        : bottom-up AABB union with per-node arrival counters       (F)
//   k_emit_nodes   : 64-byte traversal nodes (both child boxes inline, leaf runs of <= kLeafMax triangles)
#include <cub/device/device_radix_sort.cuh>
#include "nlos_ctx.h"

namespace nlos {

namespace {

constexpr int kThreads = 256;
inline int blocks_for(int64_t n, int t = kThreads) { return (int)((n + t - 1) / t); }

// a face index outside [0, V) is clamped (memory safety) and counted (the call then fails with NLOS_ERR_INVALID, nlos_abi.cu)
__device__ __forceinline__ int clampv(int i, int V) { return i < 0 ? 0 : (i >= V ? V - 1 : i); }
__device__ __forceinline__ f3 ldv(const float* __restrict__ v, int i) { return mk3(__ldg(v + 3 * (size_t)i), __ldg(v + 3 * (size_t)i + 1), __ldg(v + 3 * (size_t)i + 2)); }

__global__ void k_init_bounds(SceneBounds* sb) {
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a) { sb->lo[a] = sb->vlo[a] = f2ord(3.0e38f); sb->hi[a] = sb->vhi[a] = f2ord(-3.0e38f); }
    sb->absmax = 0u; sb->bad_faces = 0u;
  }
}

__global__ void k_absmax(const float* __restrict__ a, size_t n, SceneBounds* sb) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) m = fmaxf(m, fabsf(__ldg(a + i)));
#pragma unroll
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&sb->absmax, (unsigned)__float_as_int(m));   // m >= 0: int order == float order
}

__global__ void k_scene_bounds(const float* __restrict__ verts, int V, const int* __restrict__ faces, int F, SceneBounds* sb) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  float vl[3] = {3.0e38f, 3.0e38f, 3.0e38f}, vh[3] = {-3.0e38f, -3.0e38f, -3.0e38f};     // bounds of the vertices the faces use
  if (f < F) {
    const int i1 = faces[3 * (size_t)f], i2 = faces[3 * (size_t)f + 1], i3 = faces[3 * (size_t)f + 2];
    if (i1 < 0 || i1 >= V || i2 < 0 || i2 >= V || i3 < 0 || i3 >= V) atomicAdd(&sb->bad_faces, 1u);
    const f3 a = ldv(verts, clampv(i1, V)), b = ldv(verts, clampv(i2, V)), c = ldv(verts, clampv(i3, V));
    vl[0] = fminf(a.x, fminf(b.x, c.x)); vh[0] = fmaxf(a.x, fmaxf(b.x, c.x));
    vl[1] = fminf(a.y, fminf(b.y, c.y)); vh[1] = fmaxf(a.y, fmaxf(b.y, c.y));
    vl[2] = fminf(a.z, fminf(b.z, c.z)); vh[2] = fmaxf(a.z, fmaxf(b.z, c.z));
    const float cx = 0.5f * (vl[0] + vh[0]);
    const float cy = 0.5f * (vl[1] + vh[1]);
    const float cz = 0.5f * (vl[2] + vh[2]);
    lo[0] = hi[0] = cx; lo[1] = hi[1] = cy; lo[2] = hi[2] = cz;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o)); hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
      vl[k] = fminf(vl[k], __shfl_xor_sync(0xffffffffu, vl[k], o)); vh[k] = fmaxf(vh[k], __shfl_xor_sync(0xffffffffu, vh[k], o));
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { atomicMin(&sb->lo[k], f2ord(lo[k])); atomicMax(&sb->hi[k], f2ord(hi[k])); atomicMin(&sb->vlo[k], f2ord(vl[k])); atomicMax(&sb->vhi[k], f2ord(vh[k])); }
  }
}

__global__ void k_morton_keys(const float* __restrict__ verts, int V, const int* __restrict__ faces, int F, const SceneBounds* __restrict__ sb, uint64_t* __restrict__ keys) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const f3 a = ldv(verts, clampv(faces[3 * (size_t)f], V)), b = ldv(verts, clampv(faces[3 * (size_t)f + 1], V)), c = ldv(verts, clampv(faces[3 * (size_t)f + 2], V));
  const float cx = 0.5f * (fminf(a.x, fminf(b.x, c.x)) + fmaxf(a.x, fmaxf(b.x, c.x)));
  const float cy = 0.5f * (fminf(a.y, fminf(b.y, c.y)) + fmaxf(a.y, fmaxf(b.y, c.y)));
  const float cz = 0.5f * (fminf(a.z, fminf(b.z, c.z)) + fmaxf(a.z, fmaxf(b.z, c.z)));
  const float lx = ord2f(sb->lo[0]), ly = ord2f(sb->lo[1]), lz = ord2f(sb->lo[2]);
  const float ex = fmaxf(ord2f(sb->hi[0]) - lx, 1e-30f), ey = fmaxf(ord2f(sb->hi[1]) - ly, 1e-30f), ez = fmaxf(ord2f(sb->hi[2]) - lz, 1e-30f);
  const uint32_t code = morton30((cx - lx) / ex, (cy - ly) / ey, (cz - lz) / ez);
  keys[f] = ((uint64_t)code << 32) | (uint32_t)f;
}

// Morton-ordered triangle records and padded leaf boxes.
__global__ void k_tri_records(const float* __restrict__ verts, int V, const int* __restrict__ faces, int F, const uint64_t* __restrict__ keys,
                              const SceneBounds* __restrict__ sb, float4* __restrict__ ttris, float4* __restrict__ stris, int* __restrict__ sprim,
                              float4* __restrict__ leaf_lo, float4* __restrict__ leaf_hi) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= F) return;
  const int f = (int)(uint32_t)keys[p];
  const int i1 = clampv(faces[3 * (size_t)f], V), i2 = clampv(faces[3 * (size_t)f + 1], V), i3 = clampv(faces[3 * (size_t)f + 2], V);
  const f3 v1 = ldv(verts, i1), v2 = ldv(verts, i2), v3 = ldv(verts, i3);
  const TriRec tr = make_tri(v1, v2, v3);
  ttris[4 * (size_t)p + 0] = make_float4(tr.v0.x, tr.v0.y, tr.v0.z, __int_as_float(f));
  ttris[4 * (size_t)p + 1] = make_float4(tr.e1.x, tr.e1.y, tr.e1.z, 0.f);
  ttris[4 * (size_t)p + 2] = make_float4(tr.e2.x, tr.e2.y, tr.e2.z, 0.f);
  ttris[4 * (size_t)p + 3] = make_float4(tr.Ng.x, tr.Ng.y, tr.Ng.z, 0.f);
  // TG.cpp:157-159: faceNormal = cross(v2-v1, v3-v1); faceArea = |N|/2; faceNormal /= 2*faceArea
  const f3 N = cross3(v2 - v1, v3 - v1);
  const float A = len3(N) / 2;
  const f3 nf = N / (2 * A);
  // component-major ([4][F]): the sample kernels read these lane <-> triangle, so every 16-byte load of a warp is one contiguous 512 bytes
  // (triangle-major records cost 16 L1 wavefronts per load instead of 4 — 17 % of the forward kernel's L1 data-pipe traffic)
  stris[p] = make_float4(v1.x, v1.y, v1.z, A);
  stris[(size_t)F + p] = make_float4(v2.x, v2.y, v2.z, nf.x);
  stris[2 * (size_t)F + p] = make_float4(v3.x, v3.y, v3.z, nf.y);
  stris[3 * (size_t)F + p] = make_float4(nf.z, __int_as_float(i1), __int_as_float(i2), __int_as_float(i3));
  sprim[p] = f;
  // boxes are padded so that a float-valid triangle hit is never culled by the (float) slab test
  const float pad = __int_as_float((int)sb->absmax) * (1.0f / 65536.0f);
  leaf_lo[p] = make_float4(fminf(v1.x, fminf(v2.x, v3.x)) - pad, fminf(v1.y, fminf(v2.y, v3.y)) - pad, fminf(v1.z, fminf(v2.z, v3.z)) - pad, 0.f);
  leaf_hi[p] = make_float4(fmaxf(v1.x, fmaxf(v2.x, v3.x)) + pad, fmaxf(v1.y, fmaxf(v2.y, v3.y)) + pad, fmaxf(v1.z, fmaxf(v2.z, v3.z)) + pad, 0.f);
}

// child encoding during construction: c >= 0 internal node, c < 0 leaf ~c
__global__ void k_karras(const uint64_t* __restrict__ keys, int F, int* __restrict__ first, int* __restrict__ last, int2* __restrict__ child,
                         int* __restrict__ parent_node, int* __restrict__ parent_leaf, int* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F - 1) return;
  int a, b, s; lbvh_range(keys, F, i, a, b, s);
  first[i] = a; last[i] = b; flags[i] = 0;
  int cl, cr;
  if (a == s) { cl = ~s; parent_leaf[s] = i; } else { cl = s; parent_node[s] = i; }
  if (b == s + 1) { cr = ~(s + 1); parent_leaf[s + 1] = i; } else { cr = s + 1; parent_node[s + 1] = i; }
  child[i] = make_int2(cl, cr);
  if (i == 0) parent_node[0] = -1;
}

__device__ __forceinline__ void load_box(int c, const float4* leaf_lo, const float4* leaf_hi, const float4* node_lo, const float4* node_hi, float4& lo, float4& hi) {
  if (c < 0) { lo = __ldcg(leaf_lo + (~c)); hi = __ldcg(leaf_hi + (~c)); } else { lo = __ldcg(node_lo + c); hi = __ldcg(node_hi + c); }
}

__global__ void k_refit(int F, const int2* __restrict__ child, const int* __restrict__ parent_node, const int* __restrict__ parent_leaf,
                        const float4* leaf_lo, const float4* leaf_hi, float4* node_lo, float4* node_hi, int* flags) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= F || F < 2) return;
  int cur = parent_leaf[p];
  while (cur >= 0) {
    __threadfence();
    if (atomicAdd(&flags[cur], 1) == 0) return;      // first arrival: the sibling subtree is not finished yet
    __threadfence();
    const int2 ch = child[cur];
    float4 l0, h0, l1, h1;
    load_box(ch.x, leaf_lo, leaf_hi, node_lo, node_hi, l0, h0);
    load_box(ch.y, leaf_lo, leaf_hi, node_lo, node_hi, l1, h1);
    __stcg(node_lo + cur, make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f));
    __stcg(node_hi + cur, make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f));
    cur = parent_node[cur];
  }
}

__global__ void k_emit_nodes(int F, const int2* __restrict__ child, const int* __restrict__ first, const int* __restrict__ last,
                             const float4* __restrict__ leaf_lo, const float4* __restrict__ leaf_hi, const float4* __restrict__ node_lo,
                             const float4* __restrict__ node_hi, BvhNode* __restrict__ nodes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= F - 1) return;
  const int2 ch = child[i];
  int link[2], cnt[2]; float4 lo[2], hi[2];
  const int cc[2] = {ch.x, ch.y};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int c = cc[k];
    if (c < 0) { link[k] = ~c; cnt[k] = 1; lo[k] = leaf_lo[~c]; hi[k] = leaf_hi[~c]; }
    else {
      const int size = last[c] - first[c] + 1;
      lo[k] = node_lo[c]; hi[k] = node_hi[c];
      if (size <= kLeafMax) { link[k] = first[c]; cnt[k] = size; } else { link[k] = c; cnt[k] = 0; }
    }
  }
  BvhNode n;
  n.a = make_float4(lo[0].x, lo[0].y, lo[0].z, hi[0].x);
  n.b = make_float4(hi[0].y, hi[0].z, lo[1].x, lo[1].y);
  n.c = make_float4(lo[1].z, hi[1].x, hi[1].y, hi[1].z);
  n.d = make_int4(cnt[0] > 0 ? leaf_ref(link[0], cnt[0]) : link[0], cnt[1] > 0 ? leaf_ref(link[1], cnt[1]) : link[1], 0, 0);
  nodes[i] = n;
}

}  // namespace

void build_scene(Ctx& cx, const float* d_verts, int V, const int* d_faces, int F, const float* d_origin, int64_t L,
                 const float* d_vnormal, const float* d_valbedo, DeviceScene& out) {
  cudaStream_t st = cx.stream;
  out = DeviceScene(); out.F = F; out.V = V; out.vnormal = d_vnormal; out.valbedo = d_valbedo;
  if (F <= 0) return;
  SceneBounds* sb = cx.buf("bounds").as<SceneBounds>(1);
  uint64_t* keys_in = cx.buf("keys_in").as<uint64_t>(F);
  uint64_t* keys = cx.buf("keys").as<uint64_t>(F);
  float4* ttris = cx.buf("ttris").as<float4>(4 * (size_t)F);
  float4* stris = cx.buf("stris").as<float4>(4 * (size_t)F);
  int* sprim = cx.buf("sprim").as<int>((size_t)F);
  float4* leaf_lo = cx.buf("leaf_lo").as<float4>(F);
  float4* leaf_hi = cx.buf("leaf_hi").as<float4>(F);
  const int NI = F > 1 ? F - 1 : 1;
  float4* node_lo = cx.buf("node_lo").as<float4>(NI);
  float4* node_hi = cx.buf("node_hi").as<float4>(NI);
  int* first = cx.buf("node_first").as<int>(NI);
  int* last = cx.buf("node_last").as<int>(NI);
  int2* child = cx.buf("node_child").as<int2>(NI);
  int* parent_node = cx.buf("parent_node").as<int>(NI);
  int* parent_leaf = cx.buf("parent_leaf").as<int>(F);
  int* flags = cx.buf("node_flags").as<int>(NI);
  BvhNode* nodes = cx.buf("nodes").as<BvhNode>(NI);

  k_init_bounds<<<1, 32, 0, st>>>(sb);
  k_absmax<<<std::min(blocks_for(3 * (int64_t)V), 1024), kThreads, 0, st>>>(d_verts, 3 * (size_t)V, sb);
  if (L > 0) k_absmax<<<std::min(blocks_for(3 * L), 1024), kThreads, 0, st>>>(d_origin, 3 * (size_t)L, sb);
  k_scene_bounds<<<blocks_for(F), kThreads, 0, st>>>(d_verts, V, d_faces, F, sb);
  k_morton_keys<<<blocks_for(F), kThreads, 0, st>>>(d_verts, V, d_faces, F, sb, keys_in);
  cx.launches += 4 + (L > 0 ? 1 : 0);
  size_t tmp_bytes = 0;
  NLOS_CUDA_OK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, keys_in, keys, F, 0, 62, st));
  void* tmp = cx.buf("sort_tmp").ensure(tmp_bytes);
  NLOS_CUDA_OK(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, keys_in, keys, F, 0, 62, st));
  k_tri_records<<<blocks_for(F), kThreads, 0, st>>>(d_verts, V, d_faces, F, keys, sb, ttris, stris, sprim, leaf_lo, leaf_hi);
  cx.launches += 1;
  if (F > 1) {
    k_karras<<<blocks_for(F - 1), kThreads, 0, st>>>(keys, F, first, last, child, parent_node, parent_leaf, flags);
    k_refit<<<blocks_for(F), kThreads, 0, st>>>(F, child, parent_node, parent_leaf, leaf_lo, leaf_hi, node_lo, node_hi, flags);
    k_emit_nodes<<<blocks_for(F - 1), kThreads, 0, st>>>(F, child, first, last, leaf_lo, leaf_hi, node_lo, node_hi, nodes);
    cx.launches += 3;
  }
  NLOS_CUDA_OK(cudaGetLastError());
  out.ttris = ttris; out.stris = stris; out.sprim = sprim; out.nodes = nodes; out.bounds = sb; out.verts = d_verts;
  out.root_count = F <= kLeafMax ? F : 0;
}

}  // namespace nlos
