// mesh_kernels.cu — the cheap O(F) neighbours of the hot path as trivial kernels:
//   per-bin gradient of one vertex   smoothed_transient/transient_and_gradient.cpp:379-439, 697-840
//   normal-smoothing regulariser     smoothed_transient/stratifiedStreamedGradientRenderer.cpp:59-160
//   area ("curvature") gradient      smoothed_transient/stratifiedStreamedGradientRenderer.cpp:27-57, 162-180
//
// The two regularisers write per-vertex results with '=' from every adjacent face in the reference (last writer
// wins, and the writer depends on TBB scheduling).  Here the winner is deterministic: the adjacent face with the
// HIGHEST index, which is what a serial pass in face order produces (and what the oracle restates).
#include <algorithm>
#include "nlos_ctx.h"
#include "render_kernels.h"

namespace nlos {

namespace {

__device__ __forceinline__ f3 ldv3(const float* __restrict__ v, int i) { return mk3(__ldg(v + 3 * (size_t)i), __ldg(v + 3 * (size_t)i + 1), __ldg(v + 3 * (size_t)i + 2)); }

// ---- renderStreamedVertexGradient: thread = (sorted triangle, source); only triangles touching vertex_num do work
__global__ void k_vertex_gradient(const DeviceScene sc, const RenderParams P, int vertex_num, const double* __restrict__ taps /*K weights*/,
                                  double sigma2, double* __restrict__ acc /*[B,3]*/) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= sc.F) return;
  const float4 s3 = __ldg(sc.stris + 3 * (size_t)sc.F + p);
  const int i1 = __float_as_int(s3.y), i2 = __float_as_int(s3.z), i3 = __float_as_int(s3.w);
  if (i1 != vertex_num && i2 != vertex_num && i3 != vertex_num) return;                  // TG.cpp:729-731
  const float4 s0 = __ldg(sc.stris + p), s1 = __ldg(sc.stris + 1 * (size_t)sc.F + p), s2 = __ldg(sc.stris + 2 * (size_t)sc.F + p);
  ShadeTri st; st.v1 = xyz(s0); st.A = s0.w; st.v2 = xyz(s1); st.v3 = xyz(s2); st.nf = mk3(s1.w, s2.w, s3.x); st.i1 = i1; st.i2 = i2; st.i3 = i3;
  const float4 q0 = __ldg(sc.ttris + 4 * (size_t)p), q1 = __ldg(sc.ttris + 4 * (size_t)p + 1), q2 = __ldg(sc.ttris + 4 * (size_t)p + 2), q3 = __ldg(sc.ttris + 4 * (size_t)p + 3);
  TriRec tr; tr.v0 = xyz(q0); tr.e1 = xyz(q1); tr.e2 = xyz(q2); tr.Ng = xyz(q3); const int prim = __float_as_int(q0.w);
  const float ub_half = P.ub / 2.0f, lb_half = P.lb / 2.0f;
  for (int64_t s = blockIdx.y; s < P.L; s += gridDim.y) {
    const f3 o = xyz(__ldg(P.origin + s)), on = xyz(__ldg(P.onormal + s));
    for (int k = 0; k < P.spp; ++k) {
      SampleGeom g;
      if (!sample_self_hit(P.seed, P.src_offset + s, prim, k, o, st, tr, g)) continue;
      if (!(g.r <= ub_half && g.r >= lb_half)) continue;
      const Ray ray = make_ray(o, g.d);
      if (occluded(sc.nodes, sc.ttris, sc.root_count, ray, g.t, prim)) continue;
      const f3 n = st.nf, d = g.d; const float alb = 1.f, hl = g.r;
      float c2 = dot3(on, d), c3 = dot3(n, -d);
      if (c2 < 0) c2 = 0; if (c3 < 0) c3 = 0;
      const float ff = c2 * c3 / hl / hl;
      const double inten = alb * ff * ff;
      f3 t1 = (2 * alb * c2 * c3) * (on * c3 - n * c2 + (4 * (-d)) * c2 * c3);            // TG.cpp:793
      const float hl2 = hl * hl, hl4 = hl2 * hl2;
      t1 = t1 / (hl4 * hl);
      f3 gn = ((-2 * alb) * d) * c3 * c2 * c2; gn = gn / hl4;                              // :799-802 (always on here)
      const float ct = dot3(gn, n); gn = gn - n * ct;
      f3 t2 = n * (float)inten; t2 = (t2 + gn) / (2 * st.A);
      f3 e; float bk;
      if (vertex_num == i1) { e = st.v3 - st.v2; bk = g.u; } else if (vertex_num == i2) { e = st.v1 - st.v3; bk = g.v; } else { e = st.v2 - st.v1; bk = g.w; }
      const f3 xe = cross3(t2, e);
      for (int i = 0; i < P.K; ++i) {                                                    // :809-834 (literal tap loop)
        const double delta = (double)(((float)(-2 * P.r_grad * P.s_bin + i) * P.res) / (float)P.r_grad);
        const f3 gg = (float)(delta / sigma2 * 2) * d;
        const int64_t bin = (int64_t)floor(((double)(2.0f * hl) + delta - (double)P.lb) / (double)P.res);
        if (bin < 0 || bin >= P.numBins) continue;
        f3 gv = (t1 + gg * (float)inten) * bk + xe; gv = gv * (float)taps[i];
        atomicAdd(acc + 3 * bin, (double)(st.A * gv.x) / (double)P.spp);
        atomicAdd(acc + 3 * bin + 1, (double)(st.A * gv.y) / (double)P.spp);
        atomicAdd(acc + 3 * bin + 2, (double)(st.A * gv.z) / (double)P.spp);
      }
    }
  }
}

// ---- regularisers
__global__ void k_face_normal_area(const float* __restrict__ verts, const int* __restrict__ faces, int F, float4* __restrict__ na /*n.xyz, A*/) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const f3 a = ldv3(verts, faces[3 * (size_t)f]), b = ldv3(verts, faces[3 * (size_t)f + 1]), c = ldv3(verts, faces[3 * (size_t)f + 2]);
  f3 N = cross3(b - a, c - a); const float A = len3(N) / 2; N = N / (2 * A);             // SSG.cpp:67-73
  na[f] = make_float4(N.x, N.y, N.z, A);
}
__global__ void k_vertex_owner(const int* __restrict__ faces, int F, int* __restrict__ owner) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  atomicMax(owner + faces[3 * (size_t)f], f); atomicMax(owner + faces[3 * (size_t)f + 1], f); atomicMax(owner + faces[3 * (size_t)f + 2], f);
}
// MODE 0: normal smoothing (SSG.cpp:77-124), MODE 1: area gradient (SSG.cpp:27-57)
template <int MODE>
__global__ void k_regulariser(const float* __restrict__ verts, const int* __restrict__ faces, int F, const float4* __restrict__ na,
                              const int* __restrict__ aff, const int* __restrict__ owner, double* __restrict__ grad, double* __restrict__ value) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  double val = 0.0;
  if (f < F) {
    const float4 me = na[f];
    f3 fn = mk3(me.x, me.y, me.z);
    if (MODE == 0) {
      f3 n = fn * me.w;
      for (int i = 0; i < 3; ++i) { const int g = aff[3 * (size_t)f + i]; if (g < 0) continue; const float4 o = na[g]; n = n + mk3(o.x, o.y, o.z) * o.w; }
      n = n / len3(n);
      val = (double)me.w * (double)(1 - dot3(n, fn));                                    // SSG.cpp:94
      fn = fn - n;
    }
    const int i1 = faces[3 * (size_t)f], i2 = faces[3 * (size_t)f + 1], i3 = faces[3 * (size_t)f + 2];
    const f3 v1 = ldv3(verts, i1), v2 = ldv3(verts, i2), v3 = ldv3(verts, i3);
    if (owner[i1] == f) { const f3 g = cross3(fn, (v3 - v2) / 2); grad[3 * (size_t)i1] = g.x; grad[3 * (size_t)i1 + 1] = g.y; grad[3 * (size_t)i1 + 2] = g.z; }
    if (owner[i2] == f) { const f3 g = cross3(fn, (v1 - v3) / 2); grad[3 * (size_t)i2] = g.x; grad[3 * (size_t)i2 + 1] = g.y; grad[3 * (size_t)i2 + 2] = g.z; }
    if (owner[i3] == f) { const f3 g = cross3(fn, (v2 - v1) / 2); grad[3 * (size_t)i3] = g.x; grad[3 * (size_t)i3 + 1] = g.y; grad[3 * (size_t)i3 + 2] = g.z; }
  }
  if (MODE == 0) {
#pragma unroll
    for (int o = 16; o; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    if ((threadIdx.x & 31) == 0 && val != 0.0) atomicAdd(value, val);
  }
}

// ---- embree_intersector API: batched nearest-hit queries on the same LBVH
// MODE 0: intersect[3i..] = (primID, u, v) or intersect[3i] = -1 (c_embree_intersector.cpp:20-46);  MODE 1: intersect[i] = primID or -1 (:49-75)
template <int MODE>
__global__ void k_ray_query(const DeviceScene sc, const float* __restrict__ origins, const float* __restrict__ dirs, int64_t N, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const f3 o = mk3(__ldg(origins + 3 * i), __ldg(origins + 3 * i + 1), __ldg(origins + 3 * i + 2));
    const f3 d = mk3(__ldg(dirs + 3 * i), __ldg(dirs + 3 * i + 1), __ldg(dirs + 3 * i + 2));
    const Ray ray = make_ray(o, d);
    const HitRec h = nearest_hit(sc.nodes, sc.ttris, sc.root_count, ray);
    if (MODE == 1) out[i] = h.prim < 0 ? -1.0f : (float)h.prim;
    else if (h.prim < 0) out[3 * i] = -1.0f;
    else { out[3 * i] = (float)h.prim; out[3 * i + 1] = h.u; out[3 * i + 2] = h.v; }
  }
}
// c_embree_intersector.cpp:77-96 coord_conversion
__global__ void k_bary_to_world(const float* __restrict__ verts, const int* __restrict__ faces, const float* __restrict__ bary, int64_t N, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int fid = (int)bary[3 * i];
    if (fid < 0) continue;
    const float u = bary[3 * i + 1], v = bary[3 * i + 2];
    const int v1 = faces[3 * (size_t)fid], v2 = faces[3 * (size_t)fid + 1], v3 = faces[3 * (size_t)fid + 2];
    for (int k = 0; k < 3; ++k) out[3 * i + k] = (1 - u - v) * verts[3 * (size_t)v1 + k] + u * verts[3 * (size_t)v2 + k] + v * verts[3 * (size_t)v3 + k];
  }
}

}  // namespace

void launch_ray_query(Ctx& cx, const DeviceScene& sc, int mode, const float* origins, const float* dirs, int64_t N, float* out) {
  if (N <= 0) return;
  const int blocks = (int)std::min<int64_t>((N + 127) / 128, 148 * 32);
  if (mode == 1) k_ray_query<1><<<blocks, 128, 0, cx.stream>>>(sc, origins, dirs, N, out);
  else k_ray_query<0><<<blocks, 128, 0, cx.stream>>>(sc, origins, dirs, N, out);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
}
void launch_bary_to_world(Ctx& cx, const float* verts, const int* faces, const float* bary, int64_t N, float* out) {
  if (N <= 0) return;
  k_bary_to_world<<<(int)std::min<int64_t>((N + 255) / 256, 148 * 16), 256, 0, cx.stream>>>(verts, faces, bary, N, out);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
}

void launch_vertex_gradient(Ctx& cx, const DeviceScene& sc, const RenderParams& P, int vertex_num, const double* taps, double sigma2, double* acc) {
  if (sc.F <= 0 || P.L <= 0) return;
  const dim3 grid((unsigned)((sc.F + 127) / 128), (unsigned)std::min<int64_t>(P.L, 1024), 1);
  k_vertex_gradient<<<grid, 128, 0, cx.stream>>>(sc, P, vertex_num, taps, sigma2, acc);
  cx.launches += 1;
  NLOS_CUDA_OK(cudaGetLastError());
}

void launch_regulariser(Ctx& cx, int mode, const float* verts, int V, const int* faces, int F, const int* aff, double* grad, double* value) {
  cudaStream_t st = cx.stream;
  NLOS_CUDA_OK(cudaMemsetAsync(grad, 0, 3 * (size_t)V * sizeof(double), st));            // SSG.cpp:141 / :167
  if (value) NLOS_CUDA_OK(cudaMemsetAsync(value, 0, sizeof(double), st));
  if (F <= 0) return;
  float4* na = cx.buf("reg_na").as<float4>(F);
  int* owner = cx.buf("reg_owner").as<int>(std::max(V, 1));
  NLOS_CUDA_OK(cudaMemsetAsync(owner, 0xff, (size_t)V * sizeof(int), st));               // -1
  const int blocks = (F + 255) / 256;
  k_face_normal_area<<<blocks, 256, 0, st>>>(verts, faces, F, na);
  k_vertex_owner<<<blocks, 256, 0, st>>>(faces, F, owner);
  if (mode == 0) k_regulariser<0><<<blocks, 256, 0, st>>>(verts, faces, F, na, aff, owner, grad, value);
  else k_regulariser<1><<<blocks, 256, 0, st>>>(verts, faces, F, na, aff, owner, grad, value);
  cx.launches += 3;
  NLOS_CUDA_OK(cudaGetLastError());
}

}  // namespace nlos
