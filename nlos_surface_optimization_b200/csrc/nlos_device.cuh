// nlos_device.cuh — device-side data model shared by the .cu files of libnlos_b200.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include "nlos_core.cuh"

namespace nlos {

struct SceneBounds { unsigned lo[3], hi[3]; unsigned absmax; unsigned bad_faces; unsigned vlo[3], vhi[3]; };   // centroid bounds, max |coordinate|, count of out-of-range face indices (clamped by the build kernels), vertex bounds (ordered-uint floats)

// Mesh + acceleration structure resident in HBM for the duration of one call (rebuilt per call like the
// reference rebuilds its Embree scene, SSG.cpp:473-511).  All per-triangle arrays are in Morton order.
struct DeviceScene {
  int F = 0, V = 0;
  int root_count = 0;           // >0: the whole mesh is one leaf run (F <= kLeafMax)
  const float4* ttris = nullptr;   // [F][4]  TraceTri
  const float4* stris = nullptr;   // [4][F]  ShadeTri, component-major: (v1,A) (v2,nf.x) (v3,nf.y) (nf.z,i1,i2,i3)
  const int* sprim = nullptr;      // [F]     caller's index of the triangle at Morton position p (== ttris[4p].w)
  const BvhNode* nodes = nullptr;  // [max(F-1,1)]
  const float* vnormal = nullptr;  // [V,3] or null (caller order)
  const float* valbedo = nullptr;  // [V]   or null
  const float* verts = nullptr;    // [V,3] caller order
  const SceneBounds* bounds = nullptr;   // device copy of the scene bounds (perspective-grid forward kernel)
};


// ordered-uint encoding so that atomicMin/atomicMax work on floats of either sign
__device__ __forceinline__ unsigned f2ord(float f) { unsigned u = (unsigned)__float_as_int(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; return __int_as_float((int)u); }

// per-call render parameters passed by value to the sample kernels
struct RenderParams {
  const float4* origin;     // [L] (x,y,z,-)
  const float4* onormal;    // [L]
  int64_t L;                // sources in this call (this rank's slice)
  int64_t src_offset;       // global index of source 0 (RNG key), for sharded runs
  uint64_t seed;
  int spp;
  float lb, ub, res;        // path-length bounds and bin width (coarse)
  float res_fwd;            // res / r_fwd (float division, TG.cpp:313)
  int numBins;              // coarse bins B
  int r_fwd;                // refine scale used by the forward histogram (1 => raw histogram)
  int r_grad, s_bin;        // refine scale / sigma_bin of the gradient taps
  int K;                    // 4*r*s+1 taps
  double inv_res_fine;      // r_grad / res  (double)
  double two_over_sigma2;   // 2 / sigma^2
  float alpha;              // GGX roughness (unused when Lambertian)
  int testing_flag;
  int chunk;                // sources per blockIdx.y
  int words_per_row;        // vis words per (source, k): ceil(F/32)
  // temporal kernel of the gradient: Gaussian taps (smoothed_transient/, ggx/) or the tabulated SPAD jitter kernel (jitter/)
  int jitter;               // 1: A/B sums come from the per-source tables jA/jB [L, numBins+1] indexed by the coarse bin
  const double* jA;         // sum_i jitter_weight[i] * (-2) diff[s, b+i-offset]
  const double* jB;         // sum_i jitter_grad[i]   * (-2) diff[s, b+i-offset]
  double grad_coef;         // factor of the kernel-derivative term: 2/sigma^2 (Gaussian) or -2/res (jitter/TG.cpp:950)
  // test hook (nlos_ctx_set_external_samples): when non-null the (S,T) pair of sample k of (global source s, triangle f) is read from
  // ext_samples[ext_base + 2*((s*F + f)*spp + k)] instead of being drawn from Philox — the order in which ONE worker of the reference
  // consumes its stream, so the CUDA path can be compared with the reference's own outputs sample for sample
  const float* ext_samples; int64_t ext_count; int64_t ext_base; int F;
  int sr;                   // 1: first-generation renderer (stratified_transient_raytracer/): forward without the form-factor clamp
                            //    (SR/SST.cpp:130-137), gradient with the normal-variation term always on (SR/SSG.cpp:266-271)
};

struct Status { int code = 0; std::string msg; };

#define NLOS_CUDA_OK(expr)                                                                          \
  do {                                                                                              \
    cudaError_t e__ = (expr);                                                                       \
    if (e__ != cudaSuccess) {                                                                       \
      char b__[512];                                                                                \
      snprintf(b__, sizeof b__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
      throw std::runtime_error(b__);                                                                \
    }                                                                                               \
  } while (0)

}  // namespace nlos
