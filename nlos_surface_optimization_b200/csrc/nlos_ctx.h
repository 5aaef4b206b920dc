// nlos_ctx.h — host-side context behind the opaque nlos_ctx handle of include/nlos_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <map>
#include "nlos_device.cuh"

namespace nlos {

// grow-only device buffer; reused across calls so a steady-state iteration does no cudaMalloc
struct DevBuf {
  void* p = nullptr; size_t cap = 0;
  void* ensure(size_t bytes) {
    if (bytes > cap) {
      if (p) cudaFree(p);
      p = nullptr; cap = 0;
      size_t want = bytes + bytes / 8 + 256;
      cudaError_t e = cudaMalloc(&p, want);
      if (e != cudaSuccess) { p = nullptr; throw std::runtime_error(std::string("cudaMalloc failed: ") + cudaGetErrorString(e)); }
      cap = want;
    }
    return p;
  }
  template <class T> T* as(size_t n) { return reinterpret_cast<T*>(ensure(n * sizeof(T))); }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct Timing { float build_ms = 0, forward_ms = 0, residual_ms = 0, gradient_ms = 0, total_ms = 0; };

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev[8] = {};
  cudaEvent_t ev_copy = nullptr;   // data/weight staged on copy_stream
  cudaEvent_t ev_start = nullptr;  // start of the current call on the main stream
  cudaEvent_t ev_ext = nullptr;    // ordering against a caller's stream (nlos_ctx_wait_stream / nlos_ctx_signal_stream)
  cudaEvent_t ev_fwd = nullptr;    // forward transient final (its D2H may start while the gradient runs)
  std::string last_error;
  // options (nlos_ctx_set_*)
  uint64_t seed = 5489;          // boost::mt19937 default = the reference's built-in seed (sampler.cpp:25)
  int64_t src_offset = 0;        // global index of the first source of this call (sharded runs)
  int64_t num_sources_global = 0;  // 0 => normalise the gradient by this call's L (reference behaviour)
  int reuse_visibility = 1;      // gradient pass consumes the forward pass's visibility bits
  int64_t ext_count = 0;         // test hook: floats of the external (S,T) sample stream held in buf("ext_samples"), 0 = Philox
  int timing_enabled = 0;
  int chunk_forward = 0;         // sources per blockIdx.y in the forward pass (0 = auto)
  int chunk_gradient = 0;        // same for the gradient pass
  int forward_algo = 0;          // 0 auto (shared perspective grid where it applies), 1 BVH traversal kernel, 2 per-wall-point perspective grid, 3 shared perspective grid
  int group_side = 0;            // shared grid: a group is a tile of side x side wall spacings (0 = 4; 1 = one wall point per group)
  int grid_slices = 0;           // shared grid: slices of 1/Z (0 = 16)
  int grid_budget_mb = 0;        // shared grid: scratch memory for the lists of one batch of groups (0 = 6144 MiB)
  int grid_res = 0;              // cells per axis of the perspective grid (0 = auto from the triangle count)
  int count_work = 0;            // measurement: the next perspective-grid forward launch fills buf("work_counters") (nlos_ctx_get_work_counters)
  int work_G = 0;                // grid resolution of that launch
  int last_forward_algo = 0;     // forward kernel of the last call: 1 BVH traversal, 2 perspective grid
  int last_grid_res = 0;
  int grid_cap = 0;              // test hook: upper bound of the per-source entry budget of the perspective grid (0 = none); forces the coarsening path
  size_t vis_words = 0;          // words of the last call's visibility buffer (nlos_debug_copy_visibility_words)
  int num_sms = 0;               // multiprocessors of the device (filled at context creation)
  Timing timing;
  uint64_t launches = 0;         // kernels launched by this context (bench.py's gpu_launches)
  std::map<std::string, DevBuf> bufs;
  DevBuf& buf(const char* name) { return bufs[name]; }
  ~Ctx();
};

// lbvh.cu
void build_scene(Ctx& cx, const float* d_verts, int V, const int* d_faces, int F, const float* d_origin, int64_t L,
                 const float* d_vnormal, const float* d_valbedo, DeviceScene& out);

}  // namespace nlos

// the opaque handle of include/nlos_b200.h
struct nlos_ctx { nlos::Ctx cx; };
