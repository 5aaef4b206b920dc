// nlos_abi.cu — the extern "C" boundary declared in include/nlos_b200.h and the per-call orchestration
// that replaces the reference's drivers:
//   smoothed_transient/stratifiedStreamedTransientRenderer.cpp:20-153   (transient, intensity)
//   smoothed_transient/stratifiedStreamedGradientRenderer.cpp:183-578   (gradient, w_albedo, albedo)
//   ggx/stratifiedStreamedTransientRenderer.cpp, ggx/stratifiedStreamedGradientRenderer.cpp:27-239
// Per call: stage inputs -> K0 scene build -> K1 forward (+visibility bits) -> K3 residual -> K4/K5
// gradient -> finalize -> copy results back.  No CPU compute path exists in this library.
#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges around the phases of a call (visible to profilers, no-ops otherwise)
#include <algorithm>
#include <cstring>
#include <vector>
#include "../../include/nlos_b200.h"
#include "nlos_ctx.h"
#include "render_kernels.h"


namespace nlos {

Ctx::~Ctx() {
  if (stream) { cudaSetDevice(device); cudaStreamSynchronize(stream); }
  for (auto& kv : bufs) kv.second.release();
  for (auto& e : ev) if (e) cudaEventDestroy(e);
  if (ev_copy) cudaEventDestroy(ev_copy);
  if (ev_fwd) cudaEventDestroy(ev_fwd);
  if (ev_start) cudaEventDestroy(ev_start);
  if (ev_ext) cudaEventDestroy(ev_ext);
  if (copy_stream) cudaStreamDestroy(copy_stream);
  if (stream) cudaStreamDestroy(stream);
}

namespace {

std::string g_create_error;

struct InvalidArg : std::runtime_error { using std::runtime_error::runtime_error; };
#define NLOS_REQUIRE(cond, msg) do { if (!(cond)) throw InvalidArg(msg); } while (0)

bool is_device_ptr(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) { cudaGetLastError(); return false; }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// device view of an input array: the pointer itself when it already lives in HBM, else a staged copy
template <class T>
const T* stage_in(Ctx& cx, const char* name, const T* p, size_t n, cudaStream_t st) {
  if (!p || n == 0) return p;
  if (is_device_ptr(p)) return p;
  T* d = cx.buf(name).as<T>(n);
  NLOS_CUDA_OK(cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, st));
  return d;
}
// device view of an output array; host outputs get a scratch buffer that is copied back by finish_out
template <class T>
struct OutView { T* dev = nullptr; T* host = nullptr; size_t n = 0; };
template <class T>
OutView<T> stage_out(Ctx& cx, const char* name, T* p, size_t n, bool upload_current) {
  OutView<T> v; v.n = n;
  if (!p || n == 0) return v;
  if (is_device_ptr(p)) { v.dev = p; return v; }
  v.host = p; v.dev = cx.buf(name).as<T>(n);
  if (upload_current) NLOS_CUDA_OK(cudaMemcpyAsync(v.dev, p, n * sizeof(T), cudaMemcpyHostToDevice, cx.stream));
  return v;
}
template <class T>
bool finish_out(Ctx& cx, const OutView<T>& v, cudaStream_t on = nullptr) {
  if (!v.host) return false;
  NLOS_CUDA_OK(cudaMemcpyAsync(v.host, v.dev, v.n * sizeof(T), cudaMemcpyDeviceToHost, on ? on : cx.stream));
  return true;
}

// Every face index must address a vertex (the reference reads out of bounds silently, TG.cpp:146-150).  Host arrays are checked here,
// before anything is launched.  Device arrays cannot be checked without a round trip: the scene-build kernels clamp such indices (no
// out-of-bounds access anywhere downstream, the clamped indices are what every later kernel sees) and count them; the count is read
// back wherever the call synchronises anyway, and the call then fails with NLOS_ERR_INVALID.
void require_valid_host_faces(const int* faces_arg, int F, int V) {
  if (F <= 0 || is_device_ptr(faces_arg)) return;
  for (size_t i = 0; i < 3 * (size_t)F; ++i) NLOS_REQUIRE(faces_arg[i] >= 0 && faces_arg[i] < V, "faces: vertex index out of range");
}
void require_no_clamped_faces(Ctx& cx, const DeviceScene& sc) {      // call only after the stream has been synchronised
  if (!sc.bounds) return;
  SceneBounds h;
  NLOS_CUDA_OK(cudaMemcpy(&h, sc.bounds, sizeof h, cudaMemcpyDeviceToHost));
  NLOS_REQUIRE(h.bad_faces == 0, "faces: vertex index out of range (results of this call are invalid)");
}

// Gaussian taps exactly as the reference builds them (TG.cpp:537-544, :350-355) plus their prefix sums
struct TapTables { int K = 1; double sigma2 = 1; std::vector<double> w, wprefix, dprefix; };
TapTables make_taps(float res, int r, int s) {
  TapTables t; t.K = 4 * r * s + 1;
  const double sigma = res * s / 2.355;
  t.sigma2 = sigma * sigma;
  const double norm = 1 / sigma / std::sqrt(2 * M_PI) * res / r;
  t.wprefix.assign(t.K + 1, 0.0); t.dprefix.assign(t.K + 1, 0.0);
  for (int i = 0; i < t.K; ++i) {
    const double x = (-2 * r * s + i) * res / r / sigma;
    const double w = std::exp(-(x * x) / 2) * norm;
    const double delta = (double)(((float)(-2 * r * s + i) * res) / (float)r);     // TG.cpp:973 (float expression)
    t.w.push_back(w);
    t.wprefix[i + 1] = t.wprefix[i] + w;
    t.dprefix[i + 1] = t.dprefix[i] + w * delta;
  }
  return t;
}

struct Job {
  const float* origin = nullptr; const float* onormal = nullptr; int64_t L = 0;
  const float* verts = nullptr; int V = 0; const float* vn = nullptr; const float* va = nullptr;
  const int* faces = nullptr; int F = 0;
  bool ggx = false; float alpha = 0.f;
  int num_samples = 1; float lb = 0, ub = 0, res = 1; int numBins = 0;
  int refine = 1, sigma = 1, testing_flag = 1, loss_flag = 0;
  const double* data = nullptr; const double* weight = nullptr;
  double* transient = nullptr; double* pathlengths = nullptr; double* gradient = nullptr; double* intensity = nullptr;
  double* scalar_out = nullptr;
  int kind = -1;   // -1 forward only, 0 vertex gradient, 1 albedo scalar, 2 alpha scalar, 3 intensity
  const double* jw = nullptr; const double* jg = nullptr; int joff = 0, jlen = 0;   // jitter/ temporal kernel (jlen > 0)
  int sr = 0, w_width = 0;   // first-generation API (stratified_transient_raytracer/): unclamped forward, box-filtered residual, one tap, '=' into gradient
};

// gradient kernel: sources per block (blockIdx.y).  The default (128) amortises the 9 atomics per (triangle, chunk); small meshes
// get a smaller chunk so that the grid is MANY waves of the 148 x 7 resident blocks, not one and a bit (F = 1125, L = 4096: chunk 31 -> 1197
// blocks = 1.16 waves 2.79 ms; chunk 8 -> 4608 blocks 2.31 ms; chunk 4 -> 9216 blocks 2.24 ms; tools/sweep_chunk.py).
int auto_chunk(const Ctx& cx, const char* key, int F, int64_t L, int dflt) {
  (void)key;
  int c = dflt;
  if (cx.chunk_gradient <= 0) {
    const int64_t nx = ((int64_t)F + 127) / 128, target = 148 * 7 * 8;  // 128 = threads per block of k_gradient, 7 resident blocks per SM, 8 waves
    const int64_t ny = (target + nx - 1) / nx;
    const int64_t fit = std::max<int64_t>(4, L / std::max<int64_t>(ny, 1));
    if (fit < c) c = (int)fit;
  }
  if ((int64_t)c > L) c = (int)std::max<int64_t>(L, 1);
  // gridDim.y <= 65535
  while ((L + c - 1) / c > 65535) c *= 2;
  return c;
}

// forward kernel: sample slots (source*spp + k) per warp pass, bounded by the visibility tile in shared memory
int forward_chunk(const Ctx& cx) { int c = cx.chunk_forward > 0 ? cx.chunk_forward : 64; return std::min(std::max(c, 1), forward_max_chunk()); }

void run_job(Ctx& cx, const Job& j_in) {
  Job j = j_in;
  // sample kernels: the production build, or (test hook active) the build that reads the external (S,T) stream
  const bool use_ext = cx.ext_count > 0;
  auto fwd = use_ext ? ext::launch_forward : launch_forward;
  auto inten = use_ext ? ext::launch_intensity : launch_intensity;
  auto grad = use_ext ? ext::launch_gradient : launch_gradient;
  if (j.sr && j.kind == 0) {
    // the first-generation gradient is the tabulated-kernel path with the unit kernel: A[b] = -2 diff[b], B[b] = 0
    static const double kOne = 1.0, kZero = 0.0;
    j.jw = &kOne; j.jg = &kZero; j.jlen = 1; j.joff = 0; j.refine = 1; j.sigma = 1;
  }
  NLOS_CUDA_OK(cudaSetDevice(cx.device));
  NLOS_REQUIRE(j.L >= 0 && j.V >= 0 && j.F >= 0, "negative size");
  NLOS_REQUIRE(j.L == 0 || (j.origin && j.onormal), "origin/normal is null");
  NLOS_REQUIRE(j.F == 0 || (j.verts && j.faces), "vertices/faces is null");
  if (j.kind != 3) {
    NLOS_REQUIRE(j.numBins > 0, "numBins must be positive");
    NLOS_REQUIRE(j.res > 0.f, "pathlengthResolution must be positive");
    NLOS_REQUIRE(j.refine >= 1 && j.sigma >= 1, "refine_scale and sigma_bin must be >= 1");
    NLOS_REQUIRE(j.transient != nullptr, "transient is null");
  }
  if (j.kind >= 0 && j.kind <= 2) NLOS_REQUIRE(j.data && (j.weight || j.sr), "data/weight is null");
  if (j.sr) NLOS_REQUIRE(j.w_width >= 0 && j.kind <= 0 && !j.ggx, "first-generation API: w_width must be >= 0");
  if (j.kind == 0) NLOS_REQUIRE(j.gradient != nullptr, "gradient is null");
  if (j.kind == 3) NLOS_REQUIRE(j.intensity != nullptr, "intensity is null");

  cudaStream_t st = cx.stream;
  const bool timing = cx.timing_enabled != 0;
  if (timing) NLOS_CUDA_OK(cudaEventRecord(cx.ev[0], st));
  // the copy stream may only touch the staging buffers after every EARLIER call on the main stream has finished with them
  NLOS_CUDA_OK(cudaEventRecord(cx.ev_start, st));
  NLOS_CUDA_OK(cudaStreamWaitEvent(cx.copy_stream, cx.ev_start, 0));
  const size_t LB = (size_t)j.L * (size_t)std::max(j.numBins, 0);

  // ---- stage inputs
  const float* d_origin = stage_in(cx, "in_origin", j.origin, 3 * (size_t)j.L, st);
  const float* d_onormal = stage_in(cx, "in_onormal", j.onormal, 3 * (size_t)j.L, st);
  const float* d_verts = stage_in(cx, "in_verts", j.verts, 3 * (size_t)j.V, st);
  const int* d_faces = stage_in(cx, "in_faces", j.faces, 3 * (size_t)j.F, st);
  const float* d_vn = stage_in(cx, "in_vn", j.vn, 3 * (size_t)j.V, st);
  const float* d_va = stage_in(cx, "in_va", j.va, (size_t)j.V, st);

  // ---- outputs
  OutView<double> o_T, o_pl, o_G, o_I;
  if (j.kind != 3) {
    o_T = stage_out(cx, "out_transient", j.transient, LB, false);
    o_pl = stage_out(cx, "out_pathlengths", j.pathlengths, (size_t)j.numBins, false);
    if (o_T.dev) NLOS_CUDA_OK(cudaMemsetAsync(o_T.dev, 0, LB * sizeof(double), st));     // TG.cpp:291
    if (o_pl.dev) launch_pathlengths(cx, o_pl.dev, j.numBins, j.lb, j.res);
  }
  if (j.kind == 0) o_G = stage_out(cx, "out_gradient", j.gradient, 3 * (size_t)j.V, true);
  if (j.kind == 3) o_I = stage_out(cx, "out_intensity", j.intensity, (size_t)j.F, true);

  bool need_sync = false, fwd_recorded = false;
  if (j.F > 0 && j.L > 0) require_valid_host_faces(j.faces, j.F, j.V);
  DeviceScene sc;
  if (j.F > 0 && j.L > 0) {
    // ---- K0: scene
    nvtxRangePushA("nlos: scene build");
    build_scene(cx, d_verts, j.V, d_faces, j.F, d_origin, j.L, d_vn, d_va, sc);
    float4* origin4 = cx.buf("origin4").as<float4>((size_t)j.L);
    float4* onormal4 = cx.buf("onormal4").as<float4>((size_t)j.L);
    launch_pack4(cx, d_origin, origin4, (size_t)j.L);
    launch_pack4(cx, d_onormal, onormal4, (size_t)j.L);
    nvtxRangePop();
    if (timing) NLOS_CUDA_OK(cudaEventRecord(cx.ev[1], st));

    RenderParams P;
    std::memset(&P, 0, sizeof P);
    P.origin = origin4; P.onormal = onormal4; P.L = j.L; P.src_offset = cx.src_offset; P.seed = cx.seed;
    P.spp = 1 + (j.num_samples - 1) / j.F;                                              // TG.cpp:289
    if (P.spp < 1) P.spp = 1;
    P.lb = j.lb; P.ub = j.ub; P.res = j.res; P.numBins = std::max(j.numBins, 1);
    P.r_grad = j.refine; P.s_bin = j.sigma; P.K = 4 * j.refine * j.sigma + 1;
    P.r_fwd = (j.kind >= 0 && j.kind <= 2) ? (j.sigma < 5 ? 1 : j.refine) : j.refine;   // SSG.cpp:521-524
    if (j.kind == 3 || j.jlen > 0) P.r_fwd = 1;                                          // jitter: coarse histogram, then tabulated conv
    P.res_fwd = j.res / P.r_fwd;                                                         // TG.cpp:313
    P.alpha = j.alpha; P.testing_flag = j.testing_flag; P.sr = j.sr;
    P.F = j.F;
    if (cx.ext_count > 0) { P.ext_samples = cx.buf("ext_samples").as<float>((size_t)cx.ext_count); P.ext_count = cx.ext_count; }
    P.words_per_row = (j.F + 31) / 32;
    TapTables taps;
    if (j.kind != 3) {
      taps = make_taps(j.res, j.refine, j.sigma);
      P.inv_res_fine = (double)j.refine / (double)j.res;
      P.two_over_sigma2 = 2.0 / taps.sigma2;
      P.grad_coef = j.jlen > 0 ? -2.0 / (double)j.res : P.two_over_sigma2;
    }
    double* d_wprefix = nullptr; double* d_dprefix = nullptr;
    if (j.kind != 3) {
      d_wprefix = cx.buf("wprefix").as<double>(taps.wprefix.size());
      d_dprefix = cx.buf("dprefix").as<double>(taps.dprefix.size());
      NLOS_CUDA_OK(cudaMemcpyAsync(d_wprefix, taps.wprefix.data(), taps.wprefix.size() * sizeof(double), cudaMemcpyHostToDevice, st));
      NLOS_CUDA_OK(cudaMemcpyAsync(d_dprefix, taps.dprefix.data(), taps.dprefix.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    }

    if (j.kind == 3) {
      P.chunk = forward_chunk(cx);
      inten(cx, sc, P, j.ggx, o_I.dev);
      if (timing) { NLOS_CUDA_OK(cudaEventRecord(cx.ev[2], st)); NLOS_CUDA_OK(cudaEventRecord(cx.ev[3], st)); NLOS_CUDA_OK(cudaEventRecord(cx.ev[4], st)); }
    } else {
      // ---- K1: forward (+ visibility bits for the gradient pass)
      nvtxRangePushA("nlos: forward");
      uint32_t* vis = nullptr;
      const bool want_grad = j.kind >= 0 && j.kind <= 2;
      cx.vis_words = 0;
      if (want_grad && cx.reuse_visibility) {
        // one bit per sample; when that buffer cannot be had (very large L * spp * F) the gradient pass re-traces its rays instead
        const size_t words = (size_t)j.L * P.spp * P.words_per_row;
        try { vis = cx.buf("vis").as<uint32_t>(words); cx.vis_words = words; }
        catch (const std::exception&) { cudaGetLastError(); vis = nullptr; cx.vis_words = 0; }
      }
      P.chunk = forward_chunk(cx);                                                       // sample slots per warp pass
      const double* d_jw = nullptr; const double* d_jg = nullptr;
      if (j.jlen > 0) {
        // jitter/TG.cpp:271-356: raw coarse histogram, then T = hist (*) jitter_weight shifted by weight_offset
        d_jw = stage_in(cx, "in_jw", j.jw, (size_t)j.jlen, st);
        d_jg = stage_in(cx, "in_jg", j.jg, (size_t)j.jlen, st);
        double* hist = cx.buf("jit_hist").as<double>(LB);
        NLOS_CUDA_OK(cudaMemsetAsync(hist, 0, LB * sizeof(double), st));
        fwd(cx, sc, P, j.ggx, hist, vis, d_wprefix);
        launch_jitter_conv(cx, hist, d_jw, j.jlen, j.joff, j.numBins, j.L, o_T.dev);
      } else
      fwd(cx, sc, P, j.ggx, o_T.dev, vis, d_wprefix);
      nvtxRangePop();
      if (timing) NLOS_CUDA_OK(cudaEventRecord(cx.ev[2], st));
      NLOS_CUDA_OK(cudaEventRecord(cx.ev_fwd, st)); fwd_recorded = true;        // the transient is final here
      if (want_grad) {
        // ---- K3: residual
        nvtxRangePushA("nlos: residual + gradient");
        // host data/weight ride the copy stream while the forward kernel (already enqueued) runs
        const double* d_data = stage_in(cx, "in_data", j.data, LB, cx.copy_stream);
        const double* d_weight = j.weight ? stage_in(cx, "in_weight", j.weight, LB, cx.copy_stream) : nullptr;
        NLOS_CUDA_OK(cudaEventRecord(cx.ev_copy, cx.copy_stream));
        NLOS_CUDA_OK(cudaStreamWaitEvent(st, cx.ev_copy, 0));
        double* diff = cx.buf("diff").as<double>(LB);
        launch_residual(cx, d_data, d_weight, o_T.dev, diff, LB, j.loss_flag);
        if (j.sr && j.w_width > 0) {                                                      // SR/SSG.cpp:447-458: box mean twice
          double* tmp = cx.buf("diff_tmp").as<double>(LB);
          launch_box_filter(cx, diff, tmp, j.numBins, j.L, j.w_width);
          launch_box_filter(cx, tmp, diff, j.numBins, j.L, j.w_width);
        }
        if (timing) NLOS_CUDA_OK(cudaEventRecord(cx.ev[3], st));
        // ---- K4/K5: gradient
        P.chunk = auto_chunk(cx, "chunk_gradient", j.F, j.L, cx.chunk_gradient > 0 ? cx.chunk_gradient : 128);
        const int64_t Lnorm = cx.num_sources_global > 0 ? cx.num_sources_global : j.L;
        if (j.jlen > 0) {
          double* jA = cx.buf("jit_A").as<double>((size_t)j.L * (j.numBins + 1));
          double* jB = cx.buf("jit_B").as<double>((size_t)j.L * (j.numBins + 1));
          launch_jitter_tables(cx, diff, d_jw, d_jg, j.jlen, j.joff, j.numBins, j.L, jA, jB);
          P.jitter = 1; P.jA = jA; P.jB = jB;
        }
        if (j.sr && P.ext_samples) P.ext_base = 2 * (cx.num_sources_global > 0 ? cx.num_sources_global : j.L) * (int64_t)j.F * P.spp;   // SR/SSG.cpp:400-470: one sampler set for both passes
        if (j.kind == 0) {
          double* acc = cx.buf("grad_acc").as<double>(3 * (size_t)j.V);
          NLOS_CUDA_OK(cudaMemsetAsync(acc, 0, 3 * (size_t)j.V * sizeof(double), st));
          grad(cx, sc, P, j.ggx, 0, diff, vis, d_wprefix, d_dprefix, acc);
          if (j.sr) NLOS_CUDA_OK(cudaMemsetAsync(o_G.dev, 0, 3 * (size_t)j.V * sizeof(double), st));   // SR/SSG.cpp:414: cleared, not accumulated
          launch_finalize_gradient(cx, acc, o_G.dev, 3 * (size_t)j.V, 1.0 / (double)Lnorm);
        } else {
          double* acc = cx.buf("scalar_acc").as<double>(1);
          NLOS_CUDA_OK(cudaMemsetAsync(acc, 0, sizeof(double), st));
          grad(cx, sc, P, j.ggx, j.kind, diff, vis, d_wprefix, d_dprefix, acc);
          double h = 0;
          NLOS_CUDA_OK(cudaMemcpyAsync(&h, acc, sizeof(double), cudaMemcpyDeviceToHost, st));
          NLOS_CUDA_OK(cudaStreamSynchronize(st));
          if (j.scalar_out) *j.scalar_out = h / (double)Lnorm;                           // TG.cpp:494-498
        }
        nvtxRangePop();
        if (timing) NLOS_CUDA_OK(cudaEventRecord(cx.ev[4], st));
      } else if (timing) { NLOS_CUDA_OK(cudaEventRecord(cx.ev[3], st)); NLOS_CUDA_OK(cudaEventRecord(cx.ev[4], st)); }
    }
  } else {
    if (j.scalar_out) *j.scalar_out = 0.0;
    if (timing) for (int i = 1; i <= 4; ++i) NLOS_CUDA_OK(cudaEventRecord(cx.ev[i], st));
  }

  // ---- results back to host arrays
  bool copy_sync = false;
  if (o_T.host && fwd_recorded && j.kind >= 0 && j.kind <= 2) {
    // the transient's D2H overlaps the gradient kernels (all already enqueued on the main stream)
    NLOS_CUDA_OK(cudaStreamWaitEvent(cx.copy_stream, cx.ev_fwd, 0));
    copy_sync = finish_out(cx, o_T, cx.copy_stream);
  } else need_sync |= finish_out(cx, o_T);
  need_sync |= finish_out(cx, o_pl);
  need_sync |= finish_out(cx, o_G);
  need_sync |= finish_out(cx, o_I);
  if (timing) NLOS_CUDA_OK(cudaEventRecord(cx.ev[5], st));
  if (need_sync || timing) NLOS_CUDA_OK(cudaStreamSynchronize(st));
  if (copy_sync) NLOS_CUDA_OK(cudaStreamSynchronize(cx.copy_stream));
  if ((need_sync || timing) && is_device_ptr(j.faces)) require_no_clamped_faces(cx, sc);
  if (timing) {
    cudaEventElapsedTime(&cx.timing.build_ms, cx.ev[0], cx.ev[1]);
    cudaEventElapsedTime(&cx.timing.forward_ms, cx.ev[1], cx.ev[2]);
    cudaEventElapsedTime(&cx.timing.residual_ms, cx.ev[2], cx.ev[3]);
    cudaEventElapsedTime(&cx.timing.gradient_ms, cx.ev[3], cx.ev[4]);
    cudaEventElapsedTime(&cx.timing.total_ms, cx.ev[0], cx.ev[5]);
  }
}

int guarded(nlos_ctx* ctx, const Job& j) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try { run_job(cx, j); cx.last_error.clear(); return NLOS_OK; }
  catch (const InvalidArg& e) { cx.last_error = e.what(); return NLOS_ERR_INVALID; }
  catch (const std::exception& e) {
    cx.last_error = e.what(); cudaGetLastError();
    return cx.last_error.find("cudaMalloc") != std::string::npos ? NLOS_ERR_NOMEM : NLOS_ERR_CUDA;
  }
}

Job base_job(const float* origin, int L, const float* onormal, const float* verts, int V, const float* vn, const float* va, const int* faces, int F,
             int num_samples, float lb, float ub, float res, int numBins, int refine, int sigma) {
  Job j; j.origin = origin; j.onormal = onormal; j.L = L; j.verts = verts; j.V = V; j.vn = vn; j.va = va; j.faces = faces; j.F = F;
  j.num_samples = num_samples; j.lb = lb; j.ub = ub; j.res = res; j.numBins = numBins; j.refine = refine; j.sigma = sigma;
  return j;
}

}  // namespace
}  // namespace nlos

using namespace nlos;

extern "C" {

int nlos_ctx_create(int device, nlos_ctx** out) {
  if (!out) return NLOS_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) { g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"); cudaGetLastError(); return NLOS_ERR_CUDA; }
  if (device < 0 || device >= n) { g_create_error = "device index out of range"; return NLOS_ERR_INVALID; }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return NLOS_ERR_CUDA; }
  if (prop.major != 10) { g_create_error = "libnlos_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor); return NLOS_ERR_CUDA; }
  nlos_ctx* c = new (std::nothrow) nlos_ctx();
  if (!c) return NLOS_ERR_NOMEM;
  c->cx.device = device;
  c->cx.num_sms = prop.multiProcessorCount;
  try {
    NLOS_CUDA_OK(cudaSetDevice(device));
    NLOS_CUDA_OK(cudaStreamCreateWithFlags(&c->cx.stream, cudaStreamNonBlocking));
    NLOS_CUDA_OK(cudaStreamCreateWithFlags(&c->cx.copy_stream, cudaStreamNonBlocking));
    for (auto& ev : c->cx.ev) NLOS_CUDA_OK(cudaEventCreate(&ev));
    NLOS_CUDA_OK(cudaEventCreateWithFlags(&c->cx.ev_copy, cudaEventDisableTiming));
    NLOS_CUDA_OK(cudaEventCreateWithFlags(&c->cx.ev_fwd, cudaEventDisableTiming));
    NLOS_CUDA_OK(cudaEventCreateWithFlags(&c->cx.ev_start, cudaEventDisableTiming));
    NLOS_CUDA_OK(cudaEventCreateWithFlags(&c->cx.ev_ext, cudaEventDisableTiming));
  } catch (const std::exception& ex) { g_create_error = ex.what(); delete c; return NLOS_ERR_CUDA; }
  *out = c;
  return NLOS_OK;
}

void nlos_ctx_destroy(nlos_ctx* ctx) { delete ctx; }
const char* nlos_last_error(nlos_ctx* ctx) { return ctx ? ctx->cx.last_error.c_str() : g_create_error.c_str(); }
int nlos_ctx_synchronize(nlos_ctx* ctx) {
  if (!ctx) return NLOS_ERR_INVALID;
  cudaSetDevice(ctx->cx.device);
  cudaError_t e = cudaStreamSynchronize(ctx->cx.stream);
  if (e != cudaSuccess) { ctx->cx.last_error = cudaGetErrorString(e); return NLOS_ERR_CUDA; }
  return NLOS_OK;
}
void* nlos_ctx_stream(nlos_ctx* ctx) { return ctx ? (void*)ctx->cx.stream : nullptr; }
int nlos_ctx_wait_stream(nlos_ctx* ctx, void* stream) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  if ((cudaStream_t)stream == cx.stream) return NLOS_OK;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_CUDA_OK(cudaEventRecord(cx.ev_ext, (cudaStream_t)stream));
    NLOS_CUDA_OK(cudaStreamWaitEvent(cx.stream, cx.ev_ext, 0));
    return NLOS_OK;
  } catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}
int nlos_ctx_signal_stream(nlos_ctx* ctx, void* stream) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  if ((cudaStream_t)stream == cx.stream) return NLOS_OK;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_CUDA_OK(cudaEventRecord(cx.ev_ext, cx.stream));
    NLOS_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, cx.ev_ext, 0));
    return NLOS_OK;
  } catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}
int nlos_ctx_set_seed(nlos_ctx* ctx, uint64_t seed) { if (!ctx) return NLOS_ERR_INVALID; ctx->cx.seed = seed; return NLOS_OK; }
int nlos_ctx_set_external_samples(nlos_ctx* ctx, const float* st, int64_t n) {
  if (!ctx || n < 0 || (n > 0 && !st)) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_CUDA_OK(cudaStreamSynchronize(cx.stream));
    cx.ext_count = 0;
    if (n > 0) {
      float* d = cx.buf("ext_samples").as<float>((size_t)n);
      NLOS_CUDA_OK(cudaMemcpyAsync(d, st, (size_t)n * sizeof(float), cudaMemcpyDefault, cx.stream));
      NLOS_CUDA_OK(cudaStreamSynchronize(cx.stream));
      cx.ext_count = n;
    }
    cx.last_error.clear(); return NLOS_OK;
  } catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}

int nlos_ctx_set_source_window(nlos_ctx* ctx, int64_t src_offset, int64_t num_sources_global) {
  if (!ctx || src_offset < 0 || num_sources_global < 0) return NLOS_ERR_INVALID;
  ctx->cx.src_offset = src_offset; ctx->cx.num_sources_global = num_sources_global; return NLOS_OK;
}
int nlos_ctx_set_option(nlos_ctx* ctx, const char* key, int64_t value) {
  if (!ctx || !key) return NLOS_ERR_INVALID;
  std::string k(key);
  if (k == "reuse_visibility") ctx->cx.reuse_visibility = value != 0;
  else if (k == "chunk_forward") ctx->cx.chunk_forward = (int)value;
  else if (k == "chunk_gradient") ctx->cx.chunk_gradient = (int)value;
  else if (k == "timing") ctx->cx.timing_enabled = value != 0;
  else if (k == "forward_algo") { if (value < 0 || value > 3) { ctx->cx.last_error = "forward_algo must be 0 (auto), 1 (bvh), 2 (per-point grid) or 3 (shared grid)"; return NLOS_ERR_INVALID; } ctx->cx.forward_algo = (int)value; }
  else if (k == "count_work") ctx->cx.count_work = value != 0;
  else if (k == "grid_cap") { if (value < 0 || value > 0x7fffffff) { ctx->cx.last_error = "grid_cap out of range"; return NLOS_ERR_INVALID; } ctx->cx.grid_cap = (int)value; }
  else if (k == "group_side") { if (value < 0 || value > 64) { ctx->cx.last_error = "group_side out of range"; return NLOS_ERR_INVALID; } ctx->cx.group_side = (int)value; }
  else if (k == "grid_slices") { if (value < 0 || value > 64) { ctx->cx.last_error = "grid_slices out of range"; return NLOS_ERR_INVALID; } ctx->cx.grid_slices = (int)value; }
  else if (k == "grid_budget_mb") { if (value < 0 || value > (1 << 20)) { ctx->cx.last_error = "grid_budget_mb out of range"; return NLOS_ERR_INVALID; } ctx->cx.grid_budget_mb = (int)value; }
  else if (k == "grid_res") { if (value < 0 || value > 4096) { ctx->cx.last_error = "grid_res out of range"; return NLOS_ERR_INVALID; } ctx->cx.grid_res = (int)value; }
  else { ctx->cx.last_error = "unknown option " + k; return NLOS_ERR_INVALID; }
  return NLOS_OK;
}
int nlos_ctx_get_timing(nlos_ctx* ctx, float* ms5) {
  if (!ctx || !ms5) return NLOS_ERR_INVALID;
  const Timing& t = ctx->cx.timing;
  ms5[0] = t.build_ms; ms5[1] = t.forward_ms; ms5[2] = t.residual_ms; ms5[3] = t.gradient_ms; ms5[4] = t.total_ms;
  return NLOS_OK;
}
uint64_t nlos_ctx_launch_count(nlos_ctx* ctx) { return ctx ? ctx->cx.launches : 0; }
int nlos_ctx_get_work_counters(nlos_ctx* ctx, uint64_t* out8) {
  if (!ctx || !out8) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_CUDA_OK(cudaStreamSynchronize(cx.stream));
    for (int i = 0; i < 8; ++i) out8[i] = 0;
    if (cx.buf("work_counters").p) NLOS_CUDA_OK(cudaMemcpy(out8, cx.buf("work_counters").p, 6 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    out8[6] = (uint64_t)cx.last_grid_res; out8[7] = (uint64_t)cx.last_forward_algo;
    return NLOS_OK;
  } catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}

int nlos_streamed_render_transient(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD, const float* verticesD, int numVertices,
                                   const float* vertexNormal, const float* vertexAlbedo, const int* trianglesD, int numTriangles, int numSamples,
                                   float lb, float ub, float res, double* transient, double* pathlengths, int refine_scale, int sigma_bin, int numBins) {
  Job j = base_job(originD, numSources, normalD, verticesD, numVertices, vertexNormal, vertexAlbedo, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, refine_scale, sigma_bin);
  j.transient = transient; j.pathlengths = pathlengths; j.kind = -1;
  return guarded(ctx, j);
}

int nlos_ggx_streamed_render_transient(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD, const float* verticesD, int numVertices,
                                       const float* vertexNormal, const float* vertexAlbedo, const int* trianglesD, int numTriangles, float alpha, int numSamples,
                                       float lb, float ub, float res, double* transient, double* pathlengths, int refine_scale, int sigma_bin, int numBins) {
  Job j = base_job(originD, numSources, normalD, verticesD, numVertices, vertexNormal, vertexAlbedo, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, refine_scale, sigma_bin);
  j.transient = transient; j.pathlengths = pathlengths; j.kind = -1; j.ggx = true; j.alpha = alpha;
  return guarded(ctx, j);
}

int nlos_streamed_render_intensity(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD, const float* verticesD, int numVertices,
                                   const float* vertexNormal, const int* trianglesD, int numTriangles, int numSamples, float lb, float ub, double* intensity) {
  Job j = base_job(originD, numSources, normalD, verticesD, numVertices, vertexNormal, nullptr, trianglesD, numTriangles, numSamples, lb, ub, 1.0f, 1, 1, 1);
  j.intensity = intensity; j.kind = 3;
  return guarded(ctx, j);
}

int nlos_ggx_streamed_render_intensity(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD, const float* verticesD, int numVertices,
                                       const float* vertexNormal, const int* trianglesD, int numTriangles, float alpha, int numSamples, float lb, float ub, double* intensity) {
  Job j = base_job(originD, numSources, normalD, verticesD, numVertices, vertexNormal, nullptr, trianglesD, numTriangles, numSamples, lb, ub, 1.0f, 1, 1, 1);
  j.intensity = intensity; j.kind = 3; j.ggx = true; j.alpha = alpha;
  return guarded(ctx, j);
}

int nlos_streamed_render_gradient(nlos_ctx* ctx, const double* data, const double* weight, const float* originD, int measurement, const float* normalD,
                                  const float* verticesD, int numVertices, const float* vertexNormal, const int* trianglesD, int numTriangles, int numSamples,
                                  float lb, float ub, float res, double* transient, double* pathlengths, double* gradient, int refine_scale, int sigma_bin,
                                  int testing_flag, int loss_test, int numBins) {
  Job j = base_job(originD, measurement, normalD, verticesD, numVertices, vertexNormal, nullptr, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, refine_scale, sigma_bin);
  j.data = data; j.weight = weight; j.transient = transient; j.pathlengths = pathlengths; j.gradient = gradient;
  j.testing_flag = testing_flag; j.loss_flag = loss_test; j.kind = 0;
  return guarded(ctx, j);
}

int nlos_streamed_render_gradient_w_albedo(nlos_ctx* ctx, const double* data, const double* weight, const float* originD, int measurement, const float* normalD,
                                           const float* verticesD, int numVertices, const float* albedo, const int* trianglesD, int numTriangles, int numSamples,
                                           float lb, float ub, float res, double* transient, double* pathlengths, double* gradient, int refine_scale, int sigma_bin,
                                           int testing_flag, int loss_test, int numBins) {
  // SSG.cpp:346-393: vertexNormal = nullptr, vertexAlbedo = albedo
  Job j = base_job(originD, measurement, normalD, verticesD, numVertices, nullptr, albedo, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, refine_scale, sigma_bin);
  j.data = data; j.weight = weight; j.transient = transient; j.pathlengths = pathlengths; j.gradient = gradient;
  j.testing_flag = testing_flag; j.loss_flag = loss_test; j.kind = 0;
  return guarded(ctx, j);
}

int nlos_streamed_render_gradient_albedo(nlos_ctx* ctx, const double* data, const double* weight, const float* originD, int measurement, const float* normalD,
                                         const float* verticesD, int numVertices, const float* albedo, const int* trianglesD, int numTriangles, int numSamples,
                                         float lb, float ub, float res, double* transient, double* pathlengths, int refine_scale, int sigma_bin,
                                         int testing_flag, int loss_test, int numBins, double* result) {
  Job j = base_job(originD, measurement, normalD, verticesD, numVertices, nullptr, albedo, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, refine_scale, sigma_bin);
  j.data = data; j.weight = weight; j.transient = transient; j.pathlengths = pathlengths; j.scalar_out = result;
  j.testing_flag = testing_flag; j.loss_flag = loss_test; j.kind = 1;
  return guarded(ctx, j);
}

int nlos_ggx_streamed_render_gradient(nlos_ctx* ctx, const double* data, const double* weight, const float* originD, int measurement, const float* normalD,
                                      const float* verticesD, int numVertices, const float* vertexNormal, const int* trianglesD, int numTriangles, float alpha,
                                      int numSamples, float lb, float ub, float res, double* transient, double* pathlengths, double* gradient, int refine_scale,
                                      int sigma_bin, int testing_flag, int numBins) {
  Job j = base_job(originD, measurement, normalD, verticesD, numVertices, vertexNormal, nullptr, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, refine_scale, sigma_bin);
  j.data = data; j.weight = weight; j.transient = transient; j.pathlengths = pathlengths; j.gradient = gradient;
  j.testing_flag = testing_flag; j.loss_flag = 0; j.kind = 0; j.ggx = true; j.alpha = alpha;
  return guarded(ctx, j);
}

int nlos_ggx_streamed_render_gradient_alpha(nlos_ctx* ctx, const double* data, const double* weight, const float* originD, int measurement, const float* normalD,
                                            const float* verticesD, int numVertices, const float* vertexNormal, const int* trianglesD, int numTriangles, float alpha,
                                            int numSamples, float lb, float ub, float res, double* transient, double* pathlengths, int refine_scale, int sigma_bin,
                                            int numBins, double* result) {
  Job j = base_job(originD, measurement, normalD, verticesD, numVertices, vertexNormal, nullptr, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, refine_scale, sigma_bin);
  j.data = data; j.weight = weight; j.transient = transient; j.pathlengths = pathlengths; j.scalar_out = result;
  j.kind = 2; j.ggx = true; j.alpha = alpha;
  return guarded(ctx, j);
}


int nlos_streamed_render_vertex_gradient(nlos_ctx* ctx, int vertex_num, const float* originD, int measurement, const float* normalD, const float* verticesD,
                                         int numVertices, const int* trianglesD, int numTriangles, int numSamples, float lb, float ub, float res,
                                         double* gradient, int refine_scale, int sigma_bin, int numBins) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_REQUIRE(originD && normalD && verticesD && trianglesD && gradient, "null argument");
    NLOS_REQUIRE(measurement > 0 && numVertices > 0 && numTriangles > 0 && numBins > 0 && res > 0.f && refine_scale >= 1 && sigma_bin >= 1, "bad size");
    cudaStream_t st = cx.stream;
    const int64_t L = measurement; const int F = numTriangles;
    const float* d_origin = stage_in(cx, "in_origin", originD, 3 * (size_t)L, st);
    const float* d_onormal = stage_in(cx, "in_onormal", normalD, 3 * (size_t)L, st);
    const float* d_verts = stage_in(cx, "in_verts", verticesD, 3 * (size_t)numVertices, st);
    const int* d_faces = stage_in(cx, "in_faces", trianglesD, 3 * (size_t)F, st);
    OutView<double> o_G = stage_out(cx, "out_vgrad", gradient, 3 * (size_t)numBins, true);
    DeviceScene sc; build_scene(cx, d_verts, numVertices, d_faces, F, d_origin, L, nullptr, nullptr, sc);
    float4* origin4 = cx.buf("origin4").as<float4>((size_t)L); float4* onormal4 = cx.buf("onormal4").as<float4>((size_t)L);
    launch_pack4(cx, d_origin, origin4, (size_t)L); launch_pack4(cx, d_onormal, onormal4, (size_t)L);
    RenderParams P; std::memset(&P, 0, sizeof P);
    P.origin = origin4; P.onormal = onormal4; P.L = L; P.src_offset = cx.src_offset; P.seed = cx.seed;
    P.spp = std::max(1, 1 + (numSamples - 1) / F);
    P.lb = lb; P.ub = ub; P.res = res; P.numBins = numBins; P.r_grad = refine_scale; P.s_bin = sigma_bin; P.K = 4 * refine_scale * sigma_bin + 1;
    TapTables taps = make_taps(res, refine_scale, sigma_bin);
    double* d_taps = cx.buf("taps").as<double>(taps.w.size());
    NLOS_CUDA_OK(cudaMemcpyAsync(d_taps, taps.w.data(), taps.w.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    double* acc = cx.buf("vgrad_acc").as<double>(3 * (size_t)numBins);
    NLOS_CUDA_OK(cudaMemsetAsync(acc, 0, 3 * (size_t)numBins * sizeof(double), st));
    launch_vertex_gradient(cx, sc, P, vertex_num, d_taps, taps.sigma2, acc);
    launch_finalize_gradient(cx, acc, o_G.dev, 3 * (size_t)numBins, 1.0 / (double)L);     // TG.cpp:431-435
    if (finish_out(cx, o_G)) NLOS_CUDA_OK(cudaStreamSynchronize(st));
    cx.last_error.clear();
    return NLOS_OK;
  } catch (const InvalidArg& e) { cx.last_error = e.what(); return NLOS_ERR_INVALID; }
  catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}

static int run_regulariser(nlos_ctx* ctx, int mode, const float* verticesD, int numVertices, const int* trianglesD, int numTriangles,
                           const int* face_affinity, double* grad, double* value_out) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_REQUIRE(verticesD && trianglesD && grad && (mode == 1 || face_affinity), "null argument");
    NLOS_REQUIRE(numVertices > 0 && numTriangles >= 0, "bad size");
    cudaStream_t st = cx.stream;
    const float* d_verts = stage_in(cx, "in_verts", verticesD, 3 * (size_t)numVertices, st);
    const int* d_faces = stage_in(cx, "in_faces", trianglesD, 3 * (size_t)numTriangles, st);
    const int* d_aff = mode == 0 ? stage_in(cx, "in_aff", face_affinity, 3 * (size_t)numTriangles, st) : nullptr;
    OutView<double> o_G = stage_out(cx, "out_reg", grad, 3 * (size_t)numVertices, false);
    double* d_val = cx.buf("reg_value").as<double>(1);
    launch_regulariser(cx, mode, d_verts, numVertices, d_faces, numTriangles, d_aff, o_G.dev, d_val);
    finish_out(cx, o_G);
    double h = 0;
    const bool value_on_device = value_out && is_device_ptr(value_out);        // device-resident loop: no host round trip for the value
    if (value_on_device) NLOS_CUDA_OK(cudaMemcpyAsync(value_out, d_val, sizeof(double), cudaMemcpyDeviceToDevice, st));
    else if (value_out) NLOS_CUDA_OK(cudaMemcpyAsync(&h, d_val, sizeof(double), cudaMemcpyDeviceToHost, st));
    if ((value_out && !value_on_device) || o_G.host) NLOS_CUDA_OK(cudaStreamSynchronize(st));
    if (value_out && !value_on_device) *value_out = h;
    cx.last_error.clear();
    return NLOS_OK;
  } catch (const InvalidArg& e) { cx.last_error = e.what(); return NLOS_ERR_INVALID; }
  catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}

int nlos_streamed_render_normal_smoothing(nlos_ctx* ctx, const float* verticesD, int numVertices, const int* trianglesD, int numTriangles,
                                          const int* face_affinity, double* curvature_grad, double* value_out) {
  return run_regulariser(ctx, 0, verticesD, numVertices, trianglesD, numTriangles, face_affinity, curvature_grad, value_out);
}

int nlos_streamed_render_curvature_grad(nlos_ctx* ctx, const float* verticesD, int numVertices, const int* trianglesD, int numTriangles, double* curvature_grad) {
  return run_regulariser(ctx, 1, verticesD, numVertices, trianglesD, numTriangles, nullptr, curvature_grad, nullptr);
}

int nlos_jitter_streamed_render_transient(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD, const float* verticesD, int numVertices,
                                          const float* vertexNormal, const float* vertexAlbedo, const int* trianglesD, int numTriangles, int numSamples,
                                          float lb, float ub, float res, const double* weight, int weight_offset, int weight_length, double* transient,
                                          double* pathlengths, int numBins) {
  if (!weight || weight_length <= 0) return NLOS_ERR_INVALID;
  Job j = base_job(originD, numSources, normalD, verticesD, numVertices, vertexNormal, vertexAlbedo, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, 1, 1);
  j.transient = transient; j.pathlengths = pathlengths; j.kind = -1; j.jw = weight; j.jg = weight; j.joff = weight_offset; j.jlen = weight_length;
  return guarded(ctx, j);
}

int nlos_jitter_streamed_render_gradient(nlos_ctx* ctx, const double* data, const double* weight, const float* originD, int measurement, const float* normalD,
                                         const float* verticesD, int numVertices, const float* vertexNormal, const int* trianglesD, int numTriangles,
                                         int numSamples, float lb, float ub, float res, const double* jitter_weight, const double* jitter_grad,
                                         int weight_offset, int weight_length, double* transient, double* pathlengths, double* gradient, int testing_flag,
                                         int numBins) {
  if (!jitter_weight || !jitter_grad || weight_length <= 0) return NLOS_ERR_INVALID;
  Job j = base_job(originD, measurement, normalD, verticesD, numVertices, vertexNormal, nullptr, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, 1, 1);
  j.data = data; j.weight = weight; j.transient = transient; j.pathlengths = pathlengths; j.gradient = gradient; j.testing_flag = testing_flag;
  j.kind = 0; j.jw = jitter_weight; j.jg = jitter_grad; j.joff = weight_offset; j.jlen = weight_length;
  return guarded(ctx, j);
}

static int run_ray_query(nlos_ctx* ctx, int mode, const float* originsD, const float* directionsD, int num_ray, const float* verticesD, int num_vertices,
                         const int* trianglesD, int num_triangles, float* intersect) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_REQUIRE(num_ray >= 0 && num_vertices > 0 && num_triangles > 0, "bad size");
    if (num_ray == 0) return NLOS_OK;
    NLOS_REQUIRE(originsD && directionsD && verticesD && trianglesD && intersect, "null argument");
    cudaStream_t st = cx.stream;
    const int64_t N = num_ray;
    const float* d_o = stage_in(cx, "in_origin", originsD, 3 * (size_t)N, st);
    const float* d_d = stage_in(cx, "in_dirs", directionsD, 3 * (size_t)N, st);
    const float* d_verts = stage_in(cx, "in_verts", verticesD, 3 * (size_t)num_vertices, st);
    const int* d_faces = stage_in(cx, "in_faces", trianglesD, 3 * (size_t)num_triangles, st);
    OutView<float> o_out = stage_out(cx, "out_isect", intersect, (mode == 1 ? 1 : 3) * (size_t)N, mode == 0);   // misses leave (u,v) untouched
    DeviceScene sc; build_scene(cx, d_verts, num_vertices, d_faces, num_triangles, d_o, N, nullptr, nullptr, sc);
    launch_ray_query(cx, sc, mode, d_o, d_d, N, o_out.dev);
    if (finish_out(cx, o_out)) NLOS_CUDA_OK(cudaStreamSynchronize(st));
    cx.last_error.clear();
    return NLOS_OK;
  } catch (const InvalidArg& e) { cx.last_error = e.what(); return NLOS_ERR_INVALID; }
  catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}

// ---- first-generation API (stratified_transient_raytracer/)
int nlos_sr_streamed_render_transient(nlos_ctx* ctx, const float* originD, int numSources, const float* normalD, const float* verticesD, int numVertices,
                                      const float* vertexNormal, const float* vertexAlbedo, const int* trianglesD, int numTriangles, int numSamples,
                                      float lb, float ub, float res, double* transient, double* pathlengths, int numBins) {
  Job j = base_job(originD, numSources, normalD, verticesD, numVertices, vertexNormal, vertexAlbedo, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, 1, 1);
  j.transient = transient; j.pathlengths = pathlengths; j.kind = -1; j.sr = 1;
  return guarded(ctx, j);
}
int nlos_sr_render_transient(nlos_ctx* ctx, const float* originD, const float* normalD, const float* verticesD, int numVertices, const int* trianglesD,
                             int numTriangles, int numSamples, float lb, float ub, float res, double* transient, double* pathlengths, int numBins) {
  return nlos_sr_streamed_render_transient(ctx, originD, 1, normalD, verticesD, numVertices, nullptr, nullptr, trianglesD, numTriangles, numSamples, lb, ub, res,
                                           transient, pathlengths, numBins);
}
int nlos_sr_streamed_render_gradient(nlos_ctx* ctx, const double* data, const float* originD, int measurement, const float* normalD, const float* verticesD,
                                     int numVertices, const int* trianglesD, int numTriangles, int numSamples, float lb, float ub, float res, int w_width,
                                     double* transient, double* pathlengths, double* gradient, int numBins) {
  Job j = base_job(originD, measurement, normalD, verticesD, numVertices, nullptr, nullptr, trianglesD, numTriangles, numSamples, lb, ub, res, numBins, 1, 1);
  j.data = data; j.weight = nullptr; j.transient = transient; j.pathlengths = pathlengths; j.gradient = gradient; j.kind = 0; j.sr = 1; j.w_width = w_width;
  return guarded(ctx, j);
}

int nlos_embree3_tbb_line_intersection(nlos_ctx* ctx, const float* originsD, const float* directionsD, int num_ray, const float* verticesD, int num_vertices,
                                       const int* trianglesD, int num_triangles, float* intersect) {
  return run_ray_query(ctx, 0, originsD, directionsD, num_ray, verticesD, num_vertices, trianglesD, num_triangles, intersect);
}
int nlos_embree3_tbb_short_line_intersection(nlos_ctx* ctx, const float* originsD, const float* directionsD, int num_ray, const float* verticesD,
                                             int num_vertices, const int* trianglesD, int num_triangles, float* intersect) {
  return run_ray_query(ctx, 1, originsD, directionsD, num_ray, verticesD, num_vertices, trianglesD, num_triangles, intersect);
}
int nlos_barycentric_to_world(nlos_ctx* ctx, const float* verticesD, int num_vertices, const int* trianglesD, int num_triangles, const float* barycoord,
                              int num_ray, float* intersection_p) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_REQUIRE(num_ray >= 0 && num_vertices > 0 && num_triangles > 0, "bad size");
    if (num_ray == 0) return NLOS_OK;
    NLOS_REQUIRE(verticesD && trianglesD && barycoord && intersection_p, "null argument");
    cudaStream_t st = cx.stream;
    const float* d_verts = stage_in(cx, "in_verts", verticesD, 3 * (size_t)num_vertices, st);
    const int* d_faces = stage_in(cx, "in_faces", trianglesD, 3 * (size_t)num_triangles, st);
    const float* d_b = stage_in(cx, "in_bary", barycoord, 3 * (size_t)num_ray, st);
    OutView<float> o_out = stage_out(cx, "out_world", intersection_p, 3 * (size_t)num_ray, true);   // misses stay untouched
    launch_bary_to_world(cx, d_verts, d_faces, d_b, num_ray, o_out.dev);
    if (finish_out(cx, o_out)) NLOS_CUDA_OK(cudaStreamSynchronize(st));
    cx.last_error.clear();
    return NLOS_OK;
  } catch (const InvalidArg& e) { cx.last_error = e.what(); return NLOS_ERR_INVALID; }
  catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}

int nlos_debug_copy_visibility_words(nlos_ctx* ctx, uint32_t* out, int64_t n, int64_t* n_available) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_CUDA_OK(cudaStreamSynchronize(cx.stream));
    if (n_available) *n_available = (int64_t)cx.vis_words;
    if (out && n > 0) {
      NLOS_REQUIRE((size_t)n <= cx.vis_words, "more words requested than the last gradient call produced");
      NLOS_CUDA_OK(cudaMemcpy(out, cx.buf("vis").p, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    cx.last_error.clear();
    return NLOS_OK;
  } catch (const InvalidArg& e) { cx.last_error = e.what(); return NLOS_ERR_INVALID; }
  catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}

int nlos_debug_visibility(nlos_ctx* ctx, const float* originD, int numSources, const float* verticesD, int numVertices, const int* trianglesD, int numTriangles,
                          int numSamples, uint8_t* visibility, uint64_t* counters3) {
  if (!ctx) return NLOS_ERR_INVALID;
  Ctx& cx = ctx->cx;
  try {
    NLOS_CUDA_OK(cudaSetDevice(cx.device));
    NLOS_REQUIRE(originD && verticesD && trianglesD && visibility, "null argument");
    NLOS_REQUIRE(numSources > 0 && numVertices > 0 && numTriangles > 0, "sizes must be positive");
    cudaStream_t st = cx.stream;
    const int64_t L = numSources; const int F = numTriangles;
    const float* d_origin = stage_in(cx, "in_origin", originD, 3 * (size_t)L, st);
    const float* d_verts = stage_in(cx, "in_verts", verticesD, 3 * (size_t)numVertices, st);
    const int* d_faces = stage_in(cx, "in_faces", trianglesD, 3 * (size_t)F, st);
    DeviceScene sc; build_scene(cx, d_verts, numVertices, d_faces, F, d_origin, L, nullptr, nullptr, sc);
    float4* origin4 = cx.buf("origin4").as<float4>((size_t)L);
    launch_pack4(cx, d_origin, origin4, (size_t)L);
    RenderParams P; std::memset(&P, 0, sizeof P);
    P.origin = origin4; P.onormal = origin4; P.L = L; P.src_offset = cx.src_offset; P.seed = cx.seed;
    P.spp = std::max(1, 1 + (numSamples - 1) / F);
    P.chunk = auto_chunk(cx, "chunk_forward", F, L, 128);
    const size_t nvis = (size_t)L * F * P.spp;
    OutView<uint8_t> o_vis = stage_out(cx, "out_vis8", visibility, nvis, false);
    unsigned long long* cnt = cx.buf("vis_counters").as<unsigned long long>(3);
    NLOS_CUDA_OK(cudaMemsetAsync(cnt, 0, 3 * sizeof(unsigned long long), st));
    launch_visibility(cx, sc, P, o_vis.dev, cnt);
    finish_out(cx, o_vis);
    unsigned long long h[3] = {0, 0, 0};
    NLOS_CUDA_OK(cudaMemcpyAsync(h, cnt, sizeof h, cudaMemcpyDeviceToHost, st));
    NLOS_CUDA_OK(cudaStreamSynchronize(st));
    if (counters3) { counters3[0] = h[0]; counters3[1] = h[1]; counters3[2] = h[2]; }
    cx.last_error.clear();
    return NLOS_OK;
  } catch (const InvalidArg& e) { cx.last_error = e.what(); return NLOS_ERR_INVALID; }
  catch (const std::exception& e) { cx.last_error = e.what(); cudaGetLastError(); return NLOS_ERR_CUDA; }
}

}  // extern "C"
