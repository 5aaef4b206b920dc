// microbench.cu — measured denominators of the FP32 / atomic roofline: MEASURED_PEAKS.json only
// records HBM and bf16 peaks, which do not bound this path.
#include "../../include/nlos_b200.h"
#include "nlos_ctx.h"


namespace {

// 8 independent FFMA chains per thread: 2 flops x 8 x iters x threads
__global__ void __launch_bounds__(256) k_ffma(float* out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
    x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// FP64 RED.ADD to `naddr` addresses (power of two), hashed per thread and iteration
__global__ void __launch_bounds__(256) k_red_f64(double* buf, unsigned naddr_mask, int iters) {
  unsigned h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
  for (int i = 0; i < iters; ++i) {
    h = h * 1664525u + 1013904223u;
    atomicAdd(buf + ((h >> 8) & naddr_mask), 1.0);
  }
}

float time_ms(nlos::Ctx& cx, void (*launch)(nlos::Ctx&, void*), void* arg) {
  launch(cx, arg);                                    // warm-up
  cudaStreamSynchronize(cx.stream);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(cx.ev[6], cx.stream); launch(cx, arg); cudaEventRecord(cx.ev[7], cx.stream);
    cudaEventSynchronize(cx.ev[7]);
    float ms = 0; cudaEventElapsedTime(&ms, cx.ev[6], cx.ev[7]); best = ms < best ? ms : best;
  }
  return best;
}

struct FfmaArgs { float* out; int blocks, iters; };
void launch_ffma(nlos::Ctx& cx, void* p) { auto* a = (FfmaArgs*)p; k_ffma<<<a->blocks, 256, 0, cx.stream>>>(a->out, a->iters, 0.999f, 0.001f); cx.launches++; }
struct RedArgs { double* buf; unsigned mask; int blocks, iters; };
void launch_red(nlos::Ctx& cx, void* p) { auto* a = (RedArgs*)p; k_red_f64<<<a->blocks, 256, 0, cx.stream>>>(a->buf, a->mask, a->iters); cx.launches++; }

}  // namespace

extern "C" {

double nlos_microbench_fp32(nlos_ctx* ctx) {
  if (!ctx) return -1.0;
  nlos::Ctx& cx = ctx->cx;
  try {
    cudaSetDevice(cx.device);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cx.device);
    FfmaArgs a; a.blocks = sms * 8; a.iters = 1 << 15;
    a.out = cx.buf("mb_out").as<float>((size_t)a.blocks * 256);
    const float ms = time_ms(cx, launch_ffma, &a);
    return 2.0 * 8.0 * (double)a.iters * (double)a.blocks * 256.0 / (ms * 1e-3) / 1e12;   // TFLOP/s
  } catch (const std::exception& e) { cx.last_error = e.what(); return -1.0; }
}

double nlos_microbench_red_f64(nlos_ctx* ctx, int64_t num_addresses) {
  if (!ctx || num_addresses < 1) return -1.0;
  nlos::Ctx& cx = ctx->cx;
  try {
    cudaSetDevice(cx.device);
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cx.device);
    unsigned n = 1; while ((int64_t)n * 2 <= num_addresses) n *= 2;
    RedArgs a; a.mask = n - 1; a.blocks = sms * 8; a.iters = 256;
    a.buf = cx.buf("mb_red").as<double>(n);
    cudaMemsetAsync(a.buf, 0, (size_t)n * sizeof(double), cx.stream);
    const float ms = time_ms(cx, launch_red, &a);
    return (double)a.iters * (double)a.blocks * 256.0 / (ms * 1e-3) / 1e9;                 // G atomics/s
  } catch (const std::exception& e) { cx.last_error = e.what(); return -1.0; }
}

}  // extern "C"
