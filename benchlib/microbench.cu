// microbench.cu — measured denominators of the FP32 / atomic roofline.  MEASUREMENT ONLY: built into its own
// library (benchlib/libnlos_microbench.so, loaded by bench.py), not part of the product ABI.  MEASURED_PEAKS.json only records HBM and
// bf16 peaks, which do not bound this path, so the FP32 FFMA rate, the FP64 L2 RED.ADD rate and the shared-memory atomic rate are
// measured here, on the device and at the clocks the bench itself runs at.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC microbench.cu -o libnlos_microbench.so
#include <cuda_runtime.h>
#include <cstdint>

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

// shared-memory atomicAdd(unsigned) to `naddr` (power of two) hashed addresses per block — the cell counters of the perspective grid
__global__ void __launch_bounds__(256) k_smem_atomic(unsigned* out, unsigned naddr_mask, int iters) {
  extern __shared__ unsigned cells[];
  for (unsigned i = threadIdx.x; i <= naddr_mask; i += blockDim.x) cells[i] = 0u;
  __syncthreads();
  unsigned h = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u;
  for (int i = 0; i < iters; ++i) { h = h * 1664525u + 1013904223u; atomicAdd(&cells[(h >> 8) & naddr_mask], 1u); }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = cells[0];
}

template <class F> float best_ms(cudaStream_t st, F launch) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  launch(); cudaStreamSynchronize(st);                       // warm-up
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(a, st); launch(); cudaEventRecord(b, st); cudaEventSynchronize(b);
    float ms = 0; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best;
  }
  cudaEventDestroy(a); cudaEventDestroy(b);
  return best;
}

}  // namespace

extern "C" {

// FP32 FFMA throughput in TFLOP/s (< 0 on error)
double nlos_microbench_fp32(int device) {
  if (cudaSetDevice(device) != cudaSuccess) return -1.0;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8, iters = 1 << 15;
  float* out = nullptr; if (cudaMalloc(&out, (size_t)blocks * 256 * sizeof(float)) != cudaSuccess) return -1.0;
  cudaStream_t st; cudaStreamCreate(&st);
  const float ms = best_ms(st, [&] { k_ffma<<<blocks, 256, 0, st>>>(out, iters, 0.999f, 0.001f); });
  cudaStreamDestroy(st); cudaFree(out);
  return 2.0 * 8.0 * (double)iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e12;
}

// FP64 RED.ADD throughput in 1e9 atomics/s over num_addresses (rounded down to a power of two) hashed addresses (< 0 on error)
double nlos_microbench_red_f64(int device, int64_t num_addresses) {
  if (num_addresses < 1 || cudaSetDevice(device) != cudaSuccess) return -1.0;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  unsigned n = 1; while ((int64_t)n * 2 <= num_addresses) n *= 2;
  const int blocks = sms * 8, iters = 256;
  double* buf = nullptr; if (cudaMalloc(&buf, (size_t)n * sizeof(double)) != cudaSuccess) return -1.0;
  cudaStream_t st; cudaStreamCreate(&st);
  cudaMemsetAsync(buf, 0, (size_t)n * sizeof(double), st);
  const float ms = best_ms(st, [&] { k_red_f64<<<blocks, 256, 0, st>>>(buf, n - 1, iters); });
  cudaStreamDestroy(st); cudaFree(buf);
  return (double)iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e9;
}

// shared-memory atomicAdd(unsigned) throughput in 1e9 atomics/s, num_addresses (power of two, <= 32768) hashed counters per block (< 0 on error)
double nlos_microbench_smem_atomic(int device, int num_addresses) {
  if (num_addresses < 1 || num_addresses > 32768 || cudaSetDevice(device) != cudaSuccess) return -1.0;
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  unsigned n = 1; while ((int)n * 2 <= num_addresses) n *= 2;
  const int blocks = sms * 4, iters = 4096;
  unsigned* out = nullptr; if (cudaMalloc(&out, (size_t)blocks * sizeof(unsigned)) != cudaSuccess) return -1.0;
  cudaFuncSetAttribute(k_smem_atomic, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(n * sizeof(unsigned)));
  cudaStream_t st; cudaStreamCreate(&st);
  const float ms = best_ms(st, [&] { k_smem_atomic<<<blocks, 256, n * sizeof(unsigned), st>>>(out, n - 1, iters); });
  cudaStreamDestroy(st); cudaFree(out);
  return (double)iters * (double)blocks * 256.0 / (ms * 1e-3) / 1e9;
}

}  // extern "C"
