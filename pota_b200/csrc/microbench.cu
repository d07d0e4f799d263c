// microbench.cu — on-box FP32 FMA peak, the roofline denominator of the polynomial kernels
// ("measure an FFMA micro-benchmark on the box and use that as denominator").
#include <cuda_runtime.h>

#include "../../include/lentil_b200.h"

namespace {
constexpr int kChains = 8;
constexpr int kIters = 4096;

// 8 independent dependent-FMA chains per thread, register operands only (the 3-register FFMA form
// the table-driven kernels execute)
__global__ void __launch_bounds__(256) k_ffma_peak(float *out, float a, float b) {
  float v[kChains];
#pragma unroll
  for (int k = 0; k < kChains; ++k) v[k] = (float)(threadIdx.x + k);
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int k = 0; k < kChains; ++k) v[k] = fmaf(v[k], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kChains; ++k) s += v[k];
  if (s == 123.456f) out[0] = s;  // never true: keeps the chains alive
}
}  // namespace

extern "C" int lb_bench_fp32_peak(int device, double *tflops_out) {
  if (!tflops_out) return LB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return LB_ERR_NO_DEVICE;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  float *d = nullptr;
  cudaMalloc(&d, 4);
  const int grid = sms * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    k_ffma_peak<<<grid, block>>>(d, 1.0000001f, 1e-7f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = 2.0 * kChains * (double)kIters * grid * block;
    if (rep > 0 && ms > 0.f) best = best > flop / (ms * 1e-3) / 1e12 ? best : flop / (ms * 1e-3) / 1e12;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  cudaSetDevice(prev);
  *tflops_out = best;
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}
