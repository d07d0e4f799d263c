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

// ---- reduction throughput: red.global.add.v4.f32 into an L2-resident plane at pseudo-random pixels --------
namespace {
__global__ void __launch_bounds__(256) k_red_peak(float4 *buf, unsigned npx, int iters) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  const float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    atomicAdd(buf + (s >> 8) % npx, v);
  }
}
}  // namespace

// GB/s of 16-byte vector reductions (the splat accumulate of one gaussian AOV) into a `megabytes` MB plane.
extern "C" int lb_bench_red_peak(int device, int megabytes, double *gbytes_per_s_out) {
  if (!gbytes_per_s_out || megabytes <= 0) return LB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return LB_ERR_NO_DEVICE;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const unsigned npx = (unsigned)((size_t)megabytes * 1000000 / 16);
  float4 *d = nullptr;
  if (cudaMalloc(&d, (size_t)npx * 16) != cudaSuccess) { cudaSetDevice(prev); return LB_ERR_CUDA; }
  cudaMemset(d, 0, (size_t)npx * 16);
  const int grid = sms * 8, block = 256, iters = 256;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    k_red_peak<<<grid, block>>>(d, npx, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double gbs = 16.0 * grid * block * (double)iters / (ms * 1e-3) / 1e9;
    if (rep > 0 && gbs > best) best = gbs;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  cudaSetDevice(prev);
  *gbytes_per_s_out = best;
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}

// ---- known-answer hooks for the device primitives (tests only; the splat images pin them implicitly) --------------
#include "lens_device.cuh"
namespace {
__global__ void k_debug_primitives(size_t n, const uint32_t *__restrict__ v0, const uint32_t *__restrict__ v1, uint32_t *__restrict__ tea_out,
                                   uint32_t *__restrict__ lcg_state_out, float *__restrict__ lcg_float_out, float *__restrict__ trig_out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t t = lb::tea8(v0[i], v1[i]);
  tea_out[i] = t;
  uint32_t s = t;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    lcg_float_out[4 * i + k] = lb::lcg_rng(s);
    lcg_state_out[4 * i + k] = s;
  }
  const float x = __fmul_rn(__fsub_rn(lcg_float_out[4 * i], 0.5f), 20.0f);  // fast_sin / fast_cos argument in [-10, 10), no FMA contraction
  trig_out[2 * i] = lb::fast_sin(x);
  trig_out[2 * i + 1] = lb::fast_cos(x);
}
}  // namespace

// tea<8>(v0, v1), four LCG draws seeded with it (state + float), fast_sin / fast_cos of a value derived from the first
// draw -- computed on the device.  All pointers are HOST arrays: v0, v1 [n]; tea_out [n]; lcg_state_out, lcg_float_out
// [n][4]; trig_out [n][2].  (global.h:32-57, lens.h:17-37)
extern "C" int lb_debug_primitives(int device, size_t n, const uint32_t *v0, const uint32_t *v1, uint32_t *tea_out, uint32_t *lcg_state_out,
                                   float *lcg_float_out, float *trig_out) {
  if (!v0 || !v1 || !tea_out || !lcg_state_out || !lcg_float_out || !trig_out) return LB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return LB_ERR_NO_DEVICE;
  if (n == 0) return LB_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  uint32_t *d = nullptr;
  const size_t words = n * (2 + 1 + 4 + 4 + 2);
  if (cudaMalloc(&d, words * 4) != cudaSuccess) { cudaSetDevice(prev); return LB_ERR_CUDA; }
  uint32_t *dv0 = d, *dv1 = d + n, *dt = d + 2 * n, *ds = d + 3 * n;
  float *df = (float *)(d + 7 * n), *dg = (float *)(d + 11 * n);
  cudaMemcpy(dv0, v0, n * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dv1, v1, n * 4, cudaMemcpyHostToDevice);
  k_debug_primitives<<<(unsigned)((n + 255) / 256), 256>>>(n, dv0, dv1, dt, ds, df, dg);
  cudaMemcpy(tea_out, dt, n * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(lcg_state_out, ds, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(lcg_float_out, df, n * 16, cudaMemcpyDeviceToHost);
  cudaMemcpy(trig_out, dg, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  cudaSetDevice(prev);
  return cudaGetLastError() == cudaSuccess ? LB_OK : LB_ERR_CUDA;
}
