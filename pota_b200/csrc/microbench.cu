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

// ---- splat accumulate, three ways: what "shared-memory tile accumulation" (north_star) buys on this part ----------------
// One splat = RGBA (16 B) + filter weight (4 B) added at a pseudo-random pixel.
//   mode 0: red.global.add.v4.f32 + red.global.add.f32 into planes of `megabytes` MB (what the splat kernels do)
//   mode 1: shared-memory tile of kTile x kTile pixels, float atomics (ATOMS.CAST.SPIN loops: sm_100a has no native
//           shared-memory float add), tile flushed once with red.global.add.v4.f32 -- 28 splats per tile pixel
//   mode 2: the same tile with native 32-bit integer atomics (ATOMS.ADD), i.e. the cost floor of any smem-atomic tile
namespace {
constexpr int kTile = 96;  // 96 x 96 x 20 B = 184 KB of the 227 KB a CTA can have
__global__ void __launch_bounds__(256) k_accum_global(float4 *buf, float *wgt, unsigned npx, int iters) {
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  const float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    const unsigned px = (s >> 8) % npx;
    atomicAdd(buf + px, v);
    atomicAdd(wgt + px, 1.0f);
  }
}
template <bool kInt>
__global__ void __launch_bounds__(1024) k_accum_tile(float4 *buf, float *wgt, unsigned npx, int iters) {
  extern __shared__ float4 tile4[];
  float *tilew = reinterpret_cast<float *>(tile4 + kTile * kTile);
  for (int i = threadIdx.x; i < kTile * kTile; i += blockDim.x) { tile4[i] = make_float4(0.f, 0.f, 0.f, 0.f); tilew[i] = 0.f; }
  __syncthreads();
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  for (int i = 0; i < iters; ++i) {
    s = s * 1664525u + 1013904223u;
    const unsigned px = (s >> 8) % (unsigned)(kTile * kTile);
    if (kInt) {
      unsigned *q = reinterpret_cast<unsigned *>(tile4 + px);
      atomicAdd(q, 1u); atomicAdd(q + 1, 1u); atomicAdd(q + 2, 1u); atomicAdd(q + 3, 1u);
      atomicAdd(reinterpret_cast<unsigned *>(tilew + px), 1u);
    } else {
      float *q = reinterpret_cast<float *>(tile4 + px);
      atomicAdd(q, 1.f); atomicAdd(q + 1, 1.f); atomicAdd(q + 2, 1.f); atomicAdd(q + 3, 1.f);
      atomicAdd(tilew + px, 1.f);
    }
  }
  __syncthreads();
  const unsigned base = (blockIdx.x * 7919u * (unsigned)(kTile * kTile)) % (npx - kTile * kTile);
  for (int i = threadIdx.x; i < kTile * kTile; i += blockDim.x) {
    atomicAdd(buf + base + i, tile4[i]);
    atomicAdd(wgt + base + i, tilew[i]);
  }
}
}  // namespace

// Splats per second (1e9/s) of the accumulate alone; see the modes above.
extern "C" int lb_bench_splat_accum(int device, int mode, int megabytes, double *gsplats_per_s_out) {
  if (!gsplats_per_s_out || megabytes <= 0 || mode < 0 || mode > 2) return LB_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device >= ndev) return LB_ERR_NO_DEVICE;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const unsigned npx = (unsigned)((size_t)megabytes * 1000000 / 16);
  if (npx <= 2u * kTile * kTile) { cudaSetDevice(prev); return LB_ERR_INVALID; }
  float4 *d = nullptr;
  float *w = nullptr;
  if (cudaMalloc(&d, (size_t)npx * 16) != cudaSuccess || cudaMalloc(&w, (size_t)npx * 4) != cudaSuccess) {
    cudaFree(d);
    cudaSetDevice(prev);
    return LB_ERR_CUDA;
  }
  cudaMemset(d, 0, (size_t)npx * 16);
  cudaMemset(w, 0, (size_t)npx * 4);
  const size_t smem = (size_t)kTile * kTile * 20;
  cudaFuncSetAttribute(k_accum_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k_accum_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 256;
  const int grid = mode == 0 ? sms * 8 : sms, block = mode == 0 ? 256 : 1024;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) k_accum_global<<<grid, block>>>(d, w, npx, iters);
    else if (mode == 1) k_accum_tile<false><<<grid, block, smem>>>(d, w, npx, iters);
    else k_accum_tile<true><<<grid, block, smem>>>(d, w, npx, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double)grid * block * (double)iters / (ms * 1e-3) / 1e9;
    if (rep > 0 && ms > 0.f && g > best) best = g;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  cudaFree(w);
  cudaSetDevice(prev);
  *gsplats_per_s_out = best;
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
