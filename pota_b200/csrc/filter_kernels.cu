// filter_kernels.cu — K2b launcher (table-driven evaluator), closest-filter gather, K3 resolve.
#include "filter_kernels.cuh"
#include "unrolled_dispatch.h"

namespace lb {

constexpr int kSplatBlock = 128;

__global__ void __launch_bounds__(kSplatBlock)
k_filter_splat_table(const __grid_constant__ LensTable lens, const __grid_constant__ CamConsts<float> cam,
                     const __grid_constant__ FilterConsts fc, const __grid_constant__ AovSet aovs, const __grid_constant__ SampleIO s,
                     const WorkItem *__restrict__ work, FilterCounters *__restrict__ counters, uint64_t sample_base) {
  const TableEval<float> ev(lens);
  splat_persistent(ev, cam, fc, aovs, s, work, counters, sample_base);
}

// Closest-filter AOVs: fetch the value of the sample that won the depth key, if it belongs to this batch.
__global__ void k_closest_gather(const __grid_constant__ FilterConsts fc, const __grid_constant__ AovSet aovs,
                                 const __grid_constant__ SampleIO s, uint64_t sample_base) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (size_t)fc.xres * fc.yres) return;
  const unsigned long long key = aovs.zkey[p];
  if (key != ~0ull) {
    const uint32_t idx = (0xFFFFFFFFu - (uint32_t)key) - (uint32_t)sample_base;
    if ((size_t)idx < s.n)
      for (int a = 0; a < fc.n_aov; ++a)
        if (aovs.filter[a] == 1 && aovs.role[a] != 2) aovs.buffer[a][p] = aov_value(aovs, s, a, idx, 0.f);
  }
  const unsigned long long dkey = aovs.zkey_debug[p];
  if (dkey != ~0ull && aovs.debug_samples) {
    const uint32_t idx = (0xFFFFFFFFu - (uint32_t)dkey) - (uint32_t)sample_base;
    if ((size_t)idx < s.n)
      for (int a = 0; a < fc.n_aov; ++a)
        if (aovs.filter[a] == 1 && aovs.role[a] == 2) {
          const float v = (float)aovs.debug_samples[idx];
          aovs.buffer[a][p] = make_float4(v, v, v, v);
        }
  }
}

// driver_process_bucket, lentil_imager.cpp:112-189 (non-crypto branch): pure bandwidth, 20 B in / 16 B out per pixel
__global__ void k_resolve(const float4 *__restrict__ buffer, const float *__restrict__ weight, int filter, int role, int xres, int x0,
                          int y0, int w, int h, float4 *__restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= w || j >= h) return;
  const size_t lp = (size_t)(x0 + i) + (size_t)(y0 + j) * xres;
  float4 v = __ldg(buffer + lp);
  if (filter == 0) {
    if (role != 2) {
      const float fw = __ldg(weight + lp);
      if (fw != 0.0f) { const float inv = 1.0f / fw; v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv; }
    }
  } else {
    v.w = 1.0f;
  }
  out[(size_t)j * w + i] = v;
}

cudaError_t launch_filter_splat(int lens_kernel, const LensTable &lens, const CamConsts<float> &cam, const FilterConsts &fc,
                                const AovSet &aovs, const SampleIO &s, const WorkItem *work, FilterCounters *counters,
                                uint64_t sample_base, int num_sms, cudaStream_t stream) {
  const int grid = num_sms * 8;
  if (lens_kernel >= 0) {
    if (auto fn = unrolled_bw_launcher(lens_kernel)) return fn(cam, fc, aovs, s, work, counters, sample_base, grid, stream);
  }
  k_filter_splat_table<<<grid, kSplatBlock, 0, stream>>>(lens, cam, fc, aovs, s, work, counters, sample_base);
  return cudaGetLastError();
}

cudaError_t launch_closest_gather(const FilterConsts &fc, const AovSet &aovs, const SampleIO &s, uint64_t sample_base,
                                  cudaStream_t stream) {
  const size_t npx = (size_t)fc.xres * fc.yres;
  k_closest_gather<<<(unsigned)((npx + 255) / 256), 256, 0, stream>>>(fc, aovs, s, sample_base);
  return cudaGetLastError();
}

cudaError_t launch_resolve(const float4 *buffer, const float *weight, int filter, int role, int xres, int x0, int y0, int w, int h,
                           float4 *out, cudaStream_t stream) {
  if (w <= 0 || h <= 0) return cudaSuccess;
  dim3 grid((w + 255) / 256, h);
  k_resolve<<<grid, 256, 0, stream>>>(buffer, weight, filter, role, xres, x0, y0, w, h, out);
  return cudaGetLastError();
}

}  // namespace lb
