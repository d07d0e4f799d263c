// filter_kernels.cu — K2b launcher (table-driven evaluator), closest-filter gather, K3 resolve.
#include "filter_kernels.cuh"
#include "unrolled_dispatch.h"

namespace lb {

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
  const unsigned long long key = aovs.zkey ? aovs.zkey[p] : ~0ull;
  if (key != ~0ull) {
    const uint32_t idx = (0xFFFFFFFFu - (uint32_t)key) - (uint32_t)sample_base;
    if ((size_t)idx < s.n)
      for (int a = 0; a < fc.n_aov; ++a)
        if (aovs.filter[a] == 1 && aovs.role[a] != 2) aovs.buffer[a][p] = aov_value(aovs, s, a, idx, 0.f);
  }
  const unsigned long long dkey = (aovs.zkey_debug && aovs.debug_samples) ? aovs.zkey_debug[p] : ~0ull;
  if (dkey != ~0ull) {
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

// Ranked cryptomatte resolve, lentil_imager.cpp:122-161.  One block per bucket row because of the reference's
// `break` (:132-134): the first pixel of the row with <= rank ids ends the row, it and everything after it keep
// what the caller's bucket held.  Order: weight descending (compareTail, :11-16); equal weights stay in ascending-id
// order, which is what std::sort's insertion sort leaves for the map-ordered input of up to 16 entries.
__global__ void __launch_bounds__(256)
k_resolve_crypto(const uint32_t *__restrict__ key, const float *__restrict__ wgt, const float4 *__restrict__ total, int slots, int rank,
                 int xres, int x0, int y0, int w, float4 *__restrict__ out, uint8_t *__restrict__ brk) {
  __shared__ int first_break;
  const int j = blockIdx.x;
  if (threadIdx.x == 0) first_break = w;
  __syncthreads();
  for (int i = threadIdx.x; i < w; i += blockDim.x) {
    const size_t p = (size_t)(y0 + j) * xres + (x0 + i);
    int size = 0;
    for (int k = 0; k < slots; ++k) size += key[p * slots + k] != kCryptoFree;
    if (brk) brk[p] = size <= rank;
    else if (size <= rank) atomicMin(&first_break, i);
  }
  __syncthreads();
  const int end = first_break;
  for (int i = threadIdx.x; i < end; i += blockDim.x) {
    const size_t p = (size_t)(y0 + j) * xres + (x0 + i);
    if (brk && brk[p]) continue;
    const uint32_t *kp = key + p * slots;
    const float *wp = wgt + p * slots;
    const float tw = total[p].x;  // crypto_total_weight
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = 0; e < slots; ++e) {
      if (kp[e] == kCryptoFree) continue;
      const float ide = __uint_as_float(kp[e]), we = wp[e];
      int before = 0;
      for (int f = 0; f < slots; ++f) {
        if (f == e || kp[f] == kCryptoFree) continue;
        const float idf = __uint_as_float(kp[f]), wf = wp[f];
        before += (wf > we) || (wf == we && idf < ide);
      }
      if (before == rank) { o.x = ide; o.y = we / tw; }
      else if (before == rank + 1) { o.z = ide; o.w = we / tw; }
    }
    out[(size_t)j * w + i] = o;
  }
}

__global__ void __launch_bounds__(256)
k_crypto_merge(uint32_t *__restrict__ key, float *__restrict__ wgt, const uint32_t *__restrict__ other_key, const float *__restrict__ other_wgt,
               size_t n, int slots, FilterCounters *counters) {
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t k = other_key[t];
  if (k == kCryptoFree) return;
  crypto_insert(key, wgt, slots, (unsigned)(t / (size_t)slots), __uint_as_float(k), other_wgt[t], counters);
}

cudaError_t launch_resolve_crypto(const uint32_t *key, const float *wgt, const float4 *total, int slots, int rank, int xres, int x0, int y0,
                                  int w, int h, float4 *out, uint8_t *brk, cudaStream_t stream) {
  if (w <= 0 || h <= 0) return cudaSuccess;
  k_resolve_crypto<<<h, 256, 0, stream>>>(key, wgt, total, slots, rank, xres, x0, y0, w, out, brk);
  return cudaGetLastError();
}

cudaError_t launch_crypto_merge(uint32_t *key, float *wgt, const uint32_t *other_key, const float *other_wgt, size_t npx, int slots,
                                FilterCounters *counters, cudaStream_t stream) {
  const size_t n = npx * (size_t)slots;
  if (n == 0) return cudaSuccess;
  k_crypto_merge<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(key, wgt, other_key, other_wgt, n, slots, counters);
  return cudaGetLastError();
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

cudaError_t launch_resolve_linear(const float4 *buffer, const float *weight, int filter, int role, size_t first, size_t count, float4 *out,
                                  cudaStream_t stream) {
  // a frame-wide row: k_resolve with xres = w = count chunks of at most 2^30 pixels
  for (size_t done = 0; done < count;) {
    const size_t m = count - done < ((size_t)1 << 30) ? count - done : ((size_t)1 << 30);
    dim3 grid((unsigned)((m + 255) / 256), 1);
    k_resolve<<<grid, 256, 0, stream>>>(buffer + first + done, weight + first + done, filter, role, (int)m, 0, 0, (int)m, 1, out + done);
    done += m;
  }
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
