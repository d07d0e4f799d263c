// camera_kernels.cu — K1 launchers: table-driven camera_create_ray kernel + pinhole reverse ray.
#include "camera_kernels.cuh"
#include "lentil_internal.h"
#include "unrolled_dispatch.h"

namespace lb {

constexpr int kRayBlock = 128;

__global__ void __launch_bounds__(kRayBlock)
k_create_rays_table(const __grid_constant__ LensTable lens, const __grid_constant__ CamConsts<float> cam,
                    const __grid_constant__ RayIO io, size_t n, uint64_t ray_id_base) {
  const size_t i = (size_t)blockIdx.x * kRayBlock + threadIdx.x;
  if (i >= n) return;
  const TableEval<float> ev(lens);
  camera_create_ray(ev, cam, io, i, ray_id_base);
}

// camera_reverse_ray, lentil_camera.cpp:164-172
__global__ void k_reverse_rays(size_t n, const float4 *__restrict__ Po, float2 *__restrict__ Ps, float tan_fov) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = Po[i];
  // double coeff = 1.0 / std::max(std::abs(Po.z * tan_fov), 1e-3)
  const double coeff = 1.0 / fmax(fabs((double)p.z * (double)tan_fov), 1e-3);
  Ps[i] = make_float2((float)(p.x * coeff), (float)(p.y * coeff));
}

bool has_unrolled_kernel(int lens_model) { return unrolled_fw_launcher(lens_model) != nullptr; }

cudaError_t launch_create_rays(int lens_kernel, const LensTable &lens, const CamConsts<float> &cam, const RayIO &io, size_t n,
                               uint64_t ray_id_base, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  if (lens_kernel >= 0) {
    if (auto fn = unrolled_fw_launcher(lens_kernel)) return fn(cam, io, n, ray_id_base, stream);
  }
  const unsigned grid = (unsigned)((n + kRayBlock - 1) / kRayBlock);
  k_create_rays_table<<<grid, kRayBlock, 0, stream>>>(lens, cam, io, n, ray_id_base);
  return cudaGetLastError();
}

cudaError_t launch_reverse_rays(size_t n, const float4 *Po, float2 *Ps, float tan_fov, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  k_reverse_rays<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(n, Po, Ps, tan_fov);
  return cudaGetLastError();
}

}  // namespace lb
