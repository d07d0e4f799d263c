// unrolled_dispatch.h — lookup of the per-lens unrolled kernels generated into gen/ (one
// translation unit per lens, see pota_b200/lensgen/emit_cuda.py).  Returns nullptr when the lens
// has no unrolled kernel in this build; callers then use the table-driven kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

#include "lens_table.h"

namespace lb {
struct RayIO;
struct FilterConsts;
struct AovSet;
struct SampleIO;
struct WorkItem;
struct FilterCounters;

using FwLauncher = cudaError_t (*)(const CamConsts<float> &cam, const RayIO &io, size_t n, uint64_t ray_id_base, cudaStream_t stream);
using BwLauncher = cudaError_t (*)(const CamConsts<float> &cam, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s,
                                   const WorkItem *work, FilterCounters *counters, uint64_t sample_base, int grid, cudaStream_t stream);
// Which generation of the per-lens bodies the launchers pick (A/B timing, cross-checks in the tests):
//   LB_KERNEL_GEN=2 (default)  wavelength-folded bodies; immediates when the wavelength is the default 550 nm, coefficient table otherwise
//   LB_KERNEL_GEN=1            wavelength-folded bodies, always the coefficient table
//   LB_KERNEL_GEN=0            K2: first-generation 5-variate body (the one chromatic aberration always uses); K1 as 1
inline int kernel_generation() {
  const char *e = getenv("LB_KERNEL_GEN");
  return (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
}
FwLauncher unrolled_fw_launcher(int lens_model);
BwLauncher unrolled_bw_launcher(int lens_model);
}  // namespace lb
