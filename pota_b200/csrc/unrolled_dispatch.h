// unrolled_dispatch.h — lookup of the per-lens unrolled kernels generated into gen/ (one
// translation unit per lens, see pota_b200/lensgen/emit_cuda.py).  Returns nullptr when the lens
// has no unrolled kernel in this build; callers then use the table-driven kernels.
#pragma once
#include <cuda_runtime.h>

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
FwLauncher unrolled_fw_launcher(int lens_model);
BwLauncher unrolled_bw_launcher(int lens_model);
}  // namespace lb
