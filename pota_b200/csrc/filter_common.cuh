// filter_common.cuh — AOV accumulate primitives shared by the classify and splat kernels.
#pragma once
#include "lens_device.cuh"
#include "lentil_internal.h"

namespace lb {

// value of AOV `a` for sample i as filter_pixel gathers it (lentil_filter.cpp:206-234)
LB_DEV float4 aov_value(const AovSet &aovs, const SampleIO &s, int a, size_t i, float debug_val) {
  if (aovs.role[a] == 2 /*LB_AOV_LENTIL_DEBUG*/) return make_float4(debug_val, debug_val, debug_val, debug_val);
  const float4 *src = aovs.values[a] ? aovs.values[a] : s.rgba;
  return __ldg(src + i);
}

// closest-filter ordering key: smaller |depth| wins; among equal depths the later sample wins, which is
// what the reference's sequential `abs(depth) <= zbuffer[px]` overwrite produces (lentil.h:832-837)
LB_DEV unsigned long long closest_key(float depth, uint64_t sample_global) {
  const float ad = fabsf(depth);
  const unsigned dbits = ad == 0.0f ? 0x7F800000u : __float_as_uint(ad);  // depth 0 behaves as "empty" there
  return ((unsigned long long)dbits << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)sample_global);
}

// Camera::add_to_buffer, lentil.h:823-851.  Gaussian AOVs: one 16-byte vector reduction
// (red.global.add.v4.f32) per AOV + one scalar reduction into filter_weight_buffer for the RGBA AOV.
// Closest AOVs: 64-bit atomicMin on the depth key; the winning sample's value is fetched afterwards
// (k_closest_gather).
LB_DEV void add_to_buffer(const AovSet &aovs, int a, unsigned pixel, float4 v, float add_energy, float depth,
                          float filter_weight, const float rgb_weight[3], uint64_t sample_global) {
  if (aovs.filter[a] == 0 /*gaussian*/) {
    if (aovs.role[a] == 1 /*RGBA*/) atomicAdd(aovs.weight + pixel, filter_weight);
    float4 r;
    r.x = (v.x + add_energy) * filter_weight * rgb_weight[0];
    r.y = (v.y + add_energy) * filter_weight * rgb_weight[1];
    r.z = (v.z + add_energy) * filter_weight * rgb_weight[2];
    r.w = (v.w + add_energy) * filter_weight;
    atomicAdd(aovs.buffer[a] + pixel, r);
  } else {
    if (aovs.role[a] != 2) atomicMin(aovs.zkey + pixel, closest_key(depth, sample_global));
    else if (v.x != 0.0f) atomicMin(aovs.zkey_debug + pixel, closest_key(depth, sample_global));
  }
}

}  // namespace lb
