// filter_kernels.cuh — K2b: reverse-trace + splat of the redistributed samples, templated on the
// polynomial evaluator (table-driven or per-lens unrolled).
//
// Reference: the PolynomialOptics loop of filter_pixel, /root/reference/src/lentil_filter.cpp:248-300,
// calling Camera::trace_ray_bw_po (/root/reference/src/lentil.h:573-661) and Camera::add_to_buffer
// (lentil.h:823-851).
//
// Mapping: one WARP per redistributed source sample, fetched from the work list with an atomic
// counter (persistent CTAs).  The reference's loop
//     for (count = 0; count < samples && total < 5*samples; ++count, ++total) { ...fail -> --count }
// is sequential only through its stop rule; attempt `total` itself depends on nothing but `total`
// (seed = tea<8>(px*py+px, total+tries)).  Each round the warp runs min(32, samples-count,
// max_total-total) consecutive attempts, one per lane — the sequential loop is guaranteed to execute
// at least that many more, because `count` grows by at most one per attempt — so the set of splats is
// exactly the reference's.
#pragma once
#include "filter_common.cuh"

namespace lb {

// Camera::trace_ray_bw_po, lentil.h:573-661 (AiTraceProbe occlusion test: no scene here, never occluded)
template <typename E>
LB_DEV bool trace_ray_bw_po(const E &ev, const CamConsts<float> &cam, const float target[3], uint32_t seed_base, uint32_t total,
                            float lambda, float &sx, float &sy, unsigned &newton_its) {
  int tries = 0;
  float ax = 0.f, ay = 0.f;
  while (tries <= cam.vignetting_retries) {
    if (!cam.enable_dof) { ax = ay = 0.f; }
    else {
      uint32_t seed = tea8(seed_base, total + (uint32_t)tries);
      if (cam.blades <= 2) {
        float ux, uy;
        if (cam.bokeh_n > 0) {
          // bokehSample(rng, rng, unit_disk, rng, rng): g++ evaluates the arguments right to left, so the
          // two unused stratification numbers draw first, then column, then row
          lcg_rng(seed); lcg_rng(seed);
          const float col = lcg_rng(seed);
          const float row = lcg_rng(seed);
          bokeh_sample(cam, row, col, ux, uy);
        } else {
          const float oy = lcg_rng(seed);
          const float ox = lcg_rng(seed);
          concentric_disk_sample(ox, oy, ux, uy);
        }
        ax = ux * cam.aperture_radius;
        ay = uy * cam.aperture_radius;
      } else {
        const float r2 = lcg_rng(seed);
        const float r1 = lcg_rng(seed);
        sample_triangular_aperture(ax, ay, r1, r2, cam.aperture_radius, cam.blades);
      }
    }
    float sensor[4], out[4];
    int its;
    const float T = lt_sample_aperture(ev, cam, target, ax, ay, lambda, sensor, out, &its);
    newton_its += (unsigned)its;
    if (T <= 0.f) { ++tries; continue; }
    const float px = sensor[0] + sensor[2] * cam.bfl, py = sensor[1] + sensor[3] * cam.bfl;
    if (px * px + py * py > cam.inner_pupil_r2) { ++tries; continue; }
    sx = sensor[0] + sensor[2] * -cam.sensor_shift;  // shift sensor (lentil.h:654-655)
    sy = sensor[1] + sensor[3] * -cam.sensor_shift;
    return true;
  }
  return false;
}

// sensor position -> pixel index or -1 (lentil_filter.cpp:276-290), in double like the reference
LB_DEV int sensor_to_pixel(const FilterConsts &fc, float sx, float sy) {
  const double s0 = (double)sx / fc.sensor_half;
  const double s1 = (double)sy / fc.sensor_half * fc.aspect_full;
  const double p0 = (((s0 + 1.0) / 2.0) * (double)fc.xres_full) - (double)fc.region_min_x;
  const double p1 = (((-s1 + 1.0) / 2.0) * (double)fc.yres_full) - (double)fc.region_min_y;
  if ((p0 >= (double)fc.xres) || (p0 < 0) || (p1 >= (double)fc.yres) || (p1 < 0) || (p0 != p0) || (p1 != p1)) return -1;
  return (int)floor(p0) + (int)floor(p1) * fc.xres;
}

template <typename E>
LB_DEV void splat_work_item(const E &ev, const CamConsts<float> &cam, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s,
                            const WorkItem &w, FilterCounters *counters, uint64_t sample_base) {
  const int lane = threadIdx.x & 31;
  const size_t i = w.sample;
  const int px = __ldg(s.px + i), py = __ldg(s.py + i);
  const float depth = __ldg(s.pos_cs + i).w;
  // -camera_space_sample_position * 10.0 (lentil_filter.cpp:271)
  const float target[3] = {(float)(-(double)w.csp[0] * 10.0), (float)(-(double)w.csp[1] * 10.0), (float)(-(double)w.csp[2] * 10.0)};
  const int samples = (int)w.n_samples;
  const unsigned max_total = (unsigned)samples * 5u;
  const float inv_samples = (float)(1.0 / (double)(float)samples);
  const float weight = 1.0f * s.inv_density * inv_samples;  // filter_weight * inverse_sample_density * inv_samples (:297)
  const uint32_t seed_base = (uint32_t)(px * py + px);
  const bool chroma = fc.abb_chromatic > 0.0f;
  const int nchan = chroma ? 3 : 1;
  int count = 0;
  unsigned total = 0;
  unsigned n_splats = 0, n_attempts = 0, n_its = 0;
  while (count < samples && total < max_total) {
    const int batch = min(32, min(samples - count, (int)(max_total - total)));
    int delta = 0;
    // phase 1 (divergent): each lane reverse-traces its attempt, once per colour channel; ONE copy of the
    // Newton code (the unrolled polynomial bodies are ~20 KB of straight-line code)
    int pix0 = -1, pix1 = -1, pix2 = -1;
    if (lane < batch) {
      const uint32_t t = total + (uint32_t)lane;
      int fails = 0;
#pragma unroll 1
      for (int ch = 0; ch < nchan; ++ch) {
        float lambda = 0.55f;
        if (chroma) {  // lentil_filter.cpp:257-267
          if (ch == 0) lambda = 0.35f + (1.0f - fc.abb_chromatic) * (0.55f - 0.35f);
          else if (ch == 2) lambda = 0.55f + fc.abb_chromatic * (0.85f - 0.55f);
        }
        float sx, sy;
        ++n_attempts;
        int pixel = -1;
        if (trace_ray_bw_po(ev, cam, target, seed_base, t, lambda, sx, sy, n_its)) pixel = sensor_to_pixel(fc, sx, sy);
        if (pixel < 0) ++fails;
        else ++n_splats;
        if (ch == 0) pix0 = pixel;
        else if (ch == 1) pix1 = pixel;
        else pix2 = pixel;
      }
      delta = 1 - fails;
    }
    __syncwarp();
    // phase 2 (warp-converged): the AOV values of the source sample are warp-uniform, the pixel is per lane
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      if (ch >= nchan) break;
      const int pixel = ch == 0 ? pix0 : (ch == 1 ? pix1 : pix2);
      float rgbw[3] = {1.f, 1.f, 1.f};
      if (chroma) { rgbw[0] = ch == 0 ? 3.f : 0.f; rgbw[1] = ch == 1 ? 3.f : 0.f; rgbw[2] = ch == 2 ? 3.f : 0.f; }
      if (!__any_sync(0xffffffffu, pixel >= 0)) continue;
      for (int a = 0; a < fc.n_aov; ++a) {
        const float4 v = aov_value(aovs, s, a, i, (float)samples);
        if (pixel >= 0) add_to_buffer(aovs, a, (unsigned)pixel, v, w.add_energy, depth, weight, rgbw, sample_base + i);
      }
    }
    count += __reduce_add_sync(0xffffffffu, delta);
    total += (unsigned)batch;
  }
  n_splats = __reduce_add_sync(0xffffffffu, n_splats);
  n_attempts = __reduce_add_sync(0xffffffffu, n_attempts);
  n_its = __reduce_add_sync(0xffffffffu, n_its);
  if (lane == 0) {
    atomicAdd(&counters->splats, (unsigned long long)n_splats);
    atomicAdd(&counters->attempts, (unsigned long long)n_attempts);
    atomicAdd(&counters->newton_its, (unsigned long long)n_its);
  }
}

// persistent kernel body: warps pull work items until the list is drained
template <typename E>
LB_DEV void splat_persistent(const E &ev, const CamConsts<float> &cam, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s,
                             const WorkItem *__restrict__ work, FilterCounters *counters, uint64_t sample_base) {
  const int lane = threadIdx.x & 31;
  const unsigned n_work = *((volatile unsigned *)&counters->work_count);
  for (;;) {
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(&counters->work_next, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= n_work) break;
    const WorkItem w = work[idx];
    splat_work_item(ev, cam, fc, aovs, s, w, counters, sample_base);
  }
}

}  // namespace lb
