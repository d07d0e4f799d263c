// camera_kernels.cuh — K1: camera_create_ray for a batch, one ray (= 3 forward traces) per thread.
//
// Reference: camera_create_ray /root/reference/src/lentil_camera.cpp:78-125 calling
// Camera::trace_ray_fw_po /root/reference/src/lentil.h:283-427.  The kernel is templated on the
// polynomial evaluator so the same body serves the table-driven lens (any lens in the pack, chosen
// on the host) and the per-lens unrolled evaluators of gen/.
#pragma once
#include "lens_device.cuh"

namespace lb {

struct RayIO {
  const float *sx, *sy, *dsx, *dsy, *lensx, *lensy;
  float *origin, *dir, *dOdx, *dOdy, *dDdx, *dDdy, *weight;  // [3][n] planes
  int32_t *tries;
  size_t plane;  // plane stride in floats (n of the whole batch)
};

struct FwRay {
  float o[3], d[3];
  bool ok;
};

// one Camera::trace_ray_fw_po call (lentil.h:283-427)
template <typename E>
LB_DEV FwRay trace_ray_fw_po(const E &ev, const CamConsts<float> &cam, float sx, float sy, float &r1, float &r2,
                             bool deriv_ray, uint32_t ray_id, int &tries) {
  tries = 0;
  bool ray_succes = false;
  float out[4] = {0.f, 0.f, 0.f, 0.f};
  while (!ray_succes && tries <= cam.vignetting_retries) {
    float x = sx * cam.sensor_half, y = sy * cam.sensor_half, dx = 0.f, dy = 0.f;
    float ux = 0.f, uy = 0.f;
    if (cam.enable_dof) {
      if (!deriv_ray && tries > 0) {  // retry lens sample: counter RNG instead of the global xor128 (lentil.h:313-316)
        uint32_t seed = tea8(ray_id, (uint32_t)tries);
        r1 = lcg_rng(seed);
        r2 = lcg_rng(seed);
      }
      if (cam.bokeh_n > 0) bokeh_sample(cam, r1, r2, ux, uy);
      else if (cam.blades < 2) concentric_disk_sample(r1, r2, ux, uy);
      else sample_triangular_aperture(ux, uy, r1, r2, 1.0f, cam.blades);
    }
    const float ax = ux * cam.aperture_radius, ay = uy * cam.aperture_radius;
    if (cam.enable_dof) pt_sample_aperture(ev, x, y, dx, dy, cam.lambda, ax, ay, cam.sensor_shift);
    x += dx * cam.sensor_shift;  // move to beginning of polynomial (lentil.h:349-350)
    y += dy * cam.sensor_shift;
    const float b[5] = {x, y, dx, dy, cam.lambda};
    out[0] = out[1] = out[2] = out[3] = 0.f;  // `out.setZero()` at loop top: a failed last try leaves the evaluated values
    float tr;
    ev.out5(b, out, tr);
    const float transmittance = fmaxf(0.f, tr);
    if (transmittance <= 0.f) { ++tries; continue; }
    if (out[0] * out[0] + out[1] * out[1] > cam.outer_pupil_r2) { ++tries; continue; }
    const float px = x + dx * cam.bfl, py = y + dy * cam.bfl;
    if (px * px + py * py > cam.inner_pupil_r2) { ++tries; continue; }
    ray_succes = true;
  }
  FwRay r;
  r.ok = ray_succes;
  float pos[3], dir[3];
  outer_to_cs(cam, out, pos, dir);
  // origin/direction *= -{1,.1,.01,.001} (lentil.h:395-416), then AiV3Normalize
#pragma unroll
  for (int k = 0; k < 3; ++k) { r.o[k] = pos[k] * cam.unit_scale; dir[k] *= cam.unit_scale; }
  const float len = sqrtf(dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2]);
  const float inv = len != 0.f ? 1.0f / len : 0.f;
#pragma unroll
  for (int k = 0; k < 3; ++k) r.d[k] = dir[k] * inv;
  // NaN bailout (lentil.h:421-425)
  if (r.o[0] != r.o[0] || r.o[1] != r.o[1] || r.o[2] != r.o[2] || r.d[0] != r.d[0] || r.d[1] != r.d[1] || r.d[2] != r.d[2]) r.ok = false;
  return r;
}

// camera_create_ray (lentil_camera.cpp:78-125)
template <typename E>
LB_DEV void camera_create_ray(const E &ev, const CamConsts<float> &cam, const RayIO &io, size_t i, uint64_t ray_id_base) {
  const float sx = __ldg(io.sx + i), sy = __ldg(io.sy + i);
  const float dsx = __ldg(io.dsx + i), dsy = __ldg(io.dsy + i);
  float r1 = __ldg(io.lensx + i), r2 = __ldg(io.lensy + i);
  const uint32_t ray_id = (uint32_t)(ray_id_base + i);
  const float step = 0.001f;
  int tries = 0;
  // differential rays: baseline `step` of the reference, optionally stretched ("differentials")
  const float fd = step * cam.deriv_baseline;
  const float sx_dx = sx + (dsx * fd), sy_dy = sy + (dsy * fd);
  // main ray, then the two differential rays (lentil_camera.cpp:93,111-112) through ONE copy of the trace
  // code: the unrolled polynomial bodies are several KB of straight-line code each
  FwRay m, ax, ay;
#pragma unroll 1
  for (int t = 0; t < 3; ++t) {
    int tr;
    const FwRay r = trace_ray_fw_po(ev, cam, t == 1 ? sx_dx : sx, t == 2 ? sy_dy : sy, r1, r2, t != 0, ray_id, tr);
    if (t == 0) { m = r; tries = tr; }
    else if (t == 1) ax = r;
    else ay = r;
  }
  const float w = m.ok ? cam.exposure : 0.f * cam.exposure;
  const size_t P = io.plane;
  const float inv_fd = 1.0f / fd;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (io.origin) io.origin[k * P + i] = m.o[k];
    if (io.dir) io.dir[k * P + i] = m.d[k];
    if (io.dOdx) io.dOdx[k * P + i] = (ax.o[k] - m.o[k]) * inv_fd;
    if (io.dOdy) io.dOdy[k * P + i] = (ay.o[k] - m.o[k]) * inv_fd;
    if (io.dDdx) io.dDdx[k * P + i] = (ax.d[k] - m.d[k]) * inv_fd;
    if (io.dDdy) io.dDdy[k * P + i] = (ay.d[k] - m.d[k]) * inv_fd;
    if (io.weight) io.weight[k * P + i] = w;
  }
  if (io.tries) io.tries[i] = tries;
}

}  // namespace lb
