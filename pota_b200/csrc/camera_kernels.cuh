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
  const float len2 = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
  const float inv = len2 > 0.f ? rsqrtf(len2) : (len2 == 0.f ? 0.f : len2);  // AiV3Normalize: 1 / length, 0 for the null vector (NaN stays NaN)
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
  const uint32_t ray_id = ray_seed_word(ray_id_base + i);
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

// ---- two rays per thread: the polynomials run on the packed FP32 instructions of sm_100a ----------------------
// K1 is issue-bound, not FMA-pipe-bound (ncu r01: issue slots 80 % busy, FMA pipe 65 %): FFMA2/FMUL2/FADD2 carry two
// independent FP32 lanes per issued instruction, so evaluating the SAME polynomial body for two rays at once halves
// the issue slots of the polynomial code while every half stays bit-identical to the scalar body.  A thread owns
// rays j and j + ceil(n/2) (both halves' loads and stores stay fully coalesced) and walks their three traces in
// lockstep; whatever is not polynomial runs once per half as before.

// pt_sample_aperture for two traces in lockstep; a half that has converged (or is off) is no longer updated
template <typename E>
LB_DEV void pt_sample_aperture2(const E &ev, const float x[2], const float y[2], float dx[2], float dy[2], float lambda,
                                const float ax[2], const float ay[2], float dist, const bool on[2]) {
  float sqr_err[2] = {on[0] ? 3.402823466e+38f : 0.f, on[1] ? 3.402823466e+38f : 0.f};
  for (int k = 0; k < 5 && (sqr_err[0] > 1e-4f || sqr_err[1] > 1e-4f); k++) {
    float2 b[5];
    b[0] = make_float2(x[0] + dist * dx[0], x[1] + dist * dx[1]);
    b[1] = make_float2(y[0] + dist * dy[0], y[1] + dist * dy[1]);
    b[2] = make_float2(dx[0], dx[1]);
    b[3] = make_float2(dy[0], dy[1]);
    b[4] = make_float2(lambda, lambda);
    float2 ap[2], J[4];
    ev.ap_jac2(b, ap, J);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!(sqr_err[h] > 1e-4f)) continue;
      const float J0 = lo_hi(J[0], h), J1 = lo_hi(J[1], h), J2 = lo_hi(J[2], h), J3 = lo_hi(J[3], h);
      const float invdet = t_rcp(J0 * J3 - J1 * J2);
      const float e0 = ax[h] - lo_hi(ap[0], h), e1 = ay[h] - lo_hi(ap[1], h);
      dx[h] += (J3 * invdet) * e0;
      dx[h] += (-J1 * invdet) * e1;
      dy[h] += (-J2 * invdet) * e0;
      dy[h] += (J0 * invdet) * e1;
      sqr_err[h] = e0 * e0 + e1 * e1;
    }
  }
}

// Camera::trace_ray_fw_po (lentil.h:283-427) for two rays in lockstep, up to the outer-pupil point `out` of each half.
// `ux, uy` carry the aperture sample of each half: the main trace writes it, the derivative traces reuse it (they are
// called with the same r1, r2: lentil_camera.cpp:111-112).  A derivative trace that fails would repeat identical work
// vignetting_retries times (lentil.h:296,313): one pass suffices.  Code that is not polynomial and not needed in
// lockstep runs per half in ROLLED loops (pick/put): one copy in the instruction cache.
template <typename E>
LB_DEV void trace_pair_fw_po(const E &ev, const CamConsts<float> &cam, const float sx[2], const float sy[2], float (&r1)[2], float (&r2)[2],
                             float (&ux)[2], float (&uy)[2], bool deriv_ray, const uint32_t (&ray_id)[2], const bool valid[2],
                             float (&out)[4][2], bool (&success)[2], int (&tries)[2]) {
  bool active[2] = {valid[0], valid[1]};
  success[0] = success[1] = false;
#pragma unroll
  for (int k = 0; k < 4; ++k) out[k][0] = out[k][1] = 0.f;
  tries[0] = tries[1] = 0;
  while (active[0] || active[1]) {
    if (cam.enable_dof && !deriv_ray) {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        if (!pick(active, h)) continue;
        float a = pick(r1, h), b = pick(r2, h);
        const int tr = pick(tries, h);
        if (tr > 0) {  // retry lens sample: counter RNG instead of the global xor128 (lentil.h:313-316)
          uint32_t seed = tea8(pick(ray_id, h), (uint32_t)tr);
          a = lcg_rng(seed);
          b = lcg_rng(seed);
          put(r1, h, a);
          put(r2, h, b);
        }
        float u, v;
        if (cam.bokeh_n > 0) bokeh_sample(cam, a, b, u, v);
        else if (cam.blades < 2) concentric_disk_sample(a, b, u, v);
        else sample_triangular_aperture(u, v, a, b, 1.0f, cam.blades);
        put(ux, h, u);
        put(uy, h, v);
      }
    }
    float x[2], y[2], dx[2] = {0.f, 0.f}, dy[2] = {0.f, 0.f}, ax[2], ay[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      x[h] = sx[h] * cam.sensor_half;
      y[h] = sy[h] * cam.sensor_half;
      ax[h] = ux[h] * cam.aperture_radius;
      ay[h] = uy[h] * cam.aperture_radius;
    }
    if (cam.enable_dof) pt_sample_aperture2(ev, x, y, dx, dy, cam.lambda, ax, ay, cam.sensor_shift, active);
    float2 b[5], o2[4], t2;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      x[h] += dx[h] * cam.sensor_shift;  // move to beginning of polynomial (lentil.h:349-350)
      y[h] += dy[h] * cam.sensor_shift;
    }
    b[0] = make_float2(x[0], x[1]);
    b[1] = make_float2(y[0], y[1]);
    b[2] = make_float2(dx[0], dx[1]);
    b[3] = make_float2(dy[0], dy[1]);
    b[4] = make_float2(cam.lambda, cam.lambda);
    ev.out5_2(b, o2, t2);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!active[h]) continue;
#pragma unroll
      for (int k = 0; k < 4; ++k) out[k][h] = lo_hi(o2[k], h);
      const float transmittance = fmaxf(0.f, lo_hi(t2, h));
      const float px = x[h] + dx[h] * cam.bfl, py = y[h] + dy[h] * cam.bfl;
      const bool fail = transmittance <= 0.f || out[0][h] * out[0][h] + out[1][h] * out[1][h] > cam.outer_pupil_r2 ||
                        px * px + py * py > cam.inner_pupil_r2;
      if (!fail) { success[h] = true; active[h] = false; }
      else { ++tries[h]; if (deriv_ray || tries[h] > cam.vignetting_retries) active[h] = false; }
    }
  }
}

// outer-pupil point -> camera-space ray in scene units, normalised (lentil.h:387-425)
LB_DEV void finish_fw_ray(const CamConsts<float> &cam, const float out[4], bool success, float o[3], float d[3], bool &ok) {
  float pos[3], dir[3];
  outer_to_cs(cam, out, pos, dir);
  // origin/direction *= -{1,.1,.01,.001} (lentil.h:395-416), then AiV3Normalize
#pragma unroll
  for (int k = 0; k < 3; ++k) { o[k] = pos[k] * cam.unit_scale; dir[k] *= cam.unit_scale; }
  const float len2 = dir[0] * dir[0] + dir[1] * dir[1] + dir[2] * dir[2];
  const float inv = len2 > 0.f ? rsqrtf(len2) : (len2 == 0.f ? 0.f : len2);  // AiV3Normalize: 1 / length, 0 for the null vector (NaN stays NaN)
  ok = success;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    d[k] = dir[k] * inv;
    if (o[k] != o[k] || d[k] != d[k]) ok = false;  // NaN bailout (lentil.h:421-425)
  }
}

// camera_create_ray (lentil_camera.cpp:78-125) for rays j and j + ceil(n/2)
template <typename E>
LB_DEV void camera_create_ray_pair(const E &ev, const CamConsts<float> &cam, const RayIO &io, size_t j, size_t n, uint64_t ray_id_base) {
  const size_t half = (n + 1) / 2;
  const size_t i[2] = {j, j + half};
  const bool valid[2] = {true, i[1] < n};
  float sx[2], sy[2], dsx[2], dsy[2], r1[2], r2[2], ux[2] = {0.f, 0.f}, uy[2] = {0.f, 0.f};
  uint32_t ray_id[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const size_t k = valid[h] ? i[h] : i[0];
    sx[h] = __ldg(io.sx + k); sy[h] = __ldg(io.sy + k);
    dsx[h] = __ldg(io.dsx + k); dsy[h] = __ldg(io.dsy + k);
    r1[h] = __ldg(io.lensx + k); r2[h] = __ldg(io.lensy + k);
    ray_id[h] = ray_seed_word(ray_id_base + k);
  }
  const float step = 0.001f;
  const float fd = step * cam.deriv_baseline;
  const float inv_fd = 1.0f / fd;
  const size_t P = io.plane;
  float mo[3][2], md[3][2];
  // main rays, then the two pairs of differential rays, through ONE copy of the packed trace code
#pragma unroll 1
  for (int t = 0; t < 3; ++t) {
    float tsx[2], tsy[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      tsx[h] = t == 1 ? sx[h] + (dsx[h] * fd) : sx[h];
      tsy[h] = t == 2 ? sy[h] + (dsy[h] * fd) : sy[h];
    }
    float out[4][2];
    bool success[2];
    int tr[2];
    trace_pair_fw_po(ev, cam, tsx, tsy, r1, r2, ux, uy, t != 0, ray_id, valid, out, success, tr);
    float *po = t == 0 ? io.origin : (t == 1 ? io.dOdx : io.dOdy);
    float *pd = t == 0 ? io.dir : (t == 1 ? io.dDdx : io.dDdy);
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      if (!pick(valid, h)) continue;
      const float o4[4] = {pick(out[0], h), pick(out[1], h), pick(out[2], h), pick(out[3], h)};
      float o[3], d[3];
      bool ok;
      finish_fw_ray(cam, o4, pick(success, h), o, d, ok);
      const size_t ih = pick(i, h);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        if (t == 0) {
          put(mo[k], h, o[k]);
          put(md[k], h, d[k]);
          if (po) po[k * P + ih] = o[k];
          if (pd) pd[k * P + ih] = d[k];
          if (io.weight) io.weight[k * P + ih] = ok ? cam.exposure : 0.f * cam.exposure;
        } else {
          if (po) po[k * P + ih] = (o[k] - pick(mo[k], h)) * inv_fd;
          if (pd) pd[k * P + ih] = (d[k] - pick(md[k], h)) * inv_fd;
        }
      }
      if (t == 0 && io.tries) io.tries[ih] = pick(tr, h);
    }
  }
}

}  // namespace lb
