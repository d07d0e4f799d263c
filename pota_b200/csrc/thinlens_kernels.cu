// thinlens_kernels.cu — the thin-lens camera model (default `camera_type`, lentil_camera.cpp:20) on the GPU.
// Compiled with -fmad=false: the model is a handful of float operations per ray whose results decide the
// pixel a splat lands in, so they are evaluated operation by operation as the reference's host code does.
//
// Reference: Camera::trace_ray_fw_thinlens /root/reference/src/lentil.h:431-569 (forward) and the ThinLens
// case of filter_pixel /root/reference/src/lentil_filter.cpp:303-447 (bidirectional), helpers lens.h:477-582.
//
// Roofline: both kernels are memory/reduction bound, not FP32 bound — ~300 float operations per trace against
// 108 B per camera ray (HBM) or one 16-byte vector reduction + one 4-byte reduction per splat and AOV (L2).
#include "camera_kernels.cuh"
#include "filter_common.cuh"
#include "lentil_internal.h"

namespace lb {

struct V3f { float x, y, z; };
LB_DEV float dot3f(V3f a, V3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LB_DEV V3f normalize3f(V3f a) {  // AiV3Normalize
  float t = sqrtf(dot3f(a, a));
  if (t != 0.f) t = 1.0f / t;
  return {a.x * t, a.y * t, a.z * t};
}
LB_DEV float lerpf(float perc, float a, float b) { return a + perc * (b - a); }  // linear_interpolate, global.h:3-5

// concentricDiskSample, lens.h:477-517 (float arithmetic with double constants, as there)
LB_DEV void concentric_disk_sample_tl(float ox, float oy, double &lx, double &ly, float bias, float squarelerp) {
  if (ox == 0.0f && oy == 0.0f) { lx = 0.0; ly = 0.0; return; }
  float phi, r;
  const float a = (float)(2.0 * (double)ox - 1.0);
  const float b = (float)(2.0 * (double)oy - 1.0);
  if ((a * a) > (b * b)) { r = a; phi = (float)(0.78539816339 * (double)(b / a)); }
  else { r = b; phi = (float)((double)1.57079632679489661923f - (0.78539816339 * (double)(a / b))); }
  if (bias != 0.5f) {  // AiBias(|r|, bias) * sign(r)
    const float ar = fabsf(r);
    const float bb = (ar > 0.f) ? ((bias > 0.f) ? powf(ar, logf(bias) / logf(0.5f)) : 0.f) : 0.f;
    r = bb * (r < 0.f ? -1.f : 1.f);
  }
  const float cos_phi = fast_cos(phi);
  const float sin_phi = fast_sin(phi);
  lx = (double)(r * cos_phi);
  ly = (double)(r * sin_phi);
  if (squarelerp > 0.0f) {
    lx = (double)lerpf(squarelerp, (float)lx, a);
    ly = (double)lerpf(squarelerp, (float)ly, b);
  }
}

// unit disk sample shared by the forward and the bidirectional path (lentil.h:457-472, lentil_filter.cpp:316-322)
LB_DEV void thin_unit_disk(const CamConsts<float> &cam, const ThinConsts &tl, float r1, float r2, double &ux, double &uy) {
  if (cam.bokeh_n > 0) bokeh_sample(cam, r1, r2, ux, uy);
  else if (cam.blades < 2) concentric_disk_sample_tl(r1, r2, ux, uy, tl.abb_spherical, tl.circle_to_square);
  else sample_triangular_aperture(ux, uy, (double)r1, (double)r2, 1.0, cam.blades);
}

// empericalOpticalVignettingSquare, lens.h:532-541
LB_DEV bool optical_vignetting_square(V3f origin, V3f direction, float apertureRadius, const ThinConsts &tl) {
  const float intersection = fabsf(tl.optical_vignetting_distance / direction.z);
  const V3f p{direction.x * intersection - origin.x, direction.y * intersection - origin.y, direction.z * intersection - origin.z};
  const float power = (float)(1.0 + (double)tl.squircle);
  const float radius = apertureRadius * tl.optical_vignetting_radius;
  const float dist = powf(fabsf(p.x), power) + powf(fabsf(p.y), power);
  return !(dist > powf(radius, power));
}

// abb_coma_multipliers, lens.h:566-574 (maximal_projection is a camera constant)
LB_DEV float coma_multiplier(const ThinConsts &tl, V3f dir_from_center, double ux, double uy) {
  const float current_projection = dot3f(dir_from_center, V3f{0.0f, 0.0f, -1.0f});
  const float projection_perc =
      (float)((((double)(current_projection - tl.coma_max_projection) / (1.0 - (double)tl.coma_max_projection)) - 0.5) * 2.0);
  const float dist_from_sensor_center = (float)(1.0 - (double)projection_perc);
  const float dist_from_aperture = (float)sqrt(ux * ux + uy * uy);
  return dist_from_sensor_center * dist_from_aperture;
}

// abb_coma_perturb, lens.h:578-586: rotation about the axis orthogonal to the ray and the optical axis (double, as Eigen)
LB_DEV V3f coma_perturb(V3f dir_from_lens, V3f ray, float abb_coma, bool reverse) {
  // angle 0: the rotation (and its inverse) is exactly the identity and (float)(1*x + 0*y + 0*z) == x, so the
  // FP64 rotation can be skipped without changing a bit (abb_coma defaults to 0, lentil_camera.cpp:38)
  if (abb_coma == 0.0f) return ray;
  const V3f c{dir_from_lens.y * -1.0f - dir_from_lens.z * 0.0f, dir_from_lens.z * 0.0f - dir_from_lens.x * -1.0f,
              dir_from_lens.x * 0.0f - dir_from_lens.y * 0.0f};
  const V3f axis = normalize3f(c);
  const double ax = axis.x, ay = axis.y, az = axis.z;
  const double angle = ((double)abb_coma * 2.3456 * (double)3.14159265358979323846f) / 180.0;
  double si, co;
  sincos(angle, &si, &co);
  const double t = 1.0 - co;
  double m[3][3] = {{t * ax * ax + co, t * ax * ay - si * az, t * ax * az + si * ay},
                    {t * ax * ay + si * az, t * ay * ay + co, t * ay * az - si * ax},
                    {t * ax * az - si * ay, t * ay * az + si * ax, t * az * az + co}};
  if (reverse) {
    double a[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) a[i][j] = m[i][j];
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    m[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / det; m[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det;
    m[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det; m[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) / det;
    m[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det; m[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
    m[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / det; m[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det;
    m[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
  }
  const double rx = ray.x, ry = ray.y, rz = ray.z;
  return {(float)(m[0][0] * rx + m[0][1] * ry + m[0][2] * rz), (float)(m[1][0] * rx + m[1][1] * ry + m[1][2] * rz),
          (float)(m[2][0] * rx + m[2][1] * ry + m[2][2] * rz)};
}

// ---- forward: Camera::trace_ray_fw_thinlens, lentil.h:431-569 ------------------------------------------------
// (ux, uy): the unit-disk lens sample.  The main trace computes it; the derivative traces are called with the same
// r1, r2 (lentil_camera.cpp:111-112) and reuse it instead of sampling again.
LB_DEV FwRay trace_ray_fw_thinlens(const CamConsts<float> &cam, const ThinConsts &tl, float sx, float sy, float &r1, float &r2,
                                   double &ux, double &uy, bool deriv_ray, uint32_t ray_id, int &tries) {
  tries = 0;
  bool ray_succes = false;
  FwRay r;
  r.o[0] = r.o[1] = r.o[2] = 0.f;
  float dir[3] = {0.f, 0.f, 0.f};
  while (!ray_succes && tries <= cam.vignetting_retries) {
    float s0 = sx, s1 = sy;
    if (tl.abb_distortion > 0.0f) {  // barrelDistortion, lens.h:548-551
      const float f = (float)(1. + (double)((s0 * s0 + s1 * s1) * tl.abb_distortion));
      s0 *= f; s1 *= f;
    }
    const V3f p{(float)((double)s0 * tl.sensor_half), (float)((double)s1 * tl.sensor_half), -tl.focal_length};
    const V3f dir_from_center = normalize3f(p);
    if (cam.enable_dof && !deriv_ray) {
      if (tries > 0) {  // counter RNG instead of the global xor128 (lentil.h:460-463), as in the PO path
        uint32_t seed = tea8(ray_id, (uint32_t)tries);
        r1 = lcg_rng(seed);
        r2 = lcg_rng(seed);
      }
      thin_unit_disk(cam, tl, r1, r2, ux, uy);
      ux *= (double)tl.bokeh_anamorphic;
    }
    const V3f lens{(float)(ux * tl.aperture_radius), (float)(uy * tl.aperture_radius), 0.0f};
    const float intersection = (float)fabs(tl.focus_distance / (double)lerpf(0.0f, dir_from_center.z, 1.0f));
    const V3f focusPoint{dir_from_center.x * intersection, dir_from_center.y * intersection, dir_from_center.z * intersection};
    V3f dir_from_lens = normalize3f(V3f{focusPoint.x - lens.x, focusPoint.y - lens.y, focusPoint.z - lens.z});
    // abb_coma == 0 (the default): 0 * (finite multiplier) is +-0 and the perturbation below is the identity
    const float abb_coma_multiplied = tl.abb_coma == 0.0f ? 0.0f : tl.abb_coma * coma_multiplier(tl, dir_from_center, ux, uy);
    dir_from_lens = coma_perturb(dir_from_lens, dir_from_lens, abb_coma_multiplied, false);
    if (tl.optical_vignetting_distance > 0.0f && !deriv_ray) {
      if (!optical_vignetting_square(lens, dir_from_lens, (float)tl.aperture_radius, tl)) { ++tries; continue; }
    }
    r.o[0] = lens.x * tl.unit_scale; r.o[1] = lens.y * tl.unit_scale; r.o[2] = lens.z * tl.unit_scale;
    dir[0] = dir_from_lens.x * tl.unit_scale; dir[1] = dir_from_lens.y * tl.unit_scale; dir[2] = dir_from_lens.z * tl.unit_scale;
    ray_succes = true;
  }
  const V3f d = normalize3f(V3f{dir[0], dir[1], dir[2]});
  r.d[0] = d.x; r.d[1] = d.y; r.d[2] = d.z;
  r.ok = ray_succes;
  return r;
}

__global__ void __launch_bounds__(256)
k_create_rays_thinlens(const __grid_constant__ CamConsts<float> cam, const __grid_constant__ ThinConsts tl, const __grid_constant__ RayIO io,
                       size_t n, uint64_t ray_id_base) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float sx = __ldg(io.sx + i), sy = __ldg(io.sy + i);
  const float dsx = __ldg(io.dsx + i), dsy = __ldg(io.dsy + i);
  float r1 = __ldg(io.lensx + i), r2 = __ldg(io.lensy + i);
  const uint32_t ray_id = ray_seed_word(ray_id_base + i);
  // the thin-lens trace is exact float arithmetic on both sides: the reference's own step is kept (no baseline stretch)
  const float step = 0.001f;
  int tries, td;
  double ux = 0.0, uy = 0.0;
  const FwRay m = trace_ray_fw_thinlens(cam, tl, sx, sy, r1, r2, ux, uy, false, ray_id, tries);
  const FwRay ax = trace_ray_fw_thinlens(cam, tl, sx + (dsx * step), sy, r1, r2, ux, uy, true, ray_id, td);
  const FwRay ay = trace_ray_fw_thinlens(cam, tl, sx, sy + (dsy * step), r1, r2, ux, uy, true, ray_id, td);
  const float w = m.ok ? cam.exposure : 0.f * cam.exposure;
  const size_t P = io.plane;
  const float inv_step = 1.0f / step;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (io.origin) io.origin[k * P + i] = m.o[k];
    if (io.dir) io.dir[k * P + i] = m.d[k];
    if (io.dOdx) io.dOdx[k * P + i] = (ax.o[k] - m.o[k]) * inv_step;
    if (io.dOdy) io.dOdy[k * P + i] = (ay.o[k] - m.o[k]) * inv_step;
    if (io.dDdx) io.dDdx[k * P + i] = (ax.d[k] - m.d[k]) * inv_step;
    if (io.dDdy) io.dDdy[k * P + i] = (ay.d[k] - m.d[k]) * inv_step;
    if (io.weight) io.weight[k * P + i] = w;
  }
  if (io.tries) io.tries[i] = tries;
}

// ---- bidirectional: one attempt of the ThinLens loop of filter_pixel, lentil_filter.cpp:311-447 --------------
// the attempt's lens sample on the unit disk (:311-325); `seed` is left where the chromatic channel draw continues
LB_DEV void thinlens_lens_sample(const CamConsts<float> &cam, const ThinConsts &tl, uint32_t seed_base, uint32_t total, uint32_t &seed, double &ux,
                                 double &uy) {
  seed = tea8(seed_base, total);
  ux = 0.0; uy = 0.0;
  if (cam.bokeh_n > 0) {  // g++ evaluates bokehSample(rng, rng, ., rng, rng) right to left
    lcg_rng(seed); lcg_rng(seed);
    const float col = lcg_rng(seed);
    const float row = lcg_rng(seed);
    bokeh_sample(cam, row, col, ux, uy);
  } else if (cam.blades < 2) {
    const float oy = lcg_rng(seed);
    const float ox = lcg_rng(seed);
    concentric_disk_sample_tl(ox, oy, ux, uy, tl.abb_spherical, tl.circle_to_square);
  } else {
    const float b = lcg_rng(seed);
    const float a = lcg_rng(seed);
    sample_triangular_aperture(ux, uy, (double)a, (double)b, 1.0, cam.blades);
  }
}

// the rest of the attempt (:326-440): the scene point P through the lens point (ux, uy) onto the sensor.  Returns false for
// `--count; continue` (vignetted / off the image); pixel_x, pixel_y = region-relative pixel coordinates, rgbw = channel
// weights of the chromatic aberration.
LB_DEV bool thinlens_project(const ThinConsts &tl, const FilterConsts &fc, V3f P, double ux, double uy, uint32_t &seed, float rgbw[3],
                             float &pixel_x, float &pixel_y) {
  const float image_dist_samplepos = (-tl.focal_length * P.z) / (-tl.focal_length + P.z);
  ux *= (double)tl.bokeh_anamorphic;
  const V3f lens{(float)(ux * tl.aperture_radius), (float)(uy * tl.aperture_radius), 0.0f};
  V3f dir_from_center = normalize3f(P);
  V3f dir_lens_to_P = normalize3f(V3f{P.x - lens.x, P.y - lens.y, P.z - lens.z});
  const float abb_coma_multiplied = tl.abb_coma == 0.0f ? 0.0f : tl.abb_coma * coma_multiplier(tl, dir_from_center, ux, uy);
  dir_lens_to_P = coma_perturb(dir_lens_to_P, dir_from_center, abb_coma_multiplied, true);
  const float lenP = sqrtf(dot3f(P, P));
  const V3f Pp{lenP * dir_lens_to_P.x, lenP * dir_lens_to_P.y, lenP * dir_lens_to_P.z};
  dir_from_center = normalize3f(Pp);
  const float samplepos_image_intersection = fabsf(image_dist_samplepos / dir_from_center.z);
  const V3f ip{dir_from_center.x * samplepos_image_intersection, dir_from_center.y * samplepos_image_intersection,
               dir_from_center.z * samplepos_image_intersection};
  const V3f dir_img = normalize3f(V3f{ip.x - lens.x, ip.y - lens.y, ip.z - lens.z});
  const float fiu = fabsf(tl.image_dist_focusdist / dir_img.z);
  const V3f fu{lens.x + dir_img.x * fiu, lens.y + dir_img.y * fiu, lens.z + dir_img.z * fiu};
  const float spu_x = fu.x / fu.z, spu_y = fu.y / fu.z;
  const float dist_unperturbed = sqrtf((0.0f - spu_x) * (0.0f - spu_x) + (0.0f - spu_y) * (0.0f - spu_y));
  if (tl.optical_vignetting_distance > 0.0f) {
    dir_lens_to_P = normalize3f(V3f{Pp.x - lens.x, Pp.y - lens.y, Pp.z - lens.z});
    if (!optical_vignetting_square(lens, dir_lens_to_P, (float)tl.aperture_radius, tl)) return false;
  }
  float focusdist_intersection = fabsf(tl.image_dist_focusdist / dir_img.z);
  rgbw[0] = rgbw[1] = rgbw[2] = 1.f;
  if (tl.abb_chromatic > 0.0f) {  // channel from the counter RNG (the reference's commented variant, lentil_filter.cpp:395)
    const int channel = (int)(lcg_rng(seed) * 3) - 1;
    rgbw[0] = channel == -1 ? 3.f : 0.f; rgbw[1] = channel == 0 ? 3.f : 0.f; rgbw[2] = channel == 1 ? 3.f : 0.f;
    const float direction_shift = tl.abb_chromatic_type == 0 ? (float)abs(channel) : (float)channel;
    const float shift = direction_shift * tl.abb_chromatic * 5.0f * dist_unperturbed;
    const float aberrated = (float)(((double)-tl.focal_length * -(tl.focus_distance + (double)shift)) /
                                    ((double)-tl.focal_length + -(tl.focus_distance + (double)shift)));
    focusdist_intersection = fabsf(aberrated / dir_img.z);
  }
  const V3f fp{lens.x + dir_img.x * focusdist_intersection, lens.y + dir_img.y * focusdist_intersection, lens.z + dir_img.z * focusdist_intersection};
  float sp_x = fp.x / fp.z, sp_y = fp.y / fp.z;
  {
    const float div = (float)(tl.sensor_half / (double)-tl.focal_length);
    const float c = 1.0f / div;
    sp_x *= c; sp_y *= c;
  }
  if (tl.abb_distortion > 0.0f) {  // inverseBarrelDistortion, lens.h:553-562
    const float b = tl.abb_distortion;
    const float l = sqrtf(sp_x * sp_x + sp_y * sp_y);
    const double bd = b, ld = l;
    const float x0 = (float)pow(9. * bd * bd * ld + sqrt(3.) * sqrt(27. * bd * bd * bd * bd * ld * ld + 4. * bd * bd * bd), 1. / 3.);
    const float x = (float)((double)x0 / (pow(2., 1. / 3.) * pow(3., 2. / 3.) * bd) - pow(2. / 3., 1. / 3.) / (double)x0);
    const float f = x / l;
    sp_x *= f; sp_y *= f;
  }
  const double s0 = (double)sp_x, s1 = (double)sp_y * fc.aspect_full;
  pixel_x = (float)((((s0 + 1.0) / 2.0) * (double)(unsigned)fc.xres_full) - (double)fc.region_min_x);
  pixel_y = (float)((((-s1 + 1.0) / 2.0) * (double)(unsigned)fc.yres_full) - (double)fc.region_min_y);
  if (((double)pixel_x >= (double)fc.xres) || (pixel_x < 0) || ((double)pixel_y >= (double)fc.yres) || (pixel_y < 0)) return false;
  return true;
}

// one attempt: returns the pixel index or -1 (`--count; continue`)
LB_DEV int thinlens_attempt(const CamConsts<float> &cam, const ThinConsts &tl, const FilterConsts &fc, V3f P, uint32_t seed_base, uint32_t total,
                            float rgbw[3], int &ix, int &iy) {
  uint32_t seed;
  double ux, uy;
  thinlens_lens_sample(cam, tl, seed_base, total, seed, ux, uy);
  float pixel_x, pixel_y;
  if (!thinlens_project(tl, fc, P, ux, uy, seed, rgbw, pixel_x, pixel_y)) return -1;
  ix = (int)floorf(pixel_x);
  iy = (int)floorf(pixel_y);
  return ix + iy * fc.xres;
}

// ---- the splat kernel, two accumulate strategies ------------------------------------------------------------------
// kTile == 0: every splat is sent to L2 (red.global.add.v4.f32 + red.global.add.f32), one warp per source sample.
// kTile  > 0: SHARED-MEMORY TILE ACCUMULATION (north_star; Camera::add_to_buffer lentil.h:823-851 is the hot spot).  A CTA takes a
//   batch of kTileBatch consecutive work items -- the samples of a few neighbouring pixels, whose bokeh discs coincide -- and
//   keeps a kTile x kTile window of the RGBA AOV + the filter-weight plane in shared memory, centred where the first item's
//   disc lands.  Splats inside the window are added there (float atomics on shared memory), splats outside it and the other
//   AOVs go to L2 as before; after the batch the window's touched pixels are flushed with one vector + one scalar reduction each.
//   What it buys: the ~2000 x 16 spp x pixels splats that a highlight sends into the same few thousand pixels reach L2 once
//   per batch instead of once per splat -- same-address reductions serialise in the L2 slice -- and at sizes whose planes are
//   not L2-resident (8K, many AOVs) each of those is a DRAM read-modify-write.  What it costs: shared memory (occupancy) and
//   a CTA-wide barrier per batch; discs wider than the window (the C3 frame: 144 px) mostly miss it.
//   MEASURED (B200, profiles/r02_tile_accumulate.txt): the tile kernel loses in every regime tried -- C3 frame 4.12 vs 2.76 ms;
//   a highlight-dense frame with 40-pixel discs (650 M splats, 57 % of them through the window) 30.1 vs 18.8 ms; the same at
//   8K (planes not L2-resident, 63 % through the window) 128.5 vs 82.4 ms.  sm_100a has no shared-memory float add: each of
//   the five adds is an LDS + FADD + ATOMS.CAST.SPIN round trip the warp waits for, where the direct path issues two
//   fire-and-forget reductions that the L2 slices merge at 97 G splats/s (lb_bench_splat_accum) -- on this part the 126 MB L2
//   IS the accumulation tile.  The kernel stays as an opt-in (LB_SPLAT_TILE=1 / auto), parity-tested, default off.
constexpr int kTileBatch = 32;

struct TileView {   // the CTA's window; tile4 == nullptr: no tile (every splat to L2)
  float4 *tile4;
  float *tilew;
  int org_x, org_y, aov;  // window origin (region-relative pixels); index of the AOV kept in the window
};

// kHoistProbes: the instantiation for frames with cryptomatte AOVs (crypto_add_hoisted, filter_common.cuh); frames without them run
// the other one, which is the kernel as it was
template <int kTile, bool kHoistProbes = false>
LB_DEV void thinlens_splat_item(const CamConsts<float> &cam, const ThinConsts &tl, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s,
                                const WorkItem &w, FilterCounters *counters, uint64_t sample_base, const TileView &tv) {
  const int lane = threadIdx.x & 31;
  const size_t i = w.sample;
  const int px = __ldg(s.px + i), py = __ldg(s.py + i);
  const float depth = __ldg(s.pos_cs + i).w;
  const V3f P{w.csp[0], w.csp[1], w.csp[2]};
  const int samples = (int)w.n_samples;
  const int max_total = samples * 5;
  const float inv_samples = (float)(1.0 / (double)(float)samples);
  const float weight = 1.0f * s.inv_density * inv_samples;
  const uint32_t seed_base = (uint32_t)(px * py + px);
  int count = 0, total = 0;
  unsigned n_splats = 0, n_attempts = 0, n_tile = 0;
  float4 tv_value = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kTile > 0 && tv.tile4) tv_value = aov_value(aovs, s, tv.aov, i, (float)samples);
  // every attempt costs the same here, so plain rounds: the sequential loop is certain to run at least
  // min(samples - count, max_total - total) more attempts (count grows by at most one per attempt)
  while (count < samples && total < max_total) {
    const int batch = min(32, min(samples - count, max_total - total));
    int pixel = -1, ix = 0, iy = 0;
    float rgbw[3] = {1.f, 1.f, 1.f};
    if (lane < batch) {
      pixel = thinlens_attempt(cam, tl, fc, P, seed_base, (uint32_t)(total + lane), rgbw, ix, iy);
      ++n_attempts;
      if (pixel >= 0) ++n_splats;
    }
    __syncwarp();
    const unsigned ok = __ballot_sync(0xffffffffu, pixel >= 0);
    if (ok) {
      if (kTile > 0 && tv.tile4) {
        const int tx = ix - tv.org_x, ty = iy - tv.org_y;
        const bool inside = pixel >= 0 && (unsigned)tx < (unsigned)kTile && (unsigned)ty < (unsigned)kTile;
        if (inside) {  // add_to_buffer of the window's AOV, in shared memory
          const int t = ty * kTile + tx;
          float *q = reinterpret_cast<float *>(tv.tile4 + t);
          atomicAdd(q + 0, (tv_value.x + w.add_energy) * weight * rgbw[0]);
          atomicAdd(q + 1, (tv_value.y + w.add_energy) * weight * rgbw[1]);
          atomicAdd(q + 2, (tv_value.z + w.add_energy) * weight * rgbw[2]);
          atomicAdd(q + 3, (tv_value.w + w.add_energy) * weight);
          atomicAdd(tv.tilew + t, weight);
          ++n_tile;
        }
        // lanes inside the window skip that AOV below; the others send it to L2
        splat_all_aovs<kHoistProbes>(fc, aovs, s, i, (float)samples, pixel, w.add_energy, depth, weight, rgbw, sample_base + i, counters, inside ? tv.aov : -1);
      } else {
        splat_all_aovs<kHoistProbes>(fc, aovs, s, i, (float)samples, pixel, w.add_energy, depth, weight, rgbw, sample_base + i, counters, -1);
      }
    }
    count += __popc(ok);
    total += batch;
  }
  n_splats = __reduce_add_sync(0xffffffffu, n_splats);
  n_attempts = __reduce_add_sync(0xffffffffu, n_attempts);
  n_tile = __reduce_add_sync(0xffffffffu, n_tile);
  if (lane == 0) {
    atomicAdd(&counters->splats, (unsigned long long)n_splats);
    atomicAdd(&counters->attempts, (unsigned long long)n_attempts);
    if (n_tile) atomicAdd(&counters->tile_splats, (unsigned long long)n_tile);
  }
}

// work_heads[4]: 0 = not decided, 1 = direct kernel, 2 = tile kernel (k_thinlens_pick; the other kernel returns at once)
constexpr unsigned kPickDirect = 1u, kPickTile = 2u;

template <bool kHoistProbes>
__global__ void __launch_bounds__(128)
k_filter_splat_thinlens(const __grid_constant__ CamConsts<float> cam, const __grid_constant__ ThinConsts tl, const __grid_constant__ FilterConsts fc,
                        const __grid_constant__ AovSet aovs, const __grid_constant__ SampleIO s, const WorkItem *__restrict__ work,
                        FilterCounters *__restrict__ counters, uint64_t sample_base) {
  if (*((volatile unsigned *)&aovs.work_heads[4]) == kPickTile) return;
  const int lane = threadIdx.x & 31;
  const unsigned n_work = *((volatile unsigned *)&aovs.work_heads[0]);
  const TileView none{nullptr, nullptr, 0, 0, -1};
  for (;;) {
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(&aovs.work_heads[1], 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= n_work) break;
    const WorkItem w = work[idx];
    thinlens_splat_item<0, kHoistProbes>(cam, tl, fc, aovs, s, w, counters, sample_base, none);
  }
}

// the pixel a source sample's disc is centred on (lens point (0, 0)); false when that ray does not reach the image
LB_DEV bool thinlens_disc_centre(const ThinConsts &tl, const FilterConsts &fc, const WorkItem &w, float &cx, float &cy) {
  uint32_t seed = 0u;
  float rgbw[3];
  ThinConsts t0 = tl;
  t0.abb_chromatic = 0.0f;  // no channel draw
  return thinlens_project(t0, fc, V3f{w.csp[0], w.csp[1], w.csp[2]}, 0.0, 0.0, seed, rgbw, cx, cy);
}

template <int kTile, int kThreads>
__global__ void __launch_bounds__(kThreads)
k_filter_splat_thinlens_tile(const __grid_constant__ CamConsts<float> cam, const __grid_constant__ ThinConsts tl, const __grid_constant__ FilterConsts fc,
                             const __grid_constant__ AovSet aovs, const __grid_constant__ SampleIO s, const WorkItem *__restrict__ work,
                             FilterCounters *__restrict__ counters, uint64_t sample_base, int tile_aov) {
  if (*((volatile unsigned *)&aovs.work_heads[4]) == kPickDirect) return;
  extern __shared__ float4 smem_tile[];
  float4 *tile4 = smem_tile;
  float *tilew = reinterpret_cast<float *>(tile4 + kTile * kTile);
  __shared__ unsigned sh_first, sh_next;
  __shared__ int sh_org[2], sh_have;
  const int lane = threadIdx.x & 31;
  const unsigned n_work = *((volatile unsigned *)&aovs.work_heads[0]);
  for (int t = threadIdx.x; t < kTile * kTile; t += kThreads) { tile4[t] = make_float4(0.f, 0.f, 0.f, 0.f); tilew[t] = 0.f; }
  for (;;) {
    __syncthreads();  // the window is clean, the previous batch's shared words are no longer read
    if (threadIdx.x == 0) {
      const unsigned first = atomicAdd(&aovs.work_heads[1], (unsigned)kTileBatch);
      sh_first = first;
      sh_next = 0u;
      sh_have = 0;
      if (first < n_work) {  // window centred on the first item's disc, kept inside the frame
        float cx, cy;
        if (thinlens_disc_centre(tl, fc, work[first], cx, cy)) {
          sh_org[0] = max(0, min((int)floorf(cx) - kTile / 2, fc.xres - kTile));
          sh_org[1] = max(0, min((int)floorf(cy) - kTile / 2, fc.yres - kTile));
          sh_have = 1;
        }
      }
    }
    __syncthreads();
    const unsigned first = sh_first;
    if (first >= n_work) break;
    const unsigned n_in = min((unsigned)kTileBatch, n_work - first);
    const TileView tv{sh_have ? tile4 : nullptr, tilew, sh_org[0], sh_org[1], tile_aov};
    for (;;) {
      unsigned k = 0;
      if (lane == 0) k = atomicAdd(&sh_next, 1u);
      k = __shfl_sync(0xffffffffu, k, 0);
      if (k >= n_in) break;
      const WorkItem w = work[first + k];
      thinlens_splat_item<kTile>(cam, tl, fc, aovs, s, w, counters, sample_base, tv);
    }
    __syncthreads();
    if (tv.tile4) {  // flush the touched pixels and clean the window
      for (int t = threadIdx.x; t < kTile * kTile; t += kThreads) {
        const float4 v = tile4[t];
        const float wv = tilew[t];
        if (wv != 0.0f || v.x != 0.0f || v.y != 0.0f || v.z != 0.0f || v.w != 0.0f) {
          const int x = tv.org_x + t % kTile, y = tv.org_y + t / kTile;
          if (x < fc.xres && y < fc.yres) {
            const unsigned pixel = (unsigned)y * (unsigned)fc.xres + (unsigned)x;
            if (wv != 0.0f) atomicAdd(aovs.weight + pixel, wv);
            atomicAdd(aovs.buffer[tile_aov] + pixel, v);
          }
          tile4[t] = make_float4(0.f, 0.f, 0.f, 0.f);
          tilew[t] = 0.f;
        }
      }
    }
  }
}

// Which kernel runs this batch: the tile kernel when the discs fit its window.  One CTA looks at up to 128 work items spread over
// the list, measures each disc's extent in pixels from the projections of four lens-rim points, and takes the largest.
__global__ void __launch_bounds__(128)
k_thinlens_pick(const __grid_constant__ ThinConsts tl, const __grid_constant__ FilterConsts fc, const __grid_constant__ AovSet aovs,
                const WorkItem *__restrict__ work, int tile, int force) {
  const unsigned n_work = *((volatile unsigned *)&aovs.work_heads[0]);
  float extent = 0.f;
  if (force == 0 && n_work > 0u) {
    const unsigned idx = (unsigned)(((unsigned long long)threadIdx.x * n_work) / 128u);
    const WorkItem w = work[idx];
    ThinConsts t0 = tl;
    t0.abb_chromatic = 0.0f;
    t0.optical_vignetting_distance = 0.0f;  // the rim points are the ones vignetting removes first
    FilterConsts f0 = fc;  // no image bounds: only the offsets matter
    f0.xres = f0.yres = 1 << 20; f0.region_min_x = f0.region_min_y = -65536;  // float pixel coordinates keep 1/128 px there
    const V3f P{w.csp[0], w.csp[1], w.csp[2]};
    float x[4], y[4], rgbw[3];
    uint32_t seed = 0u;
    const double rim[4][2] = {{1.0, 0.0}, {-1.0, 0.0}, {0.0, 1.0}, {0.0, -1.0}};
    bool okk = true;
    for (int k = 0; k < 4; ++k) okk = thinlens_project(t0, f0, P, rim[k][0], rim[k][1], seed, rgbw, x[k], y[k]) && okk;
    if (okk) extent = fmaxf(fabsf(x[0] - x[1]), fabsf(y[2] - y[3]));
    else extent = 1.0e9f;
  }
  extent = fmaxf(extent, __shfl_xor_sync(0xffffffffu, extent, 16));
  extent = fmaxf(extent, __shfl_xor_sync(0xffffffffu, extent, 8));
  extent = fmaxf(extent, __shfl_xor_sync(0xffffffffu, extent, 4));
  extent = fmaxf(extent, __shfl_xor_sync(0xffffffffu, extent, 2));
  extent = fmaxf(extent, __shfl_xor_sync(0xffffffffu, extent, 1));
  __shared__ float sh[4];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = extent;
  __syncthreads();
  if (threadIdx.x == 0) {
    extent = fmaxf(fmaxf(sh[0], sh[1]), fmaxf(sh[2], sh[3]));
    // a disc no wider than the window (a few pixels of slack for the drift inside a batch) lands in it almost entirely
    unsigned pick = (extent + 8.0f <= (float)tile) ? kPickTile : kPickDirect;
    if (force > 0) pick = force == 2 ? kPickTile : kPickDirect;
    aovs.work_heads[4] = pick;
  }
}

cudaError_t launch_create_rays_thinlens(const CamConsts<float> &cam, const ThinConsts &tl, const RayIO &io, size_t n, uint64_t ray_id_base,
                                        cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  k_create_rays_thinlens<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cam, tl, io, n, ray_id_base);
  return cudaGetLastError();
}

cudaError_t launch_filter_splat_thinlens(const CamConsts<float> &cam, const ThinConsts &tl, const FilterConsts &fc, const AovSet &aovs,
                                         const SampleIO &s, const WorkItem *work, FilterCounters *counters, uint64_t sample_base, int num_sms,
                                         cudaStream_t stream) {
  // LB_SPLAT_TILE: unset / 0 = the direct kernel (default: measured faster in every regime, see the kernel comment), 1 = the
  // tile kernel, "auto" = picked on the device per batch from the discs' extent
  const int mode = [] {
    const char *e = getenv("LB_SPLAT_TILE");
    return !e || !e[0] || e[0] == '0' ? 0 : (e[0] == 'a' ? -1 : 1);
  }();
  constexpr int kTile = 64, kThreads = 256;
  constexpr size_t smem = (size_t)kTile * kTile * 20;
  int tile_aov = -1;  // the window holds the RGBA AOV (the one filter_weight_buffer belongs to, lentil.h:828)
  for (int a = 0; a < fc.n_aov && tile_aov < 0; ++a)
    if (aovs.filter[a] == 0 && aovs.role[a] == 1) tile_aov = a;
  const bool tile_possible = tile_aov >= 0 && fc.xres >= kTile && fc.yres >= kTile;
  const int force = !tile_possible || mode == 0 ? 1 : (mode == 1 ? 2 : 0);
  if (force != 1) k_thinlens_pick<<<1, 128, 0, stream>>>(tl, fc, aovs, work, kTile, force);  // (work_heads[4] == 0 also means direct)
  if (force != 2) {
    if (aovs.crypto_first >= 0) k_filter_splat_thinlens<true><<<num_sms * 8, 128, 0, stream>>>(cam, tl, fc, aovs, s, work, counters, sample_base);
    else k_filter_splat_thinlens<false><<<num_sms * 8, 128, 0, stream>>>(cam, tl, fc, aovs, s, work, counters, sample_base);
  }
  if (force != 1) {
    const cudaError_t attr = cudaFuncSetAttribute(k_filter_splat_thinlens_tile<kTile, kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (attr != cudaSuccess) return attr;
    k_filter_splat_thinlens_tile<kTile, kThreads><<<num_sms * 2, kThreads, smem, stream>>>(cam, tl, fc, aovs, s, work, counters, sample_base, tile_aov);
  }
  return cudaGetLastError();
}

}  // namespace lb
