// lens_device.cuh — device-side primitives of the lentil hot paths, templated on the arithmetic
// type (float for the per-ray kernels, double for the setup solvers).
//
// Reference code each function stands for is cited inline (/root/reference/src/...).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "lens_table.h"

namespace lb {

#define LB_DEV __device__ __forceinline__

// ---- RNG: bit-exact integer code of global.h:32-57 ---------------------------------------------
LB_DEV uint32_t tea8(uint32_t v0, uint32_t v1) {  // tea<8>, global.h:32-46
  uint32_t s0 = 0;
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    s0 += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xA341316Cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4u);
    v1 += ((v0 << 4) + 0xAD90777Du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761Eu);
  }
  return v0;
}
// 32-bit seed word of a 64-bit global ray index (retry RNG of the forward path): identical to the index below 2^32, the
// high word is mixed in above it (8 GPUs x 5.3e8 rays per frame approach 2^32)
LB_DEV uint32_t ray_seed_word(uint64_t id) { return (uint32_t)id ^ ((uint32_t)(id >> 32) * 0x9E3779B9u); }
LB_DEV float lcg_rng(uint32_t &previous) {  // rng, global.h:51-57
  previous = previous * 1664525u + 1013904223u;
  return __uint2float_rn(previous & 0x00FFFFFFu) * (1.0f / 16777216.0f);  // exact: 24-bit / 2^24
}

// ---- lens.h:17-37, float on purpose (the reference evaluates these in float) ---------------------
LB_DEV float fast_sin(float x) {
  const float PI = 3.14159265358979323846f;
  x = fmodf(x + PI, PI * 2) - PI;
  const float B = 4.0f / PI;
  const float C = -4.0f / (PI * PI);
  float y = __fadd_rn(__fmul_rn(B, x), __fmul_rn(__fmul_rn(C, x), fabsf(x)));
  const float P = 0.225f;
  return __fadd_rn(__fmul_rn(P, __fsub_rn(__fmul_rn(y, fabsf(y)), y)), y);
}
LB_DEV float fast_cos(float x) {
  const float PI = 3.14159265358979323846f;
  x = (float)((double)x + (double)PI * 0.5);  // `x += AI_PI * 0.5` promotes to double (lens.h:29)
  return fast_sin(x);                          // identical body from here on (lens.h:31-36)
}

template <typename T> LB_DEV T t_sqrt(T x);
template <> LB_DEV float t_sqrt<float>(float x) { return sqrtf(x); }
template <> LB_DEV double t_sqrt<double>(double x) { return sqrt(x); }
// 1/sqrt(x) of the tangent-frame and view-vector normalisations in sphereToCs / csToSphere.  double: sqrt then divide,
// as the host code.  float: rsqrtf (MUFU.RSQ, 2 ulp) -- the IEEE sqrt + IEEE divide pair is ~18 instructions with a
// slow-path branch each, three times per lt_sample_aperture iteration: +4.9 % splats/s, +1 % rays/s, identical splat
// counts, all parity bounds unchanged (2 ulp = 2.4e-7 relative against the 1e-4 bound).  -DLB_IEEE_RSQRT restores it.
// The divisions follow the same rule: another +8.2 % splats/s and +3.7 % rays/s (profiles/r01_k2_knobs.txt).
template <typename T> LB_DEV T t_rsqrt(T x) { return T(1) / t_sqrt(x); }
// a / b and 1 / x of the same code and of the 2x2 Newton inverses.  float: MUFU.RCP based (__fdividef, 2 ulp) unless
// -DLB_IEEE_RSQRT; -DLB_IEEE_DIV keeps only the divisions IEEE.
template <typename T> LB_DEV T t_div(T a, T b) { return a / b; }
template <typename T> LB_DEV T t_rcp(T x) { return T(1) / x; }
#ifndef LB_IEEE_RSQRT
template <> LB_DEV float t_rsqrt<float>(float x) { return rsqrtf(x); }
#ifndef LB_IEEE_DIV
template <> LB_DEV float t_div<float>(float a, float b) { return __fdividef(a, b); }
template <> LB_DEV float t_rcp<float>(float x) { return __fdividef(1.0f, x); }
#endif
#endif
template <typename T> LB_DEV T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> LB_DEV T t_max(T a, T b) { return a > b ? a : b; }

// concentric_disk_sample(..., fast_trigo = true), lens.h:309-333
template <typename T>
LB_DEV void concentric_disk_sample(T ox, T oy, T &ux, T &uy) {
  T phi, r;
  const T a = T(2.0) * ox - T(1.0);
  const T b = T(2.0) * oy - T(1.0);
  if ((a * a) > (b * b)) {
    r = a;
    phi = T(0.78539816339) * (b / a);
  } else {
    r = b;
    phi = T(3.14159265358979323846 / 2.0) - T(0.78539816339) * (a / b);
  }
  ux = r * T(fast_cos((float)phi));
  uy = r * T(fast_sin((float)phi));
}

// lens_sample_triangular_aperture, lentil.h:964-982 (angles are computed in float there)
template <typename T>
LB_DEV void sample_triangular_aperture(T &x, T &y, T r1, T r2, T radius, int blades) {
  const int tri = (int)(r1 * blades);
  r1 = r1 * blades - tri;
  const T a = t_sqrt(r1);
  const T b = (T(1.0) - r2) * a;
  const T c = r2 * a;
  const float PI = 3.14159265358979323846f;
  const float phi1 = 2.0f * PI / blades * (tri + 1);
  const float phi2 = 2.0f * PI / blades * tri;
  T s1, c1, s2, c2;
  if (sizeof(T) == 8) {
    double sd, cd;
    sincos((double)phi1, &sd, &cd); s1 = T(sd); c1 = T(cd);
    sincos((double)phi2, &sd, &cd); s2 = T(sd); c2 = T(cd);
  } else {
    float sf, cf;
    sincosf(phi1, &sf, &cf); s1 = T(sf); c1 = T(cf);
    sincosf(phi2, &sf, &cf); s2 = T(sf); c2 = T(cf);
  }
  x = radius * (b * c1 + c * c2);
  y = radius * (b * s1 + c * s2);
}

// imageData::bokehSample, imagebokeh.h:341-412: two upper_bound searches in the sorted CDF tables
LB_DEV int upper_bound_f(const float *__restrict__ a, int n, float v) {
  int lo = 0, len = n;
  while (len > 0) {
    const int half = len >> 1;
    if (!(v < __ldg(a + lo + half))) { lo += half + 1; len -= half + 1; } else len = half;
  }
  return lo;
}
// (r02, measured and withdrawn: the same upper_bound as two rounds of 16 independent loads over host-built coarse tables.  The
// thin-lens splat kernel stalls on the searches' dependent loads, but it is the L1's wavefront rate that bounds it -- every lane
// reads another row of the column table, 32 wavefronts per load instruction -- and 64 loads per attempt instead of 16 took the
// kernel from 3.4 ms to 8.8 ms on the C3 frame.  profiles/r02_thinlens_splat_ncu.txt)
// The same upper_bound through a guide table (host: BokehTables::build_guide): bucket k = floor(v * K) holds the window
// {lo, hi} = {upper_bound(a, k/K), upper_bound(a, (k+1)/K)}, which contains upper_bound(a, v) because a is sorted; the
// window holds about n/K entries (one, for the 250-pixel kernels), so the search is one guide load plus ~one table load
// instead of log2(n) dependent loads on 32 different cache lines per warp.  Bucket 0 starts at 0 and the last bucket ends
// at n, so v outside [0, 1) still gets the full search's answer; a NaN takes the full search.
LB_DEV int upper_bound_guided(const float *__restrict__ a, const uint32_t *__restrict__ guide, int n, float v) {
  if (!(v == v)) return upper_bound_f(a, n, v);
  int k = (int)(v * (float)kBokehGuide);
  k = k < 0 ? 0 : (k > kBokehGuide - 1 ? kBokehGuide - 1 : k);
  const uint32_t g = __ldg(guide + k);
  int lo = (int)(g & 0xFFFFu), len = (int)(g >> 16) - lo;
  while (len > 0) {
    const int half = len >> 1;
    if (!(v < __ldg(a + lo + half))) { lo += half + 1; len -= half + 1; } else len = half;
  }
  return lo;
}
template <typename T, typename C>
LB_DEV void bokeh_sample(const C &cam, float randomNumberRow, float randomNumberColumn, T &lx, T &ly) {
  const int n = cam.bokeh_n;
  const bool guided = cam.guide_row != nullptr;
  int r = guided ? upper_bound_guided(cam.cdf_row, cam.guide_row, n, randomNumberRow) : upper_bound_f(cam.cdf_row, n, randomNumberRow);
  if (r >= n) r = n - 1;
  const int actualPixelRow = __ldg(cam.row_idx + r);
  const int recalulatedPixelRow = actualPixelRow - ((n - 1) / 2);
  const int startPixel = actualPixelRow * n;
  int c = startPixel + (guided ? upper_bound_guided(cam.cdf_col + startPixel, cam.guide_col + (size_t)actualPixelRow * kBokehGuide, n, randomNumberColumn)
                               : upper_bound_f(cam.cdf_col + startPixel, n, randomNumberColumn));
  if (c >= startPixel + n) c = startPixel + n - 1;
  const int actualPixelColumn = __ldg(cam.col_idx + c);
  const int relativePixelColumn = actualPixelColumn - startPixel;
  const int recalulatedPixelColumn = relativePixelColumn - ((n - 1) / 2);
  const float flippedRow = (float)recalulatedPixelColumn;
  const float flippedColumn = recalulatedPixelRow * -1.0f;
  lx = T(flippedRow / (float)n) * T(2.0);
  ly = T(flippedColumn / (float)n) * T(2.0);
}

// ---- lens.h:99-221 -------------------------------------------------------------------------------
// sphereToCs with center == -R (every hot-path call site: lentil.h:389, lt_sample_aperture body).
// pos.z = normal.z*R - R is evaluated as -min(r^2,R^2)/(R*(nz+1)): same value, no cancellation in FP32.
template <typename T>
LB_DEV void sphere_to_cs(T px, T py, T dx, T dy, T R, T pos[3], T dir[3]) {
  const T r2 = px * px + py * py;
  const T nx = t_div(px, R), ny = t_div(py, R);
  const T nz = t_div(t_sqrt(t_max(T(0), R * R - r2)), t_abs(R));
  const T tz = t_sqrt(t_max(T(0), T(1) - dx * dx - dy * dy));
  // ex = normalise(nz, 0, -nx); ey = n x ex
  const T il = t_rsqrt(nz * nz + nx * nx);
  const T ex0 = nz * il, ex2 = -nx * il;
  const T ey0 = ny * ex2, ey1 = nz * ex0 - nx * ex2, ey2 = -ny * ex0;
  dir[0] = dx * ex0 + dy * ey0 + tz * nx;
  dir[1] = dy * ey1 + tz * ny;
  dir[2] = dx * ex2 + dy * ey2 + tz * nz;
  pos[0] = px;
  pos[1] = py;
  if (sizeof(T) == 8) pos[2] = nz * R + (-R);  // double path mirrors the reference expression
  else pos[2] = t_div(-(r2 < R * R ? r2 : R * R), R * (nz + T(1)));  // R*(nz-1) = R*(nz^2-1)/(nz+1)
}
// general-centre variant (setup only: inner pupil, lentil.h:1423)
template <typename T>
LB_DEV void sphere_to_cs_center(T px, T py, T dx, T dy, T center, T R, T pos[3], T dir[3]) {
  sphere_to_cs(px, py, dx, dy, R, pos, dir);
  const T nz = t_sqrt(t_max(T(0), R * R - px * px - py * py)) / t_abs(R);
  pos[2] = nz * R + center;
}
// csToSphere, lens.h:127-153
template <typename T>
LB_DEV void cs_to_sphere(const T pos[3], const T dir_in[3], T center, T R, T &odx, T &ody) {
  const T nx = t_div(pos[0], R), ny = t_div(pos[1], R), nz = t_abs(t_div(pos[2] - center, R));
  const T dl = t_rsqrt(dir_in[0] * dir_in[0] + dir_in[1] * dir_in[1] + dir_in[2] * dir_in[2]);
  const T d0 = dir_in[0] * dl, d1 = dir_in[1] * dl, d2 = dir_in[2] * dl;
  const T il = t_rsqrt(nz * nz + nx * nx);
  const T ex0 = nz * il, ex2 = -nx * il;
  const T ey0 = ny * ex2, ey1 = nz * ex0 - nx * ex2, ey2 = -ny * ex0;
  odx = d0 * ex0 + d2 * ex2;
  ody = d0 * ey0 + d1 * ey1 + d2 * ey2;
}
// cylinderToCs, lens.h:188-221
template <typename T>
LB_DEV void cylinder_to_cs(T px, T py, T dx, T dy, T center, T R, bool cyl_y, T pos[3], T dir[3]) {
  T nx = 0, ny = 0, nz;
  if (cyl_y) { nx = px / R; nz = t_sqrt(t_max(T(0), R * R - px * px)) / t_abs(R); }
  else       { ny = py / R; nz = t_sqrt(t_max(T(0), R * R - py * py)) / t_abs(R); }
  const T tz = t_sqrt(t_max(T(0), T(1) - dx * dx - dy * dy));
  T il = T(1) / t_sqrt(nz * nz + nx * nx);
  const T ex0 = nz * il, ex2 = -nx * il;
  T ey0 = ny * ex2, ey1 = nz * ex0 - nx * ex2, ey2 = -ny * ex0;
  il = T(1) / t_sqrt(ey0 * ey0 + ey1 * ey1 + ey2 * ey2);
  ey0 *= il; ey1 *= il; ey2 *= il;
  dir[0] = dx * ex0 + dy * ey0 + tz * nx;
  dir[1] = dy * ey1 + tz * ny;
  dir[2] = dx * ex2 + dy * ey2 + tz * nz;
  pos[0] = px; pos[1] = py; pos[2] = nz * R + center;
}
// csToCylinder, lens.h:156-185 (ex is NOT normalised there)
template <typename T>
LB_DEV void cs_to_cylinder(const T pos[3], const T dir_in[3], T center, T R, bool cyl_y, T &odx, T &ody) {
  T nx = 0, ny = 0;
  const T nz = t_abs((pos[2] - center) / R);
  if (cyl_y) nx = pos[0] / R; else ny = pos[1] / R;
  const T dl = T(1) / t_sqrt(dir_in[0] * dir_in[0] + dir_in[1] * dir_in[1] + dir_in[2] * dir_in[2]);
  const T d0 = dir_in[0] * dl, d1 = dir_in[1] * dl, d2 = dir_in[2] * dl;
  const T ex0 = nz, ex2 = -nx;
  T ey0 = ny * ex2, ey1 = nz * ex0 - nx * ex2, ey2 = -ny * ex0;
  const T il = T(1) / t_sqrt(ey0 * ey0 + ey1 * ey1 + ey2 * ey2);
  ey0 *= il; ey1 *= il; ey2 *= il;
  odx = d0 * ex0 + d2 * ex2;
  ody = d0 * ey0 + d1 * ey1 + d2 * ey2;
}
LB_DEV float sqrt_approx(float x);
// sphereToCs (lens.h:99-125) for the float kernels with the per-camera reciprocals of the outer pupil sphere: MUFU.SQRT / MUFU.RSQ /
// MUFU.RCP (1-2 ulp) instead of four IEEE divisions and three IEEE square roots with their fix-up sequences and slow-path branches
template <typename C>
LB_DEV void sphere_to_cs_fast(const C &cam, const float out[4], float pos[3], float dir[3]) {
  const float px = out[0], py = out[1], dx = out[2], dy = out[3];
  const float r2 = px * px + py * py;
  const float nx = px * cam.inv_outer_R, ny = py * cam.inv_outer_R;
  const float nz = sqrt_approx(fmaxf(0.0f, cam.outer_R2 - r2)) * cam.abs_inv_outer_R;
  const float tz = sqrt_approx(fmaxf(0.0f, 1.0f - dx * dx - dy * dy));
  const float il = rsqrtf(nz * nz + nx * nx);
  const float ex0 = nz * il, ex2 = -nx * il;
  const float ey0 = ny * ex2, ey1 = nz * ex0 - nx * ex2, ey2 = -ny * ex0;
  dir[0] = dx * ex0 + dy * ey0 + tz * nx;
  dir[1] = dy * ey1 + tz * ny;
  dir[2] = dx * ex2 + dy * ey2 + tz * nz;
  pos[0] = px;
  pos[1] = py;
  pos[2] = -fminf(r2, cam.outer_R2) * t_rcp(fmaf(cam.outer_R, nz, cam.outer_R));  // R*nz - R without cancellation
}
template <typename T, typename C>
LB_DEV void outer_to_cs(const C &cam, const T out[4], T pos[3], T dir[3]) {  // lentil.h:387-389
#ifndef LB_GENERIC_LT_TAIL
  if constexpr (sizeof(T) == 4) {
    if (cam.outer_geom == 0) {
      sphere_to_cs_fast(cam, out, pos, dir);
      return;
    }
  }
#endif
  if (cam.outer_geom == 0) sphere_to_cs(out[0], out[1], out[2], out[3], cam.outer_R, pos, dir);
  else cylinder_to_cs(out[0], out[1], out[2], out[3], -cam.outer_R, cam.outer_R, cam.outer_geom == 1, pos, dir);
}
template <typename T, typename C>
LB_DEV void cs_to_outer(const C &cam, const T pos[3], const T dir[3], T &odx, T &ody) {
  if (cam.outer_geom == 0) cs_to_sphere(pos, dir, -cam.outer_R, cam.outer_R, odx, ody);
  else cs_to_cylinder(pos, dir, -cam.outer_R, cam.outer_R, cam.outer_geom == 1, odx, ody);
}

// ---- generic (table-driven) polynomial evaluation --------------------------------------------------
// lens_ipow, lens.h:226-233, same multiplication tree; `e` is warp-uniform (it comes from the table)
template <typename T>
LB_DEV T ipow(T x, int e) {
  if (e == 0) return T(1);
  int odd = 0, depth = 0;
  while (e > 2) { odd |= (e & 1) << depth; ++depth; e >>= 1; }
  T p = (e == 1) ? x : x * x;
  while (depth--) p = ((odd >> depth) & 1) ? x * p * p : p * p;
  return p;
}
// flat monomial sum `+ c*x*lens_ipow(y,2)*... + ...`, left to right, as the generated headers do
template <typename T>
LB_DEV T poly_eval(const LensTable &L, int p, const T v[5]) {
  T acc = T(0);
  const int o = L.off[p], n = L.cnt[p];
  for (int i = 0; i < n; ++i) {
    const Term t = L.t[o + i];
    T m = T(t.c);
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int ek = (t.e >> (4 * k)) & 15;
      if (ek == 1) m *= v[k];
      else if (ek > 1) m *= ipow(v[k], ek);
    }
    acc += m;
  }
  return acc;
}

// Evaluator policy over the table.  An unrolled per-lens policy (gen/) offers the same members.
template <typename T>
struct TableEval {
  const LensTable &L;
  LB_DEV explicit TableEval(const LensTable &l) : L(l) {}
  // aperture position + its Jacobian wrt (dx,dy): pred_x, pred_y, dx1_domega0
  LB_DEV void ap_jac(const T b[5], T ap[2], T J[4]) const {
    ap[0] = poly_eval(L, P_AP_X, b);
    ap[1] = poly_eval(L, P_AP_Y, b);
    J[0] = poly_eval(L, P_DAPX_DDX, b);
    J[1] = poly_eval(L, P_DAPX_DDY, b);
    J[2] = poly_eval(L, P_DAPY_DDX, b);
    J[3] = poly_eval(L, P_DAPY_DDY, b);
  }
  LB_DEV void out4(const T b[5], T out[4]) const {
    out[0] = poly_eval(L, P_OUT_X, b);
    out[1] = poly_eval(L, P_OUT_Y, b);
    out[2] = poly_eval(L, P_OUT_DX, b);
    out[3] = poly_eval(L, P_OUT_DY, b);
  }
  LB_DEV T transmittance(const T b[5]) const { return poly_eval(L, P_OUT_T, b); }
  LB_DEV void out5(const T b[5], T out[4], T &Tr) const {  // lens_evaluate: the 5 polynomials of pt_evaluate.h
    out4(b, out);
    Tr = transmittance(b);
  }
  LB_DEV void out_jac(const T b[5], T K[4]) const {  // domega2_dx0
    K[0] = poly_eval(L, P_DODX_DX, b);
    K[1] = poly_eval(L, P_DODX_DY, b);
    K[2] = poly_eval(L, P_DODY_DX, b);
    K[3] = poly_eval(L, P_DODY_DY, b);
  }
  // everything one lt_sample_aperture iteration needs at one point
  LB_DEV void lt_all(const T b[5], T ap[2], T J[4], T out[4], T K[4]) const {
    ap_jac(b, ap, J);
    out4(b, out);
    out_jac(b, K);
  }
};

// ---- Newton solvers (generated bodies of polynomial-optics) ------------------
// pt_sample_aperture: solve (dx,dy) so the sensor->aperture polynomial hits (ax, ay).
// pred_dx/pred_dy of the generated body are dead in the hot path (lentil.h:344 only reads sensor).
template <typename T, typename E>
LB_DEV int pt_sample_aperture(const E &ev, T x, T y, T &dx, T &dy, T lambda, T ax, T ay, T dist) {
  T sqr_err = T(3.402823466e+38);  // FLT_MAX
  int k = 0;
  for (; k < 5 && sqr_err > T(1e-4); k++) {
    const T b[5] = {x + dist * dx, y + dist * dy, dx, dy, lambda};
    T ap[2], J[4];
    ev.ap_jac(b, ap, J);
    const T invdet = t_rcp(J[0] * J[3] - J[1] * J[2]);
    const T e0 = ax - ap[0], e1 = ay - ap[1];
    dx += (J[3] * invdet) * e0;
    dx += (-J[1] * invdet) * e1;
    dy += (-J[2] * invdet) * e0;
    dy += (J[0] * invdet) * e1;
    sqr_err = e0 * e0 + e1 * e1;
  }
  return k;
}

// lt_sample_aperture: solve sensor (x,y,dx,dy) so that the ray passes aperture point (ax,ay) and
// leaves the outer pupil towards `scene`.  The generated loop body is exposed one iteration at a time
// (lt_init / lt_continue / lt_iterate / lt_finish) so that the splat kernel can advance all lanes of a
// warp by one iteration per trip and refill finished lanes; lt_sample_aperture is the plain loop.
template <typename T>
struct LtState {
  T x, y, dx, dy;
  T sqr_err, sqr_ap_err;
  T out[4];
  int error, k;
};
template <typename T>
LB_DEV void lt_init(LtState<T> &s) {
  s.x = s.y = s.dx = s.dy = T(0);
  s.sqr_err = s.sqr_ap_err = T(1e30);
  s.out[0] = s.out[1] = s.out[2] = s.out[3] = T(0);
  s.error = 0;
  s.k = 0;
}
template <typename T>
LB_DEV bool lt_continue(const LtState<T> &s) {  // for(k<100 && (sqr_err>eps || sqr_ap_err>eps) && error==0)
  const T eps = T(1e-8);
  return s.k < 100 && (s.sqr_err > eps || s.sqr_ap_err > eps) && s.error == 0;
}
// everything of one loop trip after the 14 polynomials (s.out already holds the outer-pupil point)
template <typename T, typename C>
LB_DEV void lt_iterate_tail(const C &cam, const T scene[3], T ax, T ay, const T ap[2], const T J[4], const T K[4], LtState<T> &s) {
  const T prev_sqr_err = s.sqr_err, prev_sqr_ap_err = s.sqr_ap_err;
  const T da0 = ax - ap[0], da1 = ay - ap[1];
  s.sqr_ap_err = da0 * da0 + da1 * da1;
  const T invdetap = t_rcp(J[0] * J[3] - J[1] * J[2]);
  s.dx += (J[3] * invdetap) * da0;
  s.dx += (-J[1] * invdetap) * da1;
  s.dy += (-J[2] * invdetap) * da0;
  s.dy += (J[0] * invdetap) * da1;
  T pos[3], dir[3];
  outer_to_cs(cam, s.out, pos, dir);
  const T view[3] = {scene[0] - pos[0], scene[1] - pos[1], scene[2] - pos[2]};
  T ndx, ndy;
  cs_to_outer(cam, pos, view, ndx, ndy);
  const T do0 = ndx - s.out[2], do1 = ndy - s.out[3];
  s.sqr_err = do0 * do0 + do1 * do1;
  const T invdet = t_rcp(K[0] * K[3] - K[1] * K[2]);
  s.x += T(0.72) * (K[3] * invdet) * do0;
  s.x += T(0.72) * (-K[1] * invdet) * do1;
  s.y += T(0.72) * (-K[2] * invdet) * do0;
  s.y += T(0.72) * (K[0] * invdet) * do1;
  int error = s.error;
  if (s.sqr_err > prev_sqr_err) error |= 1;
  if (s.sqr_ap_err > prev_sqr_ap_err) error |= 2;
  if (s.out[0] != s.out[0]) error |= 4;
  if (s.out[0] * s.out[0] + s.out[1] * s.out[1] > cam.outer_pupil_r2) error |= 16;
  if (s.k < 10) error = 0;  // "error reset (k<10)", tests/aperture_sampling_debug/writout.txt:40
  s.error = error;
  s.k += 1;
}
// sqrt for non-negative finite arguments on the special-function unit (MUFU.SQRT, 1-2 ulp); sqrtf is a MUFU.RSQ plus a
// fix-up sequence with a slow-path branch
LB_DEV float sqrt_approx(float x) {  // (declared above for sphere_to_cs_fast)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// The same loop trip for the float kernels and a SPHERICAL outer pupil (every lens of the pack; warp-uniform test), with
// sphereToCs (lens.h:99-125) and csToSphere (lens.h:127-153) fused: both build the same normal and tangent frame at the
// outer-pupil point, the view vector need not be normalised before it is projected (one rsqrt of |view|^2 * |ex|^2
// scales both projections), and each 2x2 Newton step applies the inverse determinant to the residual once instead of
// to the four matrix entries.  Same mathematics as lt_iterate_tail, about 55 FP32 operations + 5 MUFU instead of ~130
// (ncu r02: with the mirror-packed polynomials the tail had become 45 % of the instructions of a trip).
template <typename C>
LB_DEV void lt_iterate_tail_sphere(const C &cam, const float scene[3], float ax, float ay, const float ap[2], const float J[4], const float K[4],
                                   LtState<float> &s) {
  const float prev_sqr_err = s.sqr_err, prev_sqr_ap_err = s.sqr_ap_err;
  const float da0 = ax - ap[0], da1 = ay - ap[1];
  s.sqr_ap_err = da0 * da0 + da1 * da1;
  const float ga = t_rcp(J[0] * J[3] - J[1] * J[2]);
  const float a0 = da0 * ga, a1 = da1 * ga;
  s.dx = fmaf(-J[1], a1, fmaf(J[3], a0, s.dx));
  s.dy = fmaf(J[0], a1, fmaf(-J[2], a0, s.dy));
  // outer-pupil point -> camera space (sphere centred at -R)
  const float px = s.out[0], py = s.out[1];
  const float r2 = px * px + py * py;
  const float nx = px * cam.inv_outer_R, ny = py * cam.inv_outer_R;
  const float nz = sqrt_approx(fmaxf(0.0f, cam.outer_R2 - r2)) * cam.abs_inv_outer_R;
  const float pz = -fminf(r2, cam.outer_R2) * t_rcp(fmaf(cam.outer_R, nz, cam.outer_R));  // R*nz - R without cancellation
  const float v0 = scene[0] - px, v1 = scene[1] - py, v2 = scene[2] - pz;
  // direction towards the scene point in the tangent frame ex = (nz, 0, -nx)/|.|, ey = n x ex
  const float f2 = nz * nz + nx * nx;
  const float w = rsqrtf((v0 * v0 + v1 * v1 + v2 * v2) * f2);
  const float ndx = (v0 * nz - v2 * nx) * w;
  const float ndy = (f2 * v1 - ny * (nx * v0 + nz * v2)) * w;
  const float do0 = ndx - s.out[2], do1 = ndy - s.out[3];
  s.sqr_err = do0 * do0 + do1 * do1;
  const float go = 0.72f * t_rcp(K[0] * K[3] - K[1] * K[2]);
  const float b0 = do0 * go, b1 = do1 * go;
  s.x = fmaf(-K[1], b1, fmaf(K[3], b0, s.x));
  s.y = fmaf(K[0], b1, fmaf(-K[2], b0, s.y));
  int error = s.error;
  if (s.sqr_err > prev_sqr_err) error |= 1;
  if (s.sqr_ap_err > prev_sqr_ap_err) error |= 2;
  if (px != px) error |= 4;
  if (r2 > cam.outer_pupil_r2) error |= 16;
  if (s.k < 10) error = 0;  // "error reset (k<10)", tests/aperture_sampling_debug/writout.txt:40
  s.error = error;
  s.k += 1;
}
template <typename T, typename E, typename C>
LB_DEV void lt_iterate(const E &ev, const C &cam, const T scene[3], T ax, T ay, T lambda, LtState<T> &s) {
  const T b[5] = {s.x, s.y, s.dx, s.dy, lambda};
  T ap[2], J[4], K[4];
  ev.lt_all(b, ap, J, s.out, K);
#ifndef LB_GENERIC_LT_TAIL
  if constexpr (sizeof(T) == 4) {
    if (cam.outer_geom == 0) {
      lt_iterate_tail_sphere(cam, scene, ax, ay, ap, J, K, s);
      return;
    }
  }
#endif
  lt_iterate_tail(cam, scene, ax, ay, ap, J, K, s);
}
// half of a packed pair (two-rays-per-thread forward kernel, camera_kernels.cuh)
LB_DEV float lo_hi(const float2 &v, int h) { return h ? v.y : v.x; }
// element h of a register-resident pair, for loops over a warp-uniform half index that must stay rolled (one copy of
// the loop body in the instruction cache): selects instead of indexed addressing
template <typename T>
LB_DEV T pick(const T (&a)[2], int h) { return h ? a[1] : a[0]; }
template <typename T>
LB_DEV void put(T (&a)[2], int h, T v) { if (h) a[1] = v; else a[0] = v; }
// after the loop: final pupil test and transmittance; returns max(0, out[4])
template <typename T, typename E, typename C>
LB_DEV T lt_finish(const E &ev, const C &cam, T lambda, const LtState<T> &s) {
  int error = s.error;
  if (s.out[0] * s.out[0] + s.out[1] * s.out[1] > cam.outer_pupil_r2) error |= 16;
  if (error != 0) return T(0);
  const T b[5] = {s.x, s.y, s.dx, s.dy, lambda};
  return t_max(T(0), ev.transmittance(b));
}
template <typename T, typename E, typename C>
LB_DEV T lt_sample_aperture(const E &ev, const C &cam, const T scene[3], T ax, T ay, T lambda, T sensor[4], T out[4], int *its) {
  LtState<T> s;
  lt_init(s);
  while (lt_continue(s)) lt_iterate(ev, cam, scene, ax, ay, lambda, s);
  if (its) *its = s.k;
  sensor[0] = s.x; sensor[1] = s.y; sensor[2] = s.dx; sensor[3] = s.dy;
  out[0] = s.out[0]; out[1] = s.out[1]; out[2] = s.out[2]; out[3] = s.out[3];
  return lt_finish(ev, cam, lambda, s);
}

}  // namespace lb
