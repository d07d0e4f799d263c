// lens_table.h — lens description handed to the kernels BY VALUE as a __grid_constant__ kernel
// parameter: it lands in the constant bank (warp-uniform LDC reads, no global state, any number
// of cameras with different lenses per process).
//
// Replaces the textual `switch(lensModel){ #include "...pt_evaluate.h" }` of
// /root/reference/src/lentil.h:1261-1263,1277-1279,1307-1309: lens selection is a host-side table
// pick (generic kernels) or a kernel pick (unrolled kernels), never a per-ray branch.
#pragma once
#include <stdint.h>

namespace lb {

// polynomial slots: the 9 fitted polynomials + the 8 derivative polynomials the Newton solvers need
enum PolyId {
  P_OUT_X = 0, P_OUT_Y, P_OUT_DX, P_OUT_DY, P_OUT_T, P_AP_X, P_AP_Y, P_AP_DX, P_AP_DY,
  P_DAPX_DDX, P_DAPX_DDY, P_DAPY_DDX, P_DAPY_DDY,      // dx1_domega0 (pt_/lt_sample_aperture)
  P_DODX_DX, P_DODX_DY, P_DODY_DX, P_DODY_DY,          // domega2_dx0 (lt_sample_aperture)
  P_COUNT
};

constexpr int kMaxTerms = 1152;

struct Term {
  float c;
  uint32_t e;  // 4-bit exponents: x | y<<4 | dx<<8 | dy<<12 | lambda<<16
};

struct LensTable {
  uint16_t off[P_COUNT];
  uint16_t cnt[P_COUNT];
  Term t[kMaxTerms];
};

constexpr int kBokehGuide = 256;  // buckets of [0, 1) per guided CDF search (a power of two: k/K is exact in float)

// Per-camera scalars every ray needs (struct Camera, lentil.h:106-160 + setup results).
template <typename T>
struct CamConsts {
  T sensor_half;         // sensor_width * 0.5
  T lambda;              // micrometres
  T aperture_radius;
  T sensor_shift;
  T outer_R;             // lens_outer_pupil_curvature_radius
  T outer_pupil_r2;      // lens_outer_pupil_radius^2
  T inner_pupil_r2;      // lens_inner_pupil_radius^2
  T inner_R;             // lens_inner_pupil_curvature_radius
  T bfl;                 // lens_back_focal_length
  float unit_scale;      // -1, -0.1, -0.01, -0.001 (lentil.h:395-416)
  float exposure;
  float deriv_baseline;  // finite-difference baseline multiplier of the differential rays (1 = lentil_camera.cpp:84)
  int32_t enable_dof;
  int32_t vignetting_retries;
  int32_t blades;
  int32_t bokeh_n;        // 0 = no bokeh image, else image side length
  int32_t outer_geom;     // 0 spherical, 1 cyl-y, 2 cyl-x
  int32_t inner_geom;
  const float *cdf_row;   // bokeh tables (imagebokeh.h:30-39), device pointers
  const int32_t *row_idx;
  const float *cdf_col;
  const int32_t *col_idx;
  const uint32_t *guide_row;  // kBokehGuide {lo, hi} search windows of cdf_row, null = full binary search
  const uint32_t *guide_col;  // bokeh_n x kBokehGuide windows of the rows of cdf_col
  T inv_outer_R, abs_inv_outer_R, outer_R2;  // 1/R, 1/|R|, R^2 of the outer pupil sphere (lt_iterate_tail_sphere)
  double lambda_exact;    // wavelength in double: the host folds it into the polynomial coefficients (gen/, EvalFA / EvalFB)
};

}  // namespace lb
