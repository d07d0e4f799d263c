// lentil_host.cu — host side of the C ABI (include/lentil_b200.h).
//
// Owns what `struct Camera` (/root/reference/src/lentil.h:92-1671) owns in the reference — parameters,
// lens constants, solver results, bokeh CDF, AOV framebuffers — but device-resident, and drives the
// kernels.  No CPU compute path exists here: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lentil_b200.h"
#include "camera_kernels.cuh"
#include "gen/lens_pack_data.inc"
#include "lentil_internal.h"

using namespace lb;

namespace {

thread_local std::string g_last_error;
int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}
#define CU(call)                                                                                              \
  do {                                                                                                        \
    cudaError_t e_ = (call);                                                                                  \
    if (e_ != cudaSuccess) return fail(LB_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
  } while (0)

struct DeviceGuard {  // run on the camera's device, leave the caller's current device untouched
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

enum { C_OUTER_R, C_INNER_R, C_LENGTH, C_BFL, C_EFL, C_AP_POS, C_AP_HOUSING, C_INNER_CURV, C_OUTER_CURV, C_FOV, C_FSTOP, C_AP_R_FSTOP };

inline float clamp_min_f(float in, const float mn) { return in < mn ? mn : in; }  // global.h:15-18

uint32_t pack_exp(const unsigned char e[5]) { return e[0] | (e[1] << 4) | (e[2] << 8) | (e[3] << 12) | (e[4] << 16); }

// lens pack entry -> kernel table: 9 fitted polynomials + 8 derivative polynomials (term order kept,
// vanished terms dropped, coefficient c*e — what the generator prints for the Jacobians)
bool build_lens_table(int model, LensTable &T) {
  if (model < 0 || model >= LP_LENS_COUNT) return false;
  const LpLens &L = LP_LENSES[model];
  memset(&T, 0, sizeof T);
  int n = 0;
  auto push_base = [&](int slot, int p) {
    T.off[slot] = (uint16_t)n;
    for (int i = 0; i < L.cnt[p]; ++i) {
      T.t[n].c = (float)LP_COEF[L.off[p] + i];
      T.t[n].e = pack_exp(LP_EXP[L.off[p] + i]);
      ++n;
    }
    T.cnt[slot] = (uint16_t)(n - T.off[slot]);
  };
  auto push_deriv = [&](int slot, int p, int var) {
    T.off[slot] = (uint16_t)n;
    for (int i = 0; i < L.cnt[p]; ++i) {
      unsigned char e[5];
      memcpy(e, LP_EXP[L.off[p] + i], 5);
      if (e[var] == 0) continue;
      const double c = LP_COEF[L.off[p] + i] * e[var];
      e[var] -= 1;
      T.t[n].c = (float)c;
      T.t[n].e = pack_exp(e);
      ++n;
    }
    T.cnt[slot] = (uint16_t)(n - T.off[slot]);
  };
  for (int p = 0; p < 9; ++p) push_base(p, p);
  push_deriv(P_DAPX_DDX, 5, 2);
  push_deriv(P_DAPX_DDY, 5, 3);
  push_deriv(P_DAPY_DDX, 6, 2);
  push_deriv(P_DAPY_DDY, 6, 3);
  push_deriv(P_DODX_DX, 2, 0);
  push_deriv(P_DODX_DY, 2, 1);
  push_deriv(P_DODY_DX, 3, 0);
  push_deriv(P_DODY_DY, 3, 1);
  return n <= kMaxTerms;
}

double term_work(const LensTable &T, int slot) {  // F(P) = sum over terms (degree + 1), SURVEY.md §8d
  double f = 0;
  for (int i = 0; i < T.cnt[slot]; ++i) {
    uint32_t e = T.t[T.off[slot] + i].e;
    int deg = 0;
    for (int k = 0; k < 5; ++k) deg += (e >> (4 * k)) & 15;
    f += deg + 1;
  }
  return f;
}

// imageData::bokehProbability, imagebokeh.h:143-338 — host build of the sorted row/column CDFs
struct arrayCompare {
  const float *values;
  bool operator()(int l, int r) const { return values[l] > values[r]; }
};
struct BokehTables {
  int n = 0;
  std::vector<float> cdf_row, cdf_col;
  std::vector<int> row_idx, col_idx;
  // Guide tables ("cutpoint method") for the two upper_bound searches of bokehSample: entry k packs
  // {lo = upper_bound(cdf, k/K), hi = upper_bound(cdf, (k+1)/K)} as two 16-bit halves, so a search for v in
  // [k/K, (k+1)/K) only has to look at cdf[lo..hi) -- the same index the full binary search returns, found with ~1
  // dependent load instead of log2(n).  Empty when a table is not sorted or n does not fit 16 bits (full search then).
  std::vector<uint32_t> guide_row, guide_col;
  static void build_guide(const float *cdf, int n, uint32_t *g) {
    int prev = 0;  // bucket 0 starts at 0 (v < 0), the last bucket ends at n (v >= 1)
    for (int k = 0; k < kBokehGuide; ++k) {
      const int hi = k + 1 == kBokehGuide ? n : (int)(std::upper_bound(cdf, cdf + n, (float)(k + 1) / (float)kBokehGuide) - cdf);
      g[k] = (uint32_t)prev | ((uint32_t)hi << 16);
      prev = hi;
    }
  }
  void build_guides() {
    guide_row.clear(); guide_col.clear();
    if (n <= 0 || n > 65535) return;
    if (!std::is_sorted(cdf_row.begin(), cdf_row.end())) return;
    for (int r = 0; r < n; ++r)
      if (!std::is_sorted(cdf_col.begin() + (size_t)r * n, cdf_col.begin() + (size_t)(r + 1) * n)) return;
    guide_row.resize(kBokehGuide);
    guide_col.resize((size_t)n * kBokehGuide);
    build_guide(cdf_row.data(), n, guide_row.data());
    for (int r = 0; r < n; ++r) build_guide(cdf_col.data() + (size_t)r * n, n, guide_col.data() + (size_t)r * kBokehGuide);
  }
  bool build(const lb_bokeh_image *img) {
    if (!img || !img->pixels) return false;
    const int x = img->width, y = img->height, nc = img->channels;
    if (x != y || x <= 0 || nc < 3) return false;  // imagebokeh.h:49-51,97-101
    n = x;
    const int npx = x * y;
    std::vector<float> lum(npx), norm(npx), rowsum(y), per_row(npx);
    float total = 0.f;
    for (int i = 0, j = 0; i < npx; ++i, j += nc) {
      lum[i] = img->pixels[j] * 0.3f + img->pixels[j + 1] * 0.59f + img->pixels[j + 2] * 0.11f;
      total += lum[i];
    }
    const float inv_total = 1.0f / total;
    for (int i = 0; i < npx; ++i) norm[i] = lum[i] * inv_total;
    for (int r = 0, k = 0; r < y; ++r) {
      rowsum[r] = 0.f;
      for (int c = 0; c < x; ++c, ++k) rowsum[r] += norm[k];
    }
    row_idx.resize(y);
    for (int i = 0; i < y; ++i) row_idx[i] = i;
    std::sort(row_idx.begin(), row_idx.end(), arrayCompare{rowsum.data()});
    cdf_row.resize(y);
    float prev = 0.f;
    for (int i = 0; i < y; ++i) prev = cdf_row[i] = prev + rowsum[row_idx[i]];
    for (int r = 0, i = 0; r < y; ++r)
      for (int c = 0; c < x; ++c, ++i) per_row[i] = (norm[i] != 0 && rowsum[r] != 0) ? norm[i] / rowsum[r] : 0.f;
    col_idx.resize(npx);
    for (int i = 0; i < npx; ++i) col_idx[i] = i;
    for (int i = 0; i < npx; i += x) std::sort(col_idx.begin() + i, col_idx.begin() + i + x, arrayCompare{per_row.data()});
    cdf_col.resize(npx);
    for (int r = 0, i = 0; r < y; ++r) {
      prev = 0.f;
      for (int c = 0; c < x; ++c, ++i) prev = cdf_col[i] = prev + per_row[col_idx[i]];
    }
    build_guides();
    return true;
  }
};

std::vector<double> logarithmic_values() {  // lens.h:395-407
  std::vector<double> v;
  for (double i = -1.0; i <= 1.0; i += 0.0001) v.push_back((i < 0 ? -1 : 1) * std::pow(i, 2.0) * (45.0 - 0.0) + 0.0);
  return v;
}

}  // namespace

// ================================================================================================
struct FilterState;  // filter_host.cu
void filter_state_destroy(FilterState *);

struct lb_camera {
  int device = 0;
  int num_sms = 148;
  lb_camera_params params{};
  lb_camera_state st{};
  LensTable lens{};
  int lens_kernel = -1;  // LensModel of the unrolled kernel in use, -1 = table-driven
  // The reference differences two float32 traces 1e-3*dsx apart (lentil_camera.cpp:84,97-118): ~8 ulp of sx, so
  // its differentials carry a few % of float quantisation even with FP64 tracing.  FP32 tracing adds
  // evaluation noise on top; differencing over a 16x longer baseline (truncation error ~1e-5 relative) puts
  // that noise below the reference's own quantisation (measured: profiles/r01_differentials.txt).
  float deriv_baseline = 16.0f;
  CamConsts<float> camf{};
  ThinConsts thin{};
  CamConsts<double> camd{};
  // bokeh CDF on device
  float *d_cdf_row = nullptr, *d_cdf_col = nullptr;
  int32_t *d_row_idx = nullptr, *d_col_idx = nullptr;
  int bokeh_n = 0;
  bool bokeh_guided = false;  // guide tables present behind d_cdf_row / d_cdf_col
  // host-path pipeline
  cudaStream_t pipe_stream[3] = {nullptr, nullptr, nullptr};
  cudaEvent_t pipe_event[3] = {nullptr, nullptr, nullptr};
  float *d_stage[3] = {nullptr, nullptr, nullptr};
  size_t stage_rays = 0;
  FilterState *filter = nullptr;
  std::mutex mu;
};

namespace {

void free_bokeh(lb_camera *c) {
  cudaFree(c->d_cdf_row); cudaFree(c->d_cdf_col); cudaFree(c->d_row_idx); cudaFree(c->d_col_idx);
  c->d_cdf_row = c->d_cdf_col = nullptr;
  c->d_row_idx = c->d_col_idx = nullptr;
  c->bokeh_n = 0;
  c->bokeh_guided = false;
}

void refresh_consts(lb_camera *c) {
  const lb_camera_state &s = c->st;
  const lb_camera_params &p = c->params;
  auto fill = [&](auto &k) {
    using T = decltype(k.lambda);
    k.sensor_half = T((double)p.sensor_width * 0.5);
    k.lambda = T(s.lambda);
    k.lambda_exact = s.lambda;
    k.inv_outer_R = T(1.0 / s.lens_outer_pupil_curvature_radius);
    k.abs_inv_outer_R = T(1.0 / std::fabs(s.lens_outer_pupil_curvature_radius));
    k.outer_R2 = T(s.lens_outer_pupil_curvature_radius * s.lens_outer_pupil_curvature_radius);
    k.aperture_radius = T(s.aperture_radius);
    k.sensor_shift = T(s.sensor_shift);
    k.outer_R = T(s.lens_outer_pupil_curvature_radius);
    k.outer_pupil_r2 = T(s.lens_outer_pupil_radius * s.lens_outer_pupil_radius);
    k.inner_pupil_r2 = T(s.lens_inner_pupil_radius * s.lens_inner_pupil_radius);
    k.inner_R = T(s.lens_inner_pupil_curvature_radius);
    k.bfl = T(s.lens_back_focal_length);
    const float scales[4] = {(float)-1.0, (float)-0.1, (float)-0.01, (float)-0.001};
    k.unit_scale = scales[std::min(std::max(p.units, 0), 3)];
    k.exposure = p.exp;
    k.deriv_baseline = c->deriv_baseline;
    k.enable_dof = p.enable_dof != 0;
    k.vignetting_retries = p.vignetting_retries;
    k.blades = p.aperture_blades_lentil;
    k.bokeh_n = (p.bokeh_enable_image && c->bokeh_n > 0) ? c->bokeh_n : 0;
    k.outer_geom = s.outer_pupil_geometry;
    k.inner_geom = s.inner_pupil_geometry;
    k.cdf_row = c->d_cdf_row; k.row_idx = c->d_row_idx; k.cdf_col = c->d_cdf_col; k.col_idx = c->d_col_idx;
    const bool guided = k.bokeh_n > 0 && c->bokeh_guided;
    k.guide_row = guided ? reinterpret_cast<const uint32_t *>(c->d_cdf_row + c->bokeh_n) : nullptr;
    k.guide_col = guided ? reinterpret_cast<const uint32_t *>(c->d_cdf_col + (size_t)c->bokeh_n * c->bokeh_n) : nullptr;
  };
  fill(c->camf);
  fill(c->camd);
  // thin-lens scalars (get_lentil_camera_params lentil.h:1216-1229; constants hoisted out of the per-ray code)
  ThinConsts &t = c->thin;
  auto clampf = [](float in, const float mn, const float mx) { if (in < mn) in = mn; if (in > mx) in = mx; return in; };  // global.h:8-12
  t.sensor_half = (double)p.sensor_width * 0.5;
  t.focus_distance = s.focus_distance;
  t.aperture_radius = s.aperture_radius;
  t.focal_length = clamp_min_f(p.focal_length_lentil, 0.01);
  t.abb_spherical = clampf(p.abb_spherical, 0.001, 0.999);
  t.circle_to_square = clampf(p.bokeh_circle_to_square, 0.01, 0.99);
  t.bokeh_anamorphic = clampf(1.0 - p.bokeh_anamorphic, 0, 1.0);
  t.abb_coma = p.abb_coma;
  t.abb_distortion = p.abb_distortion;
  t.abb_chromatic = p.abb_chromatic;
  t.abb_chromatic_type = p.abb_chromatic_type;
  t.optical_vignetting_distance = p.optical_vignetting;
  t.optical_vignetting_radius = 1.0f;
  t.squircle = 1.0 + std::log(1.0 + t.circle_to_square) * std::exp(t.circle_to_square * 3.0);  // lerp_squircle_mapping, lens.h:544-546
  {  // maximal_projection of abb_coma_multipliers (lens.h:567-568), float arithmetic as there
    const float sw = p.sensor_width;
    const float mx = (float)(1.0 * (sw * 0.5)), mz = -t.focal_length;
    float len = std::sqrt(mx * mx + mx * mx + mz * mz);
    if (len != 0) len = 1 / len;
    t.coma_max_projection = (mx * len) * 0.0f + (mx * len) * 0.0f + (mz * len) * -1.0f;
  }
  t.image_dist_focusdist = (-t.focal_length * -t.focus_distance) / (-t.focal_length + -t.focus_distance);  // lentil.h:664-666
  const float tl_scales[4] = {(float)10.0, (float)1.0, (float)0.1, (float)0.01};
  t.unit_scale = tl_scales[std::min(std::max(p.units, 0), 3)];
}

// camera_create_ray dispatch on cameraType (lentil_camera.cpp:90-94)
cudaError_t launch_rays(lb_camera *c, const RayIO &io, size_t n, uint64_t ray_id_base, cudaStream_t stream) {
  if (c->params.camera_type == LB_CAMERA_THINLENS) return launch_create_rays_thinlens(c->camf, c->thin, io, n, ray_id_base, stream);
  return launch_create_rays(c->lens_kernel, c->lens, c->camf, io, n, ray_id_base, stream);
}

template <typename T>
struct DevBuf {  // device scratch freed on every exit path
  T *p = nullptr;
  ~DevBuf() { cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, n * sizeof(T)); }
};

// get_lentil_camera_params (lentil.h:1189-1243) + camera_model_specific_setup (lentil.h:1568-1670).
// Transactional: a failure (unknown lens, bad bokeh image, CUDA error) leaves the camera in its previous valid state.
int camera_setup_impl(lb_camera *c, const lb_camera_params *p, const lb_bokeh_image *bokeh);
int camera_setup(lb_camera *c, const lb_camera_params *p, const lb_bokeh_image *bokeh) {
  struct Saved {
    lb_camera_params params; lb_camera_state st; LensTable lens; int lens_kernel; CamConsts<float> camf; ThinConsts thin; CamConsts<double> camd;
    float *cdf_row, *cdf_col; int32_t *row_idx, *col_idx; int bokeh_n; bool bokeh_guided;
  };
  Saved *old = new Saved{c->params, c->st, c->lens, c->lens_kernel, c->camf, c->thin, c->camd, c->d_cdf_row, c->d_cdf_col, c->d_row_idx, c->d_col_idx, c->bokeh_n, c->bokeh_guided};
  // the new tables are built beside the old ones; whichever set loses is freed below
  c->d_cdf_row = c->d_cdf_col = nullptr;
  c->d_row_idx = c->d_col_idx = nullptr;
  c->bokeh_n = 0;
  c->bokeh_guided = false;
  const int rc = camera_setup_impl(c, p, bokeh);
  if (rc != LB_OK) {
    free_bokeh(c);
    c->params = old->params; c->st = old->st; c->lens = old->lens; c->lens_kernel = old->lens_kernel; c->camf = old->camf; c->thin = old->thin; c->camd = old->camd;
    c->d_cdf_row = old->cdf_row; c->d_cdf_col = old->cdf_col; c->d_row_idx = old->row_idx; c->d_col_idx = old->col_idx; c->bokeh_n = old->bokeh_n; c->bokeh_guided = old->bokeh_guided;
  } else {
    cudaDeviceSynchronize();  // launches that still read the previous tables
    cudaFree(old->cdf_row); cudaFree(old->cdf_col); cudaFree(old->row_idx); cudaFree(old->col_idx);
  }
  delete old;
  return rc;
}
int camera_setup_impl(lb_camera *c, const lb_camera_params *p, const lb_bokeh_image *bokeh) {
  LensTable table;
  if (!build_lens_table(p->lens_model, table)) return fail(LB_ERR_LENS, "unknown lens_model %d", p->lens_model);
  c->lens = table;
  c->params = *p;
  lb_camera_state &s = c->st;
  memset(&s, 0, sizeof s);
  const LpLens &L = LP_LENSES[p->lens_model];
  s.lens_outer_pupil_radius = L.c[C_OUTER_R];
  s.lens_inner_pupil_radius = L.c[C_INNER_R];
  s.lens_length = L.c[C_LENGTH];
  s.lens_back_focal_length = L.c[C_BFL];
  s.lens_effective_focal_length = L.c[C_EFL];
  s.lens_aperture_pos = L.c[C_AP_POS];
  s.lens_aperture_housing_radius = L.c[C_AP_HOUSING];
  s.lens_inner_pupil_curvature_radius = L.c[C_INNER_CURV];
  s.lens_outer_pupil_curvature_radius = L.c[C_OUTER_CURV];
  s.lens_field_of_view = L.c[C_FOV];
  s.lens_fstop = L.c[C_FSTOP];
  s.lens_aperture_radius_at_fstop = L.c[C_AP_R_FSTOP];
  s.outer_pupil_geometry = L.outer_geom;
  s.inner_pupil_geometry = L.inner_geom;
  const double input_fstop = clamp_min_f(p->fstop, 0.01);
  const float focal_length = clamp_min_f(p->focal_length_lentil, 0.01);
  s.focus_distance = p->focus_dist;
  s.lambda = p->wavelength * 0.001;
  // LB_FORCE_TABLE=1 selects the table-driven kernels even when the lens has unrolled ones (tests, A/B timing)
  const char *force_table = getenv("LB_FORCE_TABLE");
  c->lens_kernel = (has_unrolled_kernel(p->lens_model) && !(force_table && force_table[0] == '1')) ? p->lens_model : -1;

  // bokeh image -> CDF tables (lentil.h:222-228)
  if (p->bokeh_enable_image) {
    BokehTables bt;
    if (!bt.build(bokeh)) return fail(LB_ERR_IMAGE, "bokeh image missing, not square or < 3 channels");
    const size_t n = bt.n, n2 = n * n;
    // the guide tables ride behind the CDFs in the same allocations
    const bool guided = !bt.guide_row.empty() && !(getenv("LB_NO_CDF_GUIDE") && getenv("LB_NO_CDF_GUIDE")[0] == '1');
    CU(cudaMalloc(&c->d_cdf_row, (n + (guided ? kBokehGuide : 0)) * 4)); CU(cudaMalloc(&c->d_row_idx, n * 4));
    CU(cudaMalloc(&c->d_cdf_col, (n2 + (guided ? n * kBokehGuide : 0)) * 4)); CU(cudaMalloc(&c->d_col_idx, n2 * 4));
    c->bokeh_guided = guided;
    if (guided) {
      CU(cudaMemcpy(c->d_cdf_row + n, bt.guide_row.data(), kBokehGuide * 4, cudaMemcpyHostToDevice));
      CU(cudaMemcpy(c->d_cdf_col + n2, bt.guide_col.data(), n * kBokehGuide * 4, cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(c->d_cdf_row, bt.cdf_row.data(), n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_row_idx, bt.row_idx.data(), n * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_cdf_col, bt.cdf_col.data(), n2 * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(c->d_col_idx, bt.col_idx.data(), n2 * 4, cudaMemcpyHostToDevice));
    c->bokeh_n = bt.n;
  }

  if (p->camera_type == LB_CAMERA_POLYNOMIAL_OPTICS) {
    s.focus_distance *= 10.0;
    refresh_consts(c);
    // -- aperture radius from the f-stop (lentil.h:1603-1613) --
    if (input_fstop == 0.0) {
      s.aperture_radius = s.lens_aperture_radius_at_fstop;
    } else {
      const int maxrays = 1000;
      DevBuf<double4> out;
      CU(out.alloc(maxrays));
      CU(cudaMemset(out.p, 0, maxrays * sizeof(double4)));  // entry 0 is not a ray (the loop of lentil.h:1395 starts at 1)
      CU(launch_fstop_rays(c->lens, c->camd, maxrays, s.lens_outer_pupil_radius, out.p, nullptr));
      std::vector<double4> h(maxrays);
      CU(cudaMemcpy(h.data(), out.p, maxrays * sizeof(double4), cudaMemcpyDeviceToHost));
      // sequential selection of trace_backwards_for_fstop (lentil.h:1395-1440)
      double best_fstop = 0.0, best_radius = 0.0, calc_radius = 0.0;
      bool returned = false;
      for (int i = 1; i < maxrays; ++i) {
        if (h[i].x == 0.0) continue;
        const double parallel_ray_height = (static_cast<double>(i) / static_cast<double>(maxrays)) * s.lens_outer_pupil_radius;
        const double theta = std::atan(h[i].y / h[i].z);
        const double fstop = 1.0 / (std::sin(theta) * 2.0);
        if (fstop < input_fstop) { calc_radius = best_radius; returned = true; break; }
        best_fstop = fstop;
        best_radius = parallel_ray_height;
      }
      if (!returned) calc_radius = best_radius;
      (void)best_fstop;
      s.aperture_radius = std::min(s.lens_aperture_radius_at_fstop, calc_radius);
    }
    // -- sensor shift: logarithmic focus search (lentil.h:1445-1460,1632-1634) --
    {
      const std::vector<double> shifts = logarithmic_values();
      const int n = (int)shifts.size();
      DevBuf<double> shifts_buf;
      DevBuf<double4> out_buf;
      CU(shifts_buf.alloc(n + 1));
      CU(out_buf.alloc(n + 1));
      double *d_shifts = shifts_buf.p;
      double4 *d_out = out_buf.p;
      CU(cudaMemcpy(d_shifts, shifts.data(), n * sizeof(double), cudaMemcpyHostToDevice));
      const double ap_y = s.lens_aperture_housing_radius * 0.25;
      CU(launch_focus_distances(c->lens, c->camd, ap_y, d_shifts, n, d_out, nullptr));
      std::vector<double4> h(n);
      CU(cudaMemcpy(h.data(), d_out, n * sizeof(double4), cudaMemcpyDeviceToHost));
      double closest_distance = 999999999.0, best_sensor_shift = 0.0;
      for (int i = 0; i < n; ++i) {
        const double new_distance = s.focus_distance - h[i].x;
        if (new_distance < closest_distance && new_distance > 0.0) { closest_distance = new_distance; best_sensor_shift = shifts[i]; }
      }
      s.sensor_shift = best_sensor_shift + p->extra_sensor_shift;
      // trace_ray_focus_check (lentil.h:1316-1357) at the chosen shift
      CU(cudaMemcpy(d_shifts + n, &s.sensor_shift, sizeof(double), cudaMemcpyHostToDevice));
      CU(launch_focus_distances(c->lens, c->camd, ap_y, d_shifts + n, 1, d_out + n, nullptr));
      double4 chk;
      CU(cudaMemcpy(&chk, d_out + n, sizeof chk, cudaMemcpyDeviceToHost));
      s.focus_check_ok = chk.y > 0.0 && !(chk.z > s.lens_outer_pupil_radius * s.lens_outer_pupil_radius) &&
                         !(chk.w > s.lens_inner_pupil_radius * s.lens_inner_pupil_radius);
      s.focus_check_distance = s.focus_check_ok ? chk.x : 0.0;
    }
    s.tan_fov = std::tan(s.lens_field_of_view / 2.0);
  } else {  // ThinLens (lentil.h:1663-1668)
    const float fov = 2.0 * std::atan(p->sensor_width / (2.0 * focal_length));
    s.tan_fov = std::tan(fov / 2.0);
    s.aperture_radius = (focal_length / (2.0 * input_fstop)) / 10.0;
  }
  refresh_consts(c);
  return LB_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

void lb_camera_params_default(lb_camera_params *p) {  // lentil_camera.cpp:19-52
  if (!p) return;
  memset(p, 0, sizeof *p);
  p->camera_type = LB_CAMERA_THINLENS;
  p->bidir_sample_mult = 5;
  p->units = LB_UNITS_CM;
  p->sensor_width = 36.0f;
  p->enable_dof = 1;
  p->fstop = 0.0f;
  p->focus_dist = 150.0f;
  p->exp = 1.0f;
  p->lens_model = 16;  // cooke__speed_panchro__1920__40mm
  p->wavelength = 550.0f;
  p->focal_length_lentil = 35.0f;
  p->abb_spherical = 0.5f;
  p->vignetting_retries = 15;
  p->bidir_add_energy_minimum_luminance = 2.0f;
  p->bidir_add_energy_transition = 1.0f;
}
int lb_lens_count(void) { return LP_LENS_COUNT; }
const char *lb_lens_name(int m) { return (m >= 0 && m < LP_LENS_COUNT) ? LP_LENSES[m].name : nullptr; }
const char *lb_last_error(void) { return g_last_error.c_str(); }
const char *lb_version(void) { return "lentil_b200 0.4.0 (sm_100a)"; }

int lb_camera_create(const lb_camera_params *params, const lb_bokeh_image *bokeh, int device, lb_camera **out) {
  if (!params || !out) return fail(LB_ERR_INVALID, "null argument");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return fail(LB_ERR_NO_DEVICE, "no CUDA device: lentil_b200 has no CPU path");
  }
  if (device < 0 || device >= ndev) return fail(LB_ERR_INVALID, "device %d out of range (%d devices)", device, ndev);
  DeviceGuard g(device);
  lb_camera *c = new lb_camera();
  c->device = device;
  if (const char *e = getenv("LB_DERIV_BASELINE")) c->deriv_baseline = (float)atof(e);
  cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device);
  int rc = camera_setup(c, params, bokeh);
  if (rc != LB_OK) { lb_camera_destroy(c); return rc; }
  *out = c;
  return LB_OK;
}

int lb_camera_update(lb_camera *c, const lb_camera_params *params, const lb_bokeh_image *bokeh) {  // node_update
  if (!c || !params) return fail(LB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(c->mu);  // the reference serialises setup_camera with a crit-sec (lentil.h:212,235)
  DeviceGuard g(c->device);
  return camera_setup(c, params, bokeh);
}

void lb_camera_destroy(lb_camera *c) {
  if (!c) return;
  DeviceGuard g(c->device);
  free_bokeh(c);
  for (int i = 0; i < 3; ++i) {
    if (c->pipe_stream[i]) cudaStreamDestroy(c->pipe_stream[i]);
    if (c->pipe_event[i]) cudaEventDestroy(c->pipe_event[i]);
    cudaFree(c->d_stage[i]);
  }
  if (c->filter) filter_state_destroy(c->filter);
  delete c;
}

int lb_camera_get_state(const lb_camera *c, lb_camera_state *out) {
  if (!c || !out) return fail(LB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(const_cast<lb_camera *>(c)->mu);
  *out = c->st;
  return LB_OK;
}
int lb_camera_set_state(lb_camera *c, double aperture_radius, double sensor_shift) {
  if (!c) return fail(LB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(c->mu);
  c->st.aperture_radius = aperture_radius;
  c->st.sensor_shift = sensor_shift;
  refresh_consts(c);
  return LB_OK;
}

int lb_camera_set_pupil_geometry(lb_camera *c, int outer_geometry, int inner_geometry) {
  if (!c || outer_geometry < 0 || outer_geometry > 2 || inner_geometry < 0 || inner_geometry > 2) return fail(LB_ERR_INVALID, "geometry must be 0, 1 or 2");
  std::lock_guard<std::mutex> lk(c->mu);
  c->st.outer_pupil_geometry = outer_geometry;
  c->st.inner_pupil_geometry = inner_geometry;
  refresh_consts(c);
  return LB_OK;
}

int lb_camera_kernel_kind(const lb_camera *c) { return (c && c->lens_kernel >= 0) ? 1 : 0; }

int lb_camera_lens_work(const lb_camera *c, lb_lens_work *w) {
  if (!c || !w) return fail(LB_ERR_INVALID, "null argument");
  const LensTable &T = c->lens;
  auto F = [&](std::initializer_list<int> slots) { double f = 0; for (int s : slots) f += term_work(T, s); return f; };
  auto N = [&](std::initializer_list<int> slots) { int n = 0; for (int s : slots) n += T.cnt[s]; return n; };
  w->terms_eval = N({P_OUT_X, P_OUT_Y, P_OUT_DX, P_OUT_DY, P_OUT_T});
  w->terms_ap = N({P_AP_X, P_AP_Y, P_AP_DX, P_AP_DY});
  w->terms_ap_jac = N({P_DAPX_DDX, P_DAPX_DDY, P_DAPY_DDX, P_DAPY_DDY});
  w->terms_out_jac = N({P_DODX_DX, P_DODX_DY, P_DODY_DX, P_DODY_DY});
  w->F_eval = F({P_OUT_X, P_OUT_Y, P_OUT_DX, P_OUT_DY, P_OUT_T});
  w->F_ap = F({P_AP_X, P_AP_Y, P_AP_DX, P_AP_DY, P_DAPX_DDX, P_DAPX_DDY, P_DAPY_DDX, P_DAPY_DDY});
  w->F_apxy = F({P_AP_X, P_AP_Y});
  w->F_apJ = F({P_DAPX_DDX, P_DAPX_DDY, P_DAPY_DDX, P_DAPY_DDY});
  w->F_out4 = F({P_OUT_X, P_OUT_Y, P_OUT_DX, P_OUT_DY});
  w->F_outJ = F({P_DODX_DX, P_DODX_DY, P_DODY_DX, P_DODY_DY});
  w->F_T = F({P_OUT_T});
  return LB_OK;
}

int lb_camera_create_rays(lb_camera *c, size_t n, uint64_t ray_id_base, const lb_ray_in *in, const lb_ray_out *out, lb_stream stream) {
  if (!c || !in || !out) return fail(LB_ERR_INVALID, "null argument");
  if (n == 0) return LB_OK;
  if (!in->sx || !in->sy || !in->dsx || !in->dsy || !in->lensx || !in->lensy) return fail(LB_ERR_INVALID, "null input array");
  std::lock_guard<std::mutex> lk(c->mu);  // the launch reads camf / the lens table / the bokeh pointers: not under a concurrent lb_camera_update
  DeviceGuard g(c->device);
  RayIO io{in->sx, in->sy, in->dsx, in->dsy, in->lensx, in->lensy, out->origin, out->dir, out->dOdx, out->dOdy,
           out->dDdx, out->dDdy, out->weight, out->tries, n};
  CU(launch_rays(c, io, n, ray_id_base, (cudaStream_t)stream));
  return LB_OK;
}

// Host-buffer variant: chunks of `chunk` rays flow through a 3-deep pipeline, each slot with its
// own stream and device staging block [6 in | 21 out | tries][chunk], so the H2D of chunk k+1, the
// kernel of chunk k and the D2H of chunk k-1 overlap (PCIe is full duplex).
int lb_camera_create_rays_host(lb_camera *c, size_t n, uint64_t ray_id_base, const lb_ray_in *in, const lb_ray_out *out) {
  if (!c || !in || !out) return fail(LB_ERR_INVALID, "null argument");
  if (n == 0) return LB_OK;
  if (!in->sx || !in->sy || !in->dsx || !in->dsy || !in->lensx || !in->lensy) return fail(LB_ERR_INVALID, "null input array");
  std::lock_guard<std::mutex> lk(c->mu);
  DeviceGuard g(c->device);
  const size_t chunk = std::min<size_t>(std::max<size_t>(n, 1), (size_t)1 << 21);  // 2 Mi rays: 48 MB in, 176 MB out per slot
  const size_t slot_floats = chunk * 28;
  if (c->stage_rays < chunk) {
    for (int i = 0; i < 3; ++i) {
      cudaFree(c->d_stage[i]);
      c->d_stage[i] = nullptr;
      CU(cudaMalloc(&c->d_stage[i], slot_floats * sizeof(float)));
      if (!c->pipe_stream[i]) CU(cudaStreamCreateWithFlags(&c->pipe_stream[i], cudaStreamNonBlocking));
      if (!c->pipe_event[i]) CU(cudaEventCreateWithFlags(&c->pipe_event[i], cudaEventDisableTiming));
    }
    c->stage_rays = chunk;
  }
  const float *src[6] = {in->sx, in->sy, in->dsx, in->dsy, in->lensx, in->lensy};
  float *dst[7] = {out->origin, out->dir, out->dOdx, out->dOdy, out->dDdx, out->dDdy, out->weight};
  // The three channels of `weight` are the same number (AtRGB white or black times the exposure, lentil.h:379,
  // lentil_camera.cpp:124): only plane 0 crosses the host link (the path is bound by its device-to-host direction,
  // 84 -> 76 B per ray), and this thread fills planes 1 and 2 from it while later chunks are in flight.
  // chunk schedule: a long call ramps up with chunk/8, chunk/4, chunk/2 so the first device-to-host copy starts early
  std::vector<size_t> starts, sizes;
  for (size_t b = 0; b < n;) {
    size_t m = chunk;
    if (n > 4 * chunk && starts.size() < 3) m = chunk >> (3 - starts.size());
    m = std::min(m, n - b);
    starts.push_back(b);
    sizes.push_back(m);
    b += m;
  }
  auto replicate_weight = [&](size_t kk) -> cudaError_t {
    const size_t b = starts[kk], m = sizes[kk];
    cudaError_t e = cudaEventSynchronize(c->pipe_event[kk % 3]);
    if (e != cudaSuccess) return e;
    memcpy(out->weight + n + b, out->weight + b, m * sizeof(float));
    memcpy(out->weight + 2 * n + b, out->weight + b, m * sizeof(float));
    return cudaSuccess;
  };
  size_t k = 0;
  for (; k < starts.size(); ++k) {
    const int s = (int)(k % 3);
    const size_t base = starts[k], m = sizes[k];
    cudaStream_t st = c->pipe_stream[s];
    float *d = c->d_stage[s];
    for (int a = 0; a < 6; ++a) CU(cudaMemcpyAsync(d + a * chunk, src[a] + base, m * sizeof(float), cudaMemcpyHostToDevice, st));
    RayIO io{};
    io.sx = d; io.sy = d + chunk; io.dsx = d + 2 * chunk; io.dsy = d + 3 * chunk; io.lensx = d + 4 * chunk; io.lensy = d + 5 * chunk;
    float *o = d + 6 * chunk;
    float **slots[7] = {&io.origin, &io.dir, &io.dOdx, &io.dOdy, &io.dDdx, &io.dDdy, &io.weight};
    for (int v = 0; v < 7; ++v) *slots[v] = dst[v] ? o + (size_t)v * 3 * chunk : nullptr;
    io.tries = out->tries ? (int32_t *)(o + 21 * chunk) : nullptr;
    io.plane = chunk;
    CU(launch_rays(c, io, m, ray_id_base + base, st));
    for (int v = 0; v < 6; ++v)
      if (dst[v])  // 3 planes of the chunk -> 3 plane ranges of the user's [3][n] array: one strided copy
        CU(cudaMemcpy2DAsync(dst[v] + base, n * sizeof(float), o + (size_t)v * 3 * chunk, chunk * sizeof(float), m * sizeof(float), 3,
                             cudaMemcpyDeviceToHost, st));
    if (out->weight) CU(cudaMemcpyAsync(out->weight + base, io.weight, m * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (out->tries) CU(cudaMemcpyAsync(out->tries + base, io.tries, m * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(c->pipe_event[s], st));
    if (out->weight && k >= 2) CU(replicate_weight(k - 2));  // two chunks stay in flight behind this wait
  }
  if (out->weight) {
    if (k >= 2) CU(replicate_weight(k - 2));
    CU(replicate_weight(k - 1));
  }
  for (int i = 0; i < 3; ++i) CU(cudaStreamSynchronize(c->pipe_stream[i]));
  return LB_OK;
}

int lb_camera_reverse_rays(lb_camera *c, size_t n, const float *Po, float *Ps, lb_stream stream) {
  if (!c || !Po || !Ps) return fail(LB_ERR_INVALID, "null argument");
  DeviceGuard g(c->device);
  CU(launch_reverse_rays(n, (const float4 *)Po, (float2 *)Ps, (float)c->st.tan_fov, (cudaStream_t)stream));
  return LB_OK;
}

}  // extern "C"

// accessors for filter_host.cu
namespace lb {
int cam_device(lb_camera *c) { return c->device; }
int cam_num_sms(lb_camera *c) { return c->num_sms; }
const lb_camera_params &cam_params(lb_camera *c) { return c->params; }
const lb_camera_state &cam_state(lb_camera *c) { return c->st; }
const LensTable &cam_lens(lb_camera *c) { return c->lens; }
int cam_lens_kernel(lb_camera *c) { return c->lens_kernel; }
const CamConsts<float> &cam_consts(lb_camera *c) { return c->camf; }
const ThinConsts &cam_thin(lb_camera *c) { return c->thin; }
FilterState *&cam_filter(lb_camera *c) { return c->filter; }
std::mutex &cam_mutex(lb_camera *c) { return c->mu; }
int lb_fail(int code, const char *msg) { return fail(code, "%s", msg); }
}  // namespace lb
