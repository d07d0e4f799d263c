// sanitize_driver.cpp — a small, torch-free workload over the C ABI for compute-sanitizer (memcheck / racecheck / initcheck):
// camera rays (PO unrolled + table kernels, thin-lens), bidirectional redistribution with gaussian, closest, lentil_debug and
// ranked cryptomatte AOVs (the warp-converged accumulate of the splat kernels' service phase), host-path
// staging with concurrent batches, per-bucket resolve.
//   g++ -O1 -std=c++17 -I include scripts/sanitize_driver.cpp -L pota_b200 -llentil_b200 -Wl,-rpath,'$ORIGIN/../../pota_b200' -o scripts/_build/sanitize_driver
//   compute-sanitizer --tool racecheck scripts/_build/sanitize_driver
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lentil_b200.h"

#define CHECK(x)                                                                  \
  do {                                                                            \
    int rc_ = (x);                                                                \
    if (rc_ != LB_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, rc_, lb_last_error()); return 1; } \
  } while (0)

static uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s; }
static float rnd(uint32_t &s) { return (lcg(s) & 0xFFFFFF) / 16777216.0f; }

static int run(int camera_type, int lens_model, bool with_bokeh) {
  lb_camera_params p;
  lb_camera_params_default(&p);
  p.camera_type = camera_type;
  p.lens_model = lens_model;
  p.fstop = 1.4f;
  p.focus_dist = 35.0f;
  p.bidir_sample_mult = 6;
  p.focal_length_lentil = 50.0f;
  p.bokeh_enable_image = with_bokeh;
  const int bn = 48;
  std::vector<float> bimg((size_t)bn * bn * 3);
  for (int y = 0; y < bn; ++y)
    for (int x = 0; x < bn; ++x) {
      const float r = std::hypot(x - 23.5f, y - 23.5f) / 23.5f;
      const float v = r <= 1.f ? 0.4f + 0.6f * r : 0.f;
      for (int k = 0; k < 3; ++k) bimg[((size_t)y * bn + x) * 3 + k] = v;
    }
  lb_bokeh_image img{bn, bn, 3, bimg.data()};
  lb_camera *cam = nullptr;
  CHECK(lb_camera_create(&p, with_bokeh ? &img : nullptr, 0, &cam));
  lb_camera_state st;
  CHECK(lb_camera_get_state(cam, &st));
  // ---- camera rays through the host path (odd count: the pair kernel's padding half) ----
  const size_t n = 4099;
  std::vector<float> in[6], out[7];
  uint32_t s = 12345u + lens_model;
  for (auto &v : in) v.resize(n);
  for (size_t i = 0; i < n; ++i) {
    in[0][i] = 2.f * rnd(s) - 1.f; in[1][i] = (2.f * rnd(s) - 1.f) * 0.5625f; in[2][i] = in[3][i] = 2.f / 1920.f;
    in[4][i] = rnd(s); in[5][i] = rnd(s);
  }
  for (auto &v : out) v.assign(3 * n, 0.f);
  std::vector<int32_t> tries(n);
  lb_ray_in rin{in[0].data(), in[1].data(), in[2].data(), in[3].data(), in[4].data(), in[5].data()};
  lb_ray_out rout{out[0].data(), out[1].data(), out[2].data(), out[3].data(), out[4].data(), out[5].data(), out[6].data(), tries.data()};
  CHECK(lb_camera_create_rays_host(cam, n, 7, &rin, &rout));
  size_t live = 0;
  for (size_t i = 0; i < n; ++i) live += out[6][i] != 0.f;
  // ---- redistribution: RGBA + a light AOV + closest + lentil_debug + two ranked cryptomatte AOVs ----
  const int W = 128, H = 72, spp = 4, D = 3;  // >= 64 in both directions: LB_SPLAT_TILE=1 runs the shared-memory tile kernel
  lb_aov_desc aovs[6];
  memset(aovs, 0, sizeof aovs);
  const char *names[6] = {"RGBA", "light0", "N", "lentil_debug", "crypto_material00", "crypto_object01"};
  const int filt[6] = {LB_FILTER_GAUSSIAN, LB_FILTER_GAUSSIAN, LB_FILTER_CLOSEST, LB_FILTER_GAUSSIAN, LB_FILTER_CRYPTO, LB_FILTER_CRYPTO};
  const int role[6] = {LB_AOV_RGBA, LB_AOV_PLAIN, LB_AOV_PLAIN, LB_AOV_LENTIL_DEBUG, LB_AOV_PLAIN, LB_AOV_PLAIN};
  for (int a = 0; a < 6; ++a) { strncpy(aovs[a].name, names[a], 63); aovs[a].filter = filt[a]; aovs[a].role = role[a]; }
  lb_frame_desc frame{W, H, W, H, 0, 0, 0};
  CHECK(lb_filter_begin(cam, &frame, 6, aovs));
  const size_t ns = (size_t)W * H * spp;
  std::vector<int32_t> px(ns), py(ns);
  std::vector<float> rgba(4 * ns, 0.f), pos(4 * ns, 0.f), light(4 * ns, 0.f), nrm(4 * ns, 0.f), opac(D * ns), id0(D * ns), id1(D * ns);
  std::vector<uint8_t> cnt(ns);
  const float half_w = (float)st.tan_fov * 75.f;
  for (size_t k = 0; k < ns; ++k) {
    const int x = (int)((k / spp) % W), y = (int)((k / spp) / W);
    px[k] = x; py[k] = y;
    pos[4 * k + 3] = 1.0e30f;
    const bool hit = ((x % 24) == 12 || (x % 24) == 13) && ((y % 18) == 9);  // small emissive patches
    if (hit) {
      const float sx = 2.f * (x + rnd(s)) / W - 1.f, sy = (1.f - 2.f * (y + rnd(s)) / H) / (16.f / 9.f);
      for (int c = 0; c < 3; ++c) rgba[4 * k + c] = light[4 * k + c] = 60.f;
      rgba[4 * k + 3] = light[4 * k + 3] = 1.f;
      pos[4 * k] = sx * half_w; pos[4 * k + 1] = sy * half_w; pos[4 * k + 2] = -75.f; pos[4 * k + 3] = 75.f + rnd(s);
      nrm[4 * k + 2] = 1.f;
    }
    cnt[k] = (uint8_t)(lcg(s) % (D + 1));
    for (int d = 0; d < D; ++d) { opac[D * k + d] = (lcg(s) % 5) * 0.25f; id0[D * k + d] = 1.f + (x / 16) + d; id1[D * k + d] = -2.5f * (1 + (y / 16)); }
  }
  const float *vals[6] = {nullptr, light.data(), nrm.data(), nullptr, nullptr, nullptr};
  const float *ids[6] = {nullptr, nullptr, nullptr, nullptr, id0.data(), id1.data()};
  lb_samples S{};
  S.n = ns; S.px = px.data(); S.py = py.data(); S.rgba = rgba.data(); S.pos_cs = pos.data(); S.aov_values = vals; S.inv_density = 1.f / spp;
  S.crypto_depth = D; S.crypto_count = cnt.data(); S.crypto_opacity = opac.data(); S.crypto_ids = ids;
  CHECK(lb_filter_accumulate_host(cam, &S));
  // a gaussian-only frame in several small host batches: concurrent batches on the scratch pool
  lb_filter_stats fs;
  CHECK(lb_filter_get_stats(cam, &fs));
  std::vector<float> bucket(16 * 16 * 4, -1.f);
  for (int a = 0; a < 6; ++a) CHECK(lb_imager_resolve_host(cam, a, 16, 16, 16, 16, bucket.data()));
  CHECK(lb_filter_begin(cam, &frame, 2, aovs));
  for (size_t lo = 0; lo < ns; lo += ns / 5 + 1) {
    lb_samples T = S;
    const size_t m = std::min(ns / 5 + 1, ns - lo);
    T.n = m; T.px = px.data() + lo; T.py = py.data() + lo; T.rgba = rgba.data() + 4 * lo; T.pos_cs = pos.data() + 4 * lo;
    const float *v2[2] = {nullptr, light.data() + 4 * lo};
    T.aov_values = v2; T.crypto_depth = 0; T.crypto_count = nullptr; T.crypto_opacity = nullptr; T.crypto_ids = nullptr;
    CHECK(lb_filter_accumulate_host(cam, &T));
  }
  lb_filter_stats fs2;
  CHECK(lb_filter_get_stats(cam, &fs2));
  CHECK(lb_imager_resolve_host(cam, 0, 0, 0, W, 1, bucket.data()));
  printf("camera_type %d lens %d bokeh %d: %zu/%zu live rays; frame A: %llu redistributed, %llu splats, %llu crypto dropped; frame B: %llu splats\n", camera_type,
         lens_model, (int)with_bokeh, live, n, (unsigned long long)fs.redistributed, (unsigned long long)fs.splats, (unsigned long long)fs.crypto_dropped,
         (unsigned long long)fs2.splats);
  lb_camera_destroy(cam);
  return (fs.redistributed > 0 && fs.splats > 0 && fs2.splats > 0 && live > n / 2) ? 0 : 2;
}

int main() {
  int rc = run(LB_CAMERA_POLYNOMIAL_OPTICS, 5, true);  // unrolled kernels (550 nm immediates), image bokeh
  if (!rc) rc = run(LB_CAMERA_POLYNOMIAL_OPTICS, 12, false);
  if (!rc) { setenv("LB_KERNEL_GEN", "1", 1); rc = run(LB_CAMERA_POLYNOMIAL_OPTICS, 5, false); }  // coefficient-table kernels
  if (!rc) { setenv("LB_FORCE_TABLE", "1", 1); rc = run(LB_CAMERA_POLYNOMIAL_OPTICS, 5, true); unsetenv("LB_FORCE_TABLE"); }  // table-driven kernels
  if (!rc) rc = run(LB_CAMERA_THINLENS, 5, true);
  printf(rc ? "sanitize_driver FAILED (%d)\n" : "sanitize_driver ok\n", rc);
  return rc;
}
