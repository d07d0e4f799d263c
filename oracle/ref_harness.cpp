// ref_harness.cpp — drives the REFERENCE'S OWN SOURCES (compiled where they lie under /root/reference/src
// behind oracle/shims/ai.h + the Eigen stand-in, see ref.mk) through the same C entry points as the
// oracle.  TEST INFRASTRUCTURE: used to pin oracle/lentil_oracle.cpp against the reference itself and to
// generate tests/golden/*.npz (tests/golden/make_golden.py).  Output only into oracle/_ref/.
//
// What is the reference here: struct Camera (src/lentil.h), camera_create_ray (src/lentil_camera.cpp),
// filter_pixel (src/lentil_filter.cpp), driver_process_bucket (src/lentil_imager.cpp), src/lens.h,
// src/global.h, src/imagebokeh.h — unmodified.  What is NOT: the Arnold SDK (shimmed), Eigen (shimmed), and
// the per-lens generated bodies, which come from the build's lens pack in the generator's format
// (oracle/_ref/gen/polynomial-optics/..., written by pota_b200.lensgen.emit).
#include <ai.h>

#include <chrono>
#include <deque>
#include <iostream>
#include <regex>
#include <thread>

#include "../include/lentil_b200.h"
#ifdef LB_ADAPTOR
// Built a second time by adaptor/Makefile with -DLB_ADAPTOR: the SAME scenario code (universe, nodes, parameters, operator
// AOV list, per-sample CreateRay, per-pixel FilterPixel, per-bucket DriverProcessBucket) then drives the node tables of the
// adaptor (adaptor/lentil_b200_*.cpp over liblentil_b200.so) instead of the reference's.  Only the accessors that reach
// into the reference's `struct Camera` differ: they go through the C ABI.
#include "../CryptomatteArnold/cryptomatte/cryptomatte.h"
#include "lentil_b200_adaptor.h"
typedef LbAdaptorCamera Camera;
#else
// The three lens_* wrappers (lentil.h:1257-1313) sit behind `private:`; this test harness calls them
// directly to pin the oracle's restatement of the generated bodies.  Layout is unaffected.
#define private public
#include "lentil.h"
#undef private
#endif

extern const AtNodeMethods *lentilMethods;         // lentil_camera.cpp:5
extern const AtNodeMethods *LentilFilterDataMtd;   // lentil_filter.cpp:6
extern const AtNodeMethods *LentilImagerMtd;       // lentil_imager.cpp:7
extern const AtNodeMethods *LentilOperatorMtd;     // lentil_operator.cpp:8
extern "C" bool NodeLoader(int i, AtNodeLib *node);  // lentil_loader.cpp:20-28

struct ref_camera {
  AtUniverse uni;
  AtNode options, camera, filter, imager, op, crypto, exr, f_gauss, f_closest, f_crypto;
  bool use_operator = false;  // the AOV list comes from the library's own lentil_operator node instead of build_operator_aovs
  void *op_user = nullptr;
  CryptomatteData cryptodata;
  AtArray aov_shaders, outputs;
  OperatorData opdata;
  Camera *cam = nullptr;
  std::vector<lb_aov_desc> aovs;
  lb_frame_desc frame{};
  int aa = 3;
  size_t scene_nodes = 0;  // uni.nodes entries that belong to the scene itself
  size_t n_user_aovs = 0;  // aovs[0 .. n_user_aovs) came with ref_filter_begin, the rest are the operator's helper outputs
#ifdef LB_ADAPTOR
  std::vector<float> host_buffer, host_weight;  // lb_filter_buffers_host copies handed out by ref_filter_buffers
#endif
};

namespace {
void set_i(AtNode &n, const char *k, int v) { n.params[k].i = v; }
void set_f(AtNode &n, const char *k, float v) { n.params[k].f = v; }
void set_b(AtNode &n, const char *k, bool v) { n.params[k].b = v; }
void set_s(AtNode &n, const char *k, const char *v) { n.params[k].s = v; }

void apply_params(ref_camera *r, const lb_camera_params *p) {  // names: lentil_camera.cpp:19-52
  AtNode &c = r->camera;
  set_i(c, "camera_type", p->camera_type);
  set_i(c, "bidir_sample_mult", p->bidir_sample_mult);
  set_i(c, "units", p->units);
  set_f(c, "sensor_width", p->sensor_width);
  set_b(c, "enable_dof", p->enable_dof != 0);
  set_f(c, "fstop", p->fstop);
  set_f(c, "focus_dist", p->focus_dist);
  set_i(c, "aperture_blades_lentil", p->aperture_blades_lentil);
  set_f(c, "exp", p->exp);
  set_i(c, "lens_model", p->lens_model);
  set_f(c, "wavelength", p->wavelength);
  set_f(c, "extra_sensor_shift", p->extra_sensor_shift);
  set_f(c, "focal_length_lentil", p->focal_length_lentil);
  set_f(c, "optical_vignetting", p->optical_vignetting);
  set_f(c, "abb_spherical", p->abb_spherical);
  set_f(c, "abb_distortion", p->abb_distortion);
  set_f(c, "abb_coma", p->abb_coma);
  set_f(c, "abb_chromatic", p->abb_chromatic);
  set_i(c, "abb_chromatic_type", p->abb_chromatic_type);
  set_f(c, "bokeh_circle_to_square", p->bokeh_circle_to_square);
  set_f(c, "bokeh_anamorphic", p->bokeh_anamorphic);
  set_b(c, "bokeh_enable_image", p->bokeh_enable_image != 0);
  set_s(c, "bokeh_image_path", "shim://bokeh");
  set_i(c, "vignetting_retries", p->vignetting_retries);
  set_f(c, "bidir_add_energy", p->bidir_add_energy);
  set_f(c, "bidir_add_energy_minimum_luminance", p->bidir_add_energy_minimum_luminance);
  set_f(c, "bidir_add_energy_transition", p->bidir_add_energy_transition);
  set_b(c, "enable_bidir_transmission", p->enable_bidir_transmission != 0);
  set_b(c, "enable_skydome", p->enable_skydome != 0);
}

// what lentil_operator.cpp:25-171 leaves in OperatorData for the camera to copy (lentil.h:988-1012)
void build_operator_aovs(ref_camera *r) {
  r->opdata.aovs.clear();
  for (size_t i = 0; i < r->aovs.size(); ++i) {
    if (r->aovs[i].filter == LB_FILTER_CRYPTO) continue;  // those arrive through options.outputs (setup_crypto_aovs, lentil.h:1015-1052)
    std::string out = std::string(r->aovs[i].name) + " RGBA lentil_replaced_filter shim_driver";
    AOVData a(&r->uni, out);
    a.original_filter = AtString(r->aovs[i].filter == LB_FILTER_CLOSEST ? "closest_filter" : "gaussian_filter");
    a.index = (int)r->opdata.aovs.size();
    r->opdata.aovs.push_back(a);
  }
  r->opdata.aovcount = (int)r->opdata.aovs.size();
}

void update_camera(ref_camera *r) {
  shim_default_universe() = &r->uni;
  AtNode &o = r->options;
  set_i(o, "xres", r->frame.xres_without_region);
  set_i(o, "yres", r->frame.yres_without_region);
  // setup_filter computes xres = region_max - region_min + 1 (lentil.h:1079-1080)
  set_i(o, "region_min_x", r->frame.region_min_x);
  set_i(o, "region_min_y", r->frame.region_min_y);
  set_i(o, "region_max_x", r->frame.region_min_x + r->frame.xres - 1);
  set_i(o, "region_max_y", r->frame.region_min_y + r->frame.yres - 1);
  set_i(o, "AA_samples", r->aa);
  set_b(o, "enable_adaptive_sampling", false);
  set_b(o, "ignore_dof", false);
  set_b(o, "enable_progressive_render", false);
  set_f(o, "meters_per_unit", 0.01f);
  if (!r->use_operator) build_operator_aovs(r);
  // cryptomatte: a `cryptomatte` AOV shader whose setup has completed (lentil.h:244-270) and ranked crypto outputs
  // written by an EXR driver (lentil.h:1104-1114)
  r->aov_shaders.ptrs.clear();
  r->outputs.strs.clear();
  bool any_crypto = false;
  for (auto &a : r->aovs) {
    if (a.filter == LB_FILTER_CRYPTO) { r->outputs.strs.push_back(std::string(a.name) + " FLOAT cryptomatte_filter shim_exr"); any_crypto = true; }
    else if (r->use_operator)  // the scene as exported: every output still on its original filter node
      r->outputs.strs.push_back(std::string(a.name) + " RGBA " + (a.filter == LB_FILTER_CLOSEST ? "shim_closest" : "shim_gaussian") + " shim_driver");
  }
  if (any_crypto) r->aov_shaders.ptrs.push_back(&r->crypto);
  AiNodeSetArray(&o, AtString("aov_shaders"), &r->aov_shaders);
  AiNodeSetArray(&o, AtString("outputs"), &r->outputs);
  if (r->use_operator) {  // operator_init + operator_cook of the node table NodeLoader hands out, on a fresh OperatorData per update
    if (r->op.local_data != &r->opdata) LentilOperatorMtd->OperatorCleanup(&r->op, r->op_user);
    while (r->uni.nodes.size() > r->scene_nodes) r->uni.nodes.pop_back();  // nodes made by the previous cook
    r->uni.created.clear();
    LentilOperatorMtd->OperatorInit(&r->op, &r->op_user);
    LentilOperatorMtd->OperatorCook(&r->op, &r->op, r->op_user, nullptr, nullptr);
  }
  lentilMethods->Update(&r->uni.session, &r->camera);
  r->cam = (Camera *)AiNodeGetLocalData(&r->camera);
}
}  // namespace

extern "C" {

int ref_camera_create(const lb_camera_params *p, const lb_bokeh_image *img, ref_camera **out) {
  ref_camera *r = new ref_camera();
  r->uni.options = &r->options;
  r->uni.camera = &r->camera;
  r->uni.nodes = {&r->options, &r->camera, &r->filter, &r->imager, &r->op, &r->crypto, &r->exr, &r->f_gauss, &r->f_closest, &r->f_crypto};
  r->scene_nodes = r->uni.nodes.size();
  r->f_gauss.name = "shim_gaussian"; r->f_gauss.entry.name = "gaussian_filter";
  r->f_closest.name = "shim_closest"; r->f_closest.entry.name = "closest_filter";
  r->f_crypto.name = "cryptomatte_filter"; r->f_crypto.entry.name = "cryptomatte_filter";
  r->crypto.name = "cryptomatte_shader"; r->crypto.entry.name = "cryptomatte"; r->crypto.local_data = &r->cryptodata;
  r->exr.name = "shim_exr"; r->exr.entry.name = "driver_exr";
  r->uni.entry_counts["imager_denoiser_oidn"] = 1;  // -> filter_width 1.0 (lentil.h:1083-1088): no footprint overlap
  for (AtNode *n : r->uni.nodes) n->universe = &r->uni;
  r->options.name = "options"; r->options.entry.name = "options";
  r->camera.name = "lentil_cam"; r->camera.entry.name = "lentil_camera";
  r->filter.name = "lentil_replaced_filter"; r->filter.entry.name = "lentil_filter";
  r->imager.name = "imager_lentil"; r->imager.entry.name = "imager_lentil";
  r->op.name = "lentil_operator"; r->op.entry.name = "lentil_operator";
  r->op.local_data = &r->opdata;
  shim_default_universe() = &r->uni;
  if (img && img->pixels) {
    ShimTexture t;
    t.w = img->width; t.h = img->height; t.c = img->channels;
    t.px.assign(img->pixels, img->pixels + (size_t)t.w * t.h * t.c);
    shim_textures()["shim://bokeh"] = t;
  }
  apply_params(r, p);
  r->frame = lb_frame_desc{8, 8, 8, 8, 0, 0};
  lb_aov_desc rgba{};
  strcpy(rgba.name, "RGBA"); rgba.filter = LB_FILTER_GAUSSIAN; rgba.role = LB_AOV_RGBA;
  r->aovs = {rgba};
  lentilMethods->PluginInitialize(nullptr);
  lentilMethods->Initialize(&r->uni.session, &r->camera);
  const int aborts = shim_abort_count();
  update_camera(r);
  if (shim_abort_count() != aborts) { *out = r; return LB_ERR_IMAGE; }
  *out = r;
  return LB_OK;
}
void ref_camera_destroy(ref_camera *r) {
  if (!r) return;
  shim_default_universe() = &r->uni;
  lentilMethods->Finish(&r->camera);
  if (r->op.local_data != &r->opdata) LentilOperatorMtd->OperatorCleanup(&r->op, r->op_user);
  delete r;
}
// From the next ref_filter_begin on, the camera's AOV list is what the library's own lentil_operator node cooks from the scene's
// outputs (every AOV on its original gaussian / closest filter node): the user's AOVs followed by lentil_debug, lentil_time and
// lentil_raydir (lentil_operator.cpp:103-128).  Off: the harness writes the list itself (build_operator_aovs).
int ref_camera_use_operator(ref_camera *r, int on) {
  if (!on && r->op.local_data != &r->opdata) {
    LentilOperatorMtd->OperatorCleanup(&r->op, r->op_user);
    r->op.local_data = &r->opdata;
  }
  r->use_operator = on != 0;
  return LB_OK;
}

int ref_camera_get_state(const ref_camera *r, lb_camera_state *s) {
#ifdef LB_ADAPTOR
  return lb_camera_get_state(r->cam->cam, s);
#else
  const Camera *c = r->cam;
  memset(s, 0, sizeof *s);
  s->aperture_radius = c->aperture_radius; s->sensor_shift = c->sensor_shift; s->tan_fov = c->tan_fov;
  s->focus_distance = c->focus_distance; s->lambda = c->lambda;
  s->lens_outer_pupil_radius = c->lens_outer_pupil_radius; s->lens_inner_pupil_radius = c->lens_inner_pupil_radius;
  s->lens_length = c->lens_length; s->lens_back_focal_length = c->lens_back_focal_length;
  s->lens_effective_focal_length = c->lens_effective_focal_length; s->lens_aperture_pos = c->lens_aperture_pos;
  s->lens_aperture_housing_radius = c->lens_aperture_housing_radius;
  s->lens_inner_pupil_curvature_radius = c->lens_inner_pupil_curvature_radius;
  s->lens_outer_pupil_curvature_radius = c->lens_outer_pupil_curvature_radius;
  s->lens_field_of_view = c->lens_field_of_view; s->lens_fstop = c->lens_fstop;
  s->lens_aperture_radius_at_fstop = c->lens_aperture_radius_at_fstop;
  return LB_OK;
#endif
}
int ref_camera_set_pupil_geometry(ref_camera *r, int outer, int inner) {
#ifdef LB_ADAPTOR
  return lb_camera_set_pupil_geometry(r->cam->cam, outer, inner);
#else
  const char *names[3] = {"spherical", "cyl-y", "cyl-x"};
  r->cam->lens_outer_pupil_geometry = names[outer];
  r->cam->lens_inner_pupil_geometry = names[inner];
  return LB_OK;
#endif
}
int ref_camera_set_state(ref_camera *r, double aperture_radius, double sensor_shift) {
#ifdef LB_ADAPTOR
  return lb_camera_set_state(r->cam->cam, aperture_radius, sensor_shift);
#else
  r->cam->aperture_radius = aperture_radius;
  r->cam->sensor_shift = sensor_shift;
  return LB_OK;
#endif
}

// camera_create_ray per sample.  nthreads > 1 calls it concurrently on the shared Camera the way Arnold's
// render threads do (the retry RNG is then the reference's racy process-global xor128, global.h:22-27).
int ref_camera_create_rays(ref_camera *r, size_t n, uint64_t, const lb_ray_in *in, const lb_ray_out *out, int nthreads) {
  shim_default_universe() = &r->uni;
  float *dst[7] = {out->origin, out->dir, out->dOdx, out->dOdy, out->dDdx, out->dDdy, out->weight};
#ifdef LB_ADAPTOR
  {  // what a host that knows its upcoming camera samples does at bucket entry; CreateRay below is per sample as ever
    std::vector<AtCameraInput> all(n);
    for (size_t i = 0; i < n; ++i) all[i] = AtCameraInput{in->sx[i], in->sy[i], in->dsx[i], in->dsy[i], in->lensx[i], in->lensy[i], 0.f};
    if (!getenv("LB_ADAPTOR_NO_PREFETCH") && lentil_b200_prefetch_rays(&r->camera, n, all.data()) != LB_OK) return LB_ERR_CUDA;
  }
#endif
  auto work = [&](size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
      AtCameraInput ci{in->sx[i], in->sy[i], in->dsx[i], in->dsy[i], in->lensx[i], in->lensy[i], 0.f};
      AtCameraOutput co;
      lentilMethods->CreateRay(&r->camera, ci, co, 0);
      const float v[7][3] = {{co.origin.x, co.origin.y, co.origin.z}, {co.dir.x, co.dir.y, co.dir.z}, {co.dOdx.x, co.dOdx.y, co.dOdx.z},
                             {co.dOdy.x, co.dOdy.y, co.dOdy.z},       {co.dDdx.x, co.dDdx.y, co.dDdx.z}, {co.dDdy.x, co.dDdy.y, co.dDdy.z},
                             {co.weight.r, co.weight.g, co.weight.b}};
      for (int k = 0; k < 7; ++k)
        if (dst[k]) { dst[k][i] = v[k][0]; dst[k][n + i] = v[k][1]; dst[k][2 * n + i] = v[k][2]; }
      if (out->tries) out->tries[i] = -1;  // not observable through the reference's interface
    }
  };
  if (nthreads <= 1) { work(0, n); return LB_OK; }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work, n * t / nthreads, n * (t + 1) / nthreads);
  for (auto &t : th) t.join();
  return LB_OK;
}

int ref_filter_begin(ref_camera *r, const lb_frame_desc *f, int n_aov, const lb_aov_desc *aovs, int aa_samples) {
  r->frame = *f;
  r->aovs.assign(aovs, aovs + n_aov);
  r->aa = aa_samples;
  r->n_user_aovs = r->aovs.size();
  update_camera(r);  // node_update -> setup_camera -> setup_lentil_aovs/setup_filter (lentil.h:211-281)
  if (r->use_operator) {  // the operator's helper outputs become AOVs n, n+1, n+2 of ref_imager_resolve / ref_filter_buffers
    const struct { const char *name; int filter, role; } extra[3] = {{"lentil_debug", LB_FILTER_CLOSEST, LB_AOV_LENTIL_DEBUG},
                                                                     {"lentil_time", LB_FILTER_GAUSSIAN, LB_AOV_PLAIN},
                                                                     {"lentil_raydir", LB_FILTER_GAUSSIAN, LB_AOV_PLAIN}};
    for (auto &e : extra) {
      lb_aov_desc d{};
      strcpy(d.name, e.name); d.filter = e.filter; d.role = e.role;
      r->aovs.push_back(d);
    }
  }
  return r->cam->redistribution ? LB_OK : LB_ERR_STATE;
}

// filter_pixel for every run of consecutive samples that share (px, py); the reference derives the inverse
// sample density from the number of samples it is handed (lentil_filter.cpp:79-87)
int ref_filter_accumulate(ref_camera *r, const lb_samples *S, int nthreads) {
  shim_default_universe() = &r->uni;
  r->camera.world_to_camera = AtMatrix();
  if (S->world_to_camera) memcpy(r->camera.world_to_camera.data, S->world_to_camera, 16 * sizeof(float));
  const size_t n = S->n;
  std::vector<float> Z(4 * n), P(4 * n), vol(4 * n, 0.f), ign(4 * n, 0.f), zero(4 * n, 0.f);
  for (size_t i = 0; i < n; ++i) {
    P[4 * i] = S->pos_cs[4 * i]; P[4 * i + 1] = S->pos_cs[4 * i + 1]; P[4 * i + 2] = S->pos_cs[4 * i + 2];
    Z[4 * i] = S->pos_cs[4 * i + 3];
    if (S->flags) { vol[4 * i] = (S->flags[i] & LB_SAMPLE_VOLUME) ? 1.f : 0.f; ign[4 * i] = (S->flags[i] & LB_SAMPLE_IGNORE) ? 1.f : 0.f; }
  }
  AtAOVSampleIterator it;
  it.aov_name = AtString("RGBA");
  it.rgba = S->rgba;
  it.aovs["P"] = {P.data(), 3};
  it.aovs["Z"] = {Z.data(), 1};
  it.aovs["volume"] = {vol.data(), 3};
  // The filter REQUIRES "FLOAT lentil_bidir_ignore" (lentil_filter.cpp:24) but READS atstring_lentil_ignore =
  // "lentil_ignore" (lentil.h:184, lentil_filter.cpp:162).  LB_SAMPLE_IGNORE stands for the value read at :162, so the
  // harness serves it under the name that line asks for (and under the required name, which nothing reads).
  it.aovs["lentil_ignore"] = {ign.data(), 1};
  it.aovs["lentil_bidir_ignore"] = {ign.data(), 1};
  it.aovs["lentil_time"] = {zero.data(), 1};
  it.aovs["lentil_raydir"] = {S->raydir ? S->raydir : zero.data(), 3};
  it.aovs["transmission"] = {S->transmission ? S->transmission : zero.data(), 4};
  for (size_t a = 0; a < r->n_user_aovs; ++a) {
    if (r->aovs[a].filter == LB_FILTER_CRYPTO) {
      it.depth_ids[r->aovs[a].name] = (S->crypto_ids && S->crypto_ids[a]) ? S->crypto_ids[a] : nullptr;
      continue;
    }
    const float *v = (S->aov_values && S->aov_values[a]) ? S->aov_values[a] : S->rgba;
    it.aovs[r->aovs[a].name] = {v, 4};
  }
  it.depth_n = S->crypto_depth;
  it.depth_count = S->crypto_count;
  it.depth_opacity = S->crypto_opacity;
  // runs of consecutive samples sharing a pixel
  std::vector<std::pair<size_t, size_t>> runs;
  for (size_t b = 0; b < n;) {
    size_t e = b + 1;
    while (e < n && S->px[e] == S->px[b] && S->py[e] == S->py[b]) ++e;
    runs.push_back({b, e});
    b = e;
  }
  auto work = [&](size_t lo, size_t hi) {
    AtAOVSampleIterator local = it;  // per-thread cursor over the shared arrays
    AtRGBA data_out;
    for (size_t k = lo; k < hi; ++k) {
      local.px = S->px[runs[k].first] + r->frame.region_min_x;  // filter_pixel subtracts region_min again (:99-100)
      local.py = S->py[runs[k].first] + r->frame.region_min_y;
      local.begin = runs[k].first; local.end = runs[k].second; local.cur = -1;
      LentilFilterDataMtd->FilterPixel(&r->filter, &local, &data_out, AI_TYPE_RGBA);
    }
  };
  // nthreads > 1: Arnold's bucket threads share the Camera's framebuffers without locks (lentil.h:828-829);
  // only used for timing, never for parity
  if (nthreads <= 1) work(0, runs.size());
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) th.emplace_back(work, runs.size() * t / nthreads, runs.size() * (t + 1) / nthreads);
    for (auto &t : th) t.join();
  }
#ifdef LB_ADAPTOR
  lentil_b200_flush(&r->camera);  // the threads' partial batches (the imager flushes them as well)
#endif
  return r->cam->redistribution ? LB_OK : LB_ERR_STATE;
}

int ref_imager_resolve(ref_camera *r, int aov, int x0, int y0, int w, int h, float *rgba_out) {
  shim_default_universe() = &r->uni;
  AtOutputIterator oit;
  oit.outs.push_back({AtString(r->aovs[aov].name), AI_TYPE_RGBA, rgba_out});
  LentilImagerMtd->DriverProcessBucket(&r->imager, &oit, nullptr, x0, y0, w, h, 0);
  return LB_OK;
}

// ---- the operator (lentil_operator.cpp:19-191) and the plugin entry point (lentil_loader.cpp:20-28) -----------------------
// NodeLoader's answers for i = 0, 1, ... as text: `<i> <name> type=<node_type> out=<output_type> version=<v> methods=<which table>`
int ref_node_loader(char *dump, size_t cap) {
  std::string out;
  int i = 0;
  for (;; ++i) {
    AtNodeLib lib;
    memset(&lib, 0, sizeof lib);
    if (!NodeLoader(i, &lib)) break;
    const char *table = lib.methods == lentilMethods ? "camera" : lib.methods == LentilFilterDataMtd ? "filter"
                      : lib.methods == LentilImagerMtd ? "imager" : lib.methods == LentilOperatorMtd ? "operator" : "?";
    out += std::to_string(i) + " " + (lib.name ? lib.name : "") + " type=" + std::to_string(lib.node_type) + " out=" + std::to_string((int)lib.output_type) +
           " version=" + lib.version + " methods=" + table + "\n";
    if (i > 64) break;
  }
  if (dump && cap) { strncpy(dump, out.c_str(), cap - 1); dump[cap - 1] = 0; }
  return i;
}

// What each node of the plugin declares to the renderer, as text: parameter declarations (type, name, default, enum strings) and
// metadata from node_parameters; for the filter node the AOVs node_initialize requires, the width node_update sets and
// filter_output_type over the data types; for the imager the render hints node_update sets.
int ref_node_interface(char *dump, size_t cap) {
  std::string out;
  for (int i = 0; i < 64; ++i) {
    AtNodeLib lib;
    memset(&lib, 0, sizeof lib);
    if (!NodeLoader(i, &lib)) break;
    out += std::string("node ") + lib.name + "\n";
    AtList list;
    AtNodeEntry entry;
    entry.name = lib.name;
    if (lib.methods->Parameters) lib.methods->Parameters(&list, &entry);
    for (auto &d : list.decls) out += "  param " + d + "\n";
    for (auto &m : entry.meta) out += "  meta " + m + "\n";
    AtUniverse uni;
    AtNode options, node;
    options.name = "options"; options.entry.name = "options";
    node.name = lib.name; node.entry.name = lib.name;
    uni.options = &options; uni.nodes = {&options, &node};
    options.universe = node.universe = &uni;
    shim_default_universe() = &uni;
    if (lib.node_type == AI_NODE_FILTER) {
      lib.methods->Initialize(&uni.session, &node);
      for (auto &a : node.required_aovs) out += "  requires " + a + "\n";
      for (int oidn = 0; oidn < 2; ++oidn) {  // an OIDN denoiser imager in the scene narrows the filter (lentil_filter.cpp:33-38)
        uni.entry_counts["imager_denoiser_oidn"] = oidn;
        lib.methods->Update(&uni.session, &node);
        out += "  filter_width oidn=" + std::to_string(oidn) + " " + shim_flt(node.filter_width) + "\n";
      }
      for (int t : {AI_TYPE_BYTE, AI_TYPE_INT, AI_TYPE_UINT, AI_TYPE_BOOLEAN, AI_TYPE_FLOAT, AI_TYPE_RGB, AI_TYPE_RGBA, AI_TYPE_VECTOR, AI_TYPE_VECTOR2,
                    AI_TYPE_STRING, AI_TYPE_POINTER, AI_TYPE_NODE, AI_TYPE_ARRAY, AI_TYPE_MATRIX})
        out += "  output_type " + std::to_string(t) + " -> " + std::to_string((int)lib.methods->FilterOutputType(&node, (uint8_t)t)) + "\n";
      lib.methods->Finish(&node);
    } else if (lib.node_type == AI_NODE_DRIVER) {
      lib.methods->Initialize(&uni.session, &node);
      lib.methods->Update(&uni.session, &node);
      for (auto &h : uni.session.hints) out += "  hint " + h + "\n";
      lib.methods->Finish(&node);
    }
    shim_default_universe() = nullptr;
  }
  if (dump && cap) { strncpy(dump, out.c_str(), cap - 1); dump[cap - 1] = 0; }
  return out.size() < cap ? LB_OK : LB_ERR_INVALID;
}

namespace {
struct ShimScene {  // a universe described line by line (ref_operator_cook)
  AtUniverse uni;
  std::deque<AtNode> nodes;
  AtNode *options = nullptr, *camera = nullptr, *op = nullptr;
  AtArray outputs, aov_shaders;
  AtNode *add(const std::string &name, const std::string &entry) {
    nodes.emplace_back();
    AtNode *n = &nodes.back();
    n->name = name; n->entry.name = entry; n->universe = &uni;
    uni.nodes.push_back(n);
    return n;
  }
};
const AtNodeMethods *operator_table() {  // through the plugin's own entry point, as Arnold finds it
  for (int i = 0; i < 64; ++i) {
    AtNodeLib lib;
    memset(&lib, 0, sizeof lib);
    if (!NodeLoader(i, &lib)) break;
    if (lib.name && std::string(lib.name) == "lentil_operator" && lib.node_type == AI_NODE_OPERATOR) return lib.methods;
  }
  return nullptr;
}
std::string dump_aov(const AOVData &a, size_t i) {
  return "aov " + std::to_string(i) + " name=" + a.name.c_str() + " type=" + std::to_string(a.type) + " original_filter=" + a.original_filter.c_str() +
         " duplicate=" + std::to_string((int)a.is_duplicate) + " crypto=" + std::to_string((int)a.is_crypto) + " index=" + std::to_string(a.index) +
         " tokens=[" + a.to.camera_tok + "|" + a.to.aov_name_tok + "|" + a.to.aov_type_tok + "|" + a.to.filter_tok + "|" + a.to.driver_tok + "|" +
         (a.to.half_flag ? "HALF" : "") + "] driver=" + (a.to.get_driver() ? AiNodeGetName(a.to.get_driver()) : "-") + " output=\"" + a.to.rebuild_output() + "\"\n";
}
}  // namespace

// Runs operator_init + operator_cook (`cooks` times, as a re-cooked scene does) on a scene given as text, one item per line:
//   node <name> <entry>        a node of the scene (filters, drivers, ...)
//   camera <entry>             node entry of the active camera (default lentil_camera)
//   aov_shader <name>          a declared node already listed in options.aov_shaders
//   output <output string>     one element of options.outputs
// and dumps what is left: cook's return values, OperatorData.aovs, the nodes of the universe with their string parameters and
// links, options.aov_shaders, and the output list + AOV registrations after the camera's sanitize / rebuild step
// (aov_data.h:166-189, lentil.h:1046-1058).
int ref_operator_cook(const char *scene_text, int cooks, char *dump, size_t cap) {
  const AtNodeMethods *mt = operator_table();
  if (!mt || !mt->OperatorInit || !mt->OperatorCook || !mt->OperatorCleanup) return LB_ERR_STATE;
  ShimScene sc;
  sc.options = sc.add("options", "options");
  std::string camera_entry = "lentil_camera";
  std::vector<std::pair<std::string, std::string>> decl;
  std::vector<std::string> outs, shader_names;
  {
    std::string text(scene_text ? scene_text : ""), line;
    size_t at = 0;
    while (at <= text.size()) {
      const size_t nl = text.find('\n', at);
      line = text.substr(at, nl == std::string::npos ? std::string::npos : nl - at);
      at = nl == std::string::npos ? text.size() + 1 : nl + 1;
      if (line.rfind("node ", 0) == 0) {
        const size_t sp = line.find(' ', 5);
        if (sp == std::string::npos) return LB_ERR_INVALID;
        decl.push_back({line.substr(5, sp - 5), line.substr(sp + 1)});
      } else if (line.rfind("camera ", 0) == 0) camera_entry = line.substr(7);
      else if (line.rfind("aov_shader ", 0) == 0) shader_names.push_back(line.substr(11));
      else if (line.rfind("output ", 0) == 0) outs.push_back(line.substr(7));
      else if (!line.empty()) return LB_ERR_INVALID;
    }
  }
  sc.camera = sc.add("scene_camera", camera_entry);
  for (auto &d : decl) sc.add(d.first, d.second);
  sc.op = sc.add("lentil_operator", "lentil_operator");
  sc.uni.options = sc.options;
  sc.uni.camera = sc.camera;
  shim_default_universe() = &sc.uni;
  sc.outputs.strs = outs;
  for (auto &nm : shader_names) sc.aov_shaders.ptrs.push_back(AiNodeLookUpByName(&sc.uni, AtString(nm.c_str())));
  AiNodeSetArray(sc.options, AtString("outputs"), &sc.outputs);
  AiNodeSetArray(sc.options, AtString("aov_shaders"), &sc.aov_shaders);

  std::string out;
  void *user = nullptr;
  out += std::string("init=") + (mt->OperatorInit(sc.op, &user) ? "1" : "0") + "\n";
  for (int k = 0; k < cooks; ++k) out += "cook" + std::to_string(k) + "=" + (mt->OperatorCook(sc.op, sc.op, user, nullptr, nullptr) ? "1" : "0") + "\n";
  if (mt->OperatorPostCook) out += std::string("post_cook=") + (mt->OperatorPostCook(sc.op, user) ? "1" : "0") + "\n";
  OperatorData *od = (OperatorData *)AiNodeGetLocalData(sc.op);
  for (size_t i = 0; i < od->aovs.size(); ++i) out += dump_aov(od->aovs[i], i);
  for (AtNode *n : sc.uni.nodes) {
    out += "node " + n->name + " entry=" + n->entry.name;
    for (auto &kv : n->params) if (!kv.second.s.empty()) out += " " + kv.first + "=" + kv.second.s;
    for (auto &kv : n->links) out += " " + kv.first + "<-" + kv.second->name;
    out += "\n";
  }
  AtArray *sh = AiNodeGetArray(sc.options, AtString("aov_shaders"));
  out += "aov_shaders";
  for (uint32_t i = 0; i < AiArrayGetNumElements(sh); ++i) out += std::string(" ") + (AiArrayGetPtr(sh, i) ? AiNodeGetName((AtNode *)AiArrayGetPtr(sh, i)) : "null");
  out += "\n";
  // what the camera makes of the list (setup_lentil_aovs + rebuild + sanitize, lentil.h:1008-1011,1046-1058)
  std::vector<AOVData> list(od->aovs.begin(), od->aovs.end());
  rebuild_arnold_outputs_from_list(&sc.uni, list);
  AtArray *final_outputs = AiNodeGetArray(sc.options, AtString("outputs"));
  for (uint32_t i = 0; i < AiArrayGetNumElements(final_outputs); ++i) out += std::string("final_output ") + AiArrayGetStr(final_outputs, i).c_str() + "\n";
  for (auto &nm : sc.uni.registered_aovs) out += "registered " + nm + "\n";
  sanitize_aov_list(list);
  for (size_t i = 0; i < list.size(); ++i) out += "framebuffer " + std::to_string(i) + " " + list[i].name.c_str() + " " + list[i].original_filter.c_str() + "\n";
  if (final_outputs != &sc.outputs) delete final_outputs;
  mt->OperatorCleanup(sc.op, user);
  shim_default_universe() = nullptr;
  if (dump && cap) { strncpy(dump, out.c_str(), cap - 1); dump[cap - 1] = 0; }
  return (int)out.size() < (int)cap ? LB_OK : LB_ERR_INVALID;
}

#ifdef LB_ADAPTOR
int ref_filter_buffers(ref_camera *r, int aov, float **buffer, float **weight) {
  const int a = (aov >= 0 && aov < (int)r->aovs.size()) ? lentil_b200_aov_index(&r->camera, r->aovs[aov].name) : -1;
  if (a < 0) return LB_ERR_INVALID;
  lentil_b200_flush(&r->camera);
  const size_t npx = (size_t)r->frame.xres * r->frame.yres;
  r->host_buffer.resize(npx * 4);
  r->host_weight.resize(npx);
  const int rc = lb_filter_buffers_host(r->cam->cam, a, r->host_buffer.data(), r->host_weight.data());
  if (buffer) *buffer = r->host_buffer.data();
  if (weight) *weight = r->host_weight.data();
  return rc;
}
// AOVData::crypto_hash_map as the reference harness dumps it: ids ascending (std::map order), free slots NaN
int ref_filter_crypto(ref_camera *r, int aov, int slots, float *ids_out, float *weights_out, float *total_out) {
  const int a = (aov >= 0 && aov < (int)r->aovs.size()) ? lentil_b200_aov_index(&r->camera, r->aovs[aov].name) : -1;
  if (a < 0) return -1;
  lentil_b200_flush(&r->camera);
  int k = 0;
  if (lb_filter_crypto_host(r->cam->cam, a, nullptr, nullptr, &k) != LB_OK) return -1;
  const size_t npx = (size_t)r->frame.xres * r->frame.yres;
  std::vector<float> ids(npx * k), wts(npx * k), plane(npx * 4);
  if (lb_filter_crypto_host(r->cam->cam, a, ids.data(), wts.data(), nullptr) != LB_OK) return -1;
  if (lb_filter_buffers_host(r->cam->cam, a, plane.data(), nullptr) != LB_OK) return -1;
  size_t mx = 0;
  for (size_t p = 0; p < npx; ++p) {
    std::vector<std::pair<float, float>> e;
    for (int j = 0; j < k; ++j) {
      uint32_t bits;
      memcpy(&bits, &ids[p * k + j], 4);
      if (bits != 0xFFFFFFFFu) e.push_back({ids[p * k + j], wts[p * k + j]});
    }
    std::sort(e.begin(), e.end(), [](const std::pair<float, float> &x, const std::pair<float, float> &y) { return x.first < y.first; });
    mx = std::max(mx, e.size());
    if (total_out) total_out[p] = plane[4 * p];
    for (int j = 0; j < slots; ++j) {
      const bool on = j < (int)e.size();
      const uint32_t nanbits = 0xFFFFFFFFu;
      if (ids_out) { if (on) ids_out[p * slots + j] = e[j].first; else memcpy(&ids_out[p * slots + j], &nanbits, 4); }
      if (weights_out) weights_out[p * slots + j] = on ? e[j].second : 0.0f;
    }
  }
  return (int)mx;
}
}  // extern "C"
#else
static AOVData *find_aov(ref_camera *r, int aov) {
  if (aov < 0 || aov >= (int)r->aovs.size()) return nullptr;
  for (auto &a : r->cam->aovs) if (a.name == AtString(r->aovs[aov].name)) return &a;
  return nullptr;
}

int ref_filter_buffers(ref_camera *r, int aov, float **buffer, float **weight) {
  AOVData *a = find_aov(r, aov);
  if (!a || a->buffer.empty()) return LB_ERR_INVALID;
  if (buffer) *buffer = &a->buffer[0].r;
  if (weight) *weight = r->cam->filter_weight_buffer.data();
  return LB_OK;
}

// AOVData::crypto_hash_map / crypto_total_weight dumped like orc_filter_crypto
int ref_filter_crypto(ref_camera *r, int aov, int slots, float *ids_out, float *weights_out, float *total_out) {
  AOVData *a = find_aov(r, aov);
  if (!a || a->crypto_hash_map.empty()) return -1;
  size_t mx = 0;
  for (size_t p = 0; p < a->crypto_hash_map.size(); ++p) {
    mx = std::max(mx, a->crypto_hash_map[p].size());
    if (total_out) total_out[p] = a->crypto_total_weight[p];
    int k = 0;
    for (auto const &kv : a->crypto_hash_map[p]) {
      if (k >= slots) break;
      if (ids_out) ids_out[p * slots + k] = kv.first;
      if (weights_out) weights_out[p * slots + k] = kv.second;
      ++k;
    }
    for (; k < slots; ++k) {
      if (ids_out) { const uint32_t nanbits = 0xFFFFFFFFu; memcpy(&ids_out[p * slots + k], &nanbits, 4); }
      if (weights_out) weights_out[p * slots + k] = 0.0f;
    }
  }
  return (int)mx;
}

// ---- primitives, straight from the reference headers -------------------------------------------------------
unsigned int ref_tea8(unsigned int v0, unsigned int v1) { return tea<8>(v0, v1); }
float ref_rng(unsigned int *state) { return rng(*state); }
void ref_xor128_seq(uint32_t *out, int n) { for (int i = 0; i < n; ++i) out[i] = xor128(); }  // continues the process-global state
float ref_fast_sin(float x) { return fast_sin(x); }
float ref_fast_cos(float x) { return fast_cos(x); }
double ref_lens_ipow(double x, int e) { return lens_ipow(x, e); }
void ref_concentric_disk_sample(double ox, double oy, int fast_trigo, double out[2]) {
  Eigen::Vector2d d(0, 0);
  concentric_disk_sample(ox, oy, d, fast_trigo != 0);
  out[0] = d(0); out[1] = d(1);
}
void ref_sphereToCs(const double inpos[2], const double indir[2], double center, double R, double outpos[3], double outdir[3]) {
  Eigen::Vector3d p(0, 0, 0), d(0, 0, 0);
  sphereToCs(Eigen::Vector2d(inpos[0], inpos[1]), Eigen::Vector2d(indir[0], indir[1]), p, d, center, R);
  for (int k = 0; k < 3; ++k) { outpos[k] = p(k); outdir[k] = d(k); }
}
void ref_csToSphere(const double inpos[3], const double indir[3], double center, double R, double outpos[2], double outdir[2]) {
  Eigen::Vector2d p(0, 0), d(0, 0);
  csToSphere(Eigen::Vector3d(inpos[0], inpos[1], inpos[2]), Eigen::Vector3d(indir[0], indir[1], indir[2]), p, d, center, R);
  for (int k = 0; k < 2; ++k) { outpos[k] = p(k); outdir[k] = d(k); }
}
void ref_cylinderToCs(const double inpos[2], const double indir[2], double center, double R, int cyl_y, double outpos[3], double outdir[3]) {
  Eigen::Vector3d p(0, 0, 0), d(0, 0, 0);
  cylinderToCs(Eigen::Vector2d(inpos[0], inpos[1]), Eigen::Vector2d(indir[0], indir[1]), p, d, center, R, cyl_y != 0);
  for (int k = 0; k < 3; ++k) { outpos[k] = p(k); outdir[k] = d(k); }
}
void ref_csToCylinder(const double inpos[3], const double indir[3], double center, double R, int cyl_y, double outpos[2], double outdir[2]) {
  Eigen::Vector2d p(0, 0), d(0, 0);
  csToCylinder(Eigen::Vector3d(inpos[0], inpos[1], inpos[2]), Eigen::Vector3d(indir[0], indir[1], indir[2]), p, d, center, R, cyl_y != 0);
  for (int k = 0; k < 2; ++k) { outpos[k] = p(k); outdir[k] = d(k); }
}
int ref_logarithmic_values(double *out, int cap) {
  auto v = logarithmic_values();
  for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
  return (int)v.size();
}
void ref_line_plane_intersection(const double o[3], const double d[3], double out[3]) {
  Eigen::Vector3d r = line_plane_intersection(Eigen::Vector3d(o[0], o[1], o[2]), Eigen::Vector3d(d[0], d[1], d[2]));
  for (int k = 0; k < 3; ++k) out[k] = r(k);
}
void ref_bokeh_sample(ref_camera *r, float r_row, float r_col, double out[2]) {
  Eigen::Vector2d l(0, 0);
  r->cam->image.bokehSample(r_row, r_col, l, 0.f, 0.f);
  out[0] = l(0); out[1] = l(1);
}
double ref_lens_evaluate(ref_camera *r, const double in[5], double out[5]) {
  Eigen::VectorXd i(5), o(5);
  for (int k = 0; k < 5; ++k) { i(k) = in[k]; o(k) = out[k]; }
  double t = r->cam->lens_evaluate(i, o);
  for (int k = 0; k < 5; ++k) out[k] = o(k);
  return t;
}
void ref_lens_pt_sample_aperture(ref_camera *r, double in[5], double out[5], double dist) {
  Eigen::VectorXd i(5), o(5);
  for (int k = 0; k < 5; ++k) { i(k) = in[k]; o(k) = out[k]; }
  r->cam->lens_pt_sample_aperture(i, o, dist);
  for (int k = 0; k < 5; ++k) { in[k] = i(k); out[k] = o(k); }
}
double ref_lens_lt_sample_aperture(ref_camera *r, const double scene[3], const double ap[2], double sensor[5], double out[5], double lambda_) {
  Eigen::VectorXd s(5), o(5);
  for (int k = 0; k < 5; ++k) { s(k) = sensor[k]; o(k) = out[k]; }
  double t = r->cam->lens_lt_sample_aperture(Eigen::Vector3d(scene[0], scene[1], scene[2]), Eigen::Vector2d(ap[0], ap[1]), s, o, lambda_);
  for (int k = 0; k < 5; ++k) { sensor[k] = s(k); out[k] = o(k); }
  return t;
}
int ref_trace_ray_bw_po(ref_camera *r, const double target[3], int px, int py, int total_samples_taken, float lambda_in, double sensor_pos[2]) {
  Eigen::Vector2d s(0, 0);
  AtMatrix m;
  AtShaderGlobals sg;
  bool ok = r->cam->trace_ray_bw_po(Eigen::Vector3d(target[0], target[1], target[2]), s, px, py, total_samples_taken, m, AtVector(0, 0, 0), &sg, lambda_in, false);
  sensor_pos[0] = s(0); sensor_pos[1] = s(1);
  return ok ? 1 : 0;
}
float ref_get_coc_thinlens(ref_camera *r, float z) { return r->cam->get_coc_thinlens(AtVector(0.f, 0.f, z)); }

}  // extern "C"
#endif  // LB_ADAPTOR
