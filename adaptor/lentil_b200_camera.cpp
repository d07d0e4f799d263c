// lentil_b200_camera.cpp — the `lentil_camera` node over liblentil_b200.so.
// Replaces /root/reference/src/lentil_camera.cpp (node table, parameters :19-52, node_initialize/update/finish :56-75,
// camera_create_ray :78-125, camera_reverse_ray :164-172) and Camera::setup_camera (src/lentil.h:211-281).
#include "lentil_b200_adaptor.h"

#include <cmath>

AI_CAMERA_NODE_EXPORT_METHODS(lentilMethods)

static const char *Units[] = {"mm", "cm", "dm", "m", "automatic", NULL};
static const char *CameraTypes[] = {"ThinLens", "PolynomialOptics", NULL};
static const char *ChromaticTypes[] = {"green_magenta", "red_cyan", NULL};
static const char *LensModelNames[] = {
#include "../include/auto_generated_lens_includes/pota_cpp_lenses.h"
    NULL};

node_parameters {  // lentil_camera.cpp:19-52, same names and defaults
  (void)Units; (void)CameraTypes; (void)ChromaticTypes; (void)LensModelNames;
  AiParameterEnum("camera_type", 0, CameraTypes);
  AiParameterInt("bidir_sample_mult", 5);
  AiParameterEnum("units", 1, Units);
  AiParameterFlt("sensor_width", 36.0);
  AiParameterBool("enable_dof", true);
  AiParameterFlt("fstop", 0.0);
  AiParameterFlt("focus_dist", 150.0);
  AiParameterInt("aperture_blades_lentil", 0);
  AiParameterFlt("exp", 1.0);
  AiParameterEnum("lens_model", 16, LensModelNames);
  AiParameterFlt("wavelength", 550.0);
  AiParameterFlt("extra_sensor_shift", 0.0);
  AiParameterFlt("focal_length_lentil", 35.0);
  AiParameterFlt("optical_vignetting", 0.0);
  AiParameterFlt("abb_spherical", 0.5);
  AiParameterFlt("abb_distortion", 0.0);
  AiParameterFlt("abb_coma", 0.0);
  AiParameterFlt("abb_chromatic", 0.0);
  AiParameterEnum("abb_chromatic_type", 0, ChromaticTypes);
  AiParameterFlt("bokeh_circle_to_square", 0.0);
  AiParameterFlt("bokeh_anamorphic", 0.0);
  AiParameterBool("bokeh_enable_image", false);
  AiParameterStr("bokeh_image_path", "");
  AiParameterInt("vignetting_retries", 15);
  AiParameterFlt("bidir_add_energy", 0.0);
  AiParameterFlt("bidir_add_energy_minimum_luminance", 2.0);
  AiParameterFlt("bidir_add_energy_transition", 1.0);
  AiParameterBool("enable_bidir_transmission", false);
  AiParameterBool("enable_skydome", false);
  AiMetaDataSetBool(nentry, nullptr, "force_update", true);
}

node_plugin_initialize { return true; }
node_plugin_cleanup {}

node_initialize {
  AiCameraInitialize(node);
  AiNodeSetLocalData(node, new LbAdaptorCamera());
}

node_update {
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(node);
  lb_adaptor_setup(c, AiNodeGetUniverse(node));
  AiCameraUpdate(node, false);
}

node_finish { delete (LbAdaptorCamera *)AiNodeGetLocalData(node); }

static inline uint64_t ray_key(const AtCameraInput &in) {
  uint32_t a, b, c, d;
  memcpy(&a, &in.sx, 4); memcpy(&b, &in.sy, 4); memcpy(&c, &in.lensx, 4); memcpy(&d, &in.lensy, 4);
  uint64_t h = ((uint64_t)a << 32 | b) * 0x9E3779B97F4A7C15ull;
  h ^= ((uint64_t)c << 32 | d) + 0xD6E8FEB86659FD93ull + (h << 6) + (h >> 2);
  return h;
}

static void answer(const LbAdaptorCamera *c, size_t i, size_t n, AtCameraOutput &o) {
  const std::vector<float> *p = c->pre_out;
  o.origin = AtVector(p[0][i], p[0][n + i], p[0][2 * n + i]);
  o.dir = AtVector(p[1][i], p[1][n + i], p[1][2 * n + i]);
  o.dOdx = AtVector(p[2][i], p[2][n + i], p[2][2 * n + i]);
  o.dOdy = AtVector(p[3][i], p[3][n + i], p[3][2 * n + i]);
  o.dDdx = AtVector(p[4][i], p[4][n + i], p[4][2 * n + i]);
  o.dDdy = AtVector(p[5][i], p[5][n + i], p[5][2 * n + i]);
  o.weight = AtRGB(p[6][i], p[6][n + i], p[6][2 * n + i]);
}

camera_create_ray {  // lentil_camera.cpp:78-125
  (void)tid;
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(node);
  if (c->pre_n) {
    auto it = c->pre_index.find(ray_key(input));
    if (it != c->pre_index.end()) {
      const AtCameraInput &q = c->pre_in[it->second];
      if (q.sx == input.sx && q.sy == input.sy && q.lensx == input.lensx && q.lensy == input.lensy && q.dsx == input.dsx && q.dsy == input.dsy) {
        answer(c, it->second, c->pre_n, output);
        return;
      }
    }
  }
  // not prefetched: one ray through the host-buffer call (all latency; a renderer that can should prefetch per bucket)
  float out[7][3];
  lb_ray_in in{&input.sx, &input.sy, &input.dsx, &input.dsy, &input.lensx, &input.lensy};
  lb_ray_out ro{out[0], out[1], out[2], out[3], out[4], out[5], out[6], nullptr};
  if (lb_camera_create_rays_host(c->cam, 1, 0, &in, &ro) != LB_OK) {
    AiMsgError("[LENTIL B200] %s", lb_last_error());
    output.weight = AI_RGB_BLACK;  // in-band failure, as a ray that cannot be traced (lentil.h:379)
    return;
  }
  output.origin = AtVector(out[0][0], out[0][1], out[0][2]);
  output.dir = AtVector(out[1][0], out[1][1], out[1][2]);
  output.dOdx = AtVector(out[2][0], out[2][1], out[2][2]);
  output.dOdy = AtVector(out[3][0], out[3][1], out[3][2]);
  output.dDdx = AtVector(out[4][0], out[4][1], out[4][2]);
  output.dDdy = AtVector(out[5][0], out[5][1], out[5][2]);
  output.weight = AtRGB(out[6][0], out[6][1], out[6][2]);
}

camera_reverse_ray {  // lentil_camera.cpp:164-172: pinhole approximation, cheap enough to stay on the host
  (void)relative_time;
  const LbAdaptorCamera *c = (const LbAdaptorCamera *)AiNodeGetLocalData(node);
  lb_camera_state s;
  lb_camera_get_state(c->cam, &s);
  const double coeff = 1.0 / std::max(std::abs((double)Po.z * s.tan_fov), 1e-3);
  Ps.x = (float)(Po.x * coeff);
  Ps.y = (float)(Po.y * coeff);
  return true;
}

extern "C" {

int lentil_b200_prefetch_rays(AtNode *camera_node, size_t n, const AtCameraInput *inputs) {
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(camera_node);
  if (!c || !c->cam) return LB_ERR_STATE;
  std::lock_guard<std::mutex> lk(c->mu);
  c->pre_n = 0;
  c->pre_index.clear();
  c->pre_in.assign(inputs, inputs + n);
  if (n == 0) return LB_OK;
  std::vector<float> soa[6];
  for (auto &v : soa) v.resize(n);
  for (size_t i = 0; i < n; ++i) {
    soa[0][i] = inputs[i].sx; soa[1][i] = inputs[i].sy; soa[2][i] = inputs[i].dsx;
    soa[3][i] = inputs[i].dsy; soa[4][i] = inputs[i].lensx; soa[5][i] = inputs[i].lensy;
  }
  for (auto &v : c->pre_out) v.assign(3 * n, 0.f);
  lb_ray_in in{soa[0].data(), soa[1].data(), soa[2].data(), soa[3].data(), soa[4].data(), soa[5].data()};
  lb_ray_out out{c->pre_out[0].data(), c->pre_out[1].data(), c->pre_out[2].data(), c->pre_out[3].data(),
                 c->pre_out[4].data(), c->pre_out[5].data(), c->pre_out[6].data(), nullptr};
  const int rc = lb_camera_create_rays_host(c->cam, n, 0, &in, &out);
  if (rc != LB_OK) return rc;
  c->pre_index.reserve(n * 2);
  for (size_t i = 0; i < n; ++i) c->pre_index.emplace(ray_key(inputs[i]), (uint32_t)i);
  c->pre_n = n;
  return LB_OK;
}

lb_camera *lentil_b200_camera_handle(AtNode *camera_node) {
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(camera_node);
  return c ? c->cam : nullptr;
}
int lentil_b200_redistribution(AtNode *camera_node) {
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(camera_node);
  return c && c->redistribution;
}
int lentil_b200_flush(AtNode *camera_node) {
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(camera_node);
  return c ? lb_adaptor_flush_all(c) : LB_ERR_STATE;
}
int lentil_b200_aov_index(AtNode *camera_node, const char *aov_name) {
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(camera_node);
  if (!c) return -1;
  for (size_t a = 0; a < c->aovs.size(); ++a)
    if (c->aovs[a].name == AtString(aov_name)) return (int)a;
  return -1;
}

}  // extern "C"
