// lentil_b200_adaptor.cpp — setup and sample batching shared by the adaptor's nodes.
#include "lentil_b200_adaptor.h"

#include <algorithm>
#include <atomic>
#include <cmath>

void LbSampleBatch::clear() {
  px.clear(); py.clear(); rgba.clear(); pos.clear(); raydir.clear(); transmission.clear(); flags.clear();
  for (auto &v : values) v.clear();
  crypto_count.clear(); crypto_opacity.clear();
  for (auto &v : crypto_ids) v.clear();
  n = 0;
}

static std::atomic<uint64_t> g_next_uid{1};
LbAdaptorCamera::LbAdaptorCamera() : uid(g_next_uid.fetch_add(1)) {}

LbAdaptorCamera::~LbAdaptorCamera() {
  for (LbSampleBatch *b : batches) delete b;
  if (cam) lb_camera_destroy(cam);
}

LbSampleBatch &lb_adaptor_thread_batch(LbAdaptorCamera *c) {
  thread_local std::unordered_map<uint64_t, LbSampleBatch *> mine;  // batches are owned (and freed) by their camera
  LbSampleBatch *&b = mine[c->uid];
  if (!b) {
    b = new LbSampleBatch();
    std::lock_guard<std::mutex> lk(c->mu);
    c->batches.push_back(b);
  }
  if (b->values.size() != c->aovs.size()) {
    b->values.assign(c->aovs.size(), {});
    b->crypto_ids.assign(c->aovs.size(), {});
  }
  return *b;
}

int lb_adaptor_flush(LbAdaptorCamera *c, LbSampleBatch &b) {
  if (b.n == 0) return LB_OK;
  lb_samples s{};
  s.n = b.n;
  s.px = b.px.data(); s.py = b.py.data();
  s.rgba = b.rgba.data(); s.pos_cs = b.pos.data(); s.raydir = b.raydir.data(); s.transmission = b.transmission.data();
  s.flags = b.flags.data();
  std::vector<const float *> vals(c->aovs.size(), nullptr), ids(c->aovs.size(), nullptr);
  for (size_t a = 0; a < c->aovs.size(); ++a) {
    if (!b.values[a].empty()) vals[a] = b.values[a].data();
    if (!b.crypto_ids[a].empty()) ids[a] = b.crypto_ids[a].data();
  }
  s.aov_values = vals.data();
  s.inv_density = b.inv_density;
  if (c->has_crypto) {
    s.crypto_depth = LB_CRYPTO_MAX_DEPTH;
    s.crypto_count = b.crypto_count.data();
    s.crypto_opacity = b.crypto_opacity.data();
    s.crypto_ids = ids.data();
  }
  s.world_to_camera = b.w2c;
  const int rc = lb_filter_accumulate_host(c->cam, &s);
  if (rc != LB_OK) AiMsgError("[LENTIL B200] %s", lb_last_error());
  b.clear();
  return rc;
}

int lb_adaptor_flush_all(LbAdaptorCamera *c) {
  std::vector<LbSampleBatch *> all;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    all = c->batches;
  }
  int rc = LB_OK;
  for (LbSampleBatch *b : all) {
    const int r = lb_adaptor_flush(c, *b);
    if (r != LB_OK) rc = r;
  }
  return rc;
}

static float clamp_min_f(float v, float lo) { return v < lo ? lo : v; }

// Camera::setup_camera (lentil.h:211-281): parameters (get_lentil_camera_params :1189-1243), bokeh image
// (imagebokeh.h:83-107; the CDF is built inside lb_camera_create), bidirectional status (:1150-1172), AOV list from the
// operator and the cryptomatte outputs (setup_lentil_aovs :988-1012, setup_crypto_aovs :1015-1052), framebuffers
// (setup_filter :1056-1117 -> lb_filter_begin).
void lb_adaptor_setup(LbAdaptorCamera *c, AtUniverse *universe) {
  c->options_node = AiUniverseGetOptions(universe);
  c->camera_node = AiUniverseGetCamera(universe);
  AtNode *cn = c->camera_node, *on = c->options_node;
  lb_camera_params p;
  lb_camera_params_default(&p);
  p.camera_type = AiNodeGetInt(cn, AtString("camera_type"));
  p.units = AiNodeGetInt(cn, AtString("units"));
  if (p.units == 4) {  // "automatic" (lentil.h:1193-1199)
    const float mpu = AiNodeGetFlt(on, AtString("meters_per_unit"));
    if (mpu == 1.0f) p.units = LB_UNITS_M;
    else if (mpu == 0.1f) p.units = LB_UNITS_DM;
    else if (mpu == 0.01f) p.units = LB_UNITS_CM;
    else if (mpu == 0.001f) p.units = LB_UNITS_MM;
  }
  p.sensor_width = AiNodeGetFlt(cn, AtString("sensor_width"));
  p.enable_dof = AiNodeGetBool(cn, AtString("enable_dof"));
  if (AiNodeGetBool(on, AtString("ignore_dof"))) p.enable_dof = 0;
  p.fstop = AiNodeGetFlt(cn, AtString("fstop"));
  p.focus_dist = AiNodeGetFlt(cn, AtString("focus_dist"));
  p.aperture_blades_lentil = AiNodeGetInt(cn, AtString("aperture_blades_lentil"));
  p.exp = AiNodeGetFlt(cn, AtString("exp"));
  p.lens_model = AiNodeGetInt(cn, AtString("lens_model"));
  p.wavelength = AiNodeGetFlt(cn, AtString("wavelength"));
  p.extra_sensor_shift = AiNodeGetFlt(cn, AtString("extra_sensor_shift"));
  p.focal_length_lentil = AiNodeGetFlt(cn, AtString("focal_length_lentil"));
  p.optical_vignetting = AiNodeGetFlt(cn, AtString("optical_vignetting"));
  p.abb_spherical = AiNodeGetFlt(cn, AtString("abb_spherical"));
  p.abb_distortion = AiNodeGetFlt(cn, AtString("abb_distortion"));
  p.abb_coma = AiNodeGetFlt(cn, AtString("abb_coma"));
  p.abb_chromatic = AiNodeGetFlt(cn, AtString("abb_chromatic"));
  p.abb_chromatic_type = AiNodeGetInt(cn, AtString("abb_chromatic_type"));
  p.bokeh_circle_to_square = AiNodeGetFlt(cn, AtString("bokeh_circle_to_square"));
  p.bokeh_anamorphic = AiNodeGetFlt(cn, AtString("bokeh_anamorphic"));
  p.bokeh_enable_image = AiNodeGetBool(cn, AtString("bokeh_enable_image"));
  p.bidir_sample_mult = AiNodeGetInt(cn, AtString("bidir_sample_mult"));
  p.bidir_add_energy_minimum_luminance = AiNodeGetFlt(cn, AtString("bidir_add_energy_minimum_luminance"));
  p.bidir_add_energy = AiNodeGetFlt(cn, AtString("bidir_add_energy"));
  p.bidir_add_energy_transition = AiNodeGetFlt(cn, AtString("bidir_add_energy_transition"));
  p.vignetting_retries = AiNodeGetInt(cn, AtString("vignetting_retries"));
  p.enable_bidir_transmission = AiNodeGetBool(cn, AtString("enable_bidir_transmission"));
  p.enable_skydome = AiNodeGetBool(cn, AtString("enable_skydome"));
  (void)clamp_min_f;
  c->params = p;

  // bokeh image through the Arnold texture API (imagebokeh.h:83-107); pixels stay on the host, the library builds the CDF
  std::vector<float> pixels;
  lb_bokeh_image img{};
  if (p.bokeh_enable_image) {
    const AtString path = AiNodeGetStr(cn, AtString("bokeh_image_path"));
    unsigned iw = 0, ih = 0, nc = 0;
    bool ok = AiTextureGetResolution(path, &iw, &ih) && AiTextureGetNumChannels(path, &nc);
    if (ok) {
      pixels.resize((size_t)iw * ih * nc);
      ok = AiTextureLoad(path, false, 0, pixels.data());
    }
    if (!ok) {
      AiMsgError("[LENTIL CAMERA PO] Couldn't open bokeh image!");
      AiRenderAbort();
      return;
    }
    img.width = (int)iw; img.height = (int)ih; img.channels = (int)nc; img.pixels = pixels.data();
  }
  const int rc = c->cam ? lb_camera_update(c->cam, &p, p.bokeh_enable_image ? &img : nullptr)
                        : lb_camera_create(&p, p.bokeh_enable_image ? &img : nullptr, /*device*/ 0, &c->cam);
  if (rc != LB_OK) {  // lentil.h:226-227: setup errors abort the render
    AiMsgError("[LENTIL B200] %s", lb_last_error());
    AiRenderAbort();
    return;
  }
  c->pre_n = 0;
  c->pre_index.clear();

  // get_bidirectional_status (lentil.h:1150-1172)
  c->redistribution = true;
  if (!p.enable_dof) c->redistribution = false;
  else if (p.bidir_sample_mult == 0) c->redistribution = false;
  else if (AiNodeGetBool(on, AtString("enable_progressive_render"))) {
    AiMsgError("[LENTIL BIDIRECTIONAL] Progressive rendering is not supported.");
    AiRenderAbort();
    c->redistribution = false;
  }
  c->aovs.clear();
  c->has_crypto = false;
  if (!c->redistribution) return;

  // cryptomatte AOV shader present? (lentil.h:244-270; its setup is assumed complete -- the wait loop belongs to the host)
  c->cryptomatte_lentil = false;
  AtArray *aov_shaders = AiNodeGetArray(on, AtString("aov_shaders"));
  for (uint32_t i = 0; i < AiArrayGetNumElements(aov_shaders); ++i) {
    AtNode *n = static_cast<AtNode *>(AiArrayGetPtr(aov_shaders, i));
    if (n && AiNodeEntryGetNameAtString(AiNodeGetNodeEntry(n)) == AtString("cryptomatte")) c->cryptomatte_lentil = true;
  }
  // setup_lentil_aovs: copy the operator's list
  AtNode *op = nullptr;
  AtNodeIterator *iter = AiUniverseGetNodeIterator(universe, AI_NODE_ALL);
  while (!AiNodeIteratorFinished(iter)) {
    AtNode *n = AiNodeIteratorGetNext(iter);
    if (AiNodeEntryGetNameAtString(AiNodeGetNodeEntry(n)) == AtString("lentil_operator")) { op = n; break; }
  }
  AiNodeIteratorDestroy(iter);
  if (!op) {
    AiMsgError("[LENTIL] Since Lentil 2.5, lentil requires an operator (lentil_operator) to function. Please insert this operator.");
    c->redistribution = false;
    return;
  }
  OperatorData *od = (OperatorData *)AiNodeGetLocalData(op);
  c->aovs.insert(c->aovs.end(), od->aovs.begin(), od->aovs.end());
  // setup_crypto_aovs: ranked cryptomatte outputs of options.outputs
  AtArray *outputs = AiNodeGetArray(on, AtString("outputs"));
  std::vector<AOVData> crypto_aovs;
  for (uint32_t i = 0; i < AiArrayGetNumElements(outputs); ++i) {
    AOVData aov(universe, std::string(AiArrayGetStr(outputs, i).c_str()));
    bool replace_filter = true, cryptomatte_aov = false;
    if (aov.to.aov_name_tok == "crypto_material" || aov.to.aov_name_tok == "crypto_asset" || aov.to.aov_name_tok == "crypto_object") {
      replace_filter = false;
      cryptomatte_aov = true;
    } else if (aov.to.aov_name_tok.find("crypto_") != std::string::npos) {
      aov.is_crypto = true;
      cryptomatte_aov = true;
    }
    if (cryptomatte_aov) {
      if (replace_filter && aov.to.aov_name_tok != "lentil_replaced_filter") aov.to.filter_tok = "lentil_replaced_filter";
      crypto_aovs.push_back(aov);
    }
  }
  c->aovs.insert(c->aovs.end(), crypto_aovs.begin(), crypto_aovs.end());
  rebuild_arnold_outputs_from_list(universe, c->aovs);
  // setup_filter
  sanitize_aov_list(c->aovs);
  c->xres_without_region = AiNodeGetInt(on, AtString("xres"));
  c->yres_without_region = AiNodeGetInt(on, AtString("yres"));
  c->region_min_x = AiNodeGetInt(on, AtString("region_min_x"));
  c->region_min_y = AiNodeGetInt(on, AtString("region_min_y"));
  c->region_max_x = AiNodeGetInt(on, AtString("region_max_x"));
  c->region_max_y = AiNodeGetInt(on, AtString("region_max_y"));
  auto unset = [](int v) { return v == INT32_MIN || v == INT32_MAX; };
  if (unset(c->region_min_x) || unset(c->region_max_x) || unset(c->region_min_y) || unset(c->region_max_y)) {
    c->region_min_x = 0; c->region_min_y = 0; c->region_max_x = c->xres_without_region; c->region_max_y = c->yres_without_region;
  }
  c->xres = c->region_max_x - c->region_min_x + 1;
  c->yres = c->region_max_y - c->region_min_y + 1;
  c->filter_width = AiNodeEntryGetCount(AiNodeEntryLookUp(AtString("imager_denoiser_oidn"))) != 0 ? 1.0f : 1.5f;
  c->imager_print_once_only = false;
  if (c->aovs.empty() || c->aovs.size() > 16) {
    if (c->aovs.size() > 16) AiMsgError("[LENTIL B200] more than 16 lentil AOVs");
    c->redistribution = false;
    return;
  }
  std::vector<lb_aov_desc> desc(c->aovs.size());
  for (size_t a = 0; a < c->aovs.size(); ++a) {
    AOVData &aov = c->aovs[a];
    aov.index = (int)a;
    memset(&desc[a], 0, sizeof(lb_aov_desc));
    strncpy(desc[a].name, aov.name.c_str(), 63);
    AtNode *driver = aov.to.get_driver();
    const bool crypto_tables = aov.to.aov_name_tok.find("crypto_") != std::string::npos && driver && AiNodeIs(driver, AtString("driver_exr"));
    if (aov.is_crypto && crypto_tables) { desc[a].filter = LB_FILTER_CRYPTO; c->has_crypto = true; }
    else desc[a].filter = aov.original_filter == AtString("closest_filter") ? LB_FILTER_CLOSEST : LB_FILTER_GAUSSIAN;  // lentil.h:827,832
    desc[a].role = aov.name == AtString("RGBA") ? LB_AOV_RGBA : (aov.name == AtString("lentil_debug") ? LB_AOV_LENTIL_DEBUG : LB_AOV_PLAIN);
  }
  lb_frame_desc frame{c->xres, c->yres, c->xres_without_region, c->yres_without_region, c->region_min_x, c->region_min_y, 0};
  {  // batches of the previous frame are dropped with its framebuffers (destroy_buffers, lentil.h:214)
    std::lock_guard<std::mutex> lk(c->mu);
    for (LbSampleBatch *b : c->batches) b->clear();
  }
  if (lb_filter_begin(c->cam, &frame, (int)desc.size(), desc.data()) != LB_OK) {
    AiMsgError("[LENTIL B200] %s", lb_last_error());
    AiRenderAbort();
    c->redistribution = false;
  }
}
