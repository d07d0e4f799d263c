// lentil_b200_filter.cpp — the `lentil_filter` node over liblentil_b200.so.
// Replaces /root/reference/src/lentil_filter.cpp: filter_pixel (:66-480) no longer splats on the CPU; for the RGBA AOV it
// copies what the reference reads from the AtAOVSampleIterator (:105-234, lentil.h:779-811) into a per-thread lb_samples
// batch and hands full batches to lb_filter_accumulate_host.  Classification, reverse tracing and the splat run on the GPU.
#include "lentil_b200_adaptor.h"

#include <cmath>

AI_FILTER_NODE_EXPORT_METHODS(LentilFilterDataMtd);

node_parameters { AiMetaDataSetBool(nentry, nullptr, "force_update", true); }

node_initialize {
  static const char *required_aovs[] = {"RGBA RGBA", "VECTOR P", "FLOAT Z", "FLOAT lentil_time", "FLOAT lentil_debug", "RGB lentil_raydir",
                                        "RGB opacity", "RGBA transmission", "FLOAT lentil_bidir_ignore", NULL};  // lentil_filter.cpp:16-25
  AiFilterInitialize(node, false, required_aovs);
}

node_update {
  AiFilterUpdate(node, AiNodeEntryGetCount(AiNodeEntryLookUp(AtString("imager_denoiser_oidn"))) != 0 ? 1.0f : 1.5f);  // :33-38
}

node_finish {}

filter_output_type {  // :45-63
  switch (input_type) {
    case AI_TYPE_RGBA: case AI_TYPE_RGB: case AI_TYPE_VECTOR: case AI_TYPE_FLOAT: return AI_TYPE_RGBA;
    default: return AI_TYPE_NONE;
  }
}

filter_pixel {
  (void)data_type;
  AtUniverse *universe = AiNodeGetUniverse(node);
  AtNode *camera_node = AiUniverseGetCamera(universe);
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(camera_node);
  *((AtRGBA *)data_out) = AI_RGBA_ZERO;  // every lentil AOV is rewritten by the imager after the frame (lentil_imager.cpp:112-189)
  const bool rgba_aov = AiAOVSampleIteratorGetAOVName(iterator) == AtString("RGBA");  // :73
  if (!rgba_aov) return;
  const int aa_samples_set_by_user = AiNodeGetInt(AiUniverseGetOptions(universe), AtString("AA_samples"));
  const bool adaptive_sampling = AiNodeGetBool(AiUniverseGetOptions(universe), AtString("enable_adaptive_sampling"));
  float inverse_sample_density = 0.0f;
  if (!adaptive_sampling) {  // :79-88: density from the number of samples handed over
    int samples_counter = 0;
    while (AiAOVSampleIteratorGetNext(iterator)) ++samples_counter;
    AiAOVSampleIteratorReset(iterator);
    const float AA_samples = std::sqrt(samples_counter) / c->filter_width;
    inverse_sample_density = 1.0 / (AA_samples * AA_samples);
    if (static_cast<int>(std::round(AA_samples)) != aa_samples_set_by_user || aa_samples_set_by_user < 3) c->redistribution = false;
  }
  if (!c->redistribution) return;

  int px, py;
  AiAOVSampleIteratorGetPixel(iterator, px, py);
  px -= c->region_min_x;  // :99-100
  py -= c->region_min_y;
  LbSampleBatch &b = lb_adaptor_thread_batch(c);
  static const AtString s_p("P"), s_z("Z"), s_raydir("lentil_raydir"), s_volume("volume"), s_time("lentil_time"), s_transmission("transmission"),
      s_ignore("lentil_ignore"), s_opacity("opacity"), s_debug("lentil_debug");
  for (int sampleid = 0; AiAOVSampleIteratorGetNext(iterator) == true; sampleid++) {
    float inv_density = inverse_sample_density;
    uint32_t flags = 0;
    if (adaptive_sampling) {  // :108-113
      inv_density = AiAOVSampleIteratorGetInvDensity(iterator);
      if (inv_density > 0.2f) flags |= LB_SAMPLE_IGNORE;
    }
    const float time = AiAOVSampleIteratorGetAOVFlt(iterator, s_time);
    AtMatrix w2c;
    AiWorldToCameraMatrix(c->camera_node, time, w2c);  // :139-141
    // one matrix and one density per batch: a sample that changes either starts a new batch
    if (b.n && (inv_density != b.inv_density || memcmp(w2c.data, b.w2c, sizeof b.w2c) != 0)) lb_adaptor_flush(c, b);
    if (b.n == 0) {
      b.inv_density = inv_density;
      memcpy(b.w2c, w2c.data, sizeof b.w2c);
    }
    const AtRGBA sample = AiAOVSampleIteratorGetRGBA(iterator);                  // :115
    const AtVector P = AiAOVSampleIteratorGetAOVVec(iterator, s_p);              // :116 (world space: the device applies :119-142)
    const float depth = AiAOVSampleIteratorGetAOVFlt(iterator, s_z);             // :117
    const AtVector raydir = AiAOVSampleIteratorGetAOVVec(iterator, s_raydir);    // :121
    if (AiColorMaxRGB(AiAOVSampleIteratorGetAOVRGB(iterator, s_volume)) > 0.0f) flags |= LB_SAMPLE_VOLUME;  // :135-137
    const AtRGBA tr = AiAOVSampleIteratorGetAOVRGBA(iterator, s_transmission);   // :152
    if (AiAOVSampleIteratorGetAOVFlt(iterator, s_ignore) > 0.0f) flags |= LB_SAMPLE_IGNORE;  // :162 (reads "lentil_ignore", lentil.h:184)
    b.px.push_back(px); b.py.push_back(py);
    b.rgba.insert(b.rgba.end(), {sample.r, sample.g, sample.b, sample.a});
    b.pos.insert(b.pos.end(), {P.x, P.y, P.z, depth});
    b.raydir.insert(b.raydir.end(), {raydir.x, raydir.y, raydir.z, 0.f});
    b.transmission.insert(b.transmission.end(), {tr.r, tr.g, tr.b, tr.a});
    b.flags.push_back(flags);
    // every AOV of the sample widened to RGBA (:206-234); RGBA itself and lentil_debug need no copy
    for (size_t a = 0; a < c->aovs.size(); ++a) {
      const AOVData &aov = c->aovs[a];
      if (aov.is_crypto || aov.name == s_debug || a == 0 && aov.name == AtString("RGBA")) continue;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      switch (aov.type) {
        case AI_TYPE_RGBA: { const AtRGBA x = AiAOVSampleIteratorGetAOVRGBA(iterator, aov.name); v[0] = x.r; v[1] = x.g; v[2] = x.b; v[3] = x.a; } break;
        case AI_TYPE_RGB: { const AtRGB x = AiAOVSampleIteratorGetAOVRGB(iterator, aov.name); v[0] = x.r; v[1] = x.g; v[2] = x.b; v[3] = 1.f; } break;
        case AI_TYPE_FLOAT: { const float x = AiAOVSampleIteratorGetAOVFlt(iterator, aov.name); v[0] = v[1] = v[2] = x; v[3] = 1.f; } break;
        case AI_TYPE_VECTOR: { const AtVector x = AiAOVSampleIteratorGetAOVVec(iterator, aov.name); v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = 1.f; } break;
      }
      b.values[a].insert(b.values[a].end(), v, v + 4);
    }
    // cryptomatte depth sub-samples (cryptomatte_construct_cache, lentil.h:779-811): only the reads stay here, the
    // transparency weighting and the id merge run on the device
    if (c->has_crypto) {
      const size_t at = b.crypto_opacity.size();
      b.crypto_opacity.resize(at + LB_CRYPTO_MAX_DEPTH, 0.f);
      uint8_t count = 0;
      bool first = true;
      for (size_t a = 0; a < c->aovs.size(); ++a) {
        if (!c->aovs[a].is_crypto) continue;
        std::vector<float> &ids = b.crypto_ids[a];
        const size_t ia = ids.size();
        ids.resize(ia + LB_CRYPTO_MAX_DEPTH, 0.f);
        int d = 0;
        while (AiAOVSampleIteratorGetNextDepth(iterator)) {
          if (d < LB_CRYPTO_MAX_DEPTH) {
            if (first) b.crypto_opacity[at + d] = AiColorToGrey(AiAOVSampleIteratorGetAOVRGB(iterator, s_opacity));
            ids[ia + d] = AiAOVSampleIteratorGetAOVFlt(iterator, c->aovs[a].name);
          }
          ++d;
        }
        if (first) count = (uint8_t)std::min(d, (int)LB_CRYPTO_MAX_DEPTH);
        first = false;
        // the depth walk leaves the real iterator on the next sample: re-seek (reset_iterator_to_id, lentil.h:1176-1184)
        AiAOVSampleIteratorReset(iterator);
        for (int i = 0; AiAOVSampleIteratorGetNext(iterator) == true; i++)
          if (i == sampleid) break;
      }
      b.crypto_count.push_back(count);
    }
    ++b.n;
  }
  if (b.n >= kLbFlushSamples) lb_adaptor_flush(c, b);
}
