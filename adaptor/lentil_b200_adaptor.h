// lentil_b200_adaptor.h — what the three Arnold nodes of the adaptor share: the stand-in for `struct Camera`
// (/root/reference/src/lentil.h:92-1671) that the reference's nodes fetch with AiNodeGetLocalData(camera_node).
//
// The adaptor is the reference plugin with its per-ray compute replaced by liblentil_b200.so: node registration,
// parameter reading, AOV bookkeeping (aov_data.h / operator_data.h stay the plugin's own, included from the reference
// tree) and the Arnold callbacks are here; camera_create_ray / filter_pixel / driver_process_bucket only marshal into
// the batch C ABI of include/lentil_b200.h.  Built against the Arnold SDK in production and against oracle/shims/ai.h
// in this repository (adaptor/Makefile), where oracle/ref_harness.cpp drives it through the same node-method tables
// as the compiled reference (tests/test_adaptor_gpu.py).
#pragma once
#include <ai.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <regex>  // aov_data.h tokenises options.outputs with std::regex and relies on its includer for the header
#include <string>
#include <unordered_map>
#include <vector>

#include "aov_data.h"       // reference: AOVData, TokenizedOutputLentil
#include "operator_data.h"  // reference: OperatorData (filled by lentil_operator.cpp)
#include "../include/lentil_b200.h"

// filter_pixel's per-sample reads (lentil_filter.cpp:105-234) as structure-of-arrays: one lb_samples batch
struct LbSampleBatch {
  std::vector<int32_t> px, py;
  std::vector<float> rgba, pos, raydir, transmission;  // [n][4]
  std::vector<uint32_t> flags;
  std::vector<std::vector<float>> values;              // per AOV slot: [n][4], empty when the slot takes no per-sample value
  std::vector<uint8_t> crypto_count;                   // [n]
  std::vector<float> crypto_opacity;                   // [n][LB_CRYPTO_MAX_DEPTH]
  std::vector<std::vector<float>> crypto_ids;          // per AOV slot: [n][LB_CRYPTO_MAX_DEPTH]
  float inv_density = 0.f;
  float w2c[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  size_t n = 0;
  void clear();
};

struct LbAdaptorCamera {
  const uint64_t uid;  // threads key their batch by this, not by the address (a later camera may reuse it)
  LbAdaptorCamera();
  lb_camera *cam = nullptr;
  lb_camera_params params{};
  AtNode *camera_node = nullptr, *options_node = nullptr;
  bool redistribution = false, cryptomatte_lentil = false, imager_print_once_only = false;
  float filter_width = 1.5f;
  int xres = 0, yres = 0, xres_without_region = 0, yres_without_region = 0, region_min_x = 0, region_min_y = 0, region_max_x = 0, region_max_y = 0;
  std::vector<AOVData> aovs;  // the lentil_replaced_filter AOVs, index == lb aov index (lentil.h:988-1117)
  bool has_crypto = false;
  std::mutex mu;
  std::vector<LbSampleBatch *> batches;  // every thread's batch, for the flush before the imager runs
  // per-bucket prefetch of camera rays: inputs the renderer is about to ask for, traced in one batch
  std::vector<float> pre_out[7];  // origin, dir, dOdx, dOdy, dDdx, dDdy, weight as [3][n] planes
  size_t pre_n = 0;
  std::unordered_map<uint64_t, uint32_t> pre_index;
  std::vector<AtCameraInput> pre_in;
  ~LbAdaptorCamera();
};

static constexpr size_t kLbFlushSamples = (size_t)1 << 20;  // a thread's batch goes to the GPU at this size (40+ MB of samples)

LbSampleBatch &lb_adaptor_thread_batch(LbAdaptorCamera *c);
int lb_adaptor_flush(LbAdaptorCamera *c, LbSampleBatch &b);
int lb_adaptor_flush_all(LbAdaptorCamera *c);
void lb_adaptor_setup(LbAdaptorCamera *c, AtUniverse *universe);  // Camera::setup_camera, lentil.h:211-281

extern "C" {
// Trace the camera rays of `n` upcoming camera_create_ray calls in one lb_camera_create_rays_host batch; the calls are
// then answered from the result (exact match on sx, sy, lensx, lensy).  Inputs that were not prefetched fall back to a
// one-ray call.
AI_EXPORT_LIB int lentil_b200_prefetch_rays(AtNode *camera_node, size_t n, const AtCameraInput *inputs);
AI_EXPORT_LIB lb_camera *lentil_b200_camera_handle(AtNode *camera_node);
AI_EXPORT_LIB int lentil_b200_redistribution(AtNode *camera_node);
AI_EXPORT_LIB int lentil_b200_flush(AtNode *camera_node);
AI_EXPORT_LIB int lentil_b200_aov_index(AtNode *camera_node, const char *aov_name);
}
