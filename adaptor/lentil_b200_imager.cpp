// lentil_b200_imager.cpp — the `imager_lentil` node over liblentil_b200.so.
// Replaces /root/reference/src/lentil_imager.cpp: driver_process_bucket (:66-193) copies each bucket of each lentil AOV out of
// the image lb_imager_resolve_host resolves once per frame and AOV.
#include "lentil_b200_adaptor.h"

AI_DRIVER_NODE_EXPORT_METHODS(LentilImagerMtd);

node_parameters {
  AiMetaDataSetStr(nentry, nullptr, AtString("subtype"), AtString("imager"));
  AiParameterBool(AtString("enable"), true);
}
node_initialize { AiDriverInitialize(node, false); }
node_update {
  AtRenderSession *rs = AiUniverseGetRenderSession(AiNodeGetUniverse(node));
  AiRenderSetHintInt(rs, AtString("imager_padding"), 0);
  AiRenderSetHintInt(rs, AtString("imager_schedule"), 0x02);  // after the full frame (lentil_imager.cpp:37-38)
}
node_finish {}
driver_supports_pixel_type { (void)node; return pixel_type == AI_TYPE_RGBA || pixel_type == AI_TYPE_RGB || pixel_type == AI_TYPE_FLOAT || pixel_type == AI_TYPE_VECTOR; }
driver_open { (void)node; (void)iterator; (void)display_window; (void)data_window; (void)bucket_size; }
driver_extension {
  static const char *extensions[] = {NULL};
  return extensions;
}
driver_needs_bucket { (void)node; (void)bucket_xo; (void)bucket_yo; (void)bucket_size_x; (void)bucket_size_y; (void)tid; return true; }
driver_prepare_bucket { (void)node; (void)bucket_xo; (void)bucket_yo; (void)bucket_size_x; (void)bucket_size_y; (void)tid; }
driver_write_bucket { (void)node; (void)iterator; (void)sample_iterator; (void)bucket_xo; (void)bucket_yo; (void)bucket_size_x; (void)bucket_size_y; (void)tid; }
driver_close { (void)node; (void)iterator; }

driver_process_bucket {
  (void)sample_iterator; (void)tid;
  AiOutputIteratorReset(iterator);
  AtUniverse *universe = AiNodeGetUniverse(node);
  LbAdaptorCamera *c = (LbAdaptorCamera *)AiNodeGetLocalData(AiUniverseGetCamera(universe));
  if (!c->redistribution) {  // :74-81
    if (!c->imager_print_once_only) {
      AiMsgInfo("[LENTIL IMAGER] Skipping imager");
      c->imager_print_once_only = true;
    }
    return;
  }
  lb_adaptor_flush_all(c);  // samples still sitting in the threads' batches (no-op once they are empty)
  AtString aov_name = AtString("");
  int aov_type = 0;
  const void *bucket_data = nullptr;
  while (AiOutputIteratorGetNext(iterator, &aov_name, &aov_type, &bucket_data)) {
    int index = -1;
    for (size_t a = 0; a < c->aovs.size(); ++a)
      if (c->aovs[a].name == aov_name) index = (int)a;  // last match, as the reference's loop (:103-106)
    if (index < 0) continue;
    if (aov_name == AtString("lentil_ignore") || aov_name == AtString("lentil_time")) continue;  // :109
    if (lb_imager_resolve_host(c->cam, index, bucket_xo, bucket_yo, bucket_size_x, bucket_size_y, (float *)bucket_data) != LB_OK)
      AiMsgError("[LENTIL B200] %s", lb_last_error());
  }
  c->imager_print_once_only = true;
}
