// lentil_b200_operator.cpp — the `lentil_operator` node of the adaptor: the scene pre-pass in front of the redistribution path
// (/root/reference/src/lentil_operator.cpp:19-191).  It decides WHICH outputs the device framebuffers of lb_filter_begin are made
// for: every entry of options.outputs whose type lentil can redistribute has its filter swapped for the one shared
// `lentil_replaced_filter` node (the original filter kind is remembered: gaussian -> summed and weight-normalised, closest ->
// depth-keyed, lentil.h:827,832), three helper outputs are appended (lentil_debug, lentil_time, lentil_raydir) and the two AOV
// shaders that fill lentil_time / lentil_raydir are linked into options.aov_shaders.  The list it leaves in OperatorData is what
// lb_adaptor_setup copies (lentil.h:988-1012) and turns into lb_aov_desc[].  No compute: host-side bookkeeping only.
//
// tests/test_operator.py runs this node and the compiled reference's over the same scenes and compares everything they leave.
#include "lentil_b200_adaptor.h"

AI_OPERATOR_NODE_EXPORT_METHODS(LentilOperatorMtd);

namespace {
const char *const kSharedFilter = "lentil_replaced_filter";

// filter kinds the redistribution can stand in for (lentil_operator.cpp:10-12,53-60)
bool filter_kind_known(const AtString &entry) {
  return entry == AtString("gaussian_filter") || entry == AtString("closest_filter") || entry == AtString("variance_filter");
}
// AOV data types with a per-sample read in filter_pixel (lentil_filter.cpp:214-234; lentil_operator.cpp:65-71)
bool type_redistributable(const std::string &tok) { return tok == "RGBA" || tok == "RGB" || tok == "FLOAT" || tok == "VECTOR"; }
// the unranked cryptomatte outputs are display-only and keep their filter (lentil_operator.cpp:75-78)
bool unranked_cryptomatte(const std::string &name) { return name == "crypto_material" || name == "crypto_asset" || name == "crypto_object"; }

// A helper output (lentil_operator.cpp:103-128): the FIRST output with its name and type tokens replaced, tokenised again from
// the rebuilt string so that every other option of that output (camera, filter, driver, HALF) carries over.
AOVData helper_output(AtUniverse *universe, const AOVData &first, const char *name, const char *type) {
  AOVData a = first;
  a.to.aov_type_tok = type;
  a.to.aov_name_tok = name;
  a.to = TokenizedOutputLentil(universe, AtString(a.to.rebuild_output().c_str()));
  a.name = AtString(name);
  a.type = string_to_arnold_type(a.to.aov_type_tok);
  return a;
}

// `<write_entry> <write_name>` fed by `<read_entry> <read_name>` reading the shader-globals variable `variable`, appended to
// options.aov_shaders (lentil_operator.cpp:131-158)
void append_state_writer(AtUniverse *universe, const char *write_entry, const char *write_name, const char *read_entry,
                         const char *read_name, const char *variable, const char *aov_name) {
  AtNode *options = AiUniverseGetOptions(universe);
  AtNode *writer = AiNode(universe, AtString(write_entry), AtString(write_name));
  AtNode *reader = AiNode(universe, AtString(read_entry), AtString(read_name));
  AiNodeSetStr(reader, AtString("variable"), AtString(variable));
  AiNodeSetStr(writer, AtString("aov_name"), AtString(aov_name));
  AiNodeLink(reader, AtString("aov_input"), writer);
  AtArray *shaders = AiNodeGetArray(options, AtString("aov_shaders"));
  const uint32_t n = AiArrayGetNumElements(shaders);
  AiArrayResize(shaders, n + 1, 1);
  AiArraySetPtr(shaders, n, (void *)writer);
  AiNodeSetArray(options, AtString("aov_shaders"), shaders);
}
}  // namespace

node_parameters { AiMetaDataSetBool(nentry, nullptr, "force_update", true); }

operator_init {
  AiNodeSetLocalData(op, new OperatorData());
  return true;
}

operator_cook {
  OperatorData *data = static_cast<OperatorData *>(AiNodeGetLocalData(op));
  AtUniverse *universe = AiNodeGetUniverse(op);
  // only scenes rendered through the lentil camera are touched (lentil_operator.cpp:30-33)
  if (AiNodeEntryGetNameAtString(AiNodeGetNodeEntry(AiUniverseGetCamera(universe))) != AtString("lentil_camera")) return false;
  if (!AiNodeLookUpByName(universe, AtString(kSharedFilter))) AiNode(universe, AtString("lentil_filter"), AtString(kSharedFilter));

  AtArray *outputs = AiNodeGetArray(AiUniverseGetOptions(universe), AtString("outputs"));
  const uint32_t n_outputs = AiArrayGetNumElements(outputs);
  for (uint32_t i = 0; i < n_outputs; ++i) {
    AOVData aov(universe, std::string(AiArrayGetStr(outputs, i).c_str()));
    const AtString kind = AiNodeEntryGetNameAtString(AiNodeGetNodeEntry(AiNodeLookUpByName(universe, AtString(aov.to.filter_tok.c_str()))));
    if (filter_kind_known(kind)) aov.original_filter = kind;
    else {
      AiMsgWarning("[LENTIL] Specified AOV filter (%s) is incompatible with Lentil. Defaulting to gaussian_filter.", kind.c_str());
      aov.original_filter = AtString("gaussian_filter");
    }
    const std::string &name = aov.to.aov_name_tok;
    bool swap = type_redistributable(aov.to.aov_type_tok);
    if (unranked_cryptomatte(name)) swap = false;
    else if (name.find("crypto_") != std::string::npos) continue;  // ranked cryptomatte outputs are picked up by the camera (lentil.h:1015-1052)
    if (swap && name != kSharedFilter) aov.to.filter_tok = kSharedFilter;
    for (const AOVData &seen : data->aovs)  // the same AOV on a second driver: one framebuffer (aov_data.h:166-175 drops it later)
      if (seen.to.aov_name_tok == name) aov.is_duplicate = true;
    data->aovs.push_back(aov);
  }
  if (data->aovs.empty()) {  // the reference indexes aovs[0] regardless (lentil_operator.cpp:107): a scene without outputs is an error here
    AiMsgError("[LENTIL B200] lentil_operator: options.outputs holds no output to derive the helper AOVs from");
    return false;
  }

  // helper outputs, all derived from the first entry the list has ever held (lentil_operator.cpp:107,118,127)
  {
    AOVData debug = helper_output(universe, data->aovs[0], "lentil_debug", "FLOAT");
    debug.original_filter = AtString("closest_filter");
    data->aovs.push_back(debug);
  }
  data->aovs.push_back(helper_output(universe, data->aovs[0], "lentil_time", "FLOAT"));
  data->aovs.push_back(helper_output(universe, data->aovs[0], "lentil_raydir", "RGB"));

  append_state_writer(universe, "aov_write_float", "lentil_time_write", "state_float", "lentil_time_read", "time", "lentil_time");
  append_state_writer(universe, "aov_write_rgb", "lentil_raydir_write", "state_vector", "lentil_raydir_read", "Rd", "lentil_raydir");
  return true;
}

operator_post_cook { return true; }

operator_cleanup {
  delete static_cast<OperatorData *>(AiNodeGetLocalData(op));
  return true;
}

void registerLentilOperator(AtNodeLib *node) {
  node->methods = LentilOperatorMtd;
  node->name = "lentil_operator";
  node->node_type = AI_NODE_OPERATOR;
  strcpy(node->version, AI_VERSION);
}
