// lentil_b200_loader.cpp — the plugin entry point of the adaptor (/root/reference/src/lentil_loader.cpp:20-28): Arnold calls
// NodeLoader(i, &lib) with i = 0, 1, ... until it returns false and installs the node each call describes.  Same four nodes,
// same names, node types and output types as the reference registers (lentil_camera.cpp:176-182, lentil_filter.cpp:486-492,
// lentil_imager.cpp:203-209, lentil_operator.cpp:186-191); the method tables are the adaptor's, over liblentil_b200.so.
#include <ai.h>

#include <cstring>

extern const AtNodeMethods *lentilMethods;        // lentil_b200_camera.cpp
extern const AtNodeMethods *LentilFilterDataMtd;  // lentil_b200_filter.cpp
extern const AtNodeMethods *LentilImagerMtd;      // lentil_b200_imager.cpp
extern const AtNodeMethods *LentilOperatorMtd;    // lentil_b200_operator.cpp

namespace {
struct NodeRow {
  const char *name;
  int node_type;
  int output_type;  // < 0: left as the caller passed it (the reference's operator registration does not set it)
  const AtNodeMethods *const *methods;  // read at call time: the tables are initialised in other translation units
};
const NodeRow kNodes[] = {
    {"lentil_camera", AI_NODE_CAMERA, AI_TYPE_UNDEFINED, &lentilMethods},
    {"lentil_filter", AI_NODE_FILTER, AI_TYPE_NONE, &LentilFilterDataMtd},
    {"imager_lentil", AI_NODE_DRIVER, AI_TYPE_NONE, &LentilImagerMtd},
    {"lentil_operator", AI_NODE_OPERATOR, -1, &LentilOperatorMtd},
};
}  // namespace

node_loader {
  if (i < 0 || i >= (int)(sizeof(kNodes) / sizeof(kNodes[0]))) return false;
  const NodeRow &row = kNodes[i];
  strcpy(node->version, AI_VERSION);
  node->methods = *row.methods;
  node->name = row.name;
  node->node_type = row.node_type;
  if (row.output_type >= 0) node->output_type = (uint8_t)row.output_type;
  return true;
}
