// ai.h — minimal stand-in for the Arnold SDK header, ONLY to compile the reference's own sources
// (/root/reference/src/*.h, lentil_camera.cpp, lentil_filter.cpp, lentil_imager.cpp, lentil_operator.cpp,
// lentil_loader.cpp) into oracle/_ref without Arnold.  TEST INFRASTRUCTURE: nothing here is linked into the product.
//
// It models just what those files touch: POD math types with the operators used,
// a node/universe store with typed parameters, the AOV sample iterator over caller-supplied sample
// arrays, an output iterator, textures served from memory, and the node-definition macros.  Semantics
// that matter numerically and are NOT verifiable here (the SDK is absent) are marked [EXTERNAL].
#pragma once
#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <set>
#include <string>
#include <vector>

#define AI_VERSION "7.0.0.0-shim"
#define AI_PI 3.14159265358979323846f
#define AI_PIOVER2 1.57079632679489661923f
#define AI_BIG 1.0e12f
#define AI_INFINITE 1.0e30f
#define AI_EPSILON 1.0e-4f
#define AI_EXPORT_LIB

enum { AI_TYPE_BYTE = 0, AI_TYPE_INT, AI_TYPE_UINT, AI_TYPE_BOOLEAN, AI_TYPE_FLOAT, AI_TYPE_RGB, AI_TYPE_RGBA, AI_TYPE_VECTOR,
       AI_TYPE_VECTOR2 = 9, AI_TYPE_STRING, AI_TYPE_POINTER, AI_TYPE_NODE, AI_TYPE_ARRAY, AI_TYPE_MATRIX, AI_TYPE_ENUM,
       AI_TYPE_UNDEFINED = 255, AI_TYPE_NONE = 255 };
enum { AI_NODE_UNDEFINED = 0, AI_NODE_OPTIONS = 1, AI_NODE_CAMERA = 2, AI_NODE_FILTER = 0x40, AI_NODE_DRIVER = 0x80,
       AI_NODE_OPERATOR = 0x1000 /* [EXTERNAL] */, AI_NODE_ALL = 0xFFFF };
enum { AI_RAY_UNDEFINED = 0, AI_RAY_SHADOW = 2 };
enum { AI_AOV_BLEND_NONE = 0 };

// ---- AtString: interned C string, compared by pointer like the SDK's --------------------------------
struct AtString {
  const char *s = nullptr;
  AtString() {}
  AtString(const char *c) { s = c ? intern(c) : nullptr; }
  const char *c_str() const { return s ? s : ""; }
  bool empty() const { return !s || !*s; }
  bool operator==(const AtString &o) const { return s == o.s || (c_str()[0] == 0 && o.c_str()[0] == 0); }
  bool operator!=(const AtString &o) const { return !(*this == o); }
  bool operator<(const AtString &o) const { return s < o.s; }
  static const char *intern(const char *c) {
    static std::set<std::string> pool;
    return pool.insert(c).first->c_str();
  }
};

// ---- vectors / colours --------------------------------------------------------------------------
struct AtVector2 {
  float x = 0, y = 0;
  AtVector2() {}
  AtVector2(float a, float b) : x(a), y(b) {}
  AtVector2 operator*(float f) const { return AtVector2(x * f, y * f); }
  AtVector2 &operator*=(float f) { x *= f; y *= f; return *this; }
  AtVector2 operator-(const AtVector2 &o) const { return AtVector2(x - o.x, y - o.y); }
  AtVector2 operator+(const AtVector2 &o) const { return AtVector2(x + o.x, y + o.y); }
  AtVector2 operator/(float f) const { return AtVector2(x / f, y / f); }
  AtVector2 &operator+=(const AtVector2 &o) { x += o.x; y += o.y; return *this; }
  AtVector2 &operator/=(float f) { const float c = 1.0f / f; x *= c; y *= c; return *this; }
  AtVector2 &operator-=(const AtVector2 &o) { x -= o.x; y -= o.y; return *this; }
};
inline AtVector2 operator*(float f, const AtVector2 &v) { return v * f; }
struct AtVector {
  float x = 0, y = 0, z = 0;
  AtVector() {}
  AtVector(float a, float b, float c) : x(a), y(b), z(c) {}
  float &operator[](int i) { return (&x)[i]; }
  const float &operator[](int i) const { return (&x)[i]; }
  AtVector operator+(const AtVector &o) const { return AtVector(x + o.x, y + o.y, z + o.z); }
  AtVector operator-(const AtVector &o) const { return AtVector(x - o.x, y - o.y, z - o.z); }
  AtVector operator-() const { return AtVector(-x, -y, -z); }
  AtVector operator*(float f) const { return AtVector(x * f, y * f, z * f); }
  AtVector operator/(float f) const { const float c = 1.0f / f; return AtVector(x * c, y * c, z * c); }  // [EXTERNAL] reciprocal multiply
  AtVector &operator*=(float f) { x *= f; y *= f; z *= f; return *this; }
  AtVector &operator/=(float f) { const float c = 1.0f / f; x *= c; y *= c; z *= c; return *this; }
  AtVector &operator+=(const AtVector &o) { x += o.x; y += o.y; z += o.z; return *this; }
  AtVector &operator-=(const AtVector &o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  bool operator==(const AtVector &o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const AtVector &o) const { return !(*this == o); }
};
inline AtVector operator*(float f, const AtVector &v) { return v * f; }
struct AtRGB {
  float r = 0, g = 0, b = 0;
  AtRGB() {}
  AtRGB(float c) : r(c), g(c), b(c) {}
  AtRGB(float a, float b_, float c) : r(a), g(b_), b(c) {}
  AtRGB operator*(float f) const { return AtRGB(r * f, g * f, b * f); }
  AtRGB operator/(float f) const { const float c = 1.0f / f; return AtRGB(r * c, g * c, b * c); }
  AtRGB operator*(const AtRGB &o) const { return AtRGB(r * o.r, g * o.g, b * o.b); }
  AtRGB operator+(const AtRGB &o) const { return AtRGB(r + o.r, g + o.g, b + o.b); }
  AtRGB operator-(const AtRGB &o) const { return AtRGB(r - o.r, g - o.g, b - o.b); }
  AtRGB &operator+=(const AtRGB &o) { r += o.r; g += o.g; b += o.b; return *this; }
  AtRGB &operator-=(const AtRGB &o) { r -= o.r; g -= o.g; b -= o.b; return *this; }
  AtRGB &operator*=(float f) { r *= f; g *= f; b *= f; return *this; }
  AtRGB &operator/=(float f) { const float c = 1.0f / f; r *= c; g *= c; b *= c; return *this; }
  bool operator==(const AtRGB &o) const { return r == o.r && g == o.g && b == o.b; }
  bool operator!=(const AtRGB &o) const { return !(*this == o); }
};
inline AtRGB operator*(float f, const AtRGB &c) { return c * f; }
struct AtRGBA {
  float r = 0, g = 0, b = 0, a = 0;
  AtRGBA() {}
  AtRGBA(float r_, float g_, float b_, float a_) : r(r_), g(g_), b(b_), a(a_) {}
  AtRGBA(float c) : r(c), g(c), b(c), a(c) {}  // [EXTERNAL] `AtRGBA = samples * redistribute` (lentil_filter.cpp:210) sets all four
  AtRGBA(const AtRGB &c, float a_ = 1.f) : r(c.r), g(c.g), b(c.b), a(a_) {}  // implicit in the SDK: lentil.h:763 assigns an AtRGB
  AtRGBA operator+(const AtRGBA &o) const { return AtRGBA(r + o.r, g + o.g, b + o.b, a + o.a); }
  AtRGBA operator-(const AtRGBA &o) const { return AtRGBA(r - o.r, g - o.g, b - o.b, a - o.a); }
  AtRGBA operator+(float f) const { return AtRGBA(r + f, g + f, b + f, a + f); }  // [EXTERNAL] alpha included
  AtRGBA operator*(float f) const { return AtRGBA(r * f, g * f, b * f, a * f); }
  AtRGBA operator/(float f) const { const float c = 1.0f / f; return AtRGBA(r * c, g * c, b * c, a * c); }
  AtRGBA operator*(const AtRGBA &o) const { return AtRGBA(r * o.r, g * o.g, b * o.b, a * o.a); }  // `rgba * AtRGB` -> AtRGBA(rgb, 1): alpha kept
  AtRGBA &operator+=(const AtRGBA &o) { r += o.r; g += o.g; b += o.b; a += o.a; return *this; }
  AtRGBA &operator-=(const AtRGBA &o) { r -= o.r; g -= o.g; b -= o.b; a -= o.a; return *this; }
  AtRGBA &operator*=(float f) { r *= f; g *= f; b *= f; a *= f; return *this; }
  AtRGBA &operator/=(float f) { const float c = 1.0f / f; r *= c; g *= c; b *= c; a *= c; return *this; }  // [EXTERNAL] reciprocal multiply
  bool operator==(const AtRGBA &o) const { return r == o.r && g == o.g && b == o.b && a == o.a; }
  bool operator!=(const AtRGBA &o) const { return !(*this == o); }
};
inline AtRGBA operator*(float f, const AtRGBA &c) { return c * f; }
#define AI_RGB_ZERO AtRGB(0.f, 0.f, 0.f)
#define AI_RGB_BLACK AtRGB(0.f, 0.f, 0.f)
#define AI_RGB_WHITE AtRGB(1.f, 1.f, 1.f)
#define AI_RGBA_ZERO AtRGBA(0.f, 0.f, 0.f, 0.f)

struct AtMatrix { float data[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}}; };
struct AtBBox2 { int minx, miny, maxx, maxy; };

inline float AiSqr(float a) { return a * a; }
inline float AiV3Dot(const AtVector &a, const AtVector &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float AiV3Length(const AtVector &a) { return std::sqrt(AiV3Dot(a, a)); }
inline AtVector AiV3Normalize(const AtVector &a) { float t = AiV3Length(a); if (t != 0) t = 1 / t; return a * t; }
inline AtVector AiV3Cross(const AtVector &a, const AtVector &b) { return AtVector(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float AiV3Dist(const AtVector &a, const AtVector &b) { return AiV3Length(a - b); }
inline bool AiV3IsSmall(const AtVector &a, float e = AI_EPSILON) { return std::abs(a.x) < e && std::abs(a.y) < e && std::abs(a.z) < e; }
inline float AiV2Dot(const AtVector2 &a, const AtVector2 &b) { return a.x * b.x + a.y * b.y; }
inline float AiV2Length(const AtVector2 &a) { return std::sqrt(AiV2Dot(a, a)); }
inline float AiV2Dist(const AtVector2 &a, const AtVector2 &b) { return AiV2Length(a - b); }
inline float AiColorMaxRGB(const AtRGB &c) { return std::max(c.r, std::max(c.g, c.b)); }
inline float AiColorMaxRGB(const AtRGBA &c) { return std::max(c.r, std::max(c.g, c.b)); }
inline float AiColorToGrey(const AtRGB &c) { return (c.r + c.g + c.b) / 3.0f; }
inline float AiBias(float a, float b) { return (a > 0) ? ((b > 0) ? std::pow(a, std::log(b) / std::log(0.5f)) : 0) : 0; }
inline float AiFastExp(float x) { return std::exp(x); }
inline AtVector AiM4PointByMatrixMult(const AtMatrix &m, const AtVector &p) {
  return AtVector(p.x * m.data[0][0] + p.y * m.data[1][0] + p.z * m.data[2][0] + m.data[3][0],
                  p.x * m.data[0][1] + p.y * m.data[1][1] + p.z * m.data[2][1] + m.data[3][1],
                  p.x * m.data[0][2] + p.y * m.data[1][2] + p.z * m.data[2][2] + m.data[3][2]);
}

// ---- messages / memory ----------------------------------------------------------------------------
inline int &shim_verbose() { static int v = 0; return v; }
inline void shim_msg(const char *lvl, const char *fmt, va_list ap) { if (!shim_verbose()) return; fprintf(stderr, "[%s] ", lvl); vfprintf(stderr, fmt, ap); fputc('\n', stderr); }
inline void AiMsgInfo(const char *fmt, ...) { va_list ap; va_start(ap, fmt); shim_msg("info", fmt, ap); va_end(ap); }
inline void AiMsgWarning(const char *fmt, ...) { va_list ap; va_start(ap, fmt); shim_msg("warn", fmt, ap); va_end(ap); }
inline void AiMsgError(const char *fmt, ...) { va_list ap; va_start(ap, fmt); shim_msg("error", fmt, ap); va_end(ap); }
inline int &shim_abort_count() { static int n = 0; return n; }
inline void AiRenderAbort() { ++shim_abort_count(); }
inline void *AiMalloc(size_t n) { return malloc(n); }
inline void AiFree(void *p) { free(p); }
inline void AiAddMemUsage(int64_t, const AtString &) {}
struct AtCritSec { int dummy; };
inline void AiCritSecInit(AtCritSec *) {}
inline void AiCritSecClose(AtCritSec *) {}
inline void AiCritSecEnter(AtCritSec *) {}
inline void AiCritSecLeave(AtCritSec *) {}

// ---- nodes ---------------------------------------------------------------------------------------------
struct AtNode;
struct AtUniverse;
struct AtArray { std::vector<std::string> strs; std::vector<void *> ptrs; };
struct AtParamValueShim { int i = 0; float f = 0; bool b = false; std::string s; void *p = nullptr; AtArray *a = nullptr; };
struct AtNodeEntry { std::string name; int count = 0; std::vector<std::string> meta; /* AiMetaDataSet* calls, in order */ };
struct AtNode {
  std::string name;
  AtNodeEntry entry;
  std::map<std::string, AtParamValueShim> params;
  void *local_data = nullptr;
  AtUniverse *universe = nullptr;
  AtMatrix world_to_camera;  // camera nodes: what AiWorldToCameraMatrix returns (identity unless the harness sets it)
  std::map<std::string, AtNode *> links;  // AiNodeLink: input name -> the node feeding it
  std::vector<std::string> required_aovs;  // AiFilterInitialize: the AOVs a filter node asks the renderer for
  float filter_width = 0.f;                // AiFilterUpdate
};
struct AtRenderSession { int dummy; std::vector<std::string> hints; /* AiRenderSetHintInt calls */ };
struct AtUniverse {
  AtNode *options = nullptr, *camera = nullptr;
  std::vector<AtNode *> nodes;
  std::map<std::string, int> entry_counts;  // AiNodeEntryLookUp/AiNodeEntryGetCount
  AtRenderSession session;
  std::deque<AtNode> created;  // nodes made by AiNode() (lentil_operator.cpp:37,131-132,147-148): owned here, listed in `nodes`
  std::vector<std::string> registered_aovs;  // AiAOVRegister calls, in order
};
inline AtUniverse *&shim_default_universe() { static AtUniverse *u = nullptr; return u; }
inline const AtParamValueShim &shim_param_get(const AtNode *n, const AtString &k) {
  static AtParamValueShim zero;
  auto it = n->params.find(k.c_str());
  return it == n->params.end() ? zero : it->second;
}
inline int AiNodeGetInt(const AtNode *n, const AtString &k) { return shim_param_get(n, k).i; }
inline float AiNodeGetFlt(const AtNode *n, const AtString &k) { return shim_param_get(n, k).f; }
inline bool AiNodeGetBool(const AtNode *n, const AtString &k) { return shim_param_get(n, k).b; }
inline AtString AiNodeGetStr(const AtNode *n, const AtString &k) { return AtString(shim_param_get(n, k).s.c_str()); }
inline AtArray *AiNodeGetArray(const AtNode *n, const AtString &k) { static AtArray empty; AtArray *a = shim_param_get(n, k).a; return a ? a : &empty; }
inline void AiNodeSetArray(AtNode *n, const AtString &k, AtArray *a) { n->params[k.c_str()].a = a; }
inline void *AiNodeGetLocalData(const AtNode *n) { return n->local_data; }
inline void AiNodeSetLocalData(AtNode *n, void *p) { n->local_data = p; }
inline AtUniverse *AiNodeGetUniverse(const AtNode *n) { return n->universe; }
inline const char *AiNodeGetName(const AtNode *n) { return n->name.c_str(); }
inline const AtNodeEntry *AiNodeGetNodeEntry(const AtNode *n) { static const AtNodeEntry none; return n ? &n->entry : &none; }  // [EXTERNAL] a null node has no entry; the stand-in gives it an unnamed one
inline AtString AiNodeEntryGetNameAtString(const AtNodeEntry *e) { return AtString(e->name.c_str()); }
inline bool AiNodeIs(const AtNode *n, const AtString &s) { return n->entry.name == s.c_str(); }
inline AtNode *AiUniverseGetOptions(const AtUniverse *u) { return u->options; }
inline AtNode *AiUniverseGetCamera(const AtUniverse *u) { return u->camera; }
inline AtRenderSession *AiUniverseGetRenderSession(AtUniverse *u) { return &u->session; }
inline void AiRenderSetHintInt(AtRenderSession *rs, const AtString &k, int v) { if (rs) rs->hints.push_back(std::string(k.c_str()) + "=" + std::to_string(v)); }
inline AtNode *AiNodeLookUpByName(const AtUniverse *u, const AtString &name) {
  for (AtNode *n : u->nodes) if (n->name == name.c_str()) return n;
  return nullptr;
}
inline const AtNodeEntry *AiNodeEntryLookUp(const AtString &name) {
  static std::map<std::string, AtNodeEntry> entries;
  AtNodeEntry &e = entries[name.c_str()];
  e.name = name.c_str();
  AtUniverse *u = shim_default_universe();
  e.count = (u && u->entry_counts.count(name.c_str())) ? u->entry_counts[name.c_str()] : 0;
  return &e;
}
inline int AiNodeEntryGetCount(const AtNodeEntry *e) { return e->count; }
struct AtNodeIterator { const AtUniverse *u; size_t i; };
inline AtNodeIterator *AiUniverseGetNodeIterator(const AtUniverse *u, unsigned) { return new AtNodeIterator{u, 0}; }
inline bool AiNodeIteratorFinished(const AtNodeIterator *it) { return it->i >= it->u->nodes.size(); }
inline AtNode *AiNodeIteratorGetNext(AtNodeIterator *it) { return it->u->nodes[it->i++]; }
inline void AiNodeIteratorDestroy(AtNodeIterator *it) { delete it; }
inline uint32_t AiArrayGetNumElements(const AtArray *a) { return (uint32_t)std::max(a->strs.size(), a->ptrs.size()); }
inline void *AiArrayGetPtr(const AtArray *a, uint32_t i) { return a->ptrs[i]; }
inline AtString AiArrayGetStr(const AtArray *a, uint32_t i) { return AtString(a->strs[i].c_str()); }
inline AtArray *AiArrayAllocate(uint32_t n, uint8_t, uint8_t) { AtArray *a = new AtArray(); a->strs.resize(n); return a; }
inline void AiArraySetStr(AtArray *a, uint32_t i, const char *s) { a->strs[i] = s; }
// pointer arrays (options.aov_shaders) grow through resize + set (lentil_operator.cpp:139-142,155-158); string arrays keep `strs`
inline void AiArrayResize(AtArray *a, uint32_t n, uint8_t) { if (!a->strs.empty()) a->strs.resize(n); else a->ptrs.resize(n, nullptr); }
inline void AiArraySetPtr(AtArray *a, uint32_t i, void *p) { if (i >= a->ptrs.size()) a->ptrs.resize(i + 1, nullptr); a->ptrs[i] = p; }
inline bool AiAOVRegister(const char *name, uint8_t, int) { if (AtUniverse *u = shim_default_universe()) u->registered_aovs.push_back(name); return true; }
// scene editing: a node of entry `type` called `name`, owned by the universe
inline AtNode *AiNode(AtUniverse *u, const AtString &type, const AtString &name = AtString(""), const AtNode * = nullptr) {
  u->created.emplace_back();
  AtNode *n = &u->created.back();
  n->name = name.c_str(); n->entry.name = type.c_str(); n->universe = u;
  u->nodes.push_back(n);
  return n;
}
inline void AiNodeSetStr(AtNode *n, const AtString &k, const AtString &v) { n->params[k.c_str()].s = v.c_str(); }
inline bool AiNodeLink(AtNode *src, const AtString &input, AtNode *target) { target->links[input.c_str()] = src; return true; }
inline const char *AiParamGetTypeName(uint8_t) { return "type"; }
inline float AiCameraGetShutterStart() { return 0.f; }
inline float AiCameraGetShutterEnd() { return 0.f; }
inline void AiCameraInitialize(AtNode *) {}
inline void AiCameraUpdate(AtNode *, bool) {}
inline void AiCameraToWorldMatrix(const AtNode *, float, AtMatrix &m) { m = AtMatrix(); }
inline void AiWorldToCameraMatrix(const AtNode *n, float, AtMatrix &m) { m = n ? n->world_to_camera : AtMatrix(); }
inline void AiFilterInitialize(AtNode *n, bool, const char **required) {
  n->required_aovs.clear();
  for (; required && *required; ++required) n->required_aovs.push_back(*required);
}
inline void AiFilterUpdate(AtNode *n, float width) { n->filter_width = width; }
inline void AiDriverInitialize(AtNode *, bool) {}
inline void AiMetaDataSetBool(AtNodeEntry *e, const char *param, const char *name, bool v) {
  if (e) e->meta.push_back(std::string(param ? param : "") + ":" + name + "=" + (v ? "true" : "false"));
}
inline void AiMetaDataSetStr(AtNodeEntry *e, const char *param, const AtString &name, const AtString &v) {
  if (e) e->meta.push_back(std::string(param ? param : "") + ":" + name.c_str() + "=" + v.c_str());
}

// ---- scene rays: there is no scene, nothing is ever occluded ---------------------------------------------
struct AtShaderGlobals { int dummy; };
struct AtRay { int dummy; };
struct AtScrSample { int dummy; };
inline AtShaderGlobals *AiShaderGlobals() { return new AtShaderGlobals(); }
inline void AiShaderGlobalsDestroy(AtShaderGlobals *sg) { delete sg; }
inline AtRay AiMakeRay(int, const AtVector &, const AtVector *, float, const AtShaderGlobals *) { return AtRay(); }
inline bool AiTraceProbe(const AtRay &, const AtShaderGlobals *) { return false; }
inline bool AiTrace(const AtRay &, const AtRGB &, AtScrSample &) { return false; }

// ---- textures served from memory (imagebokeh.h:83-107) -----------------------------------------------------
struct ShimTexture { unsigned w = 0, h = 0, c = 0; std::vector<float> px; };
inline std::map<std::string, ShimTexture> &shim_textures() { static std::map<std::string, ShimTexture> t; return t; }
inline bool AiTextureGetResolution(const AtString &p, unsigned *w, unsigned *h) {
  auto it = shim_textures().find(p.c_str());
  if (it == shim_textures().end()) return false;
  *w = it->second.w; *h = it->second.h;
  return true;
}
inline bool AiTextureGetNumChannels(const AtString &p, unsigned *c) {
  auto it = shim_textures().find(p.c_str());
  if (it == shim_textures().end()) return false;
  *c = it->second.c;
  return true;
}
inline bool AiTextureLoad(const AtString &p, bool, unsigned, void *dst) {
  auto it = shim_textures().find(p.c_str());
  if (it == shim_textures().end()) return false;
  memcpy(dst, it->second.px.data(), it->second.px.size() * sizeof(float));
  return true;
}

// ---- AOV sample iterator over caller-supplied arrays ---------------------------------------------------------
struct ShimAovArray { const float *data = nullptr; int comps = 4; };  // [n][4] floats; comps = how many are meaningful
struct AtAOVSampleIterator {
  AtString aov_name;   // the AOV this filter call is for
  int px = 0, py = 0;  // pixel
  size_t begin = 0, end = 0;
  long cur = -1;
  float inv_density = 0.f;
  const float *rgba = nullptr;  // [n][4]
  std::map<std::string, ShimAovArray> aovs;
  // depth sub-samples (AiAOVSampleIteratorGetNextDepth): per sample `depth_n` slots, of which depth_count[i] are valid
  int depth_n = 0, depth_cur = -1;
  const uint8_t *depth_count = nullptr;
  const float *depth_opacity = nullptr;              // [n][depth_n] grey opacity, returned as an RGB of three equal values
  std::map<std::string, const float *> depth_ids;    // AOV name -> [n][depth_n]
  float depth_tmp[4] = {0, 0, 0, 0};
};
inline AtString AiAOVSampleIteratorGetAOVName(const AtAOVSampleIterator *it) { return it->aov_name; }
inline void AiAOVSampleIteratorGetPixel(const AtAOVSampleIterator *it, int &x, int &y) { x = it->px; y = it->py; }
inline void AiAOVSampleIteratorReset(AtAOVSampleIterator *it) { it->cur = -1; it->depth_cur = -1; }
inline bool AiAOVSampleIteratorGetNext(AtAOVSampleIterator *it) {
  long next = it->cur < 0 ? (long)it->begin : it->cur + 1;
  if ((size_t)next >= it->end) { it->cur = (long)it->end; return false; }
  it->cur = next;
  return true;
}
// walks the depth sub-samples of the current sample; false once they are exhausted (the real iterator then sits on
// the next sample, which is why the reference re-seeks afterwards, lentil.h:806-808 -- here it stays put)
inline bool AiAOVSampleIteratorGetNextDepth(AtAOVSampleIterator *it) {
  if (it->depth_n <= 0 || it->cur < 0) return false;
  const int count = it->depth_count ? std::min<int>(it->depth_count[it->cur], it->depth_n) : it->depth_n;
  if (it->depth_cur + 1 < count) { ++it->depth_cur; return true; }
  it->depth_cur = -1;
  return false;
}
inline AtVector2 AiAOVSampleIteratorGetOffset(const AtAOVSampleIterator *) { return AtVector2(0.f, 0.f); }
inline float AiAOVSampleIteratorGetInvDensity(const AtAOVSampleIterator *it) { return it->inv_density; }
inline AtRGBA AiAOVSampleIteratorGetRGBA(const AtAOVSampleIterator *it) { const float *p = it->rgba + 4 * it->cur; return AtRGBA(p[0], p[1], p[2], p[3]); }
inline const float *shim_aov(const AtAOVSampleIterator *it, const AtString &n) {
  static const float zero[4] = {0, 0, 0, 0};
  if (it->depth_cur >= 0) {  // inside a depth walk: per-sub-sample values
    float *tmp = const_cast<float *>(it->depth_tmp);
    const size_t at = (size_t)it->depth_n * it->cur + it->depth_cur;
    if (std::string(n.c_str()) == "opacity") {
      const float o = it->depth_opacity ? it->depth_opacity[at] : 0.0f;
      tmp[0] = tmp[1] = tmp[2] = tmp[3] = o;
      return tmp;
    }
    auto d = it->depth_ids.find(n.c_str());
    if (d != it->depth_ids.end() && d->second) { tmp[0] = tmp[1] = tmp[2] = tmp[3] = d->second[at]; return tmp; }
  }
  auto f = it->aovs.find(n.c_str());
  return (f == it->aovs.end() || !f->second.data) ? zero : f->second.data + 4 * it->cur;
}
inline AtRGBA AiAOVSampleIteratorGetAOVRGBA(const AtAOVSampleIterator *it, const AtString &n) { const float *p = shim_aov(it, n); return AtRGBA(p[0], p[1], p[2], p[3]); }
inline AtRGB AiAOVSampleIteratorGetAOVRGB(const AtAOVSampleIterator *it, const AtString &n) { const float *p = shim_aov(it, n); return AtRGB(p[0], p[1], p[2]); }
inline AtVector AiAOVSampleIteratorGetAOVVec(const AtAOVSampleIterator *it, const AtString &n) { const float *p = shim_aov(it, n); return AtVector(p[0], p[1], p[2]); }
inline float AiAOVSampleIteratorGetAOVFlt(const AtAOVSampleIterator *it, const AtString &n) { return shim_aov(it, n)[0]; }
inline int AiAOVSampleIteratorGetAOVInt(const AtAOVSampleIterator *it, const AtString &n) { return (int)shim_aov(it, n)[0]; }
inline unsigned AiAOVSampleIteratorGetAOVUInt(const AtAOVSampleIterator *it, const AtString &n) { return (unsigned)shim_aov(it, n)[0]; }
inline void *AiAOVSampleIteratorGetAOVPtr(const AtAOVSampleIterator *, const AtString &) { return nullptr; }
inline AtRGB AiAOVSampleIteratorGetRGB(const AtAOVSampleIterator *it) { const float *p = it->rgba + 4 * it->cur; return AtRGB(p[0], p[1], p[2]); }
inline AtVector AiAOVSampleIteratorGetVec(const AtAOVSampleIterator *it) { const float *p = it->rgba + 4 * it->cur; return AtVector(p[0], p[1], p[2]); }
inline float AiAOVSampleIteratorGetFlt(const AtAOVSampleIterator *it) { return it->rgba[4 * it->cur]; }
inline int AiAOVSampleIteratorGetInt(const AtAOVSampleIterator *it) { return (int)it->rgba[4 * it->cur]; }
inline unsigned AiAOVSampleIteratorGetUInt(const AtAOVSampleIterator *it) { return (unsigned)it->rgba[4 * it->cur]; }

// ---- output iterator of the driver ------------------------------------------------------------------------------
struct ShimOutput { AtString name; int type; void *data; };
struct AtOutputIterator { std::vector<ShimOutput> outs; long cur = -1; };
inline void AiOutputIteratorReset(AtOutputIterator *it) { it->cur = -1; }
inline bool AiOutputIteratorGetNext(AtOutputIterator *it, AtString *name, int *type, const void **data) {
  if ((size_t)(it->cur + 1) >= it->outs.size()) return false;
  ++it->cur;
  *name = it->outs[it->cur].name; *type = it->outs[it->cur].type; *data = it->outs[it->cur].data;
  return true;
}

// ---- camera I/O, node method tables and the node-definition macros ------------------------------------------------
struct AtCameraInput { float sx, sy, dsx, dsy, lensx, lensy, relative_time; };
struct AtCameraOutput { AtVector origin, dir, dOdx, dOdy, dDdx, dDdy; AtRGB weight; };
struct AtList { int dummy; std::vector<std::string> decls; /* AiParameter* declarations, in order */ };
struct AtNodeMethods {
  void (*Parameters)(AtList *, AtNodeEntry *);
  void (*Initialize)(AtRenderSession *, AtNode *);
  void (*Update)(AtRenderSession *, AtNode *);
  void (*Finish)(AtNode *);
  bool (*PluginInitialize)(void **);
  void (*PluginCleanup)(void *);
  void (*CreateRay)(const AtNode *, const AtCameraInput &, AtCameraOutput &, uint16_t);
  bool (*ReverseRay)(const AtNode *, const AtVector &, float, AtVector2 &);
  uint8_t (*FilterOutputType)(const AtNode *, uint8_t);
  void (*FilterPixel)(AtNode *, AtAOVSampleIterator *, void *, uint8_t);
  void (*DriverProcessBucket)(AtNode *, AtOutputIterator *, AtAOVSampleIterator *, int, int, int, int, uint16_t);
  // operator nodes ([EXTERNAL] ai_operator.h of Arnold 7: init / cleanup / cook / post_cook)
  bool (*OperatorInit)(AtNode *, void **);
  bool (*OperatorCleanup)(AtNode *, void *);
  bool (*OperatorCook)(AtNode *, AtNode *, void *, const AtArray *, struct AtCookContext *);
  bool (*OperatorPostCook)(AtNode *, void *);
};
struct AtNodeLib { const AtNodeMethods *methods; const char *name; int node_type; uint8_t output_type; char version[64]; };

#define node_parameters static void Parameters(AtList *params, AtNodeEntry *nentry)
#define node_initialize static void Initialize(AtRenderSession *render_session, AtNode *node)
#define node_update static void Update(AtRenderSession *render_session, AtNode *node)
#define node_finish static void Finish(AtNode *node)
#define node_plugin_initialize static bool PluginInitialize(void **plugin_data)
#define node_plugin_cleanup static void PluginCleanup(void *plugin_data)
#define camera_create_ray static void CreateRay(const AtNode *node, const AtCameraInput &input, AtCameraOutput &output, uint16_t tid)
#define camera_reverse_ray static bool ReverseRay(const AtNode *node, const AtVector &Po, float relative_time, AtVector2 &Ps)
#define filter_output_type static uint8_t FilterOutputType(const AtNode *node, uint8_t input_type)
#define filter_pixel static void FilterPixel(AtNode *node, AtAOVSampleIterator *iterator, void *data_out, uint8_t data_type)
#define driver_supports_pixel_type static bool DriverSupportsPixelType(const AtNode *node, uint8_t pixel_type)
#define driver_extension static const char **DriverExtension()
#define driver_open static void DriverOpen(AtNode *node, AtOutputIterator *iterator, AtBBox2 display_window, AtBBox2 data_window, int bucket_size)
#define driver_needs_bucket static bool DriverNeedsBucket(AtNode *node, int bucket_xo, int bucket_yo, int bucket_size_x, int bucket_size_y, uint16_t tid)
#define driver_prepare_bucket static void DriverPrepareBucket(AtNode *node, int bucket_xo, int bucket_yo, int bucket_size_x, int bucket_size_y, uint16_t tid)
#define driver_process_bucket \
  static void DriverProcessBucket(AtNode *node, AtOutputIterator *iterator, AtAOVSampleIterator *sample_iterator, int bucket_xo, int bucket_yo, \
                                  int bucket_size_x, int bucket_size_y, uint16_t tid)
#define driver_write_bucket \
  static void DriverWriteBucket(AtNode *node, AtOutputIterator *iterator, AtAOVSampleIterator *sample_iterator, int bucket_xo, int bucket_yo, \
                                int bucket_size_x, int bucket_size_y, uint16_t tid)
#define driver_close static void DriverClose(AtNode *node, AtOutputIterator *iterator)

#define SHIM_COMMON_DECLS       \
  node_parameters;              \
  node_initialize;              \
  node_update;                  \
  node_finish;
#define AI_CAMERA_NODE_EXPORT_METHODS(tag)                                                                                        \
  SHIM_COMMON_DECLS node_plugin_initialize;                                                                                       \
  node_plugin_cleanup;                                                                                                            \
  camera_create_ray;                                                                                                              \
  camera_reverse_ray;                                                                                                             \
  static AtNodeMethods shim_mtds = {Parameters, Initialize, Update, Finish, PluginInitialize, PluginCleanup, CreateRay, ReverseRay, \
                                    nullptr,    nullptr,    nullptr};                                                             \
  const AtNodeMethods *tag = &shim_mtds;
#define AI_FILTER_NODE_EXPORT_METHODS(tag)                                                                                       \
  SHIM_COMMON_DECLS filter_output_type;                                                                                          \
  filter_pixel;                                                                                                                  \
  static AtNodeMethods shim_mtds = {Parameters, Initialize, Update, Finish, nullptr, nullptr, nullptr, nullptr, FilterOutputType, \
                                    FilterPixel, nullptr};                                                                       \
  const AtNodeMethods *tag = &shim_mtds
#define AI_DRIVER_NODE_EXPORT_METHODS(tag)                                                                                                 \
  SHIM_COMMON_DECLS driver_supports_pixel_type;                                                                                            \
  driver_extension;                                                                                                                        \
  driver_open;                                                                                                                             \
  driver_needs_bucket;                                                                                                                     \
  driver_prepare_bucket;                                                                                                                   \
  driver_process_bucket;                                                                                                                   \
  driver_write_bucket;                                                                                                                     \
  driver_close;                                                                                                                            \
  static void *shim_unused_driver[] = {(void *)DriverSupportsPixelType, (void *)DriverExtension, (void *)DriverOpen, (void *)DriverNeedsBucket, \
                                       (void *)DriverPrepareBucket,     (void *)DriverWriteBucket, (void *)DriverClose};                   \
  static AtNodeMethods shim_mtds = {Parameters, Initialize, Update, Finish, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,          \
                                    DriverProcessBucket};                                                                                  \
  const AtNodeMethods *tag = (shim_unused_driver[0] ? &shim_mtds : &shim_mtds)

// operator nodes: the reference's operator only touches `op` (lentil_operator.cpp:19-181); [EXTERNAL] signatures of Arnold 7
struct AtCookContext { int dummy; };
#define operator_init static bool OperatorInit(AtNode *op, void **user_data)
#define operator_cleanup static bool OperatorCleanup(AtNode *op, void *user_data)
#define operator_cook static bool OperatorCook(AtNode *node, AtNode *op, void *user_data, const AtArray *matching_params, AtCookContext *cook_context)
#define operator_post_cook static bool OperatorPostCook(AtNode *op, void *user_data)
#define AI_OPERATOR_NODE_EXPORT_METHODS(tag)                                                                                      \
  node_parameters;                                                                                                                \
  operator_init;                                                                                                                  \
  operator_cleanup;                                                                                                               \
  operator_cook;                                                                                                                  \
  operator_post_cook;                                                                                                             \
  static AtNodeMethods shim_mtds = {Parameters, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, \
                                    nullptr,    OperatorInit, OperatorCleanup, OperatorCook, OperatorPostCook};                   \
  const AtNodeMethods *tag = &shim_mtds
// the plugin's entry point: Arnold calls it with i = 0, 1, ... until it returns false (lentil_loader.cpp:20-28)
#define node_loader extern "C" AI_EXPORT_LIB bool NodeLoader(int i, AtNodeLib *node)

// parameter declaration macros: the real ones end in a semicolon (the reference omits it twice, lentil_camera.cpp:48-49)
inline std::string shim_text(const char *c) { return c ? c : ""; }
inline std::string shim_text(const AtString &a) { return a.c_str(); }
inline std::string shim_enum_values(const char **e) { std::string v; for (; e && *e; ++e) v += (v.empty() ? "" : ",") + std::string(*e); return v; }
inline std::string shim_flt(double d) { char b[64]; snprintf(b, sizeof b, "%.9g", d); return b; }
#define AiParameterEnum(n, d, e) params->decls.push_back("enum " + shim_text(n) + " default=" + std::to_string((int)(d)) + " values=" + shim_enum_values(e));
#define AiParameterInt(n, d) params->decls.push_back("int " + shim_text(n) + " default=" + std::to_string((int)(d)));
#define AiParameterFlt(n, d) params->decls.push_back("flt " + shim_text(n) + " default=" + shim_flt((float)(d)));
#define AiParameterBool(n, d) params->decls.push_back("bool " + shim_text(n) + " default=" + ((d) ? "true" : "false"));
#define AiParameterStr(n, d) params->decls.push_back("str " + shim_text(n) + " default=\"" + shim_text(d) + "\"");
