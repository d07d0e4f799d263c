// lentil_oracle.cpp — CPU restatement (double precision, scalar) of the reference's per-ray hot paths.
//
// TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA product is compared against; it is
// never linked into, imported by or called from the product (pota_b200/).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// Every function cites the reference code (/root/reference/src/...) it restates.  The per-lens
// polynomial bodies the reference #includes from the absent polynomial-optics checkout
// (lentil.h:1262,1278,1308,1576) are restated from the generator's published output format
// (loop shape pinned by tests/aperture_sampling_debug/writout.txt:10-40 and
// newton-w4.py:44) and evaluated from the lens pack tables (gen/lens_pack_data.inc) as flat
// monomial sums with lens_ipow, term by term, exactly as the generated code would.
//
// Parity status: PINNED against the reference's own sources compiled behind stand-in headers
// (oracle/ref.mk -> oracle/_ref/libref.so): tests/test_oracle_vs_ref.py asserts bit-identical results for
// the primitives, the setup solvers, the generated bodies, camera_create_ray, trace_ray_bw_po and whole
// filter_pixel + driver_process_bucket framebuffers; tests/golden/*.npz carry reference output to boxes
// without /root/reference.  Lens COEFFICIENTS are the build's own pack — "parity unpinned" at
// coefficient level, because the reference ships none.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../include/lentil_b200.h"
#include "gen/lens_pack_data.inc"
#include "oracle.h"

namespace {

// Arnold constants (ai_constants.h of the SDK the reference builds against)
constexpr float AI_PI_F = 3.14159265358979323846f;
constexpr float AI_BIG_F = 1.0e12f;
constexpr float AI_INFINITE_F = 1.0e30f;
constexpr float AI_EPSILON_F = 1.0e-4f;

enum { P_OUT_X, P_OUT_Y, P_OUT_DX, P_OUT_DY, P_OUT_T, P_AP_X, P_AP_Y, P_AP_DX, P_AP_DY };
enum { C_OUTER_R, C_INNER_R, C_LENGTH, C_BFL, C_EFL, C_AP_POS, C_AP_HOUSING, C_INNER_CURV, C_OUTER_CURV, C_FOV, C_FSTOP, C_AP_R_FSTOP };

struct V2 { double x, y; };
struct V3 { double x, y, z; };

// ---- global.h -------------------------------------------------------------------------------
inline float clamp_f(float in, const float mn, const float mx) {  // global.h:8-12
  if (in < mn) in = mn;
  if (in > mx) in = mx;
  return in;
}
inline float clamp_min_f(float in, const float mn) {  // global.h:15-18
  if (in < mn) in = mn;
  return in;
}
inline float linear_interpolate(float perc, float a, float b) { return a + perc * (b - a); }  // global.h:3-5

template <unsigned int N>
inline unsigned int tea(const unsigned int val0, const unsigned int val1) {  // global.h:32-46
  unsigned int v0 = val0, v1 = val1, s0 = 0;
  for (unsigned int n = 0; n < N; ++n) {
    s0 += 0x9e3779b9;
    v0 += ((v1 << 4) + 0xA341316C) ^ (v1 + s0) ^ ((v1 >> 5) + 0xC8013EA4);
    v1 += ((v0 << 4) + 0xAD90777D) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7E95761E);
  }
  return v0;
}
inline float rng(unsigned int &previous) {  // global.h:51-57
  previous = previous * 1664525u + 1013904223u;
  return float(previous & 0X00FFFFFF) / float(0x01000000u);
}
struct Xor128 {  // global.h:22-27, state made explicit
  uint32_t x = 123456789, y = 362436069, z = 521288629, w = 88675123;
  uint32_t next() {
    uint32_t t = x ^ (x << 11);
    x = y; y = z; z = w;
    return w = (w ^ (w >> 19) ^ t ^ (t >> 8));
  }
};

// ---- lens.h ---------------------------------------------------------------------------------
inline float fast_sin(float x) {  // lens.h:17-24 (AI_PI is a float constant: the fmod is evaluated in float)
  x = fmodf(x + AI_PI_F, AI_PI_F * 2) - AI_PI_F;
  const float B = 4.0f / AI_PI_F;
  const float C = -4.0f / (AI_PI_F * AI_PI_F);
  float y = B * x + C * x * std::abs(x);
  const float P = 0.225f;
  return P * (y * std::abs(y) - y) + y;
}
inline float fast_cos(float x) {  // lens.h:27-37 (x += AI_PI*0.5 promotes to double, rounds back to float)
  x += AI_PI_F * 0.5;
  x = fmodf(x + AI_PI_F, AI_PI_F * 2) - AI_PI_F;
  const float B = 4.0f / AI_PI_F;
  const float C = -4.0f / (AI_PI_F * AI_PI_F);
  float y = B * x + C * x * std::abs(x);
  const float P = 0.225f;
  return P * (y * std::abs(y) - y) + y;
}
inline double dot3(V3 u, V3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }  // lens.h:48-50
inline V3 cross3(V3 u, V3 v) { return {u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x}; }  // :52-56
inline void normalise3(V3 &v) {  // lens.h:58-61
  const double ilen = 1.0f / std::sqrt(dot3(v, v));
  v.x *= ilen; v.y *= ilen; v.z *= ilen;
}

void sphereToCs(V2 inpos, V2 indir, V3 &outpos, V3 &outdir, double center, double sphereRad) {  // lens.h:99-125
  const V3 normal{inpos.x / sphereRad, inpos.y / sphereRad,
                  std::sqrt(std::max(0.0, sphereRad * sphereRad - inpos.x * inpos.x - inpos.y * inpos.y)) / std::abs(sphereRad)};
  const V3 tempDir{indir.x, indir.y, std::sqrt(std::max(0.0, 1.0 - indir.x * indir.x - indir.y * indir.y))};
  V3 ex{normal.z, 0, -normal.x};
  normalise3(ex);
  V3 ey = cross3(normal, ex);
  outdir.x = tempDir.x * ex.x + tempDir.y * ey.x + tempDir.z * normal.x;
  outdir.y = tempDir.x * ex.y + tempDir.y * ey.y + tempDir.z * normal.y;
  outdir.z = tempDir.x * ex.z + tempDir.y * ey.z + tempDir.z * normal.z;
  outpos.x = inpos.x;
  outpos.y = inpos.y;
  outpos.z = normal.z * sphereRad + center;
}
void csToSphere(V3 inpos, V3 indir, V2 &outpos, V2 &outdir, double sphereCenter, double sphereRad) {  // lens.h:127-153
  const V3 normal{inpos.x / sphereRad, inpos.y / sphereRad, std::abs((inpos.z - sphereCenter) / sphereRad)};
  V3 tempDir = indir;
  normalise3(tempDir);
  V3 ex{normal.z, 0, -normal.x};
  normalise3(ex);
  V3 ey = cross3(normal, ex);
  outdir.x = dot3(tempDir, ex);
  outdir.y = dot3(tempDir, ey);
  outpos.x = inpos.x;
  outpos.y = inpos.y;
}
void csToCylinder(V3 inpos, V3 indir, V2 &outpos, V2 &outdir, double center, double R, bool cyl_y) {  // lens.h:156-185
  V3 normal{0, 0, 0};
  if (cyl_y) { normal.x = inpos.x / R; normal.z = std::abs((inpos.z - center) / R); }
  else       { normal.y = inpos.y / R; normal.z = std::abs((inpos.z - center) / R); }
  V3 tempDir = indir;
  normalise3(tempDir);
  V3 ex{normal.z, 0, -normal.x};
  V3 ey = cross3(normal, ex);
  normalise3(ey);
  outdir.x = dot3(tempDir, ex);
  outdir.y = dot3(tempDir, ey);
  outpos.x = inpos.x;
  outpos.y = inpos.y;
}
void cylinderToCs(V2 inpos, V2 indir, V3 &outpos, V3 &outdir, double center, double R, bool cyl_y) {  // lens.h:188-221
  V3 normal{0, 0, 0};
  if (cyl_y) { normal.x = inpos.x / R; normal.z = std::sqrt(std::max(0.0, R * R - inpos.x * inpos.x)) / std::abs(R); }
  else       { normal.y = inpos.y / R; normal.z = std::sqrt(std::max(0.0, R * R - inpos.y * inpos.y)) / std::abs(R); }
  const V3 tempDir{indir.x, indir.y, std::sqrt(std::max(0.0, 1.0 - indir.x * indir.x - indir.y * indir.y))};
  V3 ex{normal.z, 0, -normal.x};
  normalise3(ex);
  V3 ey = cross3(normal, ex);
  normalise3(ey);
  outdir.x = tempDir.x * ex.x + tempDir.y * ey.x + tempDir.z * normal.x;
  outdir.y = tempDir.x * ex.y + tempDir.y * ey.y + tempDir.z * normal.y;
  outdir.z = tempDir.x * ex.z + tempDir.y * ey.z + tempDir.z * normal.z;
  outpos.x = inpos.x;
  outpos.y = inpos.y;
  outpos.z = normal.z * R + center;
}
inline double lens_ipow(const double x, const int exp) {  // lens.h:226-233
  if (exp == 0) return 1.0f;
  if (exp == 1) return x;
  if (exp == 2) return x * x;
  const double p2 = lens_ipow(x, exp / 2);
  if (exp & 1) return x * p2 * p2;
  return p2 * p2;
}
void concentric_disk_sample(const double ox, const double oy, V2 &unit_disk, bool fast_trigo) {  // lens.h:309-333
  double phi, r;
  double a = 2.0 * ox - 1.0;
  double b = 2.0 * oy - 1.0;
  if ((a * a) > (b * b)) { r = a; phi = (0.78539816339) * (b / a); }
  else { r = b; phi = (M_PI / 2.0) - (0.78539816339) * (a / b); }
  if (!fast_trigo) { unit_disk.x = r * std::cos(phi); unit_disk.y = r * std::sin(phi); }
  else { unit_disk.x = r * fast_cos(phi); unit_disk.y = r * fast_sin(phi); }
}
std::vector<double> logarithmic_values() {  // lens.h:395-407
  double min = 0.0, max = 45.0, exponent = 2.0;
  std::vector<double> log;
  for (double i = -1.0; i <= 1.0; i += 0.0001) log.push_back((i < 0 ? -1 : 1) * std::pow(i, exponent) * (max - min) + min);
  return log;
}
V3 line_plane_intersection(V3 rayOrigin, V3 rayDirection) {  // lens.h:412-419
  V3 coord{100.0, 0.0, 100.0};
  V3 planeNormal{0.0, 1.0, 0.0};
  { double n = std::sqrt(dot3(rayDirection, rayDirection)); rayDirection = {rayDirection.x / n, rayDirection.y / n, rayDirection.z / n}; }
  { double n = std::sqrt(dot3(coord, coord)); coord = {coord.x / n, coord.y / n, coord.z / n}; }
  double s = (dot3(coord, planeNormal) - dot3(planeNormal, rayOrigin));
  double dn = dot3(planeNormal, rayDirection);
  // Eigen evaluates (rayDirection * s) / dn component-wise
  return {rayOrigin.x + (rayDirection.x * s) / dn, rayOrigin.y + (rayDirection.y * s) / dn, rayOrigin.z + (rayDirection.z * s) / dn};
}

// ---- thin-lens helpers, lens.h:477-582 (float arithmetic, as there) ---------------------------------
struct V3f { float x, y, z; };
inline float dot3f(V3f a, V3f b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3f normalize3f(V3f a) {  // AiV3Normalize
  float t = std::sqrt(dot3f(a, a));
  if (t != 0) t = 1 / t;
  return {a.x * t, a.y * t, a.z * t};
}
inline float ai_bias(float a, float b) {  // AiBias [EXTERNAL, as oracle/shims/ai.h]
  return (a > 0) ? ((b > 0) ? std::pow(a, std::log(b) / std::log(0.5f)) : 0) : 0;
}
void concentricDiskSample(float ox, float oy, V2 &lens, float bias, float squarelerp, float /*squeeze_x*/) {  // lens.h:477-517
  if (ox == 0.0 && oy == 0.0) { lens.x = 0.0; lens.y = 0.0; return; }
  float phi, r;
  const float a = 2.0 * ox - 1.0;
  const float b = 2.0 * oy - 1.0;
  if ((a * a) > (b * b)) { r = a; phi = 0.78539816339 * (b / a); }
  else { r = b; phi = (1.57079632679489661923f) - ((0.78539816339) * (a / b)); }
  if (bias != 0.5) r = ai_bias(std::abs(r), bias) * (r < 0 ? -1 : 1);
  const float cos_phi = fast_cos(phi);
  const float sin_phi = fast_sin(phi);
  lens.x = r * cos_phi;
  lens.y = r * sin_phi;
  if (squarelerp > 0.0) {
    lens.x = linear_interpolate(squarelerp, lens.x, a);
    lens.y = linear_interpolate(squarelerp, lens.y, b);
  }
}
bool empericalOpticalVignettingSquare(V3f origin, V3f direction, float apertureRadius, float opticalVignettingRadius,
                                      float opticalVignettingDistance, float squarebias) {  // lens.h:532-541
  float intersection = std::abs(opticalVignettingDistance / direction.z);
  V3f p{direction.x * intersection - origin.x, direction.y * intersection - origin.y, direction.z * intersection - origin.z};
  float power = 1.0 + squarebias;
  float radius = apertureRadius * opticalVignettingRadius;
  float dist = std::pow(std::abs(p.x), power) + std::pow(std::abs(p.y), power);
  return !(dist > std::pow(radius, power));
}
inline float lerp_squircle_mapping(float amount) { return 1.0 + std::log(1.0 + amount) * std::exp(amount * 3.0); }  // lens.h:544-546
inline void barrelDistortion(float &ux, float &uy, float distortion) {  // lens.h:548-551: uv *= 1. + dot(uv,uv)*distortion
  const float f = 1. + (ux * ux + uy * uy) * distortion;
  ux *= f; uy *= f;
}
inline void inverseBarrelDistortion(float &ux, float &uy, float distortion) {  // lens.h:553-562
  float b = distortion;
  float l = std::sqrt(ux * ux + uy * uy);
  float x0 = std::pow(9. * b * b * l + std::sqrt(3.) * std::sqrt(27. * b * b * b * b * l * l + 4. * b * b * b), 1. / 3.);
  float x = x0 / (std::pow(2., 1. / 3.) * std::pow(3., 2. / 3.) * b) - std::pow(2. / 3., 1. / 3.) / x0;
  const float f = x / l;
  ux *= f; uy *= f;
}
float abb_coma_multipliers(const float sensor_width, const float focal_length, const V3f dir_from_center, const V2 unit_disk) {  // lens.h:566-574
  const V3f maximal_perturbed_ray{(float)(1.0 * (sensor_width * 0.5)), (float)(1.0 * (sensor_width * 0.5)), -focal_length};
  float maximal_projection = dot3f(normalize3f(maximal_perturbed_ray), V3f{0.0f, 0.0f, -1.0f});
  float current_projection = dot3f(dir_from_center, V3f{0.0f, 0.0f, -1.0f});
  float projection_perc = ((current_projection - maximal_projection) / (1.0 - maximal_projection) - 0.5) * 2.0;
  float dist_from_sensor_center = 1.0 - projection_perc;
  float dist_from_aperture = std::sqrt(unit_disk.x * unit_disk.x + unit_disk.y * unit_disk.y);  // Eigen norm(), double -> float
  return dist_from_sensor_center * dist_from_aperture;
}
V3f abb_coma_perturb(const V3f dir_from_lens, const V3f ray_to_perturb, const float abb_coma, const bool reverse) {  // lens.h:578-586
  const V3f c{dir_from_lens.y * -1.0f - dir_from_lens.z * 0.0f, dir_from_lens.z * 0.0f - dir_from_lens.x * -1.0f,
              dir_from_lens.x * 0.0f - dir_from_lens.y * 0.0f};  // AiV3Cross(dir_from_lens, (0,0,-1))
  const V3f axis_tmp = normalize3f(c);
  const double ax = axis_tmp.x, ay = axis_tmp.y, az = axis_tmp.z;
  const double angle = (abb_coma * 2.3456 * AI_PI_F) / 180.0;
  const double co = std::cos(angle), si = std::sin(angle), t = 1.0 - co;
  double m[3][3] = {{t * ax * ax + co, t * ax * ay - si * az, t * ax * az + si * ay},
                    {t * ax * ay + si * az, t * ay * ay + co, t * ay * az - si * ax},
                    {t * ax * az - si * ay, t * ay * az + si * ax, t * az * az + co}};
  if (reverse) {  // rot.inverse(): general 3x3 inverse, as the Eigen stand-in of oracle/_ref does
    const double (*a)[3] = m;
    const double det = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                       a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
    double r[3][3];
    r[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) / det; r[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) / det;
    r[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) / det; r[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) / det;
    r[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) / det; r[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) / det;
    r[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) / det; r[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) / det;
    r[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) / det;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = r[i][j];
  }
  const double rx = ray_to_perturb.x, ry = ray_to_perturb.y, rz = ray_to_perturb.z;
  return {(float)(m[0][0] * rx + m[0][1] * ry + m[0][2] * rz), (float)(m[1][0] * rx + m[1][1] * ry + m[1][2] * rz),
          (float)(m[2][0] * rx + m[2][1] * ry + m[2][2] * rz)};
}

// ---- imagebokeh.h ---------------------------------------------------------------------------
struct arrayCompare {  // imagebokeh.h:21-27
  const float *values;
  explicit arrayCompare(const float *v) : values(v) {}
  bool operator()(int l, int r) const { return values[l] > values[r]; }
};
struct ImageData {  // imagebokeh.h:30-412 (Arnold texture I/O replaced by caller-supplied float pixels)
  int x = 0, y = 0, nchannels = 0;
  std::vector<float> pixelData, cdfRow, cdfColumn;
  std::vector<int> rowIndices, columnIndices;
  bool isValid() const { return (x * y * nchannels > 0 && nchannels >= 3); }  // :49-51
  bool read(const lb_bokeh_image *img) {  // :83-140
    if (!img || !img->pixels) return false;
    x = img->width; y = img->height; nchannels = img->channels;
    if (x != y) { x = y = nchannels = 0; return false; }  // :97-101
    pixelData.assign(img->pixels, img->pixels + (size_t)x * y * nchannels);
    bokehProbability();
    return true;
  }
  void bokehProbability() {  // :143-338
    if (!isValid()) return;
    int npixels = x * y;
    std::vector<float> pixelValues(npixels), normalizedPixelValues(npixels);
    int o1 = (nchannels >= 2 ? 1 : 0);
    int o2 = (nchannels >= 3 ? 2 : o1);
    float totalValue = 0.0f;
    for (int i = 0, j = 0; i < npixels; ++i, j += nchannels) {
      pixelValues[i] = pixelData[j] * 0.3f + pixelData[j + o1] * 0.59f + pixelData[j + o2] * 0.11f;
      totalValue += pixelValues[i];
    }
    float invTotalValue = 1.0f / totalValue;
    for (int i = 0; i < npixels; ++i) normalizedPixelValues[i] = pixelValues[i] * invTotalValue;
    std::vector<float> summedRowValues(y);
    for (int i = 0, k = 0; i < y; ++i) {
      summedRowValues[i] = 0.0f;
      for (int j = 0; j < x; ++j, ++k) summedRowValues[i] += normalizedPixelValues[k];
    }
    rowIndices.resize(y);
    for (int i = 0; i < y; ++i) rowIndices[i] = i;
    std::sort(rowIndices.begin(), rowIndices.end(), arrayCompare(summedRowValues.data()));
    cdfRow.resize(y);
    float prevVal = 0.0f;
    for (int i = 0; i < y; ++i) { cdfRow[i] = prevVal + summedRowValues[rowIndices[i]]; prevVal = cdfRow[i]; }
    std::vector<float> normalizedValuesPerRow(npixels);
    for (int r = 0, i = 0; r < y; ++r)
      for (int c = 0; c < x; ++c, ++i) {
        if ((normalizedPixelValues[i] != 0) && (summedRowValues[r] != 0)) normalizedValuesPerRow[i] = normalizedPixelValues[i] / summedRowValues[r];
        else normalizedValuesPerRow[i] = 0;
      }
    columnIndices.resize(npixels);
    for (int i = 0; i < npixels; i++) columnIndices[i] = i;
    for (int i = 0; i < npixels; i += x) std::sort(columnIndices.begin() + i, columnIndices.begin() + i + x, arrayCompare(normalizedValuesPerRow.data()));
    cdfColumn.resize(npixels);
    for (int r = 0, i = 0; r < y; ++r) {
      prevVal = 0.0f;
      for (int c = 0; c < x; ++c, ++i) { cdfColumn[i] = prevVal + normalizedValuesPerRow[columnIndices[i]]; prevVal = cdfColumn[i]; }
    }
  }
  void bokehSample(float randomNumberRow, float randomNumberColumn, V2 &lens) const {  // :341-412
    if (!isValid()) { lens.x = 0.0; lens.y = 0.0; return; }
    const float *pUpperBound = std::upper_bound(cdfRow.data(), cdfRow.data() + y, randomNumberRow);
    int r = 0;
    pUpperBound >= (cdfRow.data() + y) ? r = y - 1 : r = static_cast<int>(pUpperBound - cdfRow.data());
    int actualPixelRow = rowIndices[r];
    int recalulatedPixelRow = actualPixelRow - ((x - 1) / 2);
    int startPixel = actualPixelRow * x;
    const float *pUpperBoundColumn = std::upper_bound(cdfColumn.data() + startPixel, cdfColumn.data() + startPixel + x, randomNumberColumn);
    int c = 0;
    pUpperBoundColumn >= cdfColumn.data() + startPixel + x ? c = startPixel + x - 1 : c = static_cast<int>(pUpperBoundColumn - cdfColumn.data());
    int actualPixelColumn = columnIndices[c];
    int relativePixelColumn = actualPixelColumn - startPixel;
    int recalulatedPixelColumn = relativePixelColumn - ((y - 1) / 2);
    float flippedRow = static_cast<float>(recalulatedPixelColumn);
    float flippedColumn = recalulatedPixelRow * -1.0f;
    lens.x = static_cast<float>(flippedRow) / static_cast<float>(x) * 2.0;
    lens.y = static_cast<float>(flippedColumn) / static_cast<float>(y) * 2.0;
  }
};

// ---- generated lens code (L0), table form ------------------------------------------------------
struct Term { double c; unsigned char e[5]; };
using Poly = std::vector<Term>;

// value of one generated monomial sum: `+ c *x*lens_ipow(y, 2)... + ...` evaluated left to right
inline double poly_eval(const Poly &p, const double v[5]) {
  double acc = 0.0;
  bool first = true;
  for (const Term &t : p) {
    double m = t.c;
    for (int k = 0; k < 5; ++k) {
      if (t.e[k] == 1) m *= v[k];
      else if (t.e[k] > 1) m *= lens_ipow(v[k], t.e[k]);
    }
    if (first) { acc = +m; first = false; } else acc += m;
  }
  return acc;
}
Poly poly_derivative(const Poly &p, int var) {
  Poly d;
  for (const Term &t : p) {
    if (t.e[var] == 0) continue;
    Term n = t;
    n.c = t.c * t.e[var];
    n.e[var] -= 1;
    d.push_back(n);
  }
  return d;
}

}  // namespace

// ===============================================================================================
struct AOV {
  std::string name;
  int filter = LB_FILTER_GAUSSIAN;
  int role = LB_AOV_PLAIN;
  std::vector<float> buffer;  // RGBA, AOVData::buffer (aov_data.h:114-164); crypto AOVs: x = crypto_total_weight (:128)
  std::vector<std::map<float, float>> crypto_hash_map;  // aov_data.h:127
};

struct orc_camera {
  // ---- struct Camera, lentil.h:92-207 (hot-path members) ----
  int lensModel = 0, unitModel = LB_UNITS_CM, cameraType = LB_CAMERA_THINLENS;
  ImageData image;
  std::vector<float> zbuffer, zbuffer_debug, filter_weight_buffer;
  std::vector<AOV> aovs;
  bool has_crypto = false;  // cryptomatte_lentil, lentil.h:193
  double lens_outer_pupil_radius = 0, lens_inner_pupil_radius = 0, lens_length = 0, lens_back_focal_length = 0;
  double lens_effective_focal_length = 0, lens_aperture_pos = 0, lens_aperture_housing_radius = 0;
  double lens_inner_pupil_curvature_radius = 0, lens_outer_pupil_curvature_radius = 0, lens_field_of_view = 0;
  double lens_fstop = 0, lens_aperture_radius_at_fstop = 0;
  int lens_inner_pupil_geometry = 0, lens_outer_pupil_geometry = 0;  // 0 spherical, 1 cyl-y, 2 cyl-x
  double focus_distance = 0, sensor_width = 0, input_fstop = 0;
  bool enable_dof = true;
  int vignetting_retries = 15, bokeh_aperture_blades = 0;
  bool bokeh_enable_image = false;
  int bidir_sample_mult = 5;
  double bidir_add_energy_minimum_luminance = 2;
  float bidir_add_energy = 0, bidir_add_energy_transition = 1;
  bool enable_bidir_transmission = false, enable_skydome = false;
  float exposure = 1;
  double lambda = 0.55;
  float extra_sensor_shift = 0, focal_length = 35;
  float abb_chromatic = 0;
  int abb_chromatic_type = 0;
  float optical_vignetting_distance = 0, optical_vignetting_radius = 1.0f;
  float abb_spherical = 0.5f, abb_coma = 0, abb_distortion = 0, circle_to_square = 0.01f, bokeh_anamorphic = 1.0f;
  float fov = 0;
  double tan_fov = 0, aperture_radius = 0, sensor_shift = 0;
  unsigned xres = 0, yres = 0, xres_without_region = 0, yres_without_region = 0;
  int region_min_x = 0, region_min_y = 0;
  bool focus_check_ok = false;
  double focus_check_distance = 0;

  // lens pack tables
  Poly P[LP_POLY_COUNT];
  Poly dap[2][2];   // d ap_{x,y} / d {dx,dy}   -> dx1_domega0
  Poly dout[2][2];  // d out_{dx,dy} / d {x,y}  -> domega2_dx0
  // instrumentation (not in the reference): Newton iteration counters
  uint64_t fw_newton_its = 0, fw_traces = 0, bw_newton_its = 0, bw_attempts = 0;
  lb_filter_stats stats{};

  // -------------------------------------------------------------------------------------------
  bool load_lens(int model) {
    if (model < 0 || model >= LP_LENS_COUNT) return false;
    const LpLens &L = LP_LENSES[model];
    for (int p = 0; p < LP_POLY_COUNT; ++p) {
      P[p].clear();
      for (int i = 0; i < L.cnt[p]; ++i) {
        Term t;
        t.c = LP_COEF[L.off[p] + i];
        memcpy(t.e, LP_EXP[L.off[p] + i], 5);
        P[p].push_back(t);
      }
    }
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) {
        dap[i][j] = poly_derivative(P[P_AP_X + i], 2 + j);
        dout[i][j] = poly_derivative(P[P_OUT_DX + i], j);
      }
    // lens_constants.h body (lentil.h:1575-1577)
    lens_outer_pupil_radius = L.c[C_OUTER_R];
    lens_inner_pupil_radius = L.c[C_INNER_R];
    lens_length = L.c[C_LENGTH];
    lens_back_focal_length = L.c[C_BFL];
    lens_effective_focal_length = L.c[C_EFL];
    lens_aperture_pos = L.c[C_AP_POS];
    lens_aperture_housing_radius = L.c[C_AP_HOUSING];
    lens_inner_pupil_curvature_radius = L.c[C_INNER_CURV];
    lens_outer_pupil_curvature_radius = L.c[C_OUTER_CURV];
    lens_field_of_view = L.c[C_FOV];
    lens_fstop = L.c[C_FSTOP];
    lens_aperture_radius_at_fstop = L.c[C_AP_R_FSTOP];
    lens_outer_pupil_geometry = L.outer_geom;
    lens_inner_pupil_geometry = L.inner_geom;
    return true;
  }

  // lentil.h:1257-1266 + generated pt_evaluate.h
  double lens_evaluate(const double in[5], double out[5]) {
    double out_transmittance = 0.0;
    out[0] = poly_eval(P[P_OUT_X], in);
    out[1] = poly_eval(P[P_OUT_Y], in);
    out[2] = poly_eval(P[P_OUT_DX], in);
    out[3] = poly_eval(P[P_OUT_DY], in);
    out_transmittance = poly_eval(P[P_OUT_T], in);
    return std::max(0.0, out_transmittance);
  }

  // lentil.h:1272-1291 + generated pt_sample_aperture.h
  void lens_pt_sample_aperture(double in[5], double out[5], double dist) {
    double out_x = out[0], out_y = out[1], out_dx = out[2], out_dy = out[3];
    double x = in[0], y = in[1], dx = in[2], dy = in[3], lambda_ = in[4];
    double pred_x, pred_y, pred_dx = 0, pred_dy = 0;
    double sqr_err = FLT_MAX;
    for (int k = 0; k < 5 && sqr_err > 1e-4; k++) {
      const double begin[5] = {x + dist * dx, y + dist * dy, dx, dy, lambda_};
      pred_x = poly_eval(P[P_AP_X], begin);
      pred_y = poly_eval(P[P_AP_Y], begin);
      pred_dx = poly_eval(P[P_AP_DX], begin);
      pred_dy = poly_eval(P[P_AP_DY], begin);
      double J[2][2];
      J[0][0] = poly_eval(dap[0][0], begin) + 0.0f;
      J[0][1] = poly_eval(dap[0][1], begin) + 0.0f;
      J[1][0] = poly_eval(dap[1][0], begin) + 0.0f;
      J[1][1] = poly_eval(dap[1][1], begin) + 0.0f;
      double invJ[2][2];
      const double invdet = 1.0f / (J[0][0] * J[1][1] - J[0][1] * J[1][0]);
      invJ[0][0] = J[1][1] * invdet;
      invJ[1][1] = J[0][0] * invdet;
      invJ[0][1] = -J[0][1] * invdet;
      invJ[1][0] = -J[1][0] * invdet;
      const double dx1[2] = {out_x - pred_x, out_y - pred_y};
      for (int i = 0; i < 2; i++) {
        dx += invJ[0][i] * dx1[i];
        dy += invJ[1][i] * dx1[i];
      }
      sqr_err = dx1[0] * dx1[0] + dx1[1] * dx1[1];
      ++fw_newton_its;
    }
    out_dx = pred_dx;
    out_dy = pred_dy;
    out[0] = out_x; out[1] = out_y; out[2] = out_dx; out[3] = out_dy;
    in[0] = x; in[1] = y; in[2] = dx; in[3] = dy;
  }

  // lentil.h:1296-1313 + generated lt_sample_aperture.h
  double lens_lt_sample_aperture(const V3 scene, const V2 ap, double sensor[5], double out[5], const double lambda_) {
    const double scene_x = scene.x, scene_y = scene.y, scene_z = scene.z;
    const double ap_x = ap.x, ap_y = ap.y;
    double x = 0, y = 0, dx = 0, dy = 0;
    const double R = lens_outer_pupil_curvature_radius;
    int error = 0;
    {
      const double eps = 1e-8;
      double sqr_err = 1e30, sqr_ap_err = 1e30;
      double prev_sqr_err = 1e32, prev_sqr_ap_err = 1e32;
      for (int k = 0; k < 100 && (sqr_err > eps || sqr_ap_err > eps) && error == 0; k++) {
        prev_sqr_err = sqr_err, prev_sqr_ap_err = sqr_ap_err;
        const double begin[5] = {x, y, dx, dy, lambda_};
        const double pred_ap[2] = {poly_eval(P[P_AP_X], begin), poly_eval(P[P_AP_Y], begin)};
        const double delta_ap[2] = {ap_x - pred_ap[0], ap_y - pred_ap[1]};
        sqr_ap_err = delta_ap[0] * delta_ap[0] + delta_ap[1] * delta_ap[1];
        double J[2][2];
        J[0][0] = poly_eval(dap[0][0], begin) + 0.0f;
        J[0][1] = poly_eval(dap[0][1], begin) + 0.0f;
        J[1][0] = poly_eval(dap[1][0], begin) + 0.0f;
        J[1][1] = poly_eval(dap[1][1], begin) + 0.0f;
        double invApJ[2][2];
        const double invdetap = 1.0f / (J[0][0] * J[1][1] - J[0][1] * J[1][0]);
        invApJ[0][0] = J[1][1] * invdetap;
        invApJ[1][1] = J[0][0] * invdetap;
        invApJ[0][1] = -J[0][1] * invdetap;
        invApJ[1][0] = -J[1][0] * invdetap;
        for (int i = 0; i < 2; i++) {
          dx += invApJ[0][i] * delta_ap[i];
          dy += invApJ[1][i] * delta_ap[i];
        }
        out[0] = poly_eval(P[P_OUT_X], begin);
        out[1] = poly_eval(P[P_OUT_Y], begin);
        out[2] = poly_eval(P[P_OUT_DX], begin);
        out[3] = poly_eval(P[P_OUT_DY], begin);
        V3 pred_out_cs_pos{0, 0, 0}, pred_out_cs_dir{0, 0, 0};
        V2 outpos{out[0], out[1]}, outdir{out[2], out[3]};
        if (lens_outer_pupil_geometry == 1) cylinderToCs(outpos, outdir, pred_out_cs_pos, pred_out_cs_dir, -R, R, true);
        else if (lens_outer_pupil_geometry == 2) cylinderToCs(outpos, outdir, pred_out_cs_pos, pred_out_cs_dir, -R, R, false);
        else sphereToCs(outpos, outdir, pred_out_cs_pos, pred_out_cs_dir, -R, R);
        V3 view{scene_x - pred_out_cs_pos.x, scene_y - pred_out_cs_pos.y, scene_z - pred_out_cs_pos.z};
        normalise3(view);
        V2 out_new_pos{0, 0}, out_new_dir{0, 0};
        if (lens_outer_pupil_geometry == 1) csToCylinder(pred_out_cs_pos, view, out_new_pos, out_new_dir, -R, R, true);
        else if (lens_outer_pupil_geometry == 2) csToCylinder(pred_out_cs_pos, view, out_new_pos, out_new_dir, -R, R, false);
        else csToSphere(pred_out_cs_pos, view, out_new_pos, out_new_dir, -R, R);
        const double delta_out[2] = {out_new_dir.x - out[2], out_new_dir.y - out[3]};
        sqr_err = delta_out[0] * delta_out[0] + delta_out[1] * delta_out[1];
        double K[2][2];
        K[0][0] = poly_eval(dout[0][0], begin) + 0.0f;
        K[0][1] = poly_eval(dout[0][1], begin) + 0.0f;
        K[1][0] = poly_eval(dout[1][0], begin) + 0.0f;
        K[1][1] = poly_eval(dout[1][1], begin) + 0.0f;
        double invJ[2][2];
        const double invdet = 1.0f / (K[0][0] * K[1][1] - K[0][1] * K[1][0]);
        invJ[0][0] = K[1][1] * invdet;
        invJ[1][1] = K[0][0] * invdet;
        invJ[0][1] = -K[0][1] * invdet;
        invJ[1][0] = -K[1][0] * invdet;
        for (int i = 0; i < 2; i++) {
          x += 0.72 * invJ[0][i] * delta_out[i];
          y += 0.72 * invJ[1][i] * delta_out[i];
        }
        if (sqr_err > prev_sqr_err) error |= 1;
        if (sqr_ap_err > prev_sqr_ap_err) error |= 2;
        if (out[0] != out[0]) error |= 4;
        if (out[0] * out[0] + out[1] * out[1] > lens_outer_pupil_radius * lens_outer_pupil_radius) error |= 16;
        if (k < 10) error = 0;  // "error reset (k<10)", writout.txt:40
        ++bw_newton_its;
      }
    }
    if (out[0] * out[0] + out[1] * out[1] > lens_outer_pupil_radius * lens_outer_pupil_radius) error |= 16;
    const double begin[5] = {x, y, dx, dy, lambda_};
    if (error == 0) out[4] = poly_eval(P[P_OUT_T], begin);
    else out[4] = 0.0f;
    sensor[0] = x; sensor[1] = y; sensor[2] = dx; sensor[3] = dy; sensor[4] = lambda_;
    return std::max(0.0, out[4]);
  }

  void outer_to_cs(V2 outpos, V2 outdir, V3 &pos, V3 &dir) {  // lentil.h:387-389
    const double R = lens_outer_pupil_curvature_radius;
    if (lens_outer_pupil_geometry == 1) cylinderToCs(outpos, outdir, pos, dir, -R, R, true);
    else if (lens_outer_pupil_geometry == 2) cylinderToCs(outpos, outdir, pos, dir, -R, R, false);
    else sphereToCs(outpos, outdir, pos, dir, -R, R);
  }

  // lentil.h:964-982
  void lens_sample_triangular_aperture(double &x, double &y, double r1, double r2, const double radius, const int blades) {
    const int tri = (int)(r1 * blades);
    r1 = r1 * blades - tri;
    double a = std::sqrt(r1);
    double b = (1.0f - r2) * a;
    double c = r2 * a;
    double p1[2], p2[2];
    { double phi = 2.0f * AI_PI_F / blades * (tri + 1); p1[0] = std::sin(phi); p1[1] = std::cos(phi); }
    { double phi = 2.0f * AI_PI_F / blades * tri; p2[0] = std::sin(phi); p2[1] = std::cos(phi); }
    x = radius * (b * p1[1] + c * p2[1]);
    y = radius * (b * p1[0] + c * p2[0]);
  }

  // lentil.h:283-427.  Retry RNG: the reference re-draws r1,r2 from the process-global xor128
  // (lentil.h:313-316); here, as in the product, from tea<8>(ray_id, tries) + rng (documented deviation).
  void trace_ray_fw_po(int &tries, const double sx, const double sy, float origin[3], float direction[3], float weight[3],
                       double &r1, double &r2, const bool deriv_ray, uint32_t ray_id) {
    tries = 0;
    bool ray_succes = false;
    double sensor[5] = {0, 0, 0, 0, 0}, aperture[5] = {0, 0, 0, 0, 0}, out[5] = {0, 0, 0, 0, 0};
    while (!ray_succes && tries <= vignetting_retries) {
      sensor[0] = sx * (sensor_width * 0.5);
      sensor[1] = sy * (sensor_width * 0.5);
      sensor[2] = sensor[3] = 0.0;
      sensor[4] = lambda;
      for (double &v : aperture) v = 0;
      for (double &v : out) v = 0;
      V2 unit_disk{0.0, 0.0};
      if (enable_dof) {
        if (!deriv_ray && tries > 0) {
          unsigned int seed = tea<8>(ray_id, (unsigned int)tries);
          r1 = rng(seed);
          r2 = rng(seed);
        }
        if (bokeh_enable_image) image.bokehSample(r1, r2, unit_disk);
        else if (bokeh_aperture_blades < 2) concentric_disk_sample(r1, r2, unit_disk, true);
        else lens_sample_triangular_aperture(unit_disk.x, unit_disk.y, r1, r2, 1.0, bokeh_aperture_blades);
      }
      aperture[0] = unit_disk.x * aperture_radius;
      aperture[1] = unit_disk.y * aperture_radius;
      if (enable_dof) lens_pt_sample_aperture(sensor, aperture, sensor_shift);
      sensor[0] += sensor[2] * sensor_shift;
      sensor[1] += sensor[3] * sensor_shift;
      double transmittance = lens_evaluate(sensor, out);
      ++fw_traces;
      if (transmittance <= 0.0) { ++tries; continue; }
      if (out[0] * out[0] + out[1] * out[1] > lens_outer_pupil_radius * lens_outer_pupil_radius) { ++tries; continue; }
      const double px = sensor[0] + sensor[2] * lens_back_focal_length;
      const double py = sensor[1] + sensor[3] * lens_back_focal_length;
      if (px * px + py * py > lens_inner_pupil_radius * lens_inner_pupil_radius) { ++tries; continue; }
      ray_succes = true;
    }
    if (ray_succes == false) weight[0] = weight[1] = weight[2] = 0.0f;
    V3 cs_origin{0, 0, 0}, cs_direction{0, 0, 0};
    outer_to_cs(V2{out[0], out[1]}, V2{out[2], out[3]}, cs_origin, cs_direction);
    origin[0] = (float)cs_origin.x; origin[1] = (float)cs_origin.y; origin[2] = (float)cs_origin.z;
    direction[0] = (float)cs_direction.x; direction[1] = (float)cs_direction.y; direction[2] = (float)cs_direction.z;
    float s = 1.0f;  // AtVector *= double literal: float * (float)literal, lentil.h:395-416
    switch (unitModel) {
      case LB_UNITS_MM: s = (float)-1.0; break;
      case LB_UNITS_CM: s = (float)-0.1; break;
      case LB_UNITS_DM: s = (float)-0.01; break;
      case LB_UNITS_M: s = (float)-0.001; break;
    }
    for (int k = 0; k < 3; ++k) { origin[k] *= s; direction[k] *= s; }
    {  // AiV3Normalize (float): v / length, zero vector if length == 0
      float len = std::sqrt(direction[0] * direction[0] + direction[1] * direction[1] + direction[2] * direction[2]);
      if (len != 0.0f) { float inv = 1.0f / len; direction[0] *= inv; direction[1] *= inv; direction[2] *= inv; }
      else direction[0] = direction[1] = direction[2] = 0.0f;
    }
    if (origin[0] != origin[0] || origin[1] != origin[1] || origin[2] != origin[2] || direction[0] != direction[0] ||
        direction[1] != direction[1] || direction[2] != direction[2])
      weight[0] = weight[1] = weight[2] = 0.0f;
  }

  inline float get_image_dist_focusdist_thinlens() {  // lentil.h:664-666
    return (-focal_length * -focus_distance) / (-focal_length + -focus_distance);
  }
  inline float get_image_dist_focusdist_thinlens_abberated(const float shift) {  // lentil.h:668-670
    return (-focal_length * -(focus_distance + shift)) / (-focal_length + -(focus_distance + shift));
  }

  // unit disk sample of the thin-lens paths (lentil.h:457-472, lentil_filter.cpp:316-322)
  void thinlens_unit_disk(double r1, double r2, V2 &unit_disk) {
    if (bokeh_enable_image) image.bokehSample(r1, r2, unit_disk);
    else if (bokeh_aperture_blades < 2) concentricDiskSample(r1, r2, unit_disk, abb_spherical, circle_to_square, bokeh_anamorphic);
    else lens_sample_triangular_aperture(unit_disk.x, unit_disk.y, r1, r2, 1.0, bokeh_aperture_blades);
  }

  // lentil.h:431-569.  Same retry-RNG substitution as trace_ray_fw_po.
  void trace_ray_fw_thinlens(int &tries, const double sx, const double sy, float origin[3], float dir[3], float weight[3], double &r1,
                             double &r2, const bool deriv_ray, uint32_t ray_id) {
    tries = 0;
    bool ray_succes = false;
    while (!ray_succes && tries <= vignetting_retries) {
      float s0 = sx, s1 = sy;  // AtVector s(sx, sy, 0.0)
      if (abb_distortion > 0.0) { s0 = sx; s1 = sy; barrelDistortion(s0, s1, abb_distortion); }
      const V3f p{(float)(s0 * (sensor_width * 0.5)), (float)(s1 * (sensor_width * 0.5)), -focal_length};
      V3f dir_from_center = normalize3f(p);
      V2 unit_disk{0, 0};
      if (enable_dof) {
        if (!deriv_ray && tries > 0) {
          unsigned int seed = tea<8>(ray_id, (unsigned int)tries);
          r1 = rng(seed);
          r2 = rng(seed);
        }
        thinlens_unit_disk(r1, r2, unit_disk);
      }
      unit_disk.x *= bokeh_anamorphic;
      float abb_field_curvature = 0.0;
      V3f lens{(float)(unit_disk.x * aperture_radius), (float)(unit_disk.y * aperture_radius), 0.0f};
      const float intersection = std::abs(focus_distance / linear_interpolate(abb_field_curvature, dir_from_center.z, 1.0));
      const V3f focusPoint{dir_from_center.x * intersection, dir_from_center.y * intersection, dir_from_center.z * intersection};
      V3f dir_from_lens = normalize3f(V3f{focusPoint.x - lens.x, focusPoint.y - lens.y, focusPoint.z - lens.z});
      float abb_coma_multiplied = abb_coma * abb_coma_multipliers(sensor_width, focal_length, dir_from_center, unit_disk);
      dir_from_lens = abb_coma_perturb(dir_from_lens, dir_from_lens, abb_coma_multiplied, false);
      if (optical_vignetting_distance > 0.0 && !deriv_ray) {
        if (!empericalOpticalVignettingSquare(lens, dir_from_lens, aperture_radius, optical_vignetting_radius, optical_vignetting_distance,
                                              lerp_squircle_mapping(circle_to_square))) {
          ++tries;
          continue;
        }
      }
      origin[0] = lens.x; origin[1] = lens.y; origin[2] = lens.z;
      dir[0] = dir_from_lens.x; dir[1] = dir_from_lens.y; dir[2] = dir_from_lens.z;
      float s = 1.0f;  // lentil.h:538-560
      switch (unitModel) {
        case LB_UNITS_MM: s = (float)10.0; break;
        case LB_UNITS_CM: s = (float)1.0; break;
        case LB_UNITS_DM: s = (float)0.1; break;
        case LB_UNITS_M: s = (float)0.01; break;
      }
      for (int k = 0; k < 3; ++k) { origin[k] *= s; dir[k] *= s; }
      ray_succes = true;
    }
    const V3f d = normalize3f(V3f{dir[0], dir[1], dir[2]});
    dir[0] = d.x; dir[1] = d.y; dir[2] = d.z;
    if (!ray_succes) weight[0] = weight[1] = weight[2] = 0.0f;
  }

  void trace_ray_fw(int &tries, const double sx, const double sy, float origin[3], float direction[3], float weight[3], double &r1,
                    double &r2, const bool deriv_ray, uint32_t ray_id) {  // lentil_camera.cpp:90-94
    if (cameraType == LB_CAMERA_THINLENS) trace_ray_fw_thinlens(tries, sx, sy, origin, direction, weight, r1, r2, deriv_ray, ray_id);
    else trace_ray_fw_po(tries, sx, sy, origin, direction, weight, r1, r2, deriv_ray, ray_id);
  }

  // lentil_camera.cpp:78-125
  void camera_create_ray(float sx, float sy, float dsx, float dsy, float lensx, float lensy, uint32_t ray_id, float *o /*21*/, int *tries_out) {
    int tries = 0;
    double r1 = lensx, r2 = lensy;
    const float step = 0.001;
    float origin[3] = {0, 0, 0}, direction[3] = {0, 0, 0}, weight[3] = {1, 1, 1};
    trace_ray_fw(tries, sx, sy, origin, direction, weight, r1, r2, false, ray_id);
    if (tries_out) *tries_out = tries;
    float input_dx_sx = sx + (dsx * step);
    float input_dx_sy = sy + (dsy * step);
    float odxo[3] = {0, 0, 0}, odyo[3] = {0, 0, 0}, odxd[3] = {0, 0, 0}, odyd[3] = {0, 0, 0};
    float wdx[3] = {1, 1, 1}, wdy[3] = {1, 1, 1};
    trace_ray_fw(tries, input_dx_sx, sy, odxo, odxd, wdx, r1, r2, true, ray_id);
    trace_ray_fw(tries, sx, input_dx_sy, odyo, odyd, wdy, r1, r2, true, ray_id);
    const float inv_step = 1.0f / step;  // AtVector::operator/(float) multiplies by the reciprocal [EXTERNAL, as oracle/shims/ai.h]
    for (int k = 0; k < 3; ++k) {
      o[0 + k] = origin[k];
      o[3 + k] = direction[k];
      o[6 + k] = (odxo[k] - origin[k]) * inv_step;     // dOdx
      o[9 + k] = (odyo[k] - origin[k]) * inv_step;     // dOdy
      o[12 + k] = (odxd[k] - direction[k]) * inv_step; // dDdx
      o[15 + k] = (odyd[k] - direction[k]) * inv_step; // dDdy
      o[18 + k] = weight[k] * exposure;
    }
  }

  // lentil.h:573-661 (AiTraceProbe occlusion ray stubbed to "not occluded": no scene in the harness)
  bool trace_ray_bw_po(V3 target, V2 &sensor_position, const int px, const int py, const int total_samples_taken, float lambda_in) {
    int tries = 0;
    bool ray_succes = false;
    double sensor[5] = {0, 0, 0, 0, lambda_in};
    double out[5] = {0, 0, 0, 0, lambda_in};
    V2 aperture{0, 0};
    while (ray_succes == false && tries <= vignetting_retries) {
      V2 unit_disk{0.0, 0.0};
      if (!enable_dof) aperture.x = aperture.y = 0.0;
      else if (enable_dof && bokeh_aperture_blades <= 2) {
        unsigned int seed = tea<8>(px * py + px, total_samples_taken + tries);
        // g++ evaluates the rng(seed) arguments right to left: the LAST argument draws first
        if (bokeh_enable_image) {
          float s2 = rng(seed), s1 = rng(seed), col = rng(seed), row = rng(seed);
          (void)s1; (void)s2;
          image.bokehSample(row, col, unit_disk);
        } else {
          float oy = rng(seed), ox = rng(seed);
          concentric_disk_sample(ox, oy, unit_disk, true);
        }
        aperture.x = unit_disk.x * aperture_radius;
        aperture.y = unit_disk.y * aperture_radius;
      } else if (enable_dof && bokeh_aperture_blades > 2) {
        unsigned int seed = tea<8>(px * py + px, total_samples_taken + tries);
        float b = rng(seed), a = rng(seed);
        lens_sample_triangular_aperture(aperture.x, aperture.y, a, b, aperture_radius, bokeh_aperture_blades);
      }
      sensor[0] = sensor[1] = 0.0;
      ++bw_attempts;
      float transmittance = lens_lt_sample_aperture(target, aperture, sensor, out, lambda_in);
      if (transmittance <= 0) { ++tries; continue; }
      const double ppx = sensor[0] + sensor[2] * lens_back_focal_length;
      const double ppy = sensor[1] + sensor[3] * lens_back_focal_length;
      if (ppx * ppx + ppy * ppy > lens_inner_pupil_radius * lens_inner_pupil_radius) { ++tries; continue; }
      ray_succes = true;
    }
    if (!ray_succes) return false;
    sensor[0] += sensor[2] * -sensor_shift;
    sensor[1] += sensor[3] * -sensor_shift;
    sensor_position.x = sensor[0];
    sensor_position.y = sensor[1];
    return true;
  }

  float get_coc_thinlens(const float csp_z) {  // lentil.h:674-692
    float _focus_distance = focus_distance;
    float _aperture_radius = aperture_radius;
    switch (cameraType) {
      case LB_CAMERA_POLYNOMIAL_OPTICS: _focus_distance /= 10.0; break;
      case LB_CAMERA_THINLENS: _aperture_radius *= 10.0; break;
    }
    const float image_dist_samplepos = (-focal_length * csp_z) / (-focal_length + csp_z);
    const float image_dist_focusdist = (-focal_length * -_focus_distance) / (-focal_length + -_focus_distance);
    return std::abs((_aperture_radius * (image_dist_samplepos - image_dist_focusdist)) / image_dist_samplepos);
  }

  float additional_luminance_soft_trans(const float sample_luminance) {  // lentil.h:1128-1138
    if (sample_luminance > bidir_add_energy_minimum_luminance && sample_luminance < bidir_add_energy_minimum_luminance + bidir_add_energy_transition) {
      float perc = (sample_luminance - bidir_add_energy_minimum_luminance) / bidir_add_energy_transition;
      return bidir_add_energy * perc;
    } else if (sample_luminance > bidir_add_energy_minimum_luminance + bidir_add_energy_transition) {
      return bidir_add_energy;
    }
    return 0.0;
  }

  // lentil.h:823-851
  void add_to_buffer(AOV &aov, const int px, const float aov_value[4], const float fitted_bidir_add_energy, float depth,
                     const float filter_weight, const float rgb_weight[3]) {
    if (aov.filter == LB_FILTER_GAUSSIAN) {
      if (aov.role == LB_AOV_RGBA) filter_weight_buffer[px] += filter_weight;
      // AtRGBA ops: (rgba + f) -> all four channels; * f; * AtRGB -> rgb scaled, alpha kept
      float v[4];
      for (int k = 0; k < 4; ++k) v[k] = (aov_value[k] + fitted_bidir_add_energy) * filter_weight;
      for (int k = 0; k < 3; ++k) v[k] *= rgb_weight[k];
      for (int k = 0; k < 4; ++k) aov.buffer[4 * (size_t)px + k] += v[k];
    } else if (aov.filter == LB_FILTER_CLOSEST) {
      if (aov.role != LB_AOV_LENTIL_DEBUG) {
        if ((std::abs(depth) <= zbuffer[px]) || zbuffer[px] == 0.0) {
          for (int k = 0; k < 4; ++k) aov.buffer[4 * (size_t)px + k] = aov_value[k];
          zbuffer[px] = std::abs(depth);
        }
      } else {
        if ((std::abs(depth) <= zbuffer_debug[px]) || zbuffer_debug[px] == 0.0) {
          if (aov_value[0] != 0.0) {
            for (int k = 0; k < 4; ++k) aov.buffer[4 * (size_t)px + k] = aov_value[k];
            zbuffer_debug[px] = std::abs(depth);
          }
        }
      }
    }
  }

  // cryptomatte_construct_cache, lentil.h:779-811: the depth sub-samples of sample i replace the
  // AiAOVSampleIteratorGetNextDepth walk
  void cryptomatte_construct_cache(std::vector<std::map<float, float>> &crypto_hashmap_cache, const lb_samples *S, size_t i) {
    for (size_t a = 0; a < aovs.size(); ++a) {
      if (aovs[a].filter != LB_FILTER_CRYPTO) continue;
      float iterative_transparency_weight = 1.0f;
      float quota = 1.0;
      float sample_value = 0.0f;
      const int D = S->crypto_depth;
      const float *ids = (S->crypto_ids && S->crypto_ids[a]) ? S->crypto_ids[a] + (size_t)D * i : nullptr;
      const int count = !ids ? 0 : (S->crypto_count ? std::min<int>(S->crypto_count[i], D) : D);
      for (int d = 0; d < count; ++d) {
        const float sub_sample_opacity = S->crypto_opacity ? S->crypto_opacity[(size_t)D * i + d] : 0.0f;
        sample_value = ids[d];
        const float sub_sample_weight = sub_sample_opacity * iterative_transparency_weight;
        iterative_transparency_weight *= (1.0f - sub_sample_opacity);
        quota -= sub_sample_weight;
        crypto_hashmap_cache[a][sample_value] += sub_sample_weight;
      }
      if (quota > 0.0) crypto_hashmap_cache[a][sample_value] += quota;
    }
  }

  // lentil.h:814-819
  void add_to_buffer_cryptomatte(AOV &aov, int px, std::map<float, float> &cryptomatte_cache, const float sample_weight) {
    aov.buffer[4 * (size_t)px] += sample_weight;  // crypto_total_weight
    for (auto const &sample : cryptomatte_cache) aov.crypto_hash_map[px][sample.first] += sample.second * sample_weight;
  }

  // lentil.h:938-955
  void filter_and_add_to_buffer_new(int px, int py, float depth, std::vector<std::map<float, float>> &crypto_cache,
                                    const std::vector<float> &aov_values, float inv_density) {
    const unsigned pixelnumber = xres * py + px;
    float filter_weight = 1.0;
    const float white[3] = {1, 1, 1};
    for (size_t a = 0; a < aovs.size(); ++a) {
      if (aovs[a].filter == LB_FILTER_CRYPTO) add_to_buffer_cryptomatte(aovs[a], pixelnumber, crypto_cache[a], inv_density);
      else add_to_buffer(aovs[a], pixelnumber, &aov_values[4 * a], 0.0, depth, filter_weight * inv_density, white);
    }
  }

  // lentil_filter.cpp:105-301 for ONE sample of the iterator (PolynomialOptics branch)
  void filter_sample(const lb_samples *S, size_t i) {
    const double xres_d = (double)xres, yres_d = (double)yres;
    const double frame_aspect_ratio_without_region = (double)xres_without_region / (double)yres_without_region;
    const int px = S->px[i], py = S->py[i];
    const float inverse_sample_density = S->inv_density;
    bool redistribute = true;
    float sample[4] = {S->rgba[4 * i], S->rgba[4 * i + 1], S->rgba[4 * i + 2], S->rgba[4 * i + 3]};
    // sample_pos_ws (:116): world space when the batch carries a world_to_camera matrix, else camera space (identity)
    float csp[3] = {S->pos_cs[4 * i], S->pos_cs[4 * i + 1], S->pos_cs[4 * i + 2]};
    double depth = S->pos_cs[4 * i + 3];
    // :119-133 AiV3IsSmall is evaluated on the world-space P
    bool small = std::abs(csp[0]) < AI_EPSILON_F && std::abs(csp[1]) < AI_EPSILON_F && std::abs(csp[2]) < AI_EPSILON_F;
    if ((depth == AI_INFINITE_F || small) && enable_skydome) {
      float rd[3] = {0, 0, 0};
      if (S->raydir) { rd[0] = S->raydir[4 * i]; rd[1] = S->raydir[4 * i + 1]; rd[2] = S->raydir[4 * i + 2]; }
      if (rd[0] == 0 && rd[1] == 0 && rd[2] == 0) redistribute = false;
      else { csp[0] = rd[0] * 99999999.0f; csp[1] = rd[1] * 99999999.0f; csp[2] = rd[2] * 99999999.0f; }
    }
    if ((depth == AI_INFINITE_F || small) && !enable_skydome) redistribute = false;
    const uint32_t flags = S->flags ? S->flags[i] : 0u;
    if (flags & LB_SAMPLE_VOLUME) redistribute = false;  // :135-137
    if (S->world_to_camera) {  // :141-142 AiM4PointByMatrixMult(world_to_camera_matrix, sample_pos_ws), float, left to right
      const float *m = S->world_to_camera;
      const float x = csp[0], y = csp[1], z = csp[2];
      csp[0] = x * m[0] + y * m[4] + z * m[8] + m[12];
      csp[1] = x * m[1] + y * m[5] + z * m[9] + m[13];
      csp[2] = x * m[2] + y * m[6] + z * m[10] + m[14];
    }
    switch (unitModel) {  // :143-148
      case LB_UNITS_MM: for (float &v : csp) v *= (float)0.1; break;
      case LB_UNITS_CM: for (float &v : csp) v *= (float)1.0; break;
      case LB_UNITS_DM: for (float &v : csp) v *= (float)10.0; break;
      case LB_UNITS_M: for (float &v : csp) v *= (float)100.0; break;
    }
    if (S->transmission) {  // :152-159
      const float *t = &S->transmission[4 * i];
      bool transmitted = enable_bidir_transmission ? false : (std::max(t[0], std::max(t[1], t[2])) > 0.0);
      if (transmitted) { sample[0] -= t[0]; sample[1] -= t[1]; sample[2] -= t[2]; redistribute = false; }
    }
    const float sample_luminance = (sample[0] + sample[1] + sample[2]) / 3.0;
    if (flags & LB_SAMPLE_IGNORE) redistribute = false;  // :162-164
    std::vector<std::map<float, float>> crypto_cache(aovs.size());  // :167-169
    if (has_crypto) cryptomatte_construct_cache(crypto_cache, S, i);
    float fitted_bidir_add_energy = 0.0;
    if (bidir_add_energy > 0.0) fitted_bidir_add_energy = additional_luminance_soft_trans(sample_luminance);
    float luminance_mult = std::max(0.0, std::pow(std::min(sample_luminance, 20.0f), 0.5) * bidir_sample_mult);  // :177
    float circle_of_confusion = get_coc_thinlens(csp[2]);
    const float coc_squared_pixels = std::pow(circle_of_confusion * yres, 2) * std::pow(luminance_mult, 2) * 0.00001;  // :179
    const float coc_treshold = 0.4;
    if (circle_of_confusion < coc_treshold) redistribute = false;
    int samples = std::ceil(coc_squared_pixels * inverse_sample_density);
    samples = clamp_f(samples, 4, 2000);  // float overload, global.h:8-12
    float inv_samples = 1.0 / static_cast<float>(samples);
    unsigned int total_samples_taken = 0;
    unsigned int max_total_samples = samples * 5;
    // :206-234
    std::vector<float> aov_values(4 * aovs.size(), 0.0f);
    for (size_t a = 0; a < aovs.size(); ++a) {
      if (aovs[a].filter == LB_FILTER_CRYPTO) continue;  // :208
      if (aovs[a].role == LB_AOV_LENTIL_DEBUG) {
        float v = samples * redistribute;
        aov_values[4 * a] = aov_values[4 * a + 1] = aov_values[4 * a + 2] = aov_values[4 * a + 3] = v;  // AtRGBA = float
        continue;
      }
      const float *src = (S->aov_values && S->aov_values[a]) ? &S->aov_values[a][4 * i] : &S->rgba[4 * i];
      for (int k = 0; k < 4; ++k) aov_values[4 * a + k] = src[k];
    }
    ++stats.samples;
    if (cameraType == LB_CAMERA_POLYNOMIAL_OPTICS && std::abs(csp[2]) < (lens_length * 0.1)) redistribute = false;  // :240 (PO case only)
    if (redistribute == false) {
      filter_and_add_to_buffer_new(px, py, depth, crypto_cache, aov_values, inverse_sample_density);
      ++stats.passthrough;
      return;
    }
    ++stats.redistributed;
    if (cameraType == LB_CAMERA_THINLENS) {  // lentil_filter.cpp:303-447
      const V3f P{csp[0], csp[1], csp[2]};
      for (int count = 0; count < samples && total_samples_taken < max_total_samples; ++count, ++total_samples_taken) {
        ++stats.attempts;
        unsigned int seed = tea<8>((px * py + px), total_samples_taken);
        float image_dist_samplepos = (-focal_length * P.z) / (-focal_length + P.z);
        V2 unit_disk{0, 0};
        // argument evaluation order of g++ (right to left), as in trace_ray_bw_po
        if (bokeh_enable_image) { float s2 = rng(seed), s1 = rng(seed), col = rng(seed), row = rng(seed); (void)s1; (void)s2; image.bokehSample(row, col, unit_disk); }
        else if (bokeh_aperture_blades < 2) { float oy = rng(seed), ox = rng(seed); concentricDiskSample(ox, oy, unit_disk, abb_spherical, circle_to_square, bokeh_anamorphic); }
        else { float b = rng(seed), a = rng(seed); lens_sample_triangular_aperture(unit_disk.x, unit_disk.y, a, b, 1.0, bokeh_aperture_blades); }
        unit_disk.x *= bokeh_anamorphic;
        V3f lens{(float)(unit_disk.x * aperture_radius), (float)(unit_disk.y * aperture_radius), 0.0f};
        V3f dir_from_center = normalize3f(P);
        V3f dir_lens_to_P = normalize3f(V3f{P.x - lens.x, P.y - lens.y, P.z - lens.z});
        float abb_coma_multiplied = abb_coma * abb_coma_multipliers(sensor_width, focal_length, dir_from_center, unit_disk);
        dir_lens_to_P = abb_coma_perturb(dir_lens_to_P, dir_from_center, abb_coma_multiplied, true);
        const float lenP = std::sqrt(dot3f(P, P));
        V3f P_perturbed{lenP * dir_lens_to_P.x, lenP * dir_lens_to_P.y, lenP * dir_lens_to_P.z};
        dir_from_center = normalize3f(P_perturbed);
        float samplepos_image_intersection = std::abs(image_dist_samplepos / dir_from_center.z);
        V3f samplepos_image_point{dir_from_center.x * samplepos_image_intersection, dir_from_center.y * samplepos_image_intersection,
                                  dir_from_center.z * samplepos_image_intersection};
        V3f dir_img = normalize3f(V3f{samplepos_image_point.x - lens.x, samplepos_image_point.y - lens.y, samplepos_image_point.z - lens.z});
        float focusdist_intersection_unperturbed = std::abs(get_image_dist_focusdist_thinlens() / dir_img.z);
        V3f fu{lens.x + dir_img.x * focusdist_intersection_unperturbed, lens.y + dir_img.y * focusdist_intersection_unperturbed,
               lens.z + dir_img.z * focusdist_intersection_unperturbed};
        const float spu_x = fu.x / fu.z, spu_y = fu.y / fu.z;
        const float distance_to_center_unperturbed = std::sqrt((0.0f - spu_x) * (0.0f - spu_x) + (0.0f - spu_y) * (0.0f - spu_y));
        // (occlusion probe: no scene, never occluded)
        if (optical_vignetting_distance > 0.0) {
          dir_lens_to_P = normalize3f(V3f{P_perturbed.x - lens.x, P_perturbed.y - lens.y, P_perturbed.z - lens.z});
          if (!empericalOpticalVignettingSquare(lens, dir_lens_to_P, aperture_radius, optical_vignetting_radius, optical_vignetting_distance,
                                                lerp_squircle_mapping(circle_to_square))) {
            --count;
            continue;
          }
        }
        float focusdist_intersection = std::abs(get_image_dist_focusdist_thinlens() / dir_img.z);
        float rgb_weight[3] = {1, 1, 1};
        if (abb_chromatic > 0.0) {
          const float abb_chromatic_lateral = 5.0;
          // the reference draws the channel from the process-global xor128 and keeps the counter-RNG variant in a
          // comment (lentil_filter.cpp:395-396); the counter variant is used here and in the product (stated deviation)
          const int channel = static_cast<int>(rng(seed) * 3) - 1;
          rgb_weight[0] = channel == -1 ? 3 : 0; rgb_weight[1] = channel == 0 ? 3 : 0; rgb_weight[2] = channel == 1 ? 3 : 0;
          float direction_shift = abb_chromatic_type == 0 ? std::abs(channel) : channel;
          focusdist_intersection = std::abs(get_image_dist_focusdist_thinlens_abberated(direction_shift * abb_chromatic * abb_chromatic_lateral * distance_to_center_unperturbed) / dir_img.z);
        }
        V3f fp{lens.x + dir_img.x * focusdist_intersection, lens.y + dir_img.y * focusdist_intersection, lens.z + dir_img.z * focusdist_intersection};
        float sp_x = fp.x / fp.z, sp_y = fp.y / fp.z;
        {  // sensor_position /= (sensor_width*0.5)/-focal_length : AtVector2::operator/=(float), reciprocal multiply [EXTERNAL]
          const float div = (sensor_width * 0.5) / -focal_length;
          const float c = 1.0f / div;
          sp_x *= c; sp_y *= c;
        }
        if (abb_distortion > 0.0) inverseBarrelDistortion(sp_x, sp_y, abb_distortion);
        const double s0 = sp_x, s1 = sp_y * frame_aspect_ratio_without_region;
        const float pixel_x = (((s0 + 1.0) / 2.0) * xres_without_region) - region_min_x;
        const float pixel_y = (((-s1 + 1.0) / 2.0) * yres_without_region) - region_min_y;
        if ((pixel_x >= xres_d) || (pixel_x < 0) || (pixel_y >= yres_d) || (pixel_y < 0)) { --count; continue; }
        unsigned pixelnumber = (unsigned)((int)std::floor(pixel_x) + ((int)std::floor(pixel_y) * xres));
        float filter_weight = 1.0;
        for (size_t a = 0; a < aovs.size(); ++a) {
          if (aovs[a].filter == LB_FILTER_CRYPTO) add_to_buffer_cryptomatte(aovs[a], pixelnumber, crypto_cache[a], inverse_sample_density * inv_samples);
          else add_to_buffer(aovs[a], pixelnumber, &aov_values[4 * a], fitted_bidir_add_energy, depth, filter_weight * inverse_sample_density * inv_samples, rgb_weight);
        }
        ++stats.splats;
      }
      return;
    }
    for (int count = 0; count < samples && total_samples_taken < max_total_samples; ++count, ++total_samples_taken) {
      V2 sensor_position{0, 0};
      float lambda_per_sample = 0.55;
      for (int channel = -1; channel <= 1; channel++) {
        float rgb_weight[3] = {1, 1, 1};
        if (abb_chromatic > 0.0) {
          if (channel == -1) { rgb_weight[0] = 3; rgb_weight[1] = 0; rgb_weight[2] = 0; lambda_per_sample = linear_interpolate(1.0 - abb_chromatic, 0.35, 0.55); }
          else if (channel == 0) { rgb_weight[0] = 0; rgb_weight[1] = 3; rgb_weight[2] = 0; lambda_per_sample = 0.55; }
          else if (channel == 1) { rgb_weight[0] = 0; rgb_weight[1] = 0; rgb_weight[2] = 3; lambda_per_sample = linear_interpolate(abb_chromatic, 0.55, 0.85); }
        } else if (abb_chromatic == 0.0 && channel > -1) continue;
        ++stats.attempts;
        // -camera_space_sample_position_eigen*10.0 : Eigen double vector built from the float position (:251,271)
        V3 target{-(double)csp[0] * 10.0, -(double)csp[1] * 10.0, -(double)csp[2] * 10.0};
        if (!trace_ray_bw_po(target, sensor_position, px, py, total_samples_taken, lambda_per_sample)) { --count; continue; }
        const V2 s{sensor_position.x / (sensor_width * 0.5), sensor_position.y / (sensor_width * 0.5) * frame_aspect_ratio_without_region};
        const V2 pixel{(((s.x + 1.0) / 2.0) * xres_without_region) - region_min_x, (((-s.y + 1.0) / 2.0) * yres_without_region) - region_min_y};
        if ((pixel.x >= xres_d) || (pixel.x < 0) || (pixel.y >= yres_d) || (pixel.y < 0) || (pixel.x != pixel.x) || (pixel.y != pixel.y)) { --count; continue; }
        unsigned pixelnumber = (unsigned)((int)std::floor(pixel.x) + ((int)std::floor(pixel.y) * xres));
        float filter_weight = 1.0;
        for (size_t a = 0; a < aovs.size(); ++a) {
          if (aovs[a].filter == LB_FILTER_CRYPTO) add_to_buffer_cryptomatte(aovs[a], pixelnumber, crypto_cache[a], inverse_sample_density * inv_samples);
          else add_to_buffer(aovs[a], pixelnumber, &aov_values[4 * a], fitted_bidir_add_energy, depth, filter_weight * inverse_sample_density * inv_samples, rgb_weight);
        }
        ++stats.splats;
      }
    }
  }

  // ---- setup, lentil.h:1316-1670 --------------------------------------------------------------
  double camera_get_y0_intersection_distance(double shift) {  // lentil.h:1361-1386
    double sensor[5] = {0, 0, 0, 0, lambda}, aperture[5] = {0, 0, 0, 0, 0}, out[5] = {0, 0, 0, 0, 0};
    aperture[1] = lens_aperture_housing_radius * 0.25;
    lens_pt_sample_aperture(sensor, aperture, shift);
    sensor[0] += sensor[2] * shift;
    sensor[1] += sensor[3] * shift;
    lens_evaluate(sensor, out);
    V3 pos{0, 0, 0}, omega{0, 0, 0};
    outer_to_cs(V2{out[0], out[1]}, V2{out[2], out[3]}, pos, omega);
    return line_plane_intersection(pos, omega).z;
  }
  bool trace_ray_focus_check(double shift, double &test_focus_distance) {  // lentil.h:1316-1357
    double sensor[5] = {0, 0, 0, 0, lambda}, aperture[5] = {0, 0, 0, 0, 0}, out[5] = {0, 0, 0, 0, 0};
    aperture[1] = lens_aperture_housing_radius * 0.25;
    lens_pt_sample_aperture(sensor, aperture, shift);
    sensor[0] += sensor[2] * shift;
    sensor[1] += sensor[3] * shift;
    double transmittance = lens_evaluate(sensor, out);
    if (transmittance <= 0.0) return false;
    if (out[0] * out[0] + out[1] * out[1] > lens_outer_pupil_radius * lens_outer_pupil_radius) return false;
    const double px = sensor[0] + sensor[2] * lens_back_focal_length;
    const double py = sensor[1] + sensor[3] * lens_back_focal_length;
    if (px * px + py * py > lens_inner_pupil_radius * lens_inner_pupil_radius) return false;
    V3 pos{0, 0, 0}, omega{0, 0, 0};
    outer_to_cs(V2{out[0], out[1]}, V2{out[2], out[3]}, pos, omega);
    test_focus_distance = line_plane_intersection(pos, omega).z;
    return true;
  }
  void trace_backwards_for_fstop(const double fstop_target, double &calculated_fstop, double &calculated_aperture_radius) {  // lentil.h:1390-1441
    const int maxrays = 1000;
    double best_valid_fstop = 0.0, best_valid_aperture_radius = 0.0;
    for (int i = 1; i < maxrays; i++) {
      const double parallel_ray_height = (static_cast<double>(i) / static_cast<double>(maxrays)) * lens_outer_pupil_radius;
      const V3 target{0, parallel_ray_height, AI_BIG_F};
      double sensor[5] = {0, 0, 0, 0, lambda}, out[5] = {0, 0, 0, 0, 0};
      V2 aperture{0.01, parallel_ray_height};
      if (lens_lt_sample_aperture(target, aperture, sensor, out, lambda) <= 0.0) continue;
      const double px = sensor[0] + (sensor[2] * lens_back_focal_length);
      const double py = sensor[1] + (sensor[3] * lens_back_focal_length);
      if (px * px + py * py > lens_inner_pupil_radius * lens_inner_pupil_radius) continue;
      V3 out_cs_pos{0, 0, 0}, out_cs_dir{0, 0, 0};
      V2 outpos{out[0], out[1]}, outdir{out[2], out[3]};
      const double Ri = lens_inner_pupil_curvature_radius;
      if (lens_inner_pupil_geometry == 1) cylinderToCs(outpos, outdir, out_cs_pos, out_cs_dir, -Ri + lens_back_focal_length, Ri, true);
      else if (lens_inner_pupil_geometry == 2) cylinderToCs(outpos, outdir, out_cs_pos, out_cs_dir, -Ri + lens_back_focal_length, Ri, false);
      else sphereToCs(outpos, outdir, out_cs_pos, out_cs_dir, -Ri + lens_back_focal_length, Ri);
      const double theta = std::atan(out_cs_pos.y / out_cs_pos.z);
      const double fstop = 1.0 / (std::sin(theta) * 2.0);
      if (fstop < fstop_target) {
        calculated_fstop = best_valid_fstop;
        calculated_aperture_radius = best_valid_aperture_radius;
        return;
      } else {
        best_valid_fstop = fstop;
        best_valid_aperture_radius = parallel_ray_height;
      }
    }
    calculated_fstop = best_valid_fstop;
    calculated_aperture_radius = best_valid_aperture_radius;
  }
  double logarithmic_focus_search(const double focal_distance) {  // lentil.h:1445-1460
    double closest_distance = 999999999.0;
    double best_sensor_shift = 0.0;
    for (double sensorshift : logarithmic_values()) {
      double intersection_distance = camera_get_y0_intersection_distance(sensorshift);
      double new_distance = focal_distance - intersection_distance;
      if (new_distance < closest_distance && new_distance > 0.0) {
        closest_distance = new_distance;
        best_sensor_shift = sensorshift;
      }
    }
    return best_sensor_shift;
  }

  // get_lentil_camera_params (lentil.h:1189-1243) + camera_model_specific_setup (:1568-1670)
  int setup(const lb_camera_params *p, const lb_bokeh_image *img) {
    cameraType = p->camera_type;
    unitModel = p->units;
    sensor_width = p->sensor_width;
    enable_dof = p->enable_dof != 0;
    input_fstop = clamp_min_f(p->fstop, 0.01);
    focus_distance = p->focus_dist;
    bokeh_aperture_blades = p->aperture_blades_lentil;
    exposure = p->exp;
    lensModel = p->lens_model;
    lambda = p->wavelength * 0.001;  // float * double literal
    extra_sensor_shift = p->extra_sensor_shift;
    focal_length = clamp_min_f(p->focal_length_lentil, 0.01);
    abb_chromatic = p->abb_chromatic;
    // thin-lens parameters, lentil.h:1217-1229
    optical_vignetting_distance = p->optical_vignetting;
    optical_vignetting_radius = 1.0;
    abb_spherical = clamp_f(p->abb_spherical, 0.001, 0.999);
    abb_distortion = p->abb_distortion;
    abb_coma = p->abb_coma;
    abb_chromatic_type = p->abb_chromatic_type;
    circle_to_square = clamp_f(p->bokeh_circle_to_square, 0.01, 0.99);
    bokeh_anamorphic = clamp_f(1.0 - p->bokeh_anamorphic, 0, 1.0);
    bokeh_enable_image = p->bokeh_enable_image != 0;
    bidir_sample_mult = p->bidir_sample_mult;
    bidir_add_energy_minimum_luminance = p->bidir_add_energy_minimum_luminance;
    bidir_add_energy = p->bidir_add_energy;
    bidir_add_energy_transition = p->bidir_add_energy_transition;
    vignetting_retries = p->vignetting_retries;
    enable_bidir_transmission = p->enable_bidir_transmission != 0;
    enable_skydome = p->enable_skydome != 0;
    switch (cameraType) {
      case LB_CAMERA_POLYNOMIAL_OPTICS: {
        focus_distance *= 10.0;
        if (!load_lens(lensModel)) return LB_ERR_LENS;
        if (input_fstop == 0.0) aperture_radius = lens_aperture_radius_at_fstop;
        else {
          double calculated_fstop = 0.0, calculated_aperture_radius = 0.0;
          trace_backwards_for_fstop(input_fstop, calculated_fstop, calculated_aperture_radius);
          aperture_radius = std::min(lens_aperture_radius_at_fstop, calculated_aperture_radius);
        }
        double best_sensor_shift = logarithmic_focus_search(focus_distance);
        sensor_shift = best_sensor_shift + extra_sensor_shift;
        double test_focus_distance = 0.0;
        focus_check_ok = trace_ray_focus_check(sensor_shift, test_focus_distance);
        focus_check_distance = test_focus_distance;
        tan_fov = std::tan(lens_field_of_view / 2.0);
      } break;
      case LB_CAMERA_THINLENS: {
        if (!load_lens(lensModel)) return LB_ERR_LENS;
        fov = 2.0 * std::atan(sensor_width / (2.0 * focal_length));
        tan_fov = std::tan(fov / 2.0);
        aperture_radius = (focal_length / (2.0 * input_fstop)) / 10.0;
      } break;
    }
    image = ImageData();
    if (bokeh_enable_image && !image.read(img)) return LB_ERR_IMAGE;  // lentil.h:224-228
    fw_newton_its = fw_traces = bw_newton_its = bw_attempts = 0;
    return LB_OK;
  }
};

// ===============================================================================================
extern "C" {

int orc_camera_create(const lb_camera_params *params, const lb_bokeh_image *bokeh, orc_camera **out) {
  if (!params || !out) return LB_ERR_INVALID;
  orc_camera *c = new orc_camera();
  int rc = c->setup(params, bokeh);
  if (rc != LB_OK) { delete c; return rc; }
  *out = c;
  return LB_OK;
}
void orc_camera_destroy(orc_camera *c) { delete c; }

int orc_camera_get_state(const orc_camera *c, lb_camera_state *s) {
  if (!c || !s) return LB_ERR_INVALID;
  memset(s, 0, sizeof(*s));
  s->aperture_radius = c->aperture_radius;
  s->sensor_shift = c->sensor_shift;
  s->tan_fov = c->tan_fov;
  s->focus_distance = c->focus_distance;
  s->lambda = c->lambda;
  s->lens_outer_pupil_radius = c->lens_outer_pupil_radius;
  s->lens_inner_pupil_radius = c->lens_inner_pupil_radius;
  s->lens_length = c->lens_length;
  s->lens_back_focal_length = c->lens_back_focal_length;
  s->lens_effective_focal_length = c->lens_effective_focal_length;
  s->lens_aperture_pos = c->lens_aperture_pos;
  s->lens_aperture_housing_radius = c->lens_aperture_housing_radius;
  s->lens_inner_pupil_curvature_radius = c->lens_inner_pupil_curvature_radius;
  s->lens_outer_pupil_curvature_radius = c->lens_outer_pupil_curvature_radius;
  s->lens_field_of_view = c->lens_field_of_view;
  s->lens_fstop = c->lens_fstop;
  s->lens_aperture_radius_at_fstop = c->lens_aperture_radius_at_fstop;
  s->outer_pupil_geometry = c->lens_outer_pupil_geometry;
  s->inner_pupil_geometry = c->lens_inner_pupil_geometry;
  s->focus_check_ok = c->focus_check_ok;
  s->focus_check_distance = c->focus_check_distance;
  return LB_OK;
}
int orc_camera_set_pupil_geometry(orc_camera *c, int outer, int inner) {
  if (!c) return LB_ERR_INVALID;
  c->lens_outer_pupil_geometry = outer;
  c->lens_inner_pupil_geometry = inner;
  return LB_OK;
}
int orc_camera_set_state(orc_camera *c, double aperture_radius, double sensor_shift) {
  if (!c) return LB_ERR_INVALID;
  c->aperture_radius = aperture_radius;
  c->sensor_shift = sensor_shift;
  return LB_OK;
}

// camera_create_ray over a batch; HOST pointers; nthreads > 1 splits the batch statically.
// seed word of a 64-bit global ray index, as the product mixes it (csrc/lens_device.cuh ray_seed_word; deviation of the retry RNG)
static inline uint32_t ray_seed_word(uint64_t id) { return (uint32_t)id ^ ((uint32_t)(id >> 32) * 0x9E3779B9u); }
int orc_camera_create_rays(orc_camera *cam, size_t n, uint64_t ray_id_base, const lb_ray_in *in, const lb_ray_out *out, int nthreads) {
  if (!cam || !in || !out) return LB_ERR_INVALID;
  auto work = [&](orc_camera *c, size_t lo, size_t hi) {
    for (size_t i = lo; i < hi; ++i) {
      float o[21];
      int tries = 0;
      c->camera_create_ray(in->sx[i], in->sy[i], in->dsx[i], in->dsy[i], in->lensx[i], in->lensy[i], ray_seed_word(ray_id_base + i), o, &tries);
      float *dst[7] = {out->origin, out->dir, out->dOdx, out->dOdy, out->dDdx, out->dDdy, out->weight};
      for (int v = 0; v < 7; ++v)
        if (dst[v]) { dst[v][i] = o[3 * v]; dst[v][n + i] = o[3 * v + 1]; dst[v][2 * n + i] = o[3 * v + 2]; }
      if (out->tries) out->tries[i] = tries;
    }
  };
  if (nthreads <= 1) { work(cam, 0, n); return LB_OK; }
  std::vector<orc_camera> clones(nthreads, *cam);  // private counters per thread
  for (auto &c : clones) c.fw_newton_its = c.fw_traces = 0;
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t) th.emplace_back(work, &clones[t], n * t / nthreads, n * (t + 1) / nthreads);
  for (auto &t : th) t.join();
  for (auto &c : clones) { cam->fw_newton_its += c.fw_newton_its; cam->fw_traces += c.fw_traces; }
  return LB_OK;
}

int orc_filter_begin(orc_camera *c, const lb_frame_desc *f, int n_aov, const lb_aov_desc *aovs) {  // lentil.h:1056-1117
  if (!c || !f || n_aov < 0) return LB_ERR_INVALID;
  c->xres = f->xres; c->yres = f->yres;
  c->xres_without_region = f->xres_without_region; c->yres_without_region = f->yres_without_region;
  c->region_min_x = f->region_min_x; c->region_min_y = f->region_min_y;
  size_t npx = (size_t)c->xres * c->yres;
  c->zbuffer.assign(npx, 0.0f);
  c->zbuffer_debug.assign(npx, 0.0f);
  c->filter_weight_buffer.assign(npx, 0.0f);
  c->aovs.clear();
  c->has_crypto = false;
  for (int a = 0; a < n_aov; ++a) {
    AOV v;
    v.name = aovs[a].name; v.filter = aovs[a].filter; v.role = aovs[a].role;
    v.buffer.assign(4 * npx, 0.0f);
    if (v.filter == LB_FILTER_CRYPTO) { v.crypto_hash_map.assign(npx, {}); c->has_crypto = true; }  // aov_data.h:145-150
    c->aovs.push_back(std::move(v));
  }
  c->stats = lb_filter_stats{};
  return LB_OK;
}

// filter_pixel over a batch (HOST pointers).  nthreads > 1: static sample partition with PRIVATE
// framebuffers per thread summed at the end (avoids the reference's unsynchronised += race,
// lentil.h:828-829); only valid for gaussian AOVs.
int orc_filter_accumulate(orc_camera *cam, const lb_samples *S, int nthreads) {
  if (!cam || !S) return LB_ERR_INVALID;
  if (nthreads <= 1) {
    for (size_t i = 0; i < S->n; ++i) cam->filter_sample(S, i);
    return LB_OK;
  }
  std::vector<orc_camera> clones(nthreads, *cam);
  for (auto &c : clones) {
    std::fill(c.filter_weight_buffer.begin(), c.filter_weight_buffer.end(), 0.0f);
    for (auto &a : c.aovs) { std::fill(a.buffer.begin(), a.buffer.end(), 0.0f); for (auto &m : a.crypto_hash_map) m.clear(); }
    c.stats = lb_filter_stats{};
    c.bw_newton_its = c.bw_attempts = 0;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() { for (size_t i = S->n * t / nthreads; i < S->n * (t + 1) / nthreads; ++i) clones[t].filter_sample(S, i); });
  for (auto &t : th) t.join();
  for (auto &c : clones) {
    for (size_t i = 0; i < cam->filter_weight_buffer.size(); ++i) cam->filter_weight_buffer[i] += c.filter_weight_buffer[i];
    for (size_t a = 0; a < cam->aovs.size(); ++a)
      for (size_t i = 0; i < cam->aovs[a].buffer.size(); ++i) cam->aovs[a].buffer[i] += c.aovs[a].buffer[i];
    for (size_t a = 0; a < cam->aovs.size(); ++a)
      for (size_t i = 0; i < cam->aovs[a].crypto_hash_map.size(); ++i)
        for (auto const &kv : c.aovs[a].crypto_hash_map[i]) cam->aovs[a].crypto_hash_map[i][kv.first] += kv.second;
    cam->stats.samples += c.stats.samples; cam->stats.redistributed += c.stats.redistributed; cam->stats.splats += c.stats.splats;
    cam->stats.attempts += c.stats.attempts; cam->stats.passthrough += c.stats.passthrough;
    cam->bw_newton_its += c.bw_newton_its; cam->bw_attempts += c.bw_attempts;
  }
  return LB_OK;
}

int orc_filter_get_stats(orc_camera *c, lb_filter_stats *out) { if (!c || !out) return LB_ERR_INVALID; *out = c->stats; return LB_OK; }

// driver_process_bucket, lentil_imager.cpp:112-189
int orc_imager_resolve(orc_camera *c, int aov, int x0, int y0, int w, int h, float *rgba_out) {
  if (!c || aov < 0 || aov >= (int)c->aovs.size() || !rgba_out) return LB_ERR_INVALID;
  const AOV &A = c->aovs[aov];
  for (int j = 0; j < h; ++j)
    for (int i = 0; i < w; ++i) {
      int y = j + y0, x = i + x0;
      int in_idx = j * w + i;
      int linear_pixel = (x - c->region_min_x) + ((y - c->region_min_y) * (int)c->xres);
      if (A.name.find("crypto") != std::string::npos) {  // :122-161
        int rank = 0;
        if (A.name == "crypto_material01" || A.name == "crypto_asset01" || A.name == "crypto_object01") rank = 2;
        else if (A.name == "crypto_material02" || A.name == "crypto_asset02" || A.name == "crypto_object02") rank = 4;
        if (A.crypto_hash_map.empty()) return LB_ERR_INVALID;
        const std::map<float, float> &m = A.crypto_hash_map[linear_pixel];
        if ((int)m.size() <= rank) break;  // :132-134 leaves the rest of this bucket row untouched
        std::vector<std::pair<float, float>> all_vals(m.begin(), m.end());
        std::sort(all_vals.begin(), all_vals.end(), [](const std::pair<float, float> x, const std::pair<float, float> y) { return x.second > y.second; });
        float out[4] = {0, 0, 0, 0};
        int iter = 0;
        const float total = A.buffer[4 * (size_t)linear_pixel];
        for (auto const &v : all_vals) {
          if (iter == rank) { out[0] = v.first; out[1] = v.second / total; }
          else if (iter == rank + 1) { out[2] = v.first; out[3] = v.second / total; }
          iter++;
        }
        for (int k = 0; k < 4; ++k) rgba_out[4 * (size_t)in_idx + k] = out[k];
        continue;
      }
      float image[4] = {A.buffer[4 * (size_t)linear_pixel], A.buffer[4 * (size_t)linear_pixel + 1], A.buffer[4 * (size_t)linear_pixel + 2], A.buffer[4 * (size_t)linear_pixel + 3]};
      if (A.filter == LB_FILTER_GAUSSIAN) {
        if (A.role != LB_AOV_LENTIL_DEBUG) {
          if (c->filter_weight_buffer[linear_pixel] != 0.0) {
            // AtRGBA::operator/=(float): multiplies by the reciprocal
            float inv = 1.0f / c->filter_weight_buffer[linear_pixel];
            for (float &v : image) v *= inv;
          }
        }
      } else if (A.filter == LB_FILTER_CLOSEST) {
        image[3] = 1.0f;
      }
      for (int k = 0; k < 4; ++k) rgba_out[4 * (size_t)in_idx + k] = image[k];
    }
  return LB_OK;
}

// AOVData::crypto_hash_map as fixed-size tables, ascending id per pixel; unused slots id NaN (all bits set), weight 0.
// Returns the largest map size found (may exceed `slots`: such pixels are truncated).
int orc_filter_crypto(orc_camera *c, int aov, int slots, float *ids_out, float *weights_out, float *total_out) {
  if (!c || aov < 0 || aov >= (int)c->aovs.size() || c->aovs[aov].crypto_hash_map.empty()) return -1;
  const AOV &A = c->aovs[aov];
  size_t mx = 0;
  for (size_t p = 0; p < A.crypto_hash_map.size(); ++p) {
    mx = std::max(mx, A.crypto_hash_map[p].size());
    if (total_out) total_out[p] = A.buffer[4 * p];
    int k = 0;
    for (auto const &kv : A.crypto_hash_map[p]) {
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

int orc_filter_buffers(orc_camera *c, int aov, float **buffer, float **weight) {
  if (!c || aov < 0 || aov >= (int)c->aovs.size()) return LB_ERR_INVALID;
  if (buffer) *buffer = c->aovs[aov].buffer.data();
  if (weight) *weight = c->filter_weight_buffer.data();
  return LB_OK;
}

void orc_camera_counters(const orc_camera *c, uint64_t out[4]) {
  out[0] = c->fw_newton_its; out[1] = c->fw_traces; out[2] = c->bw_newton_its; out[3] = c->bw_attempts;
}

// ---- primitive entry points for pinning against oracle/_ref and for KATs ----------------------
unsigned int orc_tea8(unsigned int v0, unsigned int v1) { return tea<8>(v0, v1); }
float orc_rng(unsigned int *state) { return rng(*state); }
void orc_xor128_seq(uint32_t *out, int n) { Xor128 g; for (int i = 0; i < n; ++i) out[i] = g.next(); }
float orc_fast_sin(float x) { return fast_sin(x); }
float orc_fast_cos(float x) { return fast_cos(x); }
double orc_lens_ipow(double x, int e) { return lens_ipow(x, e); }
void orc_concentric_disk_sample(double ox, double oy, int fast_trigo, double out[2]) { V2 d{0, 0}; concentric_disk_sample(ox, oy, d, fast_trigo != 0); out[0] = d.x; out[1] = d.y; }
void orc_sphereToCs(const double inpos[2], const double indir[2], double center, double R, double outpos[3], double outdir[3]) {
  V3 p{0, 0, 0}, d{0, 0, 0}; sphereToCs(V2{inpos[0], inpos[1]}, V2{indir[0], indir[1]}, p, d, center, R);
  outpos[0] = p.x; outpos[1] = p.y; outpos[2] = p.z; outdir[0] = d.x; outdir[1] = d.y; outdir[2] = d.z;
}
void orc_csToSphere(const double inpos[3], const double indir[3], double center, double R, double outpos[2], double outdir[2]) {
  V2 p{0, 0}, d{0, 0}; csToSphere(V3{inpos[0], inpos[1], inpos[2]}, V3{indir[0], indir[1], indir[2]}, p, d, center, R);
  outpos[0] = p.x; outpos[1] = p.y; outdir[0] = d.x; outdir[1] = d.y;
}
void orc_cylinderToCs(const double inpos[2], const double indir[2], double center, double R, int cyl_y, double outpos[3], double outdir[3]) {
  V3 p{0, 0, 0}, d{0, 0, 0}; cylinderToCs(V2{inpos[0], inpos[1]}, V2{indir[0], indir[1]}, p, d, center, R, cyl_y != 0);
  outpos[0] = p.x; outpos[1] = p.y; outpos[2] = p.z; outdir[0] = d.x; outdir[1] = d.y; outdir[2] = d.z;
}
void orc_csToCylinder(const double inpos[3], const double indir[3], double center, double R, int cyl_y, double outpos[2], double outdir[2]) {
  V2 p{0, 0}, d{0, 0}; csToCylinder(V3{inpos[0], inpos[1], inpos[2]}, V3{indir[0], indir[1], indir[2]}, p, d, center, R, cyl_y != 0);
  outpos[0] = p.x; outpos[1] = p.y; outdir[0] = d.x; outdir[1] = d.y;
}
int orc_logarithmic_values(double *out, int cap) { auto v = logarithmic_values(); int n = (int)v.size(); for (int i = 0; i < n && i < cap; ++i) out[i] = v[i]; return n; }
void orc_line_plane_intersection(const double o[3], const double d[3], double out[3]) { V3 r = line_plane_intersection(V3{o[0], o[1], o[2]}, V3{d[0], d[1], d[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; }

// bokeh CDF tables (for upload by tests; the product builds its own)
int orc_bokeh_tables(orc_camera *c, int *size, const float **cdfRow, const int **rowIndices, const float **cdfColumn, const int **columnIndices) {
  if (!c || !c->image.isValid()) return LB_ERR_IMAGE;
  *size = c->image.x; *cdfRow = c->image.cdfRow.data(); *rowIndices = c->image.rowIndices.data();
  *cdfColumn = c->image.cdfColumn.data(); *columnIndices = c->image.columnIndices.data();
  return LB_OK;
}
void orc_bokeh_sample(orc_camera *c, float r_row, float r_col, double out[2]) { V2 l{0, 0}; c->image.bokehSample(r_row, r_col, l); out[0] = l.x; out[1] = l.y; }

// single-call access to the wrappers, for unit tests
double orc_lens_evaluate(orc_camera *c, const double in[5], double out[5]) { return c->lens_evaluate(in, out); }
void orc_lens_pt_sample_aperture(orc_camera *c, double in[5], double out[5], double dist) { c->lens_pt_sample_aperture(in, out, dist); }
double orc_lens_lt_sample_aperture(orc_camera *c, const double scene[3], const double ap[2], double sensor[5], double out[5], double lambda_) {
  return c->lens_lt_sample_aperture(V3{scene[0], scene[1], scene[2]}, V2{ap[0], ap[1]}, sensor, out, lambda_);
}
int orc_trace_ray_bw_po(orc_camera *c, const double target[3], int px, int py, int total_samples_taken, float lambda_in, double sensor_pos[2]) {
  V2 s{0, 0};
  bool ok = c->trace_ray_bw_po(V3{target[0], target[1], target[2]}, s, px, py, total_samples_taken, lambda_in);
  sensor_pos[0] = s.x; sensor_pos[1] = s.y;
  return ok ? 1 : 0;
}
float orc_get_coc_thinlens(orc_camera *c, float z) { return c->get_coc_thinlens(z); }

}  // extern "C"
