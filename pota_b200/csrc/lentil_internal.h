// lentil_internal.h — host-visible launchers of the CUDA kernels (internal to liblentil_b200.so).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "lens_table.h"

namespace lb {

struct RayIO;

// ---- K1 (camera_kernels.cu) -------------------------------------------------------------------
// lens_kernel < 0: table-driven evaluator; otherwise the unrolled kernel of that LensModel.
cudaError_t launch_create_rays(int lens_kernel, const LensTable &lens, const CamConsts<float> &cam, const RayIO &io, size_t n,
                               uint64_t ray_id_base, cudaStream_t stream);
bool has_unrolled_kernel(int lens_model);
cudaError_t launch_reverse_rays(size_t n, const float4 *Po, float2 *Ps, float tan_fov, cudaStream_t stream);

// ---- setup solvers, FP64 (setup_kernels.cu) -----------------------------------------------------
// camera_get_y0_intersection_distance (lentil.h:1361-1386) for n sensor shifts.
// out[i] = {intersection z, transmittance, out_x^2+out_y^2, inner-pupil r^2}
cudaError_t launch_focus_distances(const LensTable &lens, const CamConsts<double> &cam, double aperture_y, const double *shifts,
                                   int n, double4 *out, cudaStream_t stream);
// trace_backwards_for_fstop (lentil.h:1390-1441): for i in [1, n): out[i] = {valid, pos.y, pos.z, 0}
cudaError_t launch_fstop_rays(const LensTable &lens, const CamConsts<double> &cam, int n, double outer_pupil_radius, double4 *out,
                              cudaStream_t stream);

// ---- thin-lens path (thinlens_kernels.cu): trace_ray_fw_thinlens lentil.h:431-569, filter ThinLens branch
// lentil_filter.cpp:303-447.  Scalars of struct Camera the thin-lens code reads, types as in the reference.
struct ThinConsts {
  double sensor_half;       // sensor_width * 0.5
  double focus_distance;
  double aperture_radius;
  float focal_length;
  float abb_spherical, circle_to_square, bokeh_anamorphic;
  float abb_coma, abb_distortion, abb_chromatic;
  int32_t abb_chromatic_type;
  float optical_vignetting_distance, optical_vignetting_radius;
  float squircle;              // lerp_squircle_mapping(circle_to_square), lens.h:544-546 (host libm)
  float coma_max_projection;   // maximal_projection of abb_coma_multipliers, lens.h:567-568
  float image_dist_focusdist;  // get_image_dist_focusdist_thinlens(), lentil.h:664-666
  float unit_scale;            // 10, 1, 0.1, 0.01 (lentil.h:538-560)
};
struct FilterConsts;
struct AovSet;
struct SampleIO;
struct WorkItem;
struct FilterCounters;
cudaError_t launch_create_rays_thinlens(const CamConsts<float> &cam, const ThinConsts &tl, const RayIO &io, size_t n, uint64_t ray_id_base,
                                        cudaStream_t stream);
cudaError_t launch_filter_splat_thinlens(const CamConsts<float> &cam, const ThinConsts &tl, const FilterConsts &fc, const AovSet &aovs,
                                         const SampleIO &s, const WorkItem *work, FilterCounters *counters, uint64_t sample_base, int num_sms,
                                         cudaStream_t stream);

// ---- K2/K3 (filter_kernels.cu) ------------------------------------------------------------------
struct FilterConsts {
  int32_t xres, yres, xres_full, yres_full, region_min_x, region_min_y;
  int32_t n_aov;
  int32_t bidir_sample_mult;
  int32_t camera_type;
  int32_t enable_skydome, enable_bidir_transmission;
  float unit_mult;            // 0.1, 1, 10, 100 (lentil_filter.cpp:143-148)
  float focal_length;         // thin-lens focal length param (get_coc_thinlens)
  float coc_focus_distance;   // (float)focus_distance, /10 for PO (lentil.h:676-681)
  float coc_aperture_radius;  // (float)aperture_radius, *10 for thin lens
  float bidir_add_energy, bidir_add_energy_transition;
  double bidir_add_energy_minimum_luminance;
  float abb_chromatic;
  double lens_length_tenth;   // lens_length * 0.1 (lentil_filter.cpp:240)
  double sensor_half;         // sensor_width * 0.5
  double aspect_full;         // xres_without_region / yres_without_region
};

constexpr int kMaxAov = 16;
struct AovSet {
  float4 *buffer[kMaxAov];         // AOVData::buffer
  const float4 *values[kMaxAov];   // per-sample values of this batch (NULL: use rgba / lentil_debug)
  int32_t filter[kMaxAov];
  int32_t role[kMaxAov];
  float *weight;                   // filter_weight_buffer
  unsigned long long *zkey;        // closest-filter depth|~sample key per pixel
  unsigned long long *zkey_debug;
  uint16_t *debug_samples;         // per-sample `samples * redistribute` of this batch (closest lentil_debug AOV only)
  // cryptomatte AOVs (filter == LB_FILTER_CRYPTO): AOVData::crypto_hash_map as crypto_slots open-addressed
  // slots per pixel; AOVData::crypto_total_weight lives in buffer[a][pixel].x
  uint32_t *crypto_key[kMaxAov];     // [npx][crypto_slots] id bits, kCryptoFree = unused
  float *crypto_wgt[kMaxAov];        // [npx][crypto_slots]
  const float *crypto_ids[kMaxAov];  // this batch: [n][crypto_depth] ids of the depth sub-samples
  float2 *crypto_cache[kMaxAov];     // this batch: [n][max(crypto_depth,1)] merged {id, weight} of each sample, packed,
                                     // kCryptoFree-terminated (cryptomatte_construct_cache, lentil.h:779-811)
  int32_t crypto_slots, crypto_depth;
  int32_t crypto_first;              // the cryptomatte AOV whose plane holds crypto_total_weight for all of them (-1: none)
  int32_t add_zeros;                 // 1: send all-zero gaussian contributions too (LB_ADD_ZEROS=1, A/B timing); default 0
  unsigned int *work_heads;          // this batch's work list: [0] items appended by classify, [1] next work-unit ticket,
                                     // [2] sum and [3] maximum of n_samples over the items (classify), [4] the thin-lens
                                     // splat kernel picked for the batch (k_thinlens_pick); kWorkHeads words
};
constexpr int kWorkHeads = 8;
constexpr uint32_t kCryptoFree = 0xFFFFFFFFu;  // a NaN bit pattern: Cryptomatte hashes are never NaN
constexpr int kCryptoMaxDepth = 8;

struct SampleIO {
  size_t n;
  const int32_t *px, *py;
  const float4 *rgba, *pos_cs, *raydir, *transmission;
  const uint32_t *flags;
  float inv_density;
  const uint8_t *crypto_count;   // [n] or NULL
  const float *crypto_opacity;   // [n][crypto_depth] or NULL
  float w2c[4][3];               // world_to_camera_matrix of the batch (columns 0..2 of the AtMatrix; identity by default)
};

struct WorkItem {  // one redistributed source sample
  uint32_t sample;
  uint32_t n_samples;  // clamp(ceil(coc^2 * ...), 4, 2000)
  float add_energy;    // fitted_bidir_add_energy
  float csp[3];        // camera-space position after unit scaling / skydome substitution (lentil_filter.cpp:121-148)
  // PO splat kernel: the first n_samples attempts are dealt out in chunks to whichever warps are free (filter_kernels.cuh)
  uint32_t chunks_done;  // chunks completed
  uint32_t fails;        // failed attempts among the completed chunks
};

struct FilterCounters {  // device-side mirror of lb_filter_stats
  unsigned long long samples, redistributed, splats, attempts, passthrough;
  unsigned long long newton_its;
  unsigned long long crypto_dropped;
  unsigned long long tile_splats;  // splats accumulated in a shared-memory window before they reached L2 (thin-lens tile kernel)
};

cudaError_t launch_filter_classify(const FilterConsts &fc, const AovSet &aovs, const SampleIO &s, WorkItem *work,
                                   FilterCounters *counters, uint64_t sample_base, cudaStream_t stream);
cudaError_t launch_filter_splat(int lens_kernel, const LensTable &lens, const CamConsts<float> &cam, const FilterConsts &fc,
                                const AovSet &aovs, const SampleIO &s, const WorkItem *work, FilterCounters *counters,
                                uint64_t sample_base, int num_sms, cudaStream_t stream);
cudaError_t launch_closest_gather(const FilterConsts &fc, const AovSet &aovs, const SampleIO &s, uint64_t sample_base,
                                  cudaStream_t stream);
// ranked cryptomatte resolve of one bucket (lentil_imager.cpp:122-161); out is read-modify-write
// brk != NULL: no row ends early; instead brk[pixel] = 1 marks the pixels that hold <= rank ids (they are not written), so the
// caller can end bucket rows itself (out and brk are then indexed like the frame: w == xres, x0 == y0 == 0)
cudaError_t launch_resolve_crypto(const uint32_t *key, const float *wgt, const float4 *total, int slots, int rank, int xres, int x0, int y0,
                                  int w, int h, float4 *out, uint8_t *brk, cudaStream_t stream);
// multi-GPU: fold another rank's tables into this rank's
cudaError_t launch_crypto_merge(uint32_t *key, float *wgt, const uint32_t *other_key, const float *other_wgt, size_t npx, int slots,
                                FilterCounters *counters, cudaStream_t stream);
cudaError_t launch_resolve(const float4 *buffer, const float *weight, int filter, int role, int xres, int x0, int y0, int w, int h,
                           float4 *out, cudaStream_t stream);
// same for the pixel range [first, first + count) of the frame, out[0] = pixel `first`
cudaError_t launch_resolve_linear(const float4 *buffer, const float *weight, int filter, int role, size_t first, size_t count, float4 *out,
                                  cudaStream_t stream);

}  // namespace lb
