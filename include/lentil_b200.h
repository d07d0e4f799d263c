/* lentil_b200.h — C ABI of the B200-native lentil hot paths.
 *
 * Drop-in boundary for the two per-ray hot paths of zpelgrims/pota ("lentil"):
 *   camera_create_ray      /root/reference/src/lentil_camera.cpp:78-125
 *   filter_pixel           /root/reference/src/lentil_filter.cpp:66-301   (PolynomialOptics branch)
 *   driver_process_bucket  /root/reference/src/lentil_imager.cpp:66-193   (resolve)
 * Arnold calls those once per sample / pixel / bucket from its render threads; this ABI is the
 * batch form of the same contracts: plain pointers and sizes, structure-of-arrays, no C++ or
 * torch types.  Unless a function name ends in `_host`, every data pointer is a DEVICE pointer
 * on the camera's device and `stream` is a cudaStream_t (NULL = default stream).
 *
 * Error convention (reference: callbacks return void, failures are in-band, setup errors abort
 * the render — lentil.h:226-227,379,424,651): functions return LB_OK or a negative lb_status;
 * a ray that cannot be traced comes back with weight = 0, exactly as the reference does.
 * There is no CPU fallback: without a CUDA device lb_camera_create fails with LB_ERR_NO_DEVICE.
 */
#ifndef LENTIL_B200_H
#define LENTIL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define LB_API
#else
#define LB_API __attribute__((visibility("default")))
#endif

typedef struct lb_camera lb_camera; /* replaces `struct Camera`, lentil.h:92-1671 */
typedef void *lb_stream;            /* cudaStream_t */

typedef enum lb_status {
  LB_OK = 0,
  LB_ERR_INVALID = -1,   /* bad argument */
  LB_ERR_NO_DEVICE = -2, /* no CUDA device / extension unusable: there is no CPU path */
  LB_ERR_CUDA = -3,      /* CUDA runtime error, see lb_last_error() */
  LB_ERR_LENS = -4,      /* unknown lens_model */
  LB_ERR_STATE = -5,     /* call order (e.g. accumulate before lb_filter_begin) */
  LB_ERR_IMAGE = -6,     /* bokeh image not square / < 3 channels (imagebokeh.h:49-51,97-101) */
  LB_ERR_COMM = -7       /* NCCL unavailable or failed */
} lb_status;

/* enums of lentil.h:60-84 */
enum { LB_UNITS_MM = 0, LB_UNITS_CM = 1, LB_UNITS_DM = 2, LB_UNITS_M = 3 };
enum { LB_CAMERA_THINLENS = 0, LB_CAMERA_POLYNOMIAL_OPTICS = 1 };

/* Node parameters of lentil_camera (lentil_camera.cpp:19-52), read as in lentil.h:1189-1243.
 * lb_camera_params_default() fills the reference's C++ defaults. */
typedef struct lb_camera_params {
  int32_t camera_type;       /* LB_CAMERA_*: both models are traced (ThinLens is the reference's default, lentil_camera.cpp:20) */
  int32_t bidir_sample_mult; /* 5 */
  int32_t units;             /* LB_UNITS_*  ("automatic" must be resolved by the caller, lentil.h:1193-1199) */
  float sensor_width;        /* 36 */
  int32_t enable_dof;        /* 1 */
  float fstop;               /* 0 -> clamp_min 0.01 (lentil.h:1206) */
  float focus_dist;          /* 150, scene units */
  int32_t aperture_blades_lentil; /* 0 */
  float exp;                 /* exposure, 1 */
  int32_t lens_model;        /* LensModel enum value, pota_h_lenses.h:4-47 (index into the lens pack) */
  float wavelength;          /* nm, 550 */
  float extra_sensor_shift;  /* 0 */
  float focal_length_lentil; /* mm, 35 (thin-lens focal length; also used by get_coc_thinlens in PO mode) */
  float optical_vignetting;
  float abb_spherical;
  float abb_distortion;
  float abb_coma;
  float abb_chromatic;
  int32_t abb_chromatic_type;
  float bokeh_circle_to_square;
  float bokeh_anamorphic;
  int32_t bokeh_enable_image; /* 0; needs a lb_bokeh_image */
  int32_t vignetting_retries; /* 15 */
  float bidir_add_energy;
  float bidir_add_energy_minimum_luminance; /* 2 */
  float bidir_add_energy_transition;        /* 1 */
  int32_t enable_bidir_transmission;
  int32_t enable_skydome;
} lb_camera_params;

/* Bokeh kernel image as AiTextureLoad would return it (imagebokeh.h:83-107): host floats,
 * row-major, `channels` interleaved, square. */
typedef struct lb_bokeh_image {
  int32_t width, height, channels;
  const float *pixels; /* HOST pointer */
} lb_bokeh_image;

/* Scalars derived by setup (camera_model_specific_setup, lentil.h:1568-1670) + lens constants
 * (lentil.h:106-120). */
typedef struct lb_camera_state {
  double aperture_radius;
  double sensor_shift;
  double tan_fov;
  double focus_distance; /* after the x10 of lentil.h:1574 */
  double lambda;         /* micrometres */
  double lens_outer_pupil_radius, lens_inner_pupil_radius, lens_length, lens_back_focal_length;
  double lens_effective_focal_length, lens_aperture_pos, lens_aperture_housing_radius;
  double lens_inner_pupil_curvature_radius, lens_outer_pupil_curvature_radius;
  double lens_field_of_view, lens_fstop, lens_aperture_radius_at_fstop;
  int32_t outer_pupil_geometry; /* 0 spherical, 1 cyl-y, 2 cyl-x */
  int32_t inner_pupil_geometry;
  int32_t focus_check_ok; /* trace_ray_focus_check result, lentil.h:1316-1357 */
  double focus_check_distance;
} lb_camera_state;

/* ---- camera_create_ray ------------------------------------------------------------------ */

/* AtCameraInput fields the reference reads (lentil_camera.cpp:82-98), one array each. */
typedef struct lb_ray_in {
  const float *sx, *sy;       /* screen-space sample */
  const float *dsx, *dsy;     /* screen-space derivative steps */
  const float *lensx, *lensy; /* lens sample in [0,1) */
} lb_ray_in;

/* AtCameraOutput fields the reference writes (lentil_camera.cpp:115-124).  Each vector output is
 * three planes of n floats, [3][n] (x plane, y plane, z plane): stores are warp-coalesced and the
 * result carries no padding over PCIe/NVLink.  Any pointer may be NULL to skip that output.
 * `tries` = vignetting retries the main ray used (lentil.h:290), diagnostic only. */
typedef struct lb_ray_out {
  float *origin; /* [3][n] */
  float *dir;    /* [3][n] */
  float *dOdx, *dOdy, *dDdx, *dDdy; /* [3][n] each */
  float *weight; /* [3][n] r,g,b */
  int32_t *tries;
} lb_ray_out;

LB_API void lb_camera_params_default(lb_camera_params *p);
LB_API int lb_lens_count(void);
LB_API const char *lb_lens_name(int lens_model); /* LensModelNames, pota_cpp_lenses.h */
LB_API const char *lb_last_error(void);
/* "lentil_b200 <major>.<minor>.<patch> (sm_100a)".  0.2.0: lb_frame_desc, lb_samples and lb_filter_stats grew at their
 * ends (cryptomatte); 0.3.0: lb_samples grew at its end (world_to_camera); 0.4.0: lb_filter_stats grew at its end
 * (tile_splats).  Callers built against an older minor must be
 * recompiled. */
LB_API const char *lb_version(void);

/* node_initialize + node_update (lentil_camera.cpp:56-68): builds the camera on CUDA device
 * `device`, runs the setup solvers there, uploads lens tables and the bokeh CDF. */
LB_API int lb_camera_create(const lb_camera_params *params, const lb_bokeh_image *bokeh /*nullable*/, int device,
                            lb_camera **out);
LB_API int lb_camera_update(lb_camera *cam, const lb_camera_params *params, const lb_bokeh_image *bokeh);
LB_API void lb_camera_destroy(lb_camera *cam); /* node_finish, lentil_camera.cpp:70-75 */
LB_API int lb_camera_get_state(const lb_camera *cam, lb_camera_state *out);
/* Override the two solver results (tests / callers that cache them). */
LB_API int lb_camera_set_state(lb_camera *cam, double aperture_radius, double sensor_shift);
/* Override lens_outer/inner_pupil_geometry (0 spherical, 1 cyl-y, 2 cyl-x; lentil.h:119-120): the pack's
 * stand-in lenses are all spherical, anamorphic lenses of a user pack select the cylinder transforms. */
LB_API int lb_camera_set_pupil_geometry(lb_camera *cam, int outer_geometry, int inner_geometry);

/* camera_create_ray for n samples.  Retries after a vignetted main ray draw their lens sample
 * from the reference's own counter RNG, seed = tea<8>(ray_id_base + i, try) (global.h:32-57),
 * instead of the reference's process-global xor128 state (lentil.h:313-316) — see DESIGN.md. */
LB_API int lb_camera_create_rays(lb_camera *cam, size_t n, uint64_t ray_id_base, const lb_ray_in *in,
                                 const lb_ray_out *out, lb_stream stream);
/* Same contract with HOST buffers: pipelines H2D, kernel and D2H in chunks on internal streams
 * and returns when the outputs are in host memory. */
LB_API int lb_camera_create_rays_host(lb_camera *cam, size_t n, uint64_t ray_id_base, const lb_ray_in *in,
                                      const lb_ray_out *out);
/* camera_reverse_ray (lentil_camera.cpp:164-172): pinhole approximation, Po [n][4] -> Ps [n][2]. */
LB_API int lb_camera_reverse_rays(lb_camera *cam, size_t n, const float *Po, float *Ps, lb_stream stream);

/* Algorithmic flop counts of the camera's lens (F(P) = sum over terms of
 * (degree + 1)), for roofline bookkeeping. */
typedef struct lb_lens_work {
  int32_t terms_eval, terms_ap, terms_ap_jac, terms_out_jac;
  double F_eval;   /* 5 polynomials of pt_evaluate */
  double F_ap;     /* 4 aperture polynomials + 4 Jacobian polynomials (pt_sample_aperture) */
  double F_apxy;   /* lt: 2 aperture position polynomials */
  double F_apJ;    /* lt: their 4 derivatives wrt (dx,dy) */
  double F_out4;   /* lt: 4 outer-pupil polynomials */
  double F_outJ;   /* lt: 4 derivatives of out[2..3] wrt (x,y) */
  double F_T;      /* transmittance polynomial */
} lb_lens_work;
LB_API int lb_camera_lens_work(const lb_camera *cam, lb_lens_work *out);
/* 1 when the camera's lens runs the per-lens unrolled kernels, 0 for the table-driven kernels. */
LB_API int lb_camera_kernel_kind(const lb_camera *cam);
/* On-box FP32 FMA throughput in TFLOP/s (register-operand FFMA chains): roofline denominator. */
LB_API int lb_bench_fp32_peak(int device, double *tflops_out);
/* On-box throughput of 16-byte vector reductions (red.global.add.v4.f32) at random pixels of a `megabytes` MB
 * plane, in GB/s: roofline denominator of the splat accumulate. */
LB_API int lb_bench_red_peak(int device, int megabytes, double *gbytes_per_s_out);
/* The splat accumulate alone (RGBA + filter weight = 20 B per splat at pseudo-random pixels), in 1e9 splats/s:
 * mode 0 = red.global.add.v4.f32 + red.global.add.f32 into planes of `megabytes` MB (what the splat kernels do,
 * Camera::add_to_buffer lentil.h:823-851); mode 1 = a 96x96-pixel shared-memory tile updated with float atomics and
 * flushed once with vector reductions; mode 2 = the same tile with native 32-bit integer atomics (the floor of any
 * shared-memory-atomic tile).  Evidence for where the accumulate should live on this part. */
LB_API int lb_bench_splat_accum(int device, int mode, int megabytes, double *gsplats_per_s_out);

/* Known-answer hook for the device-side integer / float primitives (global.h:32-57 tea<8> and rng, lens.h:17-37
 * fast_sin / fast_cos): for n seed pairs (v0, v1) computes on the device t = tea<8>(v0, v1), four rng() draws seeded
 * with t (state after each draw and the float), and fast_sin / fast_cos of (first draw - 0.5) * 20.  HOST arrays:
 * tea_out [n], lcg_state_out / lcg_float_out [n][4], trig_out [n][2]. */
LB_API int lb_debug_primitives(int device, size_t n, const uint32_t *v0, const uint32_t *v1, uint32_t *tea_out,
                               uint32_t *lcg_state_out, float *lcg_float_out, float *trig_out);

/* ---- filter / imager ---------------------------------------------------------------------- */

/* AOVData::original_filter (lentil.h:827,832); LB_FILTER_CRYPTO marks an AOVData with is_crypto set
 * (lentil.h:1036-1038): a ranked cryptomatte AOV ("crypto_material00", ...), accumulated as per-pixel
 * {hash id -> weight} tables (aov_data.h:126-128) and resolved by rank (lentil_imager.cpp:122-161). */
enum { LB_FILTER_GAUSSIAN = 0, LB_FILTER_CLOSEST = 1, LB_FILTER_CRYPTO = 2 };
enum { LB_CRYPTO_MAX_DEPTH = 8, LB_CRYPTO_DEFAULT_SLOTS = 16, LB_CRYPTO_MAX_SLOTS = 64 };
enum { LB_AOV_PLAIN = 0, LB_AOV_RGBA = 1, LB_AOV_LENTIL_DEBUG = 2 }; /* atstring_rgba / atstring_lentil_debug */

typedef struct lb_aov_desc {
  char name[64];
  int32_t filter; /* LB_FILTER_* */
  int32_t role;   /* LB_AOV_*: RGBA also feeds filter_weight_buffer; lentil_debug stores the splat count */
} lb_aov_desc;

/* Render region (lentil.h:1070-1080): xres/yres are the region size the buffers are allocated
 * with, *_without_region the full frame. */
typedef struct lb_frame_desc {
  int32_t xres, yres;
  int32_t xres_without_region, yres_without_region;
  int32_t region_min_x, region_min_y;
  int32_t crypto_slots; /* distinct ids kept per pixel and cryptomatte AOV (the reference's std::map is unbounded);
                           0 = LB_CRYPTO_DEFAULT_SLOTS.  Contributions that find every slot taken are counted in
                           lb_filter_stats.crypto_dropped. */
} lb_frame_desc;

/* One batch of AOV samples, as filter_pixel reads them from the AtAOVSampleIterator
 * (lentil_filter.cpp:105-234).  float4 arrays are [n][4]. */
typedef struct lb_samples {
  size_t n;
  const int32_t *px, *py;   /* pixel the sample is filtered for, region-relative (:96-100) */
  const float *rgba;        /* [n][4] RGBA AOV (:115) */
  const float *pos_cs;      /* [n][4] xyz = the P AOV (:116), w = Z depth AOV (:117).  P is WORLD space when
                               world_to_camera is given -- the device then does what :121-142 does: AiV3IsSmall on the
                               world-space P, skydome substitution in world space, AiM4PointByMatrixMult.  With
                               world_to_camera == NULL (identity) P is the camera-space position before unit scaling. */
  const float *raydir;      /* [n][4] lentil_raydir AOV (same space as P), nullable (skydome only, :121-128) */
  const float *transmission;/* [n][4] transmission AOV, nullable (:152-159) */
  const uint32_t *flags;    /* LB_SAMPLE_* bits, nullable */
  const float *const *aov_values; /* [n_aov] device pointers to [n][4] values already widened to RGBA (:206-234);
                                     entry may be NULL for the RGBA role (uses `rgba`) and for lentil_debug */
  float inv_density;        /* inverse_sample_density (:84) */
  /* cryptomatte depth sub-samples, what cryptomatte_construct_cache (lentil.h:779-811) walks with
   * AiAOVSampleIteratorGetNextDepth; all NULL / 0 when the frame has no LB_FILTER_CRYPTO AOV */
  int32_t crypto_depth;           /* D <= LB_CRYPTO_MAX_DEPTH: sub-sample slots stored per sample */
  const uint8_t *crypto_count;    /* [n] valid sub-samples of each sample (<= D); NULL = D for every sample */
  const float *crypto_opacity;    /* [n][D] AiColorToGrey(opacity AOV) of each sub-sample (:790) */
  const float *const *crypto_ids; /* [n_aov] device pointers to [n][D] hash ids of that AOV (:791); NULL entries for
                                     the other AOVs */
  /* AiWorldToCameraMatrix of the batch (lentil_filter.cpp:139-142): HOST pointer (also in the device-pointer call) to an
   * AtMatrix, float[4][4] row-major, row-vector convention p' = p * M as AiM4PointByMatrixMult; NULL = identity. */
  const float *world_to_camera;
} lb_samples;
enum { LB_SAMPLE_VOLUME = 1u, LB_SAMPLE_IGNORE = 2u }; /* volume_in_sample (:136), lentil_bidir_ignore > 0 (:162) */

typedef struct lb_filter_stats {
  uint64_t samples;       /* source samples consumed */
  uint64_t redistributed; /* of which were reverse traced */
  uint64_t splats;        /* successful (source, lens sample) pairs that reached add_to_buffer */
  uint64_t attempts;      /* reverse-trace attempts (total_samples_taken summed) */
  uint64_t passthrough;   /* filter_and_add_to_buffer_new adds (:243-246) */
  uint64_t crypto_dropped;/* cryptomatte contributions lost because a pixel already held crypto_slots other ids */
  uint64_t tile_splats;   /* of `splats`: accumulated in a shared-memory window first, reaching L2 merged (thin-lens tile kernel) */
} lb_filter_stats;

/* setup_filter (lentil.h:1056-1117): allocates zeroed device framebuffers. */
LB_API int lb_filter_begin(lb_camera *cam, const lb_frame_desc *frame, int n_aov, const lb_aov_desc *aovs);
/* filter_pixel, RGBA branch, for a batch of samples: classify, reverse trace, splat. */
/* Calls on one camera execute in call order even when they are issued on different streams (they share the work list and
 * the planes): each waits on the device for the previous accumulate / reduce / resolve of that camera. */
LB_API int lb_filter_accumulate(lb_camera *cam, const lb_samples *samples, lb_stream stream);
/* Same with every pointer in lb_samples (and aov_values[i]) a HOST pointer (pinned or pageable): the samples travel in
 * chunks through a ring of six device staging blocks, the copies running ahead of the kernels of the previous chunks. */
LB_API int lb_filter_accumulate_host(lb_camera *cam, const lb_samples *samples);
LB_API int lb_filter_get_stats(lb_camera *cam, lb_filter_stats *out); /* synchronises */
/* Diagnostic: lt_sample_aperture Newton iterations executed since lb_filter_begin (synchronises). */
LB_API int lb_filter_newton_iterations(lb_camera *cam, uint64_t *out);
/* driver_process_bucket (lentil_imager.cpp:112-189) for one bucket of one AOV -> rgba_out [h][w][4].
 * Cryptomatte AOVs: rank 0 / 2 / 4 from the AOV name as there (:124-126); rgba_out = {id, coverage, id, coverage}
 * of ranks r and r+1.  As in the reference (:132-134, a `break` out of the row loop), the first pixel of a
 * bucket row that holds <= rank ids ends that row: it and the pixels after it are left as the caller
 * supplied them, so rgba_out is read-modify-write for these AOVs. */
LB_API int lb_imager_resolve(lb_camera *cam, int aov, int x0, int y0, int w, int h, float *rgba_out, lb_stream stream);
/* HOST bucket.  Meant to be called per bucket and output like driver_process_bucket: the first call after the frame
 * changed resolves the whole region of that AOV into pinned host memory (one kernel, one copy), every call copies its
 * bucket out of it -- no device work, allocation or synchronisation per bucket. */
LB_API int lb_imager_resolve_host(lb_camera *cam, int aov, int x0, int y0, int w, int h, float *rgba_out);
/* Raw accumulators (device pointers owned by the camera): AOVData::buffer [yres][xres][4],
 * filter_weight_buffer [yres][xres].  For a cryptomatte AOV the plane's x component is
 * AOVData::crypto_total_weight (aov_data.h:128), y/z/w are unused. */
LB_API int lb_filter_buffers(lb_camera *cam, int aov, float **buffer, float **filter_weight_buffer);
/* Copies of the raw accumulators in HOST memory (either pointer may be NULL); synchronises. */
LB_API int lb_filter_buffers_host(lb_camera *cam, int aov, float *buffer_out, float *filter_weight_buffer_out);
/* AOVData::crypto_hash_map of a cryptomatte AOV as fixed-size tables in HOST memory: ids_out and
 * weights_out are [yres][xres][crypto_slots]; unused slots have id NaN (bits 0xFFFFFFFF) and weight 0.
 * Returns the slot count through *slots_out (any pointer may be NULL); synchronises. */
LB_API int lb_filter_crypto_host(lb_camera *cam, int aov, float *ids_out, float *weights_out, int *slots_out);

/* ---- multi-GPU (new: the reference is single-process) -------------------------------------- */

/* One process per GPU.  Each rank accumulates its share of the source samples into full-frame
 * partial framebuffers; lb_filter_reduce combines them over NVLink: closest AOVs by a 64-bit
 * min-reduce of the depth keys (losers zero their pixel), then ONE float32 sum-reduce over the
 * contiguous block holding every AOV plane and the weight plane.  root < 0: all ranks get the
 * result (ncclAllReduce); root >= 0: only that rank (ncclReduce).  The communicator is created from
 * a 128-byte NCCL unique id produced by one rank and distributed by the caller.  Ranks must give
 * their samples disjoint global indices (lb_filter_set_sample_base) when closest AOVs are used. */
LB_API int lb_comm_unique_id(uint8_t id_out[128]);
LB_API int lb_comm_init(lb_camera *cam, int world_size, int rank, const uint8_t id[128]);
LB_API int lb_filter_set_sample_base(lb_camera *cam, uint64_t first_global_sample_index);
LB_API int lb_filter_reduce(lb_camera *cam, int root, lb_stream stream);
/* The same combine in its scalable form: after the closest-key and cryptomatte steps, one in-place
 * ncclReduceScatter per plane (all planes in one NCCL group), so every rank ends up OWNING one contiguous pixel slab of
 * every plane -- lb_filter_slab tells which.  The planes are padded to a multiple of 5040 pixels so that 1..10, 12, 14,
 * 15 and 16 ranks get equal slabs; other world sizes, and frames with cryptomatte AOVs (their ranked resolve reads
 * whole bucket rows), return LB_ERR_INVALID: use lb_filter_reduce.  Outside its slab a
 * rank's planes are partial afterwards: lb_imager_resolve refuses, lb_imager_resolve_gather is the resolve. */
LB_API int lb_filter_reduce_scatter(lb_camera *cam, lb_stream stream);
LB_API int lb_filter_slab(lb_camera *cam, size_t *first_pixel, size_t *n_pixels);
/* driver_process_bucket (lentil_imager.cpp:112-189) for the whole region, after lb_filter_reduce_scatter: every rank
 * resolves its own slab, the resolved slabs are gathered on `root` (ncclSend/ncclRecv) or on every rank (root < 0,
 * ncclAllGather).  rgba_out: device [yres][xres][4], may be NULL on ranks that do not receive.  With a single rank (no
 * communicator) it is a plain full-region resolve. */
LB_API int lb_imager_resolve_gather(lb_camera *cam, int aov, float *rgba_out, int root, lb_stream stream);
/* The multi-GPU combine AND driver_process_bucket (lentil_imager.cpp:112-189) in ONE kernel over peer memory (NVLink /
 * NVSwitch, the ranks' buffers mapped into each other through CUDA IPC): every rank owns one pixel slab of the frame; for its
 * slab it loads the partial planes of ALL ranks straight from their device memory, sums them in rank order (closest-filter
 * AOVs: takes the value of the rank that holds the smallest depth key, lentil.h:832-846), resolves, and stores the resolved
 * pixels into the image block of `root` (of every rank when root < 0).  Replaces lb_filter_reduce_scatter + n x
 * lb_imager_resolve_gather: no reduced planes are written back, no separate resolve and gather passes, every AOV in one launch.
 * Collective: every rank of the communicator calls it with the same arguments (ranks must also agree on the frames given to
 * lb_filter_begin -- the mappings are renewed when buffers are reallocated).  aov_indices: n_out AOVs (no cryptomatte AOVs).
 * images_out: NULL, or n_out device pointers ([yres][xres][4] floats; entries may be NULL) the receiving ranks copy the
 * images to; without it they are read in place through lb_imager_peer_image.  The partial planes are left untouched.
 * Slabs: even over the ranks; when one rank receives (root >= 0) and there are three or more ranks, the receiver owns no slab
 * (its NVLink ingress is the limit: it takes in the images only).  At most 8 ranks; single rank / no communicator: a plain
 * resolve of every listed AOV. */
LB_API int lb_imager_resolve_peer(lb_camera *cam, int n_out, const int *aov_indices, float *const *images_out, int root,
                                  lb_stream stream);
/* Device pointer to the [yres][xres][4] image of `aov` that the last lb_imager_resolve_peer left on this (receiving) rank;
 * valid until the next lb_imager_resolve_peer / lb_filter_begin with another frame size. */
LB_API int lb_imager_peer_image(lb_camera *cam, int aov, float **image);
LB_API int lb_comm_destroy(lb_camera *cam);

#ifdef __cplusplus
}
#endif
#endif /* LENTIL_B200_H */
