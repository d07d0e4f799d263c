// filter_classify.cu — K2a: per-source-sample front half of filter_pixel (compiled with -fmad=false:
// the splat COUNT of a sample is an integer decision taken from float/double expressions, so they
// are evaluated operation by operation exactly as the reference's host code does).
//
// Reference: /root/reference/src/lentil_filter.cpp:105-246 — read the sample, decide `redistribute`,
// derive the sample count from CoC^2 x luminance, gather the AOV values, and for samples that are not
// redistributed do filter_and_add_to_buffer_new (lentil.h:938-955) right here.  Redistributed samples
// are appended to a work list consumed by the splat kernel (filter_kernels.cu).
#include "filter_common.cuh"

namespace lb {

// Camera::get_coc_thinlens, lentil.h:674-692 (float arithmetic, as there)
__device__ float get_coc_thinlens(const FilterConsts &fc, float z) {
  const float fl = fc.focal_length;
  const float image_dist_samplepos = (-fl * z) / (-fl + z);
  const float image_dist_focusdist = (-fl * -fc.coc_focus_distance) / (-fl + -fc.coc_focus_distance);
  return fabsf((fc.coc_aperture_radius * (image_dist_samplepos - image_dist_focusdist)) / image_dist_samplepos);
}

// Camera::additional_luminance_soft_trans, lentil.h:1128-1138
__device__ float additional_luminance_soft_trans(const FilterConsts &fc, float lum) {
  const double lo = fc.bidir_add_energy_minimum_luminance;
  if ((double)lum > lo && (double)lum < lo + (double)fc.bidir_add_energy_transition) {
    const float perc = (float)(((double)lum - lo) / (double)fc.bidir_add_energy_transition);
    return fc.bidir_add_energy * perc;
  } else if ((double)lum > lo + (double)fc.bidir_add_energy_transition) {
    return fc.bidir_add_energy;
  }
  return 0.0f;
}

// cryptomatte_construct_cache (lentil.h:779-811) for AOV `a` of sample i: walks the depth sub-samples in order,
// merging equal ids as the reference's std::map<float,float> does, into a packed {id, weight} list.  Float arithmetic
// operation by operation as there (this unit is -fmad=false).
// A redistributed sample leaves the list in the batch scratch for the splat kernel.  A pass-through sample (`!redistribute`)
// adds it to its own pixel right here -- add_to_buffer_cryptomatte (lentil.h:814-819) from filter_and_add_to_buffer_new
// (:938-955), what crypto_add (filter_common.cuh) does for a splat -- and writes nothing: 33 M samples x 3 AOVs x 4 entries would
// be 3.2 GB of scratch that nothing reads again.
// (kAddPass = false: every sample writes its list and pass-through samples add it later through crypto_add, the form the kernel
// for frames WITHOUT cryptomatte AOVs keeps so that it stays the kernel it was.)
template <bool kAddPass>
__device__ void crypto_build_cache(const AovSet &aovs, const SampleIO &s, int a, size_t i, bool redistribute, unsigned pixel,
                                   FilterCounters *counters) {
  const int D = aovs.crypto_depth;
  const int stride = D > 1 ? D : 1;
  float key[kCryptoMaxDepth], wgt[kCryptoMaxDepth];
  int m = 0;
  auto add = [&](float k, float v) {
    for (int j = 0; j < m; ++j)
      if (key[j] == k) { wgt[j] += v; return; }
    // at most D distinct ids exist; an id that never compares equal to itself (NaN: Cryptomatte never emits one) could
    // otherwise claim a ninth entry through the trailing quota
    if (m == kCryptoMaxDepth) return;
    key[m] = k; wgt[m] = 0.0f + v; ++m;
  };
  float iterative_transparency_weight = 1.0f, quota = 1.0f, sample_value = 0.0f;
  const float *ids = aovs.crypto_ids[a];
  int count = !ids ? 0 : (s.crypto_count ? min((int)__ldg(s.crypto_count + i), D) : D);
  for (int d = 0; d < count; ++d) {
    const float sub_sample_opacity = s.crypto_opacity ? __ldg(s.crypto_opacity + i * (size_t)D + d) : 0.0f;
    sample_value = __ldg(ids + i * (size_t)D + d);
    const float sub_sample_weight = sub_sample_opacity * iterative_transparency_weight;
    iterative_transparency_weight *= (1.0f - sub_sample_opacity);
    quota -= sub_sample_weight;
    add(sample_value, sub_sample_weight);
  }
  if ((double)quota > 0.0) add(sample_value, quota);
  if (!kAddPass || redistribute) {
    float2 *out = aovs.crypto_cache[a] + i * (size_t)stride;
    for (int j = 0; j < stride; ++j) out[j] = j < m ? make_float2(key[j], wgt[j]) : make_float2(__uint_as_float(kCryptoFree), 0.0f);
  } else {
    const float sample_weight = s.inv_density;
    if (a == aovs.crypto_first) red_add(&aovs.buffer[a][pixel].x, sample_weight);  // crypto_total_weight, one plane for all
    for (int j = 0; j < m && j < stride; ++j)
      crypto_insert(aovs.crypto_key[a], aovs.crypto_wgt[a], aovs.crypto_slots, pixel, key[j], wgt[j] * sample_weight, counters);
  }
}

// kCryptoFrame: the frame has cryptomatte AOVs (picked by the host).
template <bool kCryptoFrame>
__global__ void __launch_bounds__(256)
k_filter_classify(const __grid_constant__ FilterConsts fc, const __grid_constant__ AovSet aovs, const __grid_constant__ SampleIO s,
                  WorkItem *__restrict__ work, FilterCounters *__restrict__ counters, uint64_t sample_base) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool active = i < s.n;
  if (active) {  // samples filtered for a pixel outside the region cannot be accumulated anywhere: dropped (the arrays come
                 // from an external host; the reference would index out of bounds, lentil.h:941)
    const int px = __ldg(s.px + i), py = __ldg(s.py + i);
    if (px < 0 || py < 0 || px >= fc.xres || py >= fc.yres) active = false;
  }
  bool redistribute = active;
  int samples = 0;
  float add_energy = 0.f;
  float csp[3] = {0.f, 0.f, 0.f};
  float depth = 0.f, debug_val = 0.f;
  unsigned pixel = 0u;
  if (active) {
    float4 sample = __ldg(s.rgba + i);
    const float4 pz = __ldg(s.pos_cs + i);
    csp[0] = pz.x; csp[1] = pz.y; csp[2] = pz.z;
    depth = pz.w;
    // lentil_filter.cpp:119-133
    const bool small = fabsf(csp[0]) < 1.0e-4f && fabsf(csp[1]) < 1.0e-4f && fabsf(csp[2]) < 1.0e-4f;
    if ((depth == 1.0e30f || small) && fc.enable_skydome) {
      float4 rd = s.raydir ? __ldg(s.raydir + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (rd.x == 0.f && rd.y == 0.f && rd.z == 0.f) redistribute = false;
      else { csp[0] = rd.x * 99999999.0f; csp[1] = rd.y * 99999999.0f; csp[2] = rd.z * 99999999.0f; }
    }
    if ((depth == 1.0e30f || small) && !fc.enable_skydome) redistribute = false;
    const uint32_t flags = s.flags ? __ldg(s.flags + i) : 0u;
    if (flags & 1u) redistribute = false;  // volume_in_sample (:135-137)
    {  // AiM4PointByMatrixMult(world_to_camera_matrix, sample_pos_ws) (:141-142); identity unless the batch carries a matrix
      const float x = csp[0], y = csp[1], z = csp[2];
      csp[0] = x * s.w2c[0][0] + y * s.w2c[1][0] + z * s.w2c[2][0] + s.w2c[3][0];
      csp[1] = x * s.w2c[0][1] + y * s.w2c[1][1] + z * s.w2c[2][1] + s.w2c[3][1];
      csp[2] = x * s.w2c[0][2] + y * s.w2c[1][2] + z * s.w2c[2][2] + s.w2c[3][2];
    }
    csp[0] *= fc.unit_mult; csp[1] *= fc.unit_mult; csp[2] *= fc.unit_mult;  // :143-148
    if (s.transmission) {  // :152-159
      const float4 t = __ldg(s.transmission + i);
      const bool transmitted = fc.enable_bidir_transmission ? false : (fmaxf(t.x, fmaxf(t.y, t.z)) > 0.0f);
      if (transmitted) { sample.x -= t.x; sample.y -= t.y; sample.z -= t.z; redistribute = false; }
    }
    const float lum = (float)((double)(sample.x + sample.y + sample.z) / 3.0);  // :161
    if (flags & 2u) redistribute = false;  // lentil_bidir_ignore (:162-164)
    if (fc.bidir_add_energy > 0.0f) add_energy = additional_luminance_soft_trans(fc, lum);
    // :177-179
    const float luminance_mult = (float)fmax(0.0, sqrt((double)fminf(lum, 20.0f)) * (double)fc.bidir_sample_mult);
    const float coc = get_coc_thinlens(fc, csp[2]);
    const float cy = coc * (float)(unsigned)fc.yres;
    const float coc_squared_pixels = (float)(((double)cy * (double)cy) * ((double)luminance_mult * (double)luminance_mult) * 0.00001);
    if (coc < 0.4f) redistribute = false;  // :183-187
    samples = (int)ceilf(coc_squared_pixels * s.inv_density);  // :197
    float sf = (float)samples;                                  // clamp(float,4,2000), global.h:8-12
    if (sf < 4.f) sf = 4.f;
    if (sf > 2000.f) sf = 2000.f;
    samples = (int)sf;
    debug_val = (float)(samples * (redistribute ? 1 : 0));  // lentil_debug value, taken at :209-211
    if (fc.camera_type == 1 && (double)fabsf(csp[2]) < fc.lens_length_tenth) redistribute = false;  // :240, PolynomialOptics case only
    const int px = __ldg(s.px + i), py = __ldg(s.py + i);
    if (!kCryptoFrame)
      for (int a = 0; a < fc.n_aov; ++a)  // lentil_filter.cpp:167-169
        if (aovs.filter[a] == 2) crypto_build_cache<false>(aovs, s, a, i, redistribute, 0u, counters);
    pixel = (unsigned)fc.xres * (unsigned)py + (unsigned)px;
    if (kCryptoFrame)
      for (int a = 0; a < fc.n_aov; ++a)
        if (aovs.filter[a] == 2) crypto_build_cache<true>(aovs, s, a, i, redistribute, pixel, counters);
    if (aovs.debug_samples) aovs.debug_samples[i] = (uint16_t)debug_val;
  }
  const int lane = threadIdx.x & 31;
  // filter_and_add_to_buffer_new (lentil.h:938-955) for the samples that are not redistributed: every AOV, own pixel,
  // weight inv_density.  The renderer hands over a pixel's samples together, so consecutive lanes mostly share the
  // pixel: gaussian AOVs are summed over each run of equal pixels inside the warp first (segmented shuffle
  // reduction) and added with ONE reduction per run -- 16 same-address reductions per pixel at 16 spp otherwise
  // serialise in L2.  Float sums are order-free here as in the reference (its threads race, lentil.h:828-829).
  const bool pass = active && !redistribute;
  if (__any_sync(0xffffffffu, pass)) {
    const unsigned key = pass ? pixel : 0xFFFFFFFFu;
    const unsigned prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = lane == 0 || prev != key;
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned above = lane < 31 ? heads >> (lane + 1) : 0u;  // head flags of the lanes after this one
    for (int a = 0; a < fc.n_aov; ++a) {
      if (aovs.filter[a] == 2) {  // cryptomatte tables: per sample; kCryptoFrame: already added where the list was built
        if (!kCryptoFrame && pass) crypto_add(aovs, a, i, true, pixel, s.inv_density, counters);
      } else if (aovs.filter[a] == 1) {  // closest: depth key
        if (pass) {
          const float white[3] = {1.f, 1.f, 1.f};
          add_to_buffer(aovs, a, pixel, aov_value(aovs, s, a, i, debug_val), 0.0f, depth, s.inv_density, white, sample_base + i);
        }
      } else {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        float w = 0.f;
        if (pass) {  // (aov_value + 0) * filter_weight * white, per sample as add_to_buffer rounds it
          const float4 v = aov_value(aovs, s, a, i, debug_val);
          w = s.inv_density;
          r = make_float4((v.x + 0.0f) * w * 1.f, (v.y + 0.0f) * w * 1.f, (v.z + 0.0f) * w * 1.f, (v.w + 0.0f) * w);
        }
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const float tx = __shfl_down_sync(0xffffffffu, r.x, off), ty = __shfl_down_sync(0xffffffffu, r.y, off);
          const float tz = __shfl_down_sync(0xffffffffu, r.z, off), tw = __shfl_down_sync(0xffffffffu, r.w, off);
          const float ww = __shfl_down_sync(0xffffffffu, w, off);
          if (lane + off < 32 && (above & ((1u << off) - 1u)) == 0u) {  // lane + off belongs to this lane's run
            r.x += tx; r.y += ty; r.z += tz; r.w += tw; w += ww;
          }
        }
        if (pass && head) {
          if (aovs.role[a] == 1 /*RGBA*/) red_add(aovs.weight + pixel, w);
          if (aovs.add_zeros || r.x != 0.0f || r.y != 0.0f || r.z != 0.0f || r.w != 0.0f) red_add(aovs.buffer[a] + pixel, r);  // see add_to_buffer
        }
      }
    }
  }
  // append redistributed samples to the work list, one atomic per warp
  const unsigned mask = __ballot_sync(0xffffffffu, redistribute);
  const unsigned warp_samples = __reduce_add_sync(0xffffffffu, redistribute ? (unsigned)samples : 0u);
  const unsigned warp_max = __reduce_max_sync(0xffffffffu, redistribute ? (unsigned)samples : 0u);
  unsigned base = 0;
  if (lane == 0 && mask) {
    base = atomicAdd(&aovs.work_heads[0], (unsigned)__popc(mask));
    atomicAdd(&aovs.work_heads[2], warp_samples);  // attempts the splat kernel is certain to run: sizes its work units
    atomicMax(&aovs.work_heads[3], warp_max);
    atomicAdd(&counters->redistributed, (unsigned long long)__popc(mask));
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  if (redistribute) {
    WorkItem w;
    w.sample = (uint32_t)i;
    w.n_samples = (uint32_t)samples;
    w.add_energy = add_energy;
    w.csp[0] = csp[0]; w.csp[1] = csp[1]; w.csp[2] = csp[2];
    w.chunks_done = w.fails = 0u;
    work[base + __popc(mask & ((1u << lane) - 1u))] = w;
  }
  // (samples consumed and pass-through adds are counted on the host: three same-address reductions per warp here cost
  // more than the rest of the kernel -- 3.1 M serialised L2 operations for a 33 M-sample batch)
}

cudaError_t launch_filter_classify(const FilterConsts &fc, const AovSet &aovs, const SampleIO &s, WorkItem *work,
                                   FilterCounters *counters, uint64_t sample_base, cudaStream_t stream) {
  if (s.n == 0) return cudaSuccess;
  const unsigned grid = (unsigned)((s.n + 255) / 256);
  if (aovs.crypto_first >= 0) k_filter_classify<true><<<grid, 256, 0, stream>>>(fc, aovs, s, work, counters, sample_base);
  else k_filter_classify<false><<<grid, 256, 0, stream>>>(fc, aovs, s, work, counters, sample_base);
  return cudaGetLastError();
}

}  // namespace lb
