// filter_common.cuh — AOV accumulate primitives shared by the classify and splat kernels.
#pragma once
#include "lens_device.cuh"
#include "lentil_internal.h"

namespace lb {

// Fire-and-forget reductions.  With LB_RED_PTX (the PO splat kernels, filter_kernels.cuh) they are spelled in PTX: there ptxas
// otherwise emits `ATOMG ... RZ` -- an atomic whose result is thrown away but still travels back, counted by ncu as op_atom --
// for the very atomicAdd calls it turns into REDG in the thin-lens and classify kernels.  Those keep the intrinsics: a volatile
// asm statement pins the reduction in the instruction stream and the thin-lens splat measured 2.96 instead of 2.76 ms with it.
#ifdef LB_RED_PTX
LB_DEV void red_add(float4 *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}
LB_DEV void red_add(float *p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v)); }
LB_DEV void red_min(unsigned long long *p, unsigned long long v) { asm volatile("red.global.min.u64 [%0], %1;" ::"l"(p), "l"(v)); }
#else
LB_DEV void red_add(float4 *p, float4 v) { atomicAdd(p, v); }
LB_DEV void red_add(float *p, float v) { atomicAdd(p, v); }
LB_DEV void red_min(unsigned long long *p, unsigned long long v) { atomicMin(p, v); }
#endif

// value of AOV `a` for sample i as filter_pixel gathers it (lentil_filter.cpp:206-234)
LB_DEV float4 aov_value(const AovSet &aovs, const SampleIO &s, int a, size_t i, float debug_val) {
  if (aovs.filter[a] == 2 /*LB_FILTER_CRYPTO: no value, lentil_filter.cpp:208*/) return make_float4(0.f, 0.f, 0.f, 0.f);
  if (aovs.role[a] == 2 /*LB_AOV_LENTIL_DEBUG*/) return make_float4(debug_val, debug_val, debug_val, debug_val);
  const float4 *src = aovs.values[a] ? aovs.values[a] : s.rgba;
  return __ldg(src + i);
}

// closest-filter ordering key: smaller |depth| wins; among equal depths the later sample wins, which is
// what the reference's sequential `abs(depth) <= zbuffer[px]` overwrite produces (lentil.h:832-837)
LB_DEV unsigned long long closest_key(float depth, uint64_t sample_global) {
  const float ad = fabsf(depth);
  const unsigned dbits = ad == 0.0f ? 0x7F800000u : __float_as_uint(ad);  // depth 0 behaves as "empty" there
  return ((unsigned long long)dbits << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)sample_global);
}

// Camera::add_to_buffer, lentil.h:823-851.  Gaussian AOVs: one 16-byte vector reduction
// (red.global.add.v4.f32) per AOV + one scalar reduction into filter_weight_buffer for the RGBA AOV.
// Closest AOVs: 64-bit atomicMin on the depth key; the winning sample's value is fetched afterwards
// (k_closest_gather).
LB_DEV void add_to_buffer(const AovSet &aovs, int a, unsigned pixel, float4 v, float add_energy, float depth,
                          float filter_weight, const float rgb_weight[3], uint64_t sample_global) {
  if (aovs.filter[a] == 0 /*gaussian*/) {
    if (aovs.role[a] == 1 /*RGBA*/) red_add(aovs.weight + pixel, filter_weight);
    float4 r;
    r.x = (v.x + add_energy) * filter_weight * rgb_weight[0];
    r.y = (v.y + add_energy) * filter_weight * rgb_weight[1];
    r.z = (v.z + add_energy) * filter_weight * rgb_weight[2];
    r.w = (v.w + add_energy) * filter_weight;
    // An all-zero contribution is not sent: adding +-0 changes no bit of a buffer (it starts at +0 and round-to-nearest sums never
    // produce -0), and most AOVs of a production set are zero for most samples (a per-light AOV holds one light's samples) --
    // at 8K x 10 AOVs the planes are not L2-resident and every reduction is a DRAM read-modify-write.  NaNs compare unequal: sent.
    if (aovs.add_zeros || r.x != 0.0f || r.y != 0.0f || r.z != 0.0f || r.w != 0.0f) red_add(aovs.buffer[a] + pixel, r);
  } else {
    if (aovs.role[a] != 2) red_min(aovs.zkey + pixel, closest_key(depth, sample_global));
    else if (v.x != 0.0f) red_min(aovs.zkey_debug + pixel, closest_key(depth, sample_global));
  }
}

// ---- cryptomatte ----------------------------------------------------------------------------------
// aov.crypto_hash_map[px][id] += w (lentil.h:817) on a fixed-size open-addressed table: claim a slot with a
// compare-and-swap on the id bits, then a float reduction on its weight.
// crypto_insert_at: the same probing from slot h, whose id has already been read as `cur` (crypto_add_hoisted).
LB_DEV void crypto_insert_at(uint32_t *__restrict__ k0, float *__restrict__ w0, int slots, uint32_t key, float w, unsigned h, uint32_t cur,
                             FilterCounters *counters) {
  for (int t = 0; t < slots; ++t) {
    if (t) cur = *(volatile uint32_t *)(k0 + h);
    if (cur == kCryptoFree) cur = atomicCAS(k0 + h, kCryptoFree, key);
    if (cur == kCryptoFree || cur == key) {
      red_add(w0 + h, w);
      return;
    }
    h = h + 1 == (unsigned)slots ? 0u : h + 1;
  }
  atomicAdd(&counters->crypto_dropped, 1ull);
}
LB_DEV uint32_t crypto_key_bits(float id) { return __float_as_uint(id == 0.0f ? 0.0f : id); }
LB_DEV unsigned crypto_home_slot(uint32_t key, int slots) { return ((key * 2654435761u) >> 15) % (unsigned)slots; }

LB_DEV void crypto_insert(uint32_t *__restrict__ keys, float *__restrict__ wgts, int slots, unsigned pixel, float id, float w,
                          FilterCounters *counters) {
  const uint32_t key = __float_as_uint(id == 0.0f ? 0.0f : id);  // -0 and +0 are one std::map key
  uint32_t *k0 = keys + (size_t)pixel * slots;
  float *w0 = wgts + (size_t)pixel * slots;
  unsigned h = ((key * 2654435761u) >> 15) % (unsigned)slots;
  for (int t = 0; t < slots; ++t) {
    uint32_t cur = *(volatile uint32_t *)(k0 + h);
    if (cur == kCryptoFree) cur = atomicCAS(k0 + h, kCryptoFree, key);
    if (cur == kCryptoFree || cur == key) {
      red_add(w0 + h, w);
      return;
    }
    h = h + 1 == (unsigned)slots ? 0u : h + 1;
  }
  atomicAdd(&counters->crypto_dropped, 1ull);
}

// add_to_buffer_cryptomatte (lentil.h:814-819) for AOV `a` and the cached {id, weight} list of source sample i.
// Called warp-converged from the splat kernels (`on`: this lane has a splat at `pixel`).
LB_DEV void crypto_add(const AovSet &aovs, int a, size_t i, bool on, unsigned pixel, float sample_weight, FilterCounters *counters) {
  const int stride = aovs.crypto_depth > 1 ? aovs.crypto_depth : 1;
  const float2 *e = aovs.crypto_cache[a] + i * (size_t)stride;
  if (on && a == aovs.crypto_first) red_add(&aovs.buffer[a][pixel].x, sample_weight);  // crypto_total_weight, one plane for every cryptomatte AOV
  for (int j = 0; j < stride; ++j) {
    const float2 kv = e[j];
    if (__float_as_uint(kv.x) == kCryptoFree) break;
    if (on) crypto_insert(aovs.crypto_key[a], aovs.crypto_wgt[a], aovs.crypto_slots, pixel, kv.x, kv.y * sample_weight, counters);
  }
}

// The same for the thin-lens splat kernels, whose attempts cost ~650 instructions: with cryptomatte AOVs they wait on the slot probes
// (ncu: long_scoreboard is the top stall, issue slots 40 % busy; profiles/r02_crypto_splat_ncu.txt), and
// one id's probe -> reduction does not depend on another's: the home slots of up to four ids are read back to back, then each id
// continues from the value read.  Same atomics as one crypto_insert per id; a slot's id never changes once claimed, so a value
// read early is still right when it matches, and a stale "free" just goes through the compare-and-swap as always.
LB_DEV void crypto_add_hoisted(const AovSet &aovs, int a, size_t i, bool on, unsigned pixel, float sample_weight, FilterCounters *counters) {
  const int stride = aovs.crypto_depth > 1 ? aovs.crypto_depth : 1;
  const float2 *e = aovs.crypto_cache[a] + i * (size_t)stride;
  if (on && a == aovs.crypto_first) red_add(&aovs.buffer[a][pixel].x, sample_weight);  // crypto_total_weight, one plane for every cryptomatte AOV
  const int slots = aovs.crypto_slots;
  const size_t row = on ? (size_t)pixel * slots : 0;  // `pixel` means nothing on a lane without a splat
  uint32_t *k0 = aovs.crypto_key[a] + row;
  float *w0 = aovs.crypto_wgt[a] + row;
  for (int j0 = 0; j0 < stride; j0 += 4) {
    float2 kv[4];
    uint32_t key[4], cur[4];
    unsigned h[4];
    int n = 0;  // warp-uniform: the list belongs to the source sample
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (j0 + k < stride && n == k) {
        kv[k] = e[j0 + k];
        if (__float_as_uint(kv[k].x) != kCryptoFree) n = k + 1;
      }
    }
    if (on) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < n) {
          key[k] = crypto_key_bits(kv[k].x);
          h[k] = crypto_home_slot(key[k], slots);
          cur[k] = *(volatile uint32_t *)(k0 + h[k]);
        }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k < n) crypto_insert_at(k0, w0, slots, key[k], kv[k].y * sample_weight, h[k], cur[k], counters);
    }
    if (n < 4) break;
  }
}

// every AOV of one splat (lentil_filter.cpp:295-298 / :442-445).  Warp-converged: the source sample i and with it
// the AOV values are warp-uniform, the target pixel is per lane (< 0: this lane has nothing to add).
// skip_aov: an AOV this lane has already accumulated elsewhere (the shared-memory window of the thin-lens tile kernel), or -1.
// kHoistProbes: crypto_add_hoisted for the cryptomatte AOVs (the thin-lens kernels; the polynomial-optics kernels are bound by their
// Newton trips and keep the form that costs them no registers).
template <bool kHoistProbes = false>
LB_DEV void splat_all_aovs(const FilterConsts &fc, const AovSet &aovs, const SampleIO &s, size_t i, float debug_val, int pixel,
                           float add_energy, float depth, float weight, const float rgb_weight[3], uint64_t sample_global,
                           FilterCounters *counters, int skip_aov = -1) {
  for (int a = 0; a < fc.n_aov; ++a) {
    if (aovs.filter[a] == 2) {
      if (kHoistProbes) crypto_add_hoisted(aovs, a, i, pixel >= 0, (unsigned)pixel, weight, counters);
      else crypto_add(aovs, a, i, pixel >= 0, (unsigned)pixel, weight, counters);
    } else {
      const float4 v = aov_value(aovs, s, a, i, debug_val);
      if (pixel >= 0 && a != skip_aov) add_to_buffer(aovs, a, (unsigned)pixel, v, add_energy, depth, weight, rgb_weight, sample_global);
    }
  }
}

}  // namespace lb
