// filter_kernels.cuh — K2b: reverse-trace + splat of the redistributed samples, templated on the
// polynomial evaluator (table-driven or per-lens unrolled).
//
// Reference: the PolynomialOptics loop of filter_pixel, /root/reference/src/lentil_filter.cpp:248-300,
// calling Camera::trace_ray_bw_po (/root/reference/src/lentil.h:573-661) and Camera::add_to_buffer
// (lentil.h:823-851).
//
// Mapping: one WARP per redistributed source sample, fetched from the work list with an atomic
// counter (persistent CTAs).  The reference's loop
//     for (count = 0; count < samples && total < 5*samples; ++count, ++total) { ...fail -> --count }
// is sequential only through its stop rule; attempt `total` itself depends on nothing but `total`
// (seed = tea<8>(px*py+px, total+tries)).  Attempt t runs iff count_before(t) = t - fails_before(t) <
// samples, and fails_before(t) >= the failures already KNOWN among completed attempts, so every attempt
// t < min(samples + known_fails, 5*samples) is certain to be executed by the sequential loop.  Lanes
// therefore pull the next certain attempt as soon as their own ends ("work refill"): the set of splats
// is exactly the reference's, and a lane never waits for a slower neighbour.
//
// The Newton solve is advanced ONE iteration per loop trip for all busy lanes: iteration counts differ
// per attempt (mean ~30, up to 100), and a per-attempt inner loop left 2/3 of the lanes idle
// (profiles/r01_k2_filter_splat_ncu.txt: 10.97 of 32 lanes active).
#pragma once
#define LB_RED_PTX 1  // the PO splat kernels issue their reductions as PTX `red` (see filter_common.cuh)
#include "filter_common.cuh"

namespace lb {

// lanes that must be waiting before the (divergent) service phase of the splat loop runs
#ifndef LB_SERVICE_BATCH
#define LB_SERVICE_BATCH 8
#endif
constexpr int kServiceBatch = LB_SERVICE_BATCH;
constexpr int kSplatBlock = 128;  // threads per CTA of every splat kernel

// aperture point of attempt `total`, try `tries` (lentil.h:594-609)
LB_DEV void bw_aperture_sample(const CamConsts<float> &cam, uint32_t seed_base, uint32_t total, int tries, float &ax, float &ay) {
  if (!cam.enable_dof) { ax = ay = 0.f; return; }
  uint32_t seed = tea8(seed_base, total + (uint32_t)tries);
  if (cam.blades <= 2) {
    float ux, uy;
    if (cam.bokeh_n > 0) {
      // bokehSample(rng, rng, unit_disk, rng, rng): g++ evaluates the arguments right to left, so the
      // two unused stratification numbers draw first, then column, then row
      lcg_rng(seed); lcg_rng(seed);
      const float col = lcg_rng(seed);
      const float row = lcg_rng(seed);
      bokeh_sample(cam, row, col, ux, uy);
    } else {
      const float oy = lcg_rng(seed);
      const float ox = lcg_rng(seed);
      concentric_disk_sample(ox, oy, ux, uy);
    }
    ax = ux * cam.aperture_radius;
    ay = uy * cam.aperture_radius;
  } else {
    const float r2 = lcg_rng(seed);
    const float r1 = lcg_rng(seed);
    sample_triangular_aperture(ax, ay, r1, r2, cam.aperture_radius, cam.blades);
  }
}

// sensor position -> pixel index or -1 (lentil_filter.cpp:276-290), in double like the reference
LB_DEV int sensor_to_pixel(const FilterConsts &fc, float sx, float sy) {
  const double s0 = (double)sx / fc.sensor_half;
  const double s1 = (double)sy / fc.sensor_half * fc.aspect_full;
  const double p0 = (((s0 + 1.0) / 2.0) * (double)fc.xres_full) - (double)fc.region_min_x;
  const double p1 = (((-s1 + 1.0) / 2.0) * (double)fc.yres_full) - (double)fc.region_min_y;
  if ((p0 >= (double)fc.xres) || (p0 < 0) || (p1 >= (double)fc.yres) || (p1 < 0) || (p0 != p0) || (p1 != p1)) return -1;
  return (int)floor(p0) + (int)floor(p1) * fc.xres;
}

LB_DEV float channel_lambda(const FilterConsts &fc, int ch) {  // lentil_filter.cpp:254-267
  if (!(fc.abb_chromatic > 0.0f)) return 0.55f;
  if (ch == 0) return 0.35f + (1.0f - fc.abb_chromatic) * (0.55f - 0.35f);
  if (ch == 2) return 0.55f + fc.abb_chromatic * (0.85f - 0.55f);
  return 0.55f;
}

// Attempts [t_begin, ...) of one work item.  t_end >= 0: the range [t_begin, t_end) lies inside the first n_samples attempts, all
// certain to run (count_before(t) <= t < samples): a CHUNK, no stop rule needed.  t_end < 0: the tail of the sequential loop
// from attempt t_begin = n_samples on, with `known_fails` failures among the attempts before it.  Returns the failures found.
template <typename E>
LB_DEV int splat_work_item(const E &ev, const CamConsts<float> &cam, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s,
                           const WorkItem &w, FilterCounters *counters, uint64_t sample_base, int t_begin, int t_end, int known_fails) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const size_t i = w.sample;
  const int px = __ldg(s.px + i), py = __ldg(s.py + i);
  const float depth = __ldg(s.pos_cs + i).w;
  // -camera_space_sample_position * 10.0 (lentil_filter.cpp:271)
  const float target[3] = {(float)(-(double)w.csp[0] * 10.0), (float)(-(double)w.csp[1] * 10.0), (float)(-(double)w.csp[2] * 10.0)};
  const int samples = (int)w.n_samples;
  const int max_total = samples * 5;
  const float inv_samples = (float)(1.0 / (double)(float)samples);
  const float weight = 1.0f * s.inv_density * inv_samples;  // filter_weight * inverse_sample_density * inv_samples (:297)
  const uint32_t seed_base = (uint32_t)(px * py + px);
  const bool chroma = fc.abb_chromatic > 0.0f;
  const int nchan = chroma ? 3 : 1;

  // warp-uniform bookkeeping of the stop rule
  int next = t_begin;   // next attempt index (`total_samples_taken`) to hand out
  const int fails_before = known_fails;  // known_fails: channel failures among COMPLETED attempts
  unsigned n_splats = 0, n_attempts = 0, n_its = 0;
  // per-lane attempt state
  enum { IDLE = 0, RUNNING = 1, PENDING = 2 };  // PENDING: Newton loop ended, try not yet finished
  int state = IDLE;
  int t = 0, ch = 0, tries = 0;
  float ax = 0.f, ay = 0.f, lambda = 0.55f;
  LtState<float> st;
  lt_init(st);
  // (r02, measured and withdrawn: the aperture points of the next 32 attempts precomputed by the converged warp into a
  // shared-memory ring instead of inside the divergent service phase -- 58.53 vs 58.58 ms on the C3 frame, no gain.)

  for (;;) {
    // ---- service phase: finish ended tries, hand out new attempts, splat -------------------------------
    // It is divergent, scalar-ish code (transmittance polynomial, FP64 pixel mapping, RNG + CDF search), so it
    // runs only when kServiceBatch lanes are waiting for it (or nothing else is left to do): its cost is
    // shared by a batch of lanes instead of being paid by the whole warp for every single attempt.
    const int limit = t_end >= 0 ? t_end : min(samples + known_fails, max_total);
    const unsigned running = __ballot_sync(0xffffffffu, state == RUNNING);
    const unsigned pending = __ballot_sync(0xffffffffu, state == PENDING);
    const unsigned idle = ~(running | pending);
    const int fillable = min(__popc(idle), max(limit - next, 0));
    if (running == 0u && pending == 0u && fillable == 0) break;
    if (__popc(pending) + fillable >= kServiceBatch || running == 0u) {
      int pixel = -1;  // >= 0: this lane splats channel `splat_ch` in this service phase
      int splat_ch = 0, new_fails = 0;
      bool need_sample = false;
      if (state == PENDING) {  // lentil.h:633-658 after lens_lt_sample_aperture returned
        const float T = lt_finish(ev, cam, lambda, st);
        bool ok = T > 0.f;
        if (ok) {
          const float qx = st.x + st.dx * cam.bfl, qy = st.y + st.dy * cam.bfl;
          ok = !(qx * qx + qy * qy > cam.inner_pupil_r2);
        }
        bool channel_done = true;
        if (ok) {
          const float sx = st.x + st.dx * -cam.sensor_shift;  // shift sensor (lentil.h:654-655)
          const float sy = st.y + st.dy * -cam.sensor_shift;
          pixel = sensor_to_pixel(fc, sx, sy);
          splat_ch = ch;
          if (pixel < 0) new_fails = 1;  // off-image: `--count; continue` (lentil_filter.cpp:282-287)
          else ++n_splats;
        } else if (++tries <= cam.vignetting_retries) {  // vignetted: next try of the same attempt (lentil.h:592,634-645)
          channel_done = false;
        } else {
          new_fails = 1;  // trace_ray_bw_po returned false (lentil_filter.cpp:271-274)
        }
        if (channel_done) {
          if (++ch < nchan) {  // next colour channel of the same attempt: same seeds, other wavelength
            tries = 0;
            lambda = channel_lambda(fc, ch);
            ++n_attempts;
            state = RUNNING;
            need_sample = true;
          } else {
            state = IDLE;
          }
        } else {
          state = RUNNING;
          need_sample = true;
        }
      }
      known_fails += __reduce_add_sync(0xffffffffu, new_fails);
      // hand the next CERTAIN attempts to idle lanes (the failures just learnt may have raised the limit)
      {
        const unsigned free_lanes = __ballot_sync(0xffffffffu, state == IDLE);
        const int navail = max((t_end >= 0 ? t_end : min(samples + known_fails, max_total)) - next, 0);
        const int rank = __popc(free_lanes & lt_mask);
        if (state == IDLE && rank < navail) {
          state = RUNNING;
          t = next + rank;
          ch = 0;
          tries = 0;
          lambda = channel_lambda(fc, 0);
          ++n_attempts;
          need_sample = true;
        }
        next += min(__popc(free_lanes), navail);
      }
      if (need_sample) {  // the one copy of the aperture sampling code
        bw_aperture_sample(cam, seed_base, (uint32_t)t, tries, ax, ay);
        lt_init(st);
      }
      __syncwarp();
      // splat (warp-converged): the AOV values of the source sample are warp-uniform, the pixel is per lane
      if (__any_sync(0xffffffffu, pixel >= 0)) {
        float rgbw[3] = {1.f, 1.f, 1.f};
        if (chroma) { rgbw[0] = splat_ch == 0 ? 3.f : 0.f; rgbw[1] = splat_ch == 1 ? 3.f : 0.f; rgbw[2] = splat_ch == 2 ? 3.f : 0.f; }
        splat_all_aovs(fc, aovs, s, i, (float)samples, pixel, w.add_energy, depth, weight, rgbw, sample_base + i, counters);
      }
    }

    // ---- one Newton iteration of lt_sample_aperture for every running lane ------------------------------
    if (state == RUNNING) {
      lt_iterate(ev, cam, target, ax, ay, lambda, st);
      ++n_its;
      if (!lt_continue(st)) state = PENDING;
    }
  }
  n_splats = __reduce_add_sync(0xffffffffu, n_splats);
  n_attempts = __reduce_add_sync(0xffffffffu, n_attempts);
  n_its = __reduce_add_sync(0xffffffffu, n_its);
  if (lane == 0) {
    atomicAdd(&counters->splats, (unsigned long long)n_splats);
    atomicAdd(&counters->attempts, (unsigned long long)n_attempts);
    atomicAdd(&counters->newton_its, (unsigned long long)n_its);
  }
  return known_fails - fails_before;
}

// (A variant with two attempt slots per lane and FFMA2/FMUL2 Newton bodies was built and measured in r01: correct, 20-30 % fewer
// instructions, but 168 registers -> 12 warps/SM and 12 % slower; profiles/r01_k2_packed_experiment.txt.  Its successor is the
// mirror-packed body of lensgen/emit_folded.py: packed over the lens symmetry instead of over two attempts, same register count.)

// Persistent kernel body.  A work item's first n_samples attempts are certain, so they can be dealt out in CHUNKS to whichever
// warps are free; the warp that completes an item's last chunk knows the item's failure count and runs the tail of the
// sequential loop (the `--count` re-draws, a handful of attempts on a few items).  Chunk size: the whole item while the list
// holds several items per warp -- a warp that stays on one item keeps all its lanes refilled and drains them once per item --
// and smaller when it does not (a rank of an 8-GPU run holds 1/8 of the frame's highlights: 1.5 items of ~5 ms per resident
// warp, i.e. whole items would run in two rounds of which the second is half empty).
template <typename E>
LB_DEV void splat_persistent(const E &ev, const CamConsts<float> &cam, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s,
                             const WorkItem *__restrict__ work_, FilterCounters *counters, uint64_t sample_base) {
  WorkItem *work = const_cast<WorkItem *>(work_);  // chunks_done / fails are updated in place
  const int lane = threadIdx.x & 31;
  const unsigned n_work = *((volatile unsigned *)&aovs.work_heads[0]);
  const unsigned total_attempts = *((volatile unsigned *)&aovs.work_heads[2]);
  const unsigned warps = gridDim.x * (blockDim.x >> 5);
  const unsigned max_samples = *((volatile unsigned *)&aovs.work_heads[3]);
  unsigned chunk = 0x7fffffffu;  // whole items
  if (n_work < 4u * warps) chunk = max(64u, ((total_attempts / (4u * warps)) + 31u) & ~31u);
  // work-unit tickets: ticket -> (item, chunk) with a fixed number of chunk slots per item (1 for whole items); an item
  // shorter than its slots leaves empty tickets, which cost one atomic each
  const unsigned slots = chunk >= max_samples ? 1u : (max_samples + chunk - 1u) / chunk;
  const unsigned long long n_tickets = (unsigned long long)n_work * slots;
  for (;;) {
    unsigned ticket = 0;
    if (lane == 0) ticket = atomicAdd(&aovs.work_heads[1], 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (ticket >= n_tickets) break;
    const unsigned idx = ticket / slots, c = ticket - idx * slots;
    const WorkItem w = work[idx];
    if ((unsigned long long)c * chunk >= (unsigned long long)w.n_samples) continue;  // an empty slot of a short item
    const unsigned n_chunks = (unsigned)(((unsigned long long)w.n_samples + chunk - 1) / chunk);
    int t0 = (int)min((unsigned long long)c * chunk, (unsigned long long)w.n_samples);
    int t1 = (int)min((unsigned long long)t0 + chunk, (unsigned long long)w.n_samples);
    int known = 0;
    // ONE call site (the body is ~30 KB of straight-line code): first the chunk, then -- only for the warp that completes the
    // item's last chunk, and only if the item had failures -- the tail of the sequential loop
    for (bool tail = false;; tail = true) {
      const int fails = splat_work_item(ev, cam, fc, aovs, s, w, counters, sample_base, t0, t1, known);
      if (tail) break;
      unsigned done = 0, all_fails = 0;
      if (lane == 0) {
        if (fails) atomicAdd(&work[idx].fails, (unsigned)fails);
        __threadfence();
        done = atomicAdd(&work[idx].chunks_done, 1u) + 1u;
        if (done == n_chunks) {
          __threadfence();
          all_fails = *((volatile unsigned *)&work[idx].fails);
        }
      }
      done = __shfl_sync(0xffffffffu, done, 0);
      all_fails = __shfl_sync(0xffffffffu, all_fails, 0);
      if (done != n_chunks || all_fails == 0u) break;
      t0 = (int)w.n_samples;  // the attempts the failures bought (lentil_filter.cpp:272-287): sequential rule from here on
      t1 = -1;
      known = (int)all_fails;
    }
  }
}

}  // namespace lb
