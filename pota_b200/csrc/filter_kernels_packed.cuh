// filter_kernels_packed.cuh — EXPERIMENTAL, not part of the default build: K2b with two attempt slots per lane and
// the Newton polynomials on the packed FP32 instructions (FFMA2/FMUL2).  Compiled into the per-lens units only when the
// library is built with LB_K2_PACKED=1 (pota_b200/lensgen/emit_cuda.py); the product path is splat_work_item in
// filter_kernels.cuh.  Measured in r01 (profiles/r01_k2_packed_experiment.txt): parity-green (same splat and attempt
// counts as the scalar kernel) but slower, because it needs 168 registers (12 warps/SM).  Kept so that the next round
// can continue from tested code: the levers are register pressure (one accumulator per polynomial, LB_PACKED_ACC=1;
// splitting lt_all2 into its aperture and outer-pupil groups) and the service batch.
#pragma once
#include "filter_kernels.cuh"

namespace lb {

// ---- two attempts per lane: the Newton polynomials run on the packed FP32 instructions ----------------------------
// ncu r01 (profiles/r01_k2_final_ncu.txt): the scalar kernel is issue-bound (issue slots 80 % busy, FMA pipe 58 %) and
// 75 % of what it issues is FFMA/FMUL of lt_all.  Here every lane carries TWO attempt slots of the same work item (64
// per warp); the solver state of the pair lives in float2 registers and one trip advances both through the FFMA2/FMUL2
// body Eval::lt_all2.  Everything that is not polynomial exists ONCE in the code and runs per half in rolled loops
// over a warp-uniform half index (the halves are read with selects): a first version that unrolled those loops was a
// third slower than the scalar kernel -- 70 KB of code, 24 % of the stall samples waiting for instruction fetch
// (the SM's instruction cache holds 32 KB).  The stop rule, the certain-attempt refill and the service phase are the
// ones above with the slot set doubled.
constexpr int kServiceBatch2 = 16;

struct LtState2 {  // LtState<float> of two solves
  float2 x, y, dx, dy, sqr_err, sqr_ap_err, out[4];
  int error[2], k[2];
};
LB_DEV void put_half(float2 &v, int h, float f) { if (h) v.y = f; else v.x = f; }
LB_DEV LtState<float> get_half(const LtState2 &S, int h) {
  LtState<float> q;
  q.x = lo_hi(S.x, h); q.y = lo_hi(S.y, h); q.dx = lo_hi(S.dx, h); q.dy = lo_hi(S.dy, h);
  q.sqr_err = lo_hi(S.sqr_err, h); q.sqr_ap_err = lo_hi(S.sqr_ap_err, h);
#pragma unroll
  for (int k = 0; k < 4; ++k) q.out[k] = lo_hi(S.out[k], h);
  q.error = pick(S.error, h); q.k = pick(S.k, h);
  return q;
}
LB_DEV void set_half(LtState2 &S, int h, const LtState<float> &q) {
  put_half(S.x, h, q.x); put_half(S.y, h, q.y); put_half(S.dx, h, q.dx); put_half(S.dy, h, q.dy);
  put_half(S.sqr_err, h, q.sqr_err); put_half(S.sqr_ap_err, h, q.sqr_ap_err);
#pragma unroll
  for (int k = 0; k < 4; ++k) put_half(S.out[k], h, q.out[k]);
  put(S.error, h, q.error); put(S.k, h, q.k);
}

// packed float helpers: two independent FP32 lanes per instruction (FMUL2/FFMA2/FADD2); division and square root stay
// the IEEE scalar sequences, once per half (independent, so they interleave)
LB_DEV float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
LB_DEV float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
LB_DEV float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
LB_DEV float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.0f, -1.0f), a); }  // a - b, one rounding
LB_DEV float2 nms2(float2 a, float2 b, float2 c) { return __ffma2_rn(__fmul2_rn(a, make_float2(-1.0f, -1.0f)), b, c); }  // c - a*b
LB_DEV float2 bc2(float a) { return make_float2(a, a); }
LB_DEV float2 sqrt2(float2 a) { return make_float2(sqrtf(a.x), sqrtf(a.y)); }
LB_DEV float2 rcp2(float2 a) { return make_float2(1.0f / a.x, 1.0f / a.y); }
LB_DEV float2 div2(float2 a, float b) { return make_float2(a.x / b, a.y / b); }
LB_DEV float2 max2(float2 a, float b) { return make_float2(fmaxf(a.x, b), fmaxf(a.y, b)); }

// lt_iterate_tail for both halves at once (spherical outer pupil): same operations as sphere_to_cs / cs_to_sphere /
// the two 2x2 Newton updates of lens_device.cuh, written on float2.  `on` halves commit, the others keep their state.
LB_DEV void lt_iterate_tail2(const CamConsts<float> &cam, const float scene[3], float2 ax, float2 ay, const float2 ap[2],
                             const float2 J[4], const float2 K[4], const float2 out[4], LtState2 &S, bool on0, bool on1) {
  // aperture error and Newton update of the direction
  const float2 da0 = sub2(ax, ap[0]), da1 = sub2(ay, ap[1]);
  const float2 sqr_ap_err = fma2(da0, da0, mul2(da1, da1));
  const float2 invdetap = rcp2(nms2(J[1], J[2], mul2(J[0], J[3])));
  const float2 m1 = bc2(-1.0f);
  float2 ndx = fma2(mul2(J[3], invdetap), da0, S.dx);
  ndx = fma2(mul2(mul2(J[1], m1), invdetap), da1, ndx);
  float2 ndy = fma2(mul2(mul2(J[2], m1), invdetap), da0, S.dy);
  ndy = fma2(mul2(J[0], invdetap), da1, ndy);
  // position part of sphere_to_cs(out, R) (lens.h:99-125; pos.z in the cancellation-free form of lens_device.cuh);
  // the solver never reads the direction it would return
  const float R = cam.outer_R;
  const float2 px = out[0], py = out[1], dx = out[2], dy = out[3];
  const float2 r2 = fma2(px, px, mul2(py, py));
  const float2 nx = div2(px, R), ny = div2(py, R);
  const float2 nz = div2(sqrt2(max2(sub2(bc2(R * R), r2), 0.f)), fabsf(R));
  const float2 rr = make_float2(fminf(r2.x, R * R), fminf(r2.y, R * R));
  const float2 den = mul2(bc2(R), add2(nz, bc2(1.0f)));
  const float2 pz = make_float2(-rr.x / den.x, -rr.y / den.y);
  // view vector to the scene point, then cs_to_sphere(pos, view, -R, R) (lens.h:127-153)
  const float2 v0 = sub2(bc2(scene[0]), px), v1 = sub2(bc2(scene[1]), py), v2 = sub2(bc2(scene[2]), pz);
  const float2 nz2 = div2(add2(pz, bc2(R)), R);
  const float2 nzp = make_float2(fabsf(nz2.x), fabsf(nz2.y));
  const float2 dl = rcp2(sqrt2(fma2(v0, v0, fma2(v1, v1, mul2(v2, v2)))));
  const float2 d0 = mul2(v0, dl), d1 = mul2(v1, dl), d2 = mul2(v2, dl);
  const float2 il2 = rcp2(sqrt2(fma2(nzp, nzp, mul2(nx, nx))));
  const float2 fx0 = mul2(nzp, il2), fx2 = mul2(mul2(nx, m1), il2);
  const float2 fy0 = mul2(ny, fx2), fy1 = nms2(nx, fx2, mul2(nzp, fx0)), fy2 = mul2(mul2(ny, m1), fx0);
  const float2 odx = fma2(d0, fx0, mul2(d2, fx2));
  const float2 ody = fma2(d0, fy0, fma2(d1, fy1, mul2(d2, fy2)));
  const float2 do0 = sub2(odx, dx), do1 = sub2(ody, dy);
  const float2 sqr_err = fma2(do0, do0, mul2(do1, do1));
  const float2 invdet = rcp2(nms2(K[1], K[2], mul2(K[0], K[3])));
  const float2 c72 = bc2(0.72f);
  float2 nxs = fma2(mul2(c72, mul2(K[3], invdet)), do0, S.x);
  nxs = fma2(mul2(c72, mul2(mul2(K[1], m1), invdet)), do1, nxs);
  float2 nys = fma2(mul2(c72, mul2(mul2(K[2], m1), invdet)), do0, S.y);
  nys = fma2(mul2(c72, mul2(K[0], invdet)), do1, nys);
  // flags and commit, per half
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (!(h ? on1 : on0)) continue;
    const float se = lo_hi(sqr_err, h), sa = lo_hi(sqr_ap_err, h), o0 = lo_hi(out[0], h), o1 = lo_hi(out[1], h);
    int error = S.error[h];
    if (se > lo_hi(S.sqr_err, h)) error |= 1;
    if (sa > lo_hi(S.sqr_ap_err, h)) error |= 2;
    if (o0 != o0) error |= 4;
    if (o0 * o0 + o1 * o1 > cam.outer_pupil_r2) error |= 16;
    if (S.k[h] < 10) error = 0;
    S.error[h] = error;
    S.k[h] += 1;
    put_half(S.x, h, lo_hi(nxs, h)); put_half(S.y, h, lo_hi(nys, h));
    put_half(S.dx, h, lo_hi(ndx, h)); put_half(S.dy, h, lo_hi(ndy, h));
    put_half(S.sqr_err, h, se); put_half(S.sqr_ap_err, h, sa);
#pragma unroll
    for (int k = 0; k < 4; ++k) put_half(S.out[k], h, lo_hi(out[k], h));
  }
}

template <typename E>
LB_DEV void splat_work_item2(const E &ev, const CamConsts<float> &cam, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s,
                             const WorkItem &w, FilterCounters *counters, uint64_t sample_base) {
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const size_t i = w.sample;
  const int px = __ldg(s.px + i), py = __ldg(s.py + i);
  const float depth = __ldg(s.pos_cs + i).w;
  const float target[3] = {(float)(-(double)w.csp[0] * 10.0), (float)(-(double)w.csp[1] * 10.0), (float)(-(double)w.csp[2] * 10.0)};
  const int samples = (int)w.n_samples;
  const int max_total = samples * 5;
  const float inv_samples = (float)(1.0 / (double)(float)samples);
  const float weight = 1.0f * s.inv_density * inv_samples;
  const uint32_t seed_base = (uint32_t)(px * py + px);
  const bool chroma = fc.abb_chromatic > 0.0f;
  const int nchan = chroma ? 3 : 1;

  int next = 0, known_fails = 0;
  unsigned n_splats = 0, n_attempts = 0, n_its = 0;
  enum { IDLE = 0, RUNNING = 1, PENDING = 2 };
  int state[2] = {IDLE, IDLE};
  int t[2] = {0, 0}, ch[2] = {0, 0}, tries[2] = {0, 0};
  float ax[2] = {0.f, 0.f}, ay[2] = {0.f, 0.f};
  float2 lambda = make_float2(0.55f, 0.55f);
  LtState2 S;
  {
    LtState<float> q;
    lt_init(q);
    set_half(S, 0, q);
    set_half(S, 1, q);
  }

  for (;;) {
    const int limit = min(samples + known_fails, max_total);
    const unsigned run0 = __ballot_sync(0xffffffffu, state[0] == RUNNING), run1 = __ballot_sync(0xffffffffu, state[1] == RUNNING);
    const unsigned pen0 = __ballot_sync(0xffffffffu, state[0] == PENDING), pen1 = __ballot_sync(0xffffffffu, state[1] == PENDING);
    const int n_pending = __popc(pen0) + __popc(pen1);
    const int n_idle = 64 - __popc(run0) - __popc(run1) - n_pending;
    const bool any_running = (run0 | run1) != 0u;
    const int fillable = min(n_idle, max(limit - next, 0));
    if (!any_running && n_pending == 0 && fillable == 0) break;
    if (n_pending + fillable >= kServiceBatch2 || !any_running) {
      int pixel[2] = {-1, -1}, splat_ch[2] = {0, 0};
      bool need_sample[2] = {false, false};
      int new_fails = 0;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {  // finish ended tries (lentil.h:633-658)
        if (pick(state, h) != PENDING) continue;
        const LtState<float> q = get_half(S, h);
        const float lam = lo_hi(lambda, h);
        const float T = lt_finish(ev, cam, lam, q);
        bool ok = T > 0.f;
        if (ok) {
          const float qx = q.x + q.dx * cam.bfl, qy = q.y + q.dy * cam.bfl;
          ok = !(qx * qx + qy * qy > cam.inner_pupil_r2);
        }
        int c = pick(ch, h), tr = pick(tries, h), st_new;
        bool channel_done = true, sample = false;
        if (ok) {
          const float sx = q.x + q.dx * -cam.sensor_shift;
          const float sy = q.y + q.dy * -cam.sensor_shift;
          const int pix = sensor_to_pixel(fc, sx, sy);
          put(pixel, h, pix);
          put(splat_ch, h, c);
          if (pix < 0) ++new_fails;
          else ++n_splats;
        } else if (++tr <= cam.vignetting_retries) {
          channel_done = false;
        } else {
          ++new_fails;
        }
        if (channel_done) {
          if (++c < nchan) {
            tr = 0;
            put_half(lambda, h, channel_lambda(fc, c));
            ++n_attempts;
            st_new = RUNNING;
            sample = true;
          } else {
            st_new = IDLE;
          }
        } else {
          st_new = RUNNING;
          sample = true;
        }
        put(ch, h, c);
        put(tries, h, tr);
        put(state, h, st_new);
        put(need_sample, h, sample);
      }
      known_fails += __reduce_add_sync(0xffffffffu, new_fails);
      {  // hand the next CERTAIN attempts to idle slots
        const unsigned free0 = __ballot_sync(0xffffffffu, state[0] == IDLE), free1 = __ballot_sync(0xffffffffu, state[1] == IDLE);
        const int navail = max(min(samples + known_fails, max_total) - next, 0);
        const int rank[2] = {__popc(free0 & lt_mask), __popc(free0) + __popc(free1 & lt_mask)};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (state[h] == IDLE && rank[h] < navail) {
            state[h] = RUNNING;
            t[h] = next + rank[h];
            ch[h] = 0;
            tries[h] = 0;
            put_half(lambda, h, channel_lambda(fc, 0));
            ++n_attempts;
            need_sample[h] = true;
          }
        }
        next += min(__popc(free0) + __popc(free1), navail);
      }
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {  // the one copy of the aperture sampling code
        if (!pick(need_sample, h)) continue;
        float a, b;
        bw_aperture_sample(cam, seed_base, (uint32_t)pick(t, h), pick(tries, h), a, b);
        put(ax, h, a);
        put(ay, h, b);
        LtState<float> q;
        lt_init(q);
        set_half(S, h, q);
      }
      __syncwarp();
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {  // splat (warp-converged)
        const int pix = pick(pixel, h);
        if (__any_sync(0xffffffffu, pix >= 0)) {
          const int sc = pick(splat_ch, h);
          float rgbw[3] = {1.f, 1.f, 1.f};
          if (chroma) { rgbw[0] = sc == 0 ? 3.f : 0.f; rgbw[1] = sc == 1 ? 3.f : 0.f; rgbw[2] = sc == 2 ? 3.f : 0.f; }
          splat_all_aovs(fc, aovs, s, i, (float)samples, pix, w.add_energy, depth, weight, rgbw, sample_base + i, counters);
        }
      }
    }

    // ---- one Newton iteration for every running slot: 14 polynomials packed, the rest per half ------------------
    if (state[0] == RUNNING || state[1] == RUNNING) {
      const float2 b[5] = {S.x, S.y, S.dx, S.dy, lambda};
      float2 ap[2], J[4], out[4], K[4];
      ev.lt_all2(b, ap, J, out, K);
      if (cam.outer_geom == 0) {  // spherical outer pupil: the rest of the trip on float2 as well
        const bool on0 = state[0] == RUNNING, on1 = state[1] == RUNNING;
        lt_iterate_tail2(cam, target, make_float2(ax[0], ax[1]), make_float2(ay[0], ay[1]), ap, J, K, out, S, on0, on1);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (!(h ? on1 : on0)) continue;
          ++n_its;
          const float eps = 1e-8f;  // lt_continue
          if (!(S.k[h] < 100 && (lo_hi(S.sqr_err, h) > eps || lo_hi(S.sqr_ap_err, h) > eps) && S.error[h] == 0)) state[h] = PENDING;
        }
        continue;
      }
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {  // cylindrical pupils: scalar tail per half
        if (pick(state, h) != RUNNING) continue;
        LtState<float> q = get_half(S, h);
        const float aph[2] = {lo_hi(ap[0], h), lo_hi(ap[1], h)};
        const float Jh[4] = {lo_hi(J[0], h), lo_hi(J[1], h), lo_hi(J[2], h), lo_hi(J[3], h)};
        const float Kh[4] = {lo_hi(K[0], h), lo_hi(K[1], h), lo_hi(K[2], h), lo_hi(K[3], h)};
#pragma unroll
        for (int k = 0; k < 4; ++k) q.out[k] = lo_hi(out[k], h);
        lt_iterate_tail(cam, target, pick(ax, h), pick(ay, h), aph, Jh, Kh, q);
        set_half(S, h, q);
        ++n_its;
        if (!lt_continue(q)) put(state, h, (int)PENDING);
      }
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
}

template <typename E>
LB_DEV void splat_persistent2(const E &ev, const CamConsts<float> &cam, const FilterConsts &fc, const AovSet &aovs, const SampleIO &s,
                              const WorkItem *__restrict__ work, FilterCounters *counters, uint64_t sample_base) {
  const int lane = threadIdx.x & 31;
  const unsigned n_work = *((volatile unsigned *)&counters->work_count);
  for (;;) {
    unsigned idx = 0;
    if (lane == 0) idx = atomicAdd(&counters->work_next, 1u);
    idx = __shfl_sync(0xffffffffu, idx, 0);
    if (idx >= n_work) break;
    const WorkItem w = work[idx];
    splat_work_item2(ev, cam, fc, aovs, s, w, counters, sample_base);
  }
}

}  // namespace lb
