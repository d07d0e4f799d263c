// filter_host.cu — host side of the filter / imager / multi-GPU entry points of the C ABI.
//
// Owns the device-resident replacements of Camera::aovs[i].buffer, filter_weight_buffer, zbuffer and
// zbuffer_debug (/root/reference/src/lentil.h:100-103, 1096-1117; aov_data.h:114-164).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is resolved at run time (dlopen), see load_nccl()

#include <algorithm>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lentil_b200.h"
#include "lentil_internal.h"

namespace lb {
int cam_device(lb_camera *c);
int cam_num_sms(lb_camera *c);
const lb_camera_params &cam_params(lb_camera *c);
const lb_camera_state &cam_state(lb_camera *c);
const LensTable &cam_lens(lb_camera *c);
int cam_lens_kernel(lb_camera *c);
const CamConsts<float> &cam_consts(lb_camera *c);
const ThinConsts &cam_thin(lb_camera *c);
struct FilterStateTag;
std::mutex &cam_mutex(lb_camera *c);
int lb_fail(int code, const char *msg);
}  // namespace lb
using namespace lb;

struct FilterState {
  lb_frame_desc frame{};
  int n_aov = 0;
  lb_aov_desc aovs[kMaxAov]{};
  size_t npx = 0;
  size_t npx_pad = 0;      // plane stride in pixels: npx rounded up to kSlabAlign so that every world size that divides
                           // kSlabAlign owns an equal, contiguous pixel slab of each plane (lb_filter_reduce_scatter)
  float *block = nullptr;  // [n_aov][npx_pad] float4 planes, then [npx_pad] weight: one contiguous reduction unit
  size_t block_floats = 0;
  unsigned long long *zkey = nullptr, *zkey_debug = nullptr;
  size_t zkey_npx = 0;     // pixels the two key planes are allocated for
  // Ordering across streams.  `ev_done` is recorded by every begin / reduce / resolve (operations that need the planes
  // complete or exclusive); an accumulate waits for it and for the previous user of its scratch slot, and records the slot's
  // event; begin / reduce / resolve wait for `ev_done` and for every slot (wait_all_accumulates).  So accumulates overlap
  // each other, everything else executes in call order.  Frames with closest-filter AOVs use one slot (their gather step
  // must see the batches in order).
  cudaEvent_t ev_done = nullptr;
  bool scattered = false;  // lb_filter_reduce_scatter has run: this rank's planes are complete only inside its slab
  float4 *gather = nullptr;  // [npx_pad] resolved pixels of one AOV (lb_imager_resolve_gather)
  bool has_closest = false, has_debug_closest = false;
  // Batch scratch (work list, its two heads, per-sample helper arrays).  A pool, so that accumulates issued on different
  // streams -- the chunks of lb_filter_accumulate_host, the flushes of several render threads -- run CONCURRENTLY: their
  // framebuffer updates are atomic and commute, and a persistent splat kernel whose work list is running dry leaves SM
  // slots to the next batch's kernel instead of idling them.
  struct Scratch {
    WorkItem *work = nullptr;
    size_t work_cap = 0;
    uint16_t *debug_samples = nullptr;
    float2 *crypto_cache = nullptr;   // [n_crypto][work_cap][crypto_cache_stride]
    int crypto_cache_stride = 0;
    unsigned int *heads = nullptr;    // [kWorkHeads], see AovSet::work_heads
    cudaEvent_t done = nullptr;       // recorded after the last kernel that used this scratch
    cudaStream_t last_stream = nullptr;
    bool used = false;
  };
  static constexpr int kScratch = 6;
  Scratch scratch[kScratch];
  int next_scratch = 0;
  cudaEvent_t ev_accum = nullptr;     // scratch for "wait for every accumulate in flight"
  // cryptomatte: per crypto AOV one table of [npx][crypto_slots] id bits followed by [npx][crypto_slots] weights
  int n_crypto = 0, crypto_slots = 0;
  int crypto_of[kMaxAov]{};          // AOV index -> crypto table index, -1 for the others
  int crypto_rank[kMaxAov]{};        // 0 / 2 / 4 from the AOV name (lentil_imager.cpp:124-126)
  uint32_t *crypto_tables = nullptr; // n_crypto tables back to back
  size_t crypto_table_words = 0;     // 2 * npx * crypto_slots
  FilterCounters *d_counters = nullptr;
  uint64_t sample_base = 0;
  uint64_t samples_seen = 0;  // source samples handed to accumulate since lb_filter_begin
  // host-path staging: two device blocks so that the copy of chunk k+1 overlaps the kernels of chunk k
  // kStage blocks in flight: a chunk's block is held until its splat kernel has finished (the kernel reads the AOV values from
  // it), and chunks that hold highlights run for milliseconds -- with three blocks the copies stalled behind them
  static constexpr int kStage = 6;  // at most (12 measured no faster); a call uses as many as fit 2 GiB (>= 3)
  char *stage[kStage] = {};
  size_t stage_bytes = 0;
  int stage_slots = 0;  // blocks allocated
  cudaEvent_t stage_ready[kStage] = {}, stage_free[kStage] = {};
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaStream_t work_stream[kStage] = {};  // lb_filter_accumulate_host: chunk k runs on work_stream[k % kStage]
  // lb_imager_resolve_host: the whole region of an AOV is resolved once into pinned host memory, buckets are served from it
  float4 *res_dev = nullptr;           // [npx]
  uint8_t *brk_dev = nullptr;          // [npx] cryptomatte: pixel holds <= rank ids (ends a bucket row, lentil_imager.cpp:132-134)
  float4 *res_host[kMaxAov] = {};      // pinned, [npx] each, allocated on first use
  uint8_t *brk_host[kMaxAov] = {};
  bool res_valid[kMaxAov] = {};
  size_t res_npx = 0;
  // NCCL
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  // peer-memory combine (lb_imager_resolve_peer): every rank maps the other ranks' planes, depth keys and image block through
  // CUDA IPC; alloc_gen counts (re)allocations of the exported buffers, peer_gen is the generation the mappings belong to
  static constexpr int kMaxPeers = 8;
  uint64_t alloc_gen = 1, peer_gen = 0;
  float *peer_block[kMaxPeers] = {};
  unsigned long long *peer_zkey[kMaxPeers] = {}, *peer_zkey_debug[kMaxPeers] = {};
  float4 *peer_img[kMaxPeers] = {};
  float4 *img_block = nullptr;   // [n_aov][npx_pad] resolved images, written by the owners of the slabs (this rank's own mapping)
  size_t img_floats = 0;
  float *barrier_word = nullptr; // 4-byte all-reduce = stream-ordered barrier across the ranks
  unsigned char *ipc_dev = nullptr;  // [world][kIpcBytes] handle exchange
};
namespace lb { FilterState *&cam_filter(lb_camera *c); }

namespace {

// plane stride granule in pixels: 5040 = 2^4 * 3^2 * 5 * 7, so 1..10, 12, 14, 15, 16 ranks own equal contiguous slabs
constexpr size_t kSlabAlign = 5040;

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev); else prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CUF(call)                                                                         \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) return lb_fail(LB_ERR_CUDA, cudaGetErrorString(e_));           \
  } while (0)

// ---- NCCL through dlopen: the process' already-loaded libnccl.so.2 (e.g. torch's) is reused ----
struct NcclApi {
  void *h = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclReduce) Reduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclReduceScatter) ReduceScatter = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclSend) Send = nullptr;
  decltype(&ncclRecv) Recv = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
};
NcclApi *load_nccl() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char *names[] = {getenv("LB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      if (!n) continue;
      api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.h) break;
    }
    if (!api.h) return;
#define SYM(f) api.f = (decltype(api.f))dlsym(api.h, "nccl" #f)
    SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllReduce); SYM(Reduce); SYM(Broadcast); SYM(ReduceScatter); SYM(AllGather); SYM(Send); SYM(Recv); SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString);
#undef SYM
  });
  return (api.h && api.GetUniqueId && api.CommInitRank && api.AllReduce && api.Reduce) ? &api : nullptr;
}

// Unmap every peer buffer, then wait until all ranks have done the same (the owner may only free an exported buffer after
// its importers have closed it).  Collective when mappings exist; a no-op otherwise.
int peer_release(FilterState *f) {
  if (f->peer_gen == 0) return LB_OK;
  for (int r = 0; r < FilterState::kMaxPeers; ++r) {
    if (r != f->rank) {
      if (f->peer_block[r]) cudaIpcCloseMemHandle(f->peer_block[r]);
      if (f->peer_zkey[r]) cudaIpcCloseMemHandle(f->peer_zkey[r]);
      if (f->peer_zkey_debug[r]) cudaIpcCloseMemHandle(f->peer_zkey_debug[r]);
      if (f->peer_img[r]) cudaIpcCloseMemHandle(f->peer_img[r]);
    }
    f->peer_block[r] = nullptr; f->peer_zkey[r] = f->peer_zkey_debug[r] = nullptr; f->peer_img[r] = nullptr;
  }
  cudaGetLastError();
  f->peer_gen = 0;
  NcclApi *n = load_nccl();
  if (n && f->comm && f->barrier_word && f->stream) {
    if (n->AllReduce(f->barrier_word, f->barrier_word + 1, 1, ncclFloat32, ncclSum, f->comm, f->stream) != ncclSuccess)
      return lb_fail(LB_ERR_COMM, "barrier before releasing peer-mapped buffers failed");
    CUF(cudaStreamSynchronize(f->stream));
  }
  return LB_OK;
}

void fill_consts(lb_camera *c, const FilterState *f, FilterConsts &fc) {
  const lb_camera_params &p = cam_params(c);
  const lb_camera_state &s = cam_state(c);
  fc.xres = f->frame.xres; fc.yres = f->frame.yres;
  fc.xres_full = f->frame.xres_without_region; fc.yres_full = f->frame.yres_without_region;
  fc.region_min_x = f->frame.region_min_x; fc.region_min_y = f->frame.region_min_y;
  fc.n_aov = f->n_aov;
  fc.bidir_sample_mult = p.bidir_sample_mult;
  fc.camera_type = p.camera_type;
  fc.enable_skydome = p.enable_skydome != 0;
  fc.enable_bidir_transmission = p.enable_bidir_transmission != 0;
  const float um[4] = {(float)0.1, (float)1.0, (float)10.0, (float)100.0};
  fc.unit_mult = um[std::min(std::max(p.units, 0), 3)];
  fc.focal_length = p.focal_length_lentil < (float)0.01 ? (float)0.01 : p.focal_length_lentil;
  // get_coc_thinlens prologue (lentil.h:676-687)
  float fd = (float)s.focus_distance, ar = (float)s.aperture_radius;
  if (p.camera_type == LB_CAMERA_POLYNOMIAL_OPTICS) fd = (float)((double)fd / 10.0);
  else ar = (float)((double)ar * 10.0);
  fc.coc_focus_distance = fd;
  fc.coc_aperture_radius = ar;
  fc.lens_length_tenth = s.lens_length * 0.1;
  fc.bidir_add_energy = p.bidir_add_energy;
  fc.bidir_add_energy_transition = p.bidir_add_energy_transition;
  fc.bidir_add_energy_minimum_luminance = p.bidir_add_energy_minimum_luminance;
  fc.abb_chromatic = p.abb_chromatic;
  fc.sensor_half = (double)p.sensor_width * 0.5;
  fc.aspect_full = (double)f->frame.xres_without_region / (double)f->frame.yres_without_region;
}

// The plane an AOV's per-pixel floats live in.  AOVData::crypto_total_weight (lentil.h:815) receives the same `sample_weight`
// for every cryptomatte AOV of the frame (lentil_filter.cpp:295-298 calls add_to_buffer for each of them per splat), so ONE
// plane -- the first cryptomatte AOV's -- holds it for all of them: one reduction per splat instead of one per cryptomatte AOV.
int plane_of(const FilterState *f, int aov) {
  if (f->crypto_of[aov] < 0) return aov;
  for (int a = 0; a < f->n_aov; ++a)
    if (f->crypto_of[a] >= 0) return a;
  return aov;
}

void fill_aovs(const FilterState *f, AovSet &A, const lb_samples *S, const FilterState::Scratch *sc) {
  memset(&A, 0, sizeof A);
  const float *const *values = S ? S->aov_values : nullptr;
  A.crypto_slots = f->crypto_slots;
  A.crypto_first = -1;
  for (int a = f->n_aov - 1; a >= 0; --a)
    if (f->crypto_of[a] >= 0) A.crypto_first = a;
  A.crypto_depth = S ? S->crypto_depth : 0;
  const bool add_zeros = getenv("LB_ADD_ZEROS") && getenv("LB_ADD_ZEROS")[0] == '1';
  A.add_zeros = add_zeros ? 1 : 0;
  for (int a = 0; a < f->n_aov; ++a) {
    A.buffer[a] = (float4 *)(f->block + (size_t)plane_of(f, a) * f->npx_pad * 4);
    A.values[a] = values ? (const float4 *)values[a] : nullptr;
    A.filter[a] = f->aovs[a].filter;
    A.role[a] = f->aovs[a].role;
    const int t = f->crypto_of[a];
    if (t >= 0) {
      A.values[a] = nullptr;
      A.crypto_key[a] = f->crypto_tables + (size_t)t * f->crypto_table_words;
      A.crypto_wgt[a] = (float *)(A.crypto_key[a] + f->crypto_table_words / 2);
      A.crypto_ids[a] = (S && S->crypto_depth > 0 && S->crypto_ids) ? S->crypto_ids[a] : nullptr;
      A.crypto_cache[a] = (sc && sc->crypto_cache) ? sc->crypto_cache + (size_t)t * sc->work_cap * sc->crypto_cache_stride : nullptr;
    }
  }
  A.weight = f->block + (size_t)f->n_aov * f->npx_pad * 4;
  A.zkey = f->has_closest ? f->zkey : nullptr;  // planes that are not in use are not zeroed either
  A.zkey_debug = f->has_debug_closest ? f->zkey_debug : nullptr;
  A.debug_samples = (f->has_debug_closest && sc) ? sc->debug_samples : nullptr;
  A.work_heads = sc ? sc->heads : nullptr;
}

int ensure_batch_capacity(FilterState *f, FilterState::Scratch *sc, size_t n, int crypto_depth) {
  const int stride = f->n_crypto ? std::max(crypto_depth, 1) : 0;
  if (!sc->heads) CUF(cudaMalloc(&sc->heads, kWorkHeads * sizeof(unsigned)));
  if (!sc->done) {
    CUF(cudaEventCreateWithFlags(&sc->done, cudaEventDisableTiming));
    CUF(cudaEventRecord(sc->done, f->stream));
  }
  if (sc->work_cap >= n && sc->crypto_cache_stride >= stride) return LB_OK;
  n = std::max(n, sc->work_cap);
  CUF(cudaEventSynchronize(sc->done));  // the previous user of this slot
  cudaFree(sc->work); cudaFree(sc->debug_samples); cudaFree(sc->crypto_cache);
  sc->work = nullptr; sc->debug_samples = nullptr; sc->crypto_cache = nullptr; sc->work_cap = 0; sc->crypto_cache_stride = 0;
  CUF(cudaMalloc(&sc->work, n * sizeof(WorkItem)));
  CUF(cudaMalloc(&sc->debug_samples, n * sizeof(uint16_t)));
  if (stride) CUF(cudaMalloc(&sc->crypto_cache, (size_t)f->n_crypto * n * stride * sizeof(float2)));
  sc->work_cap = n;
  sc->crypto_cache_stride = stride;
  return LB_OK;
}

// begin / reduce / resolve: after every accumulate in flight and after the previous operation of that kind
int wait_all_accumulates(FilterState *f, cudaStream_t stream) {
  CUF(cudaStreamWaitEvent(stream, f->ev_done, 0));
  for (auto &sc : f->scratch)
    if (sc.done) CUF(cudaStreamWaitEvent(stream, sc.done, 0));
  return LB_OK;
}

// classify -> splat -> (closest gather) for one device-resident batch
int accumulate_device(lb_camera *c, FilterState *f, const lb_samples *S, cudaStream_t stream) {
  if (S->n == 0) return LB_OK;
  if (S->n > 0xFFFFFFFFull) return lb_fail(LB_ERR_INVALID, "batch larger than 2^32 samples");
  if (S->crypto_depth < 0 || S->crypto_depth > LB_CRYPTO_MAX_DEPTH) return lb_fail(LB_ERR_INVALID, "crypto_depth outside [0, LB_CRYPTO_MAX_DEPTH]");
  // closest-filter AOVs fetch the winning sample's value per batch (launch_closest_gather): batches in order, one slot
  FilterState::Scratch *sc = &f->scratch[0];
  if (!(f->has_closest || f->has_debug_closest)) {
    // the first slot whose previous batch has finished (serial callers keep reusing slot 0 and its allocations); when all
    // are in flight, the next one in turn -- the new batch then queues behind that slot's batch on the device
    // a slot last used on this very stream needs no waiting at all (stream order), so a caller that issues frame after
    // frame on one stream stays on one slot
    int pick = -1;
    for (int i = 0; i < FilterState::kScratch && pick < 0; ++i)
      if (f->scratch[i].used && f->scratch[i].last_stream == stream) pick = i;
    for (int i = 0; i < FilterState::kScratch && pick < 0; ++i)
      if (!f->scratch[i].done || cudaEventQuery(f->scratch[i].done) == cudaSuccess) pick = i;
    cudaGetLastError();  // cudaErrorNotReady of the queries is not an error
    if (pick < 0) { pick = f->next_scratch; f->next_scratch = (f->next_scratch + 1) % FilterState::kScratch; }
    sc = &f->scratch[pick];
  }
  int rc = ensure_batch_capacity(f, sc, S->n, S->crypto_depth);
  if (rc != LB_OK) return rc;
  FilterConsts fc;
  fill_consts(c, f, fc);
  AovSet A;
  fill_aovs(f, A, S, sc);
  SampleIO io{S->n, S->px, S->py, (const float4 *)S->rgba, (const float4 *)S->pos_cs, (const float4 *)S->raydir,
              (const float4 *)S->transmission, S->flags, S->inv_density,
              f->n_crypto && S->crypto_depth > 0 ? S->crypto_count : nullptr, f->n_crypto && S->crypto_depth > 0 ? S->crypto_opacity : nullptr};
  // AiWorldToCameraMatrix of the batch (lentil_filter.cpp:139-142); identity when the caller hands camera-space positions
  for (int r = 0; r < 4; ++r)
    for (int k = 0; k < 3; ++k) io.w2c[r][k] = S->world_to_camera ? S->world_to_camera[4 * r + k] : (r == k ? 1.0f : 0.0f);
  CUF(cudaStreamWaitEvent(stream, f->ev_done, 0));  // the last begin / reduce / resolve
  CUF(cudaStreamWaitEvent(stream, sc->done, 0));    // the previous batch that used this scratch slot
  f->scattered = false;
  for (int a = 0; a < f->n_aov; ++a) f->res_valid[a] = false;
  CUF(cudaMemsetAsync(sc->heads, 0, kWorkHeads * sizeof(unsigned), stream));
  CUF(launch_filter_classify(fc, A, io, sc->work, f->d_counters, f->sample_base, stream));
  if (cam_params(c).camera_type == LB_CAMERA_THINLENS)
    CUF(launch_filter_splat_thinlens(cam_consts(c), cam_thin(c), fc, A, io, sc->work, f->d_counters, f->sample_base, cam_num_sms(c), stream));
  else
    CUF(launch_filter_splat(cam_lens_kernel(c), cam_lens(c), cam_consts(c), fc, A, io, sc->work, f->d_counters, f->sample_base,
                            cam_num_sms(c), stream));
  if (f->has_closest || f->has_debug_closest) CUF(launch_closest_gather(fc, A, io, f->sample_base, stream));
  CUF(cudaEventRecord(sc->done, stream));
  sc->last_stream = stream;
  sc->used = true;
  f->sample_base += S->n;
  f->samples_seen += S->n;
  return LB_OK;
}

__global__ void k_mask_closest(const unsigned long long *__restrict__ local_key, const unsigned long long *__restrict__ global_key,
                               float4 *__restrict__ buffer, size_t npx) {
  const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npx) return;
  if (local_key[p] != global_key[p]) buffer[p] = make_float4(0.f, 0.f, 0.f, 0.f);
}

}  // namespace

void filter_state_destroy(FilterState *f) {
  if (!f) return;
  peer_release(f);
  if (f->comm) { if (NcclApi *n = load_nccl()) n->CommDestroy(f->comm); }
  f->comm = nullptr;
  cudaFree(f->block); cudaFree(f->zkey); cudaFree(f->zkey_debug);
  for (auto &sc : f->scratch) {
    cudaFree(sc.work); cudaFree(sc.debug_samples); cudaFree(sc.crypto_cache); cudaFree(sc.heads);
    if (sc.done) cudaEventDestroy(sc.done);
  }
  peer_release(f);  // (lb_comm_destroy has normally done it, with the barrier)
  cudaFree(f->img_block); cudaFree(f->barrier_word); cudaFree(f->ipc_dev);
  cudaFree(f->crypto_tables);
  cudaFree(f->d_counters); cudaFree(f->gather); cudaFree(f->res_dev); cudaFree(f->brk_dev);
  for (int a = 0; a < kMaxAov; ++a) { cudaFreeHost(f->res_host[a]); cudaFreeHost(f->brk_host[a]); }
  for (int i = 0; i < FilterState::kStage; ++i) {
    cudaFree(f->stage[i]);
    if (f->stage_ready[i]) cudaEventDestroy(f->stage_ready[i]);
    if (f->stage_free[i]) cudaEventDestroy(f->stage_free[i]);
    if (f->work_stream[i]) cudaStreamDestroy(f->work_stream[i]);
  }
  if (f->ev_done) cudaEventDestroy(f->ev_done);
  if (f->stream) cudaStreamDestroy(f->stream);
  if (f->copy_stream) cudaStreamDestroy(f->copy_stream);
  delete f;
}

extern "C" {

int lb_filter_begin(lb_camera *c, const lb_frame_desc *frame, int n_aov, const lb_aov_desc *aovs) {
  if (!c || !frame || !aovs || n_aov <= 0 || n_aov > kMaxAov) return lb_fail(LB_ERR_INVALID, "bad filter_begin arguments");
  if (frame->xres <= 0 || frame->yres <= 0) return lb_fail(LB_ERR_INVALID, "empty frame");
  std::lock_guard<std::mutex> lk(cam_mutex(c));
  DeviceGuard g(cam_device(c));
  FilterState *&slot = cam_filter(c);
  FilterState *f = slot;
  if (!f) { f = new FilterState(); slot = f; }
  if (!f->stream) CUF(cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking));
  if (!f->copy_stream) CUF(cudaStreamCreateWithFlags(&f->copy_stream, cudaStreamNonBlocking));
  if (!f->ev_done) {
    CUF(cudaEventCreateWithFlags(&f->ev_done, cudaEventDisableTiming));
    CUF(cudaEventRecord(f->ev_done, f->stream));
  }
  f->frame = *frame;
  f->n_aov = n_aov;
  memcpy(f->aovs, aovs, n_aov * sizeof(lb_aov_desc));
  const size_t npx = (size_t)frame->xres * frame->yres;
  const size_t npx_pad = (npx + kSlabAlign - 1) / kSlabAlign * kSlabAlign;
  const size_t floats = npx_pad * (4 * (size_t)n_aov + 1);
  // destroy_buffers + reallocate (lentil.h:214,1096-1117).  cudaFree waits for work in flight on the old buffers.
  if (floats != f->block_floats || npx != f->zkey_npx) {  // buffers other ranks have mapped (lb_imager_resolve_peer): unmapped everywhere first
    const int rcp = peer_release(f);
    if (rcp != LB_OK) return rcp;
  }
  if (floats != f->block_floats) {
    cudaFree(f->block);
    f->block = nullptr; f->block_floats = 0;
    CUF(cudaMalloc(&f->block, floats * sizeof(float)));
    f->block_floats = floats;
    ++f->alloc_gen;
  }
  if (npx != f->zkey_npx) {  // sized by the pixel count alone: two frames can share `floats` and differ in npx
    cudaFree(f->zkey); cudaFree(f->zkey_debug); cudaFree(f->gather);
    f->zkey = f->zkey_debug = nullptr; f->gather = nullptr; f->zkey_npx = 0;
    CUF(cudaMalloc(&f->zkey, npx * sizeof(unsigned long long)));
    CUF(cudaMalloc(&f->zkey_debug, npx * sizeof(unsigned long long)));
    f->zkey_npx = npx;
    ++f->alloc_gen;
  }
  f->npx = npx;
  f->npx_pad = npx_pad;
  // cryptomatte tables (AOVData::allocate_cryptomatte_buffers, aov_data.h:145-150)
  int n_crypto = 0;
  for (int a = 0; a < n_aov; ++a) {
    f->crypto_of[a] = -1;
    f->crypto_rank[a] = 0;
    if (aovs[a].filter != LB_FILTER_CRYPTO) continue;
    char name[65] = {0};
    memcpy(name, aovs[a].name, 64);
    if (!strstr(name, "crypto_")) return lb_fail(LB_ERR_INVALID, "LB_FILTER_CRYPTO AOV whose name has no \"crypto_\" (lentil.h:1036)");
    for (const char *k : {"crypto_material", "crypto_asset", "crypto_object"}) {  // lentil_imager.cpp:124-126
      if (std::string(name) == std::string(k) + "01") f->crypto_rank[a] = 2;
      if (std::string(name) == std::string(k) + "02") f->crypto_rank[a] = 4;
    }
    f->crypto_of[a] = n_crypto++;
  }
  const int slots = frame->crypto_slots > 0 ? frame->crypto_slots : LB_CRYPTO_DEFAULT_SLOTS;
  if (n_crypto && slots > LB_CRYPTO_MAX_SLOTS) return lb_fail(LB_ERR_INVALID, "crypto_slots above LB_CRYPTO_MAX_SLOTS");
  const size_t table_words = n_crypto ? 2 * npx * (size_t)slots : 0;
  if (n_crypto != f->n_crypto || table_words != f->crypto_table_words) {
    cudaFree(f->crypto_tables);  // (waits for work in flight)
    f->crypto_tables = nullptr;
    for (auto &sc : f->scratch) {  // the per-batch caches are sized by the number of cryptomatte AOVs: reallocated on next use
      cudaFree(sc.work); cudaFree(sc.debug_samples); cudaFree(sc.crypto_cache);
      sc.work = nullptr; sc.debug_samples = nullptr; sc.crypto_cache = nullptr;
      sc.work_cap = 0;
      sc.crypto_cache_stride = 0;
    }
    f->n_crypto = 0; f->crypto_table_words = 0;
    if (n_crypto) CUF(cudaMalloc(&f->crypto_tables, (size_t)n_crypto * table_words * 4));
    f->n_crypto = n_crypto;
    f->crypto_table_words = table_words;
  }
  f->crypto_slots = n_crypto ? slots : 0;
  if (!f->d_counters) CUF(cudaMalloc(&f->d_counters, sizeof(FilterCounters)));
  // zero everything on the filter's stream, ordered after whatever still reads the previous frame; the first
  // accumulate waits for `ev_done` on its own stream
  cudaStream_t st = f->stream;
  { const int rcw = wait_all_accumulates(f, st); if (rcw != LB_OK) return rcw; }
  for (int t = 0; t < n_crypto; ++t) {
    uint32_t *tab = f->crypto_tables + (size_t)t * table_words;
    CUF(cudaMemsetAsync(tab, 0xFF, table_words / 2 * 4, st));               // ids: kCryptoFree
    CUF(cudaMemsetAsync(tab + table_words / 2, 0, table_words / 2 * 4, st)); // weights
  }
  f->has_closest = f->has_debug_closest = false;
  for (int a = 0; a < n_aov; ++a)
    if (aovs[a].filter == LB_FILTER_CLOSEST) (aovs[a].role == LB_AOV_LENTIL_DEBUG ? f->has_debug_closest : f->has_closest) = true;
  CUF(cudaMemsetAsync(f->block, 0, floats * sizeof(float), st));
  if (f->has_closest) CUF(cudaMemsetAsync(f->zkey, 0xFF, npx * sizeof(unsigned long long), st));
  if (f->has_debug_closest) CUF(cudaMemsetAsync(f->zkey_debug, 0xFF, npx * sizeof(unsigned long long), st));
  CUF(cudaMemsetAsync(f->d_counters, 0, sizeof(FilterCounters), st));
  CUF(cudaEventRecord(f->ev_done, st));
  f->sample_base = 0;
  f->samples_seen = 0;
  f->scattered = false;
  for (int a = 0; a < kMaxAov; ++a) f->res_valid[a] = false;
  return LB_OK;
}

int lb_filter_set_sample_base(lb_camera *c, uint64_t base) {
  if (!c || !cam_filter(c)) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  cam_filter(c)->sample_base = base;
  return LB_OK;
}

int lb_filter_accumulate(lb_camera *c, const lb_samples *S, lb_stream stream) {
  if (!c || !S) return lb_fail(LB_ERR_INVALID, "null argument");
  FilterState *f = cam_filter(c);
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (S->n && (!S->px || !S->py || !S->rgba || !S->pos_cs)) return lb_fail(LB_ERR_INVALID, "null sample array");
  std::lock_guard<std::mutex> lk(cam_mutex(c));
  DeviceGuard g(cam_device(c));
  return accumulate_device(c, f, S, (cudaStream_t)stream);
}

// Host-buffer variant: samples are staged to the device in chunks through THREE staging blocks.  Chunk k is copied on the
// copy stream and classified + splatted on work_stream[k % 3] with its own batch scratch, so the copy of chunk k+1 overlaps
// the kernels of chunk k AND the kernels of consecutive chunks overlap each other: a chunk holds only the highlights of its
// part of the frame, its persistent splat kernel runs dry early and the next chunk's kernel takes over the free SM slots.
// Source memory may be pageable (Arnold's sample buffers are): the driver then stages the copy itself, still beside the kernels.
int lb_filter_accumulate_host(lb_camera *c, const lb_samples *S) {
  if (!c || !S) return lb_fail(LB_ERR_INVALID, "null argument");
  FilterState *f = cam_filter(c);
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (S->n && (!S->px || !S->py || !S->rgba || !S->pos_cs)) return lb_fail(LB_ERR_INVALID, "null sample array");
  if (S->crypto_depth < 0 || S->crypto_depth > LB_CRYPTO_MAX_DEPTH) return lb_fail(LB_ERR_INVALID, "crypto_depth outside [0, LB_CRYPTO_MAX_DEPTH]");
  if (S->n == 0) return LB_OK;
  std::lock_guard<std::mutex> lk(cam_mutex(c));
  DeviceGuard g(cam_device(c));
  constexpr int kSlots = FilterState::kStage;
  const size_t chunk = std::min<size_t>(S->n, (size_t)1 << 21);
  int n_val = 0;
  for (int a = 0; a < f->n_aov; ++a) if (S->aov_values && S->aov_values[a]) ++n_val;
  const size_t Dn = f->n_crypto ? (size_t)S->crypto_depth : 0;
  int n_ids = 0;
  for (int a = 0; a < f->n_aov; ++a) if (Dn && S->crypto_ids && S->crypto_ids[a]) ++n_ids;
  // per-sample bytes: px,py (8) + rgba,pos (32) + raydir,transmission (32) + flags (4) + values + cryptomatte sub-samples
  const size_t per = 8 + 32 + (S->raydir ? 16 : 0) + (S->transmission ? 16 : 0) + (S->flags ? 4 : 0) + 16 * (size_t)n_val +
                     (Dn ? 1 + 4 * Dn * (1 + (size_t)n_ids) : 0);
  const size_t need = per * chunk + 256 * (16 + 2 * kMaxAov);
  const int n_slots = (int)std::min<size_t>((size_t)kSlots, std::max<size_t>(3, ((size_t)2 << 30) / need));
  if (f->stage_bytes < need || f->stage_slots < n_slots) {
    for (int i = 0; i < kSlots; ++i) {
      cudaFree(f->stage[i]); f->stage[i] = nullptr;
      if (!f->stage_ready[i]) CUF(cudaEventCreateWithFlags(&f->stage_ready[i], cudaEventDisableTiming));
      if (!f->stage_free[i]) CUF(cudaEventCreateWithFlags(&f->stage_free[i], cudaEventDisableTiming));
      if (!f->work_stream[i]) CUF(cudaStreamCreateWithFlags(&f->work_stream[i], cudaStreamNonBlocking));
    }
    f->stage_bytes = 0;
    f->stage_slots = 0;
    for (int i = 0; i < n_slots; ++i) CUF(cudaMalloc(&f->stage[i], need));
    f->stage_bytes = need;
    f->stage_slots = n_slots;
  }
  int rc = LB_OK;
  size_t k = 0;
  for (size_t base = 0; base < S->n && rc == LB_OK; base += chunk, ++k) {
    const int slot = (int)(k % n_slots);
    const size_t m = std::min(chunk, S->n - base);
    if (k >= (size_t)n_slots) CUF(cudaStreamWaitEvent(f->copy_stream, f->stage_free[slot], 0));  // the kernels of chunk k - kStage have read this block
    char *p = f->stage[slot];
    auto put = [&](const void *src, size_t elem) -> void * {
      if (!src) return nullptr;
      void *d = p;
      cudaMemcpyAsync(d, (const char *)src + base * elem, m * elem, cudaMemcpyHostToDevice, f->copy_stream);
      p += (chunk * elem + 255) & ~(size_t)255;
      return d;
    };
    lb_samples D = *S;
    D.n = m;
    D.rgba = (const float *)put(S->rgba, 16);
    D.pos_cs = (const float *)put(S->pos_cs, 16);
    D.raydir = (const float *)put(S->raydir, 16);
    D.transmission = (const float *)put(S->transmission, 16);
    const float *vals[kMaxAov] = {nullptr};
    for (int a = 0; a < f->n_aov; ++a) vals[a] = (S->aov_values && S->aov_values[a]) ? (const float *)put(S->aov_values[a], 16) : nullptr;
    D.aov_values = vals;
    D.px = (const int32_t *)put(S->px, 4);
    D.py = (const int32_t *)put(S->py, 4);
    D.flags = (const uint32_t *)put(S->flags, 4);
    const float *cids[kMaxAov] = {nullptr};
    if (Dn) {
      D.crypto_opacity = (const float *)put(S->crypto_opacity, 4 * Dn);
      for (int a = 0; a < f->n_aov; ++a) cids[a] = (S->crypto_ids && S->crypto_ids[a]) ? (const float *)put(S->crypto_ids[a], 4 * Dn) : nullptr;
      D.crypto_ids = cids;
      D.crypto_count = (const uint8_t *)put(S->crypto_count, 1);
    }
    CUF(cudaGetLastError());
    cudaStream_t ws = f->work_stream[slot];
    CUF(cudaEventRecord(f->stage_ready[slot], f->copy_stream));
    CUF(cudaStreamWaitEvent(ws, f->stage_ready[slot], 0));
    rc = accumulate_device(c, f, &D, ws);
    CUF(cudaEventRecord(f->stage_free[slot], ws));
  }
  for (int i = 0; i < kSlots; ++i) CUF(cudaStreamSynchronize(f->work_stream[i]));  // the caller may reuse its arrays; the result is complete
  return rc;
}

int lb_filter_get_stats(lb_camera *c, lb_filter_stats *out) {
  if (!c || !out) return lb_fail(LB_ERR_INVALID, "null argument");
  FilterState *f = cam_filter(c);
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  DeviceGuard g(cam_device(c));
  CUF(cudaDeviceSynchronize());
  FilterCounters h;
  CUF(cudaMemcpy(&h, f->d_counters, sizeof h, cudaMemcpyDeviceToHost));
  out->samples = f->samples_seen; out->redistributed = h.redistributed; out->splats = h.splats;
  out->attempts = h.attempts; out->passthrough = f->samples_seen - h.redistributed;
  out->crypto_dropped = h.crypto_dropped;
  out->tile_splats = h.tile_splats;
  return LB_OK;
}

int lb_filter_newton_iterations(lb_camera *c, uint64_t *out) {  // diagnostic: lt_sample_aperture iterations so far
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f || !out) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  DeviceGuard g(cam_device(c));
  CUF(cudaDeviceSynchronize());
  FilterCounters h;
  CUF(cudaMemcpy(&h, f->d_counters, sizeof h, cudaMemcpyDeviceToHost));
  *out = h.newton_its;
  return LB_OK;
}

int lb_imager_resolve(lb_camera *c, int aov, int x0, int y0, int w, int h, float *rgba_out, lb_stream stream_) {
  if (!c || !rgba_out) return lb_fail(LB_ERR_INVALID, "null argument");
  FilterState *f = cam_filter(c);
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (aov < 0 || aov >= f->n_aov) return lb_fail(LB_ERR_INVALID, "aov index out of range");
  const int rx = x0 - f->frame.region_min_x, ry = y0 - f->frame.region_min_y;
  if (rx < 0 || ry < 0 || w < 0 || h < 0 || rx + w > f->frame.xres || ry + h > f->frame.yres) return lb_fail(LB_ERR_INVALID, "bucket outside the region");
  if (f->scattered) return lb_fail(LB_ERR_STATE, "after lb_filter_reduce_scatter a rank holds only its slab: use lb_imager_resolve_gather");
  DeviceGuard g(cam_device(c));
  cudaStream_t stream = (cudaStream_t)stream_;
  { const int rcw = wait_all_accumulates(f, stream); if (rcw != LB_OK) return rcw; }
  if (f->crypto_of[aov] >= 0) {
    const uint32_t *key = f->crypto_tables + (size_t)f->crypto_of[aov] * f->crypto_table_words;
    CUF(launch_resolve_crypto(key, (const float *)(key + f->crypto_table_words / 2), (const float4 *)(f->block + (size_t)plane_of(f, aov) * f->npx_pad * 4),
                              f->crypto_slots, f->crypto_rank[aov], f->frame.xres, rx, ry, w, h, (float4 *)rgba_out, nullptr, stream));
  } else {
    CUF(launch_resolve((const float4 *)(f->block + (size_t)aov * f->npx_pad * 4), f->block + (size_t)f->n_aov * f->npx_pad * 4, f->aovs[aov].filter,
                       f->aovs[aov].role, f->frame.xres, rx, ry, w, h, (float4 *)rgba_out, stream));
  }
  CUF(cudaEventRecord(f->ev_done, stream));
  return LB_OK;
}

// driver_process_bucket with a HOST bucket.  Arnold calls it once per bucket and output (8 100 buckets of 16 x 16 pixels
// at 1080p): the first call after the frame changed resolves the WHOLE region of that AOV on the device into pinned
// host memory (one kernel, one copy), every call then copies its bucket rows out of that image on the host.
// Cryptomatte rows end at the first pixel of the BUCKET row holding <= rank ids (lentil_imager.cpp:132-134), which
// depends on the bucket: the device marks those pixels and the row copy stops at the first mark.
int lb_imager_resolve_host(lb_camera *c, int aov, int x0, int y0, int w, int h, float *rgba_out) {
  if (!c || !rgba_out) return lb_fail(LB_ERR_INVALID, "null argument");
  FilterState *f = cam_filter(c);
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (aov < 0 || aov >= f->n_aov) return lb_fail(LB_ERR_INVALID, "aov index out of range");
  const int rx = x0 - f->frame.region_min_x, ry = y0 - f->frame.region_min_y;
  if (rx < 0 || ry < 0 || w < 0 || h < 0 || rx + w > f->frame.xres || ry + h > f->frame.yres) return lb_fail(LB_ERR_INVALID, "bucket outside the region");
  if (w == 0 || h == 0) return LB_OK;
  if (f->scattered) return lb_fail(LB_ERR_STATE, "after lb_filter_reduce_scatter a rank holds only its slab: use lb_imager_resolve_gather");
  std::lock_guard<std::mutex> lk(cam_mutex(c));
  const bool crypto = f->crypto_of[aov] >= 0;
  if (!f->res_valid[aov] || f->res_npx != f->npx) {
    DeviceGuard g(cam_device(c));
    if (f->res_npx != f->npx) {
      cudaFree(f->res_dev); cudaFree(f->brk_dev);
      f->res_dev = nullptr; f->brk_dev = nullptr;
      for (int a = 0; a < kMaxAov; ++a) {
        cudaFreeHost(f->res_host[a]); cudaFreeHost(f->brk_host[a]);
        f->res_host[a] = nullptr; f->brk_host[a] = nullptr; f->res_valid[a] = false;
      }
      f->res_npx = 0;
      CUF(cudaMalloc(&f->res_dev, f->npx * sizeof(float4)));
      CUF(cudaMalloc(&f->brk_dev, f->npx));
      f->res_npx = f->npx;
    }
    if (!f->res_host[aov]) CUF(cudaMallocHost(&f->res_host[aov], f->npx * sizeof(float4)));
    if (crypto && !f->brk_host[aov]) CUF(cudaMallocHost(&f->brk_host[aov], f->npx));
    cudaStream_t st = f->stream;
    { const int rcw = wait_all_accumulates(f, st); if (rcw != LB_OK) return rcw; }
    const int xr = f->frame.xres, yr = f->frame.yres;
    if (crypto) {
      const uint32_t *key = f->crypto_tables + (size_t)f->crypto_of[aov] * f->crypto_table_words;
      CUF(cudaMemsetAsync(f->res_dev, 0, f->npx * sizeof(float4), st));
      CUF(launch_resolve_crypto(key, (const float *)(key + f->crypto_table_words / 2), (const float4 *)(f->block + (size_t)plane_of(f, aov) * f->npx_pad * 4),
                                f->crypto_slots, f->crypto_rank[aov], xr, 0, 0, xr, yr, f->res_dev, f->brk_dev, st));
      CUF(cudaMemcpyAsync(f->brk_host[aov], f->brk_dev, f->npx, cudaMemcpyDeviceToHost, st));
    } else {
      CUF(launch_resolve((const float4 *)(f->block + (size_t)aov * f->npx_pad * 4), f->block + (size_t)f->n_aov * f->npx_pad * 4, f->aovs[aov].filter,
                         f->aovs[aov].role, xr, 0, 0, xr, yr, f->res_dev, st));
    }
    CUF(cudaMemcpyAsync(f->res_host[aov], f->res_dev, f->npx * sizeof(float4), cudaMemcpyDeviceToHost, st));
    CUF(cudaEventRecord(f->ev_done, st));
    CUF(cudaStreamSynchronize(st));
    f->res_valid[aov] = true;
  }
  const float4 *img = f->res_host[aov];
  const int xr = f->frame.xres;
  for (int j = 0; j < h; ++j) {
    const size_t row = (size_t)(ry + j) * xr + rx;
    int n = w;
    if (crypto) {
      const uint8_t *b = f->brk_host[aov] + row;
      for (n = 0; n < w && !b[n]; ++n) {}
    }
    memcpy(rgba_out + (size_t)j * w * 4, img + row, (size_t)n * sizeof(float4));
  }
  return LB_OK;
}

int lb_filter_buffers(lb_camera *c, int aov, float **buffer, float **weight) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (aov < 0 || aov >= f->n_aov) return lb_fail(LB_ERR_INVALID, "aov index out of range");
  if (buffer) *buffer = f->block + (size_t)plane_of(f, aov) * f->npx_pad * 4;
  if (weight) *weight = f->block + (size_t)f->n_aov * f->npx_pad * 4;
  return LB_OK;
}

int lb_filter_buffers_host(lb_camera *c, int aov, float *buffer_out, float *weight_out) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (aov < 0 || aov >= f->n_aov) return lb_fail(LB_ERR_INVALID, "aov index out of range");
  DeviceGuard g(cam_device(c));
  CUF(cudaDeviceSynchronize());
  if (buffer_out) CUF(cudaMemcpy(buffer_out, f->block + (size_t)plane_of(f, aov) * f->npx_pad * 4, f->npx * 16, cudaMemcpyDeviceToHost));
  if (weight_out) CUF(cudaMemcpy(weight_out, f->block + (size_t)f->n_aov * f->npx_pad * 4, f->npx * 4, cudaMemcpyDeviceToHost));
  return LB_OK;
}

int lb_filter_crypto_host(lb_camera *c, int aov, float *ids_out, float *weights_out, int *slots_out) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (aov < 0 || aov >= f->n_aov || f->crypto_of[aov] < 0) return lb_fail(LB_ERR_INVALID, "not a cryptomatte AOV");
  DeviceGuard g(cam_device(c));
  CUF(cudaDeviceSynchronize());
  const uint32_t *key = f->crypto_tables + (size_t)f->crypto_of[aov] * f->crypto_table_words;
  const size_t half = f->crypto_table_words / 2;
  if (ids_out) CUF(cudaMemcpy(ids_out, key, half * 4, cudaMemcpyDeviceToHost));
  if (weights_out) CUF(cudaMemcpy(weights_out, key + half, half * 4, cudaMemcpyDeviceToHost));
  if (slots_out) *slots_out = f->crypto_slots;
  return LB_OK;
}

// ---- multi-GPU ------------------------------------------------------------------------------------
int lb_comm_unique_id(uint8_t id_out[128]) {
  NcclApi *n = load_nccl();
  if (!n) return lb_fail(LB_ERR_COMM, "libnccl.so.2 not found (set LB_NCCL_LIB)");
  ncclUniqueId id;
  if (n->GetUniqueId(&id) != ncclSuccess) return lb_fail(LB_ERR_COMM, "ncclGetUniqueId failed");
  memcpy(id_out, id.internal, 128);
  return LB_OK;
}

int lb_comm_init(lb_camera *c, int world_size, int rank, const uint8_t id_in[128]) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  NcclApi *n = load_nccl();
  if (!n) return lb_fail(LB_ERR_COMM, "libnccl.so.2 not found (set LB_NCCL_LIB)");
  DeviceGuard g(cam_device(c));
  if (f->comm) { peer_release(f); n->CommDestroy(f->comm); f->comm = nullptr; }
  ncclUniqueId id;
  memcpy(id.internal, id_in, 128);
  ncclResult_t r = n->CommInitRank(&f->comm, world_size, id, rank);
  if (r != ncclSuccess) return lb_fail(LB_ERR_COMM, n->GetErrorString ? n->GetErrorString(r) : "ncclCommInitRank failed");
  f->world = world_size;
  f->rank = rank;
  return LB_OK;
}

}  // extern "C"

namespace {
// closest AOVs and cryptomatte tables: the parts of the combine that are not an element-wise sum.
// Closest: global min of the depth keys, then every rank but the winner clears its pixel (so the sum keeps the winner's
// value).  Cryptomatte: every rank's id tables are broadcast in turn and folded into the receivers' own (ids differ per
// rank, so this is a merge, not an element-wise reduction); crypto_total_weight rides in the AOV plane and is summed.
int reduce_keys_and_tables(FilterState *f, NcclApi *n, int root, cudaStream_t stream) {
  auto check = [&](ncclResult_t r) { return r == ncclSuccess ? LB_OK : lb_fail(LB_ERR_COMM, n->GetErrorString ? n->GetErrorString(r) : "nccl error"); };
  int rc;
  for (int pass = 0; pass < 2; ++pass) {
    const bool on = pass == 0 ? f->has_closest : f->has_debug_closest;
    if (!on) continue;
    unsigned long long *local = pass == 0 ? f->zkey : f->zkey_debug;
    unsigned long long *global = nullptr;
    CUF(cudaMallocAsync(&global, f->npx * sizeof(unsigned long long), stream));
    if ((rc = check(n->AllReduce(local, global, f->npx, ncclUint64, ncclMin, f->comm, stream))) != LB_OK) return rc;
    for (int a = 0; a < f->n_aov; ++a) {
      const bool dbg = f->aovs[a].role == LB_AOV_LENTIL_DEBUG;
      if (f->aovs[a].filter != LB_FILTER_CLOSEST || dbg != (pass == 1)) continue;
      k_mask_closest<<<(unsigned)((f->npx + 255) / 256), 256, 0, stream>>>(local, global, (float4 *)(f->block + (size_t)a * f->npx_pad * 4), f->npx);
    }
    CUF(cudaMemcpyAsync(local, global, f->npx * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
    CUF(cudaFreeAsync(global, stream));
  }
  if (f->n_crypto) {
    if (!n->Broadcast) return lb_fail(LB_ERR_COMM, "ncclBroadcast unavailable");
    const size_t words = (size_t)f->n_crypto * f->crypto_table_words;
    uint32_t *mine = nullptr, *other = nullptr;
    CUF(cudaMallocAsync(&mine, words * 4, stream));   // snapshot: what this rank contributes
    CUF(cudaMallocAsync(&other, words * 4, stream));
    CUF(cudaMemcpyAsync(mine, f->crypto_tables, words * 4, cudaMemcpyDeviceToDevice, stream));
    for (int r = 0; r < f->world; ++r) {
      if ((rc = check(n->Broadcast(mine, other, words, ncclUint32, r, f->comm, stream))) != LB_OK) return rc;
      if (r == f->rank || (root >= 0 && f->rank != root)) continue;
      for (int t = 0; t < f->n_crypto; ++t) {
        uint32_t *key = f->crypto_tables + (size_t)t * f->crypto_table_words;
        const uint32_t *okey = other + (size_t)t * f->crypto_table_words;
        const size_t half = f->crypto_table_words / 2;
        CUF(launch_crypto_merge(key, (float *)(key + half), okey, (const float *)(okey + half), f->npx, f->crypto_slots, f->d_counters, stream));
      }
    }
    CUF(cudaFreeAsync(mine, stream));
    CUF(cudaFreeAsync(other, stream));
  }
  return LB_OK;
}
}  // namespace

extern "C" {

int lb_filter_reduce(lb_camera *c, int root, lb_stream stream_) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (f->world <= 1 || !f->comm) return LB_OK;  // single rank: nothing to combine
  NcclApi *n = load_nccl();
  if (!n) return lb_fail(LB_ERR_COMM, "NCCL unavailable");
  std::lock_guard<std::mutex> lk(cam_mutex(c));
  DeviceGuard g(cam_device(c));
  cudaStream_t stream = (cudaStream_t)stream_;
  auto check = [&](ncclResult_t r) { return r == ncclSuccess ? LB_OK : lb_fail(LB_ERR_COMM, n->GetErrorString ? n->GetErrorString(r) : "nccl error"); };
  { const int rcw = wait_all_accumulates(f, stream); if (rcw != LB_OK) return rcw; }
  for (int a = 0; a < f->n_aov; ++a) f->res_valid[a] = false;
  int rc = reduce_keys_and_tables(f, n, root, stream);
  if (rc != LB_OK) return rc;
  // one sum-reduce over every AOV plane + the weight plane
  if (root < 0) rc = check(n->AllReduce(f->block, f->block, f->block_floats, ncclFloat32, ncclSum, f->comm, stream));
  else rc = check(n->Reduce(f->block, f->block, f->block_floats, ncclFloat32, ncclSum, root, f->comm, stream));
  CUF(cudaEventRecord(f->ev_done, stream));
  return rc;
}

// Sum-reduce where every rank ends up OWNING one pixel slab of every plane: per plane one in-place
// ncclReduceScatter, all planes in one NCCL group.  Each rank then resolves its own slab and the resolved pixels are
// gathered (lb_imager_resolve_gather): 1/world of the resolve work per rank and no root bottleneck in the reduction.
int lb_filter_reduce_scatter(lb_camera *c, lb_stream stream_) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (f->world <= 1 || !f->comm) return LB_OK;
  NcclApi *n = load_nccl();
  if (!n || !n->ReduceScatter) return lb_fail(LB_ERR_COMM, "NCCL unavailable");
  if (f->npx_pad % (size_t)f->world) return lb_fail(LB_ERR_INVALID, "world size does not divide the plane granule (5040): use lb_filter_reduce");
  if (f->n_crypto) return lb_fail(LB_ERR_INVALID, "frames with cryptomatte AOVs combine with lb_filter_reduce (the ranked resolve reads whole bucket rows)");
  std::lock_guard<std::mutex> lk(cam_mutex(c));
  DeviceGuard g(cam_device(c));
  cudaStream_t stream = (cudaStream_t)stream_;
  auto check = [&](ncclResult_t r) { return r == ncclSuccess ? LB_OK : lb_fail(LB_ERR_COMM, n->GetErrorString ? n->GetErrorString(r) : "nccl error"); };
  { const int rcw = wait_all_accumulates(f, stream); if (rcw != LB_OK) return rcw; }
  for (int a = 0; a < f->n_aov; ++a) f->res_valid[a] = false;
  int rc = reduce_keys_and_tables(f, n, -1, stream);
  if (rc != LB_OK) return rc;
  const size_t slab = f->npx_pad / (size_t)f->world;  // pixels per rank
  if ((rc = check(n->GroupStart())) != LB_OK) return rc;
  for (int a = 0; a <= f->n_aov && rc == LB_OK; ++a) {
    const size_t per_px = a < f->n_aov ? 4 : 1;  // float4 planes, then the float weight plane
    float *plane = f->block + (size_t)a * f->npx_pad * 4;
    rc = check(n->ReduceScatter(plane, plane + (size_t)f->rank * slab * per_px, slab * per_px, ncclFloat32, ncclSum, f->comm, stream));
  }
  const int rc2 = check(n->GroupEnd());
  if (rc == LB_OK) rc = rc2;
  f->scattered = rc == LB_OK;
  CUF(cudaEventRecord(f->ev_done, stream));
  return rc;
}

int lb_filter_slab(lb_camera *c, size_t *first_pixel, size_t *n_pixels) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  const size_t slab = f->npx_pad / (size_t)std::max(f->world, 1);
  const size_t lo = std::min((size_t)f->rank * slab, f->npx), hi = std::min(lo + slab, f->npx);
  if (first_pixel) *first_pixel = lo;
  if (n_pixels) *n_pixels = hi - lo;
  return LB_OK;
}

// driver_process_bucket for the whole region after lb_filter_reduce_scatter: this rank resolves its slab, the slabs are
// gathered on `root` (ncclSend / ncclRecv) or on every rank (root < 0, ncclAllGather).  rgba_out: [yres][xres][4] device
// floats (only written where the result lands).
int lb_imager_resolve_gather(lb_camera *c, int aov, float *rgba_out, int root, lb_stream stream_) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (aov < 0 || aov >= f->n_aov) return lb_fail(LB_ERR_INVALID, "aov index out of range");
  if (root >= f->world) return lb_fail(LB_ERR_INVALID, "root out of range");
  std::lock_guard<std::mutex> lk(cam_mutex(c));
  DeviceGuard g(cam_device(c));
  cudaStream_t stream = (cudaStream_t)stream_;
  const bool single = f->world <= 1 || !f->comm;
  if (!single && !f->scattered) return lb_fail(LB_ERR_STATE, "lb_filter_reduce_scatter has not been called for this frame");
  const bool mine = single || root < 0 || root == f->rank;
  if (mine && !rgba_out) return lb_fail(LB_ERR_INVALID, "null output on a receiving rank");
  { const int rcw = wait_all_accumulates(f, stream); if (rcw != LB_OK) return rcw; }
  const float4 *plane = (const float4 *)(f->block + (size_t)aov * f->npx_pad * 4);
  const float *wplane = f->block + (size_t)f->n_aov * f->npx_pad * 4;
  if (single) {
    CUF(launch_resolve_linear(plane, wplane, f->aovs[aov].filter, f->aovs[aov].role, 0, f->npx, (float4 *)rgba_out, stream));
    CUF(cudaEventRecord(f->ev_done, stream));
    return LB_OK;
  }
  NcclApi *n = load_nccl();
  if (!n || !n->AllGather || !n->Send || !n->Recv) return lb_fail(LB_ERR_COMM, "NCCL unavailable");
  auto check = [&](ncclResult_t r) { return r == ncclSuccess ? LB_OK : lb_fail(LB_ERR_COMM, n->GetErrorString ? n->GetErrorString(r) : "nccl error"); };
  if (!f->gather) CUF(cudaMalloc(&f->gather, f->npx_pad * sizeof(float4)));
  const size_t slab = f->npx_pad / (size_t)f->world;
  const size_t lo = (size_t)f->rank * slab;
  const size_t cnt = lo < f->npx ? std::min(slab, f->npx - lo) : 0;
  CUF(launch_resolve_linear(plane, wplane, f->aovs[aov].filter, f->aovs[aov].role, lo, cnt, f->gather + lo, stream));
  int rc = LB_OK;
  if (root < 0) {
    rc = check(n->AllGather(f->gather + lo, f->gather, slab * 4, ncclFloat32, f->comm, stream));
  } else {
    if ((rc = check(n->GroupStart())) != LB_OK) return rc;
    if (f->rank == root) {
      for (int r = 0; r < f->world && rc == LB_OK; ++r)
        if (r != root) rc = check(n->Recv(f->gather + (size_t)r * slab, slab * 4, ncclFloat32, r, f->comm, stream));
    } else {
      rc = check(n->Send(f->gather + lo, slab * 4, ncclFloat32, root, f->comm, stream));
    }
    const int rc2 = check(n->GroupEnd());
    if (rc == LB_OK) rc = rc2;
  }
  if (rc == LB_OK && mine) CUF(cudaMemcpyAsync(rgba_out, f->gather, f->npx * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
  CUF(cudaEventRecord(f->ev_done, stream));
  return rc;
}

int lb_comm_destroy(lb_camera *c) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f || !f->comm) return LB_OK;
  peer_release(f);
  if (NcclApi *n = load_nccl()) n->CommDestroy(f->comm);
  f->comm = nullptr;
  f->world = 1;
  return LB_OK;
}

}  // extern "C"

// ---- multi-GPU combine + resolve in ONE kernel over peer memory ------------------------------------------------------
namespace {
struct PeerSet {
  const float *block[FilterState::kMaxPeers];                    // every rank's [n_aov][npx_pad] float4 planes + [npx_pad] weight plane
  const unsigned long long *zkey[FilterState::kMaxPeers];        // every rank's depth keys (closest-filter AOVs)
  const unsigned long long *zkey_debug[FilterState::kMaxPeers];
  float4 *img[FilterState::kMaxPeers];                           // every rank's image block [n_aov][npx_pad]
  int world, root, n_aov, n_out;
  int out_aov[kMaxAov], filter[kMaxAov], role[kMaxAov];
  size_t npx_pad;
};

// driver_process_bucket (lentil_imager.cpp:112-189) for the pixels [lo, lo + cnt) this rank owns, reading the partial
// framebuffers of ALL ranks through NVLink: gaussian AOVs are summed in rank order and divided by the summed filter weight,
// closest-filter AOVs take the value of the rank whose depth key is smallest (what the sequential z-test keeps,
// lentil.h:832-846).  The resolved pixel goes straight into the image block of `root` (of every rank when root < 0).
// Loads are 16 B and independent across ranks and AOVs (world x (n_out + 1) in flight per thread); stores are coalesced
// 512 B per warp.  Bound: the NVLink ingress of the slab owner, (world - 1) / world of the planes.
__global__ void __launch_bounds__(256) k_resolve_peer(const __grid_constant__ PeerSet P, size_t lo, size_t cnt) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = lo + i;
    float fw = 0.f;
    for (int r = 0; r < P.world; ++r) fw += __ldcs(P.block[r] + (size_t)P.n_aov * P.npx_pad * 4 + p);
    for (int k = 0; k < P.n_out; ++k) {
      const int a = P.out_aov[k];
      const size_t at = (size_t)a * P.npx_pad + p;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (P.filter[a] == 0) {
        for (int r = 0; r < P.world; ++r) {
          const float4 t = __ldcs(reinterpret_cast<const float4 *>(P.block[r]) + at);
          v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
        }
        if (P.role[a] != 2 && fw != 0.0f) { const float inv = 1.0f / fw; v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv; }
      } else {
        unsigned long long best = ~0ull;
        int who = 0;
        for (int r = 0; r < P.world; ++r) {
          const unsigned long long key = __ldcs((P.role[a] == 2 ? P.zkey_debug[r] : P.zkey[r]) + p);
          if (key < best) { best = key; who = r; }
        }
        if (best != ~0ull) v = __ldcs(reinterpret_cast<const float4 *>(P.block[who]) + at);
        v.w = 1.0f;
      }
      if (P.root >= 0) P.img[P.root][at] = v;
      else
        for (int r = 0; r < P.world; ++r) P.img[r][at] = v;
    }
  }
  __threadfence_system();
}

constexpr size_t kIpcBytes = 4 * sizeof(cudaIpcMemHandle_t) + 8;

// (re)map the peers' buffers: handles travel through a device buffer and ncclAllGather
int peer_exchange(FilterState *f, NcclApi *n, cudaStream_t stream) {
  auto check = [&](ncclResult_t r) { return r == ncclSuccess ? LB_OK : lb_fail(LB_ERR_COMM, n->GetErrorString ? n->GetErrorString(r) : "nccl error"); };
  if (!f->barrier_word) { CUF(cudaMalloc(&f->barrier_word, 8)); CUF(cudaMemset(f->barrier_word, 0, 8)); }
  const size_t want_img = (size_t)f->n_aov * f->npx_pad * 4;
  if (f->img_floats != want_img) {  // (the ranks agree on the frame, so they all take this branch together)
    const int rcp = peer_release(f);
    if (rcp != LB_OK) return rcp;
    cudaFree(f->img_block);
    f->img_block = nullptr; f->img_floats = 0;
    CUF(cudaMalloc(&f->img_block, want_img * sizeof(float)));
    f->img_floats = want_img;
    ++f->alloc_gen;
  }
  if (f->peer_gen == f->alloc_gen) return LB_OK;
  { const int rcp = peer_release(f); if (rcp != LB_OK) return rcp; }
  if (!f->ipc_dev) CUF(cudaMalloc(&f->ipc_dev, FilterState::kMaxPeers * kIpcBytes));
  unsigned char mine[kIpcBytes];
  memset(mine, 0, sizeof mine);
  void *bufs[4] = {f->block, f->zkey, f->zkey_debug, f->img_block};
  for (int k = 0; k < 4; ++k) {
    cudaIpcMemHandle_t h;
    CUF(cudaIpcGetMemHandle(&h, bufs[k]));
    memcpy(mine + k * sizeof h, &h, sizeof h);
  }
  CUF(cudaMemcpyAsync(f->ipc_dev + (size_t)f->rank * kIpcBytes, mine, kIpcBytes, cudaMemcpyHostToDevice, stream));
  int rc = check(n->AllGather(f->ipc_dev + (size_t)f->rank * kIpcBytes, f->ipc_dev, kIpcBytes, ncclInt8, f->comm, stream));
  if (rc != LB_OK) return rc;
  std::vector<unsigned char> all((size_t)f->world * kIpcBytes);
  CUF(cudaMemcpyAsync(all.data(), f->ipc_dev, all.size(), cudaMemcpyDeviceToHost, stream));
  CUF(cudaStreamSynchronize(stream));
  bool opened = true;
  for (int r = 0; r < f->world; ++r) {
    if (r == f->rank) {
      f->peer_block[r] = f->block; f->peer_zkey[r] = f->zkey; f->peer_zkey_debug[r] = f->zkey_debug; f->peer_img[r] = f->img_block;
      continue;
    }
    void *ptr[4] = {nullptr, nullptr, nullptr, nullptr};
    for (int k = 0; k < 4; ++k) {
      cudaIpcMemHandle_t h;
      memcpy(&h, all.data() + (size_t)r * kIpcBytes + k * sizeof h, sizeof h);
      if (cudaIpcOpenMemHandle(&ptr[k], h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ptr[k] = nullptr; opened = false; cudaGetLastError(); }
    }
    f->peer_block[r] = (float *)ptr[0]; f->peer_zkey[r] = (unsigned long long *)ptr[1];
    f->peer_zkey_debug[r] = (unsigned long long *)ptr[2]; f->peer_img[r] = (float4 *)ptr[3];
  }
  // the outcome is collective: a rank that could not map a peer must not leave the others waiting in the barrier below
  float flag[2] = {opened ? 0.0f : 1.0f, 0.0f};
  CUF(cudaMemcpyAsync(f->barrier_word, flag, 4, cudaMemcpyHostToDevice, stream));
  if ((rc = check(n->AllReduce(f->barrier_word, f->barrier_word + 1, 1, ncclFloat32, ncclSum, f->comm, stream))) != LB_OK) return rc;
  CUF(cudaMemcpyAsync(flag + 1, f->barrier_word + 1, 4, cudaMemcpyDeviceToHost, stream));
  CUF(cudaMemsetAsync(f->barrier_word, 0, 4, stream));
  CUF(cudaStreamSynchronize(stream));
  f->peer_gen = f->alloc_gen;  // (mappings exist: peer_release below / later closes them)
  if (flag[1] != 0.0f) {
    peer_release(f);
    return lb_fail(LB_ERR_COMM, "cudaIpcOpenMemHandle failed on some rank (no peer access between the devices?): use lb_filter_reduce_scatter");
  }
  return LB_OK;
}
}  // namespace

extern "C" {

int lb_imager_resolve_peer(lb_camera *c, int n_out, const int *aov_indices, float *const *images_out, int root, lb_stream stream_) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f) return lb_fail(LB_ERR_STATE, "lb_filter_begin has not been called");
  if (n_out <= 0 || n_out > f->n_aov || !aov_indices) return lb_fail(LB_ERR_INVALID, "bad AOV list");
  if (root >= f->world) return lb_fail(LB_ERR_INVALID, "root out of range");
  for (int k = 0; k < n_out; ++k) {
    if (aov_indices[k] < 0 || aov_indices[k] >= f->n_aov) return lb_fail(LB_ERR_INVALID, "aov index out of range");
    if (f->crypto_of[aov_indices[k]] >= 0) return lb_fail(LB_ERR_INVALID, "cryptomatte AOVs combine with lb_filter_reduce (id tables are merged, not summed)");
  }
  if (f->world > FilterState::kMaxPeers) return lb_fail(LB_ERR_INVALID, "more than 8 ranks: use lb_filter_reduce_scatter");
  if (f->scattered) return lb_fail(LB_ERR_STATE, "lb_filter_reduce_scatter has already consumed this frame's partial planes");
  std::lock_guard<std::mutex> lk(cam_mutex(c));
  DeviceGuard g(cam_device(c));
  cudaStream_t stream = (cudaStream_t)stream_;
  { const int rcw = wait_all_accumulates(f, stream); if (rcw != LB_OK) return rcw; }
  const bool single = f->world <= 1 || !f->comm;
  NcclApi *n = single ? nullptr : load_nccl();
  if (!single && (!n || !n->AllGather)) return lb_fail(LB_ERR_COMM, "NCCL unavailable");
  auto check = [&](ncclResult_t r) { return r == ncclSuccess ? LB_OK : lb_fail(LB_ERR_COMM, n->GetErrorString ? n->GetErrorString(r) : "nccl error"); };
  int rc = LB_OK;
  if (single) {
    const size_t want_img = (size_t)f->n_aov * f->npx_pad * 4;
    if (f->img_floats != want_img) {
      cudaFree(f->img_block);
      f->img_block = nullptr; f->img_floats = 0;
      CUF(cudaMalloc(&f->img_block, want_img * sizeof(float)));
      f->img_floats = want_img;
    }
  } else if ((rc = peer_exchange(f, n, stream)) != LB_OK) {
    return rc;
  }
  PeerSet P;
  memset(&P, 0, sizeof P);
  P.world = single ? 1 : f->world;
  P.root = single ? 0 : root;
  P.n_aov = f->n_aov;
  P.n_out = n_out;
  P.npx_pad = f->npx_pad;
  for (int k = 0; k < n_out; ++k) P.out_aov[k] = aov_indices[k];
  for (int a = 0; a < f->n_aov; ++a) { P.filter[a] = f->aovs[a].filter; P.role[a] = f->aovs[a].role; }
  for (int r = 0; r < P.world; ++r) {
    P.block[r] = single ? f->block : f->peer_block[r];
    P.zkey[r] = single ? f->zkey : f->peer_zkey[r];
    P.zkey_debug[r] = single ? f->zkey_debug : f->peer_zkey_debug[r];
    P.img[r] = single ? f->img_block : f->peer_img[r];
  }
  // barrier: every rank's accumulates have finished (and nobody still reads the previous frame's images) before any rank loads
  if (!single && (rc = check(n->AllReduce(f->barrier_word, f->barrier_word + 1, 1, ncclFloat32, ncclSum, f->comm, stream))) != LB_OK) return rc;
  // Who owns which pixels.  Every owner pulls its slab from all other ranks, and the receiving rank also takes in every owner's
  // resolved pixels -- with even slabs its NVLink ingress is (N-1)/N of (planes + images) against (N-1)/N of the planes for
  // the others (config C5, N = 8: 9.4 GB vs 4.8 GB, and the step waits for it).  So when one rank receives the image and
  // there are three or more ranks, that rank owns no slab: it only receives (5.3 GB), the others pull slabs of 1/(N-1).
  const int owners = (!single && root >= 0 && P.world >= 3) ? P.world - 1 : P.world;
  const int owner_idx = owners == P.world ? f->rank : (f->rank == root ? -1 : (f->rank < root ? f->rank : f->rank - 1));
  const size_t slab = ((f->npx + (size_t)owners - 1) / (size_t)owners + 255) / 256 * 256;
  const size_t lo = single || owner_idx < 0 ? 0 : (size_t)owner_idx * slab;
  const size_t cnt = owner_idx < 0 ? 0 : (lo < f->npx ? std::min(slab, f->npx - lo) : 0);
  if (cnt) {
    const unsigned grid = (unsigned)std::min<size_t>((cnt + 255) / 256, (size_t)cam_num_sms(c) * 16);
    k_resolve_peer<<<grid, 256, 0, stream>>>(P, lo, cnt);
    CUF(cudaGetLastError());
  }
  // barrier: every owner has stored its slab before the images are read, and before the next lb_filter_begin zeroes the planes
  if (!single && (rc = check(n->AllReduce(f->barrier_word, f->barrier_word + 1, 1, ncclFloat32, ncclSum, f->comm, stream))) != LB_OK) return rc;
  const bool mine = single || root < 0 || root == f->rank;
  if (mine && images_out)
    for (int k = 0; k < n_out; ++k)
      if (images_out[k])
        CUF(cudaMemcpyAsync(images_out[k], f->img_block + (size_t)aov_indices[k] * f->npx_pad, f->npx * sizeof(float4), cudaMemcpyDeviceToDevice, stream));
  CUF(cudaEventRecord(f->ev_done, stream));
  return LB_OK;
}

int lb_imager_peer_image(lb_camera *c, int aov, float **image) {
  FilterState *f = c ? cam_filter(c) : nullptr;
  if (!f || !f->img_block) return lb_fail(LB_ERR_STATE, "lb_imager_resolve_peer has not been called");
  if (aov < 0 || aov >= f->n_aov || !image) return lb_fail(LB_ERR_INVALID, "bad arguments");
  *image = reinterpret_cast<float *>(f->img_block + (size_t)aov * f->npx_pad);
  return LB_OK;
}

}  // extern "C"
