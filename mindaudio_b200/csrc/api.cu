// api.cu -- C ABI of libmafe.so (include/mafe.h): context, memory plumbing, plans, ragged batch
// layout and the dispatch of the hot path to the specialised (fbank512.cu) or generic
// (generic.cu) kernels.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <thread>
#include <vector>

#include "common.cuh"

namespace mafe {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};

}  // namespace mafe

using namespace mafe;

extern "C" {

int mafe_version(void) { return MAFE_VERSION; }
const char* mafe_last_error(void) { return g_err; }

// ---------------------------------------------------------------- context
int mafe_ctx_create(int device, mafe_ctx** out) {
  MAFE_REQUIRE(out != nullptr, "mafe_ctx_create: out is NULL");
  int n = 0;
  MAFE_CUDA_CHECK(cudaGetDeviceCount(&n));
  MAFE_REQUIRE(device >= 0 && device < n, "mafe_ctx_create: device %d out of range (have %d)", device, n);
  DeviceGuard g(device);
  cudaDeviceProp prop;
  MAFE_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("mafe: device %d is sm_%d%d; libmafe is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return MAFE_E_UNSUPPORTED;
  }
  mafe_ctx* c = new (std::nothrow) mafe_ctx();
  if (!c) { set_error("out of host memory"); return MAFE_E_OOM; }
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; set_error("cudaStreamCreate: %s", cudaGetErrorString(e)); return MAFE_E_CUDA; }
  c->stream = c->own_stream;
  *out = c;
  return MAFE_OK;
}

int mafe_ctx_destroy(mafe_ctx* ctx) {
  if (!ctx) return MAFE_OK;
  DeviceGuard g(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& l : ctx->lanes) {
    if (l.stream) { cudaStreamSynchronize(l.stream); cudaStreamDestroy(l.stream); }
    cudaFree(l.wave_dev);
    cudaFree(l.out_dev);
    mafe_batch_destroy(l.batch);
  }
  if (ctx->stage_up) cudaFreeHost(ctx->stage_up);
  if (ctx->stage_down) cudaFreeHost(ctx->stage_down);
  if (ctx->stage_up_done) cudaEventDestroy(ctx->stage_up_done);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return MAFE_OK;
}

int mafe_ctx_set_stream(mafe_ctx* ctx, void* cuda_stream) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  return MAFE_OK;
}

int mafe_ctx_sync(mafe_ctx* ctx) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  MAFE_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  return MAFE_OK;
}

int mafe_ctx_sm_count(const mafe_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
int64_t mafe_ctx_launch_count(const mafe_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mafe_ctx_profile_enable(mafe_ctx* ctx, int32_t enable) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  ctx->profile = enable != 0;
  return MAFE_OK;
}

static int prof_drain(mafe_ctx* ctx) {
  DeviceGuard g(ctx->device);
  MAFE_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  for (auto& p : ctx->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
      ctx->prof_ms[p.which] += ms;
      ctx->prof_n[p.which] += 1;
    }
    cudaEventDestroy(p.e0);
    cudaEventDestroy(p.e1);
  }
  ctx->prof_pending.clear();
  return MAFE_OK;
}

int mafe_ctx_profile_read(mafe_ctx* ctx, int32_t which, double* ms_out, int64_t* launches_out) {
  MAFE_REQUIRE(ctx && which >= 0 && which < MAFE_PROF_COUNT, "mafe_ctx_profile_read: bad argument");
  int rc = prof_drain(ctx);
  if (rc) return rc;
  if (ms_out) *ms_out = ctx->prof_ms[which];
  if (launches_out) *launches_out = ctx->prof_n[which];
  return MAFE_OK;
}

int mafe_ctx_profile_reset(mafe_ctx* ctx) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  int rc = prof_drain(ctx);
  if (rc) return rc;
  for (int i = 0; i < MAFE_PROF_COUNT; ++i) { ctx->prof_ms[i] = 0; ctx->prof_n[i] = 0; }
  return MAFE_OK;
}

// ---------------------------------------------------------------- memory
int mafe_device_malloc(mafe_ctx* ctx, size_t bytes, void** out_dev) {
  MAFE_REQUIRE(ctx && out_dev, "mafe_device_malloc: NULL argument");
  DeviceGuard g(ctx->device);
  MAFE_CUDA_CHECK(cudaMalloc(out_dev, std::max<size_t>(bytes, 16)));
  return MAFE_OK;
}
int mafe_device_free(mafe_ctx* ctx, void* dev) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  MAFE_CUDA_CHECK(cudaFree(dev));
  return MAFE_OK;
}
int mafe_pinned_malloc(mafe_ctx* ctx, size_t bytes, void** out_host) {
  MAFE_REQUIRE(ctx && out_host, "mafe_pinned_malloc: NULL argument");
  DeviceGuard g(ctx->device);
  MAFE_CUDA_CHECK(cudaHostAlloc(out_host, std::max<size_t>(bytes, 16), cudaHostAllocDefault));
  return MAFE_OK;
}
int mafe_pinned_free(mafe_ctx* ctx, void* host) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  MAFE_CUDA_CHECK(cudaFreeHost(host));
  return MAFE_OK;
}
int mafe_memcpy_h2d(mafe_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  if (bytes) MAFE_CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return MAFE_OK;
}
int mafe_memcpy_d2h(mafe_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  if (bytes) MAFE_CUDA_CHECK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return MAFE_OK;
}
// parallel host memcpy of a list of (dst, src, bytes) pieces: threads take contiguous byte ranges of the concatenation
static void parallel_copy(unsigned char* dst, const void* const* srcs, const int64_t* bytes, int32_t n, bool scatter_to_dst) {
  // scatter_to_dst == true: srcs[i] -> dst + prefix[i] (gather into one buffer)
  std::vector<int64_t> pre((size_t)n + 1, 0);
  for (int32_t i = 0; i < n; ++i) pre[i + 1] = pre[i] + (bytes[i] > 0 ? bytes[i] : 0);
  const int64_t total = pre[n];
  if (total == 0) return;
  int hw = (int)std::thread::hardware_concurrency();
  int nt = (int)std::min<int64_t>(std::max(1, std::min(8, hw / 2)), std::max<int64_t>(1, total / (4 << 20)));
  auto work = [&](int t) {
    const int64_t lo = total * t / nt, hi = total * (t + 1) / nt;
    int32_t i = (int32_t)(std::upper_bound(pre.begin(), pre.end(), lo) - pre.begin()) - 1;
    for (int64_t pos = lo; pos < hi; ++i) {
      const int64_t a = std::max(pos, pre[i]), b = std::min(hi, pre[i + 1]);
      if (b > a) memcpy(dst + a, (const unsigned char*)srcs[i] + (a - pre[i]), (size_t)(b - a));
      pos = pre[i + 1];
    }
  };
  (void)scatter_to_dst;
  if (nt == 1) { work(0); return; }
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
}

int mafe_memcpy_h2d_gather(mafe_ctx* ctx, void* dst_dev, const void* const* src_host, const int64_t* bytes, int32_t n) {
  MAFE_REQUIRE(ctx != nullptr && n >= 0 && (n == 0 || (src_host && bytes)), "mafe_memcpy_h2d_gather: bad argument");
  DeviceGuard g(ctx->device);
  int64_t total = 0;
  for (int32_t i = 0; i < n; ++i) {
    MAFE_REQUIRE(bytes[i] >= 0 && (bytes[i] == 0 || src_host[i]), "mafe_memcpy_h2d_gather: piece %d", i);
    total += bytes[i];
  }
  if (total == 0) return MAFE_OK;
  MAFE_REQUIRE(dst_dev != nullptr, "mafe_memcpy_h2d_gather: dst_dev is NULL");
  if (ctx->stage_up_done) MAFE_CUDA_CHECK(cudaEventSynchronize(ctx->stage_up_done));   // the previous upload has read the buffer
  if ((size_t)total > ctx->cap_stage_up) {
    if (ctx->stage_up) cudaFreeHost(ctx->stage_up);
    ctx->stage_up = nullptr; ctx->cap_stage_up = 0;
    const size_t cap = (size_t)total + (size_t)total / 4;
    MAFE_CUDA_CHECK(cudaHostAlloc(&ctx->stage_up, cap, cudaHostAllocDefault));
    ctx->cap_stage_up = cap;
  }
  if (!ctx->stage_up_done) MAFE_CUDA_CHECK(cudaEventCreateWithFlags(&ctx->stage_up_done, cudaEventDisableTiming));
  parallel_copy((unsigned char*)ctx->stage_up, src_host, bytes, n, true);
  MAFE_CUDA_CHECK(cudaMemcpyAsync(dst_dev, ctx->stage_up, (size_t)total, cudaMemcpyHostToDevice, ctx->stream));
  MAFE_CUDA_CHECK(cudaEventRecord(ctx->stage_up_done, ctx->stream));
  return MAFE_OK;
}

int mafe_memcpy_d2h_staged(mafe_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  if (bytes == 0) return MAFE_OK;
  MAFE_REQUIRE(dst_host && src_dev, "mafe_memcpy_d2h_staged: NULL argument");
  DeviceGuard g(ctx->device);
  if (bytes > ctx->cap_stage_down) {
    if (ctx->stage_down) cudaFreeHost(ctx->stage_down);
    ctx->stage_down = nullptr; ctx->cap_stage_down = 0;
    const size_t cap = bytes + bytes / 4;
    MAFE_CUDA_CHECK(cudaHostAlloc(&ctx->stage_down, cap, cudaHostAllocDefault));
    ctx->cap_stage_down = cap;
  }
  MAFE_CUDA_CHECK(cudaMemcpyAsync(ctx->stage_down, src_dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  MAFE_CUDA_CHECK(cudaStreamSynchronize(ctx->stream));
  const void* src = ctx->stage_down;
  const int64_t b = (int64_t)bytes;
  parallel_copy((unsigned char*)dst_host, &src, &b, 1, false);
  return MAFE_OK;
}

int mafe_memset(mafe_ctx* ctx, void* dst_dev, int value, size_t bytes) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  DeviceGuard g(ctx->device);
  if (bytes) MAFE_CUDA_CHECK(cudaMemsetAsync(dst_dev, value, bytes, ctx->stream));
  return MAFE_OK;
}

// ---------------------------------------------------------------- plans
int mafe_plan_create(mafe_ctx* ctx, const mafe_frontend_desc* d, mafe_plan** out) {
  MAFE_REQUIRE(ctx && d && out, "mafe_plan_create: NULL argument");
  MAFE_REQUIRE(d->n_fft >= 2 && d->n_fft <= 8192, "n_fft=%d unsupported (2..8192)", d->n_fft);
  MAFE_REQUIRE(d->frame_len >= 1 && d->frame_len <= d->n_fft, "frame_len=%d must be in [1, n_fft=%d]", d->frame_len,
               d->n_fft);
  MAFE_REQUIRE(d->hop >= 1, "Invalid hop_length: %d", d->hop);
  MAFE_REQUIRE(d->window != nullptr, "window table is NULL");
  MAFE_REQUIRE(d->pad_mode >= MAFE_PAD_CONSTANT && d->pad_mode <= MAFE_PAD_SYMMETRIC, "bad pad_mode %d", d->pad_mode);
  MAFE_REQUIRE(d->out_kind >= MAFE_OUT_COMPLEX && d->out_kind <= MAFE_OUT_MFCC, "bad out_kind %d", d->out_kind);
  MAFE_REQUIRE(!(d->center && d->frame_len != d->n_fft), "center=1 requires frame_len == n_fft");
  if (d->out_kind >= MAFE_OUT_MEL) {
    MAFE_REQUIRE(d->n_mels >= 1 && d->n_mels <= 4096 && d->mel_fb != nullptr, "mel filterbank missing / n_mels=%d", d->n_mels);
    MAFE_REQUIRE(d->log_kind >= MAFE_LOG_NONE && d->log_kind <= MAFE_LOG_DB, "bad log_kind %d", d->log_kind);
  }
  if (d->out_kind == MAFE_OUT_MFCC)
    MAFE_REQUIRE(d->n_mfcc >= 1 && d->n_mfcc <= d->n_mels && d->dct != nullptr,
                 "The number of MFCC coefficients must be no more than # mel bins.");
  DeviceGuard g(ctx->device);
  mafe_plan* p = new (std::nothrow) mafe_plan();
  if (!p) { set_error("out of host memory"); return MAFE_E_OOM; }
  p->d = *d;
  p->device = ctx->device;
  p->n_bins = d->n_fft / 2 + 1;
  switch (d->out_kind) {
    case MAFE_OUT_COMPLEX: p->out_dim = 2 * p->n_bins; break;
    case MAFE_OUT_POWER: p->out_dim = p->n_bins; break;
    case MAFE_OUT_MEL:
    case MAFE_OUT_LOGMEL: p->out_dim = d->n_mels; break;
    default: p->out_dim = d->n_mfcc; break;
  }
  int rc = generic_plan_init(ctx, p, d);
  if (rc == MAFE_OK && d->allow_fast_path && fast_plan_supported(d)) {
    rc = fast_plan_init(ctx, p, d);
    if (rc == MAFE_OK) {
      p->fast = true;
      p->tile_frames = fast_tile_frames();   // a multiple of the generic kernel's tile: it sub-tiles via gridDim.y
    }
  }
  // host table pointers are not retained
  p->d.window = nullptr; p->d.mel_fb = nullptr; p->d.dct = nullptr;
  if (rc != MAFE_OK) { mafe_plan_destroy(p); return rc; }
  *out = p;
  return MAFE_OK;
}

int mafe_plan_destroy(mafe_plan* p) {
  if (!p) return MAFE_OK;
  DeviceGuard g(p->device);
  generic_plan_free(p);
  if (p->fast_tables) fast_plan_free(p);
  delete p;
  return MAFE_OK;
}

int64_t mafe_plan_num_frames(const mafe_plan* p, int64_t n) {
  if (!p || n <= 0) return 0;
  const mafe_frontend_desc& d = p->d;
  if (d.center) {
    // spectrum.stft additionally requires n >= n_fft (spectrum.py:182-187): enforced by the python layer
    int64_t eff = n + 2 * (int64_t)(d.n_fft / 2);
    return eff < d.n_fft ? 0 : (eff - d.n_fft) / d.hop + 1;
  }
  if (n < d.frame_len) return 0;
  return (n - d.frame_len) / d.hop + 1;
}

int32_t mafe_plan_out_dim(const mafe_plan* p) { return p ? p->out_dim : 0; }
int32_t mafe_plan_is_fast(const mafe_plan* p) { return p && p->fast ? 1 : 0; }

// ---------------------------------------------------------------- ragged batch layout
}  // extern "C"

// grow-only device array
template <typename T>
static cudaError_t ensure_cap(T** p, size_t* cap, size_t need) {
  if (need <= *cap && *p) return cudaSuccess;
  if (*p) cudaFree(*p);
  *p = nullptr;
  size_t n = std::max<size_t>(need + need / 4, 16);
  cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
  *cap = e == cudaSuccess ? n : 0;
  return e;
}

extern "C" {

// (Re)fill a batch object: host layout (frame offsets, tile table) + upload on ctx->stream.  Offsets are
// rebased by `rebase` (the chunked host path uploads one chunk of the flat array at a time).  The device
// arrays only grow, so a batch can be refilled per chunk without any cudaMalloc/cudaFree.
static int batch_fill(mafe_ctx* ctx, const mafe_plan* plan, mafe_batch* b, const int64_t* so, int32_t n_utts,
                      const int32_t* utt_group, int64_t rebase) {
  b->n_utts = n_utts;
  b->frame_offsets_host.assign((size_t)n_utts + 1, 0);
  b->so_host.resize((size_t)n_utts + 1);
  b->tiles_host.clear();
  const int tf = plan->tile_frames;
  int32_t max_group = -1;
  int64_t max_frames = 0;
  b->so_host[0] = n_utts ? so[0] - rebase : 0;
  for (int32_t u = 0; u < n_utts; ++u) {
    int64_t len = so[u + 1] - so[u];
    MAFE_REQUIRE(len >= 0, "sample_offsets must be non-decreasing (utt %d)", u);
    b->so_host[u + 1] = so[u + 1] - rebase;
    int64_t t = mafe_plan_num_frames(plan, len);
    MAFE_REQUIRE(t <= INT32_MAX - tf, "utterance %d has too many frames", u);
    b->frame_offsets_host[u + 1] = b->frame_offsets_host[u] + t;
    max_frames = std::max(max_frames, t);
    for (int64_t f0 = 0; f0 < t; f0 += tf) b->tiles_host.push_back(Tile{u, (int32_t)f0});
    if (utt_group) max_group = std::max(max_group, utt_group[u]);
  }
  b->total_frames = b->frame_offsets_host[n_utts];
  b->max_utt_frames = max_frames;
  b->total_samples = n_utts ? so[n_utts] - so[0] : 0;
  b->wave_len = n_utts ? so[n_utts] - rebase : 0;
  b->n_tiles = (int32_t)b->tiles_host.size();
  b->n_groups = utt_group ? max_group + 1 : std::max(n_utts, 1);
  cudaStream_t st = ctx->stream;
  const size_t no = (size_t)n_utts + 1;
  MAFE_CUDA_CHECK(ensure_cap(&b->sample_offsets_dev, &b->cap_offsets, no));
  MAFE_CUDA_CHECK(ensure_cap(&b->frame_offsets_dev, &b->cap_foffsets, no));
  MAFE_CUDA_CHECK(ensure_cap(&b->tiles_dev, &b->cap_tiles, b->tiles_host.size()));
  MAFE_CUDA_CHECK(ensure_cap(&b->utt_sum_dev, &b->cap_utt_sum, (size_t)std::max(n_utts, 1)));
  MAFE_CUDA_CHECK(ensure_cap(&b->group_max_dev, &b->cap_groups, (size_t)std::max(b->n_groups, 1)));
  if (!b->queue_dev) MAFE_CUDA_CHECK(cudaMalloc((void**)&b->queue_dev, 64));
  MAFE_CUDA_CHECK(cudaMemcpyAsync(b->sample_offsets_dev, b->so_host.data(), no * 8, cudaMemcpyHostToDevice, st));
  MAFE_CUDA_CHECK(cudaMemcpyAsync(b->frame_offsets_dev, b->frame_offsets_host.data(), no * 8, cudaMemcpyHostToDevice, st));
  if (!b->tiles_host.empty())
    MAFE_CUDA_CHECK(cudaMemcpyAsync(b->tiles_dev, b->tiles_host.data(), b->tiles_host.size() * sizeof(Tile), cudaMemcpyHostToDevice, st));
  if (utt_group && n_utts) {
    MAFE_CUDA_CHECK(ensure_cap(&b->utt_group_dev, &b->cap_utt_group, (size_t)n_utts));
    MAFE_CUDA_CHECK(cudaMemcpyAsync(b->utt_group_dev, utt_group, (size_t)n_utts * 4, cudaMemcpyHostToDevice, st));
  } else if (!utt_group && b->utt_group_dev) {
    cudaFree(b->utt_group_dev); b->utt_group_dev = nullptr; b->cap_utt_group = 0;
  }
  if ((plan->d.utt_cmvn_mean || plan->d.utt_cmvn_std || plan->d.utt_scalar_norm) && n_utts > 0)
    MAFE_CUDA_CHECK(ensure_cap(&b->utt_stats_dev, &b->cap_utt_stats, (size_t)n_utts * 2 * plan->out_dim));
  if (plan->d.out_kind == MAFE_OUT_MFCC && b->total_frames > 0) {
    MAFE_CUDA_CHECK(ensure_cap(&b->scratch_dev, &b->cap_scratch, (size_t)b->total_frames * plan->d.n_mels));
    b->scratch_bytes = b->cap_scratch * sizeof(float);
  }
  // pageable sources: cudaMemcpyAsync has staged them before returning, the host vectors may be reused
  return MAFE_OK;
}

int mafe_batch_create(mafe_ctx* ctx, const mafe_plan* plan, const int64_t* so, int32_t n_utts, const int32_t* utt_group,
                      mafe_batch** out) {
  MAFE_REQUIRE(ctx && plan && out && (so || n_utts == 0), "mafe_batch_create: NULL argument");
  MAFE_REQUIRE(n_utts >= 0, "n_utts=%d", n_utts);
  DeviceGuard g(ctx->device);
  mafe_batch* b = new (std::nothrow) mafe_batch();
  if (!b) { set_error("out of host memory"); return MAFE_E_OOM; }
  b->device = ctx->device;
  // offsets are NOT rebased: wave_dev must point at sample 0 of the flat array
  int rc = batch_fill(ctx, plan, b, so, n_utts, utt_group, 0);
  if (rc != MAFE_OK) { mafe_batch_destroy(b); return rc; }
  *out = b;
  return MAFE_OK;
}

int mafe_batch_refill(mafe_ctx* ctx, const mafe_plan* plan, mafe_batch* b, const int64_t* so, int32_t n_utts,
                      const int32_t* utt_group) {
  MAFE_REQUIRE(ctx && plan && b && (so || n_utts == 0), "mafe_batch_refill: NULL argument");
  MAFE_REQUIRE(n_utts >= 0, "n_utts=%d", n_utts);
  MAFE_REQUIRE(b->device == ctx->device, "mafe_batch_refill: the batch belongs to device %d", b->device);
  DeviceGuard g(ctx->device);
  return batch_fill(ctx, plan, b, so, n_utts, utt_group, 0);
}

int mafe_batch_destroy(mafe_batch* b) {
  if (!b) return MAFE_OK;
  DeviceGuard g(b->device);
  cudaFree(b->sample_offsets_dev);
  cudaFree(b->frame_offsets_dev);
  cudaFree(b->tiles_dev);
  cudaFree(b->utt_group_dev);
  cudaFree(b->utt_sum_dev);
  cudaFree(b->group_max_dev);
  cudaFree(b->scratch_dev);
  cudaFree(b->utt_stats_dev);
  cudaFree(b->queue_dev);
  cudaFree(b->tile_recs_dev);
  cudaFree(b->utt_done_dev);
  delete b;
  return MAFE_OK;
}

int64_t mafe_batch_total_frames(const mafe_batch* b) { return b ? b->total_frames : 0; }
int64_t mafe_batch_total_samples(const mafe_batch* b) { return b ? b->total_samples : 0; }
int mafe_batch_frame_offsets(const mafe_batch* b, int64_t* out) {
  MAFE_REQUIRE(b && out, "mafe_batch_frame_offsets: NULL argument");
  memcpy(out, b->frame_offsets_host.data(), b->frame_offsets_host.size() * sizeof(int64_t));
  return MAFE_OK;
}
const int64_t* mafe_batch_frame_offsets_dev(const mafe_batch* b) { return b ? b->frame_offsets_dev : nullptr; }

// ---------------------------------------------------------------- the hot path
int mafe_frontend_run(mafe_ctx* ctx, const mafe_plan* plan, mafe_batch* batch, const void* wave_dev, int32_t wave_dtype,
                      float wave_scale, float* out_dev, int32_t db_group) {
  MAFE_REQUIRE(ctx && plan && batch, "mafe_frontend_run: NULL handle");
  MAFE_REQUIRE(ctx->device == plan->device && ctx->device == batch->device, "ctx/plan/batch are on different devices");
  MAFE_REQUIRE(wave_dtype == MAFE_WAVE_F32 || wave_dtype == MAFE_WAVE_I16, "bad wave_dtype %d", wave_dtype);
  MAFE_REQUIRE(db_group >= MAFE_DBGROUP_NONE && db_group <= MAFE_DBGROUP_MAP, "bad db_group %d", db_group);
  MAFE_REQUIRE(db_group != MAFE_DBGROUP_MAP || batch->utt_group_dev, "MAFE_DBGROUP_MAP needs a utt->group map");
  if (batch->total_frames == 0) return MAFE_OK;
  MAFE_REQUIRE(wave_dev && out_dev, "mafe_frontend_run: NULL buffer");
  DeviceGuard g(ctx->device);
  const mafe_frontend_desc& d = plan->d;
  const bool cmvn = d.utt_cmvn_mean || d.utt_cmvn_std;
  MAFE_REQUIRE(!((cmvn || d.utt_scalar_norm) && d.out_kind == MAFE_OUT_COMPLEX), "utterance normalisation needs a real-valued output kind");
  MAFE_REQUIRE(!(cmvn && d.utt_scalar_norm), "utt_scalar_norm excludes utt_cmvn_mean / utt_cmvn_std");
  int rc = MAFE_E_UNSUPPORTED;
  float* const feat_target = d.out_kind == MAFE_OUT_MFCC ? batch->scratch_dev : out_dev;
  if (plan->fast) {
    rc = fast_run(ctx, plan, batch, wave_dev, wave_dtype, wave_scale, feat_target, db_group);
    if (rc == MAFE_OK) return rc;                                // conformer fbank (CMVN fused inside) / STFT: done
    if (rc != kFastNeedsPost && rc != MAFE_E_UNSUPPORTED) return rc;
  }
  if (rc == MAFE_E_UNSUPPORTED) {   // generic route (also for int16 / unaligned input to a specialised kernel)
    rc = generic_run(ctx, plan, batch, wave_dev, wave_dtype, wave_scale, feat_target,
                     d.out_kind == MAFE_OUT_MFCC ? MAFE_OUT_LOGMEL : -1, db_group);
    if (rc) return rc;
  }
  if (d.out_kind == MAFE_OUT_MFCC) rc = dct_run(ctx, plan, batch, batch->scratch_dev, out_dev, db_group);
  else if (d.out_kind == MAFE_OUT_LOGMEL && d.log_kind == MAFE_LOG_DB) rc = db_clamp_run(ctx, plan, batch, out_dev, plan->out_dim, db_group);
  else rc = MAFE_OK;
  if (rc) return rc;
  if (cmvn)
    return mafe_cmvn_utt(ctx, out_dev, batch->frame_offsets_dev, batch->n_utts, plan->out_dim, d.utt_cmvn_mean, d.utt_cmvn_std);
  if (d.utt_scalar_norm)   // (the specialised transform that accumulates the moments itself returned MAFE_OK above)
    return mafe_cmvn_scalar(ctx, out_dev, batch->frame_offsets_dev, batch->n_utts, plan->out_dim, 0);
  return MAFE_OK;
}

int mafe_frontend_run_aux(mafe_ctx* ctx, const mafe_plan* plan, mafe_batch* batch, const void* wave_dev, int32_t wave_dtype,
                          float wave_scale, float* out_dev, int32_t db_group, float* aux_mel_dev) {
  if (!aux_mel_dev) return mafe_frontend_run(ctx, plan, batch, wave_dev, wave_dtype, wave_scale, out_dev, db_group);
  MAFE_REQUIRE(ctx && plan && batch, "mafe_frontend_run_aux: NULL handle");
  MAFE_REQUIRE(plan->d.out_kind == MAFE_OUT_LOGMEL || plan->d.out_kind == MAFE_OUT_MFCC,
               "mafe_frontend_run_aux: the plan's out_kind must be MAFE_OUT_LOGMEL or MAFE_OUT_MFCC");
  if (!plan->fast || !fast_has_aux_mel(plan) || wave_dtype != MAFE_WAVE_F32 || wave_scale != 1.0f || ((uintptr_t)wave_dev & 15) != 0) {
    set_error("mafe_frontend_run_aux: this plan / input runs on a kernel without the second output");
    return MAFE_E_UNSUPPORTED;
  }
  ctx->aux_mel = aux_mel_dev;
  const int rc = mafe_frontend_run(ctx, plan, batch, wave_dev, wave_dtype, wave_scale, out_dev, db_group);
  ctx->aux_mel = nullptr;
  return rc;
}

// ---------------------------------------------------------------- host-to-host hot path
int mafe_frontend_run_host(mafe_ctx* ctx, const mafe_plan* plan, const int64_t* so, int32_t n_utts, const void* wave_host,
                           int32_t wave_dtype, float wave_scale, float* out_host, int64_t* frame_offsets_out,
                           int32_t chunk_utts, int32_t db_group) {
  MAFE_REQUIRE(ctx && plan && (so || n_utts == 0), "mafe_frontend_run_host: NULL argument");
  MAFE_REQUIRE(wave_dtype == MAFE_WAVE_F32 || wave_dtype == MAFE_WAVE_I16, "bad wave_dtype %d", wave_dtype);
  MAFE_REQUIRE(db_group == MAFE_DBGROUP_NONE || db_group == MAFE_DBGROUP_UTT,
               "the chunked host path supports per-utterance dB groups only (a batch-wide floor couples chunks)");
  DeviceGuard g(ctx->device);
  const size_t es = wave_dtype == MAFE_WAVE_I16 ? 2 : 4;
  if (chunk_utts <= 0) chunk_utts = 512;
  if (ctx->lanes.empty()) {
    ctx->lanes.resize(3);
    for (auto& l : ctx->lanes) MAFE_CUDA_CHECK(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
  }
  std::vector<int64_t> fo((size_t)n_utts + 1, 0);
  for (int32_t u = 0; u < n_utts; ++u) fo[u + 1] = fo[u] + mafe_plan_num_frames(plan, so[u + 1] - so[u]);
  if (frame_offsets_out) memcpy(frame_offsets_out, fo.data(), fo.size() * sizeof(int64_t));
  cudaStream_t saved = ctx->stream;
  int rc = MAFE_OK;
  int c = 0;
  for (int32_t u0 = 0; u0 < n_utts && rc == MAFE_OK; u0 += chunk_utts, ++c) {
    const int32_t u1 = std::min(n_utts, u0 + chunk_utts);
    mafe_lane& l = ctx->lanes[c % ctx->lanes.size()];
    ctx->stream = l.stream;
    const int64_t s_begin = so[u0], n_samp = so[u1] - so[u0];
    const int64_t n_frames = fo[u1] - fo[u0];
    if (n_frames == 0) continue;
    // stream order protects the lane's buffers: the previous chunk of this lane has been enqueued before
    if (!l.batch) {
      l.batch = new (std::nothrow) mafe_batch();
      if (!l.batch) { set_error("out of host memory"); rc = MAFE_E_OOM; break; }
      l.batch->device = ctx->device;
    }
    cudaError_t e;
    if ((size_t)n_samp * es + 64 > l.wave_cap) {
      cudaStreamSynchronize(l.stream);
      cudaFree(l.wave_dev);
      l.wave_cap = ((size_t)n_samp * es + 64) * 5 / 4;
      if ((e = cudaMalloc(&l.wave_dev, l.wave_cap)) != cudaSuccess) { set_error("cudaMalloc: %s", cudaGetErrorString(e)); rc = MAFE_E_OOM; break; }
    }
    const size_t out_bytes = (size_t)n_frames * plan->out_dim * sizeof(float);
    if (out_bytes > l.out_cap) {
      cudaStreamSynchronize(l.stream);
      cudaFree(l.out_dev);
      l.out_cap = out_bytes * 5 / 4;
      if ((e = cudaMalloc((void**)&l.out_dev, l.out_cap)) != cudaSuccess) { set_error("cudaMalloc: %s", cudaGetErrorString(e)); rc = MAFE_E_OOM; break; }
    }
    if ((rc = batch_fill(ctx, plan, l.batch, so + u0, u1 - u0, nullptr, s_begin)) != MAFE_OK) break;
    e = cudaMemcpyAsync(l.wave_dev, (const char*)wave_host + (size_t)s_begin * es, (size_t)n_samp * es, cudaMemcpyHostToDevice, l.stream);
    if (e != cudaSuccess) { set_error("H2D: %s", cudaGetErrorString(e)); rc = MAFE_E_CUDA; break; }
    if ((rc = mafe_frontend_run(ctx, plan, l.batch, l.wave_dev, wave_dtype, wave_scale, l.out_dev, db_group)) != MAFE_OK) break;
    e = cudaMemcpyAsync(out_host + (size_t)fo[u0] * plan->out_dim, l.out_dev, out_bytes, cudaMemcpyDeviceToHost, l.stream);
    if (e != cudaSuccess) { set_error("D2H: %s", cudaGetErrorString(e)); rc = MAFE_E_CUDA; break; }
  }
  ctx->stream = saved;
  for (auto& l : ctx->lanes) {
    cudaError_t e = cudaStreamSynchronize(l.stream);
    if (e != cudaSuccess && rc == MAFE_OK) { set_error("lane sync: %s", cudaGetErrorString(e)); rc = MAFE_E_CUDA; }
  }
  return rc;
}

}  // extern "C"
