// api.cu -- C ABI of libmafe.so (include/mafe.h): context, memory plumbing, plans, ragged batch
// layout and the dispatch of the hot path to the specialised (fbank512.cu) or generic
// (generic.cu) kernels.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <new>

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
      p->tile_frames = fast_tile_frames();
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
int mafe_batch_create(mafe_ctx* ctx, const mafe_plan* plan, const int64_t* so, int32_t n_utts, const int32_t* utt_group,
                      mafe_batch** out) {
  MAFE_REQUIRE(ctx && plan && out && (so || n_utts == 0), "mafe_batch_create: NULL argument");
  MAFE_REQUIRE(n_utts >= 0, "n_utts=%d", n_utts);
  DeviceGuard g(ctx->device);
  mafe_batch* b = new (std::nothrow) mafe_batch();
  if (!b) { set_error("out of host memory"); return MAFE_E_OOM; }
  b->device = ctx->device;
  b->n_utts = n_utts;
  b->frame_offsets_host.assign((size_t)n_utts + 1, 0);
  std::vector<Tile> tiles;
  const int tf = plan->tile_frames;
  int32_t max_group = -1;
  for (int32_t u = 0; u < n_utts; ++u) {
    int64_t len = so[u + 1] - so[u];
    if (len < 0) { delete b; set_error("sample_offsets must be non-decreasing (utt %d)", u); return MAFE_E_INVALID_ARG; }
    int64_t t = mafe_plan_num_frames(plan, len);
    if (t > INT32_MAX - tf) { delete b; set_error("utterance %d has too many frames", u); return MAFE_E_INVALID_ARG; }
    b->frame_offsets_host[u + 1] = b->frame_offsets_host[u] + t;
    for (int64_t f0 = 0; f0 < t; f0 += tf) tiles.push_back(Tile{u, (int32_t)f0});
    if (utt_group) max_group = std::max(max_group, utt_group[u]);
  }
  b->total_frames = b->frame_offsets_host[n_utts];
  b->total_samples = n_utts ? so[n_utts] - so[0] : 0;
  b->wave_len = n_utts ? so[n_utts] : 0;
  b->n_tiles = (int32_t)tiles.size();
  b->n_groups = utt_group ? max_group + 1 : std::max(n_utts, 1);
  cudaStream_t st = ctx->stream;
  auto fail = [&](cudaError_t e) {
    set_error("mafe_batch_create: %s", cudaGetErrorString(e));
    mafe_batch_destroy(b);
    return e == cudaErrorMemoryAllocation ? MAFE_E_OOM : MAFE_E_CUDA;
  };
  cudaError_t e;
  size_t no = (size_t)n_utts + 1;
  if ((e = cudaMalloc((void**)&b->sample_offsets_dev, no * 8)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&b->frame_offsets_dev, no * 8)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&b->tiles_dev, std::max<size_t>(tiles.size(), 1) * sizeof(Tile))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&b->utt_sum_dev, std::max<size_t>(n_utts, 1) * 8)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&b->group_max_dev, (size_t)std::max(b->n_groups, 1) * 4)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc((void**)&b->work_counter_dev, 64)) != cudaSuccess) return fail(e);
  // offsets relative to so[0] are NOT rebased: wave_dev must point at sample 0 of the flat array
  if (n_utts) {
    if ((e = cudaMemcpyAsync(b->sample_offsets_dev, so, no * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e);
  } else {
    int64_t z = 0;
    if ((e = cudaMemcpyAsync(b->sample_offsets_dev, &z, 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e);
  }
  if ((e = cudaMemcpyAsync(b->frame_offsets_dev, b->frame_offsets_host.data(), no * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e);
  if (!tiles.empty())
    if ((e = cudaMemcpyAsync(b->tiles_dev, tiles.data(), tiles.size() * sizeof(Tile), cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e);
  if (utt_group && n_utts) {
    if ((e = cudaMalloc((void**)&b->utt_group_dev, (size_t)n_utts * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpyAsync(b->utt_group_dev, utt_group, (size_t)n_utts * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail(e);
  }
  if ((plan->d.utt_cmvn_mean || plan->d.utt_cmvn_std) && n_utts > 0) {
    if ((e = cudaMalloc((void**)&b->utt_stats_dev, (size_t)n_utts * 2 * plan->out_dim * sizeof(double))) != cudaSuccess) return fail(e);
  }
  if (plan->d.out_kind == MAFE_OUT_MFCC && b->total_frames > 0) {
    b->scratch_bytes = (size_t)b->total_frames * plan->d.n_mels * sizeof(float);
    if ((e = cudaMalloc((void**)&b->scratch_dev, b->scratch_bytes)) != cudaSuccess) return fail(e);
  }
  // the tile vector is pageable host memory: make sure the copies have consumed it before returning
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return fail(e);
  *out = b;
  return MAFE_OK;
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
  cudaFree(b->work_counter_dev);
  cudaFree(b->utt_stats_dev);
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
  MAFE_REQUIRE(!(cmvn && d.out_kind == MAFE_OUT_COMPLEX), "utterance CMVN needs a real-valued output kind");
  if (plan->fast) return fast_run(ctx, plan, batch, wave_dev, wave_dtype, wave_scale, out_dev);  // CMVN fused inside
  int rc;
  if (d.out_kind == MAFE_OUT_MFCC) {
    rc = generic_run(ctx, plan, batch, wave_dev, wave_dtype, wave_scale, batch->scratch_dev, MAFE_OUT_LOGMEL, db_group);
    if (rc) return rc;
    rc = dct_run(ctx, plan, batch, batch->scratch_dev, out_dev, db_group);
  } else {
    rc = generic_run(ctx, plan, batch, wave_dev, wave_dtype, wave_scale, out_dev, -1, db_group);
    if (rc) return rc;
    if (d.out_kind == MAFE_OUT_LOGMEL && d.log_kind == MAFE_LOG_DB)
      rc = db_clamp_run(ctx, plan, batch, out_dev, plan->out_dim, db_group);
  }
  if (rc) return rc;
  if (cmvn)
    return mafe_cmvn_utt(ctx, out_dev, batch->frame_offsets_dev, batch->n_utts, plan->out_dim, d.utt_cmvn_mean, d.utt_cmvn_std);
  return MAFE_OK;
}

}  // extern "C"
