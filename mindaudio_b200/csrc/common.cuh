// common.cuh -- shared host/device plumbing of libmafe (see include/mafe.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/mafe.h"

namespace mafe {

void set_error(const char* fmt, ...);

#define MAFE_CUDA_CHECK(expr)                                                                   \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      ::mafe::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return _e == cudaErrorMemoryAllocation ? MAFE_E_OOM : MAFE_E_CUDA;                        \
    }                                                                                           \
  } while (0)

#define MAFE_REQUIRE(cond, ...)      \
  do {                               \
    if (!(cond)) {                   \
      ::mafe::set_error(__VA_ARGS__); \
      return MAFE_E_INVALID_ARG;     \
    }                                \
  } while (0)

#define MAFE_LAUNCH_CHECK(ctx)                 \
  do {                                         \
    (ctx)->launches++;                         \
    MAFE_CUDA_CHECK(cudaGetLastError());       \
  } while (0)

// one tile of the ragged batch: up to `tile_frames` consecutive frames of one utterance
struct Tile {
  int32_t utt;
  int32_t frame0;
};

// sparse mel filterbank, two forms
struct MelCSR {         // by filter (generic kernel)
  int32_t* row_ptr = nullptr;  // [n_mels+1]
  int32_t* col = nullptr;      // [nnz]
  float* val = nullptr;        // [nnz]
  int32_t nnz = 0;
};

}  // namespace mafe

struct mafe_prof_pending {
  int which;
  cudaEvent_t e0, e1;
};

struct mafe_batch;
struct mafe_lane {  // one lane of the pipelined host path: stream + grow-only device buffers
  cudaStream_t stream = nullptr;
  void* wave_dev = nullptr;
  size_t wave_cap = 0;
  float* out_dev = nullptr;
  size_t out_cap = 0;
  mafe_batch* batch = nullptr;
};

struct mafe_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int64_t launches = 0;
  bool profile = false;
  double prof_ms[MAFE_PROF_COUNT] = {0, 0, 0, 0};
  int64_t prof_n[MAFE_PROF_COUNT] = {0, 0, 0, 0};
  std::vector<mafe_prof_pending> prof_pending;
  std::vector<mafe_lane> lanes;
  // pinned staging of mafe_memcpy_h2d_gather / mafe_memcpy_d2h_staged (grow-only; one buffer per direction)
  void* stage_up = nullptr;
  size_t cap_stage_up = 0;
  cudaEvent_t stage_up_done = nullptr;   // the upload that last read stage_up
  void* stage_down = nullptr;
  size_t cap_stage_down = 0;
  float* aux_mel = nullptr;   // mafe_frontend_run_aux: second output of the running call (mel energies), else NULL
};

namespace mafe {
// RAII bracket: records events around the launches issued while it is alive (only when profiling)
struct ProfScope {
  mafe_ctx* ctx;
  mafe_prof_pending p;
  bool on;
  ProfScope(mafe_ctx* c, int which) : ctx(c), on(c->profile) {
    if (!on) return;
    p.which = which;
    cudaEventCreate(&p.e0);
    cudaEventCreate(&p.e1);
    cudaEventRecord(p.e0, ctx->stream);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(p.e1, ctx->stream);
    ctx->prof_pending.push_back(p);
  }
};
}  // namespace mafe

struct mafe_plan {
  mafe_frontend_desc d;  // table pointers nulled after creation
  int device = 0;
  int n_bins = 0;
  int out_dim = 0;
  bool fast = false;
  // generic kernel
  std::vector<int> radices;
  int* radices_dev = nullptr;
  int n_stages = 0;
  int pairs_per_tile = 0;  // complex FFTs per CTA and pass; a batch tile holds several such groups
  bool tables_in_smem = false;   // generic kernel: window / mel CSR staged in shared memory
  int tile_frames = 0;
  size_t smem_bytes = 0;
  float2* twiddle_dev = nullptr;  // [n_fft] W_N^k
  float* window_dev = nullptr;    // [frame_len]
  mafe::MelCSR mel;
  float* dct_dev = nullptr;  // [n_mels][n_mfcc]
  float* dct_img_dev = nullptr;  // tensor-core path (dct_mma.cuh): B_hi | B_lo images, N x K floats each, or null
  int dct_n_pad = 0;             // N: n_mfcc padded to a multiple of 16
  // fast path tables (fbank512.cu)
  void* fast_tables = nullptr;
};

struct mafe_batch {
  int device = 0;
  int32_t n_utts = 0;
  int64_t total_frames = 0;
  int64_t total_samples = 0;
  int64_t wave_len = 0;  // elements of the flat waveform array the offsets index into (= last offset)
  int32_t n_tiles = 0;
  int32_t n_groups = 0;
  std::vector<int64_t> frame_offsets_host;
  std::vector<int64_t> so_host;
  std::vector<mafe::Tile> tiles_host;
  size_t cap_offsets = 0, cap_foffsets = 0, cap_tiles = 0, cap_utt_sum = 0, cap_groups = 0, cap_utt_group = 0,
         cap_utt_stats = 0, cap_scratch = 0;
  int64_t* sample_offsets_dev = nullptr;  // [n_utts+1]
  int64_t* frame_offsets_dev = nullptr;   // [n_utts+1]
  mafe::Tile* tiles_dev = nullptr;        // [n_tiles]
  int32_t* utt_group_dev = nullptr;       // [n_utts] or null
  double* utt_sum_dev = nullptr;          // [n_utts] frame-mean accumulators
  int32_t* group_max_dev = nullptr;       // [max(n_utts,1)] ordered-int keys of the dB maxima
  float* scratch_dev = nullptr;           // MFCC intermediate [total_frames][n_mels]
  size_t scratch_bytes = 0;
  double* utt_stats_dev = nullptr;        // [n_utts][2][dim] fused utterance-CMVN statistics
  int32_t* queue_dev = nullptr;           // persistent-kernel work queue head (1 int)
  void* tile_recs_dev = nullptr;          // v6 kernel: one 64-byte work record per tile (tile_prepare_kernel)
  size_t cap_tile_recs = 0;               // in records
  int32_t* utt_done_dev = nullptr;        // v6 kernel, fused CMVN: completed tiles per utterance
  size_t cap_utt_done = 0;
  int64_t max_utt_frames = 0;             // longest utterance of the batch in frames
};

namespace mafe {
// generic.cu
int generic_plan_init(mafe_ctx* ctx, mafe_plan* p, const mafe_frontend_desc* desc);
void generic_plan_free(mafe_plan* p);
int generic_run(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, const void* wave, int wave_dtype, float wave_scale,
                float* out, int out_kind_override, int db_group);
int frame_mean_prepass(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, const void* wave, int wave_dtype,
                       float wave_scale);
int db_clamp_run(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, float* data, int dim, int db_group);
int dct_run(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, const float* logmel, float* out, int db_group);
// fbank512.cu
bool fast_plan_supported(const mafe_frontend_desc* desc);
int fast_plan_init(mafe_ctx* ctx, mafe_plan* p, const mafe_frontend_desc* desc);
void fast_plan_free(mafe_plan* p);
int fast_tile_frames();
// returns MAFE_E_UNSUPPORTED (without setting an error) when this call must take the generic route instead, and
// kFastNeedsPost when the features were produced but the generic post-processing (top_db clamp / DCT / CMVN) remains
constexpr int kFastNeedsPost = 1;
int fast_run(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, const void* wave, int wave_dtype, float wave_scale,
             float* out, int db_group);
bool fast_has_aux_mel(const mafe_plan* p);
}  // namespace mafe

// ---- device helpers ----
#ifdef __CUDACC__
namespace mafe {

__device__ __forceinline__ int ordered_key(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float key_to_float(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

// np.pad index map for a centred frame: s in [-(pad), L+pad) -> [0, L) or -1 (zero)
__device__ __forceinline__ int64_t pad_index(int64_t s, int64_t L, int mode) {
  if (s >= 0 && s < L) return s;
  switch (mode) {
    case MAFE_PAD_REFLECT:
      if (L == 1) return 0;
      { int64_t period = 2 * (L - 1); int64_t m = s % period; if (m < 0) m += period; return m < L ? m : period - m; }
    case MAFE_PAD_SYMMETRIC:
      { int64_t period = 2 * L; int64_t m = s % period; if (m < 0) m += period; return m < L ? m : period - 1 - m; }
    case MAFE_PAD_EDGE:
      return s < 0 ? 0 : L - 1;
    default:
      return -1;
  }
}

// same map, cheap path first: one reflection / clamp in plain compares; the general (multiply reflected) case falls back
__device__ __forceinline__ int64_t pad_index_fast(int64_t s, int64_t L, int mode) {
  if (s >= 0 && s < L) return s;
  int64_t u;
  switch (mode) {
    case MAFE_PAD_REFLECT: u = s < 0 ? -s : 2 * (L - 1) - s; break;
    case MAFE_PAD_SYMMETRIC: u = s < 0 ? -s - 1 : 2 * L - 1 - s; break;
    case MAFE_PAD_EDGE: return s < 0 ? 0 : L - 1;
    default: return -1;
  }
  return (u >= 0 && u < L) ? u : pad_index(s, L, mode);
}

// Philox4x32-10 (Salmon et al. 2011), same constants as oracle/restated.py
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// standard-normal dither sample i of utterance utt (oracle/restated.py: dither_noise)
__device__ __forceinline__ float dither_normal(uint64_t i, uint32_t utt, uint64_t seed) {
  uint32_t c0 = (uint32_t)i, c1 = (uint32_t)(i >> 32), c2 = utt, c3 = 0;
  philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
  float u1 = ((float)(c0 >> 8) + 1.0f) * 5.9604644775390625e-08f;  // ((w0>>8)+1) * 2^-24, in (0,1], exact
  float u2 = (float)(c1 >> 8) * 5.9604644775390625e-08f;
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

}  // namespace mafe
#endif
