// generic.cu -- the general front-end kernel: any n_fft (mixed radix 2/3/4/5 + prime fallback),
// any window / pad mode / hop, every output kind of include/mafe.h.  One CTA owns a tile of
// 2*P consecutive frames of one utterance; frame pairs are packed (a + i*b) into one complex
// shared-memory Stockham FFT and separated afterwards (X_a = (Z[k]+conj Z[N-k])/2, ...).
//
// Serves (reference file:line): spectrum.stft (mindaudio/data/spectrum.py:125-278), the
// Spectrogram/MelScale pair behind spectrum.spectrogram/melspectrogram (:547-698), features.fbank /
// features.mfcc (mindaudio/data/features.py:196-373) and, as the non-specialised route, the
// conformer front-end (examples/conformer/dataset.py:117-168).  FP32 arithmetic; twiddles and
// tables are rounded from float64.
#include <math.h>

#include <algorithm>

#include "common.cuh"
#include "fft_generic.cuh"
#include "dct_mma.cuh"

namespace mafe {

constexpr int kGenericThreads = 256;

struct GenericParams {
  const void* wave;
  int wave_dtype;
  float wave_scale;
  const int64_t* sample_offsets;
  const int64_t* frame_offsets;
  const Tile* tiles;
  const double* utt_sum;
  int n_fft, frame_len, hop, center, pad_mode, n_bins;
  float pre_hi, pre_lo;  // pre-emphasis coefficient split hi+lo (double -> 2 floats)
  int preemph_on, remove_mean;
  float dither;
  uint64_t seed;
  const float* window;
  const float2* tw;
  FftStages fft;
  int pairs;
  int pair_stride;
  int sub_tiles;        // groups of 2*pairs frames per batch tile
  int tables_in_smem;   // window (+ mel CSR) staged in shared memory behind the ping-pong buffers
  int nnz;
  int out_kind;
  float power, spec_scale;
  int n_mels;
  const int* row_ptr;
  const int* col;
  const float* val;
  int log_kind;
  float log_arg, log_mult, log_offset;
  float* out;
  int out_dim;
  int* group_max;
  const int* utt_group;
  int db_group;
};

__device__ __forceinline__ float raw_sample(const GenericParams& P, int64_t off, int64_t m, uint32_t utt) {
  float v = P.wave_dtype == MAFE_WAVE_I16 ? (float)((const int16_t*)P.wave)[off + m] : ((const float*)P.wave)[off + m];
  v *= P.wave_scale;
  if (P.dither != 0.0f) v = fmaf(P.dither, dither_normal((uint64_t)m, utt, P.seed), v);
  return v;
}

// y(m): dithered, pre-emphasised sample m of the utterance (dataset.py:117-119: y[0] = x[0])
__device__ __forceinline__ float signal_sample(const GenericParams& P, int64_t off, int64_t m, uint32_t utt) {
  float v = raw_sample(P, off, m, utt);
  if (P.preemph_on && m > 0) {
    float vp = raw_sample(P, off, m - 1, utt);
    v = fmaf(-P.pre_lo, vp, fmaf(-P.pre_hi, vp, v));
  }
  return v;
}

// windowed entry n of frame f (before the scalar mean is removed); 0 outside the utterance's frames
__device__ __forceinline__ float frame_entry(const GenericParams& P, const float* __restrict__ win, int64_t off, int64_t L, int64_t T,
                                             int64_t f, int n, uint32_t utt) {
  if (f >= T || n >= P.frame_len) return 0.0f;
  int64_t s = f * P.hop + n;
  if (P.center) {
    s = pad_index_fast(s - P.n_fft / 2, L, P.pad_mode);
    if (s < 0) return 0.0f;
  }
  return signal_sample(P, off, s, utt) * win[n];
}

// ---------------------------------------------------------------------------------------------
// pre-pass: sum of all windowed frame entries per utterance (conformer/dataset.py:165 needs the
// mean of the WHOLE [T, frame_len] matrix before the FFT).  double accumulation.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGenericThreads) frame_sum_kernel(GenericParams P, double* utt_sum, int tile_frames) {
  const Tile tile = P.tiles[blockIdx.x];
  const int64_t off = P.sample_offsets[tile.utt];
  const int64_t L = P.sample_offsets[tile.utt + 1] - off;
  const int64_t T = P.frame_offsets[tile.utt + 1] - P.frame_offsets[tile.utt];
  double acc = 0.0;
  const int total = tile_frames * P.frame_len;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    int f = idx / P.frame_len, n = idx - f * P.frame_len;
    acc += (double)frame_entry(P, P.window, off, L, T, tile.frame0 + f, n, (uint32_t)tile.utt);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double warp_sums[kGenericThreads / 32];
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kGenericThreads / 32; ++w) s += warp_sums[w];
    atomicAdd(&utt_sum[tile.utt], s);
  }
}

// ---------------------------------------------------------------------------------------------
// main generic kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kGenericThreads) generic_frontend_kernel(GenericParams P) {
  extern __shared__ float2 smem[];
  const int N = P.n_fft;
  const int pairs = P.pairs;
  const int ps = P.pair_stride;   // elements between the sequences of consecutive pairs (N, or N + N/16 with radix-16 passes)
  float2* buf0 = smem;
  float2* buf1 = smem + (size_t)pairs * ps;

  // per-plan tables staged once per CTA (the CTA then walks all sub-tiles of its batch tile): window, mel CSR
  const float* win = P.window;
  const int* row_ptr = P.row_ptr;
  const int* col = P.col;
  const float* val = P.val;
  if (P.tables_in_smem) {
    float* t_win = reinterpret_cast<float*>(smem + (size_t)2 * pairs * ps);
    for (int i = threadIdx.x; i < P.frame_len; i += blockDim.x) t_win[i] = P.window[i];
    win = t_win;
    if (P.out_kind >= MAFE_OUT_MEL) {
      int* t_rp = reinterpret_cast<int*>(t_win + P.frame_len);
      int* t_col = t_rp + P.n_mels + 1;
      float* t_val = reinterpret_cast<float*>(t_col + P.nnz);
      for (int i = threadIdx.x; i <= P.n_mels; i += blockDim.x) t_rp[i] = P.row_ptr[i];
      for (int i = threadIdx.x; i < P.nnz; i += blockDim.x) { t_col[i] = P.col[i]; t_val[i] = P.val[i]; }
      row_ptr = t_rp; col = t_col; val = t_val;
    }
    __syncthreads();
  }

  const Tile tile0 = P.tiles[blockIdx.x];
  const uint32_t utt = (uint32_t)tile0.utt;
  const int64_t off = P.sample_offsets[utt];
  const int64_t L = P.sample_offsets[utt + 1] - off;
  const int64_t fo = P.frame_offsets[utt];
  const int64_t T = P.frame_offsets[utt + 1] - fo;
  float mu = 0.0f;
  if (P.remove_mean) mu = (float)(P.utt_sum[utt] / ((double)T * (double)P.frame_len));
  float vmax = -INFINITY;

  for (int sub = 0; sub < P.sub_tiles; ++sub) {
  Tile tile = tile0;
  tile.frame0 += sub * 2 * pairs;   // the batch tile holds sub_tiles groups of 2*pairs frames
  if (tile.frame0 >= T) break;
  if (sub > 0) __syncthreads();     // the previous group's buffers have been read
  float2* cur = buf0;
  float2* nxt = buf1;

  // ---- load: frame pair p -> complex sequence a + i b, windowed, scalar mean removed ----
  for (int p = 0; p < pairs; ++p) {
    const int64_t fa = tile.frame0 + 2 * p;
    // interior pair (both frames exist, no sample outside the utterance, float input without dither / pre-emphasis):
    // straight coalesced loads, four in flight per thread
    const int64_t s_lo = fa * P.hop - (P.center ? P.n_fft / 2 : 0);
    const bool interior = fa + 1 < T && s_lo >= 0 && s_lo + P.hop + P.frame_len <= L && P.wave_dtype == MAFE_WAVE_F32 &&
                          P.dither == 0.0f && !P.preemph_on;
    if (interior) {
      const float* wa = (const float*)P.wave + off + s_lo;
      const float* wb = wa + P.hop;
      const float sc_w = P.wave_scale;
#pragma unroll 8
      for (int n = threadIdx.x; n < N; n += blockDim.x) {
        float a = 0.f, b = 0.f;
        if (n < P.frame_len) {
          const float w = win[n];
          a = fmaf(__ldg(wa + n) * sc_w, w, -mu);
          b = fmaf(__ldg(wb + n) * sc_w, w, -mu);
        }
        cur[(size_t)p * ps + n] = make_float2(a, b);
      }
      continue;
    }
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
      float a = 0.f, b = 0.f;
      if (n < P.frame_len) {   // zero padding of the window up to n_fft otherwise
        a = frame_entry(P, win, off, L, T, fa, n, utt);
        b = frame_entry(P, win, off, L, T, fa + 1, n, utt);
        if (fa < T) a -= mu;
        if (fa + 1 < T) b -= mu;
      }
      cur[(size_t)p * ps + n] = make_float2(a, b);
    }
  }
  __syncthreads();

  // ---- Stockham autosort stages ----
  {
    float2* res = stockham_fft<float2>(cur, nxt, pairs, N, P.fft, P.tw, ps);
    if (res != cur) { nxt = cur; cur = res; }
  }

  // ---- separate the pair, emit ----
  const int nb = P.n_bins;
  float* pw = reinterpret_cast<float*>(nxt);  // [2*pairs][nb] power spectra (aliases the idle buffer)
  const float sc = P.spec_scale;
  for (int p = 0; p < pairs; ++p)
  for (int k = threadIdx.x; k < nb; k += blockDim.x) {
    float2 za = cur[(size_t)p * ps + k];
    float2 zr = cur[(size_t)p * ps + (k == 0 ? 0 : N - k)];
    float2 xa = make_float2(0.5f * (za.x + zr.x) * sc, 0.5f * (za.y - zr.y) * sc);
    float2 xb = make_float2(0.5f * (za.y + zr.y) * sc, -0.5f * (za.x - zr.x) * sc);
    int64_t fa = tile.frame0 + 2 * p;
    if (P.out_kind == MAFE_OUT_COMPLEX) {
      if (fa < T) reinterpret_cast<float2*>(P.out + (fo + fa) * P.out_dim)[k] = xa;
      if (fa + 1 < T) reinterpret_cast<float2*>(P.out + (fo + fa + 1) * P.out_dim)[k] = xb;
    } else {
      float pa = fmaf(xa.x, xa.x, xa.y * xa.y), pb = fmaf(xb.x, xb.x, xb.y * xb.y);
      if (P.power != 2.0f) {
        if (P.power == 1.0f) { pa = sqrtf(pa); pb = sqrtf(pb); }
        else { pa = powf(sqrtf(pa), P.power); pb = powf(sqrtf(pb), P.power); }
      }
      if (P.out_kind == MAFE_OUT_POWER) {
        if (P.log_kind == MAFE_LOG_LN_PLUS) {   // log-compressed spectrogram (deepspeech2: log1p of the magnitude)
          pa = P.log_arg == 1.0f ? log1pf(pa) : logf(pa + P.log_arg);
          pb = P.log_arg == 1.0f ? log1pf(pb) : logf(pb + P.log_arg);
        }
        if (fa < T) P.out[(fo + fa) * P.out_dim + k] = pa;
        if (fa + 1 < T) P.out[(fo + fa + 1) * P.out_dim + k] = pb;
      } else {
        pw[(2 * p) * nb + k] = pa;
        pw[(2 * p + 1) * nb + k] = pb;
      }
    }
  }
  if (P.out_kind <= MAFE_OUT_POWER) continue;
  __syncthreads();

  // ---- sparse mel projection (CSR by filter) + log ----
  const int nm = P.n_mels;
  for (int idx = threadIdx.x; idx < 2 * pairs * nm; idx += blockDim.x) {
    int f = idx / nm, m = idx - f * nm;
    if (tile.frame0 + f >= T) continue;
    const float* row = pw + f * nb;
    float acc = 0.f;
    if (P.tables_in_smem) {
      // explicit shared-memory loads (through the generic pointers the compiler emits LD, not LDS)
      const uint32_t a_rp = (uint32_t)__cvta_generic_to_shared(row_ptr), a_col = (uint32_t)__cvta_generic_to_shared(col),
                     a_val = (uint32_t)__cvta_generic_to_shared(val);
      int i0, i1;
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(i0) : "r"(a_rp + 4u * m));
      asm volatile("ld.shared.s32 %0, [%1];" : "=r"(i1) : "r"(a_rp + 4u * m + 4u));
      for (int i = i0; i < i1; ++i) {
        int c; float v;
        asm volatile("ld.shared.s32 %0, [%1];" : "=r"(c) : "r"(a_col + 4u * i));
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a_val + 4u * i));
        acc = fmaf(v, row[c], acc);
      }
    } else {
      for (int i = row_ptr[m]; i < row_ptr[m + 1]; ++i) acc = fmaf(__ldg(&val[i]), row[__ldg(&col[i])], acc);
    }
    float o = acc;
    switch (P.log_kind) {
      case MAFE_LOG_LN_EPS_IF_ZERO: o = logf(acc == 0.f ? 2.220446049250313e-16f : acc); break;
      case MAFE_LOG_LN_PLUS: o = logf(acc + P.log_arg); break;
      case MAFE_LOG_DB: o = P.log_mult * log10f(fmaxf(acc, P.log_arg)) - P.log_offset; vmax = fmaxf(vmax, o); break;
      default: break;
    }
    P.out[(fo + tile.frame0 + f) * P.out_dim + m] = o;
  }
  }   // sub-tile loop
  if (P.out_kind >= MAFE_OUT_MEL && P.log_kind == MAFE_LOG_DB && P.db_group != MAFE_DBGROUP_NONE) {
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((threadIdx.x & 31) == 0 && vmax > -INFINITY) {
      int g = P.db_group == MAFE_DBGROUP_UTT ? (int)utt : (P.db_group == MAFE_DBGROUP_BATCH ? 0 : P.utt_group[utt]);
      atomicMax(&P.group_max[g], ordered_key(vmax));
    }
  }
}

// top_db clamp in place: x = max(x, groupmax - top_db)   (spectrum.py:78-89)
__global__ void __launch_bounds__(256, 8) db_clamp_kernel(float* __restrict__ data, int dim, const Tile* __restrict__ tiles, int n_tiles,
                                                        const int64_t* __restrict__ frame_offsets, int tile_frames,
                                                        const int* __restrict__ group_max, const int* __restrict__ utt_group, int db_group,
                                                        float top_db) {
  // a CTA takes a contiguous range of tiles: the chain tile -> offsets -> group maximum is paid once per utterance
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(n_tiles, t0 + per);
  int utt = -1;
  int64_t fo = 0, T = 0;
  float floor_v = 0.f;
  for (int ti = t0; ti < t1; ++ti) {
    const Tile tile = tiles[ti];
    if (tile.utt != utt) {
      utt = tile.utt;
      fo = frame_offsets[utt];
      T = frame_offsets[utt + 1] - fo;
      const int g = db_group == MAFE_DBGROUP_UTT ? utt : (db_group == MAFE_DBGROUP_BATCH ? 0 : utt_group[utt]);
      floor_v = key_to_float(group_max[g]) - top_db;
    }
    const int nf = (int)min((int64_t)tile_frames, T - tile.frame0);
    float* base = data + (fo + tile.frame0) * dim;
    const int cnt = nf * dim;
#pragma unroll 4
    for (int i = threadIdx.x; i < cnt; i += 256) base[i] = fmaxf(base[i], floor_v);
  }
}

// MFCC: out[f][c] = sum_m clamp(logmel[f][m]) * dct[m][c]   (features.py:356-361)
// Register-tiled small GEMM.  A CTA walks groups of 4 tiles (<= 128 frames): the (clamped) log-mel rows are transposed
// into shared memory [m][frame], the DCT table (zero padded to 8*CPT columns) is loaded once per CTA; thread
// (frame quad fq, coefficient group cg) keeps 4 x CPT accumulators: per mel bin one 16-byte load of 4 frames and CPT
// table values feed 4*CPT FMAs.
constexpr int kDctFrames = 128;
constexpr int kDctXStride = kDctFrames + 4;   // 16 B aligned rows
template <int CPT>
__global__ void __launch_bounds__(256) dct_kernel(const float* __restrict__ logmel, float* __restrict__ out, int n_mels, int n_mfcc,
                                                  const float* __restrict__ dct, const Tile* __restrict__ tiles, int n_tiles,
                                                  const int64_t* __restrict__ frame_offsets, int tile_frames,
                                                  const int* __restrict__ group_max, const int* __restrict__ utt_group, int db_group,
                                                  float top_db) {
  extern __shared__ __align__(16) float sm[];
  constexpr int NCP = 8 * CPT;
  float* sd = sm;                         // [n_mels][NCP]
  float* sx = sm + n_mels * NCP;          // [n_mels][kDctXStride]
  __shared__ int64_t s_row[kDctFrames / 2];   // first output row of each tile of the group (tiles hold >= 2 frames)
  __shared__ int s_nf[kDctFrames / 2];
  __shared__ float s_floor[kDctFrames / 2];
  const int tid = threadIdx.x;
  for (int i = tid; i < n_mels * NCP; i += 256) {
    const int m = i / NCP, c = i - m * NCP;
    sd[i] = c < n_mfcc ? dct[m * n_mfcc + c] : 0.f;
  }
  const int fq = tid >> 3, cg = tid & 7;
  const int tiles_per_group = kDctFrames / tile_frames;   // tile_frames = 32 -> 4; always >= 1 (tile_frames <= 32)
  const int group_frames = tiles_per_group * tile_frames;
  const int n_groups = (n_tiles + tiles_per_group - 1) / tiles_per_group;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    __syncthreads();   // previous group's sx / s_* are no longer read
    if (tid < tiles_per_group) {
      const int ti = grp * tiles_per_group + tid;
      int nf = 0;
      int64_t row = 0;
      float fl = -INFINITY;
      if (ti < n_tiles) {
        const Tile tile = tiles[ti];
        const int64_t fo = frame_offsets[tile.utt];
        const int64_t T = frame_offsets[tile.utt + 1] - fo;
        nf = (int)min((int64_t)tile_frames, T - tile.frame0);
        row = fo + tile.frame0;
        if (db_group != MAFE_DBGROUP_NONE) {
          const int g = db_group == MAFE_DBGROUP_UTT ? tile.utt : (db_group == MAFE_DBGROUP_BATCH ? 0 : utt_group[tile.utt]);
          fl = key_to_float(group_max[g]) - top_db;
        }
      }
      s_row[tid] = row; s_nf[tid] = nf; s_floor[tid] = fl;
    }
    __syncthreads();
    for (int q = 0; q < tiles_per_group; ++q) {
      const int nf = s_nf[q];
      const float fl = s_floor[q];
      const float* src = logmel + s_row[q] * n_mels;
      for (int i = tid; i < tile_frames * n_mels; i += 256) {
        const int f = i / n_mels, m = i - f * n_mels;
        sx[m * kDctXStride + q * tile_frames + f] = f < nf ? fmaxf(src[i], fl) : 0.f;
      }
    }
    __syncthreads();
    float acc[4][CPT];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < CPT; ++i) acc[j][i] = 0.f;
    const float* xp = sx + 4 * fq;
    const float* dp = sd + cg * CPT;
#pragma unroll 4
    for (int m = 0; m < n_mels; ++m) {
      const float4 x = *reinterpret_cast<const float4*>(xp + m * kDctXStride);
      float d[CPT];
#pragma unroll
      for (int i = 0; i < CPT; ++i) d[i] = dp[m * NCP + i];
#pragma unroll
      for (int i = 0; i < CPT; ++i) {
        acc[0][i] = fmaf(x.x, d[i], acc[0][i]);
        acc[1][i] = fmaf(x.y, d[i], acc[1][i]);
        acc[2][i] = fmaf(x.z, d[i], acc[2][i]);
        acc[3][i] = fmaf(x.w, d[i], acc[3][i]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = 4 * fq + j;
      if (idx >= group_frames) continue;
      const int q = idx / tile_frames, f = idx - q * tile_frames;
      if (f >= s_nf[q]) continue;
      float* dst = out + (s_row[q] + f) * n_mfcc + cg * CPT;
#pragma unroll
      for (int i = 0; i < CPT; ++i)
        if (cg * CPT + i < n_mfcc) dst[i] = acc[j][i];
    }
  }
}

__global__ void fill_int_kernel(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static void factorize(int n, std::vector<int>& out) {
  out.clear();
  if (n >= 256 && (n & (n - 1)) == 0) {
    // power of two: register-resident radix-16 passes, one smaller pass last (8 / 4 / 2) -- 2048 = 16 * 16 * 8
    while (n % 16 == 0 && n > 16) { out.push_back(16); n /= 16; }
    if (n == 16) { out.push_back(16); n = 1; }
    if (n > 1) out.push_back(n);   // 8, 4 or 2
    return;
  }
  while (n % 4 == 0) { out.push_back(4); n /= 4; }
  while (n % 2 == 0) { out.push_back(2); n /= 2; }
  while (n % 3 == 0) { out.push_back(3); n /= 3; }
  while (n % 5 == 0) { out.push_back(5); n /= 5; }
  for (int f = 7; (long long)f * f <= n; f += 2)
    while (n % f == 0) { out.push_back(f); n /= f; }
  if (n > 1) out.push_back(n);
}

template <typename T>
static int upload(T** dev, const T* host, size_t n) {
  MAFE_CUDA_CHECK(cudaMalloc((void**)dev, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) MAFE_CUDA_CHECK(cudaMemcpy(*dev, host, n * sizeof(T), cudaMemcpyHostToDevice));
  return MAFE_OK;
}

int generic_plan_init(mafe_ctx* ctx, mafe_plan* p, const mafe_frontend_desc* d) {
  const int N = d->n_fft;
  factorize(N, p->radices);
  MAFE_REQUIRE((int)p->radices.size() <= kMaxStages, "n_fft=%d has too many factors", N);
  p->n_stages = (int)p->radices.size();
  // tile size: as many frame pairs as fit ~64 KB of ping-pong buffers, 1..8
  // radix-16 / 8 passes skew their intermediate buffers by one element per 16 (fft_generic.cuh: pad16)
  const int pair_stride = p->radices[0] >= 8 ? N + N / 16 : N;
  size_t per_pair = (size_t)pair_stride * sizeof(float2) * 2;
  int pairs = (int)std::min<size_t>(8, std::max<size_t>(1, (68 * 1024) / per_pair));
  p->pairs_per_tile = pairs;
  // a batch tile = several groups of 2*pairs frames walked by ONE CTA (~32 frames): the window and the mel CSR are
  // staged in shared memory once per CTA instead of being fetched from global memory for every 2*pairs frames
  p->tile_frames = 2 * pairs * std::max(1, 16 / pairs);
  p->smem_bytes = per_pair * pairs;
  if (p->smem_bytes > 200 * 1024) {
    set_error("n_fft=%d needs %zu bytes of shared memory per frame pair (max n_fft is 8192)", N, p->smem_bytes);
    return MAFE_E_UNSUPPORTED;
  }
  if (d->out_kind >= MAFE_OUT_MEL) {
    // the power buffer [2*pairs][n_bins] aliases one ping-pong buffer: always fits (n_bins <= N)
  }
  std::vector<float2> tw(N);
  for (int k = 0; k < N; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)N;
    tw[k] = make_float2((float)cos(a), (float)sin(a));
  }
  int rc = upload(&p->twiddle_dev, tw.data(), tw.size());
  if (rc) return rc;
  rc = upload(&p->window_dev, d->window, (size_t)d->frame_len);
  if (rc) return rc;
  if (d->out_kind >= MAFE_OUT_MEL) {
    std::vector<int> row_ptr(d->n_mels + 1, 0), col;
    std::vector<float> val;
    for (int m = 0; m < d->n_mels; ++m) {
      for (int k = 0; k < p->n_bins; ++k) {
        float w = d->mel_fb[(size_t)m * p->n_bins + k];
        if (w != 0.0f) { col.push_back(k); val.push_back(w); }
      }
      row_ptr[m + 1] = (int)col.size();
    }
    p->mel.nnz = (int)col.size();
    if ((rc = upload(&p->mel.row_ptr, row_ptr.data(), row_ptr.size()))) return rc;
    if ((rc = upload(&p->mel.col, col.data(), col.size()))) return rc;
    if ((rc = upload(&p->mel.val, val.data(), val.size()))) return rc;
  }
  if (d->out_kind == MAFE_OUT_MFCC) {
    if ((rc = upload(&p->dct_dev, d->dct, (size_t)d->n_mels * d->n_mfcc))) return rc;
    {
      // tensor-core path: DCT^T split into TF32 hi / lo parts, laid out as the UMMA canonical K-major images
      const int K = d->n_mels, N = (d->n_mfcc + 15) / 16 * 16;
      if (K % 16 == 0 && K <= 96 && N <= 64 && dct_mma_smem_bytes(K, N) <= 227 * 1024) {
        std::vector<float> img((size_t)2 * N * K, 0.f);
        for (int n = 0; n < d->n_mfcc; ++n)
          for (int k = 0; k < K; ++k) {
            const float x = d->dct[(size_t)k * d->n_mfcc + n];
            uint32_t bits;
            memcpy(&bits, &x, 4);
            bits &= 0xffffe000u;
            float hi;
            memcpy(&hi, &bits, 4);
            const size_t o = dm_canon_off(n, k, K) / 4;
            img[o] = hi;
            img[(size_t)N * K + o] = x - hi;
          }
        if ((rc = upload(&p->dct_img_dev, img.data(), img.size()))) return rc;
        p->dct_n_pad = N;
      }
    }
  }
  {
    size_t tab = sizeof(float) * (size_t)d->frame_len;
    if (d->out_kind >= MAFE_OUT_MEL) tab += sizeof(int) * (size_t)(d->n_mels + 1) + (sizeof(int) + sizeof(float)) * (size_t)p->mel.nnz;
    tab = (tab + 15) & ~(size_t)15;
    p->tables_in_smem = tab <= 48 * 1024 && p->smem_bytes + tab <= 200 * 1024;
    if (p->tables_in_smem) p->smem_bytes += tab;
  }
  if (p->smem_bytes > 48 * 1024)
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(generic_frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)p->smem_bytes));
  (void)ctx;
  return MAFE_OK;
}

void generic_plan_free(mafe_plan* p) {
  cudaFree(p->twiddle_dev);
  cudaFree(p->window_dev);
  cudaFree(p->mel.row_ptr);
  cudaFree(p->mel.col);
  cudaFree(p->mel.val);
  cudaFree(p->dct_dev);
  cudaFree(p->dct_img_dev);
}

static void fill_params(GenericParams& P, const mafe_plan* p, const mafe_batch* b, const void* wave, int wave_dtype,
                        float wave_scale) {
  const mafe_frontend_desc& d = p->d;
  P.wave = wave; P.wave_dtype = wave_dtype; P.wave_scale = wave_scale;
  P.sample_offsets = b->sample_offsets_dev; P.frame_offsets = b->frame_offsets_dev; P.tiles = b->tiles_dev;
  P.utt_sum = b->utt_sum_dev;
  P.n_fft = d.n_fft; P.frame_len = d.frame_len; P.hop = d.hop; P.center = d.center; P.pad_mode = d.pad_mode;
  P.n_bins = p->n_bins;
  P.pre_hi = (float)d.preemph; P.pre_lo = (float)(d.preemph - (double)P.pre_hi);
  P.preemph_on = d.preemph != 0.0; P.remove_mean = d.remove_frame_mean;
  P.dither = d.dither; P.seed = d.dither_seed;
  P.window = p->window_dev; P.tw = p->twiddle_dev;
  for (int i = 0; i < kMaxStages; ++i) P.fft.radices[i] = i < p->n_stages ? p->radices[i] : 1;
  P.fft.n_stages = p->n_stages; P.pairs = p->pairs_per_tile; P.pair_stride = p->radices[0] >= 8 ? p->d.n_fft + p->d.n_fft / 16 : p->d.n_fft;
  P.sub_tiles = (p->tile_frames + 2 * p->pairs_per_tile - 1) / (2 * p->pairs_per_tile);
  P.tables_in_smem = p->tables_in_smem; P.nnz = p->mel.nnz;
  P.out_kind = d.out_kind; P.power = d.power; P.spec_scale = d.spec_scale;
  P.n_mels = d.n_mels; P.row_ptr = p->mel.row_ptr; P.col = p->mel.col; P.val = p->mel.val;
  P.log_kind = d.log_kind; P.log_arg = d.log_arg; P.log_mult = d.log_mult; P.log_offset = d.log_offset;
  P.out = nullptr; P.out_dim = p->out_dim;
  P.group_max = b->group_max_dev; P.utt_group = b->utt_group_dev; P.db_group = MAFE_DBGROUP_NONE;
}

int frame_mean_prepass(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, const void* wave, int wave_dtype,
                       float wave_scale) {
  if (b->n_tiles == 0) return MAFE_OK;
  GenericParams P;
  fill_params(P, p, b, wave, wave_dtype, wave_scale);
  MAFE_CUDA_CHECK(cudaMemsetAsync(b->utt_sum_dev, 0, sizeof(double) * b->n_utts, ctx->stream));
  ProfScope ps(ctx, MAFE_PROF_FRAME_MEAN);
  frame_sum_kernel<<<b->n_tiles, kGenericThreads, 0, ctx->stream>>>(P, b->utt_sum_dev, p->tile_frames);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int generic_run(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, const void* wave, int wave_dtype, float wave_scale,
                float* out, int out_kind_override, int db_group) {
  if (b->n_tiles == 0) return MAFE_OK;
  GenericParams P;
  fill_params(P, p, b, wave, wave_dtype, wave_scale);
  if (p->d.remove_frame_mean) {
    int rc = frame_mean_prepass(ctx, p, b, wave, wave_dtype, wave_scale);
    if (rc) return rc;
  }
  P.out = out;
  if (out_kind_override >= 0) {
    P.out_kind = out_kind_override;
    P.out_dim = out_kind_override == MAFE_OUT_LOGMEL ? p->d.n_mels : p->out_dim;
  }
  P.db_group = (p->d.log_kind == MAFE_LOG_DB && p->d.top_db >= 0.f) ? db_group : MAFE_DBGROUP_NONE;
  if (P.db_group != MAFE_DBGROUP_NONE) {
    int n = std::max(b->n_groups, 1);
    fill_int_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(b->group_max_dev, n, (int)0x80000000);
    MAFE_LAUNCH_CHECK(ctx);
  }
  ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
  generic_frontend_kernel<<<b->n_tiles, kGenericThreads, p->smem_bytes, ctx->stream>>>(P);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int db_clamp_run(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, float* data, int dim, int db_group) {
  if (b->n_tiles == 0 || db_group == MAFE_DBGROUP_NONE || p->d.top_db < 0.f) return MAFE_OK;
  db_clamp_kernel<<<std::min(b->n_tiles, 8 * ctx->sm_count), 256, 0, ctx->stream>>>(data, dim, b->tiles_dev, b->n_tiles, b->frame_offsets_dev, p->tile_frames,
                                                       b->group_max_dev, b->utt_group_dev, db_group, p->d.top_db);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int dct_run(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, const float* logmel, float* out, int db_group) {
  if (b->n_tiles == 0) return MAFE_OK;
  const int nm = p->d.n_mels, nc = p->d.n_mfcc;
  MAFE_REQUIRE(nc <= 128 && p->tile_frames >= 2 && p->tile_frames <= kDctFrames, "MFCC: n_mfcc %d / tile of %d frames not supported", nc, p->tile_frames);
  if (p->dct_img_dev != nullptr && p->tile_frames == 32 && getenv("MAFE_DCT_TILED") == nullptr) {
    // tensor cores (3 x TF32): dct_mma.cuh
    DctMmaParams Q;
    Q.logmel = logmel; Q.out = out; Q.K = nm; Q.n_mfcc = nc; Q.N = p->dct_n_pad; Q.bimg = p->dct_img_dev;
    Q.tiles = b->tiles_dev; Q.n_tiles = b->n_tiles; Q.frame_offsets = b->frame_offsets_dev; Q.tile_frames = p->tile_frames;
    Q.group_max = b->group_max_dev; Q.utt_group = b->utt_group_dev;
    Q.db_group = (p->d.log_kind == MAFE_LOG_DB && p->d.top_db >= 0.f && db_group != MAFE_DBGROUP_NONE) ? db_group : MAFE_DBGROUP_NONE;
    Q.top_db = p->d.top_db;
    const size_t smem = dct_mma_smem_bytes(nm, p->dct_n_pad);
    const int n_items = (b->n_tiles + 3) / 4;
    const int grid = std::max(1, std::min((n_items + 1) / 2, ctx->sm_count));
    auto launch = [&](auto kern) -> int {
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      kern<<<grid, 256, smem, ctx->stream>>>(Q);
      return MAFE_OK;
    };
    int rc;
    switch ((nm + 15) / 16) {
      case 1: rc = launch(dct_mma_kernel<1>); break;
      case 2: rc = launch(dct_mma_kernel<2>); break;
      case 3: rc = launch(dct_mma_kernel<3>); break;
      case 4: rc = launch(dct_mma_kernel<4>); break;
      case 5: rc = launch(dct_mma_kernel<5>); break;
      default: rc = launch(dct_mma_kernel<6>); break;
    }
    if (rc) return rc;
    MAFE_LAUNCH_CHECK(ctx);
    return MAFE_OK;
  }
  const int cpt_need = (nc + 7) / 8;
  const int cpt = cpt_need <= 2 ? 2 : cpt_need <= 4 ? 4 : cpt_need <= 5 ? 5 : cpt_need <= 8 ? 8 : 16;
  const size_t smem = sizeof(float) * ((size_t)nm * 8 * cpt + (size_t)nm * kDctXStride);
  MAFE_REQUIRE(smem <= 200 * 1024, "DCT table %dx%d does not fit in shared memory", nm, nc);
  const bool clamp = p->d.log_kind == MAFE_LOG_DB && p->d.top_db >= 0.f && db_group != MAFE_DBGROUP_NONE;
  const int tpg = kDctFrames / p->tile_frames;
  const int n_groups = (b->n_tiles + tpg - 1) / tpg;
  const int grid = std::min(n_groups, 4 * ctx->sm_count);
  auto launch = [&](auto kern) -> int {
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 256, smem, ctx->stream>>>(logmel, out, nm, nc, p->dct_dev, b->tiles_dev, b->n_tiles, b->frame_offsets_dev,
                                           p->tile_frames, b->group_max_dev, b->utt_group_dev,
                                           clamp ? db_group : MAFE_DBGROUP_NONE, p->d.top_db);
    return MAFE_OK;
  };
  int rc;
  switch (cpt) {
    case 2: rc = launch(dct_kernel<2>); break;
    case 4: rc = launch(dct_kernel<4>); break;
    case 5: rc = launch(dct_kernel<5>); break;
    case 8: rc = launch(dct_kernel<8>); break;
    default: rc = launch(dct_kernel<16>); break;
  }
  if (rc) return rc;
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

}  // namespace mafe
