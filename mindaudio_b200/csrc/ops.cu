// ops.cu -- element-wise spectral ops, the CMVN family and the feature post-processing kernels.
// HBM-bound integer/float streaming work: coalesced, vectorised where alignment allows, grids
// sized in multiples of the SM count, warp/block reductions for the statistics.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace mafe {

static inline int grid_for(const mafe_ctx* ctx, int64_t work_items, int per_block, int max_waves = 32) {
  int64_t blocks = (work_items + per_block - 1) / per_block;
  int64_t cap = (int64_t)ctx->sm_count * max_waves;
  return (int)std::max<int64_t>(1, std::min(blocks, cap));
}

// ------------------------------------------------------------------ magphase (spectrum.py:720-732)
__global__ void magphase_kernel(const float2* __restrict__ z, int64_t n, float power, float* __restrict__ mag,
                                float2* __restrict__ phase) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float2 v = z[i];
    float m = hypotf(v.x, v.y);
    float zero = m == 0.f ? 1.f : 0.f;
    float den = m + zero;
    if (phase) phase[i] = make_float2(v.x / den + zero, v.y / den);
    if (mag) mag[i] = power == 1.f ? m : (power == 2.f ? m * m : powf(m, power));
  }
}

// ------------------------------------------------------------------ amplitude_to_dB (spectrum.py:59-90)
__global__ void to_db_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t group_size, int blocks_per_group,
                             float mult, float amin, float db_offset, int* group_max) {
  const int64_t g = blockIdx.x / blocks_per_group;
  const int chunk = blockIdx.x - (int)(g * blocks_per_group);
  const int64_t per = (group_size + blocks_per_group - 1) / blocks_per_group;
  const int64_t lo = chunk * per, hi = min(group_size, lo + per);
  const float* xs = x + g * group_size;
  float* os = out + g * group_size;
  float vmax = -INFINITY;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    float v = mult * log10f(fmaxf(xs[i], amin)) - db_offset;
    os[i] = v;
    vmax = fmaxf(vmax, v);
  }
  if (group_max) {
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if ((threadIdx.x & 31) == 0 && vmax > -INFINITY) atomicMax(&group_max[g], ordered_key(vmax));
  }
}

__global__ void clamp_group_kernel(float* __restrict__ out, int64_t group_size, int blocks_per_group, const int* group_max,
                                   float top_db) {
  const int64_t g = blockIdx.x / blocks_per_group;
  const int chunk = blockIdx.x - (int)(g * blocks_per_group);
  const int64_t per = (group_size + blocks_per_group - 1) / blocks_per_group;
  const int64_t lo = chunk * per, hi = min(group_size, lo + per);
  const float floor_v = key_to_float(group_max[g]) - top_db;
  float* os = out + g * group_size;
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) os[i] = fmaxf(os[i], floor_v);
}

__global__ void fill_i32_kernel(int* p, int64_t n, int v) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// ------------------------------------------------------------------ dB_to_amplitude (spectrum.py:108-113)
__global__ void from_db_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, float ref, float power) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = ref * powf(exp10f(0.1f * x[i]), power);
}

// ------------------------------------------------------------------ utterance CMVN (spec_augment.py:43-70)
// one CTA per utterance; thread (r, d) strides over frames r, r+rows, ...; double accumulators;
// the second pass re-reads the utterance's features (L2-resident: <= a few hundred KB).
constexpr int kCmvnThreads = 256;

__global__ void __launch_bounds__(kCmvnThreads) cmvn_utt_kernel(float* __restrict__ feats, const int64_t* __restrict__ fo,
                                                                int dim, int mean_norm, int std_norm) {
  __shared__ double sh[2 * kCmvnThreads];   // partial sums / sums of squares per thread
  __shared__ float s_mean[2 * kCmvnThreads]; // (mean, 1/std) per column of the current tile
  const int u = blockIdx.x;
  const int64_t f0 = fo[u];
  const int T = (int)(fo[u + 1] - f0);
  if (T <= 0) return;
  float* x = feats + f0 * dim;
  // column tiles of up to kCmvnThreads dims; rows = threads / min(dim, threads)
  for (int d0 = 0; d0 < dim; d0 += kCmvnThreads) {
    const int dw = min(dim - d0, kCmvnThreads);
    const int rows = kCmvnThreads / dw;
    const int r = threadIdx.x / dw, d = threadIdx.x - r * dw;
    double s1 = 0.0, s2 = 0.0;
    if (r < rows) {
      for (int f = r; f < T; f += rows) {
        double v = (double)x[(int64_t)f * dim + d0 + d];
        s1 += v;
        s2 += v * v;
      }
    }
    sh[threadIdx.x] = s1;
    sh[kCmvnThreads + threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < dw) {
      double a = 0.0, b = 0.0;
      for (int rr = 0; rr < rows; ++rr) { a += sh[rr * dw + threadIdx.x]; b += sh[kCmvnThreads + rr * dw + threadIdx.x]; }
      double mean = a / T;
      double var = b / T - mean * mean;  // population variance (np.std, ddof=0)
      if (var < 0.0) var = 0.0;
      float m = mean_norm ? (float)mean : 0.f;
      float inv = std_norm ? (float)(1.0 / sqrt(var)) : 1.f;  // NO eps: spec_augment.py:54 divides by the raw std
      // second pass for this column tile
      s_mean[2 * threadIdx.x] = m;
      s_mean[2 * threadIdx.x + 1] = inv;
    }
    __syncthreads();
    if (r < rows) {
      const float m = s_mean[2 * d], inv = s_mean[2 * d + 1];
      for (int f = r; f < T; f += rows) {
        int64_t i = (int64_t)f * dim + d0 + d;
        x[i] = (x[i] - m) * inv;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ scalar norm (deepspeech2/dataset.py:43-47)
__global__ void __launch_bounds__(kCmvnThreads) cmvn_scalar_kernel(float* __restrict__ feats, const int64_t* __restrict__ fo,
                                                                   int dim, int log1p_first) {
  __shared__ double s1s[kCmvnThreads / 32], s2s[kCmvnThreads / 32];
  __shared__ float s_m, s_inv;
  const int u = blockIdx.x;
  const int64_t n = (fo[u + 1] - fo[u]) * dim;
  if (n <= 0) return;
  float* x = feats + fo[u] * dim;
  double s1 = 0.0, s2 = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float v = x[i];
    if (log1p_first) v = log1pf(v);
    s1 += (double)v;
    s2 += (double)v * (double)v;
  }
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { s1s[threadIdx.x >> 5] = s1; s2s[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < kCmvnThreads / 32; ++w) { a += s1s[w]; b += s2s[w]; }
    double mean = a / (double)n, var = b / (double)n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_m = (float)mean;
    s_inv = (float)(1.0 / sqrt(var));
  }
  __syncthreads();
  const float m = s_m, inv = s_inv;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    float v = x[i];
    if (log1p_first) v = log1pf(v);
    x[i] = (v - m) * inv;
  }
}

// ------------------------------------------------------------------ global CMVN stats (compute_cmvn_stats.py:61-63)
__global__ void __launch_bounds__(kCmvnThreads) cmvn_stats_kernel(const float* __restrict__ feats, int64_t total_frames,
                                                                  int dim, double* __restrict__ stats) {
  // frames are split evenly over the grid; thread (r, d) as in cmvn_utt_kernel
  __shared__ double sh[2 * kCmvnThreads];
  const int64_t per = (total_frames + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * per, hi = min(total_frames, lo + per);
  for (int d0 = 0; d0 < dim; d0 += kCmvnThreads) {
    const int dw = min(dim - d0, kCmvnThreads);
    const int rows = kCmvnThreads / dw;
    const int r = threadIdx.x / dw, d = threadIdx.x - r * dw;
    double s1 = 0.0, s2 = 0.0;
    if (r < rows) {
      const float* p = feats + d0 + d;
#pragma unroll 8
      for (int64_t f = lo + r; f < hi; f += rows) {   // independent loads: eight in flight per thread
        double v = (double)__ldg(p + f * dim);
        s1 += v;
        s2 += v * v;
      }
    }
    sh[threadIdx.x] = s1;
    sh[kCmvnThreads + threadIdx.x] = s2;
    __syncthreads();
    if (threadIdx.x < dw) {
      double a = 0.0, b = 0.0;
      for (int rr = 0; rr < rows; ++rr) { a += sh[rr * dw + threadIdx.x]; b += sh[kCmvnThreads + rr * dw + threadIdx.x]; }
      atomicAdd(&stats[d0 + threadIdx.x], a);
      atomicAdd(&stats[dim + d0 + threadIdx.x], b);
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[2 * dim], (double)total_frames);
}

// ------------------------------------------------------------------ global CMVN apply (cmvn.py:33-36)
__global__ void cmvn_apply_kernel(float* __restrict__ feats, int64_t n, int dim, const float* __restrict__ mean,
                                  const float* __restrict__ istd) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int d = (int)(i % dim);
    float v = feats[i] - mean[d];
    if (istd) v *= istd[d];
    feats[i] = v;
  }
}
// dim % 4 == 0 and a 16-byte aligned matrix: 16 bytes per thread and step; the column of a thread's next group follows
// incrementally (the 64-bit modulo per element made the scalar kernel issue bound: 52 instructions per frame of 40)
__global__ void __launch_bounds__(256) cmvn_apply4_kernel(float4* __restrict__ feats, int64_t n4, int dim4, const float4* __restrict__ mean,
                                                         const float4* __restrict__ istd) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int d = (int)(i % dim4);
  const int step = (int)(stride % dim4);
#pragma unroll 4
  for (; i < n4; i += stride) {
    float4 v = feats[i];
    const float4 m = mean[d];
    v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
    if (istd) {
      const float4 s = istd[d];
      v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
    }
    feats[i] = v;
    d += step;
    if (d >= dim4) d -= dim4;
  }
}

// ------------------------------------------------------------------ compute_deltas (features.py:158-193, A8)
__global__ void deltas_kernel(const float* __restrict__ x, float* __restrict__ out, int n_mats, int rows, int t, int64_t xs,
                              int64_t os, int n, float inv_denom, int pad_mode) {
  const int64_t per = (int64_t)rows * t;
  const int64_t total = (int64_t)n_mats * per;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t mat = i / per, rem = i - mat * per;
    int64_t r = rem / t;
    int c = (int)(rem - r * t);
    const float* row = x + mat * xs + r * t;
    float acc = 0.f;
    for (int k = -n; k <= n; ++k) {
      int64_t j = pad_index(c + k, t, pad_mode);
      if (j >= 0) acc = fmaf((float)k, row[j], acc);
    }
    out[mat * os + rem] = acc * inv_denom;
  }
}

// ------------------------------------------------------------------ context_window (features.py:94-155 as a gather)
__global__ void context_kernel(const float* __restrict__ x, float* __restrict__ out, int n_mats, int f, int t, int csize, int ksz,
                               int mf, int roll) {
  const int64_t total = (int64_t)n_mats * f * csize * t;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int tt = (int)(i % t);
    int64_t q = i / t;
    int c = (int)(q % csize);
    int64_t fm = q / csize;  // mat * f + row
    int tap = (c + roll) % ksz;
    int src = tt + tap - mf;
    out[i] = (src >= 0 && src < t) ? x[fm * t + src] : 0.f;
  }
}

// ------------------------------------------------------------------ melscale (spectrum.py:738-774)
__global__ void melscale_kernel(const float* __restrict__ spec, float* __restrict__ out, int n_mats, int n_bins, int t, int n_mels,
                                const int* __restrict__ row_ptr, const int* __restrict__ col, const float* __restrict__ val) {
  const int64_t total = (int64_t)n_mats * n_mels * t;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int tt = (int)(i % t);
    int64_t q = i / t;
    int m = (int)(q % n_mels);
    int64_t mat = q / n_mels;
    const float* base = spec + mat * n_bins * (int64_t)t + tt;
    float acc = 0.f;
    for (int j = row_ptr[m]; j < row_ptr[m + 1]; ++j) acc = fmaf(val[j], base[(int64_t)col[j] * t], acc);
    out[i] = acc;
  }
}

// ------------------------------------------------------------------ transpose [rows][cols] -> [cols][rows]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols, int64_t os) {
  __shared__ float tile[32][33];
  const float* src = in + (int64_t)blockIdx.z * rows * cols;
  float* dst = out + (int64_t)blockIdx.z * os;
  int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int r = r0 + j, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[j][threadIdx.x] = src[(int64_t)r * cols + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) dst[(int64_t)c * rows + r] = tile[threadIdx.x][j];
  }
}

// ------------------------------------------------------------------ FP32 FMA peak microbenchmark
// 16 independent FFMA chains per thread: the roofline denominator SURVEY.md section 8d asks to measure.
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float a, float b) {
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += v[i];
  if (s == 123.456f) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace mafe

using namespace mafe;

// pad_sequence (mindaudio/utils/common.py:10-52): one thread per output element (or 4 when dim % 4 == 0)
template <int V>
__global__ void pad_sequence_kernel(const float* __restrict__ feats, const int64_t* __restrict__ fo, int n_utts, int dim, int max_len,
                                    float pad, int batch_first, float* __restrict__ out, float* __restrict__ mask) {
  const int dv = dim / V;
  const int64_t total = (int64_t)n_utts * max_len * dv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % dv);
    const int64_t r = i / dv;
    int b, t;
    if (batch_first) { b = (int)(r / max_len); t = (int)(r - (int64_t)b * max_len); }
    else { t = (int)(r / n_utts); b = (int)(r - (int64_t)t * n_utts); }
    const int64_t f0 = fo[b];
    const int len = (int)(fo[b + 1] - f0);
    const bool in = t < len;
    if (V == 4) {
      float4 v = make_float4(pad, pad, pad, pad);
      if (in) v = *reinterpret_cast<const float4*>(feats + (f0 + t) * dim + 4 * d);
      *reinterpret_cast<float4*>(out + r * dim + 4 * d) = v;
    } else {
      out[r * dim + d] = in ? feats[(f0 + t) * dim + d] : pad;
    }
    if (mask != nullptr && d == 0) mask[(int64_t)b * max_len + t] = in ? 1.f : 0.f;
  }
}

// sliding-window CMN: one thread per (channel, feature) walks the frames with running window sums in double.
// Window rule of Kaldi / torchaudio.functional.sliding_window_cmn (the window moves by at most one frame per step).
__global__ void sliding_cmn_kernel(const float* __restrict__ x, float* __restrict__ out, int n_ch, int T, int D, int cmn_window,
                                   int min_cmn_window, int center, int norm_vars) {
  const int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (id >= (int64_t)n_ch * D) return;
  const int c = (int)(id / D), d = (int)(id - (int64_t)c * D);
  const float* xc = x + (int64_t)c * T * D + d;
  float* oc = out + (int64_t)c * T * D + d;
  double sum = 0.0, sumsq = 0.0;
  int last_start = -1, last_end = -1;
  for (int t = 0; t < T; ++t) {
    int ws, we;
    if (center) { ws = t - cmn_window / 2; we = ws + cmn_window; }
    else { ws = t - cmn_window; we = t + 1; }
    if (ws < 0) { we -= ws; ws = 0; }
    if (!center && we > t) we = max(t + 1, min_cmn_window);
    if (we > T) { ws -= we - T; we = T; if (ws < 0) ws = 0; }
    if (last_start == -1) {
      for (int i = ws; i < we; ++i) { const double v = xc[(int64_t)i * D]; sum += v; sumsq += v * v; }
    } else {
      if (ws > last_start) { const double v = xc[(int64_t)last_start * D]; sum -= v; sumsq -= v * v; }
      if (we > last_end) { const double v = xc[(int64_t)last_end * D]; sum += v; sumsq += v * v; }
    }
    const int n = we - ws;
    last_start = ws; last_end = we;
    const double xv = xc[(int64_t)t * D];
    double y = xv - sum / n;
    if (norm_vars) {
      if (n == 1) y = 0.0;
      else y *= 1.0 / sqrt(sumsq / n - (sum * sum) / ((double)n * n));
    }
    oc[(int64_t)t * D] = (float)y;
  }
}

// rectangle masks: one CTA per rectangle
__global__ void mask_rects_kernel(float* __restrict__ feats, const int64_t* __restrict__ fo, int n_items, int dim,
                                  const int* __restrict__ rects, float value) {
  const int* r = rects + 5 * blockIdx.x;
  const int item = r[0];
  if (item < 0 || item >= n_items) return;
  const int64_t base = fo[item];
  const int rows = (int)(fo[item + 1] - base);
  const int r0 = max(r[1], 0), r1 = min(r[2], rows), c0 = max(r[3], 0), c1 = min(r[4], dim);
  const int w = c1 - c0;
  if (r1 <= r0 || w <= 0) return;
  const int64_t n = (int64_t)(r1 - r0) * w;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const int rr = (int)(i / w), cc = (int)(i - (int64_t)rr * w);
    feats[(base + r0 + rr) * dim + c0 + cc] = value;
  }
}

// phase vocoder (augment.py:828-871): one thread per (matrix, bin) walks the output steps; lanes = consecutive bins
__global__ void phase_vocoder_kernel(const float2* __restrict__ spec, int n_mats, int T, int F, double rate,
                                     const double* __restrict__ phi, int n_steps, float2* __restrict__ out) {
  const int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (id >= (int64_t)n_mats * F) return;
  const int mat = (int)(id / F), k = (int)(id - (int64_t)mat * F);
  const float2* sp = spec + (int64_t)mat * T * F + k;
  float2* op = out + (int64_t)mat * n_steps * F + k;
  const double phi_k = phi[k];
  const double two_pi = 6.283185307179586476925286766559;
  const float2 z0 = T > 0 ? sp[0] : make_float2(0.f, 0.f);
  float phase_acc = atan2f(z0.y, z0.x);
  for (int t = 0; t < n_steps; ++t) {
    const double step = (double)t * rate;           // np.arange(0, T, rate)[t]
    const int i0 = (int)step;
    const double alpha = step - floor(step);        // np.mod(step, 1.0), step >= 0
    const float2 c0 = i0 < T ? sp[(int64_t)i0 * F] : make_float2(0.f, 0.f);
    const float2 c1 = i0 + 1 < T ? sp[(int64_t)(i0 + 1) * F] : make_float2(0.f, 0.f);
    const float m0 = hypotf(c0.x, c0.y), m1 = hypotf(c1.x, c1.y);
    const float mag = (float)((1.0 - alpha) * (double)m0 + alpha * (double)m1);
    float sn, cs;
    sincosf(phase_acc, &sn, &cs);
    op[(int64_t)t * F] = make_float2(cs * mag, sn * mag);
    double dphase = (double)atan2f(c1.y, c1.x) - (double)atan2f(c0.y, c0.x) - phi_k;
    dphase = dphase - two_pi * rint(dphase / two_pi);
    phase_acc = (float)((double)phase_acc + (phi_k + dphase));
  }
}

// 1-D median filter along one axis (scipy.ndimage.median_filter, mode="reflect"): one thread per output element keeps
// the size/2 + 1 smallest window values in a sorted register/local array (partial insertion sort) -- rank size/2.
constexpr int kMedianMaxRank = 64;   // size <= 127
__global__ void median_filter_kernel(const float* __restrict__ x, float* __restrict__ out, int n_mats, int rows, int cols, int size,
                                     int axis) {
  const int64_t total = (int64_t)n_mats * rows * cols;
  const int len = axis == 0 ? rows : cols;
  const int64_t stride = axis == 0 ? cols : 1;
  const int rank = size / 2, keep = rank + 1;
  for (int64_t id = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; id < total; id += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(id % cols);
    const int r = (int)((id / cols) % rows);
    const int i = axis == 0 ? r : c;
    const float* line = x + (id - (int64_t)i * stride);
    float best[kMedianMaxRank];   // ascending, `filled` entries
    int filled = 0;
    for (int w = 0; w < size; ++w) {
      int j = i - rank + w;
      // reflect about the edges: d c b a | a b c d | d c b a (period 2 * len)
      if (j < 0 || j >= len) {
        const int p2 = 2 * len;
        j %= p2;
        if (j < 0) j += p2;
        if (j >= len) j = p2 - 1 - j;
      }
      const float v = line[(int64_t)j * stride];
      if (filled < keep) {
        int k = filled++;
        while (k > 0 && best[k - 1] > v) { best[k] = best[k - 1]; --k; }
        best[k] = v;
      } else if (v < best[keep - 1]) {
        int k = keep - 1;
        while (k > 0 && best[k - 1] > v) { best[k] = best[k - 1]; --k; }
        best[k] = v;
      }
    }
    out[id] = best[keep - 1];
  }
}

__device__ __forceinline__ float soft_mask_one(float x, float ref, float power, int split_zeros) {
  const float z0 = fmaxf(x, ref);
  const bool bad = z0 < 1.17549435e-38f;   // np.finfo(float32).tiny
  if (isinf(power)) return x > ref ? 1.f : 0.f;
  if (bad) return split_zeros ? 0.5f : 0.f;
  const float a = x / z0, b = ref / z0;
  const float m = power == 2.f ? a * a : (power == 1.f ? a : powf(a, power));
  const float rm = power == 2.f ? b * b : (power == 1.f ? b : powf(b, power));
  return m / (m + rm);
}
__global__ void hpss_masks_kernel(const float* __restrict__ harm, const float* __restrict__ perc, int64_t n, float margin_h,
                                  float margin_p, float power, int split_zeros, float* __restrict__ mh, float* __restrict__ mp) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float h = harm[i], p = perc[i];
    mh[i] = soft_mask_one(h, p * margin_h, power, split_zeros);
    mp[i] = soft_mask_one(p, h * margin_p, power, split_zeros);
  }
}

extern "C" {

int mafe_magphase(mafe_ctx* ctx, const float* z, int64_t n, float power, float* mag, float* phase) {
  MAFE_REQUIRE(ctx && (z || n == 0), "mafe_magphase: NULL argument");
  MAFE_REQUIRE(power >= 0.f, "power must be non-negative");
  if (n == 0) return MAFE_OK;
  cudaSetDevice(ctx->device);
  magphase_kernel<<<grid_for(ctx, n, 256 * 4), 256, 0, ctx->stream>>>((const float2*)z, n, power, mag, (float2*)phase);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_amplitude_to_db(mafe_ctx* ctx, const float* x, float* out, int64_t n_groups, int64_t group_size, float mult,
                         float amin, float db_offset, float top_db) {
  MAFE_REQUIRE(ctx && ((x && out) || n_groups * group_size == 0), "mafe_amplitude_to_db: NULL argument");
  if (n_groups <= 0 || group_size <= 0) return MAFE_OK;
  cudaSetDevice(ctx->device);
  // enough blocks per group to fill the machine, chunks of >= 4096 elements
  int64_t want = std::max<int64_t>(1, ((int64_t)ctx->sm_count * 8 + n_groups - 1) / n_groups);
  int bpg = (int)std::max<int64_t>(1, std::min<int64_t>(want, (group_size + 4095) / 4096));
  MAFE_REQUIRE(n_groups * bpg < (int64_t)INT32_MAX, "too many groups");
  int* gmax = nullptr;
  if (top_db >= 0.f) {
    MAFE_CUDA_CHECK(cudaMallocAsync((void**)&gmax, n_groups * sizeof(int), ctx->stream));
    fill_i32_kernel<<<(int)((n_groups + 255) / 256), 256, 0, ctx->stream>>>(gmax, n_groups, (int)0x80000000);
    MAFE_LAUNCH_CHECK(ctx);
  }
  to_db_kernel<<<(int)(n_groups * bpg), 256, 0, ctx->stream>>>(x, out, group_size, bpg, mult, amin, db_offset, gmax);
  MAFE_LAUNCH_CHECK(ctx);
  if (gmax) {
    clamp_group_kernel<<<(int)(n_groups * bpg), 256, 0, ctx->stream>>>(out, group_size, bpg, gmax, top_db);
    MAFE_LAUNCH_CHECK(ctx);
    MAFE_CUDA_CHECK(cudaFreeAsync(gmax, ctx->stream));
  }
  return MAFE_OK;
}

int mafe_db_to_amplitude(mafe_ctx* ctx, const float* x, float* out, int64_t n, float ref, float power) {
  MAFE_REQUIRE(ctx && ((x && out) || n == 0), "mafe_db_to_amplitude: NULL argument");
  if (n == 0) return MAFE_OK;
  cudaSetDevice(ctx->device);
  from_db_kernel<<<grid_for(ctx, n, 256 * 4), 256, 0, ctx->stream>>>(x, out, n, ref, power);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_cmvn_utt(mafe_ctx* ctx, float* feats, const int64_t* fo, int32_t n_utts, int32_t dim, int32_t mean_norm,
                  int32_t std_norm) {
  MAFE_REQUIRE(ctx && fo && (feats || n_utts == 0), "mafe_cmvn_utt: NULL argument");
  MAFE_REQUIRE(dim >= 1, "dim=%d", dim);
  if (n_utts <= 0) return MAFE_OK;
  cudaSetDevice(ctx->device);
  ProfScope ps(ctx, MAFE_PROF_CMVN);
  cmvn_utt_kernel<<<n_utts, kCmvnThreads, 0, ctx->stream>>>(feats, fo, dim, mean_norm, std_norm);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_cmvn_scalar(mafe_ctx* ctx, float* feats, const int64_t* fo, int32_t n_utts, int32_t dim, int32_t log1p_first) {
  MAFE_REQUIRE(ctx && fo && (feats || n_utts == 0), "mafe_cmvn_scalar: NULL argument");
  if (n_utts <= 0) return MAFE_OK;
  cudaSetDevice(ctx->device);
  cmvn_scalar_kernel<<<n_utts, kCmvnThreads, 0, ctx->stream>>>(feats, fo, dim, log1p_first);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_cmvn_stats_accumulate(mafe_ctx* ctx, const float* feats, int64_t total_frames, int32_t dim, double* stats) {
  MAFE_REQUIRE(ctx && stats && (feats || total_frames == 0), "mafe_cmvn_stats_accumulate: NULL argument");
  if (total_frames <= 0) return MAFE_OK;
  cudaSetDevice(ctx->device);
  int grid = (int)std::min<int64_t>((int64_t)ctx->sm_count * 8, std::max<int64_t>(1, total_frames / 64));
  ProfScope ps(ctx, MAFE_PROF_CMVN);
  cmvn_stats_kernel<<<grid, kCmvnThreads, 0, ctx->stream>>>(feats, total_frames, dim, stats);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_cmvn_apply(mafe_ctx* ctx, float* feats, int64_t total_frames, int32_t dim, const float* mean, const float* istd) {
  MAFE_REQUIRE(ctx && mean && (feats || total_frames == 0), "mafe_cmvn_apply: NULL argument");
  if (total_frames <= 0) return MAFE_OK;
  cudaSetDevice(ctx->device);
  int64_t n = total_frames * dim;
  if ((dim & 3) == 0 && ((uintptr_t)feats & 15) == 0 && ((uintptr_t)mean & 15) == 0 && (istd == nullptr || ((uintptr_t)istd & 15) == 0))
    cmvn_apply4_kernel<<<grid_for(ctx, n / 4, 256 * 4), 256, 0, ctx->stream>>>(reinterpret_cast<float4*>(feats), n / 4, dim / 4,
                                                                               reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(istd));
  else
    cmvn_apply_kernel<<<grid_for(ctx, n, 256 * 4), 256, 0, ctx->stream>>>(feats, n, dim, mean, istd);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_melscale(mafe_ctx* ctx, const float* spec, float* out, int32_t n_mats, int32_t n_bins, int32_t t, const float* fb,
                  int32_t n_mels) {
  MAFE_REQUIRE(ctx && fb, "mafe_melscale: NULL argument");
  MAFE_REQUIRE(n_bins >= 1 && n_mels >= 1, "bad filterbank shape %dx%d", n_mels, n_bins);
  if (n_mats <= 0 || t <= 0) return MAFE_OK;
  MAFE_REQUIRE(spec && out, "mafe_melscale: NULL buffer");
  cudaSetDevice(ctx->device);
  std::vector<int> row_ptr(n_mels + 1, 0), col;
  std::vector<float> val;
  for (int m = 0; m < n_mels; ++m) {
    for (int k = 0; k < n_bins; ++k)
      if (fb[(size_t)m * n_bins + k] != 0.f) { col.push_back(k); val.push_back(fb[(size_t)m * n_bins + k]); }
    row_ptr[m + 1] = (int)col.size();
  }
  if (col.empty()) { col.push_back(0); val.push_back(0.f); }
  int *d_rp = nullptr, *d_col = nullptr;
  float* d_val = nullptr;
  cudaStream_t st = ctx->stream;
  MAFE_CUDA_CHECK(cudaMallocAsync((void**)&d_rp, row_ptr.size() * 4, st));
  MAFE_CUDA_CHECK(cudaMallocAsync((void**)&d_col, col.size() * 4, st));
  MAFE_CUDA_CHECK(cudaMallocAsync((void**)&d_val, val.size() * 4, st));
  MAFE_CUDA_CHECK(cudaMemcpyAsync(d_rp, row_ptr.data(), row_ptr.size() * 4, cudaMemcpyHostToDevice, st));
  MAFE_CUDA_CHECK(cudaMemcpyAsync(d_col, col.data(), col.size() * 4, cudaMemcpyHostToDevice, st));
  MAFE_CUDA_CHECK(cudaMemcpyAsync(d_val, val.data(), val.size() * 4, cudaMemcpyHostToDevice, st));
  int64_t total = (int64_t)n_mats * n_mels * t;
  melscale_kernel<<<grid_for(ctx, total, 256), 256, 0, st>>>(spec, out, n_mats, n_bins, t, n_mels, d_rp, d_col, d_val);
  MAFE_LAUNCH_CHECK(ctx);
  MAFE_CUDA_CHECK(cudaStreamSynchronize(st));  // host CSR vectors must outlive the copies
  MAFE_CUDA_CHECK(cudaFreeAsync(d_rp, st));
  MAFE_CUDA_CHECK(cudaFreeAsync(d_col, st));
  MAFE_CUDA_CHECK(cudaFreeAsync(d_val, st));
  return MAFE_OK;
}

int mafe_transpose(mafe_ctx* ctx, const float* in, float* out, int32_t n_mats, int32_t rows, int32_t cols,
                   int64_t out_mat_stride) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  if (n_mats <= 0 || rows <= 0 || cols <= 0) return MAFE_OK;
  MAFE_REQUIRE(in && out, "mafe_transpose: NULL buffer");
  MAFE_REQUIRE(n_mats <= 65535 && (rows + 31) / 32 <= 65535, "mafe_transpose: shape too large");
  cudaSetDevice(ctx->device);
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, n_mats), block(32, 8);
  if (out_mat_stride == 0) out_mat_stride = (int64_t)rows * cols;
  MAFE_REQUIRE(out_mat_stride >= (int64_t)rows * cols, "out_mat_stride too small");
  transpose_kernel<<<grid, block, 0, ctx->stream>>>(in, out, rows, cols, out_mat_stride);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_fp32_fma_peak(mafe_ctx* ctx, double* tflops_out) {
  MAFE_REQUIRE(ctx && tflops_out, "mafe_fp32_fma_peak: NULL argument");
  cudaSetDevice(ctx->device);
  float* d = nullptr;
  MAFE_CUDA_CHECK(cudaMalloc((void**)&d, 16));
  const int iters = 4096, blocks = ctx->sm_count * 16;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0, ctx->stream);
    fma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(e1, ctx->stream);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 16.0 * iters * 256.0 * blocks;
    if (rep > 0 && ms > 0.f) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  ctx->launches += 5;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  MAFE_CUDA_CHECK(cudaGetLastError());
  *tflops_out = best;
  return MAFE_OK;
}

int mafe_compute_deltas(mafe_ctx* ctx, const float* x, float* out, int32_t n_mats, int32_t rows, int32_t t, int64_t xs,
                        int64_t os, int32_t win_length, int32_t pad_mode) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  MAFE_REQUIRE(win_length >= 3, "win_length must be no less than 3");
  MAFE_REQUIRE(pad_mode >= MAFE_PAD_CONSTANT && pad_mode <= MAFE_PAD_SYMMETRIC, "bad pad_mode %d", pad_mode);
  if (n_mats <= 0 || rows <= 0 || t <= 0) return MAFE_OK;
  MAFE_REQUIRE(x && out, "mafe_compute_deltas: NULL buffer");
  cudaSetDevice(ctx->device);
  if (xs == 0) xs = (int64_t)rows * t;
  if (os == 0) os = (int64_t)rows * t;
  int n = (win_length - 1) / 2;
  float inv = (float)(1.0 / ((double)n * (n + 1) * (2 * n + 1) / 3.0));
  deltas_kernel<<<grid_for(ctx, (int64_t)n_mats * rows * t, 256 * 2), 256, 0, ctx->stream>>>(x, out, n_mats, rows, t, xs, os, n,
                                                                                            inv, pad_mode);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_context_window(mafe_ctx* ctx, const float* x, float* out, int32_t n_mats, int32_t f, int32_t t, int32_t left,
                        int32_t right) {
  MAFE_REQUIRE(ctx && ((x && out) || n_mats == 0), "mafe_context_window: NULL argument");
  MAFE_REQUIRE(left >= 0 && right >= 0, "negative context");
  if (n_mats <= 0 || f <= 0 || t <= 0) return MAFE_OK;
  cudaSetDevice(ctx->device);
  int csize = left + right + 1, mf = std::max(left, right), ksz = 2 * mf + 1;
  int roll = right > left ? right - left : 0;
  int64_t total = (int64_t)n_mats * f * csize * t;
  context_kernel<<<grid_for(ctx, total, 256 * 2), 256, 0, ctx->stream>>>(x, out, n_mats, f, t, csize, ksz, mf, roll);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_pad_sequence(mafe_ctx* ctx, const float* feats, const int64_t* frame_offsets, int32_t n_utts, int32_t dim, int32_t max_len,
                      float padding_value, int32_t batch_first, float* out, float* mask) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  MAFE_REQUIRE(n_utts >= 0 && dim > 0 && max_len >= 0, "mafe_pad_sequence: bad shape (%d utterances, dim %d, max_len %d)", n_utts, dim, max_len);
  if (n_utts == 0 || max_len == 0) return MAFE_OK;
  MAFE_REQUIRE(frame_offsets && out, "mafe_pad_sequence: NULL buffer");
  cudaSetDevice(ctx->device);
  const bool vec = dim % 4 == 0 && ((uintptr_t)feats & 15) == 0 && ((uintptr_t)out & 15) == 0;
  const int64_t total = (int64_t)n_utts * max_len * (vec ? dim / 4 : dim);
  if (vec)
    pad_sequence_kernel<4><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(feats, frame_offsets, n_utts, dim, max_len, padding_value,
                                                                               batch_first, out, mask);
  else
    pad_sequence_kernel<1><<<grid_for(ctx, total, 256), 256, 0, ctx->stream>>>(feats, frame_offsets, n_utts, dim, max_len, padding_value,
                                                                               batch_first, out, mask);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_sliding_window_cmn(mafe_ctx* ctx, const float* x, float* out, int32_t n_channels, int32_t num_frames, int32_t num_feats,
                            int32_t cmn_window, int32_t min_cmn_window, int32_t center, int32_t norm_vars) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  MAFE_REQUIRE(cmn_window >= 1 && min_cmn_window >= 0, "sliding_window_cmn: cmn_window %d / min_cmn_window %d out of range", cmn_window,
               min_cmn_window);
  if (n_channels <= 0 || num_frames <= 0 || num_feats <= 0) return MAFE_OK;
  MAFE_REQUIRE(x && out, "mafe_sliding_window_cmn: NULL buffer");
  MAFE_REQUIRE(x != out, "mafe_sliding_window_cmn: out must not alias x");
  cudaSetDevice(ctx->device);
  const int64_t n = (int64_t)n_channels * num_feats;
  sliding_cmn_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(x, out, n_channels, num_frames, num_feats, cmn_window,
                                                                          min_cmn_window, center, norm_vars);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_mask_rects(mafe_ctx* ctx, float* feats, const int64_t* frame_offsets, int32_t n_items, int32_t dim, const int32_t* rects,
                    int32_t n_rects, float value) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  MAFE_REQUIRE(n_items >= 0 && dim > 0 && n_rects >= 0, "mafe_mask_rects: bad shape");
  if (n_rects == 0 || n_items == 0) return MAFE_OK;
  MAFE_REQUIRE(feats && frame_offsets && rects, "mafe_mask_rects: NULL buffer");
  cudaSetDevice(ctx->device);
  mask_rects_kernel<<<n_rects, 256, 0, ctx->stream>>>(feats, frame_offsets, n_items, dim, rects, value);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_phase_vocoder(mafe_ctx* ctx, const float* spec, int32_t n_mats, int32_t n_frames, int32_t n_bins, double rate,
                       const double* phi_advance, int32_t n_steps, float* out) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  MAFE_REQUIRE(rate > 0.0, "rate must be a positive number");
  MAFE_REQUIRE(n_mats >= 0 && n_frames >= 0 && n_bins > 0 && n_steps >= 0, "mafe_phase_vocoder: bad shape");
  if (n_mats == 0 || n_steps == 0) return MAFE_OK;
  MAFE_REQUIRE(spec && phi_advance && out, "mafe_phase_vocoder: NULL buffer");
  cudaSetDevice(ctx->device);
  const int64_t n = (int64_t)n_mats * n_bins;
  phase_vocoder_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>((const float2*)spec, n_mats, n_frames, n_bins, rate,
                                                                           phi_advance, n_steps, (float2*)out);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_median_filter(mafe_ctx* ctx, const float* x, float* out, int32_t n_mats, int32_t rows, int32_t cols, int32_t size,
                       int32_t axis) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  MAFE_REQUIRE(size >= 1 && size / 2 < kMedianMaxRank, "median_filter: size %d out of range (1..%d)", size, 2 * kMedianMaxRank - 1);
  MAFE_REQUIRE(axis == 0 || axis == 1, "median_filter: axis must be 0 or 1");
  if (n_mats <= 0 || rows <= 0 || cols <= 0) return MAFE_OK;
  MAFE_REQUIRE(x && out && x != out, "mafe_median_filter: NULL or aliased buffer");
  cudaSetDevice(ctx->device);
  const int64_t total = (int64_t)n_mats * rows * cols;
  median_filter_kernel<<<grid_for(ctx, total, 128), 128, 0, ctx->stream>>>(x, out, n_mats, rows, cols, size, axis);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

int mafe_hpss_masks(mafe_ctx* ctx, const float* harm, const float* perc, int64_t n, float margin_h, float margin_p, float power,
                    int32_t split_zeros, float* mask_h, float* mask_p) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  MAFE_REQUIRE(power > 0.f, "power must be strictly positive.");
  if (n <= 0) return MAFE_OK;
  MAFE_REQUIRE(harm && perc && mask_h && mask_p, "mafe_hpss_masks: NULL buffer");
  cudaSetDevice(ctx->device);
  hpss_masks_kernel<<<grid_for(ctx, n, 256 * 2), 256, 0, ctx->stream>>>(harm, perc, n, margin_h, margin_p, power, split_zeros, mask_h,
                                                                       mask_p);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}

}  // extern "C"
