// fbank512.cu -- the specialised sm_100a kernel of the headline path: Kaldi-like log-mel filterbank
// features (examples/conformer/dataset.py:117-168) with a 512-point FFT, fused end to end:
//
//   waveform tile -> [dither] -> pre-emphasis -> shared memory            (phase L)
//   povey window, scalar frame-mean removal, radix-2 fold                 (phase F, registers)
//   2 x 256-point FFT per frame PAIR, 16 lanes x 16 points, radix-16^2    (phase F)
//   pair separation + |X|^2 + sparse mel (<= 2 filters per bin)           (phase S, lanes = frames)
//   combine partial sums + ln + coalesced store                           (phase C)
//
// One CTA = 256 threads = one tile of 32 consecutive frames (16 frame pairs) of one utterance;
// ragged batches come in through the tile table.  Two real frames (a, b) ride in one complex
// sequence a + i*b; even and odd output bins come from two 256-point transforms that share one
// shared-memory slot per pair (see fft512.cuh).  Nothing but the waveform is read from and
// nothing but the features is written to HBM.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>
#include <type_traits>

#include "common.cuh"
#include "fft512.cuh"

namespace mafe {

constexpr int kFastThreads = 256;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kTileFrames = 32;
constexpr int kPairs = kTileFrames / 2;
constexpr int kNfft = 512;
constexpr int kBins = kNfft / 2 + 1;  // 257
constexpr int kMaxY = 5632;           // floats of waveform per tile (31*hop + frame_len <= kMaxY)
constexpr int kPlaneStride = 33;      // [mel][frame] planes padded against bank conflicts

struct BinEntry {   // per FFT bin: contributes w0 to filter f0 and w1 to filter f0+1 (weights pre-scaled by 1/4)
  int f0;
  float w0, w1;
  int pad;
};

struct FastTablesDev {
  float* window;      // [512] analysis window, zero beyond frame_len
  float2* w512;       // [256] W512^n
  float2* w256t;      // [16][16] W256^(t*kj) stored [kj][t]
  BinEntry* bins;     // [257]
  int2* warp_range;   // [8] filters (lo, hi) a warp emits during a sweep
  int* combine;       // [n_mels] bit0-1: number of contributing warps (0..2), bit2: plane of the first
  float* cover;       // [hop] interior coverage weights c(s mod hop) for the frame-mean pre-pass
};

struct FastParams {
  const void* wave;
  int wave_dtype;
  float wave_scale;
  const int64_t* sample_offsets;
  const int64_t* frame_offsets;
  const Tile* tiles;
  const double* utt_sum;
  int frame_len, hop, n_mels, ylen;
  float pre_hi, pre_lo;
  int preemph_on, remove_mean;
  float dither;
  uint64_t seed;
  int log_kind;
  float log_arg;
  FastTablesDev tab;
  float* out;
};

// ---- shared memory carve-up (bytes) ----
struct FastSmem {
  static constexpr size_t kZ = sizeof(float2) * kPairs * kSlotStride;  // 34944
  static constexpr size_t kY = sizeof(float) * kMaxY;                  // 22528 (also the mel planes)
  static constexpr size_t kWin = sizeof(float) * kNfft;
  static constexpr size_t kW512 = sizeof(float2) * 256;
  static constexpr size_t kW256 = sizeof(float2) * 256;
  static constexpr size_t kBinsB = sizeof(BinEntry) * 264;
  static constexpr size_t kTotal = kZ + kY + kWin + kW512 + kW256 + kBinsB;
};

__device__ __forceinline__ float fast_load_sample(const FastParams& P, int64_t g) {
  return (P.wave_dtype == MAFE_WAVE_I16 ? (float)((const int16_t*)P.wave)[g] : ((const float*)P.wave)[g]) * P.wave_scale;
}

// One half of the pair's spectrum: 256-point FFT of v (HALF = 0: even bins, 1: odd bins) by the
// 16-lane group, result left in the pair's slot; then the block-wide sweep of those bins
// (lanes = frames) into the mel planes.  Must be called by all threads of the CTA.
template <int HALF>
__device__ __forceinline__ void fft_half_and_sweep(cpx (&v)[16], float2* Zs, float2* slot, const float2* s_w256,
                                                   const BinEntry* s_bins, float* my_plane, int2 wr, int n_mels, int t,
                                                   int lane, int warp) {
  constexpr int half = HALF;
  {
    // radix-16 over j, twiddle W256^(t*kj), transpose through the slot
    fft16(v);
#pragma unroll
    for (int kj = 0; kj < 16; ++kj) {
      cpx x = v[fft16_pos(kj)];
      if (kj > 0) {
        const float2 tw = s_w256[kj * 16 + t];
        x = cmulf(x, cx(tw.x, tw.y));
      }
      slot[kj * kRowStride + t] = make_float2(x.x, x.y);
    }
    __syncwarp();
    cpx u[16];
#pragma unroll
    for (int tt = 0; tt < 16; ++tt) {
      const float2 x = slot[t * kRowStride + tt];
      u[tt] = cx(x.x, x.y);
    }
    __syncwarp();
    fft16(u);
#pragma unroll
    for (int kt = 0; kt < 16; ++kt) {
      const cpx x = u[fft16_pos(kt)];
      slot[t + 16 * kt] = make_float2(x.x, x.y);  // bin 2*(t+16kt)+half of this pair's spectrum
    }
    __syncthreads();

    // ---- phase S: lanes = frames, this warp sweeps FFT bins k = 2*kk + half, kk in its range ----
    {
      const float2* zp = Zs + (lane >> 1) * kSlotStride;
      const float sgn = (lane & 1) ? -1.f : 1.f;  // frame a: Z[k] + conj Z[N-k];  frame b: Z[k] - conj Z[N-k]
      int cur = wr.x;
      float acc_lo = 0.f, acc_hi = 0.f;
      const int kk_begin = warp * 16;
      const int kk_end = kk_begin + 16 + ((half == 0 && warp == kFastWarps - 1) ? 1 : 0);  // + Nyquist bin 256
#pragma unroll 4
      for (int kk = kk_begin; kk < kk_end; ++kk) {
        const int k = 2 * kk + half;
        const int kr = half == 0 ? ((256 - kk) & 255) : (255 - kk);   // slot index of bin 512-k
        const float2 zk = zp[kk & 255];
        const float2 zn = zp[kr];
        const float re = fmaf(sgn, zn.x, zk.x);
        const float im = fmaf(-sgn, zn.y, zk.y);
        const float pw = fmaf(re, re, im * im);
        const BinEntry be = s_bins[k];
        while (cur < be.f0) {  // warp-uniform: k is the same for all lanes
          if (cur >= 0 && cur < n_mels) {
            float* dst = my_plane + cur * kPlaneStride + lane;
            *dst = half == 0 ? acc_lo : *dst + acc_lo;
          }
          acc_lo = acc_hi; acc_hi = 0.f; ++cur;
        }
        acc_lo = fmaf(be.w0, pw, acc_lo);
        acc_hi = fmaf(be.w1, pw, acc_hi);
      }
      while (cur <= wr.y) {
        if (cur >= 0 && cur < n_mels) {
          float* dst = my_plane + cur * kPlaneStride + lane;
          *dst = half == 0 ? acc_lo : *dst + acc_lo;
        }
        acc_lo = acc_hi; acc_hi = 0.f; ++cur;
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kFastThreads, 2) fbank512_kernel(FastParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* Zs = reinterpret_cast<float2*>(smem_raw);
  float* ybuf = reinterpret_cast<float*>(smem_raw + FastSmem::kZ);
  float* planes = ybuf;  // aliased: the waveform is dead once every thread holds its points in registers
  float* s_win = reinterpret_cast<float*>(smem_raw + FastSmem::kZ + FastSmem::kY);
  float2* s_w512 = reinterpret_cast<float2*>(smem_raw + FastSmem::kZ + FastSmem::kY + FastSmem::kWin);
  float2* s_w256 = s_w512 + 256;
  BinEntry* s_bins = reinterpret_cast<BinEntry*>(s_w256 + 256);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const Tile tile = P.tiles[blockIdx.x];
  const uint32_t utt = (uint32_t)tile.utt;
  const int64_t off = P.sample_offsets[utt];
  const int64_t L = P.sample_offsets[utt + 1] - off;
  const int64_t fo = P.frame_offsets[utt];
  const int T = (int)(P.frame_offsets[utt + 1] - fo);
  const int frame0 = tile.frame0;
  const int nf = min(kTileFrames, T - frame0);

  // ---- tables -> shared ----
  for (int i = tid; i < kNfft; i += kFastThreads) s_win[i] = P.tab.window[i];
  for (int i = tid; i < 256; i += kFastThreads) { s_w512[i] = P.tab.w512[i]; s_w256[i] = P.tab.w256t[i]; }
  for (int i = tid; i < kBins; i += kFastThreads) s_bins[i] = P.tab.bins[i];

  // ---- phase L: waveform tile, dither, pre-emphasis ----
  {
    const int64_t s0 = (int64_t)frame0 * P.hop;
    const int need = (nf - 1) * P.hop + P.frame_len;  // samples actually covered by this tile's frames
    for (int i = tid; i < P.ylen; i += kFastThreads) {
      const int64_t s = s0 + i;
      float v = 0.f;
      if (i < need && s < L) {
        v = fast_load_sample(P, off + s);
        if (P.dither != 0.f) v = fmaf(P.dither, dither_normal((uint64_t)s, utt, P.seed), v);
        if (P.preemph_on) {
          // x[s-1]: neighbouring lane, except lane 0 (re-read) -- only when no dither (dither needs its own g)
          float vp = 0.f;
          if (s > 0) {
            vp = fast_load_sample(P, off + s - 1);
            if (P.dither != 0.f) vp = fmaf(P.dither, dither_normal((uint64_t)(s - 1), utt, P.seed), vp);
            v = fmaf(-P.pre_lo, vp, fmaf(-P.pre_hi, vp, v));
          }
        }
      }
      ybuf[i] = v;
    }
  }
  float neg_mu = 0.f;
  if (P.remove_mean) neg_mu = -(float)(P.utt_sum[utt] / ((double)T * (double)P.frame_len));
  __syncthreads();

  // ---- phase F: every 16-lane group transforms one frame pair ----
  const int t = lane & 15;
  const int pair = warp * 2 + (lane >> 4);
  float2* slot = Zs + pair * kSlotStride;
  cpx v0[16], v1[16];
  {
    const float* ya = ybuf + (2 * pair) * P.hop;
    const float* yb = ya + P.hop;
    const int flen = P.frame_len;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = t + 16 * j;
      const float w = s_win[n];
      // frames shorter than 512 are zero-padded: entries n >= frame_len stay 0 (no mean removal there)
      cpx lo = cx(0.f, 0.f), hi = cx(0.f, 0.f);
      if (n < flen) lo = cx(fmaf(ya[n], w, neg_mu), fmaf(yb[n], w, neg_mu));
      const int n2 = n + 256;
      if (n2 < flen) {
        const float w2 = s_win[n2];
        hi = cx(fmaf(ya[n2], w2, neg_mu), fmaf(yb[n2], w2, neg_mu));
      }
      v0[j] = lo + hi;
      const float2 tw = s_w512[n];
      v1[j] = cmulf(lo - hi, cx(tw.x, tw.y));
    }
  }
  __syncthreads();  // ybuf is dead from here on (it becomes the mel planes)

  const int2 wr = P.tab.warp_range[warp];
  float* my_plane = planes + (warp & 1) * (P.n_mels * kPlaneStride);

  fft_half_and_sweep<0>(v0, Zs, slot, s_w256, s_bins, my_plane, wr, P.n_mels, t, lane, warp);
  fft_half_and_sweep<1>(v1, Zs, slot, s_w256, s_bins, my_plane, wr, P.n_mels, t, lane, warp);

  // ---- phase C: combine the (<= 2) partial sums of every (frame, filter), log, coalesced store ----
  {
    const int nm = P.n_mels;
    float* dst = P.out + (fo + frame0) * (int64_t)nm;
    const int total = nf * nm;
    for (int e = tid; e < total; e += kFastThreads) {
      const int f = e / nm, m = e - f * nm;
      const int c = P.tab.combine[m];
      const int n = c & 3, p0 = (c >> 2) & 1;
      float acc = 0.f;
      if (n >= 1) acc = planes[p0 * (nm * kPlaneStride) + m * kPlaneStride + f];
      if (n == 2) acc += planes[(p0 ^ 1) * (nm * kPlaneStride) + m * kPlaneStride + f];
      float o;
      switch (P.log_kind) {
        case MAFE_LOG_LN_EPS_IF_ZERO: o = logf(acc == 0.f ? 2.220446049250313e-16f : acc); break;
        case MAFE_LOG_LN_PLUS: o = logf(acc + P.log_arg); break;
        default: o = acc; break;
      }
      dst[e] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// fast pre-pass: sum over all windowed frame entries of an utterance = sum_s y[s] * c(s), with
// c(s) = sum of the window over the frames covering sample s (no frame materialisation).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) frame_sum_fast_kernel(FastParams P, double* utt_sum) {
  const Tile tile = P.tiles[blockIdx.x];
  const uint32_t utt = (uint32_t)tile.utt;
  const int64_t off = P.sample_offsets[utt];
  const int T = (int)(P.frame_offsets[utt + 1] - P.frame_offsets[utt]);
  const int hop = P.hop, flen = P.frame_len;
  // this tile owns samples [frame0*hop, (frame0+32)*hop), clipped to the framed region of the utterance
  const int64_t s_lo = (int64_t)tile.frame0 * hop;
  const int64_t framed_end = (int64_t)(T - 1) * hop + flen;
  // the utterance's last tile also owns the tail [.., framed_end) that its last frame reaches into
  const int64_t s_hi = tile.frame0 + kTileFrames >= T ? framed_end : s_lo + (int64_t)kTileFrames * hop;
  double acc = 0.0;
  for (int64_t s = s_lo + threadIdx.x; s < s_hi; s += blockDim.x) {
    float v = fast_load_sample(P, off + s);
    if (P.dither != 0.f) v = fmaf(P.dither, dither_normal((uint64_t)s, utt, P.seed), v);
    if (P.preemph_on && s > 0) {
      float vp = fast_load_sample(P, off + s - 1);
      if (P.dither != 0.f) vp = fmaf(P.dither, dither_normal((uint64_t)(s - 1), utt, P.seed), vp);
      v = fmaf(-P.pre_lo, vp, fmaf(-P.pre_hi, vp, v));
    }
    // frames t with 0 <= s - t*hop < flen, 0 <= t < T
    const int t_hi = (int)min((int64_t)T - 1, s / hop);
    float c = 0.f;
    for (int tt = t_hi; tt >= 0; --tt) {
      const int64_t n = s - (int64_t)tt * hop;
      if (n >= flen) break;
      c += __ldg(&P.tab.window[n]);
    }
    acc += (double)v * (double)c;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += ws[w];
    atomicAdd(&utt_sum[utt], s);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
}  // namespace mafe
#include "fbank512_tile.cuh"
#include "fbank512_v3.cuh"
#include "fbank512_v6.cuh"
#include "stft512.cuh"
#include "fbank400.cuh"
#include "stftn16.cuh"
#include "front2048.cuh"
namespace mafe {

struct FastTablesHost {
  FastTablesDev dev;
  int ylen;
  bool stft = false;      // n_fft = 512 complex STFT plan (stft512_kernel)
  int stftn = 0;          // n_fft = 320 / 400 complex STFT plan (stftn16_kernel<20 / 25>): N1, else 0
  float2 tws[16];         // W_N1^(j1 k1) of the in-register N1-point DFT
  bool f400 = false;      // n_fft = 400 mel front-end plan (fbank400_kernel)
  float2* tw400_dev = nullptr;
  float* cover4_dev = nullptr;   // v6 FS: c' table
  F400Sweep sweep400;     // sweep program of fbank400_kernel (kernel-parameter bank)
  float2 tw25[16];
  bool tile_geom = false; // conformer geometry (400 / 160 / 512 / 80 filters): the v3 kernel and its pre-pass apply
  V3Sweep sweep;          // sweep program of the v3 kernel (kernel-parameter bank)
  bool v3 = false;        // the v3 kernel's compact planes can hold this filterbank
  int* comb3_dev = nullptr;   // v3: plane rows (A | B << 8) of every filter
  bool f2048 = false;         // n_fft = 2048 front-end (front2048_kernel)
  int j0_2048 = 0, j1_2048 = 16, mw_floats_2048 = 0;
  float2* tw2048_dev = nullptr;
  int *mstart_dev = nullptr, *mcount_dev = nullptr, *moff_dev = nullptr;
  float* mweights_dev = nullptr;
  V6Sweep sweep6;             // dense per-filter mel program of the v6 kernel (kernel-parameter bank)
  bool v6 = false;
  V5Sweep sweep5;             // half-warp variant (experimental, MAFE_HALFWARP_SWEEP)
  bool v5 = false;
  int* comb5_dev = nullptr;
};

int fast_tile_frames() { return kTileFrames; }

// per-bin form of the filterbank: <= 2 ADJACENT filters per bin, non-decreasing in k
static bool build_bins_n(const mafe_frontend_desc* d, int nb, float scale, std::vector<BinEntry>& bins) {
  const int nm = d->n_mels;
  bins.assign(nb, BinEntry{0, 0.f, 0.f, 0});
  int prev = -1;
  for (int k = 0; k < nb; ++k) {
    int first = -1, count = 0, last = -1;
    for (int m = 0; m < nm; ++m)
      if (d->mel_fb[(size_t)m * nb + k] != 0.f) { if (first < 0) first = m; last = m; ++count; }
    if (count > 2 || (count == 2 && last != first + 1)) return false;
    BinEntry e{prev, 0.f, 0.f, 0};
    if (count == 2) {
      e.f0 = first; e.w0 = d->mel_fb[(size_t)first * nb + k]; e.w1 = d->mel_fb[(size_t)last * nb + k];
    } else if (count == 1) {
      // keep f0 monotone: attach the single filter as the upper one when that keeps order, else as the lower
      if (first - 1 >= prev) { e.f0 = first - 1; e.w1 = d->mel_fb[(size_t)first * nb + k]; }
      else { e.f0 = first; e.w0 = d->mel_fb[(size_t)first * nb + k]; }
    } else {
      e.f0 = prev;  // empty bin: stay where we are
    }
    if (e.f0 < prev) return false;
    prev = e.f0;
    e.w0 *= scale; e.w1 *= scale;
    bins[k] = e;
  }
  return true;
}
// 512-point kernels: the pair separation leaves 2*X, so |2X|^2 / 4
static bool build_bins(const mafe_frontend_desc* d, std::vector<BinEntry>& bins) { return build_bins_n(d, kBins, 0.25f, bins); }

// the kernel's emit pattern, replayed on the host (it does not depend on data)
static bool build_combine_n(const std::vector<BinEntry>& bins, int nm, int nb, int bins_per_warp, std::vector<int2>& ranges,
                            std::vector<int>& comb) {
  ranges.assign(kFastWarps, int2{0, -1});
  std::vector<std::vector<int>> who(nm);
  for (int w = 0; w < kFastWarps; ++w) {
    const int k_lo = std::min(bins_per_warp * w, nb - 1);
    const int k_hi = (w == kFastWarps - 1) ? nb : std::min(bins_per_warp * w + bins_per_warp, nb);
    int lo = bins[k_lo].f0, hi = bins[k_hi - 1].f0 + 1;
    ranges[w] = int2{lo, hi};
    for (int m = std::max(lo, 0); m <= std::min(hi, nm - 1); ++m) who[m].push_back(w);
  }
  comb.assign(nm, 0);
  for (int m = 0; m < nm; ++m) {
    if (who[m].size() > 2) return false;
    if (who[m].size() == 2 && who[m][1] != who[m][0] + 1) return false;
    int n = (int)who[m].size();
    int p0 = n ? (who[m][0] & 1) : 0;
    comb[m] = n | (p0 << 2);
  }
  return true;
}
// Sweep program of the v3 kernel.  The 128 sub-transform outputs kk (FFT bins 2kk, 2kk+1; the Nyquist bin rides with the
// last warp) are split into 8 contiguous warp ranges that BALANCE the sweep cost (bins x c_bin + retired filters x
// c_ret per half): the mel triangles are narrower than an FFT bin at the low end and ~16 bins wide at the top, so equal
// bin counts would leave the first warp with 4x the retires of the last.  Exact min-max partition by dynamic programming.
static bool build_v3_program(const std::vector<BinEntry>& bins, V3Sweep& S, std::vector<int>& comb3) {
  memset(&S, 0, sizeof(S));
  comb3.assign(kV2Mels, 0);
  constexpr int NK = 128, W = kFastWarps;
  constexpr int c_bin = 17, c_ret = 8;   // instruction counts per bin / per retire (SASS)
  auto lo_of = [&](int a) { return bins[2 * a].f0; };
  auto hi_of = [&](int b) { return bins[b == NK ? kBins - 1 : 2 * b - 1].f0 + 1; };   // last warp takes bin 256 too
  auto cost = [&](int a, int b) -> long {
    const int nret = hi_of(b) - lo_of(a) + 1;
    if (nret > kV3Runs) return -1;
    // a filter may be emitted by at most two (adjacent) warps: the range must advance the filter index by >= 2
    if (a > 0 && b < NK && bins[2 * b].f0 - bins[2 * a - 1].f0 < 2) return -1;
    return (long)c_bin * (2 * (b - a) + (b == NK ? 1 : 0)) + (long)c_ret * 2 * nret;
  };
  const long INF = 1L << 60;
  std::vector<std::vector<long>> best(W + 1, std::vector<long>(NK + 1, INF));
  std::vector<std::vector<int>> from(W + 1, std::vector<int>(NK + 1, -1));
  best[0][0] = 0;
  for (int w = 1; w <= W; ++w)
    for (int b = w; b <= NK; ++b)
      for (int a = w - 1; a < b; ++a) {
        if (best[w - 1][a] == INF) continue;
        const long c = cost(a, b);
        if (c < 0) continue;
        const long v = std::max(best[w - 1][a], c);
        if (v < best[w][b]) { best[w][b] = v; from[w][b] = a; }
      }
  if (best[W][NK] == INF) return false;
  int edge[W + 1];
  edge[W] = NK;
  for (int w = W; w > 0; --w) edge[w - 1] = from[w][edge[w]];

  int rows = 0;
  std::vector<int> nrow(kV2Mels, 0);
  for (int w = 0; w < W; ++w) {
    const int a = edge[w], b = edge[w + 1];
    const int lo = lo_of(a), hi = hi_of(b);
    S.kk0[w] = (unsigned char)a;
    S.kk0[w + 1] = (unsigned char)b;
    S.row0[w] = (unsigned char)rows;
    for (int m = lo; m <= hi; ++m, ++rows)
      if (m >= 0 && m < kV2Mels) {
        if (nrow[m] >= 2) return false;
        comb3[m] |= rows << (8 * nrow[m]);
        ++nrow[m];
      }
    for (int h = 0; h < 2; ++h) {
      const int kk_end = b + ((h == 0 && b == NK) ? 1 : 0);
      int cur = lo;
      if (kk_end - a > 16 * kV3MaskWords) return false;
      for (int kk = a; kk < kk_end; ++kk) {
        const BinEntry& e = bins[2 * kk + h];
        const int nret = e.f0 - cur;
        if (nret > 3) return false;   // 2 bits per step in the retire mask
        S.step[h * kV3HalfStride + kk] = V3Step{e.w0, e.w1, nret, 0};
        S.nret_mask[h * W + w][(kk - a) / 16] |= (uint32_t)nret << (2 * ((kk - a) % 16));
        cur = e.f0;
      }
      S.tail[h][w] = (unsigned char)(hi - cur + 1);
    }
  }
  if (rows >= kV3PlaneRows) return false;
  S.zero_row = rows;
  for (int m = 0; m < kV2Mels; ++m) {
    if (nrow[m] == 0) comb3[m] = rows | (rows << 8);
    else if (nrow[m] == 1) comb3[m] |= rows << 8;
  }
  return true;
}

// Mel program of the v6 kernel: every filter = a run of aligned 4-bin chunks of the natural-order power row with
// zero-padded weights.  The 80 filters are dealt to the 8 warps as contiguous ranges; inside a warp every filter takes the
// chunk count of the warp's widest one (no per-filter dispatch in the kernel), and the split minimises the largest warp
// cost  filters x (6 + 5 chunks)  by dynamic programming.
static bool build_v6_program(const std::vector<BinEntry>& bins, V6Sweep& S) {
  memset(&S, 0, sizeof(S));
  constexpr int W = kFastWarps, M = kV2Mels;
  std::vector<float> dense((size_t)M * kBins, 0.f);
  for (int k = 0; k < kBins; ++k) {
    const BinEntry& e = bins[k];
    if (e.w0 != 0.f && e.f0 >= 0 && e.f0 < M) dense[(size_t)e.f0 * kBins + k] = e.w0;
    if (e.w1 != 0.f && e.f0 + 1 >= 0 && e.f0 + 1 < M) dense[(size_t)(e.f0 + 1) * kBins + k] = e.w1;
  }
  int c0[M], n4[M];
  for (int m = 0; m < M; ++m) {
    int k0 = -1, k1 = -1;
    for (int k = 0; k < kBins; ++k)
      if (dense[(size_t)m * kBins + k] != 0.f) { if (k0 < 0) k0 = k; k1 = k + 1; }
    if (k0 < 0) { k0 = 0; k1 = 1; }   // empty filter: one all-zero chunk
    c0[m] = k0 / 4;
    n4[m] = (k1 + 3) / 4 - c0[m];
    if (n4[m] > kV6MaxN4) return false;
  }
  auto cost = [&](int a, int b) -> long {   // filters a .. b - 1 in one warp
    int nw = 0;
    for (int m = a; m < b; ++m) nw = std::max(nw, n4[m]);
    return (long)(b - a) * (6 + 5 * nw);
  };
  const long INF = 1L << 60;
  std::vector<std::vector<long>> best(W + 1, std::vector<long>(M + 1, INF));
  std::vector<std::vector<int>> from(W + 1, std::vector<int>(M + 1, -1));
  best[0][0] = 0;
  for (int w = 1; w <= W; ++w)
    for (int b = w; b <= M; ++b)
      for (int a = w - 1; a < b; ++a) {
        if (best[w - 1][a] == INF) continue;
        const long v = std::max(best[w - 1][a], cost(a, b));
        if (v < best[w][b]) { best[w][b] = v; from[w][b] = a; }
      }
  if (best[W][M] == INF) return false;
  int edge[W + 1];
  edge[W] = M;
  for (int w = W; w > 0; --w) edge[w - 1] = from[w][edge[w]];
  int chunks = 0;
  for (int w = 0; w < W; ++w) {
    const int a = edge[w], b = edge[w + 1];
    int nw = 0;
    for (int m = a; m < b; ++m) nw = std::max(nw, n4[m]);
    S.f0[w] = (unsigned char)a;
    S.nw[w] = (unsigned char)nw;
    S.wbase[w] = (unsigned short)chunks;
    if (chunks + (b - a) * nw > kV6MaxChunks) return false;
    for (int m = a; m < b; ++m) {
      int start = c0[m];
      if (4 * (start + nw) > kV6Row) start = kV6Row / 4 - nw;   // keep the padded run inside the row
      S.start[m] = (unsigned short)(4 * start);
      for (int j = 0; j < nw; ++j) {
        float w4[4];
        for (int i = 0; i < 4; ++i) {
          const int k = 4 * (start + j) + i;
          w4[i] = k < kBins ? dense[(size_t)m * kBins + k] : 0.f;
        }
        S.w[chunks++] = make_float4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
  }
  S.f0[W] = (unsigned char)M;
  return true;
}

// Half-warp variant: 16 cost-balanced ranges (one per half-warp), a filter may be emitted by up to three of them.
static bool build_v5_program(const std::vector<BinEntry>& bins, V5Sweep& S, std::vector<int>& comb5) {
  memset(&S, 0, sizeof(S));
  comb5.assign(kV2Mels, 0);
  constexpr int NK = 128, W = kV5Ranges;
  constexpr int c_bin = 14, c_ret = 8;   // per (pair, bin) step / per retire, relative
  auto lo_of = [&](int a) { return bins[2 * a].f0; };
  auto hi_of = [&](int b) { return bins[b == NK ? kBins - 1 : 2 * b - 1].f0 + 1; };
  auto cost = [&](int a, int b) -> long {
    const int nret = hi_of(b) - lo_of(a) + 1;
    if (nret > kV3Runs || b - a + 1 > 32) return -1;
    return (long)c_bin * (2 * (b - a) + (b == NK ? 1 : 0)) + (long)c_ret * 2 * nret;
  };
  const long INF = 1L << 60;
  std::vector<std::vector<long>> best(W + 1, std::vector<long>(NK + 1, INF));
  std::vector<std::vector<int>> from(W + 1, std::vector<int>(NK + 1, -1));
  best[0][0] = 0;
  for (int w = 1; w <= W; ++w)
    for (int b = w; b <= NK; ++b)
      for (int a = w - 1; a < b; ++a) {
        if (best[w - 1][a] == INF) continue;
        const long c = cost(a, b);
        if (c < 0) continue;
        const long v = std::max(best[w - 1][a], c);
        if (v < best[w][b]) { best[w][b] = v; from[w][b] = a; }
      }
  if (best[W][NK] == INF) return false;
  int edge[W + 1];
  edge[W] = NK;
  for (int w = W; w > 0; --w) edge[w - 1] = from[w][edge[w]];
  int rows = 0;
  std::vector<int> nrow(kV2Mels, 0);
  for (int w = 0; w < W; ++w) {
    const int a = edge[w], b = edge[w + 1];
    const int lo = lo_of(a), hi = hi_of(b);
    S.kk0[w] = (unsigned char)a;
    S.kk0[w + 1] = (unsigned char)b;
    S.row0[w] = (unsigned char)rows;
    for (int m = lo; m <= hi; ++m, ++rows)
      if (m >= 0 && m < kV2Mels) {
        if (nrow[m] >= 3) return false;
        comb5[m] |= rows << (8 * nrow[m]);
        ++nrow[m];
      }
    for (int h = 0; h < 2; ++h) {
      const int kk_end = b + ((h == 0 && b == NK) ? 1 : 0);
      int cur = lo;
      for (int kk = a; kk < kk_end; ++kk) {
        const BinEntry& e = bins[2 * kk + h];
        const int nret = e.f0 - cur;
        if (nret > 3) return false;
        S.step[h * kV3HalfStride + kk] = V3Step{e.w0, e.w1, nret, 0};
        S.nret_mask[h * W + w][(kk - a) / 16] |= (uint32_t)nret << (2 * ((kk - a) % 16));
        cur = e.f0;
      }
      S.tail[h][w] = (unsigned char)(hi - cur + 1);
    }
  }
  if (rows >= kV5PlaneRows || rows > 254) return false;
  S.zero_row = rows;
  for (int m = 0; m < kV2Mels; ++m)
    for (int r = nrow[m]; r < 3; ++r) comb5[m] |= rows << (8 * r);
  return true;
}

// Same for the one-pass sweep of fbank400_kernel: bins 0 .. 200 split into 8 cost-balanced contiguous warp ranges.
static bool build_f400_program(const std::vector<BinEntry>& bins, int nm, F400Sweep& S, std::vector<int>& comb) {
  memset(&S, 0, sizeof(S));
  comb.assign(nm, 0);
  constexpr int NK = kBins400, W = kFastWarps;
  constexpr int c_bin = 17, c_ret = 7;
  auto lo_of = [&](int a) { return bins[a].f0; };
  auto hi_of = [&](int b) { return bins[b - 1].f0 + 1; };
  auto cost = [&](int a, int b) -> long {
    if (a > 0 && b < NK && bins[b].f0 - bins[a - 1].f0 < 2) return -1;   // a filter is emitted by <= 2 adjacent warps
    return (long)c_bin * (b - a) + (long)c_ret * (hi_of(b) - lo_of(a) + 1);
  };
  const long INF = 1L << 60;
  std::vector<std::vector<long>> best(W + 1, std::vector<long>(NK + 1, INF));
  std::vector<std::vector<int>> from(W + 1, std::vector<int>(NK + 1, -1));
  best[0][0] = 0;
  for (int w = 1; w <= W; ++w)
    for (int b = w; b <= NK; ++b)
      for (int a = w - 1; a < b; ++a) {
        if (best[w - 1][a] == INF) continue;
        const long c = cost(a, b);
        if (c < 0) continue;
        const long v = std::max(best[w - 1][a], c);
        if (v < best[w][b]) { best[w][b] = v; from[w][b] = a; }
      }
  if (best[W][NK] == INF) return false;
  int edge[W + 1];
  edge[W] = NK;
  for (int w = W; w > 0; --w) edge[w - 1] = from[w][edge[w]];
  int rows = 0;
  std::vector<int> nrow(nm, 0);
  for (int w = 0; w < W; ++w) {
    const int a = edge[w], b = edge[w + 1];
    const int lo = lo_of(a), hi = hi_of(b);
    S.kk0[w] = (unsigned char)a;
    S.kk0[w + 1] = (unsigned char)b;
    S.row0[w] = (unsigned char)rows;
    for (int m = lo; m <= hi; ++m, ++rows)
      if (m >= 0 && m < nm) {
        if (nrow[m] >= 2) return false;
        comb[m] |= rows << (8 * nrow[m]);
        ++nrow[m];
      }
    int cur = lo;
    if (b - a > 16 * kV3MaskWords) return false;
    for (int k = a; k < b; ++k) {
      const int nret = bins[k].f0 - cur;
      if (nret > 3) return false;   // 2 bits per step in the retire mask
      S.step[k] = V3Step{bins[k].w0, bins[k].w1, nret, 0};
      S.nret_mask[w][(k - a) / 16] |= (uint32_t)nret << (2 * ((k - a) % 16));
      cur = bins[k].f0;
    }
    S.tail[w] = (unsigned char)(hi - cur + 1);
  }
  if (rows >= kMaxRows400 || rows > 254) return false;
  S.zero_row = rows;
  for (int m = 0; m < nm; ++m) {
    if (nrow[m] == 0) comb[m] = rows | (rows << 8);
    else if (nrow[m] == 1) comb[m] |= rows << 8;
  }
  return true;
}

static bool build_combine(const std::vector<BinEntry>& bins, int nm, std::vector<int2>& ranges, std::vector<int>& comb) {
  return build_combine_n(bins, nm, kBins, 32, ranges, comb);
}

static bool stft_plan_supported(const mafe_frontend_desc* d) {
  if (!(d->out_kind == MAFE_OUT_COMPLEX || (d->out_kind == MAFE_OUT_POWER && d->power > 0.f))) return false;
  return d->n_fft == kNfft && d->frame_len == kNfft && d->hop >= 1 && d->hop <= kStftMaxHop &&
         d->preemph == 0.0 && !d->remove_frame_mean && d->dither == 0.f;
}

// n_fft = 320 (deepspeech2) / 400 complex STFT: N1 of stftn16_kernel, or 0
static int stftn_plan_supported(const mafe_frontend_desc* d) {
  if (!(d->out_kind == MAFE_OUT_COMPLEX || d->out_kind == MAFE_OUT_POWER) || d->frame_len != d->n_fft || d->preemph != 0.0 ||
      d->remove_frame_mean || d->dither != 0.f)
    return 0;
  if (d->out_kind == MAFE_OUT_POWER && !(d->power > 0.f)) return 0;
  if (d->n_fft != 320 && d->n_fft != 400) return 0;
  if (d->hop < 1 || d->hop > d->n_fft / 2) return 0;
  return d->n_fft / 16;
}

static bool f400_plan_supported(const mafe_frontend_desc* d) {
  if (!(d->n_fft == kN400 && d->frame_len == kN400 && d->hop >= 1 && d->hop <= kMaxHop400)) return false;
  if (!(d->out_kind == MAFE_OUT_MEL || d->out_kind == MAFE_OUT_LOGMEL || d->out_kind == MAFE_OUT_MFCC)) return false;
  if (d->power != 2.0f || d->preemph != 0.0 || d->remove_frame_mean || d->dither != 0.f) return false;
  if (d->n_mels < 2 || d->n_mels > kMaxMels400) return false;
  std::vector<BinEntry> bins;
  std::vector<int> comb;
  F400Sweep sw;
  return build_bins_n(d, kBins400, 1.0f, bins) && build_f400_program(bins, d->n_mels, sw, comb);
}

// n_fft = 2048: every output kind, hop <= 512, no pre-emphasis / frame-mean removal / dither (front2048.cuh)
static bool f2048_plan_supported(const mafe_frontend_desc* d) {
  if (d->n_fft != kN2048 || d->frame_len != kN2048 || d->hop < 1 || d->hop > kMaxHop2048) return false;
  if (d->preemph != 0.0 || d->remove_frame_mean || d->dither != 0.f) return false;
  if (d->out_kind == MAFE_OUT_POWER && !(d->power > 0.f)) return false;
  if (d->out_kind >= MAFE_OUT_MEL && (!(d->power > 0.f) || d->n_mels < 1)) return false;
  return true;
}

bool fast_plan_supported(const mafe_frontend_desc* d) {
  if (f2048_plan_supported(d)) return true;
  if (stft_plan_supported(d)) return true;
  if (stftn_plan_supported(d)) return true;
  if (f400_plan_supported(d)) return true;
  if (d->n_fft != kNfft || d->center || d->out_kind != MAFE_OUT_LOGMEL || d->power != 2.0f || d->spec_scale != 1.0f)
    return false;
  if (!(d->log_kind == MAFE_LOG_LN_EPS_IF_ZERO || d->log_kind == MAFE_LOG_LN_PLUS || d->log_kind == MAFE_LOG_NONE)) return false;
  if (d->n_mels < 1 || d->n_mels > 128) return false;
  const int ylen = (kTileFrames - 1) * d->hop + d->frame_len;
  if (ylen > kMaxY || 2 * d->n_mels * kPlaneStride > kMaxY) return false;
  std::vector<BinEntry> bins;
  if (!build_bins(d, bins)) return false;
  std::vector<int2> ranges;
  std::vector<int> comb;
  return build_combine(bins, d->n_mels, ranges, comb);
}

template <typename T>
static int up(T** dev, const std::vector<T>& h) {
  MAFE_CUDA_CHECK(cudaMalloc((void**)dev, std::max<size_t>(h.size(), 1) * sizeof(T)));
  MAFE_CUDA_CHECK(cudaMemcpy(*dev, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  return MAFE_OK;
}

int fast_plan_init(mafe_ctx* ctx, mafe_plan* p, const mafe_frontend_desc* d) {
  (void)ctx;
  FastTablesHost* th = new FastTablesHost();
  memset(&th->dev, 0, sizeof(th->dev));
  th->ylen = (kTileFrames - 1) * d->hop + d->frame_len;
  p->fast_tables = th;
  th->stft = stft_plan_supported(d);
  std::vector<float> win(std::max(kNfft, d->frame_len), 0.f);
  for (int i = 0; i < d->frame_len; ++i) win[i] = th->stft ? 0.5f * d->spec_scale * d->window[i] : d->window[i];
  std::vector<float2> w512(256), w256(256);
  for (int n = 0; n < 256; ++n) {
    double a = -2.0 * M_PI * n / 512.0;
    w512[n] = make_float2((float)cos(a), (float)sin(a));
  }
  for (int kj = 0; kj < 16; ++kj)
    for (int t = 0; t < 16; ++t) {
      double a = -2.0 * M_PI * (double)(t * kj) / 256.0;
      w256[kj * 16 + t] = make_float2((float)cos(a), (float)sin(a));
    }
  int rc;
  th->f2048 = f2048_plan_supported(d);
  if (th->f2048) {
    std::vector<float> w2(kN2048);
    int lo = kN2048, hi = 0;
    for (int i = 0; i < kN2048; ++i) {
      w2[i] = 0.5f * d->spec_scale * d->window[i];   // 1/2: the pair separation leaves 2X
      if (d->window[i] != 0.f) { lo = std::min(lo, i); hi = std::max(hi, i + 1); }
    }
    th->j0_2048 = lo < hi ? lo / 128 : 0;
    th->j1_2048 = lo < hi ? (hi + 127) / 128 : 0;
    std::vector<float2> tw(kN2048);
    for (int j = 0; j < kN2048; ++j) {
      const double a = -2.0 * M_PI * (double)j / (double)kN2048;
      tw[j] = make_float2((float)cos(a), (float)sin(a));
    }
    if ((rc = up(&th->dev.window, w2))) return rc;
    if ((rc = up(&th->tw2048_dev, tw))) return rc;
    if (d->out_kind >= MAFE_OUT_MEL) {
      // Per warp of 32 filters (thread = filter in the kernel): every filter's support starts on a multiple of 4 bins
      // (16-byte loads of the power rows) and is padded with zero weights to the warp's widest filter, so the warp walks
      // nch 4-bin chunks in lock step; the weights of chunk i sit at float4 index base + 32 i + lane: consecutive lanes
      // read consecutive 16-byte words.  (First version: compact per-filter arrays at unaligned starts -- the weight loads
      // ran at 2.6 x and the eight scalar power loads per chunk at 2.7 x their ideal wavefronts, ncu source page.)
      std::vector<int> ms(d->n_mels, 0), mc(d->n_mels, 0), mo(d->n_mels, 0);
      std::vector<float> mw;
      for (int m0 = 0; m0 < d->n_mels; m0 += 32) {
        const int m1 = std::min(d->n_mels, m0 + 32);
        std::vector<int> k0s(32, 0), k1s(32, 0);
        int nch = 1;
        for (int m = m0; m < m1; ++m) {
          int k0 = -1, k1 = -1;
          for (int k = 0; k < kBins2048; ++k)
            if (d->mel_fb[(size_t)m * kBins2048 + k] != 0.f) { if (k0 < 0) k0 = k; k1 = k + 1; }
          if (k0 < 0) { k0 = 0; k1 = 0; }
          k0s[m - m0] = k0 & ~3; k1s[m - m0] = k1;
          nch = std::max(nch, (k1 - (k0 & ~3) + 3) / 4);
        }
        const int base4 = (int)(mw.size() / 4);
        mw.resize(mw.size() + (size_t)nch * 32 * 4, 0.f);
        for (int m = m0; m < m1; ++m) {
          int k0 = k0s[m - m0];
          if (k0 + 4 * nch > kPRow2048) k0 = kPRow2048 - 4 * nch;   // rows hold kPRow2048 (multiple of 4) floats, the pad is zero
          ms[m] = k0; mc[m] = nch; mo[m] = base4;
          for (int i = 0; i < nch; ++i)
            for (int j = 0; j < 4; ++j) {
              const int k = k0 + 4 * i + j;
              mw[((size_t)base4 + 32 * i + (m - m0)) * 4 + j] = (k >= 0 && k < kBins2048) ? d->mel_fb[(size_t)m * kBins2048 + k] : 0.f;
            }
        }
      }
      th->mw_floats_2048 = (int)mw.size();
      if ((rc = up(&th->mstart_dev, ms))) return rc;
      if ((rc = up(&th->mcount_dev, mc))) return rc;
      if ((rc = up(&th->moff_dev, mo))) return rc;
      if ((rc = up(&th->mweights_dev, mw))) return rc;
    }
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(front2048_kernel<3, 13>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF2048SmemBudget));
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(front2048_kernel<0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF2048SmemBudget));
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(front2048_kernel<-1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kF2048SmemBudget));
    return MAFE_OK;
  }
  th->stftn = stftn_plan_supported(d);
  if (th->stftn) {
    const int N1 = th->stftn, N = 16 * N1, B = N1 == 25 ? 5 : 4;   // N1 = 5 x B: twiddles W_N1^(j1 k1), j1 = 1..4, k1 = 1..B-1
    std::vector<float> wn(N);
    for (int i = 0; i < N; ++i) wn[i] = 0.5f * d->spec_scale * d->window[i];   // 1/2: the pair separation leaves 2X
    std::vector<float2> twn(2 * N);   // t = 16..31: the rotated upper half-warp of stftn16_kernel
    for (int kj = 0; kj < N1; ++kj)
      for (int t = 0; t < 32; ++t) {
        double a = -2.0 * M_PI * (double)((t * kj) % N) / (double)N;
        twn[kj * 32 + t] = make_float2((float)cos(a), (float)sin(a));
      }
    memset(th->tws, 0, sizeof(th->tws));
    for (int j1 = 1; j1 < 5; ++j1)
      for (int k1 = 1; k1 < B; ++k1) {
        double a = -2.0 * M_PI * (double)(j1 * k1) / (double)N1;
        th->tws[(j1 - 1) * (B - 1) + (k1 - 1)] = make_float2((float)cos(a), (float)sin(a));
      }
    if ((rc = up(&th->dev.window, wn))) return rc;
    if ((rc = up(&th->tw400_dev, twn))) return rc;
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(stftn16_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StftN<20>::kTotal));
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(stftn16_kernel<25>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StftN<25>::kTotal));
    return MAFE_OK;
  }
  th->f400 = !th->stft && f400_plan_supported(d);
  if (th->f400) {
    std::vector<float> w400(kN400);
    for (int i = 0; i < kN400; ++i) w400[i] = 0.5f * d->spec_scale * d->window[i];   // 1/2: the pair separation leaves 2X
    std::vector<float2> tw400(2 * kN400);   // t = 16..31: the rotated upper half-warp
    for (int kj = 0; kj < 25; ++kj)
      for (int t = 0; t < 32; ++t) {
        double a = -2.0 * M_PI * (double)((t * kj) % 400) / 400.0;
        tw400[kj * 32 + t] = make_float2((float)cos(a), (float)sin(a));
      }
    for (int j1 = 1; j1 < 5; ++j1)
      for (int k1 = 1; k1 < 5; ++k1) {
        double a = -2.0 * M_PI * (double)(j1 * k1) / 25.0;
        th->tw25[(j1 - 1) * 4 + (k1 - 1)] = make_float2((float)cos(a), (float)sin(a));
      }
    std::vector<BinEntry> bins;
    std::vector<int> comb;
    if (!build_bins_n(d, kBins400, 1.0f, bins) || !build_f400_program(bins, d->n_mels, th->sweep400, comb)) {
      set_error("filterbank is not in per-bin form");
      return MAFE_E_UNSUPPORTED;
    }
    if ((rc = up(&th->dev.window, w400))) return rc;
    if ((rc = up(&th->tw400_dev, tw400))) return rc;
    if ((rc = up(&th->dev.combine, comb))) return rc;
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank400_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f400_smem_bytes(kMaxRows400)));
    return MAFE_OK;
  }
  if (th->stft) {
    if ((rc = up(&th->dev.window, win))) return rc;
    if ((rc = up(&th->dev.w512, w512))) return rc;
    if ((rc = up(&th->dev.w256t, w256))) return rc;
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(stft512_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StftSmem::kTotal));
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(stft512_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)StftSmem::kTotal));
    return MAFE_OK;
  }
  std::vector<BinEntry> bins;
  std::vector<int2> ranges;
  std::vector<int> comb;
  if (!build_bins(d, bins) || !build_combine(bins, d->n_mels, ranges, comb)) {
    set_error("filterbank is not in per-bin form");
    return MAFE_E_UNSUPPORTED;
  }
  if ((rc = up(&th->dev.window, win))) return rc;
  if ((rc = up(&th->dev.w512, w512))) return rc;
  if ((rc = up(&th->dev.w256t, w256))) return rc;
  if ((rc = up(&th->dev.bins, bins))) return rc;
  if ((rc = up(&th->dev.warp_range, ranges))) return rc;
  if ((rc = up(&th->dev.combine, comb))) return rc;
  {
    std::vector<float> cover(d->hop, 0.f);
    for (int r = 0; r < d->hop; ++r) {
      double c = 0.0;
      for (int n = r; n < d->frame_len; n += d->hop) c += (double)d->window[n];
      cover[r] = (float)c;
    }
    if ((rc = up(&th->dev.cover, cover))) return rc;
    if (d->hop == kV2Hop) {
      // c'(r) = c(r) - a c(r + 1) (fbank512_v6.cuh, FS), four copies, copy k shifted by k entries: a thread's 4 (8)
      // coefficients are aligned 16-byte loads from copy r mod 4
      std::vector<double> cd(d->hop, 0.0);
      for (int r = 0; r < d->hop; ++r)
        for (int n = r; n < d->frame_len; n += d->hop) cd[r] += (double)d->window[n];
      std::vector<float> c4(4 * kCwRow, 0.f);
      for (int k = 0; k < 4; ++k)
        for (int j = 0; j < kCwRow; ++j) {
          const int r = (j + k) % d->hop;
          c4[k * kCwRow + j] = (float)(cd[r] - d->preemph * cd[(r + 1) % d->hop]);
        }
      if ((rc = up(&th->cover4_dev, c4))) return rc;
    }
  }
  MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FastSmem::kTotal));
  th->tile_geom = d->frame_len == kV2Flen && d->hop == kV2Hop && d->n_mels == kV2Mels;
  if (th->tile_geom) {
    std::vector<int> comb3;
    th->v3 = build_v3_program(bins, th->sweep, comb3);
    if ((rc = up(&th->comb3_dev, comb3))) return rc;
    th->v6 = th->v3 && build_v6_program(bins, th->sweep6);
    if (th->v6) {
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v6_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6Smem::kTotal));
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v6_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6Smem::kTotal));
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v6_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6Smem::kTotal));
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v6_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6Smem::kTotal));
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v6_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6Smem::kTotal));
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v6_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6Smem::kTotal));
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v6_kernel<false, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6Smem::kTotal));
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v6_kernel<true, true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V6Smem::kTotal));
    }
    std::vector<int> comb5;
    th->v5 = th->v3 && build_v5_program(bins, th->sweep5, comb5);
    if (th->v5) {
      if ((rc = up(&th->comb5_dev, comb5))) return rc;
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v3_kernel<false, true, V5Sweep>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V5Smem::kTotal));
      MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v3_kernel<true, true, V5Sweep>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V5Smem::kTotal));
    }
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3Smem::kTotal));
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(fbank512_v3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V3Smem::kTotal));
  }
  return MAFE_OK;
}

void fast_plan_free(mafe_plan* p) {
  FastTablesHost* th = static_cast<FastTablesHost*>(p->fast_tables);
  if (!th) return;
  cudaFree(th->dev.window); cudaFree(th->dev.w512); cudaFree(th->dev.w256t);
  cudaFree(th->dev.bins); cudaFree(th->dev.warp_range); cudaFree(th->dev.combine); cudaFree(th->dev.cover);
  cudaFree(th->comb3_dev);
  cudaFree(th->comb5_dev);
  cudaFree(th->tw400_dev);
  cudaFree(th->cover4_dev);
  cudaFree(th->tw2048_dev); cudaFree(th->mstart_dev); cudaFree(th->mcount_dev); cudaFree(th->moff_dev); cudaFree(th->mweights_dev);
  delete th;
  p->fast_tables = nullptr;
}

__global__ void fill_keys_kernel(int* p, int n, int v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

bool fast_has_aux_mel(const mafe_plan* p) {   // kernels that can emit the mel energies beside the log-mel (mafe_frontend_run_aux)
  const FastTablesHost* th = static_cast<const FastTablesHost*>(p->fast_tables);
  return th != nullptr && th->f400 && p->d.out_kind >= MAFE_OUT_LOGMEL;
}

int fast_run(mafe_ctx* ctx, const mafe_plan* p, mafe_batch* b, const void* wave, int wave_dtype, float wave_scale, float* out,
             int db_group) {
  if (b->n_tiles == 0) return MAFE_OK;
  const FastTablesHost* th = static_cast<const FastTablesHost*>(p->fast_tables);
  const mafe_frontend_desc& d = p->d;
  if (th->f2048) {
    F2048Params F;
    F.wave = wave; F.wave_dtype = wave_dtype; F.wave_scale = wave_scale;
    F.sample_offsets = b->sample_offsets_dev; F.frame_offsets = b->frame_offsets_dev; F.tiles = b->tiles_dev;
    F.n_tiles = b->n_tiles; F.hop = d.hop; F.center = d.center; F.pad_mode = d.pad_mode;
    F.j0 = th->j0_2048; F.j1 = th->j1_2048;
    F.window = th->dev.window; F.tw = th->tw2048_dev;
    F.out_kind = d.out_kind == MAFE_OUT_MFCC ? MAFE_OUT_LOGMEL : d.out_kind;
    F.power = d.power; F.n_mels = d.n_mels;
    F.mstart = th->mstart_dev; F.mcount = th->mcount_dev; F.moff = th->moff_dev; F.mweights = th->mweights_dev;
    F.log_kind = (d.out_kind == MAFE_OUT_MEL || d.out_kind == MAFE_OUT_COMPLEX) ? MAFE_LOG_NONE : d.log_kind;
    F.log_arg = d.log_arg; F.log_mult = d.log_mult; F.log_offset = d.log_offset;
    F.out = out; F.out_dim = d.out_kind >= MAFE_OUT_MEL ? d.n_mels : p->out_dim;
    F.queue_head = b->queue_dev;
    F.db_group = (d.out_kind >= MAFE_OUT_LOGMEL && d.log_kind == MAFE_LOG_DB && d.top_db >= 0.f) ? db_group : MAFE_DBGROUP_NONE;
    F.group_max = b->group_max_dev; F.utt_group = b->utt_group_dev;
    // staging buffers for this hop and window support: (16 - 1) hop + 128 (j1 - j0) samples, rounded to 16 bytes; two when
    // they fit 2 CTAs per SM (fastspeech2: win 1200 -> 10 of 16 rows, 23 KB per buffer instead of 26: both fit)
    F.stage_floats = (((kHalfFrames2048 - 1) * d.hop + 128 * (th->j1_2048 - th->j0_2048)) + 3 + 4) & ~3;   // + the alignment shift (<= 3)
    F.mw_floats = d.out_kind >= MAFE_OUT_MEL ? ((th->mw_floats_2048 + 3) & ~3) : 0;
    F.n_stage = f2048_smem_total(F.stage_floats, 2, F.mw_floats) <= kF2048SmemBudget ? 2 : 1;
    const size_t smem_bytes = f2048_smem_total(F.stage_floats, F.n_stage, F.mw_floats);
    if (smem_bytes > kF2048SmemBudget) return MAFE_E_UNSUPPORTED;   // (a filterbank with an enormous support): generic route
    MAFE_CUDA_CHECK(cudaMemsetAsync(b->queue_dev, 0, sizeof(int32_t), ctx->stream));
    if (F.db_group != MAFE_DBGROUP_NONE) {
      const int n = std::max(b->n_groups, 1);
      fill_keys_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(b->group_max_dev, n, (int)0x80000000);
      MAFE_LAUNCH_CHECK(ctx);
    }
    ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
    // window rows as compile-time constants for the two usual supports: win 1200 (fastspeech2 / wavegrad) and the full window
    const int grid = std::min(2 * b->n_tiles, 2 * ctx->sm_count);
    if (F.j0 == 3 && F.j1 == 13) front2048_kernel<3, 13><<<grid, 256, smem_bytes, ctx->stream>>>(F);
    else if (F.j0 == 0 && F.j1 == 16) front2048_kernel<0, 16><<<grid, 256, smem_bytes, ctx->stream>>>(F);
    else front2048_kernel<-1, 0><<<grid, 256, smem_bytes, ctx->stream>>>(F);
    MAFE_LAUNCH_CHECK(ctx);
    return (d.out_kind >= MAFE_OUT_MEL || d.utt_scalar_norm) ? kFastNeedsPost : MAFE_OK;
  }
  if (th->f400) {
    if (wave_dtype != MAFE_WAVE_F32 || wave_scale != 1.0f || ((uintptr_t)wave & 15) != 0) return MAFE_E_UNSUPPORTED;
    F400Params F;
    F.wave = (const float*)wave; F.total_samples = b->wave_len;
    F.sample_offsets = b->sample_offsets_dev; F.frame_offsets = b->frame_offsets_dev; F.tiles = b->tiles_dev;
    F.n_tiles = b->n_tiles; F.hop = d.hop; F.center = d.center; F.pad_mode = d.pad_mode; F.n_mels = d.n_mels;
    F.log_kind = d.out_kind == MAFE_OUT_MEL ? MAFE_LOG_NONE : d.log_kind;
    F.log_arg = d.log_arg; F.log_mult = d.log_mult; F.log_offset = d.log_offset;
    F.window = th->dev.window; F.tw400 = th->tw400_dev; F.plane_rows = th->sweep400.zero_row + 1;
    F.combine = th->dev.combine; F.out = out; F.queue_head = b->queue_dev;
    F.aux_mel = ctx->aux_mel;
    F.db_group = (d.log_kind == MAFE_LOG_DB && d.top_db >= 0.f && d.out_kind != MAFE_OUT_MEL) ? db_group : MAFE_DBGROUP_NONE;
    F.group_max = b->group_max_dev; F.utt_group = b->utt_group_dev;
    for (int i = 0; i < 16; ++i) F.tw25[i] = th->tw25[i];
    MAFE_CUDA_CHECK(cudaMemsetAsync(b->queue_dev, 0, sizeof(int32_t), ctx->stream));
    if (F.db_group != MAFE_DBGROUP_NONE) {
      const int n = std::max(b->n_groups, 1);
      fill_keys_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(b->group_max_dev, n, (int)0x80000000);
      MAFE_LAUNCH_CHECK(ctx);
    }
    ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
    fbank400_kernel<<<std::min(b->n_tiles, MAFE_F400_CTAS * ctx->sm_count), kFastThreads, f400_smem_bytes(F.plane_rows), ctx->stream>>>(F, th->sweep400);
    MAFE_LAUNCH_CHECK(ctx);
    return kFastNeedsPost;
  }
  if (th->stftn) {
    if (wave_dtype != MAFE_WAVE_F32 || wave_scale != 1.0f || ((uintptr_t)wave & 15) != 0) return MAFE_E_UNSUPPORTED;
    StftNParams S;
    S.wave = (const float*)wave; S.total_samples = b->wave_len;
    S.sample_offsets = b->sample_offsets_dev; S.frame_offsets = b->frame_offsets_dev; S.tiles = b->tiles_dev;
    S.n_tiles = b->n_tiles; S.hop = d.hop; S.center = d.center; S.pad_mode = d.pad_mode;
    S.window = th->dev.window; S.twn = th->tw400_dev; S.out = out; S.queue_head = b->queue_dev;
    S.out_power = d.out_kind == MAFE_OUT_POWER; S.power = d.power;
    S.log_kind = d.out_kind == MAFE_OUT_POWER ? d.log_kind : MAFE_LOG_NONE; S.log_arg = d.log_arg;
    for (int i = 0; i < 16; ++i) S.tws[i] = th->tws[i];
    const bool scalar_norm = d.utt_scalar_norm && d.out_kind == MAFE_OUT_POWER && b->utt_stats_dev != nullptr;
    S.utt_stats = scalar_norm ? b->utt_stats_dev : nullptr;
    MAFE_CUDA_CHECK(cudaMemsetAsync(b->queue_dev, 0, sizeof(int32_t), ctx->stream));
    if (scalar_norm) MAFE_CUDA_CHECK(cudaMemsetAsync(b->utt_stats_dev, 0, sizeof(double) * 2 * b->n_utts, ctx->stream));
    {
      ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
      const int grid = std::min(b->n_tiles, (th->stftn == 20 ? StftN<20>::kCtasPerSm : StftN<25>::kCtasPerSm) * ctx->sm_count);
      if (th->stftn == 20) stftn16_kernel<20><<<grid, kFastThreads, StftN<20>::kTotal, ctx->stream>>>(S);
      else stftn16_kernel<25><<<grid, kFastThreads, StftN<25>::kTotal, ctx->stream>>>(S);
      MAFE_LAUNCH_CHECK(ctx);
    }
    if (scalar_norm) {   // the transform accumulated sum x, sum x^2 per utterance: one streaming pass normalises
      ProfScope ps(ctx, MAFE_PROF_CMVN);
      scalar_norm_apply_kernel<<<std::min(b->n_tiles, 8 * ctx->sm_count), 256, 0, ctx->stream>>>(
          out, b->tiles_dev, b->n_tiles, b->frame_offsets_dev, b->utt_stats_dev, p->out_dim, p->tile_frames);
      MAFE_LAUNCH_CHECK(ctx);
      return MAFE_OK;
    }
    return d.utt_scalar_norm ? kFastNeedsPost : MAFE_OK;
  }
  if (th->stft) {
    if (wave_dtype != MAFE_WAVE_F32 || wave_scale != 1.0f || ((uintptr_t)wave & 15) != 0) return MAFE_E_UNSUPPORTED;
    StftParams S;
    S.wave = (const float*)wave; S.total_samples = b->wave_len;
    S.sample_offsets = b->sample_offsets_dev; S.frame_offsets = b->frame_offsets_dev; S.tiles = b->tiles_dev;
    S.n_tiles = b->n_tiles; S.hop = d.hop; S.center = d.center; S.pad_mode = d.pad_mode;
    S.window = th->dev.window; S.w512 = th->dev.w512; S.w256t = th->dev.w256t;
    S.out = out; S.queue_head = b->queue_dev;
    S.out_power = d.out_kind == MAFE_OUT_POWER; S.power = d.power;
    S.log_kind = d.out_kind == MAFE_OUT_POWER ? d.log_kind : MAFE_LOG_NONE; S.log_arg = d.log_arg;
    MAFE_CUDA_CHECK(cudaMemsetAsync(b->queue_dev, 0, sizeof(int32_t), ctx->stream));
    ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
    if (S.out_power) stft512_kernel<true><<<std::min(b->n_tiles, 2 * ctx->sm_count), kFastThreads, StftSmem::kTotal, ctx->stream>>>(S);
    else stft512_kernel<false><<<std::min(b->n_tiles, 2 * ctx->sm_count), kFastThreads, StftSmem::kTotal, ctx->stream>>>(S);
    MAFE_LAUNCH_CHECK(ctx);
    return d.utt_scalar_norm ? kFastNeedsPost : MAFE_OK;
  }
  FastParams P;
  P.wave = wave; P.wave_dtype = wave_dtype; P.wave_scale = wave_scale;
  P.sample_offsets = b->sample_offsets_dev; P.frame_offsets = b->frame_offsets_dev; P.tiles = b->tiles_dev;
  P.utt_sum = b->utt_sum_dev;
  P.frame_len = d.frame_len; P.hop = d.hop; P.n_mels = d.n_mels; P.ylen = th->ylen;
  P.pre_hi = (float)d.preemph; P.pre_lo = (float)(d.preemph - (double)P.pre_hi);
  P.preemph_on = d.preemph != 0.0; P.remove_mean = d.remove_frame_mean;
  P.dither = d.dither; P.seed = d.dither_seed;
  P.log_kind = d.log_kind; P.log_arg = d.log_arg;
  P.tab = th->dev;
  P.out = out;
  const bool cmvn = d.utt_cmvn_mean || d.utt_cmvn_std;
  if (th->tile_geom && th->v3 && ((uintptr_t)wave & 15) == 0) {
    V2Params Q;
    Q.wave = wave; Q.total_samples = b->wave_len; Q.wave_scale = wave_scale;
    Q.sample_offsets = b->sample_offsets_dev; Q.frame_offsets = b->frame_offsets_dev; Q.tiles = b->tiles_dev;
    Q.n_tiles = b->n_tiles; Q.utt_sum = b->utt_sum_dev; Q.utt_stats = cmvn ? b->utt_stats_dev : nullptr;

    Q.pre_hi = P.pre_hi; Q.pre_lo = P.pre_lo; Q.preemph_on = P.preemph_on; Q.remove_mean = P.remove_mean;
    Q.dither = P.dither; Q.seed = P.seed; Q.log_kind = P.log_kind; Q.log_arg = P.log_arg;
    Q.window = th->dev.window; Q.w512 = th->dev.w512; Q.w256t = th->dev.w256t; Q.combine = th->comb3_dev;
    Q.out = out;
    Q.queue_head = b->queue_dev;
    Q.tile_recs = nullptr; Q.sum_recs = nullptr;
    Q.utt_done = nullptr; Q.lag = 0; Q.mean_norm = 0; Q.std_norm = 0;
    Q.lag_s = 0; Q.cover4 = th->cover4_dev; Q.utt_fsum = b->utt_sum_dev; Q.fsum_done = nullptr;
    static const bool halfwarp = getenv("MAFE_HALFWARP_SWEEP") != nullptr;   // experimental: 16 ranges, two frames per lane
    static const bool force_v3 = getenv("MAFE_FBANK_V3") != nullptr;         // A/B switch: the round-1 kernel
    static const bool no_tmem = getenv("MAFE_NO_TMEM") != nullptr;           // A/B switch: constants from shared memory
    static const bool no_fuse = getenv("MAFE_NO_FUSED_CMVN") != nullptr;     // A/B switch: separate CMVN apply kernel
    static const bool no_fs = getenv("MAFE_NO_FUSED_FRAMESUM") != nullptr;   // A/B switch: frame-mean pre-pass kernel
    const bool use_v6 = th->v6 && !force_v3 && !halfwarp;
    const bool fuse = use_v6 && cmvn && !no_tmem && !no_fuse;
    // frame-mean sums inside the persistent kernel (needs the fused CMVN variant; dither changes the samples per frame;
    // and a batch that amortises the lag_s sum-only items at the head of the queue; small batches keep the pre-pass)
    const bool fs = fuse && !no_fs && d.remove_frame_mean && d.dither == 0.f && th->cover4_dev != nullptr &&
                    b->n_tiles >= 12 * ctx->sm_count;
    MAFE_CUDA_CHECK(cudaMemsetAsync(b->queue_dev, 0, sizeof(int32_t), ctx->stream));
    if (d.remove_frame_mean) MAFE_CUDA_CHECK(cudaMemsetAsync(b->utt_sum_dev, 0, sizeof(double) * b->n_utts, ctx->stream));
    if (d.remove_frame_mean && !fs) {
      ProfScope ps(ctx, MAFE_PROF_FRAME_MEAN);
      if (b->n_utts >= 4 * ctx->sm_count) {
        // enough utterances to fill the machine: one CTA streams one utterance (no per-tile latency chain)
        if (wave_dtype == MAFE_WAVE_I16)
          frame_sum_utt_kernel<true><<<b->n_utts, 256, 0, ctx->stream>>>(Q, th->dev.cover, b->utt_sum_dev);
        else
          frame_sum_utt_kernel<false><<<b->n_utts, 256, 0, ctx->stream>>>(Q, th->dev.cover, b->utt_sum_dev);
      } else {
        const int pgrid = b->n_tiles;  // one CTA per tile
        if (wave_dtype == MAFE_WAVE_I16)
          frame_sum_baked_kernel<true><<<pgrid, 256, 0, ctx->stream>>>(Q, th->dev.cover, b->utt_sum_dev);
        else
          frame_sum_baked_kernel<false><<<pgrid, 256, 0, ctx->stream>>>(Q, th->dev.cover, b->utt_sum_dev);
      }
      MAFE_LAUNCH_CHECK(ctx);
    }
    if (cmvn) MAFE_CUDA_CHECK(cudaMemsetAsync(b->utt_stats_dev, 0, sizeof(double) * 2 * kV2Mels * b->n_utts, ctx->stream));
    if (use_v6) {
      const int grid3 = std::min(b->n_tiles, MAFE_V6_CTAS * ctx->sm_count);   // persistent: 3 CTAs per SM, dynamic tile queue
      if (b->cap_tile_recs < (size_t)b->n_tiles) {   // grow-only
        if (b->tile_recs_dev) MAFE_CUDA_CHECK(cudaFree(b->tile_recs_dev));
        b->tile_recs_dev = nullptr; b->cap_tile_recs = 0;
        const size_t cap = (size_t)b->n_tiles + (size_t)b->n_tiles / 4 + 16;
        MAFE_CUDA_CHECK(cudaMalloc(&b->tile_recs_dev, cap * (sizeof(TileInfo) + sizeof(SumRec))));
        b->cap_tile_recs = cap;
      }
      Q.tile_recs = b->tile_recs_dev;
      Q.sum_recs = (const unsigned char*)b->tile_recs_dev + b->cap_tile_recs * sizeof(TileInfo);
      // FS: the sums run lag_s items ahead of the transforms (same rule as the CMVN lag below)
      if (fs) Q.lag_s = (int)((b->max_utt_frames + kTileFrames - 1) / kTileFrames) + 3 * grid3 + 64;
      {
        ProfScope ps(ctx, MAFE_PROF_OTHER);
        const int pg = (b->n_tiles + 255) / 256;
        if (wave_dtype == MAFE_WAVE_I16) tile_prepare_kernel<true><<<pg, 256, 0, ctx->stream>>>(Q, (TileInfo*)b->tile_recs_dev, (SumRec*)Q.sum_recs);
        else tile_prepare_kernel<false><<<pg, 256, 0, ctx->stream>>>(Q, (TileInfo*)b->tile_recs_dev, (SumRec*)Q.sum_recs);
        MAFE_LAUNCH_CHECK(ctx);
      }
      if (fuse) {
        if (b->cap_utt_done < (size_t)b->n_utts) {   // grow-only; two counters per utterance (tiles finished, tiles summed)
          if (b->utt_done_dev) MAFE_CUDA_CHECK(cudaFree(b->utt_done_dev));
          b->utt_done_dev = nullptr; b->cap_utt_done = 0;
          const size_t cap = (size_t)b->n_utts + (size_t)b->n_utts / 4 + 16;
          MAFE_CUDA_CHECK(cudaMalloc((void**)&b->utt_done_dev, 2 * cap * sizeof(int32_t)));
          b->cap_utt_done = cap;
        }
        MAFE_CUDA_CHECK(cudaMemsetAsync(b->utt_done_dev, 0, sizeof(int32_t) * 2 * b->n_utts, ctx->stream));
        Q.utt_done = b->utt_done_dev;
        Q.fsum_done = b->utt_done_dev + b->n_utts;
        // the normalisation of a tile trails its computation by `lag` queue items: at least the longest utterance in tiles
        // (deadlock freedom), plus three rounds of the resident CTAs (a finished tile is published one iteration later; the
        // wait is then almost never entered).  ~1.4 k tiles = 14 MB of features + 28 MB of waveform in between: L2 resident
        Q.lag = (int)((b->max_utt_frames + kTileFrames - 1) / kTileFrames) + 3 * grid3 + 64;
        Q.mean_norm = d.utt_cmvn_mean; Q.std_norm = d.utt_cmvn_std;
      }
      ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
      if (fs) {
        if (wave_dtype == MAFE_WAVE_I16)
          fbank512_v6_kernel<true, true, true, true><<<grid3, kFastThreads, V6Smem::kTotal, ctx->stream>>>(Q, th->sweep6);
        else
          fbank512_v6_kernel<false, true, true, true><<<grid3, kFastThreads, V6Smem::kTotal, ctx->stream>>>(Q, th->sweep6);
        MAFE_LAUNCH_CHECK(ctx);
        return MAFE_OK;
      }
      if (fuse) {
        if (wave_dtype == MAFE_WAVE_I16)
          fbank512_v6_kernel<true, true, true><<<grid3, kFastThreads, V6Smem::kTotal, ctx->stream>>>(Q, th->sweep6);
        else
          fbank512_v6_kernel<false, true, true><<<grid3, kFastThreads, V6Smem::kTotal, ctx->stream>>>(Q, th->sweep6);
        MAFE_LAUNCH_CHECK(ctx);
        return MAFE_OK;   // features are final: no apply kernel
      }
      if (no_tmem) {
        if (wave_dtype == MAFE_WAVE_I16)
          fbank512_v6_kernel<true, false, false><<<grid3, kFastThreads, V6Smem::kTotal, ctx->stream>>>(Q, th->sweep6);
        else
          fbank512_v6_kernel<false, false, false><<<grid3, kFastThreads, V6Smem::kTotal, ctx->stream>>>(Q, th->sweep6);
      } else {
        if (wave_dtype == MAFE_WAVE_I16)
          fbank512_v6_kernel<true, true, false><<<grid3, kFastThreads, V6Smem::kTotal, ctx->stream>>>(Q, th->sweep6);
        else
          fbank512_v6_kernel<false, true, false><<<grid3, kFastThreads, V6Smem::kTotal, ctx->stream>>>(Q, th->sweep6);
      }
      MAFE_LAUNCH_CHECK(ctx);
    } else if (halfwarp && th->v5) {
      const int grid3 = std::min(b->n_tiles, 3 * ctx->sm_count);
      V2Params Q5 = Q;
      Q5.combine = th->comb5_dev;
      ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
      if (wave_dtype == MAFE_WAVE_I16)
        fbank512_v3_kernel<true, true, V5Sweep><<<grid3, kFastThreads, V5Smem::kTotal, ctx->stream>>>(Q5, th->sweep5);
      else
        fbank512_v3_kernel<false, true, V5Sweep><<<grid3, kFastThreads, V5Smem::kTotal, ctx->stream>>>(Q5, th->sweep5);
      MAFE_LAUNCH_CHECK(ctx);
    } else {
      const int grid3 = std::min(b->n_tiles, 3 * ctx->sm_count);   // persistent: 3 CTAs per SM, dynamic tile queue
      ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
      if (wave_dtype == MAFE_WAVE_I16)
        fbank512_v3_kernel<true><<<grid3, kFastThreads, V3Smem::kTotal, ctx->stream>>>(Q, th->sweep);
      else
        fbank512_v3_kernel<false><<<grid3, kFastThreads, V3Smem::kTotal, ctx->stream>>>(Q, th->sweep);
      MAFE_LAUNCH_CHECK(ctx);
    }
    if (cmvn) {
      ProfScope ps(ctx, MAFE_PROF_CMVN);
      cmvn_utt_apply_kernel<<<std::min(b->n_tiles, 8 * ctx->sm_count), 256, 0, ctx->stream>>>(
          out, b->tiles_dev, b->n_tiles, b->frame_offsets_dev, b->utt_stats_dev, d.utt_cmvn_mean, d.utt_cmvn_std);
      MAFE_LAUNCH_CHECK(ctx);
    }
    return MAFE_OK;
  }
  if (d.remove_frame_mean) {
    MAFE_CUDA_CHECK(cudaMemsetAsync(b->utt_sum_dev, 0, sizeof(double) * b->n_utts, ctx->stream));
    ProfScope ps(ctx, MAFE_PROF_FRAME_MEAN);
    frame_sum_fast_kernel<<<b->n_tiles, 256, 0, ctx->stream>>>(P, b->utt_sum_dev);
    MAFE_LAUNCH_CHECK(ctx);
  }
  {
    ProfScope ps(ctx, MAFE_PROF_FBANK_MAIN);
    fbank512_kernel<<<b->n_tiles, kFastThreads, FastSmem::kTotal, ctx->stream>>>(P);
    MAFE_LAUNCH_CHECK(ctx);
  }
  if (cmvn) return mafe_cmvn_utt(ctx, out, b->frame_offsets_dev, b->n_utts, d.n_mels, d.utt_cmvn_mean, d.utt_cmvn_std);
  return MAFE_OK;
}

}  // namespace mafe
