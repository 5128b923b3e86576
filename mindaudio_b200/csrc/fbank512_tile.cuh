// fbank512_tile.cuh -- pieces shared by the tile kernels of the conformer-geometry front-end (400-sample frames, hop
// 160, 512-point FFT, 80 mel filters; examples/conformer/dataset.py:117-168): kernel parameters, tile geometry, mbarrier
// + TMA bulk-copy helpers, the 256-point group transform, the frame-mean pre-pass kernels and the utterance-CMVN apply
// kernel.  Included by fbank512.cu; the main kernel lives in fbank512_v3.cuh.
#pragma once

namespace mafe {

constexpr int kV2Flen = 400;
constexpr int kV2Hop = 160;
constexpr int kV2Mels = 80;
constexpr int kV2Ylen = (kTileFrames - 1) * kV2Hop + kV2Flen;  // 5360 samples feed one tile
constexpr int kCwRow = 172;        // coverage-table row: 160 + 8 (a vector group may run over the end) rounded to 16 B
constexpr int kV2RawBytes = 21504;                              // (5360 + 1 prev) * 4 + alignment slack, 16 B multiple

struct V2Params {
  const void* wave;
  int64_t total_samples;  // length of the flat waveform array (elements)
  float wave_scale;
  const int64_t* sample_offsets;
  const int64_t* frame_offsets;
  const Tile* tiles;
  int n_tiles;
  const double* utt_sum;  // frame-mean accumulators (pre-pass)
  double* utt_stats;      // [n_utts][2][80] sum x, sum x^2 of the raw log-mel (or null)
  float pre_hi, pre_lo;
  int preemph_on, remove_mean;
  float dither;
  uint64_t seed;
  int log_kind;
  float log_arg;
  const float* window;   // [400]
  const float2* w512;    // [256]
  const float2* w256t;   // [16][16]
  const int* combine;    // [80]
  float* out;
  int* queue_head;               // dynamic work queue: next unclaimed tile index
  const void* tile_recs;         // v6: TileInfo[n_tiles], prepared per step by tile_prepare_kernel
  const void* sum_recs;          // v6 FS: SumRec[n_tiles]
  // v6 with the utterance CMVN applied inside the persistent kernel (fbank512_v6.cuh, FUSE):
  int* utt_done;                 // [n_utts] tiles of the utterance whose features and statistics are complete (zeroed per step)
  int lag;                       // the CTA that claims tile w also normalises tile w - lag; >= the longest utterance in tiles
  int mean_norm, std_norm;
  // v6 with the frame-mean sums accumulated inside the persistent kernel (FS): the pre-pass kernel is gone
  int lag_s;                     // the CTA that claims item w sums tile w and transforms tile w - lag_s (0: pre-pass sums)
  const float* cover4;           // [4][kCwRow] c'(r) = c(r) - preemph c(r + 1), copy k shifted by k entries
  double* utt_fsum;              // [n_utts] sum_s x[s] c'(s) (zeroed per step)
  int* fsum_done;                // [n_utts] tiles of the utterance whose samples have been summed (zeroed per step)
};

struct TileInfo {  // geometry of one work item, prepared by thread 0 one iteration ahead
  int64_t out_row;    // first output row (frame) of the tile
  int64_t s0;         // first sample of the tile inside its utterance
  int64_t cov_end, end_elem, base_elem;  // scalar patch-up range / element index of raw[0]
  int utt, nf, shift;
  float neg_mu;
  uint32_t bytes;     // size of the bulk copy (multiple of 16; 0: nothing can be bulk-copied)
  int T;              // frames of the tile's utterance
};
static_assert(sizeof(TileInfo) <= 64, "TileInfo slot");

struct SumRec {  // FS (fbank512_v6.cuh): what the frame-sum duty of one tile needs, prepared per step by tile_prepare_kernel
  int64_t ga;       // element index of the first 16 B-aligned sample group of the tile's periodic region
  int nvec;         // 16-byte groups from there on
  int r0;           // (sample index of ga inside the utterance) mod 160
  int n_head;       // samples of the periodic region before ga (< group size), n_tail: behind the last group
  int n_tail;
  int utt;
  int edge;         // first / last tile of an utterance: samples outside the periodic region exist (general rule)
};
static_assert(sizeof(SumRec) == 32, "SumRec slot");

// ---------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA bulk copy (SASS: SYNCS.*, UBLKCP)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// geometry of the bytes a tile needs from the flat waveform array
template <bool I16>
struct TileSrc {
  int64_t g0;        // element index of the "previous sample" slot (sample s0 - 1 of the utterance), may be -1
  int64_t ga_byte;   // 16 B aligned start of the bulk copy
  uint32_t bytes;    // bulk copy size (multiple of 16), 0 if nothing can be bulk-copied
  int shift;         // element index of g0 inside the raw buffer
  int64_t cov_end;   // first element NOT covered by the bulk copy (scalar patch-up from here)
  int64_t end_elem;  // one past the last element the tile needs
};

template <bool I16>
__device__ __forceinline__ TileSrc<I16> tile_src(const V2Params& P, const Tile tile, int64_t off, int T) {
  constexpr int ES = I16 ? 2 : 4;
  TileSrc<I16> r;
  const int nf = min(kTileFrames, T - tile.frame0);
  const int64_t s0 = (int64_t)tile.frame0 * kV2Hop;
  const int need = (nf - 1) * kV2Hop + kV2Flen;
  r.g0 = off + s0 - 1;
  r.end_elem = off + s0 + need;
  const int64_t first = r.g0 < 0 ? 0 : r.g0;
  r.ga_byte = (first * ES) & ~(int64_t)15;
  const int64_t total_bytes16 = (P.total_samples * ES) & ~(int64_t)15;
  int64_t gb = (r.end_elem * ES + 15) & ~(int64_t)15;
  if (gb > total_bytes16) gb = total_bytes16;
  r.bytes = gb > r.ga_byte ? (uint32_t)(gb - r.ga_byte) : 0u;
  r.shift = (int)(r.g0 - r.ga_byte / ES);  // -1 only when g0 == -1 (first tile of the first utterance)
  r.cov_end = r.bytes ? gb / ES : first;
  return r;
}

// Geometry of every work item, computed once per step for the whole batch (one thread per tile) after the frame-mean
// pre-pass: the persistent kernel's producer thread then only copies a 64-byte record and issues the bulk copy -- the
// 64-bit address arithmetic and the double division of the frame mean left its critical path.
template <bool I16>
__global__ void __launch_bounds__(256) tile_prepare_kernel(const V2Params P, TileInfo* __restrict__ recs, SumRec* __restrict__ srecs) {
  constexpr int ES = I16 ? 2 : 4;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n_tiles) return;
  const Tile tile = P.tiles[i];
  const int64_t off = P.sample_offsets[tile.utt];
  const int64_t fo0 = P.frame_offsets[tile.utt], fo1 = P.frame_offsets[tile.utt + 1];
  const int T = (int)(fo1 - fo0);
  const TileSrc<I16> src = tile_src<I16>(P, tile, off, T);
  TileInfo r;
  r.out_row = fo0 + tile.frame0;
  r.s0 = (int64_t)tile.frame0 * kV2Hop;
  r.cov_end = src.cov_end;
  r.end_elem = src.end_elem;
  r.base_elem = src.ga_byte / ES;
  r.utt = tile.utt;
  r.nf = min(kTileFrames, T - tile.frame0);
  r.shift = src.shift;
  // FS: the sums do not exist yet -- the record carries scale / (400 T); the persistent kernel turns it into -mu
  if (P.lag_s > 0) r.neg_mu = (float)((double)P.wave_scale / ((double)T * (double)kV2Flen));
  else r.neg_mu = P.remove_mean ? -(float)(P.utt_sum[tile.utt] / ((double)T * (double)kV2Flen)) : 0.f;
  r.bytes = src.bytes;
  r.T = T;
  recs[i] = r;
  if (P.lag_s > 0) {
    constexpr int V = I16 ? 8 : 4;
    const int64_t s_lo = r.s0;
    const int64_t framed_end = (int64_t)(T - 1) * kV2Hop + kV2Flen;
    const bool last = s_lo + (int64_t)kTileFrames * kV2Hop >= (int64_t)T * kV2Hop;
    const int64_t s_hi = last ? framed_end : s_lo + (int64_t)kTileFrames * kV2Hop;
    // periodic region of c'(s): frames q, q - 1, q - 2 of s and of s + 1 all exist
    int64_t f_lo = s_lo > 2 * kV2Hop ? s_lo : 2 * kV2Hop, f_hi = (int64_t)T * kV2Hop - 1;
    if (f_hi > s_hi) f_hi = s_hi;
    if (f_hi < f_lo) { f_lo = s_hi; f_hi = s_hi; }
    const int64_t ga0 = (off + f_lo + V - 1) / V * V;   // the flat array itself is 16 B aligned
    int64_t sa = ga0 - off;
    if (sa > f_hi) sa = f_hi;
    SumRec q;
    q.ga = off + sa;
    q.nvec = (int)((f_hi - sa) / V);
    q.r0 = (int)(sa % kV2Hop);
    q.n_head = (int)(sa - f_lo);
    q.n_tail = (int)(f_hi - (sa + (int64_t)q.nvec * V));
    q.utt = tile.utt;
    q.edge = (s_lo < f_lo || f_hi < s_hi) ? 1 : 0;
    srecs[i] = q;
  }
}

// 256-point transform of v by the 16-lane group; result (bin 2*(t+16kt)+HALF at slot[t+16kt])
__device__ __forceinline__ void fft256_group(cpx (&v)[16], float2* slot, const float2* s_w256, int t) {
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
    slot[t + 16 * kt] = make_float2(x.x, x.y);
  }
  if (t == 0) slot[256] = make_float2(u[fft16_pos(0)].x, u[fft16_pos(0)].y);   // output 0 again: partner of itself (v3 sweep)
}

template <bool I16>
__device__ __forceinline__ float raw_elem(const unsigned char* raw, int idx, float scale) {
  if (I16) return (float)reinterpret_cast<const int16_t*>(raw)[idx] * scale;
  return reinterpret_cast<const float*>(raw)[idx] * scale;
}

// ---------------------------------------------------------------------------------------------
// frame-mean pre-pass for the tile geometry: sum over all windowed frame entries of an utterance
//   = sum_s y[s] * c(s),  c(s) = sum of the window over the frames covering sample s.
// Inside an utterance c(s) only depends on s mod hop (table cw[160]); only the first / last tile of an
// utterance needs the general rule.  One CTA per tile, the tile owns hop*32 samples (+ the tail).
// ---------------------------------------------------------------------------------------------
template <bool I16>
__device__ __forceinline__ float gload(const void* wave, int64_t g, float scale) {
  if (I16) return (float)__ldg((const int16_t*)wave + g) * scale;
  return __ldg((const float*)wave + g) * scale;
}

template <bool I16>
__global__ void __launch_bounds__(256) frame_sum_baked_kernel(const V2Params P, const float* __restrict__ cw, double* utt_sum) {
  __shared__ float s_cw[kV2Hop];
  __shared__ double ws[8];
  const int tid = threadIdx.x;
  if (tid < kV2Hop) s_cw[tid] = cw[tid];
  __syncthreads();
  for (int ti = blockIdx.x; ti < P.n_tiles; ti += gridDim.x) {
    const Tile tile = P.tiles[ti];
    const uint32_t utt = (uint32_t)tile.utt;
    const int64_t off = P.sample_offsets[utt];
    const int T = (int)(P.frame_offsets[utt + 1] - P.frame_offsets[utt]);
    const int64_t s_lo = (int64_t)tile.frame0 * kV2Hop;
    float acc = 0.f;
    if (tile.frame0 >= 2 && tile.frame0 + kTileFrames < T && P.dither == 0.f && P.preemph_on) {
      // interior tile: 5120 owned samples, coverage = cw[s mod 160]
      const int64_t g = off + s_lo;
      int r = tid % kV2Hop;
#pragma unroll 4
      for (int i = tid; i < kTileFrames * kV2Hop; i += 256) {
        const float x = gload<I16>(P.wave, g + i, P.wave_scale);
        float xp = __shfl_up_sync(0xffffffffu, x, 1);
        if ((tid & 31) == 0) xp = gload<I16>(P.wave, g + i - 1, P.wave_scale);
        const float y = fmaf(-P.pre_lo, xp, fmaf(-P.pre_hi, xp, x));
        acc = fmaf(y, s_cw[r], acc);
        r += 256 - kV2Hop;           // (i + 256) mod 160
        if (r >= kV2Hop) r -= kV2Hop;
      }
    } else {
      const int64_t framed_end = (int64_t)(T - 1) * kV2Hop + kV2Flen;
      const int64_t s_hi = tile.frame0 + kTileFrames >= T ? framed_end : s_lo + (int64_t)kTileFrames * kV2Hop;
      for (int64_t s = s_lo + tid; s < s_hi; s += 256) {
        float v = gload<I16>(P.wave, off + s, P.wave_scale);
        if (P.dither != 0.f) v = fmaf(P.dither, dither_normal((uint64_t)s, utt, P.seed), v);
        if (P.preemph_on && s > 0) {
          float vp = gload<I16>(P.wave, off + s - 1, P.wave_scale);
          if (P.dither != 0.f) vp = fmaf(P.dither, dither_normal((uint64_t)(s - 1), utt, P.seed), vp);
          v = fmaf(-P.pre_lo, vp, fmaf(-P.pre_hi, vp, v));
        }
        const int t_hi = (int)min((int64_t)T - 1, s / kV2Hop);
        float c = 0.f;
        for (int tt = t_hi; tt >= 0; --tt) {
          const int64_t n = s - (int64_t)tt * kV2Hop;
          if (n >= kV2Flen) break;
          c += __ldg(&P.window[n]);
        }
        acc = fmaf(v, c, acc);
      }
    }
    double d = (double)acc;
    for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
    if ((tid & 31) == 0) ws[tid >> 5] = d;
    __syncthreads();
    if (tid == 0) {
      double t8 = 0.0;
      for (int w = 0; w < 8; ++w) t8 += ws[w];
      atomicAdd(&utt_sum[utt], t8);
    }
    __syncthreads();
  }
}

// Same sum with ONE CTA PER UTTERANCE (used when the batch has enough utterances to fill the machine): each CTA
// streams its whole utterance with deep load pipelining and issues a single atomic-free store.
template <bool I16>
__global__ void __launch_bounds__(256) frame_sum_utt_kernel(const V2Params P, const float* __restrict__ cw, double* utt_sum) {
  // The coverage table four times, copy k shifted by k entries (s_cw4[k][j] = cw[(j + k) mod 160]): a thread's 4 (8)
  // coefficients cw[r .. r+3] then are ONE aligned 16-byte load from copy r mod 4 -- lanes read consecutive quads.
  // (With a single table the lane stride of 4 words made every scalar load a 4-way bank conflict: ncu showed 70 % of
  // this kernel's shared-memory wavefronts were conflicts and the L1 data pipe, not HBM, at 81 %.)
  __shared__ __align__(16) float s_cw4[4][kCwRow];
  __shared__ double ws[8];
  const int tid = threadIdx.x;
  for (int i = tid; i < 4 * kCwRow; i += 256) s_cw4[i / kCwRow][i % kCwRow] = cw[(i % kCwRow + i / kCwRow) % kV2Hop];
  const float* s_cw = s_cw4[0];
  __syncthreads();
  const uint32_t utt = blockIdx.x;
  const int64_t off = P.sample_offsets[utt];
  const int T = (int)(P.frame_offsets[utt + 1] - P.frame_offsets[utt]);
  if (T <= 0) { if (tid == 0) utt_sum[utt] = 0.0; return; }
  const int64_t framed_end = (int64_t)(T - 1) * kV2Hop + kV2Flen;
  // interior [320, T*160): every sample is covered by frames q, q-1, (q-2) that all exist -> c = cw[s mod 160]
  const int64_t in_lo = 2 * kV2Hop, in_hi = (int64_t)T * kV2Hop;
  float acc = 0.f;
  auto general = [&](int64_t s) {
    float v = gload<I16>(P.wave, off + s, P.wave_scale);
    if (P.dither != 0.f) v = fmaf(P.dither, dither_normal((uint64_t)s, utt, P.seed), v);
    if (P.preemph_on && s > 0) {
      float vp = gload<I16>(P.wave, off + s - 1, P.wave_scale);
      if (P.dither != 0.f) vp = fmaf(P.dither, dither_normal((uint64_t)(s - 1), utt, P.seed), vp);
      v = fmaf(-P.pre_lo, vp, fmaf(-P.pre_hi, vp, v));
    }
    const int t_hi = (int)min((int64_t)T - 1, s / kV2Hop);
    float c = 0.f;
    for (int tt = t_hi; tt >= 0; --tt) {
      const int64_t n = s - (int64_t)tt * kV2Hop;
      if (n >= kV2Flen) break;
      c += __ldg(&P.window[n]);
    }
    acc = fmaf(v, c, acc);
  };
  if (P.dither == 0.f && P.preemph_on && in_hi > in_lo) {
    for (int64_t s = tid; s < in_lo; s += 256) general(s);
    for (int64_t s = in_hi + tid; s < framed_end; s += 256) general(s);
    // interior: 16-byte loads (V samples) from the first aligned sample on; the predecessor of a group is the last
    // sample of the previous lane's group (shuffle), lane 0 reloads it
    constexpr int V = I16 ? 8 : 4;
    const int64_t g = off;
    int64_t s_a = in_lo + (V - (int)((g + in_lo) % V)) % V;   // (g + s_a) % V == 0: the flat array is 16 B aligned
    if (s_a > in_hi) s_a = in_hi;
    const int64_t nvec = (in_hi - s_a) / V;
    const int64_t s_b = s_a + nvec * V;
    auto scalar = [&](int64_t s) {
      const float x = gload<I16>(P.wave, g + s, P.wave_scale);
      const float xp = gload<I16>(P.wave, g + s - 1, P.wave_scale);
      acc = fmaf(fmaf(-P.pre_lo, xp, fmaf(-P.pre_hi, xp, x)), s_cw[(int)(s % kV2Hop)], acc);
    };
    for (int64_t s = in_lo + tid; s < s_a; s += 256) scalar(s);
    for (int64_t s = s_b + tid; s < in_hi; s += 256) scalar(s);
    int r = (int)((s_a + (int64_t)V * tid) % kV2Hop);
    constexpr int rstep = (V * 256) % kV2Hop;   // a multiple of 4: r mod 4 is fixed per thread
    const float* cw_row = s_cw4[r & 3] - (r & 3);   // cw_row[r + i] = cw[(r + i) mod 160], 16 B aligned at r - (r & 3) ... see below
    const int64_t q_warp = tid & ~31;
#pragma unroll 4
    for (int64_t q0 = 0; q0 + q_warp < nvec; q0 += 256) {   // warp-uniform trip count: the shuffle stays convergent
      const int64_t q = q0 + tid;
      const bool ok = q < nvec;
      float x[V];
      if (I16) {
        int4 raw = make_int4(0, 0, 0, 0);
        if (ok) raw = __ldg(reinterpret_cast<const int4*>((const int16_t*)P.wave + g + s_a) + q);
        const int w4[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          x[2 * i % V] = (float)(int16_t)(w4[i] & 0xffff) * P.wave_scale;
          x[(2 * i + 1) % V] = (float)(int16_t)(w4[i] >> 16) * P.wave_scale;
        }
      } else {
        float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) raw = __ldg(reinterpret_cast<const float4*>((const float*)P.wave + g + s_a) + q);
        x[0] = raw.x * P.wave_scale; x[1] = raw.y * P.wave_scale; x[2] = raw.z * P.wave_scale; x[3 % V] = raw.w * P.wave_scale;
      }
      float xp = __shfl_up_sync(0xffffffffu, x[V - 1], 1);
      if ((tid & 31) == 0 && ok) xp = gload<I16>(P.wave, g + s_a + V * q - 1, P.wave_scale);
      if (ok) {
        float c[V];
#pragma unroll
        for (int i = 0; i < V; i += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(cw_row + r + i);
          c[i] = c4.x; c[i + 1] = c4.y; c[i + 2] = c4.z; c[i + 3] = c4.w;
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          acc = fmaf(fmaf(-P.pre_lo, xp, fmaf(-P.pre_hi, xp, x[i])), c[i], acc);
          xp = x[i];
        }
      }
      r += rstep;
      if (r >= kV2Hop) r -= kV2Hop;
    }
  } else {
    for (int64_t s = tid; s < framed_end; s += 256) general(s);
  }
  double d = (double)acc;
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if ((tid & 31) == 0) ws[tid >> 5] = d;
  __syncthreads();
  if (tid == 0) {
    double t8 = 0.0;
    for (int w = 0; w < 8; ++w) t8 += ws[w];
    utt_sum[utt] = t8;
  }
}

// ---------------------------------------------------------------------------------------------
// utterance CMVN from the fused statistics: x = (x - mean) / std, tile-parallel, float4
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 8) cmvn_utt_apply_kernel(float* __restrict__ feats, const Tile* __restrict__ tiles, int n_tiles,
                                                             const int64_t* __restrict__ frame_offsets,
                                                             const double* __restrict__ utt_stats, int mean_norm, int std_norm) {
  __shared__ float s_mean[kV2Mels], s_inv[kV2Mels];
  for (int ti = blockIdx.x; ti < n_tiles; ti += gridDim.x) {
    const Tile tile = tiles[ti];
    const int64_t fo = frame_offsets[tile.utt];
    const int T = (int)(frame_offsets[tile.utt + 1] - fo);
    const int nf = min(kTileFrames, T - tile.frame0);
    if (threadIdx.x < kV2Mels) {
      const double s1 = utt_stats[((size_t)tile.utt * 2) * kV2Mels + threadIdx.x];
      const double s2 = utt_stats[((size_t)tile.utt * 2 + 1) * kV2Mels + threadIdx.x];
      const double mean = s1 / T;
      double var = s2 / T - mean * mean;
      if (var < 0.0 || T == 1) var = 0.0;   // one frame: np.std is exactly 0 (the reference divides by it: nan / inf), whatever the rounding of s2
      s_mean[threadIdx.x] = mean_norm ? (float)mean : 0.f;
      s_inv[threadIdx.x] = std_norm ? (float)(1.0 / sqrt(var)) : 1.f;
    }
    __syncthreads();
    float4* p = reinterpret_cast<float4*>(feats + (fo + tile.frame0) * (int64_t)kV2Mels);
    for (int q = threadIdx.x; q < nf * (kV2Mels / 4); q += blockDim.x) {
      const int m = 4 * (q % (kV2Mels / 4));
      float4 v = p[q];
      v.x = (v.x - s_mean[m]) * s_inv[m];
      v.y = (v.y - s_mean[m + 1]) * s_inv[m + 1];
      v.z = (v.z - s_mean[m + 2]) * s_inv[m + 2];
      v.w = (v.w - s_mean[m + 3]) * s_inv[m + 3];
      p[q] = v;
    }
    __syncthreads();
  }
}

}  // namespace mafe
