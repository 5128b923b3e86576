// front2048.cuh -- n_fft = 2048 front-end (fastspeech2 / wavegrad preprocessing: melspectrogram n_fft 2048, win 1200,
// hop 300, 128 mel at 22.05 kHz, examples/fastspeech2/preprocess.py:50-64; the STFT of features.harmonic / hpss; any
// spectrum.stft / spectrogram / melspectrogram / fbank / mfcc call with n_fft = 2048 and hop <= 512).  Included by
// fbank512.cu.
//
// A frame PAIR (a, b) rides in one 2048-point complex transform a + i b, done by 128 threads with 16 points each:
//   2048 = 16 x 16 x 8:   n = 128 n1 + 8 n2 + n3,   k = k1 + 16 k2 + 256 k3
//   stage A  thread t = 8 n2 + n3:  16-point DFT over n1 (registers, packed FADD2 / FFMA2) -> x W2048^(t k1)
//   stage B  thread (k1, n3):       16-point DFT over n2                                   -> x W128^(n3 k2)
//   stage C  thread (k1, k2) x 2:   8-point DFT over n3                                    -> Z[k]
// with two exchanges through a 17 KB shared-memory scratch (layouts padded / swizzled: every 8-byte access of a
// half-warp is conflict free), then the pair separation X_a = Z[k] + conj Z[N-k], X_b = -i (Z[k] - conj Z[N-k]) (the
// window carries the 1/2) read back from the skewed Z array, and -- for the mel kinds -- a dense per-filter dot product
// over the two power rows (thread = filter).  A CTA = two such 128-thread groups (named barriers, nothing CTA-wide inside
// a pair); persistent CTAs claim half-tiles (16 frames) from a queue and stage their samples -- centre padding resolved
// while staging -- in shared memory once, so every sample is read from L2 / HBM once per half-tile instead of once per
// frame that covers it (n_fft / hop = 6.8 times).  Only the part of a frame that meets the window's support (whole rows of 128
// samples: 1280 of 2048 for win 1200) is staged.  The per-thread window entries live in registers.
#pragma once
#include "packed.cuh"

namespace mafe {

constexpr int kN2048 = 2048;
constexpr int kBins2048 = 1025;
constexpr int kHalfFrames2048 = 16;                 // frames per work item
constexpr int kMaxHop2048 = 512;
constexpr int kStage2048 = (kHalfFrames2048 - 1) * kMaxHop2048 + kN2048;   // 9728 samples
constexpr int kScr2048 = 2176;                      // float2 per group: 16 x 136 (stage A -> B), 8 x 257 (B -> C), 2048 x 17 / 16 (Z)
constexpr int kPRow2048 = 1028;                     // floats per power row

struct F2048Params {
  const void* wave;
  int wave_dtype;
  float wave_scale;
  const int64_t* sample_offsets;
  const int64_t* frame_offsets;
  const Tile* tiles;
  int n_tiles;
  int hop, center, pad_mode;
  int j0, j1;                 // rows j (n = t + 128 j) that meet the window's support
  const float* window;        // [2048], times 1/2 (pair separation) and spec_scale
  const float2* tw;           // [2048] W2048^j
  int out_kind;
  float power;
  int n_mels;
  const int* mstart;          // [n_mels] first bin the filter reads (multiple of 4)
  const int* mcount;          // [n_mels] 4-bin chunks its warp of 32 filters walks
  const int* moff;            // [n_mels] float4 index of that warp's weight block: chunk i of lane l at moff + 32 i + l
  const float* mweights;
  int log_kind;
  float log_arg, log_mult, log_offset;
  float* out;
  int out_dim;
  int* queue_head;
  int db_group;
  int* group_max;
  const int* utt_group;
  // shared-memory geometry (set by f2048_smem): staging buffers sized for THIS hop -- two of them when they fit 2 CTAs/SM,
  // so that the next half-tile's samples arrive (cp.async) while the current one is transformed --, mel weights behind
  int stage_floats, n_stage, mw_floats;
};

// dynamic shared memory: [n_stage x stage_floats] samples | scratch 2 x kScr2048 float2 | power rows 2 x 2 x kPRow2048 |
// mel weights mw_floats | work words
__host__ __device__ inline size_t f2048_off_scr(int stage_floats, int n_stage) { return sizeof(float) * (size_t)stage_floats * n_stage; }
__host__ __device__ inline size_t f2048_off_p(int stage_floats, int n_stage) { return f2048_off_scr(stage_floats, n_stage) + sizeof(float2) * 2 * kScr2048; }
__host__ __device__ inline size_t f2048_off_mw(int stage_floats, int n_stage) { return f2048_off_p(stage_floats, n_stage) + sizeof(float) * 2 * 2 * kPRow2048; }
__host__ __device__ inline size_t f2048_off_work(int stage_floats, int n_stage, int mw_floats) { return f2048_off_mw(stage_floats, n_stage) + sizeof(float) * (size_t)mw_floats; }
__host__ __device__ inline size_t f2048_smem_total(int stage_floats, int n_stage, int mw_floats) { return f2048_off_work(stage_floats, n_stage, mw_floats) + 32; }
constexpr size_t kF2048SmemBudget = (228 * 1024) / 2 - 1024;   // 2 CTAs per SM

// forward 8-point DFT, natural order in and out (radix-2 split + two 4-point DFTs)
__device__ __forceinline__ void dft8p(c2* v) {
  const float h = 0.70710678118654752440f;
  c2 a0 = add2(v[0], v[4]), a1 = add2(v[1], v[5]), a2 = add2(v[2], v[6]), a3 = add2(v[3], v[7]);
  c2 b0 = sub2(v[0], v[4]), b1 = sub2(v[1], v[5]), b2 = sub2(v[2], v[6]), b3 = sub2(v[3], v[7]);
  b1 = mul2(add2(b1, mni(b1)), bc(h));     // W8^1 = (h, -h)
  b2 = mni(b2);                            // W8^2 = -i
  b3 = mul2(sub2(mni(b3), b3), bc(h));     // W8^3 = (-h, -h)
  dft4p(a0, a1, a2, a3);
  dft4p(b0, b1, b2, b3);
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// named barrier of one 128-thread group.  The non-aligned form: after a loop whose trip count differs inside a warp
// (k = t, t + 128, ... < 1025) the lanes need not have reconverged when they arrive (synccheck flags bar.sync there).
__device__ __forceinline__ void group_bar(int g) {
  __syncwarp();
  asm volatile("barrier.sync %0, 128;" ::"r"(1 + g) : "memory");
}

// J0 >= 0: the rows [J0, J1) that meet the window's support are compile-time constants (win 1200: rows 3..12, full window:
// 0..15) -- the load loop has no predicates and the zero rows fold out of the first butterfly layer; J0 < 0: run-time range.
template <int J0, int J1>
__global__ void __launch_bounds__(256, 2) front2048_kernel(const F2048Params P) {
  extern __shared__ __align__(16) unsigned char smem[];
  float* stage_base = reinterpret_cast<float*>(smem);
  float* s_mw = reinterpret_cast<float*>(smem + f2048_off_mw(P.stage_floats, P.n_stage));
  int* s_work = reinterpret_cast<int*>(smem + f2048_off_work(P.stage_floats, P.n_stage, P.mw_floats));
  const int tid = threadIdx.x, g = tid >> 7, t = tid & 127;
  float2* scr = reinterpret_cast<float2*>(smem + f2048_off_scr(P.stage_floats, P.n_stage)) + g * kScr2048;
  // power rows of the group's pair, interleaved: prow[k] = (|X_a[k]|^p, |X_b[k]|^p) -- one 16-byte load = two ready operand pairs
  float2* prow = reinterpret_cast<float2*>(smem + f2048_off_p(P.stage_floats, P.n_stage)) + g * kPRow2048;
  for (int i = tid; i < P.mw_floats; i += 256) s_mw[i] = P.mweights[i];
  if (t < kPRow2048 - kBins2048) prow[kBins2048 + t] = make_float2(0.f, 0.f);   // the pad is read (times a zero weight) by the last filters
  const int j0 = J0 >= 0 ? J0 : P.j0, j1 = J0 >= 0 ? J1 : P.j1;

  // this thread's window entries w[t + 128 j]
  float win[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) win[j] = (j >= j0 && j < j1) ? P.window[t + 128 * j] * P.wave_scale : 0.f;   // the staged samples are raw
  // this thread's twiddles live in its TMEM lane (tcgen05.ld, 12-cycle latency, off the shared-memory pipe): columns 0..29
  // W2048^(t k1), k1 = 1..15 (stage A), columns 32..61 W128^(n3 k2), k2 = 1..15 (stage B).  Threads t and t + 128 (warps w
  // and w + 4) share a TMEM lane and need the same values.
  uint32_t* s_tm = reinterpret_cast<uint32_t*>(s_work) + 4;
  const int warp = tid >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(s_tm)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = *s_tm + ((uint32_t)(32 * (warp & 3)) << 16);
  if (warp < 4) {
    float c8[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k1 = 4 * c + 1 + i;
        const float2 w = k1 < 16 ? P.tw[(t * k1) & (kN2048 - 1)] : make_float2(0.f, 0.f);
        c8[2 * i] = w.x; c8[2 * i + 1] = w.y;
      }
      tm_st8(tb + 8 * c, c8);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int k2 = 4 * c + 1 + i;
        const float2 w = k2 < 16 ? P.tw[(16 * (t & 7) * k2) & (kN2048 - 1)] : make_float2(0.f, 0.f);
        c8[2 * i] = w.x; c8[2 * i + 1] = w.y;
      }
      tm_st8(tb + 32 + 8 * c, c8);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int n_mels = P.n_mels, out_dim = P.out_dim, log_kind = P.log_kind;
  const bool mel_kind = P.out_kind >= MAFE_OUT_MEL && t < n_mels;
  const int mel_s = mel_kind ? P.mstart[t] : 0, mel_n = mel_kind ? P.mcount[t] : 1, mel_o = mel_kind ? P.moff[t] : 0;
  const int hop = P.hop;
  const int lo = 128 * j0, span = 128 * (j1 - j0);   // staged part of a frame (window support, whole rows)
  const int n_items = 2 * P.n_tiles;   // half-tiles of 16 frames (the batch tiles hold 32)
  float vmax = -INFINITY;
  int vmax_utt = -1;
  auto flush_max = [&]() {   // dB output: group maxima for the top_db clamp (spectrum.py:78-89)
    if (vmax_utt < 0) return;
    float m = vmax;
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0 && m > -INFINITY) {
      const int grp = P.db_group == MAFE_DBGROUP_UTT ? vmax_utt : (P.db_group == MAFE_DBGROUP_BATCH ? 0 : P.utt_group[vmax_utt]);
      atomicMax(&P.group_max[grp], ordered_key(m));
    }
    vmax = -INFINITY;
  };
  const bool want_max = P.out_kind >= MAFE_OUT_MEL && P.log_kind == MAFE_LOG_DB && P.db_group != MAFE_DBGROUP_NONE;

  // geometry of a work item (half-tile)
  struct Item { int frame0, nf, T; int utt; int64_t off, L, fo, s_lo; int n_need, sh; bool ok, direct; };
  auto geometry = [&](int item) {
    Item it;
    it.ok = false; it.direct = false; it.frame0 = 0; it.nf = 0; it.T = 0; it.utt = 0; it.off = 0; it.L = 0; it.fo = 0; it.s_lo = 0; it.n_need = 0; it.sh = 0;
    if (item >= n_items) return it;
    const Tile tile = P.tiles[item >> 1];
    it.frame0 = tile.frame0 + kHalfFrames2048 * (item & 1);
    it.utt = tile.utt;
    it.off = P.sample_offsets[tile.utt];
    it.L = P.sample_offsets[tile.utt + 1] - it.off;
    it.fo = P.frame_offsets[tile.utt];
    it.T = (int)(P.frame_offsets[tile.utt + 1] - it.fo);
    if (it.frame0 >= it.T) return it;
    it.ok = true;
    it.nf = min(kHalfFrames2048, it.T - it.frame0);
    // only the rows that meet the window's support are staged: samples [lo, lo + span) of every frame
    it.s_lo = (int64_t)it.frame0 * hop - (P.center ? kN2048 / 2 : 0) + lo;
    it.n_need = (it.nf - 1) * hop + span;
    it.direct = it.s_lo >= 0 && it.s_lo + it.n_need <= it.L && P.wave_dtype == MAFE_WAVE_F32;   // plain float copy
    // plain copies land `sh` floats into the staging buffer, so that source and destination share their 16-byte phase
    if (it.direct) it.sh = (int)((reinterpret_cast<uintptr_t>((const float*)P.wave + it.off + it.s_lo) >> 2) & 3);
    return it;
  };
  auto stage_async = [&](const Item& it, float* dst) {   // float32 input; 4-byte cp.async: the source has no alignment guarantee
    const uint32_t d = smem_u32(dst);
    if (it.direct) {   // dst = buffer + it.sh: 16-byte copies between the (<= 3)-element head and tail
      const float* w = (const float*)P.wave + it.off + it.s_lo;
      const int head = min((4 - it.sh) & 3, it.n_need), n16 = (it.n_need - head) >> 2, tail0 = head + 4 * n16;
      if (tid < head) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u * tid), "l"(w + tid) : "memory");
      for (int q = tid; q < n16; q += 256)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 4u * (head + 4 * q)), "l"(w + head + 4 * q) : "memory");
      if (tid < it.n_need - tail0) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u * (tail0 + tid)), "l"(w + tail0 + tid) : "memory");
    } else {   // first / last half-tiles of an utterance: centre padding resolved per element, zeros stored directly
      const float* w = (const float*)P.wave + it.off;
      for (int i = tid; i < it.n_need; i += 256) {
        int64_t sidx = it.s_lo + i;
        if (sidx < 0 || sidx >= it.L) sidx = P.center ? pad_index_fast(sidx, it.L, P.pad_mode) : -1;
        if (sidx >= 0 && sidx < it.L) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u * i), "l"(w + sidx) : "memory");
        else dst[i] = 0.f;
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto stage_sync = [&](const Item& it, float* dst) {     // centre padding / PCM16 resolved while staging
    if (it.direct) {
      const float* w = (const float*)P.wave + it.off + it.s_lo;
#pragma unroll 1
      for (int i0 = tid; i0 < it.n_need; i0 += 13 * 256) {   // 13 independent loads in flight per thread and round
        float x[13];
#pragma unroll
        for (int r = 0; r < 13; ++r) x[r] = i0 + 256 * r < it.n_need ? __ldg(w + i0 + 256 * r) : 0.f;
#pragma unroll
        for (int r = 0; r < 13; ++r)
          if (i0 + 256 * r < it.n_need) dst[i0 + 256 * r] = x[r];
      }
      return;
    }
    for (int i = tid; i < it.n_need; i += 256) {
      int64_t sidx = it.s_lo + i;
      if (sidx < 0 || sidx >= it.L) sidx = P.center ? pad_index_fast(sidx, it.L, P.pad_mode) : -1;
      float v = 0.f;
      if (sidx >= 0 && sidx < it.L)
        v = P.wave_dtype == MAFE_WAVE_I16 ? (float)((const int16_t*)P.wave)[it.off + sidx] : ((const float*)P.wave)[it.off + sidx];
      dst[i] = v;
    }
  };
  const bool two = P.n_stage == 2 && P.wave_dtype == MAFE_WAVE_F32;   // double buffer: the next item is staged asynchronously
  if (tid == 0) s_work[0] = atomicAdd(P.queue_head, 1);
  __syncthreads();
  Item cur = geometry(s_work[0]);
  int cur_item = s_work[0];
  bool cur_async = false;
  if (two && cur.ok) { stage_async(cur, stage_base + cur.sh); cur_async = true; }
  for (uint32_t iter = 0; cur_item < n_items; ++iter) {
    const int buf = two ? (int)(iter & 1) : 0;
    float* stage = stage_base + (size_t)buf * P.stage_floats + cur.sh;
    if (tid == 0) s_work[1 + (iter & 1)] = atomicAdd(P.queue_head, 1);   // claim the next item
    __syncthreads();   // the other staging buffer (previous item) is no longer read; the claim is visible
    const int nxt_item = s_work[1 + (iter & 1)];
    const Item nxt = geometry(nxt_item);
    bool nxt_async = false;
    if (two && nxt.ok) { stage_async(nxt, stage_base + (size_t)(buf ^ 1) * P.stage_floats + nxt.sh); nxt_async = true; }
    if (cur.ok) {
      if (cur_async) {
        if (nxt_async) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
        stage_sync(cur, stage);
      }
    }
    __syncthreads();
    if (cur.ok) {
    const int frame0 = cur.frame0, nf = cur.nf;
    const int64_t fo = cur.fo;
    const uint32_t utt = (uint32_t)cur.utt;
    if (want_max && vmax_utt != (int)utt) { flush_max(); vmax_utt = (int)utt; }   // warp uniform: every thread sees the same utt

    // ---- the 8 frame pairs of the half-tile: group g takes pairs g, g + 2, g + 4, g + 6 ----
    for (int pair = g; pair < kHalfFrames2048 / 2; pair += 2) {
      const int fa = 2 * pair;                 // frame index inside the half-tile
      if (fa >= nf) break;
      const bool has_b = fa + 1 < nf;
      c2 v[16];
      {  // load + window; frame b = frame a + hop
        const float* ya = stage + fa * hop + t - lo;   // stage[i] = sample lo + i of the half-tile's first frame
        const float* yb = ya + (has_b ? hop : 0);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j >= j0 && j < j1) v[j] = mul2(pk(ya[128 * j], yb[128 * j]), bc(win[j]));
          else v[j] = pk(0.f, 0.f);
        }
      }
      // stage A: DFT over n1, twiddle W2048^(t k1), scratch [k1][t] (row stride 136)
      fft16p(v);
      sts_c2(scr + t, v[fft16_pos(0)]);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float tw[8];
        tm_ld8(tb + 8 * c, tw);
        tm_wait8(tw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int k1 = 4 * c + 1 + i;
          if (k1 < 16) sts_c2(scr + k1 * 136 + t, cmul(v[fft16_pos(k1)], tw[2 * i], tw[2 * i + 1]));
        }
      }
      group_bar(g);
      // stage B: thread (k1, n3): DFT over n2, twiddle W128^(n3 k2) = W2048^(16 n3 k2), scratch [n3][k1 16 + (k2 ^ 8 (k1 & 1))]
      {
        const int k1 = t >> 3, n3 = t & 7;
#pragma unroll
        for (int n2 = 0; n2 < 16; ++n2) v[n2] = lds_c2(scr + k1 * 136 + 8 * n2 + n3);
        group_bar(g);                          // everyone has read the stage-A layout: the scratch is re-used
        fft16p(v);
        // column k2 ^ 8 (k1 & 1): two bases, compile-time offsets (k2 < 8 -> dlo + k2, else dhi + k2 - 8)
        const int sw = (k1 & 1) << 3;
        float2* dlo = scr + n3 * 257 + k1 * 16 + sw;
        float2* dhi = scr + n3 * 257 + k1 * 16 + (sw ^ 8);
        sts_c2(dlo, v[fft16_pos(0)]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float tw[8];
          tm_ld8(tb + 32 + 8 * c, tw);
          tm_wait8(tw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k2 = 4 * c + 1 + i;
            if (k2 < 16) sts_c2(k2 < 8 ? dlo + k2 : dhi + (k2 - 8), cmul(v[fft16_pos(k2)], tw[2 * i], tw[2 * i + 1]));
          }
        }
      }
      group_bar(g);
      // stage C: combos (k1, k2) = cc >> 4, cc & 15 for cc = t, t + 128: DFT over n3 -> Z[k1 + 16 k2 + 256 k3], skewed k + k / 16
      {
        c2 w0[8], w1[8];
        const int k1a = t >> 4, k2 = t & 15, k1b = k1a + 8;
#pragma unroll
        for (int n3 = 0; n3 < 8; ++n3) {
          w0[n3] = lds_c2(scr + n3 * 257 + k1a * 16 + (k2 ^ ((k1a & 1) << 3)));
          w1[n3] = lds_c2(scr + n3 * 257 + k1b * 16 + (k2 ^ ((k1b & 1) << 3)));
        }
        group_bar(g);
        dft8p(w0);
        dft8p(w1);
        // skewed index k + k / 16 of k = k1 + 16 k2 + 256 k3 (k1 < 16): k1 + 17 k2 + 272 k3 -- one base, compile-time offsets
        float2* zb = scr + k1a + 17 * k2;
#pragma unroll
        for (int k3 = 0; k3 < 8; ++k3) {
          sts_c2(zb + 272 * k3, w0[k3]);
          sts_c2(zb + 272 * k3 + 8, w1[k3]);
        }
      }
      group_bar(g);
      // ---- pair separation, emit ----
      // X_a[k] = Z[k] + conj Z[N-k], X_b[k] = -i (Z[k] - conj Z[N-k]); thread t owns bins t + 128 i, i < 8, thread 0 also bin 1024
      auto sep = [&](int k, c2& xa, c2& xb) {
        const int kn = (kN2048 - k) & (kN2048 - 1);
        const c2 zk = lds_c2(scr + k + (k >> 4)), zn = cnj(lds_c2(scr + kn + (kn >> 4)));
        xa = add2(zk, zn);
        xb = mni(sub2(zk, zn));
      };
      auto powers = [&](c2 xa, c2 xb, float& pa, float& pb) {   // |X|^power (+ log1p of the deepspeech2 kind)
        const c2 qa = mul2(xa, xa), qb = mul2(xb, xb);
        pa = re(qa) + im(qa); pb = re(qb) + im(qb);
        if (P.power != 2.0f) {
          if (P.power == 1.0f) { pa = sqrtf(pa); pb = sqrtf(pb); }
          else { pa = powf(sqrtf(pa), P.power); pb = powf(sqrtf(pb), P.power); }
        }
      };
      const int64_t row_a = fo + frame0 + fa;
      if (P.out_kind == MAFE_OUT_COMPLEX) {
        c2* oa = reinterpret_cast<c2*>(P.out + row_a * P.out_dim);
        c2* ob = reinterpret_cast<c2*>(P.out + (row_a + 1) * P.out_dim);
        c2 xa[8], xb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) sep(t + 128 * i, xa[i], xb[i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          oa[t + 128 * i] = xa[i];
          if (has_b) ob[t + 128 * i] = xb[i];
        }
        if (t == 0) {
          sep(kBins2048 - 1, xa[0], xb[0]);
          oa[kBins2048 - 1] = xa[0];
          if (has_b) ob[kBins2048 - 1] = xb[0];
        }
      } else if (P.out_kind == MAFE_OUT_POWER) {
        float* oa = P.out + row_a * P.out_dim;
        float* ob = oa + P.out_dim;
        const int n_own = t == 0 ? 9 : 8;
#pragma unroll 1
        for (int i = 0; i < n_own; ++i) {        // rolled: the general power / log forms are long (they were 90 KB of code unrolled)
          const int k = t + 128 * i;
          c2 xa, xb;
          sep(k, xa, xb);
          float pa, pb;
          powers(xa, xb, pa, pb);
          if (P.log_kind == MAFE_LOG_LN_PLUS) {
            pa = P.log_arg == 1.0f ? log1pf(pa) : logf(pa + P.log_arg);
            pb = P.log_arg == 1.0f ? log1pf(pb) : logf(pb + P.log_arg);
          }
          oa[k] = pa;
          if (has_b) ob[k] = pb;
        }
      } else {
        if (P.power == 2.0f) {                   // the usual mel front-end: branch-free, all 16 loads of a thread in flight
          c2 xa[8], xb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) sep(t + 128 * i, xa[i], xb[i]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const c2 qa = mul2(xa[i], xa[i]), qb = mul2(xb[i], xb[i]);
            prow[t + 128 * i] = make_float2(re(qa) + im(qa), re(qb) + im(qb));
          }
          if (t == 0) {
            sep(kBins2048 - 1, xa[0], xb[0]);
            const c2 qa = mul2(xa[0], xa[0]), qb = mul2(xb[0], xb[0]);
            prow[kBins2048 - 1] = make_float2(re(qa) + im(qa), re(qb) + im(qb));
          }
        } else {
          const int n_own = t == 0 ? 9 : 8;
#pragma unroll 1
          for (int i = 0; i < n_own; ++i) {
            const int k = t + 128 * i;
            c2 xa, xb;
            sep(k, xa, xb);
            float pa, pb;
            powers(xa, xb, pa, pb);
            prow[k] = make_float2(pa, pb);
          }
        }
        group_bar(g);
        // mel projection: thread = filter, dense dot product over the filter's support for both frames: chunk i = 4 bins =
        // two 16-byte loads of (a, b) pairs against one 16-byte load of weights
        float* orow = P.out + row_a * out_dim;
        auto mel_filter = [&](int m, int s, int nch, int off4) {
          const float4* w4 = reinterpret_cast<const float4*>(s_mw) + off4 + (t & 31);
          const float4* r4 = reinterpret_cast<const float4*>(prow + s);
          c2 acc = pk(0.f, 0.f), acc1 = pk(0.f, 0.f);
          float4 ww = w4[0], p01 = r4[0], p23 = r4[1];      // nch >= 1; the next chunk's loads are issued before this chunk's FMAs
          for (int i = 1; i < nch; ++i) {
            const float4 nw = w4[32 * i], n01 = r4[2 * i], n23 = r4[2 * i + 1];
            acc = fma2(pk(p01.x, p01.y), bc(ww.x), acc);
            acc1 = fma2(pk(p01.z, p01.w), bc(ww.y), acc1);
            acc = fma2(pk(p23.x, p23.y), bc(ww.z), acc);
            acc1 = fma2(pk(p23.z, p23.w), bc(ww.w), acc1);
            ww = nw; p01 = n01; p23 = n23;
          }
          acc = fma2(pk(p01.x, p01.y), bc(ww.x), acc);
          acc1 = fma2(pk(p01.z, p01.w), bc(ww.y), acc1);
          acc = fma2(pk(p23.x, p23.y), bc(ww.z), acc);
          acc1 = fma2(pk(p23.z, p23.w), bc(ww.w), acc1);
          acc = add2(acc, acc1);
          float o[2] = {re(acc), im(acc)};
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            float x = o[f];
            switch (log_kind) {
              case MAFE_LOG_LN_EPS_IF_ZERO: x = logf(x == 0.f ? 2.220446049250313e-16f : x); break;
              case MAFE_LOG_LN_PLUS: x = logf(x + P.log_arg); break;
              case MAFE_LOG_DB: x = P.log_mult * log10f(fmaxf(x, P.log_arg)) - P.log_offset; break;
              default: break;
            }
            if (f == 0 || has_b) {
              orow[(int64_t)f * out_dim + m] = x;
              if (want_max) vmax = fmaxf(vmax, x);
            }
          }
        };
        // the first 128 filters: table entries in registers (loaded once per kernel, not once per pair); more: from the tables
        if (mel_kind) mel_filter(t, mel_s, mel_n, mel_o);
        for (int m = 128 + t; m < n_mels; m += 128) mel_filter(m, P.mstart[m], P.mcount[m], P.moff[m]);
      }
      // The next pair's stage-A stores go to the scratch, whose last readers (pair separation) all passed a barrier since:
      // the mel kinds (power rows -> barrier -> projection) need no barrier here, the spectrum kinds read Z until now.
#ifndef MAFE_F2048_KEEP_BAR7
      if (P.out_kind == MAFE_OUT_COMPLEX || P.out_kind == MAFE_OUT_POWER)
#endif
        group_bar(g);
    }
    }   // cur.ok
    cur = nxt; cur_item = nxt_item; cur_async = nxt_async;
  }
  if (want_max) flush_max();
  __syncthreads();   // every warp has issued its last TMEM load
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(*s_tm) : "memory");
}

}  // namespace mafe
