// fbank512_baked.cuh -- version 2 of the headline kernel, specialised for the conformer example's
// exact configuration (400-sample frames, hop 160, 512-point FFT, 80 Kaldi triangles;
// examples/conformer/dataset.py:117-168).  Included by fbank512.cu.
//
//  * persistent CTAs (2 per SM) striding over the tile table;
//  * the waveform tile of the NEXT work item is prefetched by the TMA engine
//    (cp.async.bulk global -> shared, mbarrier completion) while the current one is computed;
//  * frame geometry is compile time (no predicates in the load/fold code);
//  * the sparse mel projection is STRAIGHT-LINE code: which filters an FFT bin feeds (kF0, generated
//    by tools/gen_kaldi80_table.py) is baked in, only the weights are run-time data and sit in the
//    kernel-parameter constant bank, so an FFMA reads them as an immediate constant operand;
//  * per-utterance CMVN statistics (sum x, sum x^2 per mel bin) are accumulated on the fly.
#pragma once

namespace mafe {

#include "kaldi80_f0.inc"

constexpr int kV2Flen = 400;
constexpr int kV2Hop = 160;
constexpr int kV2Mels = 80;
constexpr int kV2Ylen = (kTileFrames - 1) * kV2Hop + kV2Flen;  // 5360 samples feed one tile
constexpr int kV2RawBytes = 21504;                              // (5360 + 1 prev) * 4 + alignment slack, 16 B multiple
constexpr int kV2StageStride = 84;                              // [frame][80] staging rows, padded (16 B multiple)

struct BakedWeights {  // kernel-parameter resident: c[0][..] operands
  float2 w[kBins];     // (w0, w1) of every FFT bin, pre-scaled by 1/4
};

struct SweepStep;
struct SweepHdr;

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
  const SweepStep* sweep_steps;  // [2][8][kMaxSteps]
  const SweepHdr* sweep_hdr;     // [8]
  float* out;
  int* queue_head;               // dynamic work queue: next unclaimed tile index
};

// table-driven sweep: one step per FFT bin of a (half, warp) range
constexpr int kMaxSteps = 18;
constexpr int kPlaneRows = kV2Mels + 2;  // filters -1 .. 80: guard rows take the out-of-range emits
struct SweepStep {
  float w0, w1;       // weights of filters cur / cur+1 (pre-scaled by 1/4)
  uint32_t offs;      // byte offset of Z[k] (low 16 bits) and Z[512-k] (high 16 bits) inside the pair's slot
  int nflush;         // filters to retire BEFORE this bin is accumulated
};
struct SweepHdr {
  int lo;             // first filter the warp emits
  int nsteps[2];      // bins of the even / odd half
  int tail[2];        // filters still to retire after the last bin of the half
  int pad[3];
};

struct TileInfo {  // geometry of one work item, prepared by thread 0 one iteration ahead
  int64_t out_row;    // first output row (frame) of the tile
  int64_t s0;         // first sample of the tile inside its utterance
  int64_t cov_end, end_elem, base_elem;  // scalar patch-up range / element index of raw[0]
  int utt, nf, shift;
  float neg_mu;
};
static_assert(sizeof(TileInfo) <= 64, "TileInfo slot");

// Shared-memory layout.  OCC3 = false: two raw (TMA landing) buffers, 2 CTAs/SM.  OCC3 = true: no dedicated raw
// buffer -- the next tile's waveform lands in the upper part of the Z region once the last sweep has read it
// (the lower part is the output staging), 62.8 KB per CTA -> 3 CTAs/SM.
template <bool OCC3>
struct V2SmemT {
  static constexpr size_t kRawBase = 0;
  static constexpr size_t kY = OCC3 ? 0 : 2 * kV2RawBytes;              // float[5472]: pre-emphasised tile, then the mel planes + zero row
  static constexpr size_t kYBytes = sizeof(float) * 5632;                // 5360 samples padded 16 per 320 (5616), planes + zero row (5445)
  static constexpr size_t kZ = kY + kYBytes;                             // float2[16][273], then the output staging
  static constexpr size_t kZBytes = sizeof(float2) * kPairs * kSlotStride;
  static constexpr size_t kRawInZ = kZBytes - kV2RawBytes;               // OCC3: raw landing zone = last 21 504 B of the Z region
  static constexpr size_t kWin = kZ + kZBytes;                           // float[400]
  static constexpr size_t kW512 = kWin + sizeof(float) * 400;            // float2[256]
  static constexpr size_t kW256 = kW512 + sizeof(float2) * 256;          // float2[256]
  static constexpr size_t kBar = kW256 + sizeof(float2) * 256;           // 2 mbarriers + 2 claimed indices
  static constexpr size_t kInfo = kBar + 32;                             // 2 x TileInfo
  static constexpr size_t kSteps = kInfo + 2 * 64;                       // SweepStep[2][8][kMaxSteps] (table sweep only)
  static constexpr size_t kHdr = kSteps + (OCC3 ? 0 : sizeof(SweepStep) * 2 * kFastWarps * kMaxSteps);  // SweepHdr[8]
  static constexpr size_t kTotal = kHdr + sizeof(SweepHdr) * kFastWarps;
};
typedef V2SmemT<false> V2Smem;
static_assert(V2SmemT<true>::kRawInZ % 128 == 0 && V2SmemT<true>::kRawInZ >= (kTileFrames * kV2StageStride + 6 * kV2Mels) * 4,
              "raw landing zone must not overlap the staging area");
static_assert(V2Smem::kZ % 16 == 0 && V2Smem::kBar % 8 == 0, "smem alignment");
static_assert(2 * kPlaneRows * kPlaneStride + kPlaneStride <= 5632 && kV2Ylen + 16 * (kV2Ylen / 320) <= 5632, "y buffer size");
static_assert((kTileFrames * kV2StageStride + 6 * kV2Mels) * 4 <= (int)V2Smem::kZBytes, "staging must fit in the Z buffer");

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

// ---------------------------------------------------------------------------------------------
// straight-line sparse mel sweep: warp W, bins k = 2*kk + HALF for kk in [16W, 16W+16) (+ Nyquist)
// ---------------------------------------------------------------------------------------------
template <int HALF>
__device__ __forceinline__ void emit_filter(float* plane, int lane, int m, float v) {
  float* dst = plane + m * kPlaneStride + lane;
  if (HALF == 0) *dst = v; else *dst += v;
}

template <int HALF, int CUR, int HI>
__device__ __forceinline__ void sweep_tail(float acc_lo, float acc_hi, float* plane, int lane) {
  if constexpr (CUR <= HI) {
    if constexpr (CUR >= 0 && CUR < kV2Mels) emit_filter<HALF>(plane, lane, CUR, acc_lo);
    sweep_tail<HALF, CUR + 1, HI>(acc_hi, 0.f, plane, lane);
  }
}

template <int HALF, int KK, int KK_END, int CUR, int HI>
__device__ __forceinline__ void sweep_step(float acc_lo, float acc_hi, const float2* zp, float sgn, float* plane, int lane,
                                           const BakedWeights& W) {
  if constexpr (KK == KK_END) {
    sweep_tail<HALF, CUR, HI>(acc_lo, acc_hi, plane, lane);
  } else {
    constexpr int k = 2 * KK + HALF;
    if constexpr (CUR < kF0[k]) {
      if constexpr (CUR >= 0 && CUR < kV2Mels) emit_filter<HALF>(plane, lane, CUR, acc_lo);
      sweep_step<HALF, KK, KK_END, CUR + 1, HI>(acc_hi, 0.f, zp, sgn, plane, lane, W);
    } else {
      constexpr int kr = HALF == 0 ? ((256 - KK) & 255) : (255 - KK);  // slot index of bin 512 - k
      const float2 zk = zp[KK & 255];
      const float2 zn = zp[kr];
      const float re = fmaf(sgn, zn.x, zk.x);
      const float im = fmaf(-sgn, zn.y, zk.y);
      const float pw = fmaf(re, re, im * im);
      acc_lo = fmaf(W.w[k].x, pw, acc_lo);
      acc_hi = fmaf(W.w[k].y, pw, acc_hi);
      sweep_step<HALF, KK + 1, KK_END, CUR, HI>(acc_lo, acc_hi, zp, sgn, plane, lane, W);
    }
  }
}

template <int HALF, int WARP>
__device__ __forceinline__ void sweep_warp(const float2* zp, float sgn, float* planes, int lane, const BakedWeights& W) {
  constexpr int k_lo = 32 * WARP;
  constexpr int k_hi = WARP == kFastWarps - 1 ? kBins : 32 * WARP + 32;
  constexpr int lo = kF0[k_lo], hi = kF0[k_hi - 1] + 1;
  constexpr int kk_end = 16 * WARP + 16 + ((HALF == 0 && WARP == kFastWarps - 1) ? 1 : 0);
  float* plane = planes + (WARP & 1) * (kPlaneRows * kPlaneStride);
  sweep_step<HALF, 16 * WARP, kk_end, lo, hi>(0.f, 0.f, zp, sgn, plane, lane, W);
}

template <int HALF>
__device__ __forceinline__ void sweep_dispatch(int warp, const float2* zp, float sgn, float* planes, int lane,
                                               const BakedWeights& W) {
  switch (warp) {
    case 0: sweep_warp<HALF, 0>(zp, sgn, planes, lane, W); break;
    case 1: sweep_warp<HALF, 1>(zp, sgn, planes, lane, W); break;
    case 2: sweep_warp<HALF, 2>(zp, sgn, planes, lane, W); break;
    case 3: sweep_warp<HALF, 3>(zp, sgn, planes, lane, W); break;
    case 4: sweep_warp<HALF, 4>(zp, sgn, planes, lane, W); break;
    case 5: sweep_warp<HALF, 5>(zp, sgn, planes, lane, W); break;
    case 6: sweep_warp<HALF, 6>(zp, sgn, planes, lane, W); break;
    default: sweep_warp<HALF, 7>(zp, sgn, planes, lane, W); break;
  }
}

// Table-driven sweep (one copy of the code for every warp: it stays resident in the instruction cache).
// Control flow depends only on the table, i.e. it is warp-uniform; the votes tell the compiler so.
template <int HALF>
__device__ __forceinline__ void sweep_table(const SweepStep* steps, int nsteps, int tail, const unsigned char* zp, float sgn,
                                            float* dst) {
  float acc_lo = 0.f, acc_hi = 0.f;
  for (int s = 0; s < nsteps; ++s) {
    const SweepStep st = steps[s];
    int nf = st.nflush;
    while (__any_sync(0xffffffffu, nf > 0)) {
      if (HALF == 0) *dst = acc_lo; else *dst += acc_lo;
      acc_lo = acc_hi; acc_hi = 0.f; dst += kPlaneStride; --nf;
    }
    const float2 zk = *reinterpret_cast<const float2*>(zp + (st.offs & 0xffffu));
    const float2 zn = *reinterpret_cast<const float2*>(zp + (st.offs >> 16));
    const float re = fmaf(sgn, zn.x, zk.x);
    const float im = fmaf(-sgn, zn.y, zk.y);
    const float pw = fmaf(re, re, im * im);
    acc_lo = fmaf(st.w0, pw, acc_lo);
    acc_hi = fmaf(st.w1, pw, acc_hi);
  }
  while (__any_sync(0xffffffffu, tail > 0)) {
    if (HALF == 0) *dst = acc_lo; else *dst += acc_lo;
    acc_lo = acc_hi; acc_hi = 0.f; dst += kPlaneStride; --tail;
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

template <bool I16, bool TABLE, bool OCC3>
__device__ __forceinline__ void fbank512_baked_body(const V2Params& P, const BakedWeights& W) {
  typedef V2SmemT<OCC3> V2Smem;
  extern __shared__ __align__(128) unsigned char smem[];
  auto raw_buf = [&](int b) -> unsigned char* {
    return OCC3 ? smem + V2Smem::kZ + V2Smem::kRawInZ : smem + (size_t)b * kV2RawBytes;
  };
  float* ybuf = reinterpret_cast<float*>(smem + V2Smem::kY);
  float* planes = ybuf;
  float2* Zs = reinterpret_cast<float2*>(smem + V2Smem::kZ);
  float* stage = reinterpret_cast<float*>(smem + V2Smem::kZ);
  float* s_win = reinterpret_cast<float*>(smem + V2Smem::kWin);
  float2* s_w512 = reinterpret_cast<float2*>(smem + V2Smem::kW512);
  float2* s_w256 = reinterpret_cast<float2*>(smem + V2Smem::kW256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + V2Smem::kBar);
  TileInfo* info = reinterpret_cast<TileInfo*>(smem + V2Smem::kInfo);
  SweepStep* s_steps = reinterpret_cast<SweepStep*>(smem + V2Smem::kSteps);
  SweepHdr* s_hdr = reinterpret_cast<SweepHdr*>(smem + V2Smem::kHdr);
  constexpr int ES = I16 ? 2 : 4;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kV2Flen; i += kFastThreads) s_win[i] = P.window[i];
  for (int i = tid; i < 256; i += kFastThreads) { s_w512[i] = P.w512[i]; s_w256[i] = P.w256t[i]; }
  if (TABLE) for (int i = tid; i < 2 * kFastWarps * kMaxSteps; i += kFastThreads) s_steps[i] = P.sweep_steps[i];
  if (tid < kFastWarps) s_hdr[tid] = P.sweep_hdr[tid];
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Thread 0 prepares every work item ONE ITERATION AHEAD, and it does so in STAGES spread over the
  // iteration: each stage only issues loads whose results are consumed after a later barrier, so the chain
  // claim (atomicAdd) -> tile -> offsets/mean -> TMA never stalls warp 0 (and with it the whole CTA).
  // Tiles are claimed from a global counter: uneven tiles (utterance tails) balance by themselves.
  int* s_work = reinterpret_cast<int*>(smem + V2Smem::kBar) + 4;  // [2] claimed tile index per buffer
  int nx_w = P.n_tiles;          // stage registers: only meaningful in thread 0
  Tile nx_tile = {0, 0};
  int64_t nx_off = 0, nx_fo0 = 0, nx_fo1 = 0;
  double nx_sum = 0.0;
  auto issue_tile = [&](int slot) {   // final stage: geometry + TMA for the tile claimed as nx_w
    const int T = (int)(nx_fo1 - nx_fo0);
    const TileSrc<I16> src = tile_src<I16>(P, nx_tile, nx_off, T);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (src.bytes) {
      mbar_expect_tx(&bars[slot], src.bytes);
      tma_bulk_g2s(raw_buf(slot), (const unsigned char*)P.wave + src.ga_byte, src.bytes, &bars[slot]);
    } else {
      mbar_arrive(&bars[slot]);
    }
    TileInfo ti_;
    ti_.out_row = nx_fo0 + nx_tile.frame0;
    ti_.s0 = (int64_t)nx_tile.frame0 * kV2Hop;
    ti_.cov_end = src.cov_end;
    ti_.end_elem = src.end_elem;
    ti_.base_elem = src.ga_byte / ES;
    ti_.utt = nx_tile.utt;
    ti_.nf = min(kTileFrames, T - nx_tile.frame0);
    ti_.shift = src.shift;
    ti_.neg_mu = P.remove_mean ? -(float)(nx_sum / ((double)T * (double)kV2Flen)) : 0.f;
    info[slot] = ti_;
  };
  auto load_offsets = [&]() {
    nx_off = P.sample_offsets[nx_tile.utt];
    nx_fo0 = P.frame_offsets[nx_tile.utt];
    nx_fo1 = P.frame_offsets[nx_tile.utt + 1];
    if (P.remove_mean) nx_sum = P.utt_sum[nx_tile.utt];
  };

  if (tid == 0) {   // prologue: the first tile, all stages back to back
    nx_w = atomicAdd(P.queue_head, 1);
    s_work[0] = nx_w;
    if (nx_w < P.n_tiles) { nx_tile = P.tiles[nx_w]; load_offsets(); issue_tile(0); }
  }
  __syncthreads();

  uint32_t phase0 = 0, phase1 = 0;
  int buf = 0;
  for (;; buf ^= 1) {
    if (s_work[buf] >= P.n_tiles) break;
    if (tid == 0) nx_w = atomicAdd(P.queue_head, 1);   // stage 1 (issue): claim the next tile
    const TileInfo cur = info[buf];
    const uint32_t utt = (uint32_t)cur.utt;
    const int nf = cur.nf;

    // wait for this tile's bytes
    if (buf == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1; } else { mbar_wait(&bars[1], phase1); phase1 ^= 1; }
    unsigned char* rb = raw_buf(buf);
    // scalar patch-up of what the 16 B-granular bulk copy could not cover (end of the flat array)
    if (cur.cov_end < cur.end_elem) {
      for (int64_t e = cur.cov_end + tid; e < cur.end_elem; e += kFastThreads) {
        if (I16) reinterpret_cast<int16_t*>(rb)[e - cur.base_elem] = ((const int16_t*)P.wave)[e];
        else reinterpret_cast<float*>(rb)[e - cur.base_elem] = ((const float*)P.wave)[e];
      }
      __syncthreads();
    }

    // ---- pass P: [dither] + pre-emphasis, raw -> ybuf (y[0] = x[0] at the start of an utterance) ----
    // Lanes touch consecutive words (one shared-memory wavefront per 32 samples); ybuf is PADDED by 16 floats per
    // 320 samples (pidx) so that the two frame pairs a warp folds sit 16 banks apart.
    {
      const int64_t s0 = cur.s0;
      const int need = (nf - 1) * kV2Hop + kV2Flen;
      const int sh = cur.shift + 1;  // raw index of sample s0
      int rem = tid, pad = 0;        // i mod 320, 16 * (i / 320)  (tid < 256 < 320)
      if (!I16 && nf == kTileFrames && s0 > 0 && P.dither == 0.f && P.preemph_on) {
        // interior tile, float input: each thread owns groups of 4 samples.  The 5 raw values a group needs
        // (4 samples + predecessor) start at raw float 4q + shift: two ALIGNED 16-byte loads + a tile-uniform
        // select -- conflict-free shared-memory traffic at 3.5 instructions per sample.
        const float4* r4 = reinterpret_cast<const float4*>(rb);
        const int shift = cur.shift;                // 0..3 here (s0 > 0)
        int rem80 = tid % 80, gpad = 16 * (tid / 80);   // q mod 80, 16 * (q / 80) for q = tid + 256 k
#pragma unroll
        for (int k = 0; k < (kV2Ylen / 4 + kFastThreads - 1) / kFastThreads; ++k) {
          const int q = tid + k * kFastThreads;
          if (q < kV2Ylen / 4) {
            const float4 A = r4[q], B = r4[q + 1];
            float x0, x1, x2, x3, x4;
            if (shift == 0) { x0 = A.x; x1 = A.y; x2 = A.z; x3 = A.w; x4 = B.x; }
            else if (shift == 1) { x0 = A.y; x1 = A.z; x2 = A.w; x3 = B.x; x4 = B.y; }
            else if (shift == 2) { x0 = A.z; x1 = A.w; x2 = B.x; x3 = B.y; x4 = B.z; }
            else { x0 = A.w; x1 = B.x; x2 = B.y; x3 = B.z; x4 = B.w; }
            x0 *= P.wave_scale; x1 *= P.wave_scale; x2 *= P.wave_scale; x3 *= P.wave_scale; x4 *= P.wave_scale;
            float4 y;
            y.x = fmaf(-P.pre_lo, x0, fmaf(-P.pre_hi, x0, x1));
            y.y = fmaf(-P.pre_lo, x1, fmaf(-P.pre_hi, x1, x2));
            y.z = fmaf(-P.pre_lo, x2, fmaf(-P.pre_hi, x2, x3));
            y.w = fmaf(-P.pre_lo, x3, fmaf(-P.pre_hi, x3, x4));
            *reinterpret_cast<float4*>(ybuf + 4 * q + gpad) = y;
          }
          rem80 += 16; gpad += 48;                  // q += 256 = 3 * 80 + 16
          if (rem80 >= 80) { rem80 -= 80; gpad += 16; }
        }
      } else if (nf == kTileFrames && s0 > 0 && P.dither == 0.f && P.preemph_on) {
        // interior tile (PCM16 input): lanes touch consecutive samples, predecessor through a shuffle
#pragma unroll
        for (int k = 0; k < (kV2Ylen + kFastThreads - 1) / kFastThreads; ++k) {   // fixed trip count: the shuffle stays convergent
          const int i = tid + k * kFastThreads;
          const bool ok = i < kV2Ylen;
          const float x = ok ? raw_elem<I16>(rb, sh + i, P.wave_scale) : 0.f;
          float xp = __shfl_up_sync(0xffffffffu, x, 1);
          if (lane == 0 && ok) xp = raw_elem<I16>(rb, sh + i - 1, P.wave_scale);
          if (ok) ybuf[i + pad] = fmaf(-P.pre_lo, xp, fmaf(-P.pre_hi, xp, x));
          rem += kFastThreads;
          if (rem >= 320) { rem -= 320; pad += 16; }
        }
      } else {
        for (int i = tid; i < kV2Ylen; i += kFastThreads) {
          float v = 0.f, vp = 0.f;
          if (i < need) {
            v = raw_elem<I16>(rb, sh + i, P.wave_scale);
            if (P.dither != 0.f) v = fmaf(P.dither, dither_normal((uint64_t)(s0 + i), utt, P.seed), v);
            if (i > 0 || s0 > 0) {
              vp = raw_elem<I16>(rb, sh + i - 1, P.wave_scale);
              if (P.dither != 0.f) vp = fmaf(P.dither, dither_normal((uint64_t)(s0 + i - 1), utt, P.seed), vp);
            }
          }
          ybuf[i + pad] = P.preemph_on ? fmaf(-P.pre_lo, vp, fmaf(-P.pre_hi, vp, v)) : v;
          rem += kFastThreads;
          if (rem >= 320) { rem -= 320; pad += 16; }
        }
      }
    }
    const float neg_mu = cur.neg_mu;
    __syncthreads();
    if (tid == 0) {   // stage 2: the claim has arrived during pass P -> publish it, fetch the tile record
      s_work[buf ^ 1] = nx_w;
      if (nx_w < P.n_tiles) nx_tile = P.tiles[nx_w];
    }

    // ---- phase F: frame pair -> registers, window, mean removal, radix-2 fold ----
    const int t = lane & 15;
    const int pair = warp * 2 + (lane >> 4);
    float2* slot = Zs + pair * kSlotStride;
    cpx v0[16], v1[16];
    {
      // pair p starts at padded index 336 p; sample n of frame a sits at n + 16 (n >= 320), of frame b (= a + 160)
      // at 160 + n + 16 (n >= 160): with n = t + 16 j these shifts depend on j only (compile time)
      const float* ya = ybuf + pair * 336;
      const float* yb = ya + kV2Hop;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = t + 16 * j;
        const float w = s_win[n];
        const int sb = j >= 10 ? 16 : 0;                       // frame b, index n
        const cpx lo = cx(fmaf(ya[n], w, neg_mu), fmaf(yb[n + sb], w, neg_mu));
        const float2 tw = s_w512[n];
        if (j < 9) {  // n + 256 < 400 for every lane exactly when j <= 8
          const float w2 = s_win[n + 256];
          const int sa2 = j >= 4 ? 16 : 0;                     // frame a, index n + 256 >= 320  <=>  j >= 4
          const cpx hi = cx(fmaf(ya[n + 256 + sa2], w2, neg_mu), fmaf(yb[n + 256 + 16], w2, neg_mu));   // frame b: 160 + n + 256 >= 320 always
          v0[j] = lo + hi;
          v1[j] = cmulf(lo - hi, cx(tw.x, tw.y));
        } else {
          v0[j] = lo;
          v1[j] = cmulf(lo, cx(tw.x, tw.y));
        }
      }
    }
    __syncthreads();  // ybuf is dead: it becomes the mel planes
    if (tid == 0 && nx_w < P.n_tiles) load_offsets();   // stage 3: offsets + frame-mean sum of the next tile's utterance
    if (tid >= kFastThreads - 32) planes[2 * kPlaneRows * kPlaneStride + lane] = 0.f;   // the all-zero row (see phase C1)

    const float2* zp = Zs + (lane >> 1) * kSlotStride;
    const float sgn = (lane & 1) ? -1.f : 1.f;
    const SweepHdr hdr = s_hdr[warp];
    float* plane_dst = planes + (warp & 1) * (kPlaneRows * kPlaneStride) + (hdr.lo + 1) * kPlaneStride + lane;
    fft256_group(v0, slot, s_w256, t);
    __syncthreads();
    // stage 4: everything has arrived during the FFT -> geometry into shared memory, TMA into the other raw
    // buffer (free since the previous tile's pass P); it lands while the rest of this tile is computed
    if (!OCC3 && tid == 0 && nx_w < P.n_tiles) issue_tile(buf ^ 1);
    if (TABLE) sweep_table<0>(s_steps + warp * kMaxSteps, hdr.nsteps[0], hdr.tail[0], (const unsigned char*)zp, sgn, plane_dst);
    else sweep_dispatch<0>(warp, zp, sgn, planes + kPlaneStride, lane, W);
    __syncthreads();
    fft256_group(v1, slot, s_w256, t);
    __syncthreads();
    if (TABLE) sweep_table<1>(s_steps + (kFastWarps + warp) * kMaxSteps, hdr.nsteps[1], hdr.tail[1], (const unsigned char*)zp, sgn, plane_dst);
    else sweep_dispatch<1>(warp, zp, sgn, planes + kPlaneStride, lane, W);
    __syncthreads();
    // OCC3: the Z region has been read for the last time -> the next tile's waveform may land in its upper part
    if (OCC3 && tid == 0 && nx_w < P.n_tiles) issue_tile(buf ^ 1);

    // ---- phase C1: combine the (<= 2) partial sums, log, stage [frame][80]; thread = (frame group g, filter m) ----
    // Branch-free: a filter with fewer than 2 contributing warps reads the all-zero row instead of a plane row.
    float* part = stage + kTileFrames * kV2StageStride;  // [3][2][80] per-group CMVN partial sums
    if (tid < 3 * kV2Mels) {
      const int g = tid / kV2Mels, m = tid - g * kV2Mels;
      const int c = P.combine[m];
      const int n = c & 3, p0 = (c >> 2) & 1;
      const float* zero_row = planes + 2 * kPlaneRows * kPlaneStride;
      const float* pa = n >= 1 ? planes + p0 * (kPlaneRows * kPlaneStride) + (m + 1) * kPlaneStride : zero_row;
      const float* pb = n == 2 ? planes + (p0 ^ 1) * (kPlaneRows * kPlaneStride) + (m + 1) * kPlaneStride : zero_row;
      float s1 = 0.f, s2 = 0.f;
      auto body = [&](auto log_fn) {
        // frames g, g + 3, ..., g + 30: fixed trip count, immediate offsets; only the last step can leave the tile
        // (f = 32 reads the pad column of the plane row and is discarded)
        const float* qa = pa + g;
        const float* qb = pb + g;
        float* sd = stage + g * kV2StageStride + m;
        const int left = nf - g;
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          const float o = log_fn(qa[3 * i] + qb[3 * i]);
          if (i < 10 || g < 2) sd[3 * i * kV2StageStride] = o;
          const float ov = 3 * i < left ? o : 0.f;
          s1 += ov;
          s2 = fmaf(ov, ov, s2);
        }
      };
      // ln x = lg2 x * ln 2 with the raw MUFU (mel energies are never subnormal: 0 is replaced by DBL_EPSILON)
      auto fast_ln = [](float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r * 0.69314718055994530942f; };
      if (P.log_kind == MAFE_LOG_LN_EPS_IF_ZERO) body([&](float a) { return fast_ln(a == 0.f ? 2.220446049250313e-16f : a); });
      else if (P.log_kind == MAFE_LOG_LN_PLUS) body([&](float a) { return fast_ln(a + P.log_arg); });
      else body([](float a) { return a; });
      part[(g * 2) * kV2Mels + m] = s1;
      part[(g * 2 + 1) * kV2Mels + m] = s2;
    }
    __syncthreads();

    // ---- phase C2: coalesced float4 stores + per-utterance CMVN statistics ----
    {
      float4* dst = reinterpret_cast<float4*>(P.out + cur.out_row * (int64_t)kV2Mels);
      const int total4 = nf * (kV2Mels / 4);
      for (int q = tid; q < total4; q += kFastThreads) {
        const int f = q / (kV2Mels / 4), m4 = q - f * (kV2Mels / 4);
        dst[q] = *reinterpret_cast<const float4*>(stage + f * kV2StageStride + 4 * m4);
      }
      if (P.utt_stats != nullptr && tid < 2 * kV2Mels) {
        const int m = tid % kV2Mels, which = tid / kV2Mels;
        const double s = (double)part[which * kV2Mels + m] + (double)part[(2 + which) * kV2Mels + m] +
                         (double)part[(4 + which) * kV2Mels + m];
        atomicAdd(&P.utt_stats[((size_t)utt * 2 + which) * kV2Mels + m], s);
      }
    }
    __syncthreads();  // stage (Z) and planes (y) are rewritten by the next iteration
  }
}

template <bool I16, bool TABLE>
__global__ void __launch_bounds__(kFastThreads, 2) fbank512_baked_kernel(const V2Params P, const BakedWeights W) {
  fbank512_baked_body<I16, TABLE, false>(P, W);
}

// 3 CTAs per SM: 80 registers/thread (a few dozen spills), 62.8 KB of shared memory per CTA
template <bool I16>
__global__ void __launch_bounds__(kFastThreads, 3) fbank512_baked_occ3_kernel(const V2Params P, const BakedWeights W) {
  fbank512_baked_body<I16, false, true>(P, W);
}

// ---------------------------------------------------------------------------------------------
// frame-mean pre-pass for the baked geometry: sum over all windowed frame entries of an utterance
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
  __shared__ float s_cw[kV2Hop];
  __shared__ double ws[8];
  const int tid = threadIdx.x;
  if (tid < kV2Hop) s_cw[tid] = cw[tid];
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
    int r = (int)((in_lo + tid) % kV2Hop);
    const int64_t g = off;
#pragma unroll 8
    for (int64_t s = in_lo + tid; s < in_hi; s += 256) {
      const float x = gload<I16>(P.wave, g + s, P.wave_scale);
      float xp = __shfl_up_sync(0xffffffffu, x, 1);
      if ((tid & 31) == 0) xp = gload<I16>(P.wave, g + s - 1, P.wave_scale);
      acc = fmaf(fmaf(-P.pre_lo, xp, fmaf(-P.pre_hi, xp, x)), s_cw[r], acc);
      r += 256 - kV2Hop;
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
__global__ void __launch_bounds__(256) cmvn_utt_apply_kernel(float* __restrict__ feats, const Tile* __restrict__ tiles, int n_tiles,
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
      if (var < 0.0) var = 0.0;
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
