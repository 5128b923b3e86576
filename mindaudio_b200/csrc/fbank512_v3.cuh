// fbank512_v3.cuh -- version 3 of the headline kernel (conformer front-end: 400-sample frames, hop 160, 512-point
// FFT, 80 mel triangles; examples/conformer/dataset.py:117-168).  Included by fbank512.cu after fbank512_tile.cuh,
// whose tile geometry / TMA / FFT helpers it shares.
//
// Why a third version: ncu's source view of v2 (profiles/r01_fbank512_v2occ3_segments.txt) charged 28 % of all stall samples to
// `no_instruction`: v2's code is 93 KB (every warp ran its own straight-line mel sweep, both 256-point transforms and
// five unrolled copies of pass P were inlined) against a 32 KB L1.5 instruction cache, and the 80-register cap of
// 3 CTAs/SM spilled the second transform's inputs.  v3 is the same algorithm with ONE copy of each phase:
//   * the two 256-point groups (even / odd bins) run through ONE rolled loop  load + fold -> FFT -> sweep, so that code
//     exists once and only 16 complex values are live in registers; the pre-emphasised tile stays in shared memory
//     until the second group has been loaded, and the mel planes get their own compact region (one row per
//     (warp, emitted filter): 97 rows instead of two full planes) so that 3 CTAs/SM still fit;
//   * the mel sweep is a rolled loop over the warp's bins; its program (two weights + retire count per bin) sits in
//     the kernel-parameter constant bank and is read with warp-uniform LDCs -- any filterbank with at most two
//     adjacent filters per bin runs, nothing is baked at compile time;
//   * pass P is a rolled loop specialised on the tile's alignment shift.
#pragma once

namespace mafe {

constexpr int kV3HalfStride = 130;   // >= 129 sub-transform outputs (bins 2kk + g) of one half
constexpr int kV3Runs = 32;          // >= filters one warp emits
// Sweep program: warp w owns the sub-transform outputs kk0[w] .. kk0[w+1]-1 of both halves (g = 0: FFT bins 2kk incl.
// the Nyquist bin for the last warp; g = 1: bins 2kk+1).  A lane keeps two accumulators, filters (cur, cur + 1); step
// (g, kk) first RETIRES nret filters (stores acc of filter cur to its plane row, cur advances), then accumulates the
// bin with weights (w0, w1); after the warp's last bin tail[g][w] more filters are retired.  The split is cost balanced
// on the host (build_v3_program); everything is data independent.
struct __align__(16) V3Step {
  float w0, w1;   // weights of filters cur / cur + 1, pre-scaled by 1/4 (the pair separation leaves 2 X)
  int nret;
  int pad;
};
constexpr int kV3MaskWords = 4;      // 16 steps x 2 bits per word: a warp range may hold up to 64 bins
struct V3Sweep {  // kernel-parameter resident (constant bank 0)
  V3Step step[2 * kV3HalfStride];
  uint32_t nret_mask[2 * kFastWarps][kV3MaskWords];   // the retire counts of a (half, warp) range, 2 bits per step: kept in
                                                      // a register so that the retire branch does not wait for a load
  unsigned char tail[2][kFastWarps];
  unsigned char kk0[kFastWarps + 1];
  unsigned char row0[kFastWarps];      // first plane row of a warp: it emits filters lo .. hi into consecutive rows
  int zero_row;                        // an all-zero plane row (filters with < 2 contributing warps read it)
};
// Half-warp variant (experimental): 16 bin ranges, one per HALF-warp; a lane carries both frames of a pair.  A filter may
// be emitted by up to three ranges (comb5[m] = rowA | rowB << 8 | rowC << 16).
constexpr int kV5Ranges = 2 * kFastWarps;
constexpr int kV5PlaneRows = 116;     // 80 + 2 guards + <= 2 per range boundary + zero row
struct V5Sweep {
  V3Step step[2 * kV3HalfStride];
  uint32_t nret_mask[2 * kV5Ranges][2];   // (half, range): ranges hold <= 32 steps
  unsigned char tail[2][kV5Ranges];
  unsigned char kk0[kV5Ranges + 1];
  unsigned char row0[kV5Ranges];
  int zero_row;
};
static_assert(sizeof(V5Sweep) + sizeof(V2Params) < 8000, "kernel parameters");
static_assert(sizeof(V3Sweep) + sizeof(V2Params) < 8000, "kernel parameters (large-parameter ABI, <= 32764 B)");

constexpr int kV3PlaneRows = 98;       // sum over warps of emitted filters (80 + 2 guards + <= 2 shared per boundary) + zero row
struct V3Smem {
  static constexpr size_t kY = 0;                                         // float[5632]: pre-emphasised tile (padded 16 per 320)
  static constexpr size_t kZ = kY + sizeof(float) * 5632;                 // float2[16][273]; upper part = TMA landing zone
  static constexpr size_t kZBytes = sizeof(float2) * kPairs * kSlotStride;
  static constexpr size_t kRawInZ = kZBytes - kV2RawBytes;
  static constexpr size_t kRaw = kZ + kRawInZ;
  static constexpr size_t kPlanes = kZ + kZBytes;                         // float[98][33]
  static constexpr size_t kWin = kPlanes + ((sizeof(float) * kV3PlaneRows * kPlaneStride + 15) & ~(size_t)15);
  static constexpr size_t kW512 = kWin + sizeof(float) * 400;
  static constexpr size_t kW256 = kW512 + sizeof(float2) * 256;
  static constexpr size_t kBar = kW256 + sizeof(float2) * 256;            // 2 mbarriers + 2 claimed indices
  static constexpr size_t kInfo = kBar + 32;                              // 2 x TileInfo
  static constexpr size_t kTotal = kInfo + 2 * 64;
};
static_assert(3 * (V3Smem::kTotal + 1024) <= 228 * 1024, "3 CTAs per SM");
// Half-warp variant: more plane rows (16 ranges), the W512 twiddles stay in global memory (L1) to make room.
struct V5Smem {
  static constexpr size_t kY = 0;
  static constexpr size_t kZ = kY + sizeof(float) * 5632;
  static constexpr size_t kZBytes = sizeof(float2) * kPairs * kSlotStride;
  static constexpr size_t kRawInZ = kZBytes - kV2RawBytes;
  static constexpr size_t kRaw = kZ + kRawInZ;
  static constexpr size_t kPlanes = kZ + kZBytes;
  static constexpr size_t kWin = kPlanes + ((sizeof(float) * kV5PlaneRows * kPlaneStride + 15) & ~(size_t)15);
  static constexpr size_t kW512 = kWin + sizeof(float) * 400;            // (not staged)
  static constexpr size_t kW256 = kW512;
  static constexpr size_t kBar = kW256 + sizeof(float2) * 256;
  static constexpr size_t kInfo = kBar + 32;
  static constexpr size_t kTotal = kInfo + 2 * 64;
};
static_assert(3 * (V5Smem::kTotal + 1024) <= 228 * 1024, "3 CTAs per SM (half-warp variant)");
static_assert(V5Smem::kRaw % 128 == 0 && V5Smem::kBar % 8 == 0 && V5Smem::kWin % 16 == 0, "smem alignment");
static_assert(V3Smem::kRawInZ % 128 == 0 && V3Smem::kRawInZ >= 6 * kV2Mels * 4, "landing zone vs the CMVN partial sums");
static_assert(V3Smem::kZ % 16 == 0 && V3Smem::kBar % 8 == 0 && V3Smem::kWin % 16 == 0, "smem alignment");

// One copy for every warp and both halves.  FFT bin k = 2 kk + g sits at sub-index kk of the group's slot, its conjugate
// partner 512 - k at sub-index 256 - g - kk of the same sub-transform (fft256_group duplicates output 0 at 256).
// The loop is written on 32-bit shared addresses with opaque increments: left to itself the compiler rewrites the
// pointer updates into closed-form exit values and pays ~10 extra instructions per run for it.
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
template <int D>
__device__ __forceinline__ void bump(uint32_t& a) { asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "n"(D)); }
template <int D>
__device__ __forceinline__ void bump(int& a) { asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "n"(D)); }

__device__ __forceinline__ void sweep_v3(const V3Sweep& S, int g, int warp, const float2* __restrict__ zp, float sgn,
                                         float* __restrict__ dst) {
  const int kk0 = S.kk0[warp];
  uint32_t ak = smem_u32(zp + kk0);
  uint32_t an = smem_u32(zp + (256 - g) - kk0);
  int si = g * kV3HalfStride + kk0;
  const int si_end = g * kV3HalfStride + S.kk0[warp + 1] + ((g == 0 && warp == kFastWarps - 1) ? 1 : 0);
  float acc_lo = 0.f, acc_hi = 0.f;
  auto retire = [&](int n) {
#pragma unroll 1
    do {
      float v = acc_lo;
      if (g) v += *dst;
      *dst = v;
      acc_lo = acc_hi; acc_hi = 0.f; dst += kPlaneStride;
    } while (--n);
  };
  const int gw = g * kFastWarps + warp;
  int word = 0;
#pragma unroll 1
  do {
    uint32_t m = S.nret_mask[gw][word++];   // retire counts of the next 16 steps
    const int chunk_end = min(si + 16, si_end);
#pragma unroll 1
    do {
      const float2 w = *reinterpret_cast<const float2*>(&S.step[si].w0);
      const int nr = m & 3u;
      m >>= 2;
      if (nr) retire(nr);
      const float2 zk = lds_f2(ak);
      const float2 zn = lds_f2(an);
      bump<1>(si); bump<8>(ak); bump<-8>(an);
      const float re = fmaf(sgn, zn.x, zk.x);
      const float im = fmaf(-sgn, zn.y, zk.y);
      const float pw = fmaf(re, re, im * im);
      acc_lo = fmaf(w.x, pw, acc_lo);
      acc_hi = fmaf(w.y, pw, acc_hi);
    } while (si != chunk_end);
  } while (si != si_end);
  retire(S.tail[g][warp]);   // >= 1: the last filter is always retired after the last bin
}

// Half-warp sweep: half-warp hw = 2 warp + (lane >> 4) owns the sub-range kk0[hw] .. kk0[hw + 1] - 1 of the current group
// g; its 16 lanes are the 16 frame pairs and carry BOTH frames (plane columns 2t, 2t + 1), so the two loads, the weights
// and the loop control of a step serve two frames.  The halves of a warp run their own (cost-balanced) programs: the
// loops diverge, nothing is shared between them.
__device__ __forceinline__ void sweep_v5(const V5Sweep& S, int g, int warp, int lane, const float2* __restrict__ Zs,
                                         float* __restrict__ planes) {
  const int hw = 2 * warp + (lane >> 4), t = lane & 15;
  const float2* zp = Zs + t * kSlotStride;
  const int kk0 = S.kk0[hw];
  uint32_t ak = smem_u32(zp + kk0);
  uint32_t an = smem_u32(zp + (256 - g) - kk0);
  int si = g * kV3HalfStride + kk0;
  const int si_end = g * kV3HalfStride + S.kk0[hw + 1] + ((g == 0 && hw == kV5Ranges - 1) ? 1 : 0);
  float* dst = planes + (int)S.row0[hw] * kPlaneStride + 2 * t;
  float a_lo = 0.f, a_hi = 0.f, b_lo = 0.f, b_hi = 0.f;
  auto retire = [&](int n) {
#pragma unroll 1
    do {
      float va = a_lo, vb = b_lo;
      if (g) { va += dst[0]; vb += dst[1]; }
      dst[0] = va;
      dst[1] = vb;
      a_lo = a_hi; a_hi = 0.f; b_lo = b_hi; b_hi = 0.f;
      dst += kPlaneStride;
    } while (--n);
  };
  const int gw = g * kV5Ranges + hw;
  int word = 0;
#pragma unroll 1
  do {
    uint32_t m = S.nret_mask[gw][word++];
    const int chunk_end = min(si + 16, si_end);
#pragma unroll 1
    do {
      const float2 w = *reinterpret_cast<const float2*>(&S.step[si].w0);
      const int nr = m & 3u;
      m >>= 2;
      if (nr) retire(nr);
      const float2 zk = lds_f2(ak);
      const float2 zn = lds_f2(an);
      bump<1>(si); bump<8>(ak); bump<-8>(an);
      const float ra = zk.x + zn.x, ia = zk.y - zn.y;
      const float rb = zk.x - zn.x, ib = zk.y + zn.y;
      const float pa = fmaf(ra, ra, ia * ia);
      const float pb = fmaf(rb, rb, ib * ib);
      a_lo = fmaf(w.x, pa, a_lo);
      a_hi = fmaf(w.y, pa, a_hi);
      b_lo = fmaf(w.x, pb, b_lo);
      b_hi = fmaf(w.y, pb, b_hi);
    } while (si != chunk_end);
  } while (si != si_end);
  retire((int)S.tail[g][hw]);
}

// frame pair -> registers for group g: window, mean removal, radix-2 fold.  g = 0: even bins (lo + hi);
// g = 1: odd bins ((lo - hi) W512^n).  Pair p starts at padded index 336 p of ybuf; sample n of frame a sits at
// n + 16 (n >= 320), of frame b (= a + 160) at 160 + n + 16 (n >= 160): with n = t + 16 j the shifts depend on j only.
__device__ __forceinline__ void load_fold_v3(cpx (&v)[16], const float* __restrict__ ya, const float* __restrict__ s_win,
                                             const float2* __restrict__ s_w512, int t, float neg_mu, int g) {
  const float* yb = ya + kV2Hop;
  const float s = g ? -1.f : 1.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int n = t + 16 * j;
    const float w = s_win[n];
    const int sb = j >= 10 ? 16 : 0;
    const cpx lo = cx(fmaf(ya[n], w, neg_mu), fmaf(yb[n + sb], w, neg_mu));
    if (j < 9) {  // n + 256 < 400 for every lane exactly when j <= 8
      const float w2 = s_win[n + 256];
      const int sa2 = j >= 4 ? 16 : 0;
      const cpx hi = cx(fmaf(ya[n + 256 + sa2], w2, neg_mu), fmaf(yb[n + 256 + 16], w2, neg_mu));
      v[j] = cx(fmaf(s, hi.x, lo.x), fmaf(s, hi.y, lo.y));
    } else {
      v[j] = lo;
    }
  }
  if (g) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 tw = s_w512[t + 16 * j];
      v[j] = cmulf(v[j], cx(tw.x, tw.y));
    }
  }
}

// pass P, interior tile, float input: each thread owns groups of 4 samples.  The 5 raw values a group needs (4 samples
// + predecessor) start at raw float 4q + SHIFT: two ALIGNED 16-byte loads and a compile-time pick.
template <int SHIFT, bool UNIT = false>
__device__ __forceinline__ void pass_p_f32(const float4* __restrict__ r4, float* __restrict__ ybuf, int tid, float scale,
                                           float pre_hi, float pre_lo) {
  int rem80 = tid % 80, gpad = 16 * (tid / 80);   // q mod 80, 16 * (q / 80) for q = tid + 256 k
#pragma unroll 1
  for (int q = tid; q < kV2Ylen / 4; q += kFastThreads) {
    const float4 A = r4[q], B = r4[q + 1];
    float x0, x1, x2, x3, x4;
    if (SHIFT == 0) { x0 = A.x; x1 = A.y; x2 = A.z; x3 = A.w; x4 = B.x; }
    else if (SHIFT == 1) { x0 = A.y; x1 = A.z; x2 = A.w; x3 = B.x; x4 = B.y; }
    else if (SHIFT == 2) { x0 = A.z; x1 = A.w; x2 = B.x; x3 = B.y; x4 = B.z; }
    else { x0 = A.w; x1 = B.x; x2 = B.y; x3 = B.z; x4 = B.w; }
    if (!UNIT) { x0 *= scale; x1 *= scale; x2 *= scale; x3 *= scale; x4 *= scale; }
    float4 y;
    y.x = fmaf(-pre_lo, x0, fmaf(-pre_hi, x0, x1));
    y.y = fmaf(-pre_lo, x1, fmaf(-pre_hi, x1, x2));
    y.z = fmaf(-pre_lo, x2, fmaf(-pre_hi, x2, x3));
    y.w = fmaf(-pre_lo, x3, fmaf(-pre_hi, x3, x4));
    *reinterpret_cast<float4*>(ybuf + 4 * q + gpad) = y;
    rem80 += 16; gpad += 48;                  // q += 256 = 3 * 80 + 16
    if (rem80 >= 80) { rem80 -= 80; gpad += 16; }
  }
}

template <bool I16, bool HW = false, typename SweepT = V3Sweep>
__global__ void __launch_bounds__(kFastThreads, 3) fbank512_v3_kernel(const __grid_constant__ V2Params P,
                                                                      const __grid_constant__ SweepT S) {
  using SM = typename std::conditional<HW, V5Smem, V3Smem>::type;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* rb = smem + SM::kRaw;   // the waveform tile lands in the upper part of Z
  float* ybuf = reinterpret_cast<float*>(smem + SM::kY);
  float* planes = reinterpret_cast<float*>(smem + SM::kPlanes);
  float2* Zs = reinterpret_cast<float2*>(smem + SM::kZ);
  float* stage = reinterpret_cast<float*>(smem + SM::kZ);
  float* s_win = reinterpret_cast<float*>(smem + SM::kWin);
  float2* s_w512 = reinterpret_cast<float2*>(smem + SM::kW512);
  float2* s_w256 = reinterpret_cast<float2*>(smem + SM::kW256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBar);
  int* s_work = reinterpret_cast<int*>(smem + SM::kBar) + 4;   // [2] claimed tile index per parity
  TileInfo* info = reinterpret_cast<TileInfo*>(smem + SM::kInfo);
  constexpr int ES = I16 ? 2 : 4;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kV2Flen; i += kFastThreads) s_win[i] = P.window[i];
  for (int i = tid; i < 256; i += kFastThreads) { if (!HW) s_w512[i] = P.w512[i]; s_w256[i] = P.w256t[i]; }
  const float2* w512 = HW ? P.w512 : s_w512;   // half-warp variant: the W512 twiddles are read through L1
  if (tid < kPlaneStride) planes[S.zero_row * kPlaneStride + tid] = 0.f;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // Thread 0 prepares every work item ONE ITERATION AHEAD, in STAGES spread over the iteration: each stage only issues
  // loads whose results are consumed after a later barrier, so the chain claim (atomicAdd) -> tile -> offsets / mean
  // -> TMA never stalls warp 0.  Tiles are claimed from a global counter: uneven tiles balance by themselves.
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
      tma_bulk_g2s(rb, (const unsigned char*)P.wave + src.ga_byte, src.bytes, &bars[slot]);
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

  const int t = lane & 15;
  const int pair = warp * 2 + (lane >> 4);
  float2* slot = Zs + pair * kSlotStride;
  const float* ya = ybuf + pair * 336;
  const float2* zp = Zs + (lane >> 1) * kSlotStride;
  const float sgn = (lane & 1) ? -1.f : 1.f;
  float* plane_dst = planes + (int)S.row0[warp] * kPlaneStride + lane;
  // phase C1 role of this thread: (frame group cg, filter cm) and the plane rows holding the filter's partial sums
  const int cg = tid / kV2Mels, cm = tid - cg * kV2Mels;
  const int crow = tid < 3 * kV2Mels ? P.combine[cm] : 0;   // row A | row B << 8

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
    // scalar patch-up of what the 16 B-granular bulk copy could not cover (end of the flat array)
    if (cur.cov_end < cur.end_elem) {
      for (int64_t e = cur.cov_end + tid; e < cur.end_elem; e += kFastThreads) {
        if (I16) reinterpret_cast<int16_t*>(rb)[e - cur.base_elem] = ((const int16_t*)P.wave)[e];
        else reinterpret_cast<float*>(rb)[e - cur.base_elem] = ((const float*)P.wave)[e];
      }
      __syncthreads();
    }

    // ---- pass P: [dither] + pre-emphasis, raw -> ybuf (y[0] = x[0] at the start of an utterance) ----
    // Lanes touch consecutive words; ybuf is PADDED by 16 floats per 320 samples so that the two frame pairs a warp
    // folds sit 16 banks apart.
    {
      const int64_t s0 = cur.s0;
      const int need = (nf - 1) * kV2Hop + kV2Flen;
      const int sh = cur.shift + 1;  // raw index of sample s0
      const bool interior = nf == kTileFrames && s0 > 0 && P.dither == 0.f && P.preemph_on;
      if (!I16 && interior) {
        const float4* r4 = reinterpret_cast<const float4*>(rb);
        switch (cur.shift) {   // 0..3 here (s0 > 0), tile uniform
          case 0: pass_p_f32<0>(r4, ybuf, tid, P.wave_scale, P.pre_hi, P.pre_lo); break;
          case 1: pass_p_f32<1>(r4, ybuf, tid, P.wave_scale, P.pre_hi, P.pre_lo); break;
          case 2: pass_p_f32<2>(r4, ybuf, tid, P.wave_scale, P.pre_hi, P.pre_lo); break;
          default: pass_p_f32<3>(r4, ybuf, tid, P.wave_scale, P.pre_hi, P.pre_lo); break;
        }
      } else if (interior) {
        // PCM16 input: lanes touch consecutive samples, predecessor through a shuffle (fixed trip count: convergent)
        int rem = tid, pad = 0;        // i mod 320, 16 * (i / 320)  (tid < 256 < 320)
#pragma unroll 1
        for (int k = 0; k < (kV2Ylen + kFastThreads - 1) / kFastThreads; ++k) {
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
        int rem = tid, pad = 0;
#pragma unroll 1
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

    // ---- the two 256-point groups: even bins (g = 0), odd bins (g = 1) ----
#pragma unroll 1
    for (int g = 0; g < 2; ++g) {
      cpx v[16];
      load_fold_v3(v, ya, s_win, w512, t, neg_mu, g);
      fft256_group(v, slot, s_w256, t);
      __syncthreads();
      // stage 3: offsets + frame-mean sum of the next tile's utterance (consumed after the sweep)
      if (g == 1 && tid == 0 && nx_w < P.n_tiles) load_offsets();
      if constexpr (HW) sweep_v5(S, g, warp, lane, Zs, planes);
      else sweep_v3(S, g, warp, zp, sgn, plane_dst);
      __syncthreads();
    }
    // the Z region has been read for the last time -> the next tile's waveform may land in its upper part
    if (tid == 0 && nx_w < P.n_tiles) issue_tile(buf ^ 1);

    // ---- phase C: combine the (<= 2) partial sums, log, store; thread = (frame group g, filter m) ----
    // Branch-free: a filter with fewer than 2 contributing warps reads the all-zero row for the missing ones.
    // The stores go straight to global memory: for one frame the 80 threads of a group write 320 contiguous bytes.
    float* part = stage;  // [3][2][80] per-group CMVN partial sums (lower part of the Z region, free after the sweeps)
    if (tid < 3 * kV2Mels) {
      const int g = cg, m = cm;
      const float* qa = planes + (crow & 0xff) * kPlaneStride + g;
      const float* qb = planes + (HW ? ((crow >> 8) & 0xff) : (crow >> 8)) * kPlaneStride + g;
      const float* qc = planes + (HW ? (crow >> 16) : 0) * kPlaneStride + g;   // half-warp variant: a third contributing range
      float* od = P.out + (cur.out_row + g) * (int64_t)kV2Mels + m;
      const int left = nf - g;
      // one code path for the three log kinds: ln(a == 0 ? eps : a), ln(a + c), a
      const bool use_log = P.log_kind != MAFE_LOG_NONE;
      const float add = P.log_kind == MAFE_LOG_LN_PLUS ? P.log_arg : 0.f;
      const float zero_sub = P.log_kind == MAFE_LOG_LN_EPS_IF_ZERO ? 2.220446049250313e-16f : 0.f;
      float s1 = 0.f, s2 = 0.f;
      // frames g, g + 3, ..., g + 30: fixed trip count, immediate offsets (f = 32 reads the pad column and is discarded)
#pragma unroll
      for (int i = 0; i < 11; ++i) {
        const float a = HW ? (qa[3 * i] + qb[3 * i]) + qc[3 * i] : qa[3 * i] + qb[3 * i];
        float x = a + add;
        x = x == 0.f ? zero_sub : x;
        // ln x = lg2 x * ln 2 with the raw MUFU (mel energies are never subnormal: 0 is replaced by DBL_EPSILON)
        float l;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
        const float o = use_log ? l * 0.69314718055994530942f : a;
        const bool in = 3 * i < left;
        if (in) od[3 * i * kV2Mels] = o;
        const float ov = in ? o : 0.f;
        s1 += ov;
        s2 = fmaf(ov, ov, s2);
      }
      if (P.utt_stats != nullptr) {
        part[(g * 2) * kV2Mels + m] = s1;
        part[(g * 2 + 1) * kV2Mels + m] = s2;
      }
    }
    __syncthreads();   // the next tile's geometry (thread 0, above) and the partial sums are visible
    // per-utterance CMVN statistics: one double atomic per (filter, moment)
    if (P.utt_stats != nullptr && tid < 2 * kV2Mels) {
      const int m = tid % kV2Mels, which = tid / kV2Mels;
      const double sum = (double)part[which * kV2Mels + m] + (double)part[(2 + which) * kV2Mels + m] +
                         (double)part[(4 + which) * kV2Mels + m];
      atomicAdd(&P.utt_stats[((size_t)utt * 2 + which) * kV2Mels + m], sum);
    }
  }
}

}  // namespace mafe
