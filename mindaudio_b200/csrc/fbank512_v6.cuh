// fbank512_v6.cuh -- version 6 of the headline kernel (conformer front-end: 400-sample frames, hop 160, 512-point
// FFT, 80 mel triangles; examples/conformer/dataset.py:117-168).  Included by fbank512.cu after fbank512_v3.cuh, whose
// tile geometry, TMA staging, work queue and pass P it shares.
//
// Why: v3 issues 707 warp-instructions per frame at 76 % issue-slot utilisation with the FP32 pipe only 37 % busy
// (profiles/r01_fbank512_v3_ncu_details.txt) -- it is bound by instruction ISSUE.  v6 is the same algorithm on a diet:
//   * every butterfly, twiddle and fold runs on PACKED complex registers (packed.cuh: FADD2 / FMUL2 / FFMA2 with free
//     swap / sign / broadcast operand modifiers): 107 instead of 220 FP instructions per 16-point stage + twiddles;
//   * the pair separation (A = Z[k] + conj Z[N-k], B = Z[k] - conj Z[N-k]) and |.|^2 move OUT of the mel sweep INTO the
//     FFT epilogue: the conjugate partner of a lane's bin lives at a FIXED (lane, register) of the same 16-lane group,
//     so it arrives by one shuffle pair; the lane then owns the POWER of its bins for both frames.  Both halves (even
//     bins, g = 0; odd bins, g = 1) are transformed back to back -- the even powers wait in 16 registers -- and the
//     lane writes rows  P[frame][bin]  (natural bin order) into the pair's own transpose slot: no second buffer;
//   * the mel projection is a DENSE per-filter dot product over P: lanes = the 32 frames, a warp owns whole filters,
//     a filter = 1..5 aligned 16-byte chunks of its frame's P row (LDS.128) against zero-padded weights from the
//     kernel-parameter bank, accumulated with FFMA2.  No running accumulators, no retire control, no partial sums to
//     combine (one warp computes a filter entirely): 183 chunks + 80 filter heads per frame instead of 257 sweep steps
//     x 22 instructions;
//   * 4 instead of 6 CTA barriers per tile.
#pragma once
#include "packed.cuh"

namespace mafe {

constexpr int kV6Row = 276;          // floats per P row (one frame); 276 % 32 == 20: LDS.128 by 8 consecutive rows is conflict free
constexpr int kV6Slot = kV6Row;      // float2 per pair slot = the pair's two P rows = its FFT transpose scratch (16 x 17 float2)
constexpr int kV6MaxChunks = 256;    // float4 weight chunks of all filters
constexpr int kV6MaxN4 = 8;          // chunks per filter
struct V6Sweep {          // kernel-parameter resident (constant bank 0)
  float4 w[kV6MaxChunks];                // warp by warp, filter by filter, nw[warp] chunks each: weights pre-scaled by 1/4 (the pair
                                         // separation leaves 2 X), zero outside the filter's support
  unsigned short start[kV2Mels];         // first float of the filter's first chunk in a P row (multiple of 4)
  unsigned short wbase[kFastWarps];      // first weight chunk of a warp
  unsigned char f0[kFastWarps + 8];      // warp w owns filters f0[w] .. f0[w + 1] - 1 (contiguous; cost balanced on the host)
  unsigned char nw[kFastWarps];          // chunks per filter in warp w (its widest filter; narrower ones are zero padded)
};

struct V6Smem {
  static constexpr size_t kY = 0;                                         // float[5632]: pre-emphasised tile (padded 16 per 320)
  static constexpr size_t kZ = kY + sizeof(float) * 5632;                 // float2[16][276]: FFT scratch / P rows; upper part = TMA landing zone
  static constexpr size_t kZBytes = sizeof(float2) * kPairs * kV6Slot;
  static constexpr size_t kRawInZ = kZBytes - kV2RawBytes;
  static constexpr size_t kRaw = kZ + kRawInZ;
  static constexpr size_t kPlanes = kZ + kZBytes;                         // float[80][33]
  static constexpr size_t kWin = kPlanes + ((sizeof(float) * kV2Mels * kPlaneStride + 15) & ~(size_t)15);
  static constexpr size_t kW512 = kWin + sizeof(float) * 400;
  static constexpr size_t kW256 = kW512 + sizeof(float2) * 256;
  static constexpr size_t kBar = kW256 + sizeof(float2) * 256;            // 2 mbarriers + 2 claimed indices
  static constexpr size_t kInfo = kBar + 32;                              // 2 x TileInfo + the record of the tile being normalised
  static constexpr size_t kNorm = kInfo + 3 * 64;                         // float[2][80]: mean, 1 / std of that tile's utterance
  static constexpr size_t kABar = kNorm + sizeof(float) * 2 * kV2Mels;    // mbarrier of its bulk copy
  static constexpr size_t kTotal = kABar + 16;
};
static_assert(3 * (V6Smem::kTotal + 1024) <= 228 * 1024, "3 CTAs per SM");
static_assert(V6Smem::kInfo % 16 == 0 && V6Smem::kNorm % 16 == 0 && V6Smem::kABar % 8 == 0, "smem alignment (cp.async / float4 / mbarrier)");
static_assert(V6Smem::kRaw % 128 == 0 && V6Smem::kRawInZ >= 6 * kV2Mels * 4, "landing zone vs the CMVN partial sums");
static_assert(V6Smem::kZ % 16 == 0 && V6Smem::kBar % 8 == 0 && V6Smem::kWin % 16 == 0 && V6Smem::kPlanes % 16 == 0, "smem alignment");
static_assert(15 * kRowStride + 15 < kV6Slot, "the 16 x 17 transpose scratch fits the slot");

// ---- per-lane constants in TENSOR MEMORY ----------------------------------------------------------------------------
// The window entries w[t + 16 j] and the twiddles W256^(t kj), W512^(t + 16 j) a lane needs depend on t = lane & 15 only
// and never change.  Read from shared memory they are 36 of the kernel's 164 shared-memory wavefronts per frame -- and
// ncu shows the shared-memory data pipe as THE bound of the packed kernel (l1tex__data_pipe_lsu_wavefronts 80 % of peak,
// issue slots 57 %, profiles/r02_fbank512_v6a_*).  Registers cannot hold them (87 words at an 80-register cap).  TMEM can:
// 512 columns x 128 lanes x 32 bit per SM, idle in a kernel without MMAs, with its OWN read path (tcgen05.ld, SASS LDTM):
// every thread owns one TMEM lane (32 (warp % 4) + lane) and keeps its constants in 96 columns of it.
constexpr int kTmCols = 128;          // allocation per CTA (power of two >= 32); 3 CTAs per SM use 384 of the 512 columns
constexpr int kTmWin = 0;             // columns  0..31: w[t + 16 c], c = 0..24
constexpr int kTmW256 = 32;           // columns 32..63: W256^(t kj) = (cos, sin), kj = 1..15, at 32 + 2 (kj - 1)
constexpr int kTmW512 = 64;           // columns 64..95: W512^(t + 16 j) = (cos, sin), j = 0..15
__device__ __forceinline__ void tm_ld8(uint32_t taddr, float (&v)[8]) {   // issue only: tm_wait8 before the first use
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tm_ld1(uint32_t taddr, float& v) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=f"(v) : "r"(taddr));
}
// wait for the outstanding TMEM loads; the loaded registers pass through the statement so that no use is scheduled above it
__device__ __forceinline__ void tm_wait8(float (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
}
__device__ __forceinline__ void tm_wait1(float& v) { asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(v)); }
// Warps w and w + 4 share a TMEM lane quarter: the constants are common, the parking space is per warp (h = warp >> 2).
constexpr int kTmPark = 96;           // columns 96 + 16 h ..: the even-bin powers of a lane wait here while the odd half is transformed
constexpr int kTmNyq = 26;            // columns 26 + 2 h, 27 + 2 h (unused tail of the window block): powers of bin 256
__device__ __forceinline__ void tm_st16(uint32_t taddr, const float (&a)[8], const float (&b)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]), "f"(b[0]), "f"(b[1]), "f"(b[2]),
      "f"(b[3]), "f"(b[4]), "f"(b[5]), "f"(b[6]), "f"(b[7])
      : "memory");
}
__device__ __forceinline__ void tm_st2(uint32_t taddr, float a, float b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, float (&a)[8], float (&b)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]), "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]), "=f"(b[0]), "=f"(b[1]),
        "=f"(b[2]), "=f"(b[3]), "=f"(b[4]), "=f"(b[5]), "=f"(b[6]), "=f"(b[7])
      : "r"(taddr));
}
__device__ __forceinline__ void tm_ld2(uint32_t taddr, float& a, float& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=f"(a), "=f"(b) : "r"(taddr));
}
__device__ __forceinline__ void tm_wait2(float& a, float& b) { asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a), "+f"(b)); }
__device__ __forceinline__ void tm_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v[0]),
               "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// frame pair -> registers for group g: window, mean removal, radix-2 fold (g = 0: lo + hi -> even bins; g = 1:
// (lo - hi) W512^n -> odd bins).  `y` points at the pair's first sample in the padded tile (+16 floats per 320 samples):
// lane t needs samples t + 16 J, J = 0..34 (frame a: J < 25, frame b = a + 160: J >= 10); the pad of J is 16 (J / 20).
__device__ __forceinline__ float v6_y(const float* y, int J) { return y[16 * J + 16 * (J / 20)]; }
__device__ __forceinline__ void load_fold_v6(c2 (&v)[16], const float* __restrict__ yt, const float* __restrict__ wt,
                                             const float2* __restrict__ w512t, float neg_mu, int g) {
  const c2 nm = bc(neg_mu);
  const c2 s = bc(g ? -1.f : 1.f);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const c2 lo = fma2(pk(v6_y(yt, j), v6_y(yt, j + 10)), bc(wt[16 * j]), nm);
    if (j < 9) {  // n + 256 < 400 for every lane exactly when j <= 8
      const c2 hi = fma2(pk(v6_y(yt, j + 16), v6_y(yt, j + 26)), bc(wt[16 * j + 256]), nm);
      v[j] = fma2(hi, s, lo);
    } else {
      v[j] = lo;
    }
  }
  if (g) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float2 tw = w512t[16 * j];
      v[j] = cmul(v[j], tw.x, tw.y);
    }
  }
}

// the same with the window and the W512 twiddles from the lane's TMEM columns
__device__ __forceinline__ void load_fold_v6_tm(c2 (&v)[16], const float* __restrict__ yt, uint32_t tb, float neg_mu, int g) {
  const c2 nm = bc(neg_mu);
  const c2 s = bc(g ? -1.f : 1.f);
  float wl[8], wh[8], w24;
  tm_ld8(tb + kTmWin, wl);        // w[t + 16 j], j = 0..7
  tm_ld8(tb + kTmWin + 16, wh);   // w[t + 16 j + 256], j = 0..7
  tm_wait8(wl);
  tm_wait8(wh);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const c2 lo = fma2(pk(v6_y(yt, j), v6_y(yt, j + 10)), bc(wl[j]), nm);
    const c2 hi = fma2(pk(v6_y(yt, j + 16), v6_y(yt, j + 26)), bc(wh[j]), nm);
    v[j] = fma2(hi, s, lo);
  }
  tm_ld8(tb + kTmWin + 8, wl);    // j = 8..15
  tm_ld1(tb + kTmWin + 24, w24);  // w[t + 16 * 8 + 256]
  tm_wait8(wl);
  tm_wait1(w24);
  {
    const c2 lo = fma2(pk(v6_y(yt, 8), v6_y(yt, 18)), bc(wl[0]), nm);
    const c2 hi = fma2(pk(v6_y(yt, 24), v6_y(yt, 34)), bc(w24), nm);
    v[8] = fma2(hi, s, lo);
  }
#pragma unroll
  for (int j = 9; j < 16; ++j) v[j] = fma2(pk(v6_y(yt, j), v6_y(yt, j + 10)), bc(wl[j - 8]), nm);
  if (g) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float tw[8];
      tm_ld8(tb + kTmW512 + 8 * c, tw);
      tm_wait8(tw);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[4 * c + i] = cmul(v[4 * c + i], tw[2 * i], tw[2 * i + 1]);
    }
  }
}

// 256-point transform of v by the 16-lane group (16 points per lane); output kk = t + 16 kt of the group's sub-transform
// ends in u[fft16_pos(kt)] of lane t.
__device__ __forceinline__ void fft256_v6(c2 (&v)[16], c2 (&u)[16], float2* slot, const float2* s_w256, int t) {
  fft16p(v);
#pragma unroll
  for (int kj = 0; kj < 16; ++kj) {
    c2 x = v[fft16_pos(kj)];
    if (kj > 0) {
      const float2 tw = s_w256[kj * 16 + t];
      x = cmul(x, tw.x, tw.y);
    }
    sts_c2(slot + kj * kRowStride + t, x);
  }
  __syncwarp();
#pragma unroll
  for (int tt = 0; tt < 16; ++tt) u[tt] = lds_c2(slot + t * kRowStride + tt);
  __syncwarp();
  fft16p(u);
}

// the same with the W256 twiddles from the lane's TMEM columns
__device__ __forceinline__ void fft256_v6_tm(c2 (&v)[16], c2 (&u)[16], float2* slot, uint32_t tb, int t) {
  fft16p(v);
  sts_c2(slot + t, v[fft16_pos(0)]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float tw[8];
    tm_ld8(tb + kTmW256 + 8 * c, tw);   // kj = 4 c + 1 .. 4 c + 4
    tm_wait8(tw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kj = 4 * c + 1 + i;
      if (kj < 16) sts_c2(slot + kj * kRowStride + t, cmul(v[fft16_pos(kj)], tw[2 * i], tw[2 * i + 1]));
    }
  }
  __syncwarp();
#pragma unroll
  for (int tt = 0; tt < 16; ++tt) u[tt] = lds_c2(slot + t * kRowStride + tt);
  __syncwarp();
  fft16p(u);
}

// Mel projection of one warp: lanes = frames, filters m0 .. m1 - 1, N aligned 16-byte chunks of the lane's P row per filter
// against the filter's (zero padded) weights; two filters per iteration so that two FFMA2 chains are in flight.
template <int N>
__device__ __forceinline__ float v6_dot(const float* __restrict__ row, const float4* __restrict__ w) {
  c2 acc = pk(0.f, 0.f);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float4 p = *reinterpret_cast<const float4*>(row + 4 * j);
    const float4 ww = w[j];
    acc = fma2(pk(p.x, p.y), pk(ww.x, ww.y), acc);
    acc = fma2(pk(p.z, p.w), pk(ww.z, ww.w), acc);
  }
  return re(acc) + im(acc);
}
template <int N>
__device__ __forceinline__ void v6_mel(const V6Sweep& S, const float* __restrict__ my_row, float* __restrict__ plane_col,
                                       int m0, int m1, int wbase) {
  const float4* w = S.w + wbase;
  int m = m0;
#pragma unroll 1
  for (; m + 1 < m1; m += 2, w += 2 * N) {
    const float a = v6_dot<N>(my_row + S.start[m], w);
    const float b = v6_dot<N>(my_row + S.start[m + 1], w + N);
    plane_col[m * kPlaneStride] = a;
    plane_col[(m + 1) * kPlaneStride] = b;
  }
  if (m < m1) plane_col[m * kPlaneStride] = v6_dot<N>(my_row + S.start[m], w);
}

// FUSE: the utterance CMVN is applied INSIDE this kernel.  The tile queue runs lag items past the last tile; the CTA that
// claims item w also owns the normalisation of tile w - lag: by then that tile's utterance is (almost always) complete --
// every finished tile bumps utt_done[utt] with release semantics, the normalising CTA acquires it (and spins in the rare
// case it is early; lag >= the longest utterance in tiles makes the wait deadlock free: everything it waits for has been
// claimed by a resident CTA).  The tile's 10 KB of raw log-mel come back from L2 (written ~lag tiles = 20 MB ago) by one
// bulk copy into the then idle waveform buffer, are normalised with the utterance's mean / inverse deviation and stored
// for good: the separate apply kernel -- a second read and write of the whole feature matrix through HBM -- is gone.
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_relaxed_inc(int* p) { asm volatile("red.relaxed.gpu.global.add.s32 [%0], 1;" ::"l"(p) : "memory"); }
__device__ __forceinline__ double ld_cg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}

// FS: the frame-mean sums inside the persistent kernel.  The reference subtracts ONE scalar per utterance -- the mean of the
// whole windowed frame matrix (dataset.py:165) -- so every sample must be seen before any frame of the utterance can be
// transformed; round 1 / early round 2 ran a streaming pre-pass kernel over the whole batch (0.96 ms, 5.5 GB: the waveform
// came from HBM twice).  Now the queue runs lag_s items ahead of the transforms: the CTA that claims item w first adds
//     sum_s x[s] c'(s),   c'(s) = c(s) - a c(s + 1)   (c = window summed over the frames covering s, a = pre-emphasis:
//                                                        sum_s y[s] c(s) with y[s] = x[s] - a x[s - 1] telescopes)
// over the samples tile w owns to utt_fsum[utt] (release + counter, like the CMVN duty), then transforms tile w - lag_s,
// whose utterance is complete by then; the transform's bulk copy finds the samples in L2 (read ~lag_s tiles = 30 MB ago).
// Inside an utterance c' has period 160: 240 threads take 16-byte groups 160 samples apart, so a thread's coefficients are
// the same for all its groups of a tile (ONE 16-byte shared-memory load per thread and tile, from the copy of the table
// shifted by r mod 4); the first 320 samples and the tail of an utterance use the general rule.
#ifndef MAFE_FS_POLICY
#define MAFE_FS_POLICY 1
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

// This thread's part of sum_s x[s] c'(s) over the samples tile `sr` owns (FS above): issue() = every global load,
// finish() = coefficients + FMAs.
template <bool I16>
struct V6Sum {
  static constexpr int V = I16 ? 8 : 4;                       // samples per 16-byte group
  static constexpr int GP = kV2Hop / V;                       // groups per period (40 / 20)
  static constexpr int NT = (kFastThreads / GP) * GP;         // threads taking part (240)
  static constexpr int PSTEP = NT / GP;                       // periods per sweep of the CTA (6 / 12)
  static constexpr int KMAX = (kTileFrames + PSTEP - 1) / PSTEP;
  int4 raw[KMAX];
  int h0, h1;        // scalar head / tail samples around the 16 B-aligned groups (raw bits), coefficient indices
  int i0, i1, r;
  float acc;

  static __device__ __forceinline__ int ld_raw(const void* wave, int64_t g) {
    if (I16) return (int)__ldg((const int16_t*)wave + g);
    return __float_as_int(__ldg((const float*)wave + g));
  }
  static __device__ __forceinline__ float to_f(int v) { return I16 ? (float)v : __int_as_float(v); }

  __device__ __forceinline__ void issue(const V2Params& P, const SumRec& sr, const TileInfo* __restrict__ grec, int tid) {
    acc = 0.f;
    h0 = 0; h1 = 0; i0 = 0; i1 = 0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) raw[k] = make_int4(0, 0, 0, 0);
    const int64_t ga = sr.ga;
    const int nvec = sr.nvec, r0 = sr.r0;
    if (sr.edge) {   // first / last tile of an utterance: the general rule outside the periodic region (2 of ~33 tiles)
      const TileInfo si = *grec;
      const int T = si.T;
      const int need = (si.nf - 1) * kV2Hop + kV2Flen;
      const int64_t off = si.end_elem - si.s0 - need;      // the utterance's first sample in the flat array
      const int64_t s_lo = si.s0;
      const int64_t framed_end = (int64_t)(T - 1) * kV2Hop + kV2Flen;
      const bool last = s_lo + (int64_t)kTileFrames * kV2Hop >= (int64_t)T * kV2Hop;
      const int64_t s_hi = last ? framed_end : s_lo + (int64_t)kTileFrames * kV2Hop;
      const int64_t f_lo = ga - off - sr.n_head, f_hi = ga - off + (int64_t)nvec * V + sr.n_tail;
      auto cgen = [&](int64_t q) -> float {
        if (q >= framed_end) return 0.f;
        const int64_t th = q / kV2Hop;
        const int t_hi = (int)(th < T - 1 ? th : T - 1);
        float c = 0.f;
        for (int tt = t_hi; tt >= 0; --tt) {
          const int64_t n = q - (int64_t)tt * kV2Hop;
          if (n >= kV2Flen) break;
          c += __ldg(&P.window[n]);
        }
        return c;
      };
      auto edge = [&](int64_t q) {
        const float x = to_f(ld_raw(P.wave, off + q));
        const float c1 = cgen(q + 1);
        acc = fmaf(x, fmaf(-P.pre_lo, c1, fmaf(-P.pre_hi, c1, cgen(q))), acc);
      };
      for (int64_t q = s_lo + tid; q < f_lo; q += kFastThreads) edge(q);
      for (int64_t q = f_hi + tid; q < s_hi; q += kFastThreads) edge(q);
    }
    if (tid < sr.n_head) {
      h0 = ld_raw(P.wave, ga - sr.n_head + tid);
      i0 = r0 - sr.n_head + tid;
      if (i0 < 0) i0 += kV2Hop;
    }
    if (tid < sr.n_tail) {
      h1 = ld_raw(P.wave, ga + (int64_t)nvec * V + tid);
      i1 = (r0 + nvec * V + tid) % kV2Hop;
    }
    const int rho = tid % GP, p0 = tid / GP;
    r = r0 + V * rho;
    if (r >= kV2Hop) r -= kV2Hop;
    if (tid < NT) {
      const int4* src = reinterpret_cast<const int4*>((const unsigned char*)P.wave + ga * (I16 ? 2 : 4));
#if MAFE_FS_POLICY
      const uint64_t pol = l2_policy_evict_last();   // the transform's bulk copy wants these lines again ~lag_s tiles from now
#endif
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int q = rho + GP * (p0 + PSTEP * k);
#if MAFE_FS_POLICY
        if (q < nvec) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0, %1, %2, %3}, [%4], %5;"
                                   : "=r"(raw[k].x), "=r"(raw[k].y), "=r"(raw[k].z), "=r"(raw[k].w) : "l"(src + q), "l"(pol));
#else
        if (q < nvec) asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                                   : "=r"(raw[k].x), "=r"(raw[k].y), "=r"(raw[k].z), "=r"(raw[k].w) : "l"(src + q));
#endif
      }
    }
  }

  __device__ __forceinline__ float finish(const float* __restrict__ s_cp4) {
    const float* crow = s_cp4 + (r & 3) * kCwRow + (r - (r & 3));   // crow[i] = c'((r + i) mod 160), 16 B aligned
    float c[V];
#pragma unroll
    for (int i = 0; i < V; i += 4) {
      const float4 c4 = *reinterpret_cast<const float4*>(crow + i);
      c[i] = c4.x; c[i + 1] = c4.y; c[i + 2] = c4.z; c[i + 3] = c4.w;
    }
    float a = fmaf(to_f(h0), s_cp4[i0], fmaf(to_f(h1), s_cp4[i1], acc));   // absent samples are zero bits
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {   // absent groups are zero
      const int w4[4] = {raw[k].x, raw[k].y, raw[k].z, raw[k].w};
      if (I16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          a = fmaf((float)(int16_t)(w4[i] & 0xffff), c[(2 * i) % V], a);
          a = fmaf((float)(int16_t)(w4[i] >> 16), c[(2 * i + 1) % V], a);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) a = fmaf(__int_as_float(w4[i]), c[i % V], a);
      }
    }
    return a;
  }
};

template <bool I16, bool TM, bool FUSE, bool FS = false>
#ifndef MAFE_V6_CTAS
#define MAFE_V6_CTAS 3
#endif
__global__ void __launch_bounds__(kFastThreads, MAFE_V6_CTAS) fbank512_v6_kernel(const __grid_constant__ V2Params P,
                                                                      const __grid_constant__ V6Sweep S) {
  using SM = V6Smem;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* rb = smem + SM::kRaw;   // the waveform tile lands in the upper part of Z
  float* ybuf = reinterpret_cast<float*>(smem + SM::kY);
  float* planes = reinterpret_cast<float*>(smem + SM::kPlanes);
  float2* Zs = reinterpret_cast<float2*>(smem + SM::kZ);
  float* prows = reinterpret_cast<float*>(smem + SM::kZ);
  float* stage = reinterpret_cast<float*>(smem + SM::kZ);
  float* s_win = reinterpret_cast<float*>(smem + SM::kWin);
  float2* s_w512 = reinterpret_cast<float2*>(smem + SM::kW512);
  float2* s_w256 = reinterpret_cast<float2*>(smem + SM::kW256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kBar);
  int* s_work = reinterpret_cast<int*>(smem + SM::kBar) + 4;   // [2] claimed tile index per parity
  TileInfo* info = reinterpret_cast<TileInfo*>(smem + SM::kInfo);
  TileInfo* ainfo = info + 2;                                       // FUSE: record of the tile being normalised
  float* s_norm = reinterpret_cast<float*>(smem + SM::kNorm);
  uint64_t* abar = reinterpret_cast<uint64_t*>(smem + SM::kABar);
  constexpr int ES = I16 ? 2 : 4;
  static_assert(!FS || (TM && FUSE), "the frame-sum duty uses the shared memory the TMEM variant leaves idle");
  static_assert(2 * sizeof(SumRec) + 8 * sizeof(float) <= sizeof(float) * 400 && sizeof(float) * 4 * kCwRow <= 2 * sizeof(float2) * 256, "FS tables");
  SumRec* sinfo = reinterpret_cast<SumRec*>(smem + SM::kWin);              // FS: [2] records of the tiles being summed
  float* s_ws = reinterpret_cast<float*>(smem + SM::kWin + 2 * sizeof(SumRec));     // FS: [8] warp partial sums
  float* s_cp4 = reinterpret_cast<float*>(smem + SM::kW512);               // FS: c' table, 4 shifted copies

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t* s_tm = reinterpret_cast<uint32_t*>(smem + SM::kBar) + 6;   // TMEM base address of this CTA
  if (FS) {
    for (int i = tid; i < 4 * kCwRow; i += kFastThreads) s_cp4[i] = P.cover4[i];
  }
  if (!TM) {
    for (int i = tid; i < kV2Flen; i += kFastThreads) s_win[i] = P.window[i];
    for (int i = tid; i < 256; i += kFastThreads) { s_w512[i] = P.w512[i]; s_w256[i] = P.w256t[i]; }
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(abar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (TM && warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tm)), "n"(kTmCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (TM) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  uint32_t tb = 0;
  if (TM) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    tb = *s_tm + ((uint32_t)(32 * (warp & 3)) << 16);   // this warp's lane quarter
    if (warp < 4) {   // warps w and w + 4 share a quarter and need the same constants (they depend on lane & 15 only)
      const int tt = lane & 15;
      float c8[8];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 8; ++i) c8[i] = 8 * c + i < 25 ? P.window[tt + 16 * (8 * c + i)] : 0.f;
        tm_st8(tb + kTmWin + 8 * c, c8);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kj = 4 * c + 1 + i;
          const float2 w = kj < 16 ? P.w256t[kj * 16 + tt] : make_float2(0.f, 0.f);
          c8[2 * i] = w.x; c8[2 * i + 1] = w.y;
        }
        tm_st8(tb + kTmW256 + 8 * c, c8);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 w = P.w512[tt + 16 * (4 * c + i)];
          c8[2 * i] = w.x; c8[2 * i + 1] = w.y;
        }
        tm_st8(tb + kTmW512 + 8 * c, c8);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  const uint32_t tpark = tb + kTmPark + 16 * (warp >> 2), tnyq = tb + kTmNyq + 2 * (warp >> 2);

  // Thread 0 prepares every work item ONE ITERATION AHEAD, in stages spread over the iteration: claim (atomicAdd) at the
  // top, publish after pass P, copy the tile's 64-byte record (tile_prepare_kernel) into the other info slot with
  // cp.async after the FFT phase, bulk copy of the waveform tile after the sweep.  Tiles are claimed from a global
  // counter: uneven tiles balance by themselves.
  const TileInfo* recs = reinterpret_cast<const TileInfo*>(P.tile_recs);
  int nx_w = P.n_tiles;
  const int lag_s = FS ? P.lag_s : 0;
  auto fetch_to = [&](TileInfo* dstp, int tile_i) {   // asynchronous 64-byte copy global -> shared
    const uint32_t dst = smem_u32(dstp);
    const unsigned char* srcp = reinterpret_cast<const unsigned char*>(recs + tile_i);
#pragma unroll
    for (int q = 0; q < 4; ++q)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * q), "l"(srcp + 16 * q) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto fetch_rec = [&](int slot_i) { fetch_to(&info[slot_i], nx_w - lag_s); };
  const SumRec* srecs = reinterpret_cast<const SumRec*>(P.sum_recs);
  auto fetch_srec = [&](int slot_i) {   // 32-byte sum record of tile nx_w; the tile's samples start their way into L2
    const uint32_t dst = smem_u32(&sinfo[slot_i]);
    const unsigned char* srcp = reinterpret_cast<const unsigned char*>(srecs + nx_w);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(srcp) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(srcp + 16) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  auto next_main = [&]() { return nx_w - lag_s >= 0 && nx_w - lag_s < P.n_tiles; };   // the next item transforms a tile
  auto issue_tile = [&](int slot_i) {  // the record has landed: bulk copy of the tile's bytes
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    const uint32_t bytes = info[slot_i].bytes;
    const int64_t ga_byte = info[slot_i].base_elem * ES;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (bytes) {
      mbar_expect_tx(&bars[slot_i], bytes);
      if (FS && MAFE_FS_POLICY)   // the last use of these lines
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                         smem_u32(rb)), "l"((const unsigned char*)P.wave + ga_byte), "r"(bytes), "r"(smem_u32(&bars[slot_i])), "l"(l2_policy_evict_first()) : "memory");
      else
        tma_bulk_g2s(rb, (const unsigned char*)P.wave + ga_byte, bytes, &bars[slot_i]);
    } else {
      mbar_arrive(&bars[slot_i]);
    }
  };

  // queue: item w sums tile w (FS), transforms tile w - lag_s, normalises tile w - lag_s - lag (FUSE)
  const int n_items = P.n_tiles + lag_s + (FUSE ? P.lag : 0);
  if (tid == 0) {   // prologue: the first item, all stages back to back (with FS it is a sum-only item: lag_s > grid)
    nx_w = atomicAdd(P.queue_head, 1);
    s_work[0] = nx_w;
    if (FS && nx_w < P.n_tiles) fetch_srec(0);
    if (next_main()) { fetch_rec(0); issue_tile(0); }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  // service thread (the last thread of the CTA: its warp has only half a share of phase C): utterance of the tile this CTA
  // finished last and not yet published; state of the current duty
  const bool svc = tid == kFastThreads - 1;
  int prev_utt = -1, a_flag = 0, a_need = 0;
  auto publish = [&]() {   // the fence orders everything this CTA wrote so far before the counter bump (cumulative over barriers)
    __threadfence();
    red_relaxed_inc(P.utt_done + prev_utt);
    prev_utt = -1;
  };
  uint32_t a_cnt = 0;                          // normalisation duties done by this CTA (parity of abar)
  int m_utt = 0, m_need = 0, m_flag = 0;       // FS (thread 0): utterance of the next tile to transform, its tile count, tiles summed

  const int t = lane & 15;
  const int pair = warp * 2 + (lane >> 4);
  float2* slot = Zs + pair * kV6Slot;
  const float* yt = ybuf + pair * 336 + t;
  const float* wt = s_win + t;
  const float2* w512t = s_w512 + t;
  float* row_a = prows + (2 * pair) * kV6Row + 2 * t;   // this lane's first bin pair (2t, 2t + 1) of frame a; frame b one row on
  const float* my_row = prows + lane * kV6Row;          // sweep: lane = frame
  // partner lanes of the pair separation: even half (g = 0): bin kk <-> 256 - kk -> lane (16 - t) & 15; odd half: 255 - kk -> 15 - t
  const int src_even = (lane & 16) | ((16 - t) & 15), src_odd = (lane & 16) | (15 - t);
  // phase C role of this thread: (frame group cg, filter cm)
  const int cg = tid / kV2Mels, cm = tid - cg * kV2Mels;
  const int f_begin = S.f0[warp], f_end = S.f0[warp + 1], nw = S.nw[warp], wbase = S.wbase[warp];

  // iteration it uses info / barrier slot it & 1; the slot's mbarrier completes its ((it - it0) >> 1)-th phase
  uint32_t it0 = FS ? 0xffffffffu : 0u;
  for (uint32_t it = 0;; ++it) {
    const int buf = it & 1;
    const int w = s_work[buf];
    if (w >= n_items) break;
    const int u = w - lag_s;                           // the tile this item transforms
    const bool main_tile = u >= 0 && u < P.n_tiles;
    const bool duty = FUSE && u >= P.lag;              // normalise tile u - lag
    const bool sum_duty = FS && w < P.n_tiles;         // sum the samples of tile w
    if (tid == 0) nx_w = atomicAdd(P.queue_head, 1);   // stage 1 (issue): claim the next item
    if (FS && sum_duty) {
      V6Sum<I16> fsum;
      fsum.issue(P, sinfo[buf], recs + w, tid);
      float fs = fsum.finish(s_cp4);
      for (int o = 16; o > 0; o >>= 1) fs += __shfl_xor_sync(0xffffffffu, fs, o);
      if (lane == 0) s_ws[warp] = fs;                  // read by the service thread after the third barrier
    }
    if (FUSE && svc) {
      if (duty) {                          // the duty tile's record -> ainfo (asynchronous)
        const uint32_t dst = smem_u32(ainfo);
        const unsigned char* srcp = reinterpret_cast<const unsigned char*>(recs + (u - P.lag));
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * q), "l"(srcp + 16 * q) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
    }
    const TileInfo cur = info[buf];
    const uint32_t utt = (uint32_t)cur.utt;
    const int nf = cur.nf;
    float neg_mu = 0.f;
    if (main_tile) {
    // wait for this tile's bytes
    if (FS && it0 == 0xffffffffu) it0 = it;   // FS: the first items of a CTA only sum -- the slots' phases count from its first transform
    mbar_wait(&bars[buf], ((it - it0) >> 1) & 1);
    // scalar patch-up of what the 16 B-granular bulk copy could not cover (end of the flat array)
    if (cur.cov_end < cur.end_elem) {
      for (int64_t e = cur.cov_end + tid; e < cur.end_elem; e += kFastThreads) {
        if (I16) reinterpret_cast<int16_t*>(rb)[e - cur.base_elem] = ((const int16_t*)P.wave)[e];
        else reinterpret_cast<float*>(rb)[e - cur.base_elem] = ((const float*)P.wave)[e];
      }
      __syncthreads();
    }

    // ---- pass P: [dither] + pre-emphasis, raw -> ybuf (y[0] = x[0] at the start of an utterance) ----
    {
      const int64_t s0 = cur.s0;
      const int need = (nf - 1) * kV2Hop + kV2Flen;
      const int sh = cur.shift + 1;  // raw index of sample s0
      const bool interior = nf == kTileFrames && s0 > 0 && P.dither == 0.f && P.preemph_on;
      if (!I16 && interior) {
        const float4* r4 = reinterpret_cast<const float4*>(rb);
        if (P.wave_scale == 1.0f) {
          switch (cur.shift) {   // 0..3 here (s0 > 0), tile uniform
            case 0: pass_p_f32<0, true>(r4, ybuf, tid, 1.0f, P.pre_hi, P.pre_lo); break;
            case 1: pass_p_f32<1, true>(r4, ybuf, tid, 1.0f, P.pre_hi, P.pre_lo); break;
            case 2: pass_p_f32<2, true>(r4, ybuf, tid, 1.0f, P.pre_hi, P.pre_lo); break;
            default: pass_p_f32<3, true>(r4, ybuf, tid, 1.0f, P.pre_hi, P.pre_lo); break;
          }
        } else {
          switch (cur.shift) {
            case 0: pass_p_f32<0>(r4, ybuf, tid, P.wave_scale, P.pre_hi, P.pre_lo); break;
            case 1: pass_p_f32<1>(r4, ybuf, tid, P.wave_scale, P.pre_hi, P.pre_lo); break;
            case 2: pass_p_f32<2>(r4, ybuf, tid, P.wave_scale, P.pre_hi, P.pre_lo); break;
            default: pass_p_f32<3>(r4, ybuf, tid, P.wave_scale, P.pre_hi, P.pre_lo); break;
          }
        }
      } else if (interior) {
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
    neg_mu = cur.neg_mu;
    }   // main_tile (pass P)
    __syncthreads();
    if (tid == 0) {
      s_work[buf ^ 1] = nx_w;   // stage 2: the claim has arrived during pass P -> publish it
      if (FS) {                 // FS: both records of the next item travel during the FFT phase (its -mu needs two more round trips)
        if (nx_w < P.n_tiles) fetch_srec(buf ^ 1);
        if (next_main()) fetch_rec(buf ^ 1);
      }
    }
    if (FUSE && svc && duty) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      a_need = (ainfo->T + kTileFrames - 1) / kTileFrames;
      // Relaxed (L2) load, consumed after the FFT phase.  The data reads that depend on it -- the bulk copy and the .cg
      // loads of the statistics -- are issued after the branch on its value and go to L2, where the writer's fence put
      // the data before the counter: no acquire fence on the fast path (the slow path spins with ld.acquire).
      a_flag = ld_relaxed(P.utt_done + ainfo->utt);
    }
    if (main_tile) {

    // ---- FFT phase: both 256-point halves of the pair, powers of the lane's bins, P rows ----
    {
      float ea[8], eb[8];          // even-bin powers (frames a, b) of kk = t + 16 kt, parked while the odd half runs
      float nyq_a = 0.f, nyq_b = 0.f;
#pragma unroll 1
      for (int g = 0; g < 2; ++g) {
        c2 v[16], u[16];
        if (TM) {
          load_fold_v6_tm(v, yt, tb, neg_mu, g);
          fft256_v6_tm(v, u, slot, tb, t);
        } else {
          load_fold_v6(v, yt, wt, w512t, neg_mu, g);
          fft256_v6(v, u, slot, s_w256, t);
        }
        const int src = g ? src_odd : src_even;
        const bool self = g == 0 && t == 0;   // lane 0 of the even half is its own partner, one register further
        float pa[8], pb[8];
#pragma unroll
        for (int kt = 0; kt < 8; ++kt) {
          const c2 zk = u[fft16_pos(kt)];
          const c2 send = u[fft16_pos(15 - kt)];
          float pr = __shfl_sync(0xffffffffu, re(send), src);
          float pi = __shfl_sync(0xffffffffu, im(send), src);
          if (self) { pr = re(u[fft16_pos((16 - kt) & 15)]); pi = im(u[fft16_pos((16 - kt) & 15)]); }
          const c2 zn = pk(pr, -pi);                      // conj Z[N - k]
          const c2 sa = add2(zk, zn), sb = sub2(zk, zn);  // 2 A[k], 2 i B[k]
          const c2 qa = mul2(sa, sa), qb = mul2(sb, sb);
          pa[kt] = re(qa) + im(qa);
          pb[kt] = re(qb) + im(qb);
        }
        if (g == 0) {
          // bin 256 (kk = 128, lane 0): its own partner: A = 2 re, B = 2 im
          const c2 zq = u[fft16_pos(8)];
          nyq_a = 4.f * re(zq) * re(zq);
          nyq_b = 4.f * im(zq) * im(zq);
          if (TM) {   // park the even-bin powers in the lane's TMEM columns: 18 registers less through the odd half
            tm_st16(tpark, pa, pb);
            tm_st2(tnyq, nyq_a, nyq_b);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          } else {
#pragma unroll
            for (int kt = 0; kt < 8; ++kt) { ea[kt] = pa[kt]; eb[kt] = pb[kt]; }
          }
        } else {
          if (TM) {
            tm_ld16(tpark, ea, eb);
            tm_ld2(tnyq, nyq_a, nyq_b);
            tm_wait8(ea);
            tm_wait8(eb);
            tm_wait2(nyq_a, nyq_b);
          }
          __syncwarp();   // every lane of the group has read the scratch: the slot becomes the pair's two P rows
#pragma unroll
          for (int kt = 0; kt < 8; ++kt) {
            *reinterpret_cast<float2*>(row_a + 32 * kt) = make_float2(ea[kt], pa[kt]);
            *reinterpret_cast<float2*>(row_a + kV6Row + 32 * kt) = make_float2(eb[kt], pb[kt]);
          }
          if (t == 0) { row_a[256] = nyq_a; row_a[kV6Row + 256] = nyq_b; }
        }
      }
    }
    }   // main_tile (FFT phase)
    if (FUSE && svc && duty && a_flag != a_need) {
      // early (rare): first publish what this CTA still holds back -- the utterance may be waiting for exactly that -- then
      // spin; every tile waited for has been claimed by a resident CTA (lag >= the longest utterance), which publishes it
      // within one iteration or before it waits itself
      if (prev_utt >= 0) publish();
      do { a_flag = ld_acquire(P.utt_done + ainfo->utt); } while (a_flag != a_need);
    }
    __syncthreads();
    // stage 3: the next tile's record travels to the other info slot during the mel projection
    if (!FS && tid == 0 && nx_w < P.n_tiles) fetch_rec(buf ^ 1);
    if (FS && tid == 0) {   // the records have landed: how many tiles of the next tile's utterance have been summed? (used after the mel stage)
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (nx_w < P.n_tiles) {   // the next item's samples: on their way into L2 now, read by the whole CTA one iteration later
        const SumRec& nr = sinfo[buf ^ 1];
        const uint32_t pbytes = (uint32_t)nr.nvec * 16u;
#if MAFE_FS_POLICY
        if (pbytes) asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"((const unsigned char*)P.wave + nr.ga * ES), "r"(pbytes), "l"(l2_policy_evict_last()) : "memory");
#else
        if (pbytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"((const unsigned char*)P.wave + nr.ga * ES), "r"(pbytes) : "memory");
#endif
      }
      if (next_main()) {
        m_utt = info[buf ^ 1].utt;
        m_need = (info[buf ^ 1].T + kTileFrames - 1) / kTileFrames;
        m_flag = ld_relaxed(P.fsum_done + m_utt);
      }
    }
    double st_s1 = 0.0, st_s2 = 0.0;
    if (FUSE && duty) {
      if (svc) {   // the duty tile's raw log-mel rows come back from L2 into the (now idle) waveform buffer
        const uint32_t bytes = (uint32_t)ainfo->nf * (kV2Mels * 4);
        asm volatile("fence.proxy.async;" ::: "memory");
        mbar_expect_tx(abar, bytes);
        tma_bulk_g2s(ybuf, P.out + ainfo->out_row * (int64_t)kV2Mels, bytes, abar);
      }
      if (tid < kV2Mels) {   // moments of its utterance: loaded now, used after the mel projection (L2 latency hidden)
        st_s1 = __ldcg(&P.utt_stats[((size_t)ainfo->utt * 2) * kV2Mels + tid]);
        st_s2 = __ldcg(&P.utt_stats[((size_t)ainfo->utt * 2 + 1) * kV2Mels + tid]);
      }
    }

    if (main_tile) {
    // ---- mel projection: lane = frame, this warp's filters ----
    switch (nw) {   // warp uniform
      case 1: v6_mel<1>(S, my_row, planes + lane, f_begin, f_end, wbase); break;
      case 2: v6_mel<2>(S, my_row, planes + lane, f_begin, f_end, wbase); break;
      case 3: v6_mel<3>(S, my_row, planes + lane, f_begin, f_end, wbase); break;
      case 4: v6_mel<4>(S, my_row, planes + lane, f_begin, f_end, wbase); break;
      case 5: v6_mel<5>(S, my_row, planes + lane, f_begin, f_end, wbase); break;
      case 6: v6_mel<6>(S, my_row, planes + lane, f_begin, f_end, wbase); break;
      case 7: v6_mel<7>(S, my_row, planes + lane, f_begin, f_end, wbase); break;
      default: v6_mel<8>(S, my_row, planes + lane, f_begin, f_end, wbase); break;
    }
    }   // main_tile (mel projection)
    if (FUSE && duty && tid < kV2Mels) {
      // mean / inverse deviation of the duty tile's utterance.  The moments are double (cancellation in s2 / T - mean^2);
      // the double DIVISIONS and the square root of cmvn_utt_apply_kernel would sit on the critical path of three warps
      // in every tile, so: 1 / T by one Newton step on the float reciprocal (exact to ~1e-14), 1 / sqrt(var) by MUFU.RSQ
      // + one Newton step in float (<= 1 ulp of the float that is stored anyway).
      const int T = ainfo->T;
      const double dT = (double)T;
      double r = (double)(1.0f / (float)T);
      r = r * (2.0 - dT * r);
      const double mean = st_s1 * r;
      // one frame: np.std is exactly 0 (the reference divides by it: nan / inf), whatever the rounding of the float partial sums
      const float var = T == 1 ? 0.f : (float)fmax(fma(st_s2, r, -mean * mean), 0.0);
      float y = rsqrtf(var);
      if (var > 0.f) y = y * (1.5f - 0.5f * var * y * y);
      s_norm[tid] = P.mean_norm ? (float)mean : 0.f;
      s_norm[kV2Mels + tid] = P.std_norm ? y : 1.f;
    }
    __syncthreads();
    // the Z region has been read for the last time -> the next tile's waveform may land in its upper part
    if (tid == 0 && next_main()) issue_tile(buf ^ 1);
    double m_sum = 0.0;
    if (FS && tid == 0 && next_main()) {
      // its utterance's sum: complete unless this CTA is early (rare: lag_s covers the longest utterance + three rounds of the
      // resident CTAs).  The load is issued after the branch on the counter and served by L2, where the writers' fences put
      // the sum before the counter; it is consumed after phase C.
      while (m_flag != m_need) m_flag = ld_acquire(P.fsum_done + m_utt);
      m_sum = ld_cg_f64(P.utt_fsum + m_utt);
    }
    // the tile this CTA finished in the previous iteration: its feature stores and statistics atomics were issued before
    // barriers passed since -> publish it (here, where the service thread's warp has slack); FS: and this item's sample sum
    if (FS) {
      float t8 = 0.f;
      if (svc && sum_duty) t8 = ((s_ws[0] + s_ws[1]) + (s_ws[2] + s_ws[3])) + ((s_ws[4] + s_ws[5]) + (s_ws[6] + s_ws[7]));
      if (svc && (sum_duty || prev_utt >= 0)) {
        const int s_utt = sinfo[buf].utt;
        if (sum_duty) atomicAdd(P.utt_fsum + s_utt, (double)t8);
        __threadfence();
        if (sum_duty) red_relaxed_inc(P.fsum_done + s_utt);
        if (prev_utt >= 0) { red_relaxed_inc(P.utt_done + prev_utt); prev_utt = -1; }
      }
    } else if (FUSE && svc && prev_utt >= 0) publish();
    float* part = stage;  // [3][2][80] per-group CMVN partial sums (lower part of the Z region, free after the sweep)
    if (main_tile) {

    // ---- phase C: log, store; thread = (frame group g, filter m).  For one frame the 80 threads of a group write
    // 320 contiguous bytes straight to global memory. ----
    if (tid < 3 * kV2Mels) {
      const int g = cg, m = cm;
      const float* q = planes + m * kPlaneStride + g;
      float* od = P.out + (cur.out_row + g) * (int64_t)kV2Mels + m;
      const int left = nf - g;
      const bool use_log = P.log_kind != MAFE_LOG_NONE;
      const float add = P.log_kind == MAFE_LOG_LN_PLUS ? P.log_arg : 0.f;
      const float zero_sub = P.log_kind == MAFE_LOG_LN_EPS_IF_ZERO ? 2.220446049250313e-16f : 0.f;
      float s1 = 0.f, s2 = 0.f;
      if (nf == kTileFrames && P.log_kind == MAFE_LOG_LN_EPS_IF_ZERO) {
        // full tile, conformer log: frames g + 3 i, i < 11 (g < 2) or 10; no per-element predicates
#pragma unroll
        for (int i = 0; i < 11; ++i) {
          if (i < 10 || g < 2) {
            const float a = q[3 * i];
            const float x = a == 0.f ? 2.220446049250313e-16f : a;
            float l;
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
            const float o = l * 0.69314718055994530942f;
            od[3 * i * kV2Mels] = o;
            s1 += o;
            s2 = fmaf(o, o, s2);
          }
        }
      } else {
      // frames g, g + 3, ..., g + 30: fixed trip count (f = 32 reads the pad column and is discarded)
#pragma unroll
      for (int i = 0; i < 11; ++i) {
        const float a = q[3 * i];
        float x = a + add;
        x = x == 0.f ? zero_sub : x;
        float l;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
        const float o = use_log ? l * 0.69314718055994530942f : a;
        const bool in = 3 * i < left;
        if (in) od[3 * i * kV2Mels] = o;
        const float ov = in ? o : 0.f;
        s1 += ov;
        s2 = fmaf(ov, ov, s2);
      }
      }
      if (P.utt_stats != nullptr) {
        part[(g * 2) * kV2Mels + m] = s1;
        part[(g * 2 + 1) * kV2Mels + m] = s2;
      }
    }
    }   // main_tile (phase C)
    if (FUSE && duty) {   // normalise the duty tile: y -> (y - mean) / std, 16 bytes per thread and step
      mbar_wait(abar, a_cnt & 1);
      ++a_cnt;
      const int n4 = ainfo->nf * (kV2Mels / 4);
      const float4* y4 = reinterpret_cast<const float4*>(ybuf);
      const float4* nm4 = reinterpret_cast<const float4*>(s_norm);
      float4* o4 = reinterpret_cast<float4*>(P.out + ainfo->out_row * (int64_t)kV2Mels);
      for (int q = tid; q < n4; q += kFastThreads) {
        const int m4 = q % (kV2Mels / 4);
        const float4 v = y4[q], mu = nm4[m4], iv = nm4[kV2Mels / 4 + m4];
        o4[q] = make_float4((v.x - mu.x) * iv.x, (v.y - mu.y) * iv.y, (v.z - mu.z) * iv.z, (v.w - mu.w) * iv.w);
      }
    }
    if (FS && tid == 0 && next_main()) info[buf ^ 1].neg_mu = -((float)m_sum * info[buf ^ 1].neg_mu);   // the record carried scale / (400 T)
    __syncthreads();   // the next tile's geometry (thread 0, above) and the partial sums are visible
    if (FUSE && svc && main_tile) prev_utt = cur.utt;
    // per-utterance CMVN statistics: one double atomic per (filter, moment)
    if (main_tile && P.utt_stats != nullptr && tid < 2 * kV2Mels) {
      const int m = tid % kV2Mels, which = tid / kV2Mels;
      const double sum = (double)part[which * kV2Mels + m] + (double)part[(2 + which) * kV2Mels + m] +
                         (double)part[(4 + which) * kV2Mels + m];
      atomicAdd(&P.utt_stats[((size_t)utt * 2 + which) * kV2Mels + m], sum);
    }
  }
  if (FUSE) {   // the last tile of this CTA
    __syncthreads();
    if (svc && prev_utt >= 0) publish();
  }
  if (TM) {
    __syncthreads();   // every warp has issued its last TMEM load
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*s_tm), "n"(kTmCols) : "memory");
  }
}

}  // namespace mafe
