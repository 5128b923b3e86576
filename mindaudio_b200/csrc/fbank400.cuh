// fbank400.cuh -- specialised kernel for n_fft = 400 mel front-ends: spectrum.melspectrogram (spectrum.py:609-698),
// features.fbank (features.py:196-270), features.mfcc (:273-373) with their defaults (hann 400, reflect centre
// padding, hop 160/200, HTK triangles) -- the ECAPA-TDNN / fastspeech2-style front-end of BASELINE.json configs[0],[3].
// Included by fbank512.cu.
//
// Same machinery as the 512-point kernel (persistent CTAs, dynamic tile queue, staged next-tile preparation, TMA
// bulk load of the waveform tile, frame pairs packed a + i*b, lanes = frames for the sparse mel sweep), with a
// 400 = 25 x 16 FFT: every lane of a 16-lane group transforms 25 points in registers (5 x 5), the 16-point stage
// runs as 50 independent 16-point DFTs per warp (two rounds of 32 lanes), and the mel sweep is the table-driven,
// cost-balanced rolled loop of fbank512_v3.cuh (any filterbank with <= 2 adjacent filters per bin).  Centre padding
// is written in place by a short pad pass in the first / last tile of an utterance, so the frame loads are one code
// path.  Output: mel energies or log-mel (dB / ln), stored straight from the combine phase; the top_db clamp and
// the DCT of MFCC run in the existing follow-up kernels (db_clamp_kernel, dct_kernel).
#pragma once
#include "fft400.cuh"
#include "packed.cuh"

namespace mafe {

constexpr int kN400 = 400;
constexpr int kBins400 = 201;
constexpr int kSlot400 = 425;                     // 25 rows x 17 (transpose) >= 401 outputs; 425*8 B = 18 banks mod 32
constexpr int kRaw400Bytes = 26496;               // (31 * 200 + 400) * 4 + slack, 128 B multiple
constexpr int kZ400Bytes = kPairs * kSlot400 * 8; // 54 400
constexpr int kRaw400InZ = kZ400Bytes - kRaw400Bytes;   // 27 904: the waveform tile lands in the upper part of Z
constexpr int kMaxHop400 = 200;
constexpr int kMaxMels400 = 128;
constexpr int kMaxRows400 = kMaxMels400 + 2 + 2 * (kFastWarps - 1) + 1;   // compact planes: emitted filters + zero row
static_assert(kRaw400InZ % 128 == 0, "raw landing zone alignment");

// Sweep program (see V3Sweep in fbank512_v3.cuh): warp w owns bins kk0[w] .. kk0[w+1]-1; step k first retires nret
// filters, then accumulates bin k into filters (cur, cur + 1) with weights (w0, w1).  Cost balanced on the host.
struct F400Sweep {
  V3Step step[kBins400 + 3];
  uint32_t nret_mask[kFastWarps][kV3MaskWords];   // retire counts of a warp's range, 2 bits per step (register resident)
  unsigned char tail[kFastWarps];
  unsigned char kk0[kFastWarps + 1];
  unsigned char row0[kFastWarps];
  int zero_row;
};

struct F400Params {
  const float* wave;
  int64_t total_samples;
  const int64_t* sample_offsets;
  const int64_t* frame_offsets;
  const Tile* tiles;
  int n_tiles;
  int hop, center, pad_mode, n_mels;
  int log_kind;
  float log_arg, log_mult, log_offset;
  const float* window;     // [400], pre-scaled by spec_scale * wave_scale * 1/2
  const float2* tw400;     // [25][32]  W400^(t kj), t = 0..31 (t >= 16: the rotated upper half-warp, see stftn16.cuh)
  const int* combine;      // [n_mels]: plane rows A | B << 8 holding the filter's partial sums
  float* out;              // [total_frames][n_mels]
  float* aux_mel;          // mafe_frontend_run_aux: the mel energies before the log, same layout; or null
  int* queue_head;
  int* group_max;          // dB maxima (ordered-int keys) or null
  const int* utt_group;
  int db_group;
  int plane_rows;          // rows of the compact mel planes incl. the zero row
  float2 tw25[16];         // W25^(j1 k1), j1,k1 = 1..4: kernel-parameter constant bank
};

struct F400TileInfo {
  int64_t out_row, p_lo, u_lo, off, L, cov_end, end_elem, base_elem;
  int nf, shift, n_loaded, lpad, utt, tile_len;
};
static_assert(sizeof(F400TileInfo) <= 96, "F400TileInfo slot");

// dynamic shared memory: [Z 54400][planes rows*33*4][bars 32][info 192]   (window and W400 twiddles: tensor memory)
__host__ __device__ inline size_t f400_planes_bytes(int rows) { return ((size_t)rows * kPlaneStride * 4 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t f400_smem_bytes(int rows) { return kZ400Bytes + f400_planes_bytes(rows) + 32 + 192; }
static_assert((kRaw400InZ / 4) % 32 == 0, "the tile starts on bank 0");
#ifndef MAFE_F400_CTAS
#define MAFE_F400_CTAS 2
#endif
static_assert(MAFE_F400_CTAS * (kZ400Bytes + ((kMaxRows400 * kPlaneStride * 4 + 15) & ~15) + 32 + 192 + 1024) <= 227 * 1024, "CTAs per SM");

__global__ void __launch_bounds__(kFastThreads, MAFE_F400_CTAS) fbank400_kernel(const __grid_constant__ F400Params P,
                                                                   const __grid_constant__ F400Sweep S) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int nm = P.n_mels;
  float2* Zs = reinterpret_cast<float2*>(smem);
  float* rawz = reinterpret_cast<float*>(smem + kRaw400InZ);
  float* planes = reinterpret_cast<float*>(smem + kZ400Bytes);
  unsigned char* tail_base = smem + kZ400Bytes + f400_planes_bytes(P.plane_rows);
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail_base);
  int* s_work = reinterpret_cast<int*>(bars) + 4;
  F400TileInfo* info = reinterpret_cast<F400TileInfo*>(tail_base + 32);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = P.hop;
  // The upper half-warp's pair starts 2 * hop floats after the lower one (hop 160: same bank).  When the two would collide it
  // reads its frame rotated by one block of 16 samples; the per-lane constants in tensor memory (window entries, W400
  // twiddles) are those of t + 16, which absorbs the rotation -- see stftn16.cuh (round 1: a second, skewed copy of the tile).
  const int pair_banks = (2 * hop) & 31;
  const int rot = ((pair_banks < 8 || pair_banks > 24) && (lane >> 4)) ? 1 : 0;
  if (tid < kPlaneStride) planes[S.zero_row * kPlaneStride + tid] = 0.f;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // per-lane constants in tensor memory (helpers in fbank512_v6.cuh): columns 0..24 w[t + 16 j], 32 + 2 (kj - 1) W400^(t kj)
  uint32_t* s_tm = reinterpret_cast<uint32_t*>(bars) + 6;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(s_tm)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = *s_tm + ((uint32_t)(32 * (warp & 3)) << 16);
  if (warp < 4) {
    const int tt = lane & 15;
    float c8[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c8[i] = 8 * c + i < 25 ? P.window[tt + 16 * ((8 * c + i + rot) % 25)] : 0.f;
      tm_st8(tb + 8 * c, c8);
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kj = 4 * c + 1 + i;
        const float2 w = kj < 25 ? P.tw400[kj * 32 + tt + 16 * rot] : make_float2(0.f, 0.f);
        c8[2 * i] = w.x; c8[2 * i + 1] = w.y;
      }
      tm_st8(tb + 32 + 8 * c, c8);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- staged preparation of the next tile by thread 0 (see fbank512_v3.cuh) ----
  // Centre padding: the first tile of an utterance lands `pad` floats into the buffer and the last one leaves room
  // behind the data, so that a short "pad pass" can write the reflected / replicated / zero samples in place and the
  // frame loads are the same code for every tile.
  int nx_w = P.n_tiles;
  Tile nx_tile = {0, 0};
  int64_t nx_off = 0, nx_off1 = 0, nx_fo0 = 0, nx_fo1 = 0;
  auto load_offsets = [&]() {
    nx_off = P.sample_offsets[nx_tile.utt];
    nx_off1 = P.sample_offsets[nx_tile.utt + 1];
    nx_fo0 = P.frame_offsets[nx_tile.utt];
    nx_fo1 = P.frame_offsets[nx_tile.utt + 1];
  };
  auto issue_tile = [&](int slot) {
    const int T = (int)(nx_fo1 - nx_fo0);
    const int64_t L = nx_off1 - nx_off;
    const int nf = min(kTileFrames, T - nx_tile.frame0);
    const int pad = P.center ? kN400 / 2 : 0;
    const int64_t p_lo = (int64_t)nx_tile.frame0 * hop - pad;
    const int64_t p_hi = p_lo + (int64_t)(nf - 1) * hop + kN400;
    const int64_t u_lo = p_lo < 0 ? 0 : p_lo, u_hi = p_hi > L ? L : p_hi;
    const int lpad = (int)(u_lo - p_lo);   // 0 or pad (200 floats = 800 B: keeps the 16 B alignment of the bulk copy)
    const int64_t g_lo = nx_off + u_lo, g_hi = nx_off + (u_hi > u_lo ? u_hi : u_lo);
    const int64_t ga = (g_lo * 4) & ~(int64_t)15;
    const int64_t total16 = (P.total_samples * 4) & ~(int64_t)15;
    int64_t gb = (g_hi * 4 + 15) & ~(int64_t)15;
    if (gb > total16) gb = total16;
    const uint32_t bytes = gb > ga ? (uint32_t)(gb - ga) : 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (bytes) {
      mbar_expect_tx(&bars[slot], bytes);
      tma_bulk_g2s(rawz + lpad, (const unsigned char*)P.wave + ga, bytes, &bars[slot]);
    } else {
      mbar_arrive(&bars[slot]);
    }
    F400TileInfo ti_;
    ti_.out_row = nx_fo0 + nx_tile.frame0;
    ti_.p_lo = p_lo; ti_.u_lo = u_lo; ti_.off = nx_off; ti_.L = L;
    ti_.base_elem = ga / 4;
    ti_.cov_end = bytes ? gb / 4 : g_lo;
    ti_.end_elem = g_hi;
    ti_.nf = nf;
    ti_.shift = (int)(g_lo - ga / 4);
    ti_.n_loaded = (int)(g_hi - g_lo);
    ti_.lpad = lpad;
    ti_.utt = nx_tile.utt;
    ti_.tile_len = (nf - 1) * hop + kN400;
    info[slot] = ti_;
  };
  if (tid == 0) {
    nx_w = atomicAdd(P.queue_head, 1);
    s_work[0] = nx_w;
    if (nx_w < P.n_tiles) { nx_tile = P.tiles[nx_w]; load_offsets(); issue_tile(0); }
  }
  __syncthreads();

  uint32_t phase0 = 0, phase1 = 0;
  int buf = 0;
  const int t = lane & 15;
  const int pair = warp * 2 + (lane >> 4);
  // combine role of this thread: (frame group cg, filter cm)
  const int G = kFastThreads / nm;                  // frame groups (nm <= 128 -> G >= 2)
  const int cg = tid / nm, cm = tid - cg * nm;
  const int crow = cg < G ? P.combine[cm] : 0;
  for (;; buf ^= 1) {
    if (s_work[buf] >= P.n_tiles) break;
    if (tid == 0) nx_w = atomicAdd(P.queue_head, 1);   // stage 1: claim
    const F400TileInfo cur = info[buf];
    if (buf == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1; } else { mbar_wait(&bars[1], phase1); phase1 ^= 1; }
    float* xr = rawz + cur.shift;   // xr[i] = padded sample p_lo + i of the utterance; data at xr[lpad .. lpad + n_loaded)
    if (cur.cov_end < cur.end_elem) {   // bytes the 16 B-granular bulk copy could not cover (end of the flat array)
      for (int64_t e = cur.cov_end + tid; e < cur.end_elem; e += kFastThreads) {
        rawz[cur.lpad + (e - cur.base_elem)] = P.wave[e];
      }
      __syncthreads();
    }
    if (cur.lpad > 0 || cur.lpad + cur.n_loaded < cur.tile_len) {
      // ---- pad pass (first / last tile of an utterance): padded samples outside the loaded range, in place ----
      const int d_end = cur.lpad + cur.n_loaded;
      const int n_fill = cur.lpad + (cur.tile_len - d_end);
#pragma unroll 1
      for (int e = tid; e < n_fill; e += kFastThreads) {
        const int i = e < cur.lpad ? e : d_end + (e - cur.lpad);
        const int64_t u = pad_index_fast(cur.p_lo + i, cur.L, P.pad_mode);
        float x = 0.f;
        if (u >= 0) {
          const int64_t r = u - cur.u_lo;
          x = (r >= 0 && r < cur.n_loaded) ? xr[cur.lpad + r] : __ldg(P.wave + cur.off + u);
        }
        xr[i] = x;
      }
      __syncthreads();
    }

    // ---- load: frame pair -> 25 windowed complex points per lane (packed: (a, b) * w) ----
    c2 v[25];
    {
      const float* xa = xr + (2 * pair) * hop + t + 16 * rot;
      const float* xb = xa + hop;
      const int last = 16 * 24 - 16 * 25 * rot;   // the last register block wraps to block 0 in the rotated half
      const bool fa_ok = 2 * pair < cur.nf, fb_ok = 2 * pair + 1 < cur.nf;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float w8[8];
        tm_ld8(tb + 8 * c, w8);
        float a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (8 * c + i < 25) {
            const int o = 8 * c + i == 24 ? last : 16 * (8 * c + i);
            a[i] = fa_ok ? xa[o] : 0.f;
            b[i] = fb_ok ? xb[o] : 0.f;
          }
        tm_wait8(w8);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (8 * c + i < 25) v[8 * c + i] = mul2(pk(a[i], b[i]), bc(w8[i]));
      }
    }
    __syncthreads();   // the waveform has been consumed: the Z region may be written
    if (tid == 0) {    // stage 2: publish the claim, fetch the tile record
      s_work[buf ^ 1] = nx_w;
      if (nx_w < P.n_tiles) nx_tile = P.tiles[nx_w];
    }

    // ---- stage 1: 25-point DFT in registers, twiddle W400^(t kj), rows [kj][t] of the pair's slot ----
    fft25p(v, P.tw25);
    {
      float2* slot = Zs + pair * kSlot400;
      sts_c2(&slot[t], v[0]);
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        float w8[8];
        tm_ld8(tb + 32 + 8 * c, w8);
        tm_wait8(w8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kj = 4 * c + 1 + i;
          sts_c2(&slot[kj * kRowStride + t], cmul(v[fft25_pos(kj)], w8[2 * i], w8[2 * i + 1]));
        }
      }
    }
    __syncwarp();
    if (tid == 0 && nx_w < P.n_tiles) load_offsets();   // stage 3

    // ---- stage 2: the warp's 2 x 25 sixteen-point DFTs over t, two rounds of 32 lanes ----
    {
      const int q0 = lane / 25, kj0 = lane - 25 * q0;                 // task = lane        (0..31)
      const int task1 = 32 + lane, q1 = task1 / 25, kj1 = task1 - 25 * q1;   // task = 32 + lane (valid for lane < 18)
      const bool second = lane < 18;
      const float2* s0 = Zs + (warp * 2 + q0) * kSlot400 + kj0 * kRowStride;
      const float2* s1 = Zs + (warp * 2 + (second ? q1 : 0)) * kSlot400 + (second ? kj1 : 0) * kRowStride;
      c2 u0[16], u1[16];
#pragma unroll
      for (int tt = 0; tt < 16; ++tt) u0[tt] = lds_c2(s0 + tt);
#pragma unroll
      for (int tt = 0; tt < 16; ++tt) u1[tt] = second ? lds_c2(s1 + tt) : 0ull;   // idle lanes issue no shared-memory wavefronts
      __syncwarp();
      fft16p(u0);
      float2* d0 = Zs + (warp * 2 + q0) * kSlot400 + kj0;
#pragma unroll
      for (int kt = 0; kt < 16; ++kt) sts_c2(d0 + 25 * kt, u0[fft16_pos(kt)]);
      if (kj0 == 0) sts_c2(d0 + kN400, u0[fft16_pos(0)]);   // bin 0 again: its own partner
      fft16p(u1);
      float2* d1 = Zs + (warp * 2 + (second ? q1 : 0)) * kSlot400 + (second ? kj1 : 0);
#pragma unroll
      for (int kt = 0; kt < 16; ++kt)
        if (second) sts_c2(d1 + 25 * kt, u1[fft16_pos(kt)]);
    }
    __syncthreads();   // every pair's spectrum is in its slot

    // ---- sweep: lanes = frames, this warp's bin range (one rolled loop, program in the parameter bank) ----
    {
      const float2* zp = Zs + (lane >> 1) * kSlot400;
      const float sgn = (lane & 1) ? -1.f : 1.f;
      float* dst = planes + (int)S.row0[warp] * kPlaneStride + lane;
      const int k0 = S.kk0[warp];
      uint32_t ak = smem_u32(zp + k0);
      uint32_t an = smem_u32(zp + kN400 - k0);
      int si = k0;
      const int si_end = S.kk0[warp + 1];
      float acc_lo = 0.f, acc_hi = 0.f;
      auto retire = [&](int n) {
#pragma unroll 1
        do {
          *dst = acc_lo;
          acc_lo = acc_hi; acc_hi = 0.f; dst += kPlaneStride;
        } while (--n);
      };
      int word = 0;
#pragma unroll 1
      do {
        uint32_t m = S.nret_mask[warp][word++];   // retire counts of the next 16 steps
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
      retire(S.tail[warp]);
    }
    __syncthreads();   // Z has been read for the last time
    if (tid == 0 && nx_w < P.n_tiles) issue_tile(buf ^ 1);   // stage 4: next tile's waveform into the upper Z region

    // ---- combine, log, store straight to global memory (a frame group writes nm contiguous floats per frame),
    //      track the dB maximum ----
    {
      float vmax = -INFINITY;
      if (cg < G && cg < cur.nf) {
        // frames cg, cg + G, ... of the tile: running pointers (the loop used to rebuild two 64-bit addresses and walk the
        // log-kind branches per element: ~35 instructions for 2 loads, 1 add, 1 log, 1-2 stores), log kind resolved outside
        const float* pa = planes + (crow & 0xff) * kPlaneStride + cg;
        const float* pb = planes + (crow >> 8) * kPlaneStride + cg;
        const int64_t stride = (int64_t)G * nm;
        float* od = P.out + (cur.out_row + cg) * (int64_t)nm + cm;
        float* ad = P.aux_mel ? P.aux_mel + (cur.out_row + cg) * (int64_t)nm + cm : nullptr;
        const int nf = cur.nf;
        auto run = [&](auto emit) {
          if (ad) {
#pragma unroll 2
            for (int f = cg; f < nf; f += G, pa += G, pb += G, od += stride, ad += stride) {
              const float e = *pa + *pb;
              *ad = e;
              *od = emit(e);
            }
          } else {
#pragma unroll 2
            for (int f = cg; f < nf; f += G, pa += G, pb += G, od += stride) *od = emit(*pa + *pb);
          }
        };
        switch (P.log_kind) {
          case MAFE_LOG_DB: {
            const float fl = P.log_arg, mult = P.log_mult * 0.30102999566398119521f, noff = -P.log_offset;
            run([&](float e) {
              float l2;
              asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(fmaxf(e, fl)));
              const float o = fmaf(mult, l2, noff);
              vmax = fmaxf(vmax, o);
              return o;
            });
            break;
          }
          case MAFE_LOG_LN_PLUS: {
            const float add = P.log_arg;
            run([&](float e) {
              float l2;
              asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(e + add));
              return l2 * 0.69314718055994530942f;
            });
            break;
          }
          case MAFE_LOG_LN_EPS_IF_ZERO:
            run([&](float e) {
              float l2;
              asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(e == 0.f ? 2.220446049250313e-16f : e));
              return l2 * 0.69314718055994530942f;
            });
            break;
          default: run([](float e) { return e; }); break;
        }
      }
      if (P.log_kind == MAFE_LOG_DB && P.db_group != MAFE_DBGROUP_NONE) {
        for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        if (lane == 0 && vmax > -INFINITY) {
          const int grp = P.db_group == MAFE_DBGROUP_UTT ? cur.utt : (P.db_group == MAFE_DBGROUP_BATCH ? 0 : P.utt_group[cur.utt]);
          atomicMax(&P.group_max[grp], ordered_key(vmax));
        }
      }
    }
    __syncthreads();   // planes are free again; s_work / info of the next tile are visible
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(*s_tm) : "memory");
}

}  // namespace mafe
