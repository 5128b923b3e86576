// stftn16.cuh -- specialised STFT kernel for n_fft = 16 * N1 with N1 = 20 (n_fft 320: deepspeech2's
// mindaudio.stft(n_fft=320, hop_length=160, win_length=320), examples/deepspeech2/dataset.py:39-41) and N1 = 25
// (n_fft 400): spectrum.stft, mindaudio/data/spectrum.py:125-278, complex64 [frame][n_fft/2 + 1] out.
// Included by fbank512.cu.
//
// Same machinery as fbank400.cuh (persistent CTAs, tile queue, staged next-tile preparation, TMA bulk load of the
// waveform tile, in-place pad pass for the first / last tile of an utterance,
// frame pairs packed a + i*b): each lane of a 16-lane group transforms N1 points in registers (4 x 5 or 5 x 5),
// W_N twiddle, the 16-point stage runs as 2 * N1 independent 16-point DFTs per warp in two rounds of 32 lanes.
// Every warp owns its two pairs from the transform to the store; rows are written with lanes = consecutive bins --
// complex64 (stft) or |X|^power (spectrogram: msaudio.Spectrogram, window pre-scaled by the normalisation).
#pragma once
#include "fft400.cuh"
#include "packed.cuh"

namespace mafe {

template <int N1>
struct StftN {
  static constexpr int kN = 16 * N1;
  static constexpr int kBins = kN / 2 + 1;
  static constexpr int kSlot = N1 * kRowStride;                 // N1 rows x 17 >= kN + 1 outputs (340 / 425 complex)
  static constexpr int kMaxHop = kN / 2;
  static constexpr int kRawBytes = (((kTileFrames - 1) * kMaxHop + kN) * 4 + 64 + 127) & ~127;
  static constexpr int kZBytes = ((kPairs * kSlot * 8) + 127) & ~127;
  static constexpr size_t kRaw = kZBytes;                       // the waveform tile (own region: the next tile is
                                                                // fetched while this one is transformed and stored)
  static constexpr size_t kBar = kRaw + kRawBytes + 64;         // (window and W_N twiddles: tensor memory, see the kernel)
  static constexpr int kCtasPerSm = (3 * (kBar + 32 + 2 * 96 + 1024) <= 227 * 1024) ? 3 : 2;
  static constexpr size_t kInfo = kBar + 32;
  static constexpr size_t kTotal = kInfo + 2 * 96;
  static_assert(kRaw % 128 == 0 && kCtasPerSm * (kTotal + 1024) <= 228 * 1024, "raw landing zone / CTAs per SM");
  static_assert(kSlot >= kN + 1, "slot holds the spectrum and the copy of bin 0");
};

struct StftNParams {
  const float* wave;
  int64_t total_samples;
  const int64_t* sample_offsets;
  const int64_t* frame_offsets;
  const Tile* tiles;
  int n_tiles;
  int hop, center, pad_mode;
  const float* window;     // [kN], pre-scaled by 1/2 (the pair separation leaves 2X)
  const float2* twn;       // [N1][32]  W_N^(t kj), t = 0..31 (t >= 16: the rotated upper half-warp, see the load)
  float* out;              // [total_frames][kBins] complex64, or float |X|^power (out_power != 0)
  int out_power;           // 0: complex STFT; 1: power / magnitude spectrogram (spectrum.spectrogram, spectrum.py:547-606)
  float power;
  int log_kind;            // MAFE_LOG_LN_PLUS on the power kind: ln(|X|^power + log_arg) -- deepspeech2's log1p(magnitude),
  float log_arg;           // examples/deepspeech2/dataset.py:42-43, written by the transform itself
  int* queue_head;
  double* utt_stats;       // [n_utts][2] sum x, sum x^2 of the written values (utt_scalar_norm), or NULL
  float2 tws[16];          // W_N1^(j1 k1): 16 (N1 = 25) or 12 (N1 = 20) entries, kernel-parameter constant bank
};

template <int N1>
__global__ void __launch_bounds__(kFastThreads, StftN<N1>::kCtasPerSm) stftn16_kernel(const __grid_constant__ StftNParams P) {
  typedef StftN<N1> G;
  constexpr int kN = G::kN, kBins = G::kBins, kSlot = G::kSlot;
  extern __shared__ __align__(128) unsigned char smem[];
  float2* Zs = reinterpret_cast<float2*>(smem);
  float* rawz = reinterpret_cast<float*>(smem + G::kRaw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G::kBar);
  int* s_work = reinterpret_cast<int*>(bars) + 4;
  F400TileInfo* info = reinterpret_cast<F400TileInfo*>(smem + G::kInfo);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = P.hop;
  // The upper half-warp's pair starts 2 * hop floats after the lower one.  When that offset puts both on (nearly) the same
  // banks (hop 160: exactly), the upper half reads its frame ROTATED by one block of 16 samples: register j holds sample
  // t + 16 ((j + 1) mod N1).  The N1-point DFT of the rotated sequence is the wanted one times W_N1^(-kj); the lane's window
  // entries and twiddles -- per-lane constants in tensor memory -- are those of t + 16, which absorbs it: W_N^((t + 16) kj).
  // (Round 1 had the TMA engine deliver a second, skewed copy of the tile: 21 KB of shared memory and twice the L2 reads.)
  const int pair_banks = (2 * hop) & 31;
  const int rot = ((pair_banks < 8 || pair_banks > 24) && (lane >> 4)) ? 1 : 0;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Per-lane constants in tensor memory (round 2; helpers and rationale in fbank512_v6.cuh): the window entries w[t + 16 j]
  // (columns 0..N1-1) and the twiddles W_N^(t kj), kj = 1..N1-1 (columns 32 + 2 (kj - 1)) depend on t = lane & 15 only.
  // From shared memory they were 58 of the ~310 shared-memory wavefronts a warp spends per tile on the kernel's busiest unit.
  uint32_t* s_tm = reinterpret_cast<uint32_t*>(bars) + 6;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(s_tm)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = *s_tm + ((uint32_t)(32 * (warp & 3)) << 16);
  constexpr int kTwBlocks = (2 * (N1 - 1) + 7) / 8;   // eight-word blocks of twiddles (4 per block)
  if (warp < 4) {
    const int tt = lane & 15;
    float c8[8];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 8; ++i) c8[i] = 8 * c + i < N1 ? P.window[tt + 16 * ((8 * c + i + rot) % N1)] : 0.f;
      tm_st8(tb + 8 * c, c8);
    }
#pragma unroll
    for (int c = 0; c < kTwBlocks; ++c) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kj = 4 * c + 1 + i;
        const float2 w = kj < N1 ? P.twn[kj * 32 + tt + 16 * rot] : make_float2(0.f, 0.f);
        c8[2 * i] = w.x; c8[2 * i + 1] = w.y;
      }
      tm_st8(tb + 32 + 8 * c, c8);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- staged preparation of the next tile by thread 0 (see fbank512_v3.cuh / fbank400.cuh) ----
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
    const int pad = P.center ? kN / 2 : 0;
    const int64_t p_lo = (int64_t)nx_tile.frame0 * hop - pad;
    const int64_t p_hi = p_lo + (int64_t)(nf - 1) * hop + kN;
    const int64_t u_lo = p_lo < 0 ? 0 : p_lo, u_hi = p_hi > L ? L : p_hi;
    const int lpad = (int)(u_lo - p_lo);   // 0 or pad (160 / 200 floats: keeps the 16 B alignment of the bulk copy)
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
    ti_.tile_len = (nf - 1) * hop + kN;
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

    // ---- load: frame pair -> N1 windowed complex points per lane (packed: (a, b) * w) ----
    c2 v[N1];
    {
      const float* xa = xr + (2 * pair) * hop + t + 16 * rot;
      const float* xb = xa + hop;
      const int last = 16 * (N1 - 1) - 16 * N1 * rot;   // the last register block wraps to block 0 in the rotated half
      const bool fa_ok = 2 * pair < cur.nf, fb_ok = 2 * pair + 1 < cur.nf;
#pragma unroll
      for (int c = 0; c < (N1 + 7) / 8; ++c) {
        float w8[8];
        tm_ld8(tb + 8 * c, w8);
        float a[8], b[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (8 * c + i < N1) {
            const int o = 8 * c + i == N1 - 1 ? last : 16 * (8 * c + i);
            a[i] = fa_ok ? xa[o] : 0.f;
            b[i] = fb_ok ? xb[o] : 0.f;
          }
        tm_wait8(w8);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (8 * c + i < N1) v[8 * c + i] = mul2(pk(a[i], b[i]), bc(w8[i]));
      }
    }
    __syncthreads();   // the waveform has been consumed: the buffer may be refilled
    if (tid == 0) {    // stage 2: publish the claim, fetch the tile record
      s_work[buf ^ 1] = nx_w;
      if (nx_w < P.n_tiles) nx_tile = P.tiles[nx_w];
    }

    // ---- stage 1: N1-point DFT in registers, twiddle W_N^(t kj), rows [kj][t] of the pair's slot ----
    if (N1 == 25) fft25p(v, P.tws); else fft20p(v, P.tws);
    {
      float2* slot = Zs + pair * kSlot;
      sts_c2(&slot[t], v[0]);
#pragma unroll
      for (int c = 0; c < kTwBlocks; ++c) {
        float w8[8];
        tm_ld8(tb + 32 + 8 * c, w8);
        tm_wait8(w8);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kj = 4 * c + 1 + i;
          if (kj < N1) sts_c2(&slot[kj * kRowStride + t], cmul(v[N1 == 25 ? fft25_pos(kj) : fft20_pos(kj)], w8[2 * i], w8[2 * i + 1]));
        }
      }
    }
    __syncwarp();
    if (tid == 0 && nx_w < P.n_tiles) load_offsets();   // stage 3

    // ---- stage 2: the warp's 2 x N1 sixteen-point DFTs over t, two rounds of 32 lanes ----
    {
      constexpr int kSecond = 2 * N1 - 32;                             // tasks of the second round (18 / 8)
      const int q0 = lane / N1, kj0 = lane - N1 * q0;                  // task = lane        (0..31)
      const int task1 = 32 + lane, q1 = task1 / N1, kj1 = task1 - N1 * q1;   // task = 32 + lane (valid for lane < kSecond)
      const bool second = lane < kSecond;
      const float2* s0 = Zs + (warp * 2 + q0) * kSlot + kj0 * kRowStride;
      const float2* s1 = Zs + (warp * 2 + (second ? q1 : 0)) * kSlot + (second ? kj1 : 0) * kRowStride;
      c2 u0[16], u1[16];
#pragma unroll
      for (int tt = 0; tt < 16; ++tt) u0[tt] = lds_c2(s0 + tt);
#pragma unroll
      for (int tt = 0; tt < 16; ++tt) u1[tt] = second ? lds_c2(s1 + tt) : 0ull;   // idle lanes issue no shared-memory wavefronts
      __syncwarp();
      fft16p(u0);
      float2* d0 = Zs + (warp * 2 + q0) * kSlot + kj0;
#pragma unroll
      for (int kt = 0; kt < 16; ++kt) sts_c2(d0 + N1 * kt, u0[fft16_pos(kt)]);
      if (kj0 == 0) sts_c2(d0 + kN, u0[fft16_pos(0)]);   // bin 0 again: its own partner
      fft16p(u1);
      float2* d1 = Zs + (warp * 2 + (second ? q1 : 0)) * kSlot + (second ? kj1 : 0);
#pragma unroll
      for (int kt = 0; kt < 16; ++kt)
        if (second) sts_c2(d1 + N1 * kt, u1[fft16_pos(kt)]);
    }
    __syncwarp();
    if (tid == 0 && nx_w < P.n_tiles) issue_tile(buf ^ 1);   // stage 4: the next tile lands while this one is stored

    // ---- emit: this warp's two pairs, lanes = consecutive bins ----
    float m1 = 0.f, m2 = 0.f;   // moments of the values this lane writes (scalar normalisation of the utterance)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int pq = warp * 2 + q;
      const int fa = 2 * pq;
      const float2* zp = Zs + pq * kSlot;
      const bool wa = fa < cur.nf, wb = fa + 1 < cur.nf;
      if (P.out_power == 0) {   // rows of kBins complex64
        float2* oa = reinterpret_cast<float2*>(P.out + (cur.out_row + fa) * (int64_t)(2 * kBins));
        float2* ob = oa + kBins;
#pragma unroll 2
        for (int k = lane; k < kBins; k += 32) {
          const c2 zk = lds_c2(zp + k), zn = lds_c2(zp + kN - k);
          // window carries the 1/2:  A = Z[k] + conj Z[N-k],  B = (Z[k] - conj Z[N-k]) / i = swap(Z[N-k]) + (-i) Z[k]
          if (wa) sts_c2(oa + k, add2(zk, cnj(zn)));
          if (wb) sts_c2(ob + k, add2(swp(zn), mni(zk)));
        }
      } else {                  // rows of kBins floats: |X|^power
        float* oa = P.out + (cur.out_row + fa) * (int64_t)kBins;
        float* ob = oa + kBins;
#pragma unroll 2
        for (int k = lane; k < kBins; k += 32) {
          const c2 zk = lds_c2(zp + k), zn = lds_c2(zp + kN - k);
          const c2 A = add2(zk, cnj(zn)), B = sub2(zk, cnj(zn));   // |B| is all the power kinds need of frame b
          const c2 A2 = mul2(A, A), B2 = mul2(B, B);
          float pa = re(A2) + im(A2), pb = re(B2) + im(B2);
          if (P.log_kind == MAFE_LOG_LN_PLUS && P.power == 1.0f) {
            // deepspeech2: ln(|X| + c).  MUFU square root and logarithm (2^-22 relative): the IEEE sqrtf / log1pf pair made
            // this emit compute bound (0.82 ms against 0.51 ms for the complex output that writes twice the bytes)
            float sa, sb, la, lb;
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sa) : "f"(pa));
            asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sb) : "f"(pb));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la) : "f"(sa + P.log_arg));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb) : "f"(sb + P.log_arg));
            pa = la * 0.69314718055994530942f;
            pb = lb * 0.69314718055994530942f;
          } else {
          if (P.power != 2.0f) {
            if (P.power == 1.0f) { pa = sqrtf(pa); pb = sqrtf(pb); }
            else { pa = powf(sqrtf(pa), P.power); pb = powf(sqrtf(pb), P.power); }
          }
          if (P.log_kind == MAFE_LOG_LN_PLUS) {
            if (P.log_arg == 1.0f) { pa = log1pf(pa); pb = log1pf(pb); }
            else { pa = logf(pa + P.log_arg); pb = logf(pb + P.log_arg); }
          }
          }
          if (wa) { oa[k] = pa; m1 += pa; m2 = fmaf(pa, pa, m2); }
          if (wb) { ob[k] = pb; m1 += pb; m2 = fmaf(pb, pb, m2); }
        }
      }
    }
    if (P.utt_stats != nullptr) {   // warp sums -> two double atomics per warp and tile
      double d1 = (double)m1, d2 = (double)m2;
      for (int o = 16; o > 0; o >>= 1) { d1 += __shfl_xor_sync(0xffffffffu, d1, o); d2 += __shfl_xor_sync(0xffffffffu, d2, o); }
      if (lane == 0) { atomicAdd(&P.utt_stats[2 * (size_t)cur.utt], d1); atomicAdd(&P.utt_stats[2 * (size_t)cur.utt + 1], d2); }
    }
    __syncthreads();   // the Z slots may be overwritten; s_work / info of the next tile are visible
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(*s_tm) : "memory");
}

// (x - mean) / std over ALL elements of an utterance from the moments accumulated by the transform (deepspeech2/dataset.py:44-47)
__global__ void __launch_bounds__(256, 8) scalar_norm_apply_kernel(float* __restrict__ feats, const Tile* __restrict__ tiles, int n_tiles,
                                                                 const int64_t* __restrict__ frame_offsets, const double* __restrict__ utt_stats,
                                                                 int dim, int tile_frames) {
  // a CTA takes a CONTIGUOUS range of tiles: consecutive tiles mostly belong to one utterance, so the dependent chain
  // tile -> offsets -> moments (three round trips to L2 with a grid-stride order) is paid once per utterance, not per tile
  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(n_tiles, t0 + per);
  int utt = -1, T = 0;
  int64_t fo = 0;
  float m = 0.f, inv = 0.f;
  for (int ti = t0; ti < t1; ++ti) {
    const Tile tile = tiles[ti];
    if (tile.utt != utt) {
      utt = tile.utt;
      fo = frame_offsets[utt];
      T = (int)(frame_offsets[utt + 1] - fo);
      const double n = (double)T * (double)dim;
      const double mean = utt_stats[2 * (size_t)utt] / n;
      double var = utt_stats[2 * (size_t)utt + 1] / n - mean * mean;
      if (var < 0.0) var = 0.0;
      m = (float)mean;
      inv = (float)(1.0 / sqrt(var));
    }
    const int nf = min(tile_frames, T - tile.frame0);
    float* p = feats + (fo + tile.frame0) * (int64_t)dim;
    const int cnt = nf * dim;
#pragma unroll 4
    for (int i = threadIdx.x; i < cnt; i += 256) p[i] = (p[i] - m) * inv;
  }
}

}  // namespace mafe
