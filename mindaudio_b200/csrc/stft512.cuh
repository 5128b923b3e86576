// stft512.cuh -- specialised STFT kernel for n_fft = 512 (spectrum.stft, mindaudio/data/spectrum.py:125-278,
// e.g. BASELINE.json configs[1]: hop 256, hann, centred): complex64 [frame][257] out.  Included by fbank512.cu.
//
// HBM-bound path (hop*4 B in, 2056 B out per frame).  Same building blocks as the fbank kernel: persistent CTAs
// (2 per SM), dynamic tile queue, staged next-tile preparation, TMA bulk prefetch of the waveform tile, frame PAIRS
// packed as a + i*b, radix-2 fold + two 256-point register FFTs per pair.  Differences: no pre-emphasis / mean / mel; every
// WARP owns two frame pairs from the load to the store (only __syncwarp between the phases of a half), and centre
// padding (constant / reflect / edge / symmetric) is written in place by a short pad pass in the first / last tile
// of an utterance, so the frame loads are one code path.
#pragma once

namespace mafe {

constexpr int kStftMaxHop = 256;
constexpr int kStftRawBytes = ((kTileFrames - 1) * kStftMaxHop + kNfft) * 4 + 64;  // 33 856: aligned superset of a tile

struct StftParams {
  const float* wave;
  int64_t total_samples;
  const int64_t* sample_offsets;
  const int64_t* frame_offsets;
  const Tile* tiles;
  int n_tiles;
  int hop, center, pad_mode;
  const float* window;  // [512], pre-scaled by 1/2 (the pair separation leaves 2X)
  const float2* w512;
  const float2* w256t;
  float* out;           // [total_frames][257] complex64, or float |X|^power (out_power != 0: spectrum.spectrogram)
  int out_power;
  float power;
  int log_kind;         // MAFE_LOG_LN_PLUS on the power kind: ln(|X|^power + log_arg)
  float log_arg;
  int* queue_head;
};

struct StftTileInfo {
  int64_t out_row;     // first output frame of the tile
  int64_t p_lo;        // padded-signal index of the tile's first sample (may be negative)
  int64_t u_lo;        // utterance index of the first loaded sample
  int64_t off, L;      // utterance start in the flat array / length
  int64_t cov_end, end_elem, base_elem;
  int nf, shift, n_loaded, lpad, tile_len, pad_;
};
static_assert(sizeof(StftTileInfo) <= 96, "StftTileInfo slot");

constexpr int kStftSkew = kStftRawBytes / 4;   // floats from copy 0 to copy 1 of the tile
static_assert(kStftSkew % 32 == 16 && (kStftSkew * 4) % 16 == 0, "copy 1 must sit 16 banks from copy 0 and keep the 16 B alignment of the bulk copy");
struct StftSmem {
  static constexpr size_t kRaw = 0;                    // the waveform tile TWICE: the two frame pairs a warp folds start
                                                       // a multiple of 32 floats apart (hop 256), the upper half-warp
                                                       // reads the skewed copy -> no 2-way bank conflicts
  static constexpr size_t kZ = 2 * kStftRawBytes + 64;
  static constexpr size_t kZBytes = sizeof(float2) * kPairs * kSlotStride;
  static constexpr size_t kWin = kZ + kZBytes;
  static constexpr size_t kW512 = kWin + sizeof(float) * kNfft;
  static constexpr size_t kW256 = kW512 + sizeof(float2) * 256;
  static constexpr size_t kBar = kW256 + sizeof(float2) * 256;
  static constexpr size_t kInfo = kBar + 32;
  static constexpr size_t kTotal = kInfo + 2 * 96;
};
static_assert(StftSmem::kZ % 16 == 0 && StftSmem::kBar % 8 == 0, "smem alignment");

template <bool POWER>   // POWER: |X|^power rows of 257 floats (spectrogram) instead of complex64 rows
__global__ void __launch_bounds__(kFastThreads, 2) stft512_kernel(const __grid_constant__ StftParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* rb = reinterpret_cast<float*>(smem + StftSmem::kRaw);
  float2* Zs = reinterpret_cast<float2*>(smem + StftSmem::kZ);
  float* s_win = reinterpret_cast<float*>(smem + StftSmem::kWin);
  float2* s_w512 = reinterpret_cast<float2*>(smem + StftSmem::kW512);
  float2* s_w256 = reinterpret_cast<float2*>(smem + StftSmem::kW256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + StftSmem::kBar);
  int* s_work = reinterpret_cast<int*>(smem + StftSmem::kBar) + 4;
  StftTileInfo* info = reinterpret_cast<StftTileInfo*>(smem + StftSmem::kInfo);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = P.hop;
  (void)s_win; (void)s_w512; (void)s_w256;   // round 1 staged the tables here; they live in TMEM now
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Per-lane constants in tensor memory (round 2, see fbank512_v6.cuh): the window entries w[t + 16 j], w[t + 16 j + 256], the
  // W512 twiddles of the fold and the W256 twiddles of the transform depend on t = lane & 15 only.  From shared memory they
  // were 7 of the 11 loads per point of the fold on the kernel's busiest unit (the L1 / shared data pipe); from the lane's TMEM
  // columns they are 12 eight-word loads per pair.  Columns: 0..15 w[n], 16..31 w[n + 256], 32..63 W512^n, 64..93 W256^(t kj).
  uint32_t* s_tm = reinterpret_cast<uint32_t*>(smem + StftSmem::kBar) + 6;
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
      for (int i = 0; i < 8; ++i) c8[i] = P.window[tt + 16 * ((8 * c + i) & 15) + ((8 * c + i) >> 4) * 256];
      tm_st8(tb + 8 * c, c8);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 w = P.w512[tt + 16 * (4 * c + i)];
        c8[2 * i] = w.x; c8[2 * i + 1] = w.y;
      }
      tm_st8(tb + 32 + 8 * c, c8);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kj = 4 * c + 1 + i;
        const float2 w = kj < 16 ? P.w256t[kj * 16 + tt] : make_float2(0.f, 0.f);
        c8[2 * i] = w.x; c8[2 * i + 1] = w.y;
      }
      tm_st8(tb + 64 + 8 * c, c8);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- staged preparation of the next tile by thread 0 (see fbank512_v3.cuh) ----
  // Centre padding: the first tile of an utterance lands `pad` floats into the buffer and the last one leaves room
  // behind the data; the pad pass fills both in place.
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
    const int pad = P.center ? kNfft / 2 : 0;
    const int64_t p_lo = (int64_t)nx_tile.frame0 * hop - pad;
    const int64_t p_hi = p_lo + (int64_t)(nf - 1) * hop + kNfft;
    const int64_t u_lo = p_lo < 0 ? 0 : p_lo, u_hi = p_hi > L ? L : p_hi;   // utterance samples inside the tile's span
    const int lpad = (int)(u_lo - p_lo);   // 0 or pad (256 floats = 1 KB: keeps the 16 B alignment of the bulk copy)
    // bulk copy of the 16 B aligned superset of flat elements [off+u_lo, off+u_hi)
    const int64_t g_lo = nx_off + u_lo, g_hi = nx_off + (u_hi > u_lo ? u_hi : u_lo);
    const int64_t ga = (g_lo * 4) & ~(int64_t)15;
    const int64_t total16 = (P.total_samples * 4) & ~(int64_t)15;
    int64_t gb = (g_hi * 4 + 15) & ~(int64_t)15;
    if (gb > total16) gb = total16;
    const uint32_t bytes = gb > ga ? (uint32_t)(gb - ga) : 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (bytes) {
      mbar_expect_tx(&bars[slot], 2 * bytes);
      tma_bulk_g2s(rb + lpad, (const unsigned char*)P.wave + ga, bytes, &bars[slot]);
      tma_bulk_g2s(rb + kStftSkew + lpad, (const unsigned char*)P.wave + ga, bytes, &bars[slot]);
    } else {
      mbar_arrive(&bars[slot]);
    }
    StftTileInfo ti_;
    ti_.out_row = nx_fo0 + nx_tile.frame0;
    ti_.p_lo = p_lo;
    ti_.u_lo = u_lo;
    ti_.off = nx_off;
    ti_.L = L;
    ti_.base_elem = ga / 4;
    ti_.cov_end = bytes ? gb / 4 : g_lo;
    ti_.end_elem = g_hi;
    ti_.nf = nf;
    ti_.shift = (int)(g_lo - ga / 4);
    ti_.n_loaded = (int)(g_hi - g_lo);
    ti_.lpad = lpad;
    ti_.tile_len = (nf - 1) * hop + kNfft;
    ti_.pad_ = 0;
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
  float2* slot = Zs + pair * kSlotStride;
  for (;; buf ^= 1) {
    if (s_work[buf] >= P.n_tiles) break;
    if (tid == 0) nx_w = atomicAdd(P.queue_head, 1);   // stage 1: claim
    const StftTileInfo cur = info[buf];
    if (buf == 0) { mbar_wait(&bars[0], phase0); phase0 ^= 1; } else { mbar_wait(&bars[1], phase1); phase1 ^= 1; }
    float* xr = rb + cur.shift;   // xr[i] = padded sample p_lo + i of the utterance; data at xr[lpad .. lpad + n_loaded)
    if (cur.cov_end < cur.end_elem) {   // tail of the flat array the 16 B granule could not cover
      for (int64_t e = cur.cov_end + tid; e < cur.end_elem; e += kFastThreads) {
        const float x = P.wave[e];
        rb[cur.lpad + (e - cur.base_elem)] = x;
        rb[kStftSkew + cur.lpad + (e - cur.base_elem)] = x;
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
        xr[kStftSkew + i] = x;
      }
      __syncthreads();
    }
    // the upper half-warp's pair starts 2 * hop floats after the lower one: read the skewed copy when that offset would
    // put both on (nearly) the same banks
    const int pair_banks = (2 * hop) & 31;
    const bool use_skew = pair_banks < 8 || pair_banks > 24;
    const float* xa = xr + ((lane >> 4) && use_skew ? kStftSkew : 0) + (2 * pair) * hop + t;
    const float* xb = xa + hop;
    const bool fa_ok = 2 * pair < cur.nf, fb_ok = 2 * pair + 1 < cur.nf;

    float2 hA[2][4], hB[2][4];   // even bins of the warp's two pairs, held until the odd bins exist
    // ---- fold: frame pair -> registers (window only; a = frame 2*pair, b = frame 2*pair+1), both halves at once ----
    c2 v0[16], v1[16];
    {
      auto fold = [&](auto H256) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {          // 8 points per round: their window entries and twiddles, 4 TMEM loads
          float w0[8], w1[8], ta[8], tc[8];
          tm_ld8(tb + 8 * c, w0);
          tm_ld8(tb + 16 + 8 * c, w1);
          tm_ld8(tb + 32 + 16 * c, ta);
          tm_ld8(tb + 40 + 16 * c, tc);
          tm_wait8(w0); tm_wait8(w1); tm_wait8(ta); tm_wait8(tc);
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int j = 8 * c + jj;
            const float a0 = fa_ok ? xa[16 * j] : 0.f, a1 = fa_ok ? xa[16 * j + 256] : 0.f;
            // hop 256: frame b starts where the second half of frame a starts
            const float b0 = decltype(H256)::value ? (fb_ok ? a1 : 0.f) : (fb_ok ? xb[16 * j] : 0.f);
            const float b1 = fb_ok ? xb[16 * j + 256] : 0.f;
            const c2 lo = mul2(pk(a0, b0), bc(w0[jj])), hi = mul2(pk(a1, b1), bc(w1[jj]));
            v0[j] = add2(lo, hi);
            const float* tw = jj < 4 ? ta : tc;
            v1[j] = cmul(sub2(lo, hi), tw[2 * (jj & 3)], tw[2 * (jj & 3) + 1]);
          }
        }
      };
      if (hop == 256) fold(std::true_type{}); else fold(std::false_type{});
    }
    __syncthreads();   // every warp has read the waveform tile for the last time: the buffer may be refilled
    if (tid == 0) {    // stage 2: publish the claim, fetch the tile record
      s_work[buf ^ 1] = nx_w;
      if (nx_w < P.n_tiles) nx_tile = P.tiles[nx_w];
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      c2 (&v)[16] = half ? v1 : v0;
      {   // 256-point transform of the 16-lane group, packed arithmetic; outputs to the slot (bin 2 (t + 16 kt) + half)
        fft16p(v);
        sts_c2(slot + t, v[fft16_pos(0)]);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float tw[8];
          tm_ld8(tb + 64 + 8 * c, tw);
          tm_wait8(tw);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kj = 4 * c + 1 + i;
            if (kj < 16) sts_c2(slot + kj * kRowStride + t, cmul(v[fft16_pos(kj)], tw[2 * i], tw[2 * i + 1]));
          }
        }
        __syncwarp();
        c2 u[16];
#pragma unroll
        for (int tt = 0; tt < 16; ++tt) u[tt] = lds_c2(slot + t * kRowStride + tt);
        __syncwarp();
        fft16p(u);
#pragma unroll
        for (int kt = 0; kt < 16; ++kt) sts_c2(slot + t + 16 * kt, u[fft16_pos(kt)]);
        if (t == 0) sts_c2(slot + 256, u[fft16_pos(0)]);   // output 0 again: partner of itself
      }
      __syncwarp();
      if (half == 0 && tid == 0 && nx_w < P.n_tiles) load_offsets();   // stage 3
      // ---- emit: this warp's two pairs.  Lane l owns sub-indices kk = l + 32 i: the even bins 2kk (half 0) wait in
      // registers until the odd bins 2kk + 1 (half 1) exist, then both go out as one 16-byte piece -- whole 32-byte
      // sectors per warp store instead of two half-filled passes (which cost +54 % DRAM traffic: partially written
      // sectors were evicted and re-filled between the passes). ----
      auto pw_of = [&](float re, float im) -> float {   // |X|^power of the power-spectrogram output kind
        float p = fmaf(re, re, im * im);
        if (P.power != 2.0f) p = P.power == 1.0f ? sqrtf(p) : powf(sqrtf(p), P.power);
        if (P.log_kind == MAFE_LOG_LN_PLUS) p = P.log_arg == 1.0f ? log1pf(p) : logf(p + P.log_arg);
        return p;
      };
      if (half == 0) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float2* zp = Zs + (warp * 2 + q) * kSlotStride;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kk = lane + 32 * i;
            const float2 zk = zp[kk];
            const float2 zn = zp[256 - kk];
            // window carries the 1/2:  A = Z[k] + conj Z[N-k],  B = (Z[k] - conj Z[N-k]) / i
            hA[q][i] = make_float2(zk.x + zn.x, zk.y - zn.y);
            hB[q][i] = make_float2(zk.y + zn.y, zn.x - zk.x);
            if (POWER) { hA[q][i].x = pw_of(hA[q][i].x, hA[q][i].y); hB[q][i].x = pw_of(hB[q][i].x, hB[q][i].y); }
          }
          if (lane == 0) {   // Nyquist bin 256 (kk = 128, its own partner)
            const int fa = 2 * (warp * 2 + q);
            const float2 z = zp[128];
            if (POWER) {
              float* oa = P.out + (cur.out_row + fa) * (int64_t)kBins;
              if (fa < cur.nf) oa[256] = pw_of(2.f * z.x, 0.f);
              if (fa + 1 < cur.nf) oa[kBins + 256] = pw_of(2.f * z.y, 0.f);
            } else {
              float2* oa = reinterpret_cast<float2*>(P.out + (cur.out_row + fa) * (int64_t)(2 * kBins));
              if (fa < cur.nf) oa[256] = make_float2(2.f * z.x, 0.f);
              if (fa + 1 < cur.nf) oa[kBins + 256] = make_float2(2.f * z.y, 0.f);
            }
          }
        }
      } else if (POWER) {
        // rows of 257 floats (1028 B): bins (2kk, 2kk + 1) go out as one 8-byte piece on the even rows
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int fa = 2 * (warp * 2 + q);
          const float2* zp = Zs + (warp * 2 + q) * kSlotStride;
          float* oa = P.out + (cur.out_row + fa) * (int64_t)kBins;
          float* ob = oa + kBins;
          const bool a_aligned = ((cur.out_row + fa) & 1) == 0;
          const bool wa = fa < cur.nf, wb = fa + 1 < cur.nf;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kk = lane + 32 * i;
            const float2 zk = zp[kk];
            const float2 zn = zp[255 - kk];
            const float pa1 = pw_of(zk.x + zn.x, zk.y - zn.y), pb1 = pw_of(zk.y + zn.y, zn.x - zk.x);
            if (a_aligned) {
              if (wa) *reinterpret_cast<float2*>(oa + 2 * kk) = make_float2(hA[q][i].x, pa1);
              if (wb) { ob[2 * kk] = hB[q][i].x; ob[2 * kk + 1] = pb1; }
            } else {
              if (wa) { oa[2 * kk] = hA[q][i].x; oa[2 * kk + 1] = pa1; }
              if (wb) *reinterpret_cast<float2*>(ob + 2 * kk) = make_float2(hB[q][i].x, pb1);
            }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const int fa = 2 * (warp * 2 + q);
          const float2* zp = Zs + (warp * 2 + q) * kSlotStride;
          float2* oa = reinterpret_cast<float2*>(P.out + (cur.out_row + fa) * (int64_t)(2 * kBins));
          float2* ob = oa + kBins;
          const bool a_aligned = ((cur.out_row + fa) & 1) == 0;   // rows are 2056 B: every other one is 16 B aligned
          const bool wa = fa < cur.nf, wb = fa + 1 < cur.nf;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int kk = lane + 32 * i;
            const float2 zk = zp[kk];
            const float2 zn = zp[255 - kk];
            const float2 A1 = make_float2(zk.x + zn.x, zk.y - zn.y);
            const float2 B1 = make_float2(zk.y + zn.y, zn.x - zk.x);
            const float4 va = make_float4(hA[q][i].x, hA[q][i].y, A1.x, A1.y);
            const float4 vb = make_float4(hB[q][i].x, hB[q][i].y, B1.x, B1.y);
            if (a_aligned) {
              if (wa) *reinterpret_cast<float4*>(oa + 2 * kk) = va;
              if (wb) { ob[2 * kk] = hB[q][i]; ob[2 * kk + 1] = B1; }
            } else {
              if (wa) { oa[2 * kk] = hA[q][i]; oa[2 * kk + 1] = A1; }
              if (wb) *reinterpret_cast<float4*>(ob + 2 * kk) = vb;
            }
          }
        }
      }
      __syncwarp();
      if (half == 0 && tid == 0 && nx_w < P.n_tiles) issue_tile(buf ^ 1);   // stage 4: the tile buffer was released by the barrier above
    }
    __syncthreads();   // s_work / info of the next tile are visible; the Z slots may be overwritten
  }
  __syncthreads();   // every warp has issued its last TMEM load
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(*s_tm) : "memory");
}

}  // namespace mafe
