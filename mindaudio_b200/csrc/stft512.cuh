// stft512.cuh -- specialised STFT kernel for n_fft = 512 (spectrum.stft, mindaudio/data/spectrum.py:125-278,
// e.g. BASELINE.json configs[1]: hop 256, hann, centred): complex64 [frame][257] out.  Included by fbank512.cu.
//
// HBM-bound path (hop*4 B in, 2056 B out per frame).  Same building blocks as the fbank kernel: persistent CTAs,
// dynamic tile queue, staged next-tile preparation, TMA bulk prefetch of the waveform tile, frame PAIRS packed
// as a + i*b, radix-2 fold + two 256-point register FFTs per pair.  Differences: no pre-emphasis / mean / mel;
// every WARP owns two frame pairs from the load to the store (only __syncwarp inside a tile, one __syncthreads
// per tile for the double-buffered waveform), and centre padding (constant / reflect / edge / symmetric) is an
// index map that is only evaluated in the first / last tile of an utterance.
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
  float* out;           // [total_frames][257] complex64
  int* queue_head;
};

struct StftTileInfo {
  int64_t out_row;     // first output frame of the tile
  int64_t p_lo;        // padded-signal index of the tile's first sample (may be negative)
  int64_t u_lo;        // utterance index of raw element `shift`
  int64_t off, L;      // utterance start in the flat array / length
  int64_t cov_end, end_elem, base_elem;
  int nf, shift, n_loaded, edge;
};
static_assert(sizeof(StftTileInfo) <= 96, "StftTileInfo slot");

struct StftSmem {
  static constexpr size_t kRaw0 = 0;
  static constexpr size_t kRaw1 = kStftRawBytes;
  static constexpr size_t kZ = 2 * kStftRawBytes;
  static constexpr size_t kZBytes = sizeof(float2) * kPairs * kSlotStride;
  static constexpr size_t kWin = kZ + kZBytes;
  static constexpr size_t kW512 = kWin + sizeof(float) * kNfft;
  static constexpr size_t kW256 = kW512 + sizeof(float2) * 256;
  static constexpr size_t kBar = kW256 + sizeof(float2) * 256;
  static constexpr size_t kInfo = kBar + 32;
  static constexpr size_t kTotal = kInfo + 2 * 96;
};
static_assert(StftSmem::kZ % 16 == 0 && StftSmem::kBar % 8 == 0, "smem alignment");

__global__ void __launch_bounds__(kFastThreads, 2) stft512_kernel(const StftParams P) {
  extern __shared__ __align__(128) unsigned char smem[];
  auto raw_buf = [&](int b) -> float* { return reinterpret_cast<float*>(smem + (size_t)b * kStftRawBytes); };
  float2* Zs = reinterpret_cast<float2*>(smem + StftSmem::kZ);
  float* s_win = reinterpret_cast<float*>(smem + StftSmem::kWin);
  float2* s_w512 = reinterpret_cast<float2*>(smem + StftSmem::kW512);
  float2* s_w256 = reinterpret_cast<float2*>(smem + StftSmem::kW256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + StftSmem::kBar);
  int* s_work = reinterpret_cast<int*>(smem + StftSmem::kBar) + 4;
  StftTileInfo* info = reinterpret_cast<StftTileInfo*>(smem + StftSmem::kInfo);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int hop = P.hop;
  for (int i = tid; i < kNfft; i += kFastThreads) s_win[i] = P.window[i];
  for (int i = tid; i < 256; i += kFastThreads) { s_w512[i] = P.w512[i]; s_w256[i] = P.w256t[i]; }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // ---- staged preparation of the next tile by thread 0 (see fbank512_v3.cuh) ----
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
    // bulk copy of the 16 B aligned superset of flat elements [off+u_lo, off+u_hi)
    const int64_t g_lo = nx_off + u_lo, g_hi = nx_off + u_hi;
    const int64_t ga = (g_lo * 4) & ~(int64_t)15;
    const int64_t total16 = (P.total_samples * 4) & ~(int64_t)15;
    int64_t gb = (g_hi * 4 + 15) & ~(int64_t)15;
    if (gb > total16) gb = total16;
    const uint32_t bytes = gb > ga ? (uint32_t)(gb - ga) : 0u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (bytes) {
      mbar_expect_tx(&bars[slot], bytes);
      tma_bulk_g2s(raw_buf(slot), (const unsigned char*)P.wave + ga, bytes, &bars[slot]);
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
    ti_.n_loaded = (int)(u_hi - u_lo);
    ti_.edge = (p_lo < 0 || p_hi > L) ? 1 : 0;
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
    float* rb = raw_buf(buf);
    if (cur.cov_end < cur.end_elem) {   // tail of the flat array the 16 B granule could not cover
      for (int64_t e = cur.cov_end + tid; e < cur.end_elem; e += kFastThreads) rb[e - cur.base_elem] = P.wave[e];
      __syncthreads();
    }

    // ---- fold: frame pair -> registers (window only; a = frame 2*pair, b = frame 2*pair+1) ----
    cpx v0[16], v1[16];
    {
      // padded-signal index of frame a's first sample, relative to the tile start
      const int ia = (2 * pair) * hop, ib = ia + hop;
      const bool fa_ok = 2 * pair < cur.nf, fb_ok = 2 * pair + 1 < cur.nf;
      const float* xr = rb + cur.shift;   // xr[i] = utterance sample u_lo + i
      auto sample = [&](int i) -> float {   // padded sample at tile-relative index i (edge tiles only)
        const int64_t u = pad_index_fast(cur.p_lo + i, cur.L, P.pad_mode);
        if (u < 0) return 0.f;
        const int64_t r = u - cur.u_lo;
        return (r >= 0 && r < cur.n_loaded) ? xr[r] : __ldg(P.wave + cur.off + u);
      };
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = t + 16 * j;
        float a0, b0, a1, b1;
        if (!cur.edge) {
          a0 = xr[ia + n]; a1 = xr[ia + n + 256];
          b0 = fb_ok ? xr[ib + n] : 0.f; b1 = fb_ok ? xr[ib + n + 256] : 0.f;
        } else {
          a0 = fa_ok ? sample(ia + n) : 0.f; a1 = fa_ok ? sample(ia + n + 256) : 0.f;
          b0 = fb_ok ? sample(ib + n) : 0.f; b1 = fb_ok ? sample(ib + n + 256) : 0.f;
        }
        const float w0 = s_win[n], w1 = s_win[n + 256];
        const cpx lo = cx(a0 * w0, b0 * w0), hi = cx(a1 * w1, b1 * w1);
        const float2 tw = s_w512[n];
        v0[j] = lo + hi;
        v1[j] = cmulf(lo - hi, cx(tw.x, tw.y));
      }
    }
    __syncthreads();   // every warp has read raw[buf]: it may be refilled two iterations from now
    if (tid == 0) {    // stage 2: publish the claim, fetch the tile record
      s_work[buf ^ 1] = nx_w;
      if (nx_w < P.n_tiles) nx_tile = P.tiles[nx_w];
    }

#pragma unroll
    for (int half = 0; half < 2; ++half) {
      if (half == 0) fft256_group(v0, slot, s_w256, t); else fft256_group(v1, slot, s_w256, t);
      __syncwarp();
      if (half == 0 && tid == 0 && nx_w < P.n_tiles) load_offsets();   // stage 3
      // ---- emit: this warp's two pairs; lanes sweep the 128 (+1) bins of this half ----
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int pq = warp * 2 + q;
        const int fa = 2 * pq, fb = fa + 1;
        const float2* zp = Zs + pq * kSlotStride;
        float2* oa = reinterpret_cast<float2*>(P.out + (cur.out_row + fa) * (int64_t)(2 * kBins));
        float2* ob = oa + kBins;
        const int n_kk = half == 0 ? 129 : 128;
        for (int kk = lane; kk < n_kk; kk += 32) {
          const int k = 2 * kk + half;
          const int kr = half == 0 ? ((256 - kk) & 255) : (255 - kk);
          const float2 zk = zp[kk & 255];
          const float2 zn = zp[kr];
          // window carries the 1/2:  A = Z[k] + conj Z[N-k],  B = (Z[k] - conj Z[N-k]) / i
          if (fa < cur.nf) oa[k] = make_float2(zk.x + zn.x, zk.y - zn.y);
          if (fb < cur.nf) ob[k] = make_float2(zk.y + zn.y, zn.x - zk.x);
        }
      }
      __syncwarp();
      if (half == 0 && tid == 0 && nx_w < P.n_tiles) issue_tile(buf ^ 1);   // stage 4 (raw[buf^1] was released by the barrier above)
    }
    __syncthreads();   // s_work / info of the next tile are visible; the Z slots may be overwritten
  }
}

}  // namespace mafe
