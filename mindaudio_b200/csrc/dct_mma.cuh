// dct_mma.cuh -- the MFCC projection on the 5th-generation tensor cores: out[f][c] = sum_m clamp(logmel[f][m]) dct[m][c]
// (mindaudio/data/features.py:356-361; the clamp is amplitude_to_dB's batch-coupled top_db floor, spectrum.py:78-89).
// Included by generic.cu.
//
// A dense [frames x n_mels] x [n_mels x n_mfcc] contraction -- the one GEMM-shaped step of the path (north star: "the mel
// and DCT projections ... may use tcgen05 tensor cores only if tf32 stays inside tolerance").  One TF32 pass does not (11-bit
// operands: ~5e-4 relative on a coefficient); THREE do: x = hi + lo with hi = x truncated to TF32 (exact), lo = x - hi
// (exact in FP32), and
//     A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo          (dropped: lo x lo ~ 2^-22, truncation of lo ~ 2^-21 relative)
// which is FP32-class accuracy with FP32 accumulation in tensor memory.  The FP32 FMA version (dct_kernel, generic.cu) is
// issue bound at ~190 warp-instructions per frame (0.41 ms for 1.23 M frames); here the CUDA cores only move and split the
// operand (~20 warp-instructions per frame) and the contraction is 3 x K/8 tcgen05.mma per 128 frames.
//
// CTA = two independent groups of 128 threads (one CTA per SM: a group's A_hi / A_lo images are 2 x 128 x K x 4 bytes).  Per
// group and work item (four 32-frame tiles of the batch = up to 128 frames):
//   1. the rows come from global memory as 16-byte chunks, are clamped, split and stored in the UMMA canonical K-major
//      layout without swizzle: 8-row x 16-byte core matrices, consecutive along K (LBO = 128 B), 8-row groups SBO apart.
//      A warp step covers 8 rows x 4 chunks: whole 32-byte sectors from global memory, conflict-free 16-byte stores;
//   2. fence.proxy.async, group barrier, ONE thread issues the MMAs (M = 128, N = n_mfcc padded to 16, K = 8 per
//      instruction) against the DCT images B_hi / B_lo (same layout, built on the host, staged once per CTA) and commits
//      them to the group's mbarrier;
//   3. every thread owns one row = one TMEM lane: tcgen05.ld of its n_mfcc accumulator columns, stores to global memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace mafe {

constexpr int kDmRows = 128;        // rows (frames) per MMA tile = TMEM lanes
constexpr int kDmTmemCols = 128;    // allocation per CTA: group g accumulates in columns 64 g ..

struct DctMmaParams {
  const float* logmel;          // [total_frames][K]
  float* out;                   // [total_frames][n_mfcc]
  int K, n_mfcc, N;             // K = n_mels (multiple of 16), N = n_mfcc padded to a multiple of 16
  const float* bimg;            // device: B_hi image then B_lo image, N x K floats each, canonical layout
  const Tile* tiles;
  int n_tiles;
  const int64_t* frame_offsets;
  int tile_frames;              // 32
  const int* group_max;
  const int* utt_group;
  int db_group;
  float top_db;
};

// byte offset of element (row r, k) in the canonical K-major no-swizzle image with K columns
__host__ __device__ inline uint32_t dm_canon_off(int r, int k, int K) {
  return (uint32_t)((r >> 3) * (K * 32) + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ uint32_t dm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp, SmemDescriptor): start address, leading (K) and stride (8-row
// group) byte offsets in 16-byte units, version 1 (Blackwell), no swizzle
__device__ __forceinline__ uint64_t dm_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor (InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ inline uint32_t dm_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void dm_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void dm_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

// frames per batch tile: dct_run takes this kernel for 32-frame tiles only.  A compile-time constant on purpose: with the
// run-time value every row index paid an integer division (40 per thread and item, ~1/3 of the kernel's instructions).
constexpr int kDmTile = 32;
template <int QUADS>   // QUADS = ceil(K / 16): groups of four 16-byte chunks per row
__global__ void __launch_bounds__(256, 1) dct_mma_kernel(const DctMmaParams P) {
  extern __shared__ __align__(128) unsigned char dm_smem[];
  const int K = P.K, N = P.N;
  const uint32_t b_bytes = (uint32_t)N * K * 4;              // one B image
  const uint32_t a_bytes = (uint32_t)kDmRows * K * 4;        // one A image
  unsigned char* sB = dm_smem;                               // [B_hi | B_lo]
  const int tid = threadIdx.x, g = tid >> 7, t = tid & 127, lane = tid & 31, warp = tid >> 5;
  unsigned char* sA = dm_smem + 2 * b_bytes + (size_t)g * 2 * a_bytes;   // this group's [A_hi | A_lo]
  unsigned char* meta = dm_smem + 2 * b_bytes + 4 * (size_t)a_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(meta) + g;    // [2]
  uint32_t* s_tm = reinterpret_cast<uint32_t*>(meta + 16);
  // geometry of the item's four tiles, per group and pipeline slot: index 8 * slot + tile (first output row, frames, clamp floor)
  int64_t* s_row = reinterpret_cast<int64_t*>(meta + 32) + 16 * g;
  int* s_nf = reinterpret_cast<int*>(meta + 288) + 16 * g;
  float* s_floor = reinterpret_cast<float*>(meta + 416) + 16 * g;

  for (uint32_t i = tid; i < 2 * b_bytes / 16; i += 256)
    reinterpret_cast<float4*>(sB)[i] = reinterpret_cast<const float4*>(P.bimg)[i];
  if (t == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(dm_smem_u32(mbar)));
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dm_smem_u32(s_tm)), "n"(kDmTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the B images are read by the tensor core (async proxy)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *s_tm + 64u * g;                                     // accumulator: lanes 0..127, columns 64 g ..
  const uint32_t tmem_my = tmem_d + ((uint32_t)(32 * (warp & 3)) << 16);       // this warp's lane quarter
  const uint32_t idesc = dm_idesc(kDmRows, N);
  const uint32_t sbo = (uint32_t)K * 32;
  const uint32_t a_hi = dm_smem_u32(sA), a_lo = a_hi + a_bytes, bb_hi = dm_smem_u32(sB), bb_lo = bb_hi + b_bytes;

  const int tpi = kDmRows / kDmTile;                       // batch tiles per work item (4)
  const int n_items = (P.n_tiles + tpi - 1) / tpi;
  const int n_workers = gridDim.x * 2;
  const int per = (n_items + n_workers - 1) / n_workers;         // contiguous range of items per group
  const int w0 = (blockIdx.x * 2 + g) * per, w1 = min(n_items, w0 + per);
  const int chunks = K >> 2;                                     // 16-byte chunks per row
  uint32_t phase = 0;
  // Per item: all 4 QUADS 16-byte row loads of a thread in flight at once, split, MMAs, epilogue; the geometry of an item's
  // four tiles is resolved by four threads one item ahead (double buffered).  The two groups of the CTA overlap each other's phases.  (Requesting the rows of item i + 1 before the
  // epilogue of item i -- a software pipeline inside the group -- measured 15 % SLOWER: 0.33 vs 0.29 ms.)
  auto resolve = [&](int item, int slot) {   // threads t < tpi
    const int ti = item * tpi + t;
    int nf = 0;
    int64_t row = 0;
    float fl = -INFINITY;
    if (item < w1 && ti < P.n_tiles) {
      const Tile tile = P.tiles[ti];
      const int64_t fo = P.frame_offsets[tile.utt];
      const int64_t T = P.frame_offsets[tile.utt + 1] - fo;
      nf = (int)min((int64_t)kDmTile, T - tile.frame0);
      row = fo + tile.frame0;
      if (P.db_group != MAFE_DBGROUP_NONE) {
        const int grp = P.db_group == MAFE_DBGROUP_UTT ? tile.utt : (P.db_group == MAFE_DBGROUP_BATCH ? 0 : P.utt_group[tile.utt]);
        fl = key_to_float(P.group_max[grp]) - P.top_db;
      }
    }
    s_row[8 * slot + t] = row; s_nf[8 * slot + t] = nf; s_floor[8 * slot + t] = fl;
  };
  const int r8 = lane & 7, c = lane >> 3;
  constexpr int ITERS = 4 * QUADS;                 // (8 rows x 4 chunks) blocks per warp: 16 row groups x QUADS / 4 warps
  float4 x[ITERS];
  auto request = [&](int slot) {
#pragma unroll
    for (int j = 0; j < ITERS; ++j) {
      const int blk = (t >> 5) + 4 * j;
      const int rg = blk / QUADS, cq = blk - rg * QUADS;
      const int r = 8 * rg + r8, k16 = 4 * cq + c;
      const int q = r / kDmTile, f = r - q * kDmTile;
      x[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k16 < chunks && f < s_nf[8 * slot + q])
        x[j] = __ldg(reinterpret_cast<const float4*>(P.logmel + (s_row[8 * slot + q] + f) * K) + k16);
    }
  };
  if (w0 < w1 && t < tpi) resolve(w0, 0);
  for (int item = w0; item < w1; ++item) {
    const int slot = (item - w0) & 1;
    dm_bar(g);                                       // this item's geometry (resolved one item ago) is visible
    request(slot);
    if (t < tpi) resolve(item + 1, slot ^ 1);        // the tile -> offsets -> clamp-floor chain of the NEXT item: three dependent
                                                     // round trips to L2, off the critical path (the other slot was last read
                                                     // before the previous item's closing barrier)
    // ---- 1. clamp -> hi / lo -> canonical images ----
#pragma unroll
    for (int j = 0; j < ITERS; ++j) {
      const int blk = (t >> 5) + 4 * j;
      const int rg = blk / QUADS, cq = blk - rg * QUADS;
      const int r = 8 * rg + r8, k16 = 4 * cq + c;
      if (k16 < chunks) {
        const int q = r / kDmTile, f = r - q * kDmTile;
        float4 v = x[j];
        if (f < s_nf[8 * slot + q]) {               // rows beyond the tile stay zero (no clamp: the floor may be positive)
          const float fl = s_floor[8 * slot + q];
          v.x = fmaxf(v.x, fl); v.y = fmaxf(v.y, fl); v.z = fmaxf(v.z, fl); v.w = fmaxf(v.w, fl);
        }
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u); l.w = v.w - h.w;
        const uint32_t off = (uint32_t)rg * sbo + (uint32_t)k16 * 128u + (uint32_t)r8 * 16u;
        *reinterpret_cast<float4*>(sA + off) = h;
        *reinterpret_cast<float4*>(sA + a_bytes + off) = l;
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    dm_bar(g);
    // ---- 2. 3 x K / 8 MMAs by one thread, committed to the group's mbarrier ----
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int ks = 0; ks < (K >> 3); ++ks) {
        const uint32_t ko = (uint32_t)ks * 256u;               // two 16-byte chunks along K per instruction
        const uint64_t dah = dm_desc(a_hi + ko, 128u, sbo), dal = dm_desc(a_lo + ko, 128u, sbo);
        const uint64_t dbh = dm_desc(bb_hi + ko, 128u, sbo), dbl = dm_desc(bb_lo + ko, 128u, sbo);
        dm_mma(tmem_d, dah, dbh, idesc, ks > 0 ? 1u : 0u);
        dm_mma(tmem_d, dal, dbh, idesc, 1u);
        dm_mma(tmem_d, dah, dbl, idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(dm_smem_u32(mbar)) : "memory");
    }
    {
      uint32_t ok;
      do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(dm_smem_u32(mbar)), "r"(phase) : "memory");
      } while (!ok);
      phase ^= 1;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- 3. one row per thread: accumulator columns -> global memory ----
    {
      const int q = t / kDmTile, f = t - q * kDmTile;
      const bool valid = f < s_nf[8 * slot + q];
      float* dst = P.out + (s_row[8 * slot + q] + f) * P.n_mfcc;
      const bool vec = (P.n_mfcc & 3) == 0;
      for (int c0 = 0; c0 < P.n_mfcc; c0 += 8) {
        float v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "r"(tmem_my + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;"
                     : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]));
        if (valid) {
          if (vec && c0 + 8 <= P.n_mfcc) {
            *reinterpret_cast<float4*>(dst + c0) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(dst + c0 + 4) = make_float4(v[4], v[5], v[6], v[7]);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              if (c0 + i < P.n_mfcc) dst[c0 + i] = v[i];
          }
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    dm_bar(g);   // the images, the accumulator and the geometry may be overwritten
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*s_tm), "n"(kDmTmemCols) : "memory");
}

inline size_t dct_mma_smem_bytes(int K, int N) { return (size_t)2 * N * K * 4 + (size_t)4 * kDmRows * K * 4 + 544; }

}  // namespace mafe
