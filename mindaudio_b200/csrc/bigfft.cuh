// bigfft.cuh -- arbitrary-length DFT of whole signals in global memory (complex128): Bluestein's chirp-z over a
// power-of-two radix-2 Stockham transform.  Used by mafe_resample_fft (scipy.signal.resample as called by
// mindaudio/data/processing.py:132-186 and through it by augment.pitch_shift, mindaudio/data/augment.py:874-901).
// The per-element steps are __host__ __device__ so tests/host/resample_host_check.cu runs the very same index maps
// on the CPU against a direct O(N^2) DFT.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace mafe {

struct cd { double x, y; };   // complex128, double2 layout

__host__ __device__ inline cd cd_mul(cd a, cd b) { return cd{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }

__host__ __device__ inline void sincospi_d(double t, double* s, double* c) {
#ifdef __CUDA_ARCH__
  sincospi(t, s, c);
#else
  *s = sin(M_PI * t);
  *c = cos(M_PI * t);
#endif
}

// exp(sign * i * pi * n^2 / N), the exponent reduced exactly in integers (n^2 mod 2N) before it meets floating point
__host__ __device__ inline cd chirp(int64_t n, int64_t N, int sign) {
  const int64_t r = (int64_t)(((unsigned long long)n * (unsigned long long)n) % (unsigned long long)(2 * N));
  double s, c;
  sincospi_d((double)r / (double)N, &s, &c);
  return cd{c, sign * s};
}

// One radix-2 Stockham (autosort) pass, element i of M/2; p = 1, 2, 4, ..., M/2.  sign -1 forward, +1 inverse (unscaled).
__host__ __device__ inline void stockham2(const cd* __restrict__ src, cd* __restrict__ dst, int64_t i, int64_t p, int64_t half,
                                          int sign) {
  const int64_t k = i & (p - 1);
  const cd u0 = src[i];
  double s, c;
  sincospi_d((double)k / (double)p, &s, &c);
  const cd u1 = cd_mul(src[i + half], cd{c, sign * s});
  const int64_t j = ((i - k) << 1) + k;
  dst[j] = cd{u0.x + u1.x, u0.y + u1.y};
  dst[j + p] = cd{u0.x - u1.x, u0.y - u1.y};
}

// Bluestein kernel sequence b[m] = exp(-sign * i pi n^2 / N) for n = m (m < N) and n = M - m (M - m < N), else 0
__host__ __device__ inline cd bluestein_b(int64_t m, int64_t N, int64_t M, int sign) {
  if (m < N) return chirp(m, N, -sign);
  if (M - m < N) return chirp(M - m, N, -sign);
  return cd{0.0, 0.0};
}

// Spectrum of scipy.signal.resample's real-input branch at full length `num` (Hermitian extension of the one-sided
// array it hands to irfft): X = DFT of the n_x input samples (all n_x bins available), k in [0, num).
__host__ __device__ inline cd resample_bin(const cd* __restrict__ X, int64_t k, int64_t n_x, int64_t num) {
  const int64_t m = num < n_x ? num : n_x;
  const bool mirror = k > num - k;                 // the negative-frequency half: conj of the partner bin
  const int64_t kk = mirror ? num - k : k;
  if (kk >= m / 2 + 1) return cd{0.0, 0.0};
  cd v = X[kk];
  if ((m & 1) == 0 && num != n_x && kk == m / 2) {  // the unpaired bin at m/2
    const double f = num < n_x ? 2.0 : 0.5;
    v.x *= f;
    v.y *= f;
  }
  if (mirror) v.y = -v.y;
  return v;
}

inline int64_t bluestein_size(int64_t N) {
  int64_t M = 1;
  while (M < 2 * N - 1) M <<= 1;
  return M < 2 ? 2 : M;
}

}  // namespace mafe
