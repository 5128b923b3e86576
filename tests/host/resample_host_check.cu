// Host run of the device resampler's element steps (mindaudio_b200/csrc/bigfft.cuh: chirp, Bluestein kernel sequence,
// radix-2 Stockham passes, scipy.signal.resample's spectrum rule) against a direct O(N^2) evaluation of
// irfft(rfft(x)[:m/2+1] (unpaired bin adjusted), num) * num / n_x.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../mindaudio_b200/csrc/bigfft.cuh"
using namespace mafe;

static std::vector<cd> fft_pow2(std::vector<cd> a, int sign) {
  const int64_t M = (int64_t)a.size(), half = M / 2;
  std::vector<cd> t(M);
  for (int64_t p = 1; p < M; p <<= 1) {
    for (int64_t i = 0; i < half; ++i) stockham2(a.data(), t.data(), i, p, half, sign);
    a.swap(t);
  }
  return a;
}

// DFT of length N with the given sign through the same Bluestein steps the kernels run
static std::vector<cd> bluestein(const std::vector<cd>& in, int64_t N, int64_t M, int sign) {
  std::vector<cd> a(M, cd{0, 0}), b(M);
  for (int64_t m = 0; m < N; ++m) a[m] = cd_mul(in[m], chirp(m, N, sign));
  for (int64_t m = 0; m < M; ++m) b[m] = bluestein_b(m, N, M, sign);
  std::vector<cd> af = fft_pow2(a, -1), bf = fft_pow2(b, -1);
  for (int64_t m = 0; m < M; ++m) af[m] = cd_mul(af[m], bf[m]);
  std::vector<cd> c = fft_pow2(af, +1), out(N);
  for (int64_t k = 0; k < N; ++k) {
    cd v = cd_mul(c[k], chirp(k, N, sign));
    out[k] = cd{v.x / M, v.y / M};
  }
  return out;
}

static double check(int64_t n_x, int64_t num) {
  std::vector<double> x(n_x);
  for (auto& v : x) v = rand() / (double)RAND_MAX - 0.5;
  int64_t big = n_x > num ? n_x : num;
  const int64_t M = bluestein_size(big);
  std::vector<cd> xin(n_x);
  for (int64_t n = 0; n < n_x; ++n) xin[n] = cd{x[n], 0.0};
  std::vector<cd> X = bluestein(xin, n_x, M, -1);
  std::vector<cd> Y(num);
  for (int64_t k = 0; k < num; ++k) Y[k] = resample_bin(X.data(), k, n_x, num);
  std::vector<cd> y = bluestein(Y, num, M, +1);
  // direct evaluation
  const int64_t m = num < n_x ? num : n_x, m2 = m / 2 + 1;
  std::vector<cd> R(m2);
  for (int64_t k = 0; k < m2; ++k) {
    double sr = 0, si = 0;
    for (int64_t n = 0; n < n_x; ++n) {
      const double a = -2.0 * M_PI * (double)((k * n) % n_x) / (double)n_x;
      sr += x[n] * cos(a);
      si += x[n] * sin(a);
    }
    R[k] = cd{sr, si};
  }
  if (m % 2 == 0 && num != n_x) { const double f = num < n_x ? 2.0 : 0.5; R[m / 2].x *= f; R[m / 2].y *= f; }
  double worst = 0;
  for (int64_t j = 0; j < num; ++j) {
    double acc = R[0].x;                                       // irfft: imaginary part of DC ignored
    for (int64_t k = 1; k < m2; ++k) {
      const double a = 2.0 * M_PI * (double)((k * j) % num) / (double)num;
      if (2 * k == num) acc += R[k].x * cos(a);                // Nyquist of the output: once, real part only
      else acc += 2.0 * (R[k].x * cos(a) - R[k].y * sin(a));
    }
    const double ref = acc / (double)num * ((double)num / (double)n_x);
    const double got = y[j].x / (double)n_x;                   // what bs_post_real_kernel writes
    worst = fmax(worst, fabs(got - ref));
  }
  printf("n_x %lld -> num %lld (M %lld): max abs err %.3g\n", (long long)n_x, (long long)num, (long long)M, worst);
  return worst;
}

int main() {
  srand(7);
  const int64_t cases[][2] = {{37, 50}, {48, 36}, {48, 72}, {50, 25}, {64, 64}, {31, 17}, {40, 41}, {41, 40}, {1, 3}, {3, 1},
                              {2, 2}, {2, 5}, {1000, 1499}, {1499, 1000}, {1024, 512}, {997, 1009}};
  double worst = 0;
  for (auto& c : cases) worst = fmax(worst, check(c[0], c[1]));
  // plain power-of-two transform against the direct DFT
  {
    const int64_t M = 64;
    std::vector<cd> a(M);
    for (auto& v : a) v = cd{rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};
    std::vector<cd> f = fft_pow2(a, -1);
    double w2 = 0;
    for (int64_t k = 0; k < M; ++k) {
      double sr = 0, si = 0;
      for (int64_t n = 0; n < M; ++n) {
        const double ang = -2.0 * M_PI * (double)((k * n) % M) / (double)M;
        sr += a[n].x * cos(ang) - a[n].y * sin(ang);
        si += a[n].x * sin(ang) + a[n].y * cos(ang);
      }
      w2 = fmax(w2, fmax(fabs(f[k].x - sr), fabs(f[k].y - si)));
    }
    printf("stockham 64: max abs err %.3g\n", w2);
    worst = fmax(worst, w2);
  }
  return worst < 1e-11 ? 0 : 1;
}
