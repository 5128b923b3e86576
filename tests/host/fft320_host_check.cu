// Host emulation of the 320-point complex FFT of mindaudio_b200/csrc/stftn16.cuh (20-point in-register DFT per lane,
// W320 twiddle, 16-point DFT over the lanes) vs a float64 DFT.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../mindaudio_b200/csrc/fft400.cuh"
using namespace mafe;
struct f2 { float x, y; };

int main() {
  const int N = 320, N1 = 20;
  std::vector<double> ar(N), ai(N);
  srand(12);
  for (int n = 0; n < N; ++n) { ar[n] = rand() / (double)RAND_MAX - 0.5; ai[n] = rand() / (double)RAND_MAX - 0.5; }
  f2 tw20[12];
  for (int j1 = 1; j1 < 5; ++j1)
    for (int k1 = 1; k1 < 4; ++k1) {
      double a = -2.0 * M_PI * (j1 * k1) / 20.0;
      tw20[(j1 - 1) * 3 + (k1 - 1)] = f2{(float)cos(a), (float)sin(a)};
    }
  {
    cpx v[20];
    for (int n = 0; n < N1; ++n) v[n] = cx((float)ar[n], (float)ai[n]);
    fft20(v, tw20);
    double worst = 0;
    for (int k = 0; k < N1; ++k) {
      double sr = 0, si = 0;
      for (int n = 0; n < N1; ++n) {
        double a = -2.0 * M_PI * k * n / N1;
        sr += ar[n] * cos(a) - ai[n] * sin(a);
        si += ar[n] * sin(a) + ai[n] * cos(a);
      }
      worst = fmax(worst, fmax(fabs(v[fft20_pos(k)].x - sr), fabs(v[fft20_pos(k)].y - si)));
    }
    printf("fft20 max abs err %.3g\n", worst);
    if (worst > 1e-5) return 1;
  }
  std::vector<cpx> Z(N), rows(N1 * 17);
  for (int t = 0; t < 16; ++t) {
    cpx v[20];
    for (int j = 0; j < N1; ++j) v[j] = cx((float)ar[t + 16 * j], (float)ai[t + 16 * j]);
    fft20(v, tw20);
    for (int kj = 0; kj < N1; ++kj) {
      double a = -2.0 * M_PI * (t * kj) / (double)N;
      rows[kj * 17 + t] = cmulf(v[fft20_pos(kj)], cx((float)cos(a), (float)sin(a)));
    }
  }
  for (int kj = 0; kj < N1; ++kj) {
    cpx u[16];
    for (int t = 0; t < 16; ++t) u[t] = rows[kj * 17 + t];
    fft16(u);
    for (int kt = 0; kt < 16; ++kt) Z[kj + N1 * kt] = u[fft16_pos(kt)];
  }
  double worst = 0, scale = 0;
  for (int k = 0; k < N; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      double a = -2.0 * M_PI * (double)((long long)k * n % N) / N;
      sr += ar[n] * cos(a) - ai[n] * sin(a);
      si += ar[n] * sin(a) + ai[n] * cos(a);
    }
    worst = fmax(worst, fmax(fabs(Z[k].x - sr), fabs(Z[k].y - si)));
    scale = fmax(scale, hypot(sr, si));
  }
  printf("fft320 max abs err %.3g (max |X| %.3g, rel %.3g)\n", worst, scale, worst / scale);
  return worst / scale < 1e-6 ? 0 : 1;
}
