// Host emulation of the lane/group index maps of mindaudio_b200/csrc/fft512.cuh: a 512-point complex
// FFT computed exactly the way the kernel does it (radix-2 fold, two 256-point transforms by a
// group of 16 "lanes", transpose through a [16][17] slot) is compared with a float64 direct DFT.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../mindaudio_b200/csrc/fft512.cuh"
using namespace mafe;

int main() {
  const int N = 512;
  std::vector<double> ar(N), ai(N);
  srand(7);
  for (int n = 0; n < N; ++n) {
    ar[n] = n < 400 ? (rand() / (double)RAND_MAX - 0.5) : 0.0;
    ai[n] = n < 400 ? (rand() / (double)RAND_MAX - 0.5) : 0.0;
  }
  // reference
  std::vector<double> Xr(N), Xi(N);
  for (int k = 0; k < N; ++k) {
    double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      double a = -2.0 * M_PI * (double)((long long)k * n % N) / N;
      sr += ar[n] * cos(a) - ai[n] * sin(a);
      si += ar[n] * sin(a) + ai[n] * cos(a);
    }
    Xr[k] = sr; Xi[k] = si;
  }
  // fft16 alone
  {
    cpx v[16];
    for (int n = 0; n < 16; ++n) v[n] = cx((float)ar[n], (float)ai[n]);
    fft16(v);
    double worst = 0;
    for (int k = 0; k < 16; ++k) {
      double sr = 0, si = 0;
      for (int n = 0; n < 16; ++n) {
        double a = -2.0 * M_PI * k * n / 16.0;
        sr += ar[n] * cos(a) - ai[n] * sin(a);
        si += ar[n] * sin(a) + ai[n] * cos(a);
      }
      worst = fmax(worst, fmax(fabs(v[fft16_pos(k)].x - sr), fabs(v[fft16_pos(k)].y - si)));
    }
    printf("fft16 max abs err %.3g\n", worst);
    if (worst > 1e-5) return 1;
  }
  // kernel-style 512
  std::vector<cpx> Z(N);
  for (int half = 0; half < 2; ++half) {
    cpx slot[16 * kRowStride];
    cpx regs[16][16];
    for (int t = 0; t < 16; ++t) {           // lane t holds y_half[t + 16 j]
      for (int j = 0; j < 16; ++j) {
        int n = t + 16 * j;
        cpx lo = cx((float)ar[n], (float)ai[n]), hi = cx((float)ar[n + 256], (float)ai[n + 256]);
        cpx y = half == 0 ? lo + hi : lo - hi;
        if (half == 1) {
          double a = -2.0 * M_PI * n / 512.0;
          y = cmulf(y, cx((float)cos(a), (float)sin(a)));
        }
        regs[t][j] = y;
      }
      fft16(regs[t]);
      for (int kj = 0; kj < 16; ++kj) {
        double a = -2.0 * M_PI * (t * kj) / 256.0;
        cpx v = cmulf(regs[t][fft16_pos(kj)], cx((float)cos(a), (float)sin(a)));
        slot[kj * kRowStride + t] = v;
      }
    }
    for (int u = 0; u < 16; ++u) {           // lane u gathers row u, transforms over t
      cpx v[16];
      for (int tt = 0; tt < 16; ++tt) v[tt] = slot[u * kRowStride + tt];
      fft16(v);
      for (int kt = 0; kt < 16; ++kt) Z[2 * (u + 16 * kt) + half] = v[fft16_pos(kt)];
    }
  }
  double worst = 0, scale = 0;
  for (int k = 0; k < N; ++k) {
    worst = fmax(worst, fmax(fabs(Z[k].x - Xr[k]), fabs(Z[k].y - Xi[k])));
    scale = fmax(scale, hypot(Xr[k], Xi[k]));
  }
  printf("fft512 max abs err %.3g (max |X| %.3g, rel %.3g)\n", worst, scale, worst / scale);
  return worst / scale < 1e-6 ? 0 : 1;
}
