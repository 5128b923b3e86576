// fft512.cuh -- register-resident 512-point complex FFT building blocks for the specialised
// front-end kernels (fbank512.cu).
//
// One 512-point complex sequence z[n] = a[n] + i*b[n] carries a PAIR of real frames (a, b); it is
// folded once (radix-2 DIF: y0[n] = z[n] + z[n+256] -> even bins, y1[n] = (z[n] - z[n+256]) W512^n
// -> odd bins) into two 256-point transforms.  A 256-point transform is done by a GROUP OF 16
// LANES, 16 points per lane, as radix-16 (registers) -> twiddle -> transpose through shared memory
// -> radix-16 (registers).  The radix-16 butterfly is 4x4 (two radix-4 layers).
//
// Functions are __host__ __device__ so the index maps are unit-tested on the CPU
// (tests/test_fft512_host.py compiles tests/fft512_host_check.cu with nvcc's host compiler).
#pragma once
#include <cuda_runtime.h>

#ifndef MAFE_HD
#define MAFE_HD __host__ __device__ __forceinline__
#endif

namespace mafe {

struct cpx {
  float x, y;
};

MAFE_HD cpx cx(float x, float y) { cpx r; r.x = x; r.y = y; return r; }
MAFE_HD cpx operator+(cpx a, cpx b) { return cx(a.x + b.x, a.y + b.y); }
MAFE_HD cpx operator-(cpx a, cpx b) { return cx(a.x - b.x, a.y - b.y); }
// a * b
MAFE_HD cpx cmulf(cpx a, cpx b) { return cx(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x)); }
// a * (-i)
MAFE_HD cpx mul_neg_i(cpx a) { return cx(a.y, -a.x); }

// forward 4-point DFT, in place: (a0,a1,a2,a3) -> (X0,X1,X2,X3)
MAFE_HD void dft4(cpx& a0, cpx& a1, cpx& a2, cpx& a3) {
  cpx s02 = a0 + a2, d02 = a0 - a2;
  cpx s13 = a1 + a3, d13 = mul_neg_i(a1 - a3);
  a0 = s02 + s13;
  a1 = d02 + d13;
  a2 = s02 - s13;
  a3 = d02 - d13;
}

// position of output bin k (0..15) inside v[] after fft16(): X[4*k1 + k2] lives at v[k1 + 4*k2]
MAFE_HD constexpr int fft16_pos(int k) { return (k >> 2) + 4 * (k & 3); }

// forward 16-point DFT of v[0..15] (natural order in); output bin k at v[fft16_pos(k)]
MAFE_HD void fft16(cpx* v) {
  // layer 1: for each n1, DFT4 over n2 of x[n1 + 4 n2]  ->  A[n1][k2] stored at v[n1 + 4 k2]
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) dft4(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  // twiddle A[n1][k2] *= W16^(n1*k2)
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  // k2 = 1: W^1, W^2, W^3
  v[1 + 4] = cmulf(v[1 + 4], cx(c1, -s1));
  v[2 + 4] = cx(h * (v[2 + 4].x + v[2 + 4].y), h * (v[2 + 4].y - v[2 + 4].x));  // * (h, -h)
  v[3 + 4] = cmulf(v[3 + 4], cx(s1, -c1));
  // k2 = 2: W^2, W^4, W^6
  v[1 + 8] = cx(h * (v[1 + 8].x + v[1 + 8].y), h * (v[1 + 8].y - v[1 + 8].x));
  v[2 + 8] = mul_neg_i(v[2 + 8]);
  v[3 + 8] = cx(h * (v[3 + 8].y - v[3 + 8].x), -h * (v[3 + 8].x + v[3 + 8].y));  // * (-h, -h)
  // k2 = 3: W^3, W^6, W^9
  v[1 + 12] = cmulf(v[1 + 12], cx(s1, -c1));
  v[2 + 12] = cx(h * (v[2 + 12].y - v[2 + 12].x), -h * (v[2 + 12].x + v[2 + 12].y));
  v[3 + 12] = cmulf(v[3 + 12], cx(-c1, s1));
  // layer 2: for each k2, DFT4 over n1  ->  X[4 k1 + k2] at v[k1 + 4 k2]
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) dft4(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
}

// forward 8-point DFT of v[0..7] (natural order in, natural order out): radix-2 DIF split + two DFT4
MAFE_HD void dft8(cpx* v) {
  const float h = 0.70710678118654752440f;
  cpx a0 = v[0] + v[4], a1 = v[1] + v[5], a2 = v[2] + v[6], a3 = v[3] + v[7];
  cpx b0 = v[0] - v[4], b1 = v[1] - v[5], b2 = v[2] - v[6], b3 = v[3] - v[7];
  b1 = cx(h * (b1.x + b1.y), h * (b1.y - b1.x));     // * W8^1 = (h, -h)
  b2 = mul_neg_i(b2);                                // * W8^2 = -i
  b3 = cx(h * (b3.y - b3.x), -h * (b3.x + b3.y));    // * W8^3 = (-h, -h)
  dft4(a0, a1, a2, a3);                              // X[0], X[2], X[4], X[6]
  dft4(b0, b1, b2, b3);                              // X[1], X[3], X[5], X[7]
  v[0] = a0; v[2] = a1; v[4] = a2; v[6] = a3;
  v[1] = b0; v[3] = b1; v[5] = b2; v[7] = b3;
}

// shared-memory transpose slot of one 256-point transform: row stride 17 complex (conflict-free
// for 8-byte accesses by 16 lanes), 16 rows -> 272 complex; slots are spaced kSlotStride apart so
// that the same bin of 16 different slots falls into 16 different bank pairs.
constexpr int kRowStride = 17;
constexpr int kSlotStride = 273;

}  // namespace mafe
