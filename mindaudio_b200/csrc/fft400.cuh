// fft400.cuh -- register-resident building blocks of the 400-point complex FFT (400 = 25 x 16) used by the
// specialised n_fft = 400 front-end kernel (fbank400.cuh): features.fbank / mfcc / melspectrogram defaults
// (mindaudio/data/features.py:196-373, spectrum.py:609-698) and the ECAPA-TDNN front-end.
//
// Index maps (forward DFT, z[n] = a[n] + i b[n] carries a frame PAIR):
//   n = t + 16 j  (t = 0..15 lanes of the group, j = 0..24 in registers),   k = kj + 25 kt
//   stage 1 (in thread):  Y[t][kj] = sum_j z[t + 16 j] W25^(j kj)          25-point DFT = 5 x 5
//   twiddle:              Y[t][kj] *= W400^(t kj)
//   stage 2 (16 points):  X[kj + 25 kt] = sum_t Y[t][kj] W16^(t kt)        fft16() of fft512.cuh
// __host__ __device__ so tests/host/fft400_host_check.cu exercises the maps on the CPU.
#pragma once
#include "fft512.cuh"

namespace mafe {

// forward 5-point DFT in place
MAFE_HD void dft5(cpx& x0, cpx& x1, cpx& x2, cpx& x3, cpx& x4) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
  const float s1 = -0.95105651629515357212f, s2 = -0.58778525229247312917f;  // -sin(2pi/5), -sin(4pi/5)
  const cpx a1 = x1 + x4, b1 = x1 - x4, a2 = x2 + x3, b2 = x2 - x3;
  const cpx m1 = cx(fmaf(c2, a2.x, fmaf(c1, a1.x, x0.x)), fmaf(c2, a2.y, fmaf(c1, a1.y, x0.y)));
  const cpx m2 = cx(fmaf(c1, a2.x, fmaf(c2, a1.x, x0.x)), fmaf(c1, a2.y, fmaf(c2, a1.y, x0.y)));
  // i * (s1 b1 + s2 b2)  and  i * (s2 b1 - s1 b2)
  const cpx n1 = cx(-fmaf(s2, b2.y, s1 * b1.y), fmaf(s2, b2.x, s1 * b1.x));
  const cpx n2 = cx(-fmaf(-s1, b2.y, s2 * b1.y), fmaf(-s1, b2.x, s2 * b1.x));
  x0 = cx(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
  x1 = m1 + n1;
  x4 = m1 - n1;
  x2 = m2 + n2;
  x3 = m2 - n2;
}

// position of output bin k (0..24) inside v[] after fft25(): X[k1 + 5 k2] lives at v[5 k1 + k2]
MAFE_HD constexpr int fft25_pos(int k) { return 5 * (k % 5) + k / 5; }

// forward 25-point DFT of v[0..24] (natural order in: v[j], j = 5 j2 + j1); output bin k at v[fft25_pos(k)].
// tw25[(j1-1)*4 + (k1-1)] = W25^(j1 k1) for j1, k1 = 1..4 (16 entries, caller-provided: constant/shared memory).
template <typename TW>
MAFE_HD void fft25(cpx* v, const TW* tw25) {
#pragma unroll
  for (int j1 = 0; j1 < 5; ++j1) dft5(v[j1], v[j1 + 5], v[j1 + 10], v[j1 + 15], v[j1 + 20]);   // over j2 -> A[j1][k1] at v[j1 + 5 k1]
#pragma unroll
  for (int j1 = 1; j1 < 5; ++j1)
#pragma unroll
    for (int k1 = 1; k1 < 5; ++k1) {
      const TW w = tw25[(j1 - 1) * 4 + (k1 - 1)];
      v[j1 + 5 * k1] = cmulf(v[j1 + 5 * k1], cx(w.x, w.y));
    }
#pragma unroll
  for (int k1 = 0; k1 < 5; ++k1) dft5(v[5 * k1], v[5 * k1 + 1], v[5 * k1 + 2], v[5 * k1 + 3], v[5 * k1 + 4]);  // over j1 -> X[k1 + 5 k2] at v[5 k1 + k2]
}

// ---- 20-point DFT (20 = 4 x 5) for the 320-point transform (320 = 20 x 16; deepspeech2's stft) ----
// position of output bin k (0..19) inside v[] after fft20(): X[k1 + 4 k2] lives at v[5 k1 + k2]
MAFE_HD constexpr int fft20_pos(int k) { return 5 * (k % 4) + k / 4; }

// forward 20-point DFT of v[0..19] (natural order in: v[j], j = 5 j2 + j1, j1 = 0..4, j2 = 0..3); output bin k at
// v[fft20_pos(k)].  tw20[(j1-1)*3 + (k1-1)] = W20^(j1 k1) for j1 = 1..4, k1 = 1..3 (12 entries).
template <typename TW>
MAFE_HD void fft20(cpx* v, const TW* tw20) {
#pragma unroll
  for (int j1 = 0; j1 < 5; ++j1) dft4(v[j1], v[j1 + 5], v[j1 + 10], v[j1 + 15]);   // over j2 -> A[j1][k1] at v[j1 + 5 k1]
#pragma unroll
  for (int j1 = 1; j1 < 5; ++j1)
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1) {
      const TW w = tw20[(j1 - 1) * 3 + (k1 - 1)];
      v[j1 + 5 * k1] = cmulf(v[j1 + 5 * k1], cx(w.x, w.y));
    }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft5(v[5 * k1], v[5 * k1 + 1], v[5 * k1 + 2], v[5 * k1 + 3], v[5 * k1 + 4]);  // over j1 -> X[k1 + 4 k2] at v[5 k1 + k2]
}

}  // namespace mafe
