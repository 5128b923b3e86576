// packed.cuh -- complex arithmetic on Blackwell's packed FP32 pipe (sm_100: add / mul / fma .f32x2).
//
// A complex value rides in ONE 64-bit register pair (re = low half, im = high half).  SASS: FADD2 / FMUL2 / FFMA2 take
// a 64-bit pair per operand with free per-operand modifiers -- half swap (.LO_HI), per-half sign (.NP / .PN), scalar
// broadcast (R.F32) -- which ptxas folds out of the `mov.b64` pack / unpack idioms below.  So
//   complex add / sub            = 1 instruction  (2 scalar)
//   multiply by -i / +i / conj   = 0 instructions (operand modifier of the consumer)
//   complex * twiddle            = 2 instructions (FMUL2 + FFMA2 with the swapped, sign-patterned operand; 4 scalar)
// The front-end kernels are bound by issue slots, not by the FP32 pipe (profiles/r01_fbank512_v3_ncu_details.txt: issue 76 %,
// FMA pipe 37 %), so halving the FP instruction count of the butterflies is what the packed forms buy.
#pragma once
#include <cuda_runtime.h>

namespace mafe {

typedef unsigned long long c2;   // (re, im) float pair

__device__ __forceinline__ c2 pk(float x, float y) {
  c2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float re(c2 v) {
  float a;
  asm("{ .reg .f32 t; mov.b64 {%0, t}, %1; }" : "=f"(a) : "l"(v));
  return a;
}
__device__ __forceinline__ float im(c2 v) {
  float b;
  asm("{ .reg .f32 t; mov.b64 {t, %0}, %1; }" : "=f"(b) : "l"(v));
  return b;
}
__device__ __forceinline__ c2 add2(c2 a, c2 b) {
  c2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c2 sub2(c2 a, c2 b) {
  c2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c2 mul2(c2 a, c2 b) {
  c2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ c2 fma2(c2 a, c2 b, c2 c) {
  c2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ c2 bc(float x) { return pk(x, x); }                    // (x, x)
__device__ __forceinline__ c2 swp(c2 a) { return pk(im(a), re(a)); }              // (im, re)
__device__ __forceinline__ c2 mni(c2 a) { return pk(im(a), -re(a)); }             // a * (-i)
__device__ __forceinline__ c2 cnj(c2 a) { return pk(re(a), -im(a)); }             // conj(a)
// a * (c + i s)
__device__ __forceinline__ c2 cmul(c2 a, float c, float s) { return fma2(swp(a), pk(-s, s), mul2(a, bc(c))); }

// forward 4-point DFT in place
__device__ __forceinline__ void dft4p(c2& a0, c2& a1, c2& a2, c2& a3) {
  const c2 s02 = add2(a0, a2), d02 = sub2(a0, a2), s13 = add2(a1, a3), d13 = mni(sub2(a1, a3));
  a0 = add2(s02, s13);
  a1 = add2(d02, d13);
  a2 = sub2(s02, s13);
  a3 = sub2(d02, d13);
}

// forward 16-point DFT of v[0..15] (natural order in); output bin k at v[fft16_pos(k)] (same map as fft512.cuh::fft16)
__device__ __forceinline__ void fft16p(c2* v) {
#pragma unroll
  for (int n1 = 0; n1 < 4; ++n1) dft4p(v[n1], v[n1 + 4], v[n1 + 8], v[n1 + 12]);
  const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  v[5] = cmul(v[5], c1, -s1);                                // W16^1
  v[6] = mul2(add2(v[6], mni(v[6])), bc(h));                 // W16^2 = (h, -h): h (x + y, y - x)
  v[7] = cmul(v[7], s1, -c1);                                // W16^3
  v[9] = mul2(add2(v[9], mni(v[9])), bc(h));                 // W16^2
  v[10] = mni(v[10]);                                        // W16^4 = -i
  v[11] = mul2(sub2(mni(v[11]), v[11]), bc(h));              // W16^6 = (-h, -h): h (y - x, -x - y)
  v[13] = cmul(v[13], s1, -c1);                              // W16^3
  v[14] = mul2(sub2(mni(v[14]), v[14]), bc(h));              // W16^6
  v[15] = cmul(v[15], -c1, s1);                              // W16^9
#pragma unroll
  for (int k2 = 0; k2 < 4; ++k2) dft4p(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);
}

__device__ __forceinline__ c2 pli(c2 a) { return pk(-im(a), re(a)); }             // a * (+i)

// forward 5-point DFT in place (same operation order as fft400.cuh::dft5)
__device__ __forceinline__ void dft5p(c2& x0, c2& x1, c2& x2, c2& x3, c2& x4) {
  const float c1 = 0.30901699437494742410f, c2_ = -0.80901699437494742410f;
  const float s1 = -0.95105651629515357212f, s2 = -0.58778525229247312917f;  // -sin(2pi/5), -sin(4pi/5)
  const c2 a1 = add2(x1, x4), b1 = sub2(x1, x4), a2 = add2(x2, x3), b2 = sub2(x2, x3);
  const c2 m1 = fma2(bc(c2_), a2, fma2(bc(c1), a1, x0));
  const c2 m2 = fma2(bc(c1), a2, fma2(bc(c2_), a1, x0));
  const c2 t1 = fma2(bc(s2), b2, mul2(bc(s1), b1));     // n1 = i * (s1 b1 + s2 b2)
  const c2 t2 = fma2(bc(-s1), b2, mul2(bc(s2), b1));    // n2 = i * (s2 b1 - s1 b2)
  x0 = add2(add2(x0, a1), a2);
  x1 = add2(m1, pli(t1));
  x4 = sub2(m1, pli(t1));
  x2 = add2(m2, pli(t2));
  x3 = sub2(m2, pli(t2));
}

// forward 25-point DFT (5 x 5); index maps and twiddle table of fft400.cuh::fft25 (output bin k at v[fft25_pos(k)])
__device__ __forceinline__ void fft25p(c2* v, const float2* tw25) {
#pragma unroll
  for (int j1 = 0; j1 < 5; ++j1) dft5p(v[j1], v[j1 + 5], v[j1 + 10], v[j1 + 15], v[j1 + 20]);
#pragma unroll
  for (int j1 = 1; j1 < 5; ++j1)
#pragma unroll
    for (int k1 = 1; k1 < 5; ++k1) {
      const float2 w = tw25[(j1 - 1) * 4 + (k1 - 1)];
      v[j1 + 5 * k1] = cmul(v[j1 + 5 * k1], w.x, w.y);
    }
#pragma unroll
  for (int k1 = 0; k1 < 5; ++k1) dft5p(v[5 * k1], v[5 * k1 + 1], v[5 * k1 + 2], v[5 * k1 + 3], v[5 * k1 + 4]);
}

// forward 20-point DFT (4 x 5); fft400.cuh::fft20 (output bin k at v[fft20_pos(k)])
__device__ __forceinline__ void fft20p(c2* v, const float2* tw20) {
#pragma unroll
  for (int j1 = 0; j1 < 5; ++j1) dft4p(v[j1], v[j1 + 5], v[j1 + 10], v[j1 + 15]);
#pragma unroll
  for (int j1 = 1; j1 < 5; ++j1)
#pragma unroll
    for (int k1 = 1; k1 < 4; ++k1) {
      const float2 w = tw20[(j1 - 1) * 3 + (k1 - 1)];
      v[j1 + 5 * k1] = cmul(v[j1 + 5 * k1], w.x, w.y);
    }
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft5p(v[5 * k1], v[5 * k1 + 1], v[5 * k1 + 2], v[5 * k1 + 3], v[5 * k1 + 4]);
}

__device__ __forceinline__ c2 lds_c2(const float2* p) { return *reinterpret_cast<const c2*>(p); }
__device__ __forceinline__ void sts_c2(float2* p, c2 v) { *reinterpret_cast<c2*>(p) = v; }

}  // namespace mafe
