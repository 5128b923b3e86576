// fft_generic.cuh -- shared-memory Stockham autosort FFT, runtime mixed radix (2/3/4/5 + prime
// fallback), forward transform, `pairs` independent length-N complex sequences per CTA.
#pragma once
#include "common.cuh"

namespace mafe {

constexpr int kMaxStages = 16;

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward DFT quarter turn)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

struct FftStages {
  int radices[kMaxStages];
  int n_stages;
};

// Forward DFT of `pairs` sequences held in cur[p*N + n]; returns the buffer holding the result
// (cur or nxt).  All threads of the CTA must call it; starts and ends with data visible to all.
__device__ __forceinline__ float2* stockham_fft(float2* cur, float2* nxt, int pairs, int N, const FftStages& S,
                                               const float2* __restrict__ tw) {
  int Ns = 1;
  for (int s = 0; s < S.n_stages; ++s) {
    const int R = S.radices[s];
    const int M = N / R;               // butterflies per FFT
    const int tstep = N / (Ns * R);    // twiddle index step: W_{Ns*R}^{k} = W_N^{k*tstep}
    if (R <= 5) {
      for (int idx = threadIdx.x; idx < pairs * M; idx += blockDim.x) {
        int p = idx / M, j = idx - p * M;
        int k = j % Ns;
        const float2* in = cur + (size_t)p * N;
        float2* outp = nxt + (size_t)p * N;
        int j0 = (j - k) * R + k;
        float2 v[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          if (r < R) {
            float2 x = in[j + r * M];
            if (r > 0 && k > 0) x = cmul(x, __ldg(&tw[(k * r * tstep) % N]));
            v[r] = x;
          }
        }
        if (R == 2) {
          outp[j0] = cadd(v[0], v[1]);
          outp[j0 + Ns] = csub(v[0], v[1]);
        } else if (R == 4) {
          float2 s0 = cadd(v[0], v[2]), d0 = csub(v[0], v[2]);
          float2 s1 = cadd(v[1], v[3]), d1 = mul_mi(csub(v[1], v[3]));
          outp[j0] = cadd(s0, s1);
          outp[j0 + Ns] = cadd(d0, d1);
          outp[j0 + 2 * Ns] = csub(s0, s1);
          outp[j0 + 3 * Ns] = csub(d0, d1);
        } else if (R == 3) {
          const float c = -0.5f, sn = -0.86602540378443864676f;  // W_3 = c + i*sn
          float2 t1 = cadd(v[1], v[2]);
          float2 t2 = make_float2(fmaf(c, t1.x, v[0].x), fmaf(c, t1.y, v[0].y));
          float2 d = csub(v[1], v[2]);
          float2 t3 = make_float2(-sn * d.y, sn * d.x);  // i*sn*d
          outp[j0] = cadd(v[0], t1);
          outp[j0 + Ns] = cadd(t2, t3);
          outp[j0 + 2 * Ns] = csub(t2, t3);
        } else {  // R == 5
          const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
          const float s1 = -0.95105651629515357212f, s2 = -0.58778525229247312917f;  // sin(-2pi/5), sin(-4pi/5)
          float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
          float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
          float2 m1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
          float2 m2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
          // i*(s1*b1 + s2*b2) and i*(s2*b1 - s1*b2)
          float2 n1 = make_float2(-(s1 * b1.y + s2 * b2.y), s1 * b1.x + s2 * b2.x);
          float2 n2 = make_float2(-(s2 * b1.y - s1 * b2.y), s2 * b1.x - s1 * b2.x);
          outp[j0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
          outp[j0 + Ns] = cadd(m1, n1);
          outp[j0 + 2 * Ns] = cadd(m2, n2);
          outp[j0 + 3 * Ns] = csub(m2, n2);
          outp[j0 + 4 * Ns] = csub(m1, n1);
        }
      }
    } else {
      // prime radix fallback: one thread per output, O(R) inputs each
      const int rstep = N / R;  // W_R^{1} = W_N^{rstep}
      for (int idx = threadIdx.x; idx < pairs * N; idx += blockDim.x) {
        int p = idx / N, o = idx - p * N;
        int j = o / R, t = o - j * R;  // butterfly j, output t
        int k = j % Ns;
        const float2* in = cur + (size_t)p * N;
        float2 acc = make_float2(0.f, 0.f);
        for (int r = 0; r < R; ++r) {
          long long e = ((long long)k * r * tstep + (long long)r * t * rstep) % N;
          acc = cadd(acc, cmul(in[j + r * M], __ldg(&tw[(int)e])));
        }
        nxt[(size_t)p * N + (j - k) * R + k + t * Ns] = acc;
      }
    }
    __syncthreads();
    float2* tmp = cur; cur = nxt; nxt = tmp;
    Ns *= R;
  }

  return cur;
}

}  // namespace mafe
