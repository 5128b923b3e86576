// fft_generic.cuh -- shared-memory Stockham autosort FFT, runtime mixed radix (2/3/4/5 + prime
// fallback), forward transform, `pairs` independent length-N complex sequences per CTA.
#pragma once
#include "common.cuh"
#include "fft512.cuh"

namespace mafe {

constexpr int kMaxStages = 16;

template <typename C2> struct Real;
template <> struct Real<float2> { typedef float type; };
template <> struct Real<double2> { typedef double type; };
template <typename C2, typename R> __device__ __forceinline__ C2 mk(R x, R y) { C2 r; r.x = x; r.y = y; return r; }

template <typename C2> __device__ __forceinline__ C2 cmul(C2 a, C2 b) {
  return mk<C2>(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
template <typename C2> __device__ __forceinline__ C2 cadd(C2 a, C2 b) { return mk<C2>(a.x + b.x, a.y + b.y); }
template <typename C2> __device__ __forceinline__ C2 csub(C2 a, C2 b) { return mk<C2>(a.x - b.x, a.y - b.y); }
// multiply by -i (forward DFT quarter turn)
template <typename C2> __device__ __forceinline__ C2 mul_mi(C2 a) { return mk<C2>(a.y, -a.x); }

struct FftStages {
  int radices[kMaxStages];
  int n_stages;
};

// padded position inside a sequence: one spare element per 16 (radix-16 / 8 passes write 16 consecutive outputs
// per thread; without the skew every lane of a warp would hit the same bank pair)
__device__ __forceinline__ int pad16(int i) { return i + (i >> 4); }

// Forward DFT of `pairs` sequences held in cur[p*stride + n]; returns the buffer holding the result
// (cur or nxt).  All threads of the CTA must call it; starts and ends with data visible to all.
// Radices 16 / 8 (float, power-of-two N only; chosen by the host for N >= 256) run as register-resident fft16 /
// dft8 butterflies: 3 passes for N = 2048 instead of 6 radix-4/2 passes.  Their intermediate buffers are skewed by
// pad16(); the first input and the final output are in natural order, so callers are not affected (`stride` must
// leave room for N + N/16 elements per sequence).
template <typename C2>
__device__ __forceinline__ C2* stockham_fft(C2* cur, C2* nxt, int pairs, int N, const FftStages& S,
                                            const C2* __restrict__ tw, int stride = 0) {
  typedef typename Real<C2>::type Rt;
  if (stride == 0) stride = N;
  const bool skewed = S.radices[0] >= 8;   // the host uses 16 / 8 for every pass or for none... except a final 4 / 2
  int Ns = 1;
  for (int s = 0; s < S.n_stages; ++s) {
    const int R = S.radices[s];
    const int M = N / R;               // butterflies per FFT
    const int tstep = N / (Ns * R);    // twiddle index step: W_{Ns*R}^{k} = W_N^{k*tstep}
    const bool in_skew = skewed && s > 0, out_skew = skewed && s + 1 < S.n_stages;
    if (sizeof(Rt) == 4 && (R == 16 || R == 8)) {
      const int lgM = 31 - __clz(M);
      for (int idx = threadIdx.x; idx < pairs * M; idx += blockDim.x) {
        const int p = idx >> lgM, j = idx & (M - 1), k = j & (Ns - 1);
        const float2* in = reinterpret_cast<const float2*>(cur) + (size_t)p * stride;
        float2* outp = reinterpret_cast<float2*>(nxt) + (size_t)p * stride;
        const int j0 = (j - k) * R + k;
        cpx v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          if (r < R) {
            const int i = j + r * M;
            const float2 x = in[in_skew ? pad16(i) : i];
            v[r] = cx(x.x, x.y);
            if (r > 0 && k > 0) {
              const float2 w = reinterpret_cast<const float2*>(tw)[(k * r * tstep) & (N - 1)];
              v[r] = cmulf(v[r], cx(w.x, w.y));
            }
          }
        }
        if (R == 16) {
          fft16(v);
#pragma unroll
          for (int t = 0; t < 16; ++t) {
            const int o = j0 + t * Ns;
            outp[out_skew ? pad16(o) : o] = make_float2(v[fft16_pos(t)].x, v[fft16_pos(t)].y);
          }
        } else {
          dft8(v);
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const int o = j0 + t * Ns;
            outp[out_skew ? pad16(o) : o] = make_float2(v[t].x, v[t].y);
          }
        }
      }
    } else if (R <= 5) {
      // power-of-two sizes (N, hence M and Ns for radices 4 / 2): shifts and masks instead of integer division --
      // the division-heavy index arithmetic was 60 % of this loop's instructions (ncu, n_fft = 2048)
      const bool pow2 = (N & (N - 1)) == 0 && (R == 2 || R == 4);
      const int lgM = 31 - __clz(M);
      for (int idx = threadIdx.x; idx < pairs * M; idx += blockDim.x) {
        int p, j, k;
        if (pow2) { p = idx >> lgM; j = idx & (M - 1); k = j & (Ns - 1); }
        else { p = idx / M; j = idx - p * M; k = j % Ns; }
        const C2* in = cur + (size_t)p * stride;
        C2* outp = nxt + (size_t)p * stride;
        int j0 = (j - k) * R + k;
        C2 v[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          if (r < R) {
            C2 x = in[in_skew ? pad16(j + r * M) : j + r * M];
            if (r > 0 && k > 0) x = cmul(x, tw[pow2 ? ((k * r * tstep) & (N - 1)) : ((k * r * tstep) % N)]);
            v[r] = x;
          }
        }
        if (R == 2) {
          outp[j0] = cadd(v[0], v[1]);
          outp[j0 + Ns] = csub(v[0], v[1]);
        } else if (R == 4) {
          C2 s0 = cadd(v[0], v[2]), d0 = csub(v[0], v[2]);
          C2 s1 = cadd(v[1], v[3]), d1 = mul_mi(csub(v[1], v[3]));
          outp[j0] = cadd(s0, s1);
          outp[j0 + Ns] = cadd(d0, d1);
          outp[j0 + 2 * Ns] = csub(s0, s1);
          outp[j0 + 3 * Ns] = csub(d0, d1);
        } else if (R == 3) {
          const Rt c = (Rt)-0.5, sn = (Rt)-0.86602540378443864676;  // W_3 = c + i*sn
          C2 t1 = cadd(v[1], v[2]);
          C2 t2 = mk<C2>(fma(c, t1.x, v[0].x), fma(c, t1.y, v[0].y));
          C2 d = csub(v[1], v[2]);
          C2 t3 = mk<C2>(-sn * d.y, sn * d.x);  // i*sn*d
          outp[j0] = cadd(v[0], t1);
          outp[j0 + Ns] = cadd(t2, t3);
          outp[j0 + 2 * Ns] = csub(t2, t3);
        } else {  // R == 5
          const Rt c1 = (Rt)0.30901699437494742410, c2 = (Rt)-0.80901699437494742410;
          const Rt s1 = (Rt)-0.95105651629515357212, s2 = (Rt)-0.58778525229247312917;  // sin(-2pi/5), sin(-4pi/5)
          C2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
          C2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
          C2 m1 = mk<C2>(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
          C2 m2 = mk<C2>(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
          // i*(s1*b1 + s2*b2) and i*(s2*b1 - s1*b2)
          C2 n1 = mk<C2>(-(s1 * b1.y + s2 * b2.y), s1 * b1.x + s2 * b2.x);
          C2 n2 = mk<C2>(-(s2 * b1.y - s1 * b2.y), s2 * b1.x - s1 * b2.x);
          outp[j0] = mk<C2>(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
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
        const C2* in = cur + (size_t)p * N;
        C2 acc = mk<C2>((Rt)0, (Rt)0);
        for (int r = 0; r < R; ++r) {
          long long e = ((long long)k * r * tstep + (long long)r * t * rstep) % N;
          acc = cadd(acc, cmul(in[j + r * M], tw[(int)e]));
        }
        nxt[(size_t)p * N + (j - k) * R + k + t * Ns] = acc;
      }
    }
    __syncthreads();
    C2* tmp = cur; cur = nxt; nxt = tmp;
    Ns *= R;
  }

  return cur;
}

}  // namespace mafe
