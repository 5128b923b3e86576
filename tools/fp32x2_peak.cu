#include <cstdio>
#include <cstdint>
typedef unsigned long long c2;
__device__ __forceinline__ c2 fma2(c2 a, c2 b, c2 c){ c2 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ c2 add2(c2 a, c2 b){ c2 r; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ float add1(float a, float b){ float r; asm volatile("add.rn.f32 %0, %1, %2;":"=f"(r):"f"(a),"f"(b)); return r;}
// MODE 0: scalar FFMA x8 chains; 1: FFMA2 x8 chains; 2: FADD2 x8; 3: mix 4 FFMA2 + 4 FFMA; 4: FADD scalar; 5: mix 4 FADD2 + 8 FADD
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
  float a[8]; c2 p[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = ((c2)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] + 1.f); }
  c2 ps = ((c2)__float_as_uint(s) << 32) | __float_as_uint(s);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0) { for (int i = 0; i < 8; ++i) a[i] = fma1(a[i], s, a[i]); }
      if (MODE == 1) { for (int i = 0; i < 8; ++i) p[i] = fma2(p[i], ps, p[i]); }
      if (MODE == 2) { for (int i = 0; i < 8; ++i) p[i] = add2(p[i], ps); }
      if (MODE == 3) { for (int i = 0; i < 4; ++i) { p[i] = fma2(p[i], ps, p[i]); a[i] = fma1(a[i], s, a[i]); } }
      if (MODE == 4) { for (int i = 0; i < 8; ++i) a[i] = add1(a[i], s); }
      if (MODE == 5) { for (int i = 0; i < 4; ++i) { p[i] = add2(p[i], ps); a[2*i] = add1(a[2*i], s); a[2*i+1] = add1(a[2*i+1], s); } }
    }
  }
  float acc = 0; for (int i = 0; i < 8; ++i) acc += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32));
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE> void run(const char* name, double lane_ops_per_iter) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4000;
  k<MODE><<<148 * 8, 256>>>(out, 100, 1.0001f);
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(out, iters, 1.0001f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = (double)148 * 8 * 256 * iters * lane_ops_per_iter;
  printf("%-28s %.3f ms  %.1f T lane-ops/s  (instr issue %.2f per clk per SMSP at 1.965 GHz)\n", name, ms, ops / ms / 1e9, 0.0);
  cudaFree(out);
}
int main() {
  run<0>("FFMA  x64/iter", 64);          // lane-ops = fma count per thread per iter
  run<1>("FFMA2 x64/iter", 128);
  run<2>("FADD2 x64/iter", 128);
  run<3>("FFMA2 x32 + FFMA x32", 96);
  run<4>("FADD  x64/iter", 64);
  run<5>("FADD2 x32 + FADD x64", 128);
  return 0;
}
