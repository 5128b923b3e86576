// istft.cu -- inverse STFT (mindaudio/data/spectrum.py:346-474): windowed inverse real DFT of every
// frame (two frames packed per complex Stockham FFT, inverse via conj-FFT-conj), then a gather-form
// overlap-add with the window-sum-square normalisation (spectrum.py:339-343, 477-494).
// Arithmetic is FLOAT64 like the reference (its output dtype is float64 and the window-tail
// division y /= wss amplifies FP32 rounding by 1/w^2); the float32 result is the rounded answer.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "fft_generic.cuh"

namespace mafe {

constexpr int kIstftThreads = 256;

// spec: [U][T][F] complex64 (frame-major).  seg: [U][T][N] windowed time-domain frames.
__global__ void __launch_bounds__(kIstftThreads) istft_frames_kernel(const float2* __restrict__ spec, int T, int N, int F,
                                                                     int pairs, int tiles_per_utt, FftStages S,
                                                                     const double2* __restrict__ tw,
                                                                     const double* __restrict__ window,
                                                                     double* __restrict__ seg) {
  extern __shared__ double2 smem[];
  double2* cur = smem;
  double2* nxt = smem + (size_t)pairs * N;
  const int u = blockIdx.x / tiles_per_utt;
  const int frame0 = (blockIdx.x - u * tiles_per_utt) * 2 * pairs;
  const float2* su = spec + (size_t)u * T * F;
  const bool even = (N & 1) == 0;
  for (int idx = threadIdx.x; idx < pairs * N; idx += blockDim.x) {
    int p = idx / N, n = idx - p * N;
    int fa = frame0 + 2 * p, fb = fa + 1;
    int k = n < F ? n : N - n;
    float sgn = n < F ? 1.f : -1.f;  // Hermitian extension: X[N-k] = conj(X[k])
    float2 xa = fa < T ? su[(size_t)fa * F + k] : make_float2(0.f, 0.f);
    float2 xb = fb < T ? su[(size_t)fb * F + k] : make_float2(0.f, 0.f);
    xa.y *= sgn; xb.y *= sgn;
    if (n == 0 || (even && n == N / 2)) { xa.y = 0.f; xb.y = 0.f; }  // c2r ignores these imaginary parts
    // conj(xa + i*xb) = (ar - bi) + i*(-ai - br)
    cur[idx] = make_double2((double)xa.x - (double)xb.y, -(double)xa.y - (double)xb.x);
  }
  __syncthreads();
  double2* res = stockham_fft(cur, nxt, pairs, N, S, tw);
  const double invn = 1.0 / (double)N;
  double* segu = seg + (size_t)u * T * N;
  for (int idx = threadIdx.x; idx < pairs * N; idx += blockDim.x) {
    int p = idx / N, n = idx - p * N;
    int fa = frame0 + 2 * p, fb = fa + 1;
    double2 w = res[idx];
    double wn = window[n] * invn;
    if (fa < T) segu[(size_t)fa * N + n] = w.x * wn;
    if (fb < T) segu[(size_t)fb * N + n] = -w.y * wn;
  }
}

// y[u][s] = sum_t seg[u][t][s - t*hop] / wss[s]   where wss[s] = sum_t w^2[s - t*hop] (> 1e-9)
__global__ void overlap_add_kernel(const double* __restrict__ seg, int U, int T, int N, int hop, const double* __restrict__ window,
                                   double* __restrict__ y) {
  const int64_t n_out = (int64_t)N + (int64_t)hop * (T - 1);
  const int64_t total = (int64_t)U * n_out;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t u = i / n_out, s = i - u * n_out;
    int64_t t_hi = min((int64_t)T - 1, s / hop);
    int64_t t_lo = s - N + 1 <= 0 ? 0 : (s - N + 1 + hop - 1) / hop;
    double acc = 0.0, wss = 0.0;
    const double* su = seg + (size_t)u * T * N;
    for (int64_t t = t_lo; t <= t_hi; ++t) {
      int n = (int)(s - t * hop);
      double w = window[n];
      acc += su[t * N + n];
      wss = fma(w, w, wss);
    }
    y[i] = wss > 1e-9 ? acc / wss : acc;
  }
}

}  // namespace mafe

using namespace mafe;

static void factorize_local(int n, std::vector<int>& out) {
  out.clear();
  while (n % 4 == 0) { out.push_back(4); n /= 4; }
  while (n % 2 == 0) { out.push_back(2); n /= 2; }
  while (n % 3 == 0) { out.push_back(3); n /= 3; }
  while (n % 5 == 0) { out.push_back(5); n /= 5; }
  for (int f = 7; (long long)f * f <= n; f += 2)
    while (n % f == 0) { out.push_back(f); n /= f; }
  if (n > 1) out.push_back(n);
}

extern "C" int mafe_istft(mafe_ctx* ctx, const float* spec, int32_t U, int32_t T, int32_t N, int32_t hop,
                          const double* window_host, double* y) {
  MAFE_REQUIRE(ctx && window_host, "mafe_istft: NULL argument");
  MAFE_REQUIRE(N >= 2 && N <= 4096, "istft n_fft=%d unsupported (2..4096)", N);
  MAFE_REQUIRE(hop >= 1, "Invalid hop_length: %d", hop);
  if (U <= 0 || T <= 0) return MAFE_OK;
  MAFE_REQUIRE(spec && y, "mafe_istft: NULL buffer");
  cudaSetDevice(ctx->device);
  const int F = N / 2 + 1;
  std::vector<int> rad;
  factorize_local(N, rad);
  MAFE_REQUIRE((int)rad.size() <= kMaxStages, "n_fft=%d has too many factors", N);
  FftStages S;
  for (int i = 0; i < kMaxStages; ++i) S.radices[i] = i < (int)rad.size() ? rad[i] : 1;
  S.n_stages = (int)rad.size();
  std::vector<double2> tw(N);
  for (int k = 0; k < N; ++k) {
    double a = -2.0 * M_PI * (double)k / (double)N;
    tw[k] = make_double2(cos(a), sin(a));
  }
  double2* tw_dev = nullptr;
  double* win_dev = nullptr;
  double* seg = nullptr;
  cudaStream_t st = ctx->stream;
  MAFE_CUDA_CHECK(cudaMallocAsync((void**)&tw_dev, sizeof(double2) * N, st));
  MAFE_CUDA_CHECK(cudaMallocAsync((void**)&win_dev, sizeof(double) * N, st));
  MAFE_CUDA_CHECK(cudaMallocAsync((void**)&seg, sizeof(double) * (size_t)U * T * N, st));
  MAFE_CUDA_CHECK(cudaMemcpyAsync(tw_dev, tw.data(), sizeof(double2) * N, cudaMemcpyHostToDevice, st));
  MAFE_CUDA_CHECK(cudaMemcpyAsync(win_dev, window_host, sizeof(double) * N, cudaMemcpyHostToDevice, st));
  size_t per_pair = (size_t)N * sizeof(double2) * 2;
  int pairs = (int)std::min<size_t>(8, std::max<size_t>(1, (64 * 1024) / per_pair));
  size_t smem = per_pair * pairs;
  if (smem > 48 * 1024)
    MAFE_CUDA_CHECK(cudaFuncSetAttribute(istft_frames_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int tiles_per_utt = (T + 2 * pairs - 1) / (2 * pairs);
  istft_frames_kernel<<<U * tiles_per_utt, kIstftThreads, smem, st>>>((const float2*)spec, T, N, F, pairs, tiles_per_utt, S,
                                                                       tw_dev, win_dev, seg);
  MAFE_LAUNCH_CHECK(ctx);
  int64_t total = (int64_t)U * ((int64_t)N + (int64_t)hop * (T - 1));
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((total + 255) / 256, (int64_t)ctx->sm_count * 32));
  overlap_add_kernel<<<grid, 256, 0, st>>>(seg, U, T, N, hop, win_dev, y);
  MAFE_LAUNCH_CHECK(ctx);
  // the pageable twiddle / window host vectors must outlive the async copies
  MAFE_CUDA_CHECK(cudaStreamSynchronize(st));
  MAFE_CUDA_CHECK(cudaFreeAsync(seg, st));
  MAFE_CUDA_CHECK(cudaFreeAsync(win_dev, st));
  MAFE_CUDA_CHECK(cudaFreeAsync(tw_dev, st));
  return MAFE_OK;
}
