// resample.cu -- Fourier-method resampling of whole signals on the device ("next" row f2: the `resample` step of
// augment.pitch_shift, mindaudio/data/augment.py:874-901 -> processing.resample res_type="fft"/"scipy",
// mindaudio/data/processing.py:132-186 -> scipy.signal.resample).  complex128 throughout: the reference computes in
// float64 and the lengths are arbitrary (ceil(len * ratio)), so both DFTs go through Bluestein (bigfft.cuh).
#include "bigfft.cuh"
#include "common.cuh"

namespace mafe {

// a[row][m] = x[row][m] * chirp (m < n) else 0; real input
__global__ void __launch_bounds__(256) bs_pre_real_kernel(const double* __restrict__ x, int64_t n, int64_t M, cd* __restrict__ a) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  cd v{0.0, 0.0};
  if (m < n) {
    const cd w = chirp(m, n, -1);
    const double s = x[(int64_t)blockIdx.y * n + m];
    v = cd{s * w.x, s * w.y};
  }
  a[(int64_t)blockIdx.y * M + m] = v;
}

// inverse side: a[row][k] = Y[k] * chirp(+) (k < num) else 0, Y built from the forward spectrum on the fly
__global__ void __launch_bounds__(256) bs_pre_spec_kernel(const cd* __restrict__ X, int64_t n_x, int64_t num, int64_t M,
                                                         cd* __restrict__ a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= M) return;
  cd v{0.0, 0.0};
  if (k < num) v = cd_mul(resample_bin(X + (int64_t)blockIdx.y * n_x, k, n_x, num), chirp(k, num, +1));
  a[(int64_t)blockIdx.y * M + k] = v;
}

__global__ void __launch_bounds__(256) bs_b_kernel(int64_t N, int64_t M, int sign, cd* __restrict__ b) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) b[m] = bluestein_b(m, N, M, sign);
}

__global__ void __launch_bounds__(256) stockham2_kernel(const cd* __restrict__ src, cd* __restrict__ dst, int64_t p, int64_t half,
                                                       int sign) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const int64_t row = (int64_t)blockIdx.y * 2 * half;
  stockham2(src + row, dst + row, i, p, half, sign);
}

__global__ void __launch_bounds__(256) bs_mul_kernel(cd* __restrict__ a, const cd* __restrict__ bf, int64_t M) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m < M) {
    cd* p = a + (int64_t)blockIdx.y * M + m;
    *p = cd_mul(*p, bf[m]);
  }
}

// forward: X[row][k] = chirp(-) * conv[k] / M
__global__ void __launch_bounds__(256) bs_post_spec_kernel(const cd* __restrict__ conv, int64_t n, int64_t M, cd* __restrict__ X) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const cd v = cd_mul(conv[(int64_t)blockIdx.y * M + k], chirp(k, n, -1));
  const double s = 1.0 / (double)M;
  X[(int64_t)blockIdx.y * n + k] = cd{v.x * s, v.y * s};
}

// inverse: out[row][j] = Re(chirp(+) * conv[j]) / M / n_x   (irfft's 1/num times scipy's num / n_x)
__global__ void __launch_bounds__(256) bs_post_real_kernel(const cd* __restrict__ conv, int64_t num, int64_t M, double scale,
                                                          double* __restrict__ out) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= num) return;
  const cd c = conv[(int64_t)blockIdx.y * M + j];
  const cd w = chirp(j, num, +1);
  out[(int64_t)blockIdx.y * num + j] = (c.x * w.x - c.y * w.y) * scale;
}

namespace {

// in-place semantics over a ping-pong pair: returns the buffer that holds the result
cd* fft_pow2(mafe_ctx* ctx, cd* a, cd* tmp, int64_t M, int rows, int sign, cudaError_t* err) {
  const int64_t half = M / 2;
  const dim3 grid((unsigned)((half + 255) / 256), (unsigned)rows);
  for (int64_t p = 1; p < M; p <<= 1) {
    stockham2_kernel<<<grid, 256, 0, ctx->stream>>>(a, tmp, p, half, sign);
    ctx->launches++;
    cd* t = a; a = tmp; tmp = t;
  }
  *err = cudaGetLastError();
  return a;
}

inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace
}  // namespace mafe

using namespace mafe;

extern "C" int mafe_resample_workspace(int32_t rows, int64_t n_in, int64_t n_out, size_t* bytes) {
  MAFE_REQUIRE(bytes != nullptr, "mafe_resample_workspace: bytes is NULL");
  MAFE_REQUIRE(rows >= 0 && n_in >= 1 && n_out >= 1, "mafe_resample_workspace: bad shape");
  MAFE_REQUIRE(n_in <= (1LL << 26) && n_out <= (1LL << 26), "mafe_resample_fft: signals longer than 2^26 samples are not supported");
  const int64_t M = bluestein_size(n_in > n_out ? n_in : n_out);
  // two ping-pong planes of rows x M, the kernel sequence and its ping-pong (M each), the forward spectrum rows x n_in
  *bytes = align_up((size_t)rows * M * sizeof(cd)) * 2 + align_up((size_t)M * sizeof(cd)) * 2 +
           align_up((size_t)rows * n_in * sizeof(cd));
  return MAFE_OK;
}

extern "C" int mafe_resample_fft(mafe_ctx* ctx, const double* x, int32_t rows, int64_t n_in, int64_t n_out, double* out, void* work,
                                 size_t work_bytes) {
  MAFE_REQUIRE(ctx != nullptr, "ctx is NULL");
  size_t need = 0;
  int rc = mafe_resample_workspace(rows, n_in, n_out, &need);
  if (rc != MAFE_OK) return rc;
  if (rows == 0) return MAFE_OK;
  MAFE_REQUIRE(x && out && work, "mafe_resample_fft: NULL buffer");
  MAFE_REQUIRE(work_bytes >= need, "mafe_resample_fft: workspace of %zu bytes, %zu needed", work_bytes, need);
  MAFE_REQUIRE(rows <= 65535, "mafe_resample_fft: at most 65535 signals per call");
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  const int64_t M = bluestein_size(n_in > n_out ? n_in : n_out);   // one transform size serves both directions
  char* w = (char*)work;
  cd* A = (cd*)w; w += align_up((size_t)rows * M * sizeof(cd));
  cd* T = (cd*)w; w += align_up((size_t)rows * M * sizeof(cd));
  cd* B = (cd*)w; w += align_up((size_t)M * sizeof(cd));
  cd* BT = (cd*)w; w += align_up((size_t)M * sizeof(cd));
  cd* X = (cd*)w;
  const dim3 gM((unsigned)((M + 255) / 256), (unsigned)rows), gM1((unsigned)((M + 255) / 256), 1);
  cudaError_t err = cudaSuccess;

  // ---- forward DFT of length n_in (sign -1)
  bs_pre_real_kernel<<<gM, 256, 0, st>>>(x, n_in, M, A);
  bs_b_kernel<<<gM1, 256, 0, st>>>(n_in, M, -1, B);
  ctx->launches += 2;
  cd* Bf = fft_pow2(ctx, B, BT, M, 1, -1, &err);
  MAFE_CUDA_CHECK(err);
  cd* Af = fft_pow2(ctx, A, T, M, rows, -1, &err);
  MAFE_CUDA_CHECK(err);
  cd* Aother = Af == A ? T : A;
  bs_mul_kernel<<<gM, 256, 0, st>>>(Af, Bf, M);
  ctx->launches++;
  cd* conv = fft_pow2(ctx, Af, Aother, M, rows, +1, &err);
  MAFE_CUDA_CHECK(err);
  bs_post_spec_kernel<<<dim3((unsigned)((n_in + 255) / 256), (unsigned)rows), 256, 0, st>>>(conv, n_in, M, X);
  MAFE_LAUNCH_CHECK(ctx);

  // ---- inverse DFT of length n_out (sign +1) of the resampled spectrum, real part
  bs_pre_spec_kernel<<<gM, 256, 0, st>>>(X, n_in, n_out, M, A);
  bs_b_kernel<<<gM1, 256, 0, st>>>(n_out, M, +1, B);
  ctx->launches += 2;
  Bf = fft_pow2(ctx, B, BT, M, 1, -1, &err);
  MAFE_CUDA_CHECK(err);
  Af = fft_pow2(ctx, A, T, M, rows, -1, &err);
  MAFE_CUDA_CHECK(err);
  Aother = Af == A ? T : A;
  bs_mul_kernel<<<gM, 256, 0, st>>>(Af, Bf, M);
  ctx->launches++;
  conv = fft_pow2(ctx, Af, Aother, M, rows, +1, &err);
  MAFE_CUDA_CHECK(err);
  bs_post_real_kernel<<<dim3((unsigned)((n_out + 255) / 256), (unsigned)rows), 256, 0, st>>>(conv, n_out, M,
                                                                                           1.0 / ((double)M * (double)n_in), out);
  MAFE_LAUNCH_CHECK(ctx);
  return MAFE_OK;
}
