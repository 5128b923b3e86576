// Host replay of the v3 kernel's mel sweep PROGRAM (mindaudio_b200/csrc/fbank512.cu: build_bins, build_v3_program;
// fbank512_v3.cuh: sweep_v3, phase C): the per-bin weights, retire counts (steps and register masks), warp ranges, plane
// rows and the two-row combine table are executed here exactly as the kernel executes them -- with one "lane" -- and the
// result is compared with the dense filterbank product  mel[m] = sum_k fb[m][k] * P[k].   argv[1]: float32 [80][257].
// With argv[2] = n_mels the one-pass program of fbank400_kernel (build_f400_program, 201 bins) is replayed instead.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../mindaudio_b200/csrc/fbank512.cu"
using namespace mafe;

static int check400(const char* path, int nm) {
  std::vector<float> fb((size_t)nm * kBins400);
  FILE* fh = fopen(path, "rb");
  if (!fh || fread(fb.data(), sizeof(float), fb.size(), fh) != fb.size()) { printf("cannot read %s\n", path); return 2; }
  fclose(fh);
  mafe_frontend_desc d;
  memset(&d, 0, sizeof(d));
  d.n_mels = nm;
  d.mel_fb = fb.data();
  std::vector<BinEntry> bins;
  if (!build_bins_n(&d, kBins400, 1.0f, bins)) { printf("build_bins_n: not a two-adjacent-filters bank\n"); return 3; }
  static F400Sweep S;
  std::vector<int> comb;
  if (!build_f400_program(bins, nm, S, comb)) { printf("build_f400_program: bank does not fit\n"); return 3; }
  srand(6);
  std::vector<double> P(kBins400);
  for (auto& v : P) v = 1e-3 + rand() / (double)RAND_MAX;
  std::vector<double> rows(kMaxRows400, 0.0);
  std::vector<int> covered(kBins400, 0);
  for (int w = 0; w < kFastWarps; ++w) {
    int si = S.kk0[w], row = S.row0[w], word = 0;
    const int si_end = S.kk0[w + 1];
    double lo = 0, hi = 0;
    auto retire = [&](int n) {
      for (; n > 0; --n) {
        if (row >= S.zero_row) { printf("row overflow (warp %d)\n", w); exit(1); }
        rows[row++] = lo; lo = hi; hi = 0;
      }
    };
    while (si != si_end) {
      uint32_t m = S.nret_mask[w][word++];
      const int chunk_end = std::min(si + 16, si_end);
      while (si != chunk_end) {
        const V3Step& st = S.step[si];
        const int nr = m & 3u;
        m >>= 2;
        if (nr != st.nret) { printf("mask / step retire count differ at step %d\n", si); return 1; }
        retire(nr);
        covered[si]++;
        lo += (double)st.w0 * P[si];
        hi += (double)st.w1 * P[si];
        ++si;
      }
    }
    retire(S.tail[w]);
    const int expect_rows = (w + 1 < kFastWarps ? S.row0[w + 1] : S.zero_row) - S.row0[w];
    if (row - S.row0[w] != expect_rows) { printf("warp %d retired %d rows, owns %d\n", w, row - S.row0[w], expect_rows); return 1; }
  }
  for (int k = 0; k < kBins400; ++k) if (covered[k] != 1) { printf("bin %d swept %d times\n", k, covered[k]); return 1; }
  double worst = 0;
  for (int m = 0; m < nm; ++m) {
    const double got = rows[comb[m] & 0xff] + rows[comb[m] >> 8];
    double ref = 0;
    for (int k = 0; k < kBins400; ++k) ref += (double)fb[(size_t)m * kBins400 + k] * P[k];
    worst = std::max(worst, std::abs(got - ref) / std::max(1e-12, std::abs(ref)));
  }
  printf("f400: %d mels, rows %d, max rel err %.3g\n", nm, S.zero_row, worst);
  return worst < 1e-6 ? 0 : 1;
}

// the 16-range (half-warp) program: same replay, three-row combine
static int check_v5(const std::vector<float>& fb, const std::vector<BinEntry>& bins) {
  static V5Sweep S;
  std::vector<int> comb5;
  if (!build_v5_program(bins, S, comb5)) { printf("build_v5_program: bank does not fit\n"); return 3; }
  srand(9);
  std::vector<double> P(kBins);
  for (auto& v : P) v = 1e-3 + rand() / (double)RAND_MAX;
  const int W = kV5Ranges;
  std::vector<double> rows(kV5PlaneRows, 0.0);
  std::vector<int> written(kV5PlaneRows, 0);
  int covered[2][kV3HalfStride] = {};
  int max_steps = 0;
  for (int g = 0; g < 2; ++g)
    for (int w = 0; w < W; ++w) {
      int si = g * kV3HalfStride + S.kk0[w];
      const int si_end = g * kV3HalfStride + S.kk0[w + 1] + ((g == 0 && w == W - 1) ? 1 : 0);
      max_steps = std::max(max_steps, si_end - si);
      int row = S.row0[w], word = 0;
      double lo = 0, hi = 0;
      auto retire = [&](int n) {
        for (; n > 0; --n) {
          if (row >= S.zero_row) { printf("v5 row overflow (g %d range %d)\n", g, w); exit(1); }
          if (g) { if (!written[row]) { printf("v5 row %d accumulated before it was written\n", row); exit(1); } rows[row] += lo; }
          else { rows[row] = lo; written[row]++; }
          lo = hi; hi = 0; ++row;
        }
      };
      while (si != si_end) {
        uint32_t m = S.nret_mask[g * W + w][word++];
        const int chunk_end = std::min(si + 16, si_end);
        while (si != chunk_end) {
          const V3Step& st = S.step[si];
          const int nr = m & 3u;
          m >>= 2;
          if (nr != st.nret) { printf("v5 mask / step retire count differ at step %d\n", si); return 1; }
          retire(nr);
          const int kk = si - g * kV3HalfStride, k = 2 * kk + g;
          covered[g][kk]++;
          lo += (double)st.w0 * P[k];
          hi += (double)st.w1 * P[k];
          ++si;
        }
      }
      retire(S.tail[g][w]);
      const int expect_rows = (w + 1 < W ? S.row0[w + 1] : S.zero_row) - S.row0[w];
      if (row - S.row0[w] != expect_rows) { printf("v5 range %d half %d retired %d rows, owns %d\n", w, g, row - S.row0[w], expect_rows); return 1; }
    }
  for (int k = 0; k < kBins; ++k)
    if (covered[k & 1][k >> 1] != 1) { printf("v5 bin %d swept %d times\n", k, covered[k & 1][k >> 1]); return 1; }
  double worst = 0;
  for (int m = 0; m < kV2Mels; ++m) {
    const double got = rows[comb5[m] & 0xff] + rows[(comb5[m] >> 8) & 0xff] + rows[(comb5[m] >> 16) & 0xff];
    double ref = 0;
    for (int k = 0; k < kBins; ++k) ref += 0.25 * (double)fb[(size_t)m * kBins + k] * P[k];
    worst = std::max(worst, std::abs(got - ref) / std::max(1e-12, std::abs(ref)));
  }
  printf("v5: rows %d, longest range %d steps, max rel err %.3g\n", S.zero_row, max_steps, worst);
  return worst < 1e-6 ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: %s filterbank.f32 [n_mels: 400-point program]\n", argv[0]); return 2; }
  if (argc > 2) return check400(argv[1], atoi(argv[2]));
  std::vector<float> fb((size_t)kV2Mels * kBins);
  FILE* fh = fopen(argv[1], "rb");
  if (!fh || fread(fb.data(), sizeof(float), fb.size(), fh) != fb.size()) { printf("cannot read %s\n", argv[1]); return 2; }
  fclose(fh);
  mafe_frontend_desc d;
  memset(&d, 0, sizeof(d));
  d.n_mels = kV2Mels;
  d.mel_fb = fb.data();
  std::vector<BinEntry> bins;
  if (!build_bins(&d, bins)) { printf("build_bins: not a two-adjacent-filters bank\n"); return 3; }
  static V3Sweep S;
  std::vector<int> comb3;
  if (!build_v3_program(bins, S, comb3)) { printf("build_v3_program: bank does not fit the compact planes\n"); return 3; }

  srand(5);
  std::vector<double> P(kBins);   // |2X|^2 of one frame
  for (auto& v : P) v = 1e-3 + rand() / (double)RAND_MAX;
  const int W = kFastWarps;
  std::vector<double> rows(kV3PlaneRows, 0.0);
  std::vector<int> written(kV3PlaneRows, 0);
  int covered[2][kV3HalfStride] = {};
  for (int g = 0; g < 2; ++g)
    for (int w = 0; w < W; ++w) {
      int si = g * kV3HalfStride + S.kk0[w];
      const int si_end = g * kV3HalfStride + S.kk0[w + 1] + ((g == 0 && w == W - 1) ? 1 : 0);
      int row = S.row0[w], word = 0;
      double lo = 0, hi = 0;
      auto retire = [&](int n) {
        for (; n > 0; --n) {
          if (row >= S.zero_row) { printf("row overflow (g %d warp %d)\n", g, w); exit(1); }
          if (g) { if (!written[row]) { printf("row %d accumulated before it was written\n", row); exit(1); } rows[row] += lo; }
          else { rows[row] = lo; written[row]++; }
          lo = hi; hi = 0; ++row;
        }
      };
      while (si != si_end) {
        uint32_t m = S.nret_mask[g * W + w][word++];
        const int chunk_end = std::min(si + 16, si_end);
        while (si != chunk_end) {
          const V3Step& st = S.step[si];
          const int nr = m & 3u;
          m >>= 2;
          if (nr != st.nret) { printf("mask / step retire count differ at step %d\n", si); return 1; }
          retire(nr);
          const int kk = si - g * kV3HalfStride, k = 2 * kk + g;
          if (k >= kBins) { printf("step %d is bin %d\n", si, k); return 1; }
          covered[g][kk]++;
          lo += (double)st.w0 * P[k];
          hi += (double)st.w1 * P[k];
          ++si;
        }
      }
      retire(S.tail[g][w]);
      const int expect_rows = (w + 1 < W ? S.row0[w + 1] : S.zero_row) - S.row0[w];
      if (row - S.row0[w] != expect_rows) { printf("warp %d half %d retired %d rows, owns %d\n", w, g, row - S.row0[w], expect_rows); return 1; }
    }
  for (int k = 0; k < kBins; ++k)
    if (covered[k & 1][k >> 1] != 1) { printf("bin %d swept %d times\n", k, covered[k & 1][k >> 1]); return 1; }
  if (rows[S.zero_row] != 0.0) { printf("zero row written\n"); return 1; }
  double worst = 0;
  for (int m = 0; m < kV2Mels; ++m) {
    const double got = rows[comb3[m] & 0xff] + rows[comb3[m] >> 8];
    double ref = 0;
    for (int k = 0; k < kBins; ++k) ref += 0.25 * (double)fb[(size_t)m * kBins + k] * P[k];
    worst = std::max(worst, std::abs(got - ref) / std::max(1e-12, std::abs(ref)));
  }
  int lo_r = 255, hi_r = 0;
  for (int w = 0; w < W; ++w) { lo_r = std::min(lo_r, S.kk0[w + 1] - S.kk0[w]); hi_r = std::max(hi_r, S.kk0[w + 1] - S.kk0[w]); }
  printf("rows %d, warp ranges %d..%d sub-bins, max rel err %.3g\n", S.zero_row, lo_r, hi_r, worst);
  if (worst >= 1e-6) return 1;
  return check_v5(fb, bins);
}
