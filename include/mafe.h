/*
 * mafe.h -- C ABI of libmafe.so: MindAudio Front-End feature path on B200 (sm_100a).
 *
 * The reference (mindspore-lab/mindaudio) has NO native code and NO FFI on this
 * path: its boundary is the Python call surface of mindaudio/data/spectrum.py and
 * mindaudio/data/features.py (numpy in / numpy out) plus the conformer example's
 * front-end (examples/conformer/dataset.py:56-168) and the CMVN family.  This header
 * is the thinnest C cut under that surface; every entry point names the reference
 * interface it serves.  The Python mirror of the reference API that binds these
 * symbols with ctypes lives in mindaudio_b200/ (see INTEGRATION.md).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no exceptions across the boundary.
 *   - every function returns a status (MAFE_OK == 0, negative = error); the message of
 *     the last error on the calling thread is mafe_last_error().
 *   - the caller owns every buffer; the library owns only the opaque handles it returns.
 *   - "dev" pointers are CUDA device pointers on the ctx's device; "host" pointers are
 *     ordinary host memory.  Work is enqueued on the ctx's stream and is asynchronous
 *     unless stated otherwise; mafe_ctx_sync() waits for it.
 *   - ragged batches: utterance u owns samples [sample_offsets[u], sample_offsets[u+1])
 *     of one flat waveform array and frames [frame_offsets[u], frame_offsets[u+1]) of
 *     one flat, frame-major output array.  A dense [B, L] batch is the special case of
 *     equal lengths.
 *   - fork safety: nothing touches CUDA before the first mafe_ctx_create() of a process.
 */
#ifndef MAFE_H_
#define MAFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MAFE_VERSION 102 /* 0.1.2: mafe_frontend_run_aux (0.1.1: mafe_frontend_desc gained utt_scalar_norm) */

/* ---- status codes ---- */
#define MAFE_OK 0
#define MAFE_E_INVALID_ARG (-1)
#define MAFE_E_CUDA (-2)
#define MAFE_E_OOM (-3)
#define MAFE_E_UNSUPPORTED (-4)

/* ---- enums (plain ints in the structs) ---- */
/* np.pad / mindspore BorderType modes used by stft (spectrum.py:132,212-232) and Spectrogram */
#define MAFE_PAD_CONSTANT 0
#define MAFE_PAD_REFLECT 1
#define MAFE_PAD_EDGE 2
#define MAFE_PAD_SYMMETRIC 3

/* what one frame produces */
#define MAFE_OUT_COMPLEX 0 /* complex64 [n_bins]           -- spectrum.stft            */
#define MAFE_OUT_POWER 1   /* float |X|^power [n_bins]     -- spectrum.spectrogram     */
#define MAFE_OUT_MEL 2     /* float mel energies [n_mels]  -- spectrum.melspectrogram  */
#define MAFE_OUT_LOGMEL 3  /* float log-mel [n_mels]       -- features.fbank / conformer fbank */
#define MAFE_OUT_MFCC 4    /* float DCT-II of log-mel [n_mfcc] -- features.mfcc        */

/* log applied to mel energies */
#define MAFE_LOG_NONE 0
#define MAFE_LOG_LN_EPS_IF_ZERO 1 /* ln(x == 0 ? DBL_EPSILON : x)  conformer/dataset.py:154-155 */
#define MAFE_LOG_LN_PLUS 2        /* ln(x + log_arg)               features.py:349-350 (log_mels); also valid on
                                   * MAFE_OUT_POWER: ln(|X|^power + log_arg) -- deepspeech2's log1p(magnitude),
                                   * examples/deepspeech2/dataset.py:42-43, written by the transform itself */
#define MAFE_LOG_DB 3             /* mult*log10(max(x, amin)) - mult*log10(max(amin, ref)) spectrum.py:73-76 */

/* waveform element types */
#define MAFE_WAVE_F32 0
#define MAFE_WAVE_I16 1 /* PCM16; scaled by wave_scale on load (io.py:741-745 gives /32768) */

/* top_db clamp grouping (spectrum.py:78-89) */
#define MAFE_DBGROUP_NONE 0  /* top_db = None                                             */
#define MAFE_DBGROUP_UTT 1   /* 2-D input: clamp per utterance                            */
#define MAFE_DBGROUP_BATCH 2 /* 3-D input: ONE floor for the whole call (batch coupling)  */
#define MAFE_DBGROUP_MAP 3   /* explicit utt -> group map (4-D input: per leading item)   */

typedef struct mafe_ctx mafe_ctx;
typedef struct mafe_plan mafe_plan;
typedef struct mafe_batch mafe_batch;

/*
 * Immutable description of one front-end configuration.  All table pointers are HOST
 * pointers read during mafe_plan_create() only.
 */
typedef struct mafe_frontend_desc {
  int32_t n_fft;       /* DFT size (spectrum.py:127; conformer: 512, dataset.py:166)                  */
  int32_t frame_len;   /* samples taken per frame; < n_fft => zero-padded at the END (np.fft.rfft(n=)) */
  int32_t hop;         /* frame shift                                                                  */
  int32_t center;      /* 1: pad n_fft/2 both sides so frame t is centred at t*hop (spectrum.py:181)   */
  int32_t pad_mode;    /* MAFE_PAD_* for center=1                                                      */
  int32_t out_kind;    /* MAFE_OUT_*                                                                   */
  const float* window; /* [frame_len] analysis window, already padded/centred by the caller            */

  double preemph;            /* 0 = off; y[0]=x[0], y[n]=x[n]-c*x[n-1] over the whole utterance (dataset.py:117-119) */
  int32_t remove_frame_mean; /* 1: subtract ONE scalar = mean of all windowed frame entries (dataset.py:165)        */
  float dither;              /* 0 = off (reference path); else x += dither * N(0,1), Philox4x32-10       */
  uint64_t dither_seed;

  float power;      /* MAFE_OUT_POWER/MEL...: |X|^power (2 = re^2+im^2, 1 = magnitude)       */
  float spec_scale; /* multiplies X before |.|^power (Spectrogram normalized=True); 1 = off */

  int32_t n_mels;       /* rows of mel_fb                                                        */
  const float* mel_fb;  /* [n_mels][n_fft/2+1] dense filterbank (rows = filters)                 */
  int32_t log_kind;     /* MAFE_LOG_*                                                            */
  float log_arg;        /* LN_PLUS: added constant; DB: amin                                     */
  float log_mult;       /* DB: 10 (power) or 20 (magnitude)                                      */
  float log_offset;     /* DB: mult*log10(max(amin, ref)) subtracted                             */
  float top_db;         /* DB: < 0 disables the clamp                                            */

  int32_t n_mfcc;   /* MAFE_OUT_MFCC: columns of dct                                            */
  const float* dct; /* [n_mels][n_mfcc] (create_dct layout, features.py:337)                    */

  /* fused per-utterance CMVN of the frame-major output (examples/ECAPA-TDNN/spec_augment.py:43-70):
   * (x - mean_t) / std_t per feature dim over the utterance's frames; both 0 = off */
  int32_t utt_cmvn_mean;
  int32_t utt_cmvn_std;

  int32_t allow_fast_path; /* 1: use a specialised kernel when the configuration has one        */

  /* fused per-utterance SCALAR normalisation of the output (examples/deepspeech2/dataset.py:44-47):
   * (x - mean) / std over ALL elements of the utterance's [T, out_dim] matrix; excludes utt_cmvn_*. */
  int32_t utt_scalar_norm;
} mafe_frontend_desc;

/* ---- library / errors ---- */
int mafe_version(void);
const char* mafe_last_error(void);

/* ---- context: one device + one stream ---- */
int mafe_ctx_create(int device, mafe_ctx** out);
int mafe_ctx_destroy(mafe_ctx* ctx);
/* Run on a caller-owned cudaStream_t (e.g. torch's current stream); NULL = the ctx's own stream. */
int mafe_ctx_set_stream(mafe_ctx* ctx, void* cuda_stream);
int mafe_ctx_sync(mafe_ctx* ctx);
int mafe_ctx_sm_count(const mafe_ctx* ctx);
/* Number of kernels this ctx has launched since creation (bench.py's gpu_launches). */
int64_t mafe_ctx_launch_count(const mafe_ctx* ctx);

/* Per-kernel device timing for bench.py's roofline: when enabled, every launch of the hot-path kernels is
 * bracketed by CUDA events on the ctx stream.  mafe_ctx_profile_read() synchronises, then returns the summed
 * milliseconds and the launch count of kernel class `which` (MAFE_PROF_*) since the last reset. */
#define MAFE_PROF_FBANK_MAIN 0   /* fbank512_kernel / generic_frontend_kernel */
#define MAFE_PROF_FRAME_MEAN 1   /* utterance frame-mean pre-pass            */
#define MAFE_PROF_CMVN 2         /* CMVN kernels                             */
#define MAFE_PROF_OTHER 3
#define MAFE_PROF_COUNT 4
int mafe_ctx_profile_enable(mafe_ctx* ctx, int32_t enable);
int mafe_ctx_profile_read(mafe_ctx* ctx, int32_t which, double* ms_out, int64_t* launches_out);
int mafe_ctx_profile_reset(mafe_ctx* ctx);

/* FP32 FMA peak of the device (TFLOP/s) from a register-resident FFMA microbenchmark: the denominator of the
 * compute roofline of the fbank path (SURVEY.md section 8d asks for it to be measured, not assumed). */
int mafe_fp32_fma_peak(mafe_ctx* ctx, double* tflops_out);

/* ---- memory plumbing for numpy callers (no torch needed) ---- */
int mafe_device_malloc(mafe_ctx* ctx, size_t bytes, void** out_dev);
int mafe_device_free(mafe_ctx* ctx, void* dev);
int mafe_pinned_malloc(mafe_ctx* ctx, size_t bytes, void** out_host);
int mafe_pinned_free(mafe_ctx* ctx, void* host);
int mafe_memcpy_h2d(mafe_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes); /* async on ctx stream */
int mafe_memcpy_d2h(mafe_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes); /* async on ctx stream */
int mafe_memset(mafe_ctx* ctx, void* dst_dev, int value, size_t bytes);
/* numpy-in / numpy-out plumbing of the reference-facing Python layer.
 * h2d_gather: n host arrays (src_host[i], bytes[i]; pageable is fine) land back to back at dst_dev.  The library's host
 *   threads copy them in parallel into a pinned staging buffer of the ctx, ONE cudaMemcpyAsync uploads it (a Python-side
 *   np.concatenate + pageable upload of 128 utterances cost 25 ms; this is ~5).  Returns when the staging copy is complete
 *   and the upload is enqueued: the sources may be reused at once.
 * d2h_staged: device -> pinned staging (full PCIe rate), stream synchronised, then a parallel copy into dst_host (fresh
 *   numpy pages are first touched by several threads).  Synchronous. */
int mafe_memcpy_h2d_gather(mafe_ctx* ctx, void* dst_dev, const void* const* src_host, const int64_t* bytes, int32_t n);
int mafe_memcpy_d2h_staged(mafe_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);

/* ---- plans ---- */
int mafe_plan_create(mafe_ctx* ctx, const mafe_frontend_desc* desc, mafe_plan** out);
int mafe_plan_destroy(mafe_plan* plan);
/* frames produced for an utterance of n_samples (0 if too short); spectrum.py:298, dataset.py:127 */
int64_t mafe_plan_num_frames(const mafe_plan* plan, int64_t n_samples);
/* floats per frame of output (complex counts 2 per bin) */
int32_t mafe_plan_out_dim(const mafe_plan* plan);
/* 1 if the specialised sm_100a kernel serves this plan, 0 if the generic mixed-radix kernel does */
int32_t mafe_plan_is_fast(const mafe_plan* plan);

/* ---- ragged batch layout (host offsets -> device tables) ---- */
/* sample_offsets_host: int64[n_utts+1].  utt_group_host: int32[n_utts] or NULL (MAFE_DBGROUP_MAP). */
int mafe_batch_create(mafe_ctx* ctx, const mafe_plan* plan, const int64_t* sample_offsets_host, int32_t n_utts,
                      const int32_t* utt_group_host, mafe_batch** out);
int mafe_batch_destroy(mafe_batch* batch);
/* Re-lay an existing batch object for new offsets (any plan of the same ctx).  Its device tables only grow, so a caller that
 * keeps batch objects between calls pays no cudaMalloc / cudaFree per batch (a per-utterance API call otherwise spends more
 * time in them than in the kernels).  Stream ordered on the ctx stream: work enqueued earlier with the old layout is safe. */
int mafe_batch_refill(mafe_ctx* ctx, const mafe_plan* plan, mafe_batch* batch, const int64_t* sample_offsets_host, int32_t n_utts,
                      const int32_t* utt_group_host);
int64_t mafe_batch_total_frames(const mafe_batch* batch);
int64_t mafe_batch_total_samples(const mafe_batch* batch);
/* frame_offsets_host_out: int64[n_utts+1] */
int mafe_batch_frame_offsets(const mafe_batch* batch, int64_t* frame_offsets_host_out);
/* device copy of the frame offsets (int64[n_utts+1]) for the CMVN calls below */
const int64_t* mafe_batch_frame_offsets_dev(const mafe_batch* batch);

/*
 * The hot path.  wave_dev: flat waveform (MAFE_WAVE_*), wave_scale multiplies every sample on
 * load (1.0f; 32768.0f reproduces read()*(1<<15), dataset.py:389-390, from float input in [-1,1)).
 * out_dev: [total_frames][out_dim] floats, frame-major (time-major), ragged by frame offsets.
 * db_group: MAFE_DBGROUP_* (only for log_kind == MAFE_LOG_DB with top_db >= 0).
 * Serves: spectrum.stft (spectrum.py:125-278), spectrum.spectrogram (:547-606),
 * spectrum.melspectrogram (:609-698), features.fbank (features.py:196-270), features.mfcc
 * (:273-373, without deltas/context), conformer compute_fbank_feats (dataset.py:159-168).
 */
int mafe_frontend_run(mafe_ctx* ctx, const mafe_plan* plan, mafe_batch* batch, const void* wave_dev, int32_t wave_dtype,
                      float wave_scale, float* out_dev, int32_t db_group);

/*
 * mafe_frontend_run of a MAFE_OUT_LOGMEL / MAFE_OUT_MFCC plan that ALSO writes the mel energies the features are the
 * log / DCT of -- exactly what the same front-end with out_kind MAFE_OUT_MEL writes -- to aux_mel_dev
 * [total_frames][n_mels].  For pipelines that call melspectrogram() and fbank() / mfcc() on the same waveforms with
 * the same front-end parameters (spectrum.py:609-698 next to features.py:196-270 / :273-373; BASELINE configs[3]):
 * one transform instead of two.  aux_mel_dev == NULL: plain mafe_frontend_run.  Returns MAFE_E_UNSUPPORTED, with
 * nothing launched, when the plan's kernel cannot emit the second output (today: anything but the n_fft 400 tile
 * kernel on 16-byte aligned float32 input with wave_scale 1); the caller then runs a MAFE_OUT_MEL plan beside it.
 */
int mafe_frontend_run_aux(mafe_ctx* ctx, const mafe_plan* plan, mafe_batch* batch, const void* wave_dev, int32_t wave_dtype,
                          float wave_scale, float* out_dev, int32_t db_group, float* aux_mel_dev);

/*
 * The same path with HOST buffers (what a numpy / data-loader caller has): wave_host is the flat waveform
 * in host memory (pinned memory makes the copies asynchronous), out_host receives [total_frames][out_dim].
 * Utterances are processed in chunks of chunk_utts (<= 0: 512) on three internal streams, so the H2D copy of
 * chunk i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap.  Synchronous: returns when out_host
 * is complete.  frame_offsets_host_out: int64[n_utts+1] or NULL.  db_group: MAFE_DBGROUP_NONE / _UTT only.
 * On an error the call still waits for the chunks it has enqueued; the contents of out_host are then unspecified.
 */
int mafe_frontend_run_host(mafe_ctx* ctx, const mafe_plan* plan, const int64_t* sample_offsets_host, int32_t n_utts,
                           const void* wave_host, int32_t wave_dtype, float wave_scale, float* out_host,
                           int64_t* frame_offsets_host_out, int32_t chunk_utts, int32_t db_group);

/* ---- spectral element-wise ops on device arrays ---- */
/* spectrum.magphase iscomplex=True (spectrum.py:720-732): n complex64 -> mag^power, unit phase (0 -> 1+0j). */
int mafe_magphase(mafe_ctx* ctx, const float* z_dev, int64_t n, float power, float* mag_dev, float* phase_dev);
/* spectrum.amplitude_to_dB (spectrum.py:59-90): n_groups contiguous groups of group_size elements;
 * top_db < 0 disables the clamp.  in == out allowed. */
int mafe_amplitude_to_db(mafe_ctx* ctx, const float* x_dev, float* out_dev, int64_t n_groups, int64_t group_size,
                         float mult, float amin, float db_offset, float top_db);
/* spectrum.dB_to_amplitude (spectrum.py:108-113): ref * (10^(0.1 x))^power */
int mafe_db_to_amplitude(mafe_ctx* ctx, const float* x_dev, float* out_dev, int64_t n, float ref, float power);

/* spectrum.melscale (spectrum.py:738-774): spec [n_mats][n_bins][t] (time contiguous) -> out [n_mats][n_mels][t];
 * mel_fb_host: HOST [n_mels][n_bins] dense filterbank. */
int mafe_melscale(mafe_ctx* ctx, const float* spec_dev, float* out_dev, int32_t n_mats, int32_t n_bins, int32_t t,
                  const float* mel_fb_host, int32_t n_mels);
/* layout helper: in [n_mats][rows][cols] -> out [n_mats][cols][rows] (frame-major <-> the reference's [.., freq, time]);
 * out_mat_stride: floats between consecutive output matrices (>= rows*cols; 0 = dense). */
int mafe_transpose(mafe_ctx* ctx, const float* in_dev, float* out_dev, int32_t n_mats, int32_t rows, int32_t cols,
                   int64_t out_mat_stride);

/* ---- istft (spectrum.py:346-474) ---- */
/* spec_dev: [n_utts][n_frames][n_fft/2+1] complex64 frame-major; window: HOST double[n_fft] (padded);
 * y_dev: double [n_utts][n_fft + hop*(n_frames-1)] = overlap-added, window-sum-square normalised signal
 * (the caller trims n_fft/2 / length as the reference does).  float64 arithmetic, like the reference. */
int mafe_istft(mafe_ctx* ctx, const float* spec_dev, int32_t n_utts, int32_t n_frames, int32_t n_fft, int32_t hop,
               const double* window_host, double* y_dev);

/* ---- CMVN family ---- */
/* Per-utterance, per-dim mean (and population-std) normalisation in place
 * (examples/ECAPA-TDNN/spec_augment.py:43-70).  feats_dev: [total_frames][dim] frame-major. */
int mafe_cmvn_utt(mafe_ctx* ctx, float* feats_dev, const int64_t* frame_offsets_dev, int32_t n_utts, int32_t dim,
                  int32_t mean_norm, int32_t std_norm);
/* Per-utterance scalar (x - mean)/std over ALL elements of the utterance, optional log1p first
 * (examples/deepspeech2/dataset.py:43-47). */
int mafe_cmvn_scalar(mafe_ctx* ctx, float* feats_dev, const int64_t* frame_offsets_dev, int32_t n_utts, int32_t dim,
                     int32_t log1p_first);
/* Global statistics (examples/conformer/compute_cmvn_stats.py:61-63,108-112): stats_dev is
 * double[2*dim+1] = {sum x[dim], sum x^2[dim], frame count}; ACCUMULATES (caller zeroes). */
int mafe_cmvn_stats_accumulate(mafe_ctx* ctx, const float* feats_dev, int64_t total_frames, int32_t dim,
                               double* stats_dev);
/* (x - mean) * istd in place (mindaudio/models/layers/cmvn.py:33-36); istd_dev NULL = mean only. */
int mafe_cmvn_apply(mafe_ctx* ctx, float* feats_dev, int64_t total_frames, int32_t dim, const float* mean_dev,
                    const float* istd_dev);

/* ---- feature post-processing ("next" row f1: features.py:69-193) ---- */
/* compute_deltas: n_mats matrices of [rows][t] (time contiguous), matrix strides in floats (0 = dense) so the
 * result can be written straight into the concatenated [.., 3*rows, t] array of features.py:264-267. */
int mafe_compute_deltas(mafe_ctx* ctx, const float* x_dev, float* out_dev, int32_t n_mats, int32_t rows, int32_t t,
                        int64_t x_mat_stride, int64_t out_mat_stride, int32_t win_length, int32_t pad_mode);
/* context_window: in [n_mats][f][t] -> out [n_mats][f*(l+r+1)][t] (features.py:94-155 gather form). */
int mafe_context_window(mafe_ctx* ctx, const float* x_dev, float* out_dev, int32_t n_mats, int32_t f, int32_t t,
                        int32_t left, int32_t right);

/* ---- collate ("next" row f3: mindaudio/utils/common.py:10-52 pad_sequence, examples/conformer/dataset.py:563-569,
 * 616-621) ---- */
/* Ragged features [total_frames][dim] -> padded batch out_dev [n_utts][max_len][dim] (batch_first != 0) or
 * [max_len][n_utts][dim].  Utterance i owns rows frame_offsets[i] .. frame_offsets[i+1]; rows beyond its length get
 * padding_value, longer utterances are truncated to max_len.  mask_dev (may be NULL): float [n_utts][max_len], 1 for
 * frames of the utterance, 0 for padding (= ~make_pad_mask(lengths, max_len) as float32). */
int mafe_pad_sequence(mafe_ctx* ctx, const float* feats_dev, const int64_t* frame_offsets_dev, int32_t n_utts, int32_t dim,
                      int32_t max_len, float padding_value, int32_t batch_first, float* out_dev, float* mask_dev);

/* ---- feature-domain ops of the pipelines ("next" row f4) ---- */
/* sliding_window_cmn (mindaudio/data/processing.py:380-407 -> msaudio.SlidingWindowCmn; Kaldi's sliding-window CMN as
 * in torchaudio.functional.sliding_window_cmn): x, out [n_channels][num_frames][num_feats]; out must NOT alias x
 * (frames leave the window after they have been normalised). */
int mafe_sliding_window_cmn(mafe_ctx* ctx, const float* x_dev, float* out_dev, int32_t n_channels, int32_t num_frames,
                            int32_t num_feats, int32_t cmn_window, int32_t min_cmn_window, int32_t center, int32_t norm_vars);
/* Rectangle masking, in place (SpecAugment: examples/conformer/dataset.py:493-534; mindaudio/data/augment.py:28-98
 * frequencymasking / timemasking).  feats_dev is a ragged [total_rows][dim] array with row offsets
 * (frame_offsets_dev[n_items + 1]); rects_dev holds n_rects x (item, row0, row1, col0, col1) int32, half-open ranges
 * relative to the item, clipped to the item's extent; every covered element is set to value. */
int mafe_mask_rects(mafe_ctx* ctx, float* feats_dev, const int64_t* frame_offsets_dev, int32_t n_items, int32_t dim,
                    const int32_t* rects_dev, int32_t n_rects, float value);

/* ---- phase vocoder ("next" row f2: mindaudio/data/augment.py:795-871 time_stretch / _phase_vocoder) ---- */
/* spec_dev [n_mats][n_frames][n_bins] complex64 frame-major -> out_dev [n_mats][n_steps][n_bins]: time step t reads
 * frames floor(t*rate), +1 (zero beyond the end), interpolates the magnitudes and accumulates the phase advance
 * (float32 accumulator updated in float64, as numpy does for `phase_acc += ...`).  phi_advance_dev: double[n_bins]
 * expected phase advance per bin (linspace(0, pi*hop, n_bins)). */
int mafe_phase_vocoder(mafe_ctx* ctx, const float* spec_dev, int32_t n_mats, int32_t n_frames, int32_t n_bins, double rate,
                       const double* phi_advance_dev, int32_t n_steps, float* out_dev);

/* ---- harmonic / percussive separation ("next" row f2: mindaudio/data/features.py:438-559 soft_mask / hpss / harmonic) ---- */
/* scipy.ndimage.median_filter(x, size, mode="reflect") along one axis of n_mats row-major [rows][cols] matrices:
 * window [i - size/2, i - size/2 + size), rank size/2, boundary d c b a | a b c d | d c b a.  axis 0 = along rows
 * (frequency: the percussive filter), 1 = along columns (time: the harmonic filter).  out must not alias x. */
int mafe_median_filter(mafe_ctx* ctx, const float* x_dev, float* out_dev, int32_t n_mats, int32_t rows, int32_t cols,
                       int32_t size, int32_t axis);
/* soft masks of hpss (features.py:438-469, 513-528) from the two median-filtered magnitudes, elementwise over n values:
 * mask_h = soft_mask(harm, perc * margin_h), mask_p = soft_mask(perc, harm * margin_p), power > 0 (INFINITY = hard
 * mask), split_zeros as in the reference. */
int mafe_hpss_masks(mafe_ctx* ctx, const float* harm_dev, const float* perc_dev, int64_t n, float margin_h, float margin_p,
                    float power, int32_t split_zeros, float* mask_h_dev, float* mask_p_dev);

/* ---- WAV decode ("next" row f3: mindaudio/data/io.py:347-747 read / _fmt_chunk / _data_chunk / _skip_unknown_chunk) ---- */
enum {  /* mafe_wav_info.sample_kind: the container of one item of the data chunk (io.py:444-470) */
  MAFE_WAV_U8 = 1,   /* PCM, bit depth 1..8: unsigned bytes */
  MAFE_WAV_I8 = 2,
  MAFE_WAV_I16 = 3,
  MAFE_WAV_I24 = 4,  /* 3-byte container, returned left-justified in an int32 (io.py:505-512) */
  MAFE_WAV_I32 = 5,
  MAFE_WAV_I40 = 6,  /* 5/6/7-byte containers, left-justified in an int64 */
  MAFE_WAV_I48 = 7,
  MAFE_WAV_I56 = 8,
  MAFE_WAV_I64 = 9,
  MAFE_WAV_F32 = 10,
  MAFE_WAV_F64 = 11
};
enum { MAFE_WAV_OUT_F32 = 0, MAFE_WAV_OUT_F64 = 1, MAFE_WAV_OUT_I16 = 2 };
enum {  /* mafe_wav_info.warnings: the reference's WavFileWarning cases */
  MAFE_WAV_WARN_UNKNOWN_CHUNK = 1,  /* "Chunk (non-data) not understood, skipping it." (io.py:730-736) */
  MAFE_WAV_WARN_EOF = 2,            /* "Reached EOF prematurely" (io.py:684-691) */
  MAFE_WAV_WARN_INCOMPLETE_ID = 4   /* "Incomplete chunk ID ... ignoring it." (io.py:697-700) */
};
enum {  /* mafe_wav_info.error_kind when mafe_wav_parse fails: the exception class the reference raises */
  MAFE_WAV_ERR_VALUE = 1, MAFE_WAV_ERR_TYPE = 2, MAFE_WAV_ERR_UNBOUND = 3, MAFE_WAV_ERR_ZERODIV = 4,
  MAFE_WAV_ERR_STRUCT = 5,  /* struct.error: a size / header field cut short by the end of the file */
  MAFE_WAV_ERR_OS = 6       /* OSError: the file could not be opened / mapped (mafe_wav_files_open) */
};
typedef struct mafe_wav_info {
  int32_t format_tag;        /* 1 PCM, 3 IEEE float (WAVE_FORMAT_EXTENSIBLE resolved through its GUID, io.py:364-382) */
  int32_t channels;
  int32_t sample_rate;
  int32_t bytes_per_second;
  int32_t block_align;
  int32_t bit_depth;
  int32_t big_endian;        /* RIFX */
  int32_t sample_kind;       /* MAFE_WAV_* */
  int32_t bytes_per_sample;  /* block_align / channels */
  int32_t warnings;          /* MAFE_WAV_WARN_* bits */
  int32_t error_kind;        /* MAFE_WAV_ERR_* when the call failed */
  int32_t reserved;
  int64_t data_offset;       /* byte offset in the buffer of the first item `read` returns */
  int64_t n_items;           /* items `read` returns, over all channels */
  int64_t data_chunk_bytes;  /* the data chunk's size field */
} mafe_wav_info;
/* Host only, no device work: walk the RIFF/RIFX container held in bytes[0..n_bytes) exactly as `read(file, offset,
 * duration)` does (duration_s = 0 stands for None).  filelike != 0 selects the reference's path for file objects
 * without a C-level descriptor (io.BytesIO: read(size) of the whole chunk, `duration` ignored, io.py:500-503). */
int mafe_wav_parse(const void* bytes, int64_t n_bytes, double offset_s, double duration_s, int32_t filelike, mafe_wav_info* info);
/* Host only: walk n_files containers (blobs[k], blob_bytes[k]: the files' contents; path-like semantics, offset 0,
 * no duration) on n_threads host threads (<= 0: up to 16), fill infos[n_files] and the prefix sums of the payload sizes
 * payload_offsets[n_files + 1]; when stage != NULL also pack the payloads back to back into it (a pinned buffer of the
 * caller; stage_bytes >= payload_offsets[n_files]), so that one copy moves the whole batch to the device.  On a
 * malformed file: returns MAFE_E_INVALID_ARG, *failed_index = the first such file, infos[*failed_index].error_kind and
 * mafe_last_error() describe it. */
int mafe_wav_stage(const void* const* blobs, const int64_t* blob_bytes, int32_t n_files, int32_t n_threads,
                   mafe_wav_info* infos, int64_t* payload_offsets, void* stage, int64_t stage_bytes, int32_t* failed_index);
/* The same from paths (io.read's `open(file, "rb")` branch, io.py:646-647): the files are memory-mapped by the host
 * threads, walked like mafe_wav_stage (infos, payload_offsets as there), and kept mapped in *out until
 * mafe_wav_files_close; mafe_wav_files_pack copies the payloads from the page cache into the caller's pinned buffer. */
typedef struct mafe_wav_files mafe_wav_files;
int mafe_wav_files_open(const char* const* paths, int32_t n_files, int32_t n_threads, mafe_wav_files** out,
                        mafe_wav_info* infos, int64_t* payload_offsets, int32_t* failed_index);
int mafe_wav_files_pack(mafe_wav_files* files, void* stage, int64_t stage_bytes);
int mafe_wav_files_close(mafe_wav_files* files);
/* Device: n_items items of kind sample_kind at payload_dev (byte pointer, any alignment) -> out_dev as float32 / float64
 * in `read`'s unified output format (int16 / 32768, int32 and 24-bit / 2^31, everything else unchanged; io.py:741-746)
 * times `scale` (the conformer pipeline's `* (1 << 15)`, examples/conformer/dataset.py:389-390), or as raw int16 in
 * host byte order (MAFE_WAV_OUT_I16, PCM16 only: the front-end's MAFE_WAVE_I16 input). */
int mafe_wav_decode(mafe_ctx* ctx, const void* payload_dev, int64_t n_items, int32_t sample_kind, int32_t big_endian,
                    int32_t out_dtype, double scale, void* out_dev);

/* ---- Fourier-method resampling ("next" row f2: the `resample` step of augment.pitch_shift, mindaudio/data/augment.py:
 * 874-901 -> processing.resample res_type "fft" / "scipy", mindaudio/data/processing.py:132-186 -> scipy.signal.resample) ---- */
/* x_dev [rows][n_in] float64 -> out_dev [rows][n_out] float64 = irfft(rfft(x)[: m/2 + 1] (unpaired bin at m/2 doubled
 * when shrinking, halved when growing; m = min(n_in, n_out)), n_out) * n_out / n_in.  Lengths are arbitrary: both DFTs
 * run as Bluestein chirp-z transforms over a power-of-two FFT in complex128.  work_dev: scratch of at least
 * mafe_resample_workspace(...) bytes (256-byte aligned). */
int mafe_resample_workspace(int32_t rows, int64_t n_in, int64_t n_out, size_t* bytes);
int mafe_resample_fft(mafe_ctx* ctx, const double* x_dev, int32_t rows, int64_t n_in, int64_t n_out, double* out_dev,
                      void* work_dev, size_t work_bytes);

#ifdef __cplusplus
}
#endif
#endif /* MAFE_H_ */
