"""GPU parity: fbank / mfcc / deltas / context / conformer front-end / CMVN vs oracle and goldens."""
import os

import numpy as np
import pytest

from oracle import restated as R
from tests.util import TOL_LOGMEL, logmel_err, mfcc_err, mixed_err, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ma():
    import __graft_entry__ as entry
    entry.build()
    import mindaudio_b200
    return mindaudio_b200


def test_fbank_cfg1_and_ecapa(ma, golden):
    x = golden.wav()
    out = ma.fbank(x, n_mels=80, n_fft=400, hop_length=160)            # BASELINE.json configs[0]
    assert out.shape == (80, 600) and out.dtype == np.float64
    assert mixed_err(golden.take("features_msop/fbank_cfg1", out), golden["features_msop/fbank_cfg1"]) <= TOL_LOGMEL
    assert mixed_err(out, R.fbank(x, n_mels=80, n_fft=400, hop_length=160)) <= TOL_LOGMEL
    assert ma.fbanks(x, n_mels=80, n_fft=400, hop_length=160).shape == (80, 600)
    xe = synth(4, (4, 48000))
    out = ma.fbank(xe, deltas=False, n_mels=80, left_frames=0, right_frames=0, n_fft=400, hop_length=160)
    assert out.shape == (4, 80, 301) and out.dtype == np.float32
    assert mixed_err(golden.take("features_msop/fbank_ecapa_syn4", out), golden["features_msop/fbank_ecapa_syn4"]) <= TOL_LOGMEL
    # batch coupling of the top_db floor (spectrum.py:81-86): batched != per utterance when levels differ
    xs = xe * np.array([1.0, 1e-3, 1.0, 1.0], dtype=np.float32)[:, None]
    assert mixed_err(ma.fbank(xs, n_mels=80), R.fbank(xs, n_mels=80)) <= TOL_LOGMEL
    x3 = synth(6, (2, 3, 8000)) * np.array([1.0, 1e-3], dtype=np.float32)[:, None, None]
    assert mixed_err(ma.fbank(x3, n_mels=23), R.fbank(x3, n_mels=23)) <= TOL_LOGMEL       # 4-D mel: per batch item


def test_fbank_mfcc_deltas_context(ma, golden):
    xm = synth(11, (2, 16000))
    out = ma.fbank(xm, deltas=True, context=True)
    ref = golden["features_msop/fbank_default_dc_syn11"]
    assert out.shape == (2, 1320, 81)
    assert mixed_err(golden.take("features_msop/fbank_default_dc_syn11", out), ref) <= TOL_LOGMEL
    out = ma.mfcc(xm)
    assert out.shape == (2, 660, 81)                        # docstring's 101 frames is stale (features.py:326-329)
    assert mfcc_err(golden.take("features_msop/mfcc_default_syn11", out), golden["features_msop/mfcc_default_syn11"]) <= TOL_LOGMEL
    x = golden.wav()
    out = ma.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160)      # configs[3]
    assert out.shape == (40, 600)
    assert mfcc_err(golden.take("features_msop/mfcc_cfg4", out), golden["features_msop/mfcc_cfg4"]) <= TOL_LOGMEL
    out = ma.mfcc(x, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160, log_mels=True)
    assert mfcc_err(golden.take("features_msop/mfcc_cfg4_logmels", out), golden["features_msop/mfcc_cfg4_logmels"]) <= TOL_LOGMEL
    out = ma.mfcc(xm, norm="none", left_frames=2, right_frames=4)
    assert mfcc_err(out, R.mfcc(xm, norm="none", left_frames=2, right_frames=4)) <= TOL_LOGMEL
    x3 = synth(6, (2, 3, 8000))
    assert mfcc_err(ma.mfcc(x3), R.mfcc(x3)) <= TOL_LOGMEL                                # 4-D context route


@pytest.mark.parametrize("n_mels,n_mfcc", [(80, 40), (32, 13), (48, 20), (64, 64), (96, 39), (16, 16)])
def test_mfcc_tensor_core_sizes(ma, n_mels, n_mfcc):
    """MFCC on the 32-frame tile kernels (n_fft 400 / 2048) with n_mels % 16 == 0: the DCT runs on the tensor cores
    (dct_mma_kernel: tcgen05 kind::tf32, three passes over the hi / lo split) -- every K / 16 instantiation, coefficient
    counts that are not multiples of 8, ragged last tiles, the top_db clamp on and off."""
    x = synth(22, (5, 20000)) * np.array([1.0, 0.02, 1.0, 0.3, 1e-3], dtype=np.float32)[:, None]
    for n_fft in (400, 2048):
        kw = dict(deltas=False, context=False, n_fft=n_fft, n_mels=n_mels, n_mfcc=n_mfcc)
        assert mfcc_err(ma.mfcc(x, **kw), R.mfcc(x, **kw)) <= TOL_LOGMEL, n_fft
        assert mfcc_err(ma.mfcc(x[1, :4321], log_mels=True, **kw), R.mfcc(x[1, :4321], log_mels=True, **kw)) <= TOL_LOGMEL, n_fft


@pytest.mark.parametrize("n_fft,n_mels,n_mfcc", [(1024, 40, 13), (600, 64, 20), (2048, 128, 64), (400, 128, 128), (512, 23, 23)])
def test_mfcc_other_sizes(ma, n_fft, n_mels, n_mfcc):
    """MFCC through every DCT tiling (coefficients per thread 2/4/5/8/16) and generic-kernel tile sizes (2..16 frames)."""
    x = synth(21, (3, 12000)) * np.array([1.0, 0.02, 1.0], dtype=np.float32)[:, None]
    kw = dict(deltas=False, context=False, n_fft=n_fft, n_mels=n_mels, n_mfcc=n_mfcc)
    assert mfcc_err(ma.mfcc(x, **kw), R.mfcc(x, **kw)) <= TOL_LOGMEL
    assert mfcc_err(ma.mfcc(x[0], log_mels=True, **kw), R.mfcc(x[0], log_mels=True, **kw)) <= TOL_LOGMEL


def test_deltas_and_context_standalone(ma, golden):
    fb = R.fbank(golden.wav(), n_mels=80, n_fft=400, hop_length=160)
    d = ma.compute_deltas(fb[:, :100], win_length=7, pad_mode="reflect")
    assert mixed_err(golden.take("features_msop/deltas_syn", d), golden["features_msop/deltas_syn"]) <= 1e-5
    for mode in ("edge", "constant", "symmetric"):
        z = np.random.default_rng(3).standard_normal((2, 13, 40)).astype(np.float32)
        assert mixed_err(ma.compute_deltas(z, 5, mode), R.compute_deltas(z, 5, mode)) <= 1e-5
    c = ma.context_window(fb[:10, :60].astype(np.float32), 3, 5)
    assert c.shape == (1, 90, 60)            # reference quirk: 2-D input keeps the batch axis it added
    assert np.array_equal(c, golden["features_msop/context_3_5"])
    z = np.random.default_rng(1).standard_normal((3, 7, 50)).astype(np.float32)
    for l, r in ((3, 5), (4, 4), (5, 3), (0, 0), (5, 5), (0, 3), (2, 0)):
        assert np.array_equal(ma.context_window(z, l, r), R.context_window(z, l, r))
    z4 = np.random.default_rng(1).standard_normal((2, 3, 7, 20)).astype(np.float32)
    assert np.array_equal(ma.context_window(z4, 2, 3), R.context_window(z4, 2, 3))


@pytest.mark.parametrize("fast", [False, True])
def test_conformer_fbank_golden(ma, golden, fast):
    x = golden.wav() * (1 << 15)
    out = ma.compute_fbank_feats(x, 16000, 25, 10, 80, allow_fast_path=fast)
    assert out.shape == (598, 80) and out.dtype == np.float64
    assert mixed_err(out, golden["conformer_cmvn/conformer_fbank"]) <= TOL_LOGMEL
    for i, n in enumerate((16000, 23456, 400, 559, 560)):
        w = np.round(synth(3, (n,)) * 32768).astype(np.float64)
        got = ma.compute_fbank_feats(w, 16000, 25, 10, 80, allow_fast_path=fast)
        ref = golden["conformer_cmvn/conformer_syn3_%d" % i]
        assert got.shape == ref.shape and mixed_err(got, ref) <= TOL_LOGMEL
    assert ma.compute_fbank_feats(np.ones(399), 16000, 25, 10, 80, allow_fast_path=fast).shape == (0, 80)


@pytest.mark.parametrize("fast", [False, True])
def test_conformer_ragged_batch(ma, fast):
    from mindaudio_b200._engine import get_engine
    rng = np.random.default_rng(3)
    lens = [int(v) for v in rng.integers(16000, 320001, size=6)] + [400, 399, 0, 561, 100000]
    waves = [np.round(synth(100 + i, (n,)) * 32768).astype(np.float32) for i, n in enumerate(lens)]
    pipe = ma.FbankPipeline(cmvn=None, allow_fast_path=fast)
    assert pipe.plan.is_fast == fast
    so = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    out, fo = get_engine().run_frontend(pipe.plan, np.concatenate(waves), so)
    for u, w in enumerate(waves):
        ref = R.conformer_fbank(w.astype(np.float64)) if len(w) else np.zeros((0, 80))
        got = out[fo[u]:fo[u + 1]]
        assert got.shape == ref.shape
        assert logmel_err(got, ref) <= 1.0, u
        if ref.size:
            assert np.mean(np.abs(got - ref) > TOL_LOGMEL * np.maximum(1, np.abs(ref))) <= 1e-4   # floor term is rare
    # int16 staging gives the same features as float32 staging of the same integers
    i16 = np.concatenate(waves).astype(np.int16)
    out16, _ = get_engine().run_frontend(pipe.plan, i16, so)
    assert np.array_equal(out16, out)


def test_conformer_dither(ma):
    from mindaudio_b200._engine import get_engine
    w = np.round(synth(8, (16000,)) * 32768).astype(np.float32)
    for fast in (False, True):
        a = ma.compute_fbank_feats(w, 16000, 25, 10, 80, dither=1.0, seed=1234, allow_fast_path=fast)
        b = ma.compute_fbank_feats(w, 16000, 25, 10, 80, dither=1.0, seed=1234, allow_fast_path=fast)
        c = ma.compute_fbank_feats(w, 16000, 25, 10, 80, dither=1.0, seed=99, allow_fast_path=fast)
        assert np.array_equal(a, b) and not np.array_equal(a, c)         # deterministic, seeded
        ref = R.conformer_fbank(w.astype(np.float64), dither=1.0, seed=1234, utt_id=0)
        assert mixed_err(a, ref) <= TOL_LOGMEL
        off = ma.compute_fbank_feats(w, 16000, 25, 10, 80, dither=0.0, allow_fast_path=fast)
        assert mixed_err(off, R.conformer_fbank(w.astype(np.float64))) <= TOL_LOGMEL


def test_cmvn_family(ma, golden):
    g = lambda n: golden["conformer_cmvn/" + n]
    feats = [g("conformer_fbank")] + [g("conformer_syn3_%d" % i) for i in range(5)]
    stats = ma.compute_cmvn_stats(feats)
    assert stats.frame_num == int(g("cmvn_frame_num"))
    assert np.max(np.abs(stats.mean_stat - g("cmvn_mean_stat")) / np.abs(g("cmvn_mean_stat"))) <= 1e-6
    assert np.max(np.abs(stats.var_stat - g("cmvn_var_stat")) / np.abs(g("cmvn_var_stat"))) <= 1e-6
    mean, istd = stats.mean_istd()
    assert np.max(np.abs(mean - g("cmvn_mean"))) <= 1e-5 and np.max(np.abs(istd / g("cmvn_istd") - 1)) <= 1e-5
    out = ma.GlobalCMVN(g("cmvn_mean"), g("cmvn_istd"))(feats[0])
    assert out.dtype == np.float32 and mixed_err(out, g("global_cmvn_applied")) <= 1e-5
    out = ma.GlobalCMVN(g("cmvn_mean"), g("cmvn_istd"), norm_var=False)(feats[0][None])
    assert mixed_err(out[0], feats[0].astype(np.float32) - g("cmvn_mean").astype(np.float32)) <= 1e-5
    batch = np.stack([feats[0][:300], feats[0][298:598]])
    out = ma.InputNormalization(mean_norm=True, std_norm=False, norm_type="sentence").construct(batch.copy())
    assert mixed_err(out, g("utt_cmvn_mean_only")) <= 1e-5
    out = ma.InputNormalization(mean_norm=True, std_norm=True, norm_type="sentence")(batch.copy())
    assert mixed_err(out, g("utt_cmvn_mean_std")) <= 1e-5
    # ragged utterance CMVN with an offsets array (no padding)
    flat = np.concatenate(feats[:3])
    fo = np.concatenate([[0], np.cumsum([f.shape[0] for f in feats[:3]])])
    out = ma.utterance_cmvn(flat, fo)
    ref = np.concatenate([R.utt_cmvn(f) for f in feats[:3]])
    assert mixed_err(out, ref) <= 1e-5
    # DeepSpeech2 chain: stft 320/160 -> magphase -> log1p -> scalar norm (deepspeech2/dataset.py:36-47)
    spec = ma.stft(golden.wav(), n_fft=320, hop_length=160, win_length=320)
    mag, _ = ma.magphase(spec, power=1.0, iscomplex=True)
    out = ma.scalar_norm(mag)
    assert mixed_err(golden.take("spectrum/ds2_norm", out), golden["spectrum/ds2_norm"]) <= 1e-4
    import json, os, tempfile
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "global_cmvn")
        ma.save_cmvn_json(stats, p)
        assert set(json.load(open(p))) == {"mean_stat", "var_stat", "frame_num"}     # compute_cmvn_stats.py:121-128
        m2, i2 = ma.load_cmvn(p, True)
        assert np.allclose(m2, mean) and np.allclose(i2, istd)


def test_pipeline_with_utt_cmvn_matches_reference_chain(ma):
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(5)
    lens = [int(v) for v in rng.integers(16000, 80000, size=5)]
    waves = [np.round(synth(200 + i, (n,)) * 32768).astype(np.float32) for i, n in enumerate(lens)]
    pipe = ma.FbankPipeline(cmvn="utt", mean_norm=True, std_norm=True)
    wave = torch.from_numpy(np.concatenate(waves)).cuda()
    batch = pipe.layout(lens)
    out = pipe(wave, batch=batch)
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    fo = batch.frame_offsets
    for u, w in enumerate(waves):
        raw = R.conformer_fbank(w.astype(np.float64))
        assert logmel_err(out[fo[u]:fo[u + 1]] * raw.std(axis=0) + raw.mean(axis=0), raw) <= 2.0
        assert mixed_err(out[fo[u]:fo[u + 1]], R.utt_cmvn(raw)) <= 1e-3
    batch.close()


def test_host_pipeline_chunked_matches_oracle(ma):
    """mafe_frontend_run_host: chunked H2D / kernels / D2H on three streams (ragged, chunk boundaries, int16)."""
    rng = np.random.default_rng(9)
    lens = [int(v) for v in rng.integers(400, 60000, size=23)] + [399, 0, 400]
    waves = [np.round(synth(300 + i, (n,)) * 32768).astype(np.float32) for i, n in enumerate(lens)]
    for cmvn in (None, "utt"):
        pipe = ma.FbankPipeline(cmvn=cmvn)
        out, fo = pipe.features(waves, chunk_utts=4)
        assert fo[-1] == out.shape[0]
        for u, w in enumerate(waves):
            ref = R.conformer_fbank(w.astype(np.float64)) if len(w) else np.zeros((0, 80))
            got = out[fo[u]:fo[u + 1]]
            assert got.shape == ref.shape
            if cmvn and ref.shape[0] > 1:
                # CMVN divides by the per-bin std: judge the DE-normalised features with the log-mel criterion
                # (factor 2: the statistics carry the same FP32 rounding as the features)
                assert logmel_err(got * ref.std(axis=0) + ref.mean(axis=0), ref) <= 2.0, u
            elif not cmvn:
                assert logmel_err(got, ref) <= 1.0, u
        out16, fo16 = pipe.features([w.astype(np.int16) for w in waves], chunk_utts=7)
        assert np.array_equal(fo16, fo)
        if cmvn is None:
            assert np.array_equal(out16, out)                     # PCM16 staging is lossless for integer samples
        else:
            # chunking changes the atomic summation order; the in-kernel frame-mean sums group 8 (PCM16) or 4 (float32)
            # samples per thread, so the utterance mean differs in its last bits
            assert np.allclose(out16, out, atol=1e-4, equal_nan=True)


def test_many_utterances_take_the_per_utterance_prepass(ma):
    """>= 4 x SM-count utterances: the frame-mean pre-pass runs one CTA per utterance (frame_sum_utt_kernel)."""
    rng = np.random.default_rng(21)
    n = 640
    lens = [int(v) for v in rng.integers(400, 6000, size=n)]
    lens[5], lens[17], lens[100] = 399, 0, 52000
    waves = [np.round(synth(1000 + i, (m,)) * 32768).astype(np.float32) for i, m in enumerate(lens)]
    pipe = ma.FbankPipeline(cmvn=None)
    out, fo = pipe.features(waves, chunk_utts=n)      # one chunk: all utterances in one batch
    for u in list(range(0, n, 37)) + [5, 17, 100, n - 1]:
        w = waves[u]
        ref = R.conformer_fbank(w.astype(np.float64)) if len(w) else np.zeros((0, 80))
        got = out[fo[u]:fo[u + 1]]
        assert got.shape == ref.shape
        assert logmel_err(got, ref) <= 1.0, u


def test_frame_mean_sums_inside_the_persistent_kernel(ma):
    """Large batches with utterance CMVN: the frame-mean sums are accumulated by the persistent kernel itself (lagged sum
    duty, fbank512_v6.cuh FS) -- first / last tiles, one- to three-frame utterances, tile-aligned lengths, PCM16 input."""
    rng = np.random.default_rng(33)
    n = 900
    lens = [int(v) for v in rng.integers(400, 60000, size=n)]
    special = {3: 400, 11: 559, 12: 560, 13: 720, 14: 880, 40: 400 + 160 * 31, 41: 400 + 160 * 32, 42: 400 + 160 * 63,
               43: 400 + 160 * 64, 44: 399 + 160 * 32, 45: 0, 46: 399, 47: 120000}
    for k, v in special.items():
        lens[k] = v
    assert sum(max(0, (m - 400) // 160 + 1) for m in lens) / 32 > 12 * 148       # enough tiles for the fused sums
    waves = [np.round(synth(5000 + i, (m,)) * 32768).astype(np.float32) for i, m in enumerate(lens)]
    check = sorted(set(list(special) + list(range(0, n, 45)) + [n - 1]))
    refs = {u: (R.conformer_fbank(waves[u].astype(np.float64)) if lens[u] >= 400 else np.zeros((0, 80))) for u in check}
    pipe = ma.FbankPipeline(cmvn="utt")
    for dtype in (np.float32, np.int16):
        out, fo = pipe.features([w.astype(dtype) for w in waves], chunk_utts=n)
        for u in check:
            ref, got = refs[u], out[fo[u]:fo[u + 1]]
            assert got.shape == ref.shape, u
            if ref.shape[0] > 1:
                assert logmel_err(got * ref.std(axis=0) + ref.mean(axis=0), ref) <= 2.0, (u, lens[u], dtype)


def test_kernel_variants_agree(tmp_path):
    """The fused default kernels against their unfused / FP32 counterparts (A/B environment switches, read once per
    process -> subprocesses): frame-mean sums inside the persistent kernel vs the pre-pass kernel, utterance CMVN inside
    the kernel vs the apply kernel, per-lane constants in TMEM vs shared memory, tensor-core DCT vs FP32 DCT."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for name, env in (("default", {}), ("prepass", {"MAFE_NO_FUSED_FRAMESUM": "1"}), ("apply", {"MAFE_NO_FUSED_CMVN": "1"}),
                      ("smem", {"MAFE_NO_TMEM": "1"}), ("fp32dct", {"MAFE_DCT_TILED": "1"})):
        path = str(tmp_path / (name + ".npz"))
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, os.path.join(root, "tests", "variant_probe.py"), path], check=True, cwd=root, env=e, timeout=300)
        outs[name] = np.load(path)
    ref = outs["default"]
    ok = np.isfinite(ref["feats"])                                  # one-frame utterances normalise to nan / inf in every variant
    for name in ("prepass", "apply", "smem"):
        o = outs[name]
        assert np.array_equal(o["fo"], ref["fo"]), name
        assert np.array_equal(np.isfinite(o["feats"]), ok), name
        # normalised features: the variants differ in summation order (atomics, float vs double partial sums) only
        d = np.abs(o["feats"][ok] - ref["feats"][ok])
        assert np.mean(d > 1e-4) <= 1e-4 and d.max() <= 5e-3, (name, float(d.max()), float(np.mean(d > 1e-4)))
    d = np.abs(outs["fp32dct"]["mfcc"] - ref["mfcc"])
    assert d.max() <= 1e-4 * max(1.0, float(np.abs(ref["mfcc"]).max())), float(d.max())   # 3 x TF32 vs FP32 FMA


def test_gather_upload_and_staged_download(ma):
    """mafe_memcpy_h2d_gather / mafe_memcpy_d2h_staged (numpy-in / numpy-out plumbing): many pieces of ragged sizes incl.
    empty ones, single-threaded (small) and multi-threaded (>= 8 MiB) totals, back-to-back calls reusing the staging."""
    from mindaudio_b200._engine import get_engine
    eng = get_engine()
    rng = np.random.default_rng(12)
    for sizes in ([0, 5, 0, 1], [1000, 0, 3, 777777, 1], list(rng.integers(1, 300000, size=97)) + [0, 4 << 20]):
        arrs = [rng.integers(-32768, 32767, size=int(n)).astype(np.int16) for n in sizes]
        ref = np.concatenate(arrs) if arrs else np.zeros(0, np.int16)
        with eng.lock:
            dev = eng.buf("wave", max(ref.nbytes, 16))
            for _ in range(2):                      # the second call waits for the first upload before refilling the staging
                eng.h2d_gather(dev, arrs)
            back = np.empty_like(ref)
            if ref.nbytes:
                eng.d2h_staged(back, dev)
            eng.sync()
        assert np.array_equal(back, ref)
    # the public API through the large-array route (>= 4 MiB in and out)
    x = synth(31, (24, 48000))
    out = ma.stft(x, n_fft=512, hop_length=128)
    assert np.abs(out - R.stft(x, n_fft=512, hop_length=128)).max() <= 1e-5 * np.abs(out).max()


def test_pad_sequence_and_padded_pipeline(ma):
    """Scope row f3: pad_sequence on the device (bit exact vs the restated reference), and front-end + collate in one
    device round trip (xs_pad / xs_lengths / xs_masks of examples/conformer/dataset.py:563-569, 616-621)."""
    rng = np.random.default_rng(5)
    seqs = [rng.standard_normal((n, 80)).astype(np.float32) for n in (17, 3, 40, 1, 25)]
    for kw in (dict(), dict(padding_max_len=30), dict(batch_first=False, padding_max_len=48), dict(padding_max_len=2)):
        out = ma.pad_sequence(seqs, padding_value=0.0, atype=np.float32, **kw)
        ref = R.pad_sequence(seqs, padding_value=0.0, atype=np.float32, **kw)
        assert out.dtype == ref.dtype and out.shape == ref.shape and np.array_equal(out, ref)
    odd = [rng.standard_normal((n, 7)).astype(np.float32) for n in (5, 9)]       # dim % 4 != 0: scalar kernel
    assert np.array_equal(ma.pad_sequence(odd, padding_value=-2.5, atype=np.float32), R.pad_sequence(odd, padding_value=-2.5, atype=np.float32))
    labels = [rng.integers(0, 50, size=n).astype(np.int32) for n in (4, 9, 2)]   # host route, as the reference
    assert np.array_equal(ma.pad_sequence(labels, padding_value=-1, padding_max_len=10), R.pad_sequence(labels, padding_value=-1, padding_max_len=10))
    assert np.array_equal(ma.make_pad_mask(np.array([5, 3, 2]), 8), R.make_pad_mask(np.array([5, 3, 2]), 8))

    lens = [16000, 5361, 400, 30000, 8000]
    waves = [np.round(synth(70 + i, (n,)) * 32768).astype(np.float32) for i, n in enumerate(lens)]
    pipe = ma.FbankPipeline(cmvn=None)
    flat, fo = pipe.features(waves)
    feats = [flat[fo[i]:fo[i + 1]] for i in range(len(lens))]
    for max_len in (None, 120, 200):
        xs_pad, xs_lengths, xs_masks = pipe.features_padded(waves, max_len=max_len)
        ml = max_len if max_len is not None else max(f.shape[0] for f in feats)
        r_pad, r_len, r_mask = R.conformer_collate_x(feats, ml)
        assert np.array_equal(xs_pad, r_pad) and np.array_equal(xs_lengths, r_len) and np.array_equal(xs_masks, r_mask)
    # + SpecAugment between the front-end and the padding (dataset.py:560-569), same random draws as the reference
    import random
    conf = {"num_t_mask": 2, "num_f_mask": 2, "max_t": 50, "max_f": 10}
    xs_pad, _, _ = pipe.features_padded(waves, max_len=150, spec_aug_conf=conf, rng=random.Random(5))
    masked = R.spec_aug([f.copy() for f in feats], conf, random.Random(5))
    assert np.array_equal(xs_pad, R.conformer_collate_x(masked, 150)[0])


def test_feature_domain_ops(ma):
    """Scope row f4: sliding-window CMN, spectral centroid, SpecAugment / frequency / time masking."""
    import random
    rng = np.random.default_rng(11)
    x = (rng.standard_normal((3, 90, 40)) * 3.0 + 1.0).astype(np.float32)
    for kw in (dict(cmn_window=20, min_cmn_window=5), dict(cmn_window=15, center=True, norm_vars=True), dict(), dict(cmn_window=8, min_cmn_window=3, norm_vars=True)):
        out, ref = ma.sliding_window_cmn(x, **kw), R.sliding_window_cmn(x, **kw)
        assert out.dtype == ref.dtype and out.shape == ref.shape
        assert np.max(np.abs(out - ref)) <= 1e-5 * max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(ma.sliding_window_cmn(x[0, :, :3].astype(np.float64), 30, 10) - R.sliding_window_cmn(x[0, :, :3].astype(np.float64), 30, 10))) <= 1e-5

    w = synth(5, (2, 9000))
    for kw in (dict(), dict(n_fft=512, win_length=400, hop_length=160)):
        out, ref = ma.spectral_centroid(w, 16000, **kw), R.spectral_centroid(w, 16000, **kw)
        assert out.shape == ref.shape and np.max(np.abs(out - ref) / np.abs(ref)) <= 1e-4
    assert ma.spectral_centroid(w[0], 16000).shape == R.spectral_centroid(w[0], 16000).shape

    conf = {"num_t_mask": 2, "num_f_mask": 2, "max_t": 50, "max_f": 10}
    xs = [rng.standard_normal((n, 80)).astype(np.float32) + 5.0 for n in (120, 33, 400, 7)]
    a = R.spec_aug([v.copy() for v in xs], conf, random.Random(77))
    b = ma.spec_aug([v.copy() for v in xs], conf, random.Random(77))
    assert all(np.array_equal(u, v) for u, v in zip(a, b))                    # bit exact: same draws, same rectangles
    spec = np.abs(rng.standard_normal((2, 3, 64, 50))).astype(np.float32)
    assert np.array_equal(ma.frequencymasking(spec, frequency_mask_param=9, mask_start=5, mask_value=-1.0), R.mask_along_axis(spec, 9, 5, -1.0, -2))
    assert np.array_equal(ma.timemasking(spec[0, 0], frequency_mask_param=20, mask_start=30), R.mask_along_axis(spec[0, 0], 20, 30, 0.0, -1))
    m = ma.frequencymasking(spec, iid_masks=True, frequency_mask_param=9, rng=np.random.default_rng(1))
    changed = (m != spec).any(axis=-1)                                        # [2, 3, 64]: masked frequency rows
    assert changed.sum(axis=-1).max() <= 9 and m.shape == spec.shape
    with pytest.raises(ValueError):
        ma.frequencymasking(spec, frequency_mask_param=65)


def test_bench_size_batch_properties(ma):
    """At the bench's scale (thousands of ragged utterances in ONE launch) the oracle cannot follow; size-independent
    properties instead: (1) independence -- an utterance's features do not depend on its batch mates (a sample of them
    recomputed in a small batch agree to FP32 rounding: the frame mean is summed by another kernel there); (2) utterance CMVN leaves zero mean / unit variance per
    mel bin; (3) the frame count follows floor((L - 400) / 160) + 1 for every utterance."""
    torch = pytest.importorskip("torch")
    g = torch.Generator(device="cuda").manual_seed(3)
    rng = np.random.default_rng(3)
    n = 4096
    lens = rng.integers(16000, 160001, size=n).astype(np.int64)
    wave = torch.clamp(0.05 * torch.randn(int(lens.sum()), generator=g, device="cuda"), -1, 1) * 32768.0
    so = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=so[1:])
    raw_pipe, cmvn_pipe = ma.FbankPipeline(cmvn=None), ma.FbankPipeline(cmvn="utt")
    b = raw_pipe.layout(lens.tolist())
    out = raw_pipe(wave, batch=b)
    outn = cmvn_pipe(wave, batch=cmvn_pipe.layout(lens.tolist()))
    torch.cuda.synchronize()
    fo = b.frame_offsets
    assert np.array_equal(np.diff(fo), (lens - 400) // 160 + 1)
    pick = rng.choice(n, size=24, replace=False)
    sub = torch.cat([wave[so[u]:so[u + 1]] for u in pick])
    sub_out = raw_pipe(sub, lengths=[int(lens[u]) for u in pick])
    torch.cuda.synchronize()
    pos = 0
    for u in pick:
        t = int(fo[u + 1] - fo[u])
        assert logmel_err(out[fo[u]:fo[u + 1]].cpu().numpy(), sub_out[pos:pos + t].cpu().numpy().astype(np.float64)) <= 1.0, u
        pos += t
    for u in pick[:8]:
        x = outn[fo[u]:fo[u + 1]].double()
        assert float(x.mean(dim=0).abs().max()) < 1e-3 and float((x.std(dim=0, unbiased=False) - 1).abs().max()) < 1e-3, u
    assert bool(torch.isfinite(out).all()) and bool(torch.isfinite(outn).all())


def test_hpss_harmonic(ma):
    """Scope row f2: soft_mask / hpss / harmonic (features.py:438-559).  The median filters are exact selections; the
    soft masks are a handful of FP32 operations (1e-6)."""
    x = synth(17, (2, 9000))
    spec = R.stft(x, n_fft=512)
    for kw in (dict(), dict(kernel_size=(13, 7), margin=(1.0, 3.0)), dict(power=1.0, kernel_size=8), dict(mask=True, margin=2.0)):
        out, ref = ma.hpss(spec, **kw), R.hpss(spec, **kw)
        for u, v in zip(out, ref):
            assert u.shape == v.shape and u.dtype == v.dtype
            assert np.max(np.abs(u - v)) <= 2e-6 * max(1.0, np.max(np.abs(v)))
    mag = np.abs(spec[1]).astype(np.float32)
    out, ref = ma.hpss(mag, power=np.inf, mask=True), R.hpss(mag, power=np.inf, mask=True)
    assert all(u.dtype == v.dtype and np.array_equal(u, v) for u, v in zip(out, ref))
    m, mr = ma.soft_mask(mag, mag[::-1].copy(), power=2, split_zeros=True), R.soft_mask(mag, mag[::-1].copy(), power=2, split_zeros=True)
    assert m.dtype == mr.dtype and np.max(np.abs(m - mr)) <= 2e-6
    y, yr = ma.harmonic(x[0].astype(np.float64), margin=3.0), R.harmonic(x[0].astype(np.float64), margin=3.0)
    assert y.shape == yr.shape and y.dtype == yr.dtype and np.max(np.abs(y - yr)) <= 1e-5 * np.max(np.abs(yr))
    with pytest.raises(TypeError):
        ma.hpss(spec, margin=0.5)



def test_mel_and_features_one_transform(ma):
    """Extension mafe_frontend_run_aux (BASELINE configs[3] asks for melspectrogram AND mfcc of the same waveforms): the
    n_fft 400 kernel writes the mel energies beside their dB, so (mel, fbank / mfcc) come from ONE transform.  Both arrays
    must equal the separate reference-signature calls bit for bit (same kernel, same arithmetic) and meet the oracle;
    kernels without the second output (n_fft 512 / 2048 here) take the two-plan route with the same results."""
    rel = lambda a, b: float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
    x = synth(23, (3, 16000))
    for kw in (dict(n_mels=80, n_fft=400, hop_length=160), dict(n_mels=40), dict(n_mels=64, n_fft=512, hop_length=128),
               dict(n_mels=128, n_fft=2048, win_length=1200, hop_length=300)):
        mel, fb = ma.mel_and_fbank(x, **kw)
        mkw = dict(kw)
        ref_mel = ma.melspectrogram(x, **mkw)
        assert mel.shape == ref_mel.shape and mel.dtype == ref_mel.dtype and np.array_equal(mel, ref_mel), kw
        ref_fb = ma.fbank(x, **kw)
        assert fb.shape == ref_fb.shape and np.array_equal(fb, ref_fb), kw
        assert rel(mel, R.melspectrogram(x, **mkw)) <= 1e-5, kw
    for xs in (x, x[0], x.reshape(1, 3, 16000), x.astype(np.float64)):       # 2-D / 1-D / 3-D dB groups, float64 in and out
        mel, mf = ma.mel_and_mfcc(xs, n_mels=80, n_mfcc=40, hop_length=160)
        ref = ma.mfcc(xs, deltas=False, context=False, n_mels=80, n_mfcc=40, hop_length=160)
        assert mf.shape == ref.shape and mf.dtype == ref.dtype and np.array_equal(mf, ref)
        ref_mel = ma.melspectrogram(xs, n_mels=80, hop_length=160)
        assert mel.shape == ref_mel.shape and mel.dtype == ref_mel.dtype and np.array_equal(mel, ref_mel)
    mel, mf = ma.mel_and_mfcc(x, n_mels=40, n_mfcc=13, log_mels=True)
    assert np.array_equal(mf, ma.mfcc(x, deltas=False, context=False, n_mels=40, n_mfcc=13, log_mels=True))
    assert np.array_equal(mel, ma.melspectrogram(x, n_mels=40))
    with pytest.raises(ValueError):
        ma.mel_and_mfcc(x, n_mels=20, n_mfcc=21)


def test_padded_pipeline_reference_order(ma):
    """the reference collate sorts by frame count, longest first, BEFORE spec_aug / pad_sequence
    (examples/conformer/dataset.py:483-489).  sort_by_length=True must equal a call on the pre-sorted list (same seed),
    and return the order."""
    import random
    rng = np.random.default_rng(21)
    waves = [np.round(synth(100 + i, (int(n),)) * 32768).astype(np.float32) for i, n in enumerate(rng.integers(4000, 30000, size=6))]
    conf = {"num_t_mask": 2, "num_f_mask": 2, "max_t": 20, "max_f": 10}
    pipe = ma.FbankPipeline(cmvn="utt")
    xs, ln, mk, order = pipe.features_padded(waves, spec_aug_conf=conf, rng=random.Random(5), sort_by_length=True)
    frames = np.array([(len(w) - 400) // 160 + 1 for w in waves])
    assert np.array_equal(order, np.argsort(frames)[::-1]) and np.array_equal(ln, frames[order])
    xs2, ln2, mk2 = pipe.features_padded([waves[i] for i in order], spec_aug_conf=conf, rng=random.Random(5))
    assert np.array_equal(ln, ln2) and np.array_equal(mk, mk2) and np.allclose(xs, xs2, atol=2e-5)


def test_ds2_features_fused(ma, golden):
    """deepspeech2 front-end (examples/deepspeech2/dataset.py:36-47) through the fused output kind: the transform kernel
    writes log1p(|X|) itself (MAFE_OUT_POWER + MAFE_LOG_LN_PLUS), the scalar normalisation follows on the device.  Golden =
    the reference's own chain (stft -> magphase -> log1p -> (m - mean) / std) on the sample WAV."""
    x = golden.wav()
    out = ma.ds2_features(x)
    assert out.shape == (161, 600) and out.dtype == np.float32
    assert mixed_err(golden.take("spectrum/ds2_norm", out), golden["spectrum/ds2_norm"]) <= 1e-4
    raw = ma.ds2_features(x, normalize=False)
    ref = np.log1p(np.abs(R.stft(x, n_fft=320, hop_length=160, win_length=320)))
    assert np.max(np.abs(raw - ref)) <= 1e-5 * max(1.0, np.max(ref))
    for kw in (dict(n_fft=512, hop_length=128, win_length=400), dict(n_fft=2048, hop_length=300, win_length=1200), dict(n_fft=1024, hop_length=256, win_length=1024)):
        raw = ma.ds2_features(x, normalize=False, **kw)
        ref = np.log1p(np.abs(R.stft(x, **kw)))
        assert raw.shape == ref.shape and np.max(np.abs(raw - ref)) <= 1e-5 * max(1.0, np.max(ref)), kw
        nrm = ma.ds2_features(x, normalize=True, **kw)            # other transform kernels: normalisation as a post step
        assert np.max(np.abs(nrm - (ref - ref.mean()) / ref.std())) <= 1e-4, kw
    # a ragged batch in one launch: every utterance is normalised by its own moments
    ws = [x[:40000], x[1000:9000], x[:512], x[5:70005]]
    for kw in (dict(), dict(n_fft=512, hop_length=128, win_length=400)):
        outs = ma.ds2_features(ws, **kw)
        for w, o in zip(ws, outs):
            ref = np.log1p(np.abs(R.stft(w, **(kw or dict(n_fft=320, hop_length=160, win_length=320)))))
            assert o.shape == ref.shape and np.max(np.abs(o - (ref - ref.mean()) / ref.std())) <= 1e-4, (kw, len(w))
