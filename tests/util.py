"""Shared helpers for the parity suites (test infrastructure)."""
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# North-star tolerances (BASELINE.json / BASELINE.md section 4):
#   log-mel / MFCC:  |d| <= 1e-4 * max(1, |ref|)
#   STFT bins:       max|dX| / max|X| <= 1e-5  (per utterance)
TOL_LOGMEL = 1e-4
TOL_STFT = 1e-5


def synth(seed, shape, scale=0.05):
    """Synthetic waveform of SURVEY.md section 8d: clip(0.05*N(0,1), -1, 1) float32."""
    rng = np.random.default_rng(seed)
    return np.clip(scale * rng.standard_normal(shape), -1.0, 1.0).astype(np.float32)


def mixed_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref)))) if ref.size else 0.0


# every log-mel comparison of a test session: (test id, elements, fraction outside the PLAIN north-star bound, max mixed error);
# tests/conftest.py writes it to gpurun_out/parity_report.json at the end of a GPU session
PARITY_REPORT = []
MAX_FRAC_OUTSIDE_PLAIN = 1e-4


def plain_violations(got, ref, tol=TOL_LOGMEL):
    """(fraction of elements with |d| > tol * max(1, |ref|), largest |d| / max(1, |ref|)): the north-star bound as written."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if not ref.size:
        return 0.0, 0.0
    e = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
    return float(np.mean(e > tol)), float(e.max())


def record_parity(kind, got, ref, tol=TOL_LOGMEL):
    frac, worst = plain_violations(got, ref, tol)
    PARITY_REPORT.append({"test": os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0], "kind": kind,
                          "elements": int(np.asarray(ref).size), "frac_outside_plain_bound": frac, "max_mixed_err": worst,
                          "bound": tol})
    return frac, worst


def logmel_err(got, ref, to_ln=1.0):
    """Log-mel criterion, returned as a ratio (pass when <= 1):

        |d| <= 1e-4 * max(1, |ref|)  +  2e-7 * sqrt(E_peak(t) / E_ref)

    The first term is the north-star tolerance.  The second is the FP32 quantisation floor of the
    FFT: every bin carries an absolute amplitude error of ~1e-7 of the frame's peak spectral
    amplitude (the STFT criterion allows 1e-5), which is visible in ln E only for mel bins more
    than ~60 dB below the frame's peak E_peak(t) (|d ln E| = 2 dA / A).  ``to_ln`` converts the
    feature unit to nepers (ln 10 / 10 for dB features).  Last axis = mel, second-to-last = frame
    for time-major [T, M] features; pass the array transposed otherwise."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if not ref.size:
        return 0.0
    # the plain bound first: at most MAX_FRAC_OUTSIDE_PLAIN of the elements may need the floor term at all
    frac, worst = record_parity("logmel", got, ref)
    assert frac <= MAX_FRAC_OUTSIDE_PLAIN, "%.3g of %d elements are outside |d| <= 1e-4 max(1, |ref|) (worst %.3g)" % (frac, ref.size, worst)
    ln_ref = ref * to_ln
    peak = np.max(ln_ref, axis=-1, keepdims=True)
    floor = 2e-7 * np.exp(0.5 * (peak - ln_ref)) / to_ln
    tol = TOL_LOGMEL * np.maximum(1.0, np.abs(ref)) + floor
    return float(np.max(np.abs(got - ref) / tol))


def mfcc_err(got, ref):
    """MFCC criterion: the DCT mixes every mel bin of a frame, so the error of one coefficient is
    judged against the frame's cepstral scale: max_t max_c |d[c,t]| / max(1, max_c |ref[c,t]|).
    (A per-coefficient relative error is ill-conditioned: coefficients cross zero while c0 ~ 500;
    an FP32 FFT leaves ~1e-3 dB on bins 70 dB below a frame's peak -- SURVEY.md App. C2.)"""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if not ref.size:
        return 0.0
    scale = np.maximum(1.0, np.max(np.abs(ref), axis=-2, keepdims=True))
    return float(np.max(np.abs(got - ref) / scale))


def stft_err(got, ref):
    ref = np.asarray(ref)
    return float(np.max(np.abs(np.asarray(got) - ref)) / max(np.max(np.abs(ref)), 1e-30)) if ref.size else 0.0


class Golden:
    """tests/golden/*.npz written by oracle/make_goldens.py (reference python executed in the dev
    container).  Long arrays are stored on a frame subset: ``take(name, full)`` applies the same
    subset to a freshly computed full-size result."""

    def __init__(self):
        self.files = {n[:-4]: np.load(os.path.join(GOLDEN_DIR, n)) for n in os.listdir(GOLDEN_DIR) if n.endswith(".npz")}

    def __getitem__(self, key):
        fname, name = key.split("/")
        return self.files[fname][name]

    def wav(self):
        """BAC009S0002W0122 as io.read returns it: float64 = int16 / 32768."""
        return self["spectrum/wav_i16"].astype(np.float64) / 32768.0

    def take(self, key, full):
        fname, name = key.split("/")
        f = self.files[fname]
        full = np.asarray(full)
        if name + "__cols" in f.files:
            assert tuple(f[name + "__shape"]) == full.shape, (key, tuple(f[name + "__shape"]), full.shape)
            return full[..., f[name + "__cols"]]
        return full
