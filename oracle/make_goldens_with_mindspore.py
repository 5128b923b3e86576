"""Regenerate the MindSpore-backed goldens with the REAL ``mindspore.dataset.audio`` -- TEST INFRASTRUCTURE.

    pip install mindspore==2.3.0          # the reference's pin (requirements.txt:1, .github/workflows/ut_test.yaml:33-37)
    MAFE_REFERENCE_ROOT=/path/to/mindaudio python -m oracle.make_goldens_with_mindspore [--out tests/golden]

Why: ``spectrogram`` / ``melspectrogram`` / ``melscale`` / ``fbank`` / ``mfcc`` / ``compute_deltas`` /
``sliding_window_cmn`` / ``spectral_centroid`` compute inside the MindSpore 2.3.0 wheel (call sites
``mindaudio/data/spectrum.py:594-606, 673-694``, ``features.py:62, 191, 337``, ``processing.py:380-407``), which is neither
vendored under the reference tree nor installable in the build container (no network).  ``tests/golden/features_msop.npz``
was therefore frozen from the reference's python wrappers around a float64 RESTATEMENT of those ops (``oracle/ms_shim``,
``msop=1``): "parity unpinned vs the MindSpore binary".  This script closes that gap wherever MindSpore exists: it loads
the reference's ``io.py`` / ``spectrum.py`` / ``features.py`` / ``processing.py`` UNCHANGED on top of the real
``mindspore`` package, runs the very recipe of ``oracle/make_goldens.py::features_goldens`` (same keys, same inputs),
writes ``features_mindspore.npz`` (``msop=0``) and a diff report against the restated fixture
(``features_mindspore_diff.json``: per key the shape, max |d|, the mixed error |d| / max(1, |ref|) and the fraction of
elements outside the north-star 1e-4).  A key whose mixed error is <= 1e-5 pins the restatement; anything larger names the
op whose MindSpore arithmetic differs (expected candidates, SURVEY.md App. C1: float32 vs float64 mel filterbank / DCT
tables).  Copy ``features_mindspore.npz`` over ``features_msop.npz`` to make the GPU suite judge against the binary.

``--self-test`` runs the same code path on the restated shim (no MindSpore needed): the mechanics are exercised in the
build container by ``tests/test_oracle_golden.py::test_mindspore_golden_script_self_test``.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)


def _load_reference_modules(ref_root, names):
    """The reference's data modules executed as they are, by file path, on whatever ``mindspore`` is importable
    (``mindaudio/__init__.py`` is NOT executed: it imports the model zoo)."""
    if not hasattr(np, "float_"):
        np.float_ = np.float64            # spectrum.py:425,482 (numpy >= 2 removed the alias)
    pkg = types.ModuleType("mindaudio")
    pkg.__path__ = [os.path.join(ref_root, "mindaudio")]
    sub = types.ModuleType("mindaudio.data")
    sub.__path__ = [os.path.join(ref_root, "mindaudio", "data")]
    saved = {k: v for k, v in sys.modules.items() if k == "mindaudio" or k.startswith("mindaudio.")}
    for k in saved:
        del sys.modules[k]
    sys.modules["mindaudio"], sys.modules["mindaudio.data"] = pkg, sub
    mods = {}
    try:
        for name in names:
            full = "mindaudio.data." + name
            spec = importlib.util.spec_from_file_location(full, os.path.join(ref_root, "mindaudio", "data", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
            setattr(sub, name, mod)
            mods[name] = mod
    finally:
        for k in list(sys.modules):
            if k == "mindaudio" or k.startswith("mindaudio."):
                del sys.modules[k]
        sys.modules.update(saved)
    return mods


def extra_goldens(mods, x):
    """The remaining MindSpore-backed rows -- not in features_msop.npz."""
    ft, pr = mods["features"], mods.get("processing")
    e = {}
    fb = ft.fbank(x, n_mels=80, n_fft=400, hop_length=160)
    e["deltas_default"] = ft.compute_deltas(fb[:, :200])
    e["spectral_centroid"] = ft.spectral_centroid(x.astype(np.float32), 16000, 400, 400, 160, 0, "hann")
    if pr is not None:
        feats = np.ascontiguousarray(fb[:, :300].T)           # [frames, feats]
        e["sliding_cmn_600_100"] = pr.sliding_window_cmn(feats, 600, 100, False, False)
        e["sliding_cmn_center_var"] = pr.sliding_window_cmn(feats, 100, 20, True, True)
    return e


def diff_report(new, old):
    rep = {}
    for k in sorted(new):
        if k.endswith(("__cols", "__shape")) or k == "msop" or k not in old:
            continue
        a, b = np.asarray(new[k], dtype=np.float64), np.asarray(old[k], dtype=np.float64)
        if a.shape != b.shape:
            rep[k] = {"shape_mindspore": list(a.shape), "shape_restated": list(b.shape)}
            continue
        d = np.abs(a - b)
        mixed = d / np.maximum(1.0, np.abs(b))
        rep[k] = {"shape": list(a.shape), "max_abs": float(d.max()) if d.size else 0.0,
                  "max_mixed": float(mixed.max()) if d.size else 0.0,
                  "frac_outside_1e-4": float(np.mean(mixed > 1e-4)) if d.size else 0.0,
                  "pins_restatement": bool((mixed.max() if d.size else 0.0) <= 1e-5)}
    return rep


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "golden"))
    ap.add_argument("--self-test", action="store_true", help="run on the restated shim instead of MindSpore (mechanics only)")
    args = ap.parse_args(argv)
    sys.path.insert(0, REPO)
    from oracle import make_goldens as mg
    from oracle import ref_loader
    ref_root = ref_loader.REF_ROOT
    if not ref_loader.available():
        raise SystemExit("reference tree not found at %s (set MAFE_REFERENCE_ROOT)" % ref_root)
    if args.self_test:
        ref_loader._install_shim()
        names = ("io", "spectrum", "features")                # processing.py needs more of mindspore than the shim has
    else:
        try:
            import mindspore                                   # noqa: F401
        except ImportError as exc:
            raise SystemExit("mindspore is not importable (%s): install mindspore==2.3.0 (requirements.txt:1) or use "
                             "--self-test" % exc)
        if getattr(sys.modules["mindspore"], "__file__", "").startswith(os.path.join(HERE, "ms_shim")):
            raise SystemExit("the stub oracle/ms_shim/mindspore is on sys.path, not the real package")
        names = ("io", "spectrum", "features", "processing")
    mods = _load_reference_modules(ref_root, names)
    x, sr = mods["io"].read(ref_loader.sample_wav())
    assert sr == 16000 and x.shape == (95984,)
    f = mg.features_goldens(mods["spectrum"], mods["features"], x, msop=1 if args.self_test else 0)
    f.update(extra_goldens(mods, x))
    thinned = mg.thin(f)
    os.makedirs(args.out, exist_ok=True)
    name = "features_selftest" if args.self_test else "features_mindspore"
    np.savez_compressed(os.path.join(args.out, name + ".npz"), **thinned)
    old_path = os.path.join(REPO, "tests", "golden", "features_msop.npz")
    rep = diff_report(thinned, dict(np.load(old_path))) if os.path.isfile(old_path) else {}
    with open(os.path.join(args.out, name + "_diff.json"), "w") as fh:
        json.dump(rep, fh, indent=1)
    for k, v in rep.items():
        print("%-28s %s" % (k, json.dumps(v)))
    unpinned = [k for k, v in rep.items() if not v.get("pins_restatement", False)]
    print("keys compared: %d; restatement pinned (mixed <= 1e-5) on %d; differing: %s" % (len(rep), len(rep) - len(unpinned), unpinned))
    return rep


if __name__ == "__main__":
    main()
