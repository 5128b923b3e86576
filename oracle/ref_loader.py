"""Load the reference's OWN python for the hot path, unchanged -- TEST INFRASTRUCTURE.

Only works where ``/root/reference`` exists (the dev container).  Used by
``oracle/make_goldens.py`` and ``tests/test_oracle_vs_reference.py`` to pin
``oracle/restated.py``; never imported by the product or by ``-m gpu`` tests.

Mechanics: a stub ``mindspore`` package (``oracle/ms_shim``) on
``sys.path``, ``np.float_ = np.float64`` (numpy >= 2 removed it;
``spectrum.py:425,482`` use it), package shells for ``mindaudio`` /
``mindaudio.data`` in ``sys.modules`` so relative imports resolve without
executing ``mindaudio/__init__.py`` (which imports the model zoo), and
``ast``-extraction of the conformer front-end functions from
``examples/conformer/dataset.py`` (its module top imports
``mindspore.dataset.engine``).
"""
from __future__ import annotations

import ast
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("MAFE_REFERENCE_ROOT", "/root/reference")
_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ms_shim")
_cache = {}


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "mindaudio", "data", "spectrum.py"))


def _install_shim():
    import numpy as np
    if not hasattr(np, "float_"):
        np.float_ = np.float64
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo not in sys.path:
        sys.path.insert(0, repo)
    real = sys.modules.get("mindspore")
    if real is not None and not getattr(real, "__file__", "").startswith(_SHIM):
        raise RuntimeError("a real mindspore is already imported; run make_goldens_with_mindspore.py instead")
    if _SHIM not in sys.path:
        sys.path.insert(0, _SHIM)


def load_data_modules():
    """Returns (io, spectrum, features) modules executing the reference files as-is."""
    if "data" in _cache:
        return _cache["data"]
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    _install_shim()
    saved = {k: sys.modules.get(k) for k in ("mindaudio", "mindaudio.data")}
    pkg = types.ModuleType("mindaudio")
    pkg.__path__ = [os.path.join(REF_ROOT, "mindaudio")]
    sub = types.ModuleType("mindaudio.data")
    sub.__path__ = [os.path.join(REF_ROOT, "mindaudio", "data")]
    sys.modules["mindaudio"], sys.modules["mindaudio.data"] = pkg, sub
    mods = []
    try:
        for name in ("io", "spectrum", "features"):
            full = "mindaudio.data." + name
            spec = importlib.util.spec_from_file_location(
                full, os.path.join(REF_ROOT, "mindaudio", "data", name + ".py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            spec.loader.exec_module(mod)
            setattr(sub, name, mod)
            mods.append(mod)
    finally:
        # leave no 'mindaudio' shells behind: the product package may want that name
        for k in list(sys.modules):
            if k == "mindaudio" or k.startswith("mindaudio."):
                del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
    _cache["data"] = tuple(mods)
    return _cache["data"]


def _extract_functions(path, names, extra_globals=None):
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    keep = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    missing = set(names) - {n.name for n in keep}
    if missing:
        raise RuntimeError("not found in %s: %s" % (path, sorted(missing)))
    module = ast.Module(body=keep, type_ignores=[])
    import math
    import numpy as np
    g = {"np": np, "math": math, "__name__": "ref_extract"}
    g.update(extra_globals or {})
    exec(compile(module, path, "exec"), g)
    return g


def load_conformer_frontend():
    """The nine front-end functions of ``examples/conformer/dataset.py:56-168``, as written."""
    if "conf" not in _cache:
        names = ["inverse_mel_scale", "mel_scale", "mel_scale_scalar", "get_mel_banks", "preemphasis",
                 "enframe", "get_spectrum", "fbank", "compute_fbank_feats"]
        g = _extract_functions(os.path.join(REF_ROOT, "examples", "conformer", "dataset.py"), names)
        _cache["conf"] = types.SimpleNamespace(**{n: g[n] for n in names})
    return _cache["conf"]


def load_input_normalization():
    """``examples/ECAPA-TDNN/spec_augment.py:22-70`` class, as written."""
    if "inorm" not in _cache:
        g = _extract_functions(os.path.join(REF_ROOT, "examples", "ECAPA-TDNN", "spec_augment.py"),
                               ["InputNormalization"])
        _cache["inorm"] = g["InputNormalization"]
    return _cache["inorm"]


def load_json_cmvn():
    """``mindaudio/utils/load_files.py:9-29`` function, as written."""
    if "jc" not in _cache:
        import json
        g = _extract_functions(os.path.join(REF_ROOT, "mindaudio", "utils", "load_files.py"),
                               ["_load_json_cmvn"], {"json": json})
        _cache["jc"] = g["_load_json_cmvn"]
    return _cache["jc"]


def load_collate():
    """``pad_sequence`` (mindaudio/utils/common.py:10-52) and ``make_pad_mask`` (mindaudio/utils/mask.py:44-67), as written."""
    if "collate" not in _cache:
        from typing import List, Tuple
        g1 = _extract_functions(os.path.join(REF_ROOT, "mindaudio", "utils", "common.py"), ["pad_sequence"],
                                {"List": List, "Tuple": Tuple})
        g2 = _extract_functions(os.path.join(REF_ROOT, "mindaudio", "utils", "mask.py"), ["make_pad_mask"], {"List": List})
        _cache["collate"] = types.SimpleNamespace(pad_sequence=g1["pad_sequence"], make_pad_mask=g2["make_pad_mask"])
    return _cache["collate"]


def load_spec_aug():
    """The ``spec_aug`` METHOD of the conformer collate class (``examples/conformer/dataset.py:493-534``), as written,
    exec'd as a plain function (``self`` is unused); it draws from the ``random`` MODULE."""
    if "specaug" not in _cache:
        import random
        path = os.path.join(REF_ROOT, "examples", "conformer", "dataset.py")
        with open(path) as fh:
            tree = ast.parse(fh.read(), filename=path)
        fn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "spec_aug"]
        if not fn:
            raise RuntimeError("spec_aug not found in %s" % path)
        g = {"random": random, "__name__": "ref_extract"}
        exec(compile(ast.Module(body=[fn[0]], type_ignores=[]), path, "exec"), g)
        _cache["specaug"] = g["spec_aug"]
    return _cache["specaug"]


def load_phase_vocoder():
    """``_phase_vocoder`` (mindaudio/data/augment.py:828-871), as written."""
    if "pvoc" not in _cache:
        g = _extract_functions(os.path.join(REF_ROOT, "mindaudio", "data", "augment.py"), ["_phase_vocoder"])
        _cache["pvoc"] = g["_phase_vocoder"]
    return _cache["pvoc"]


def load_pitch_shift():
    """``time_stretch`` / ``_phase_vocoder`` / ``pitch_shift`` (mindaudio/data/augment.py:795-901) and ``resample``
    (mindaudio/data/processing.py:132-186), as written, wired to the reference's own ``stft`` / ``istft`` / ``_pad_shape``."""
    if "pitch" not in _cache:
        import scipy
        import scipy.signal
        _, sp, _ = load_data_modules()
        gr = _extract_functions(os.path.join(REF_ROOT, "mindaudio", "data", "processing.py"), ["resample"],
                                {"scipy": scipy})
        g = _extract_functions(os.path.join(REF_ROOT, "mindaudio", "data", "augment.py"),
                               ["time_stretch", "_phase_vocoder", "pitch_shift"],
                               {"stft": sp.stft, "istft": sp.istft, "_pad_shape": sp._pad_shape, "resample": gr["resample"]})
        _cache["pitch"] = types.SimpleNamespace(resample=gr["resample"], pitch_shift=g["pitch_shift"], time_stretch=g["time_stretch"])
    return _cache["pitch"]


def sample_wav(name="BAC009S0002W0122.wav"):
    return os.path.join(REF_ROOT, "tests", "samples", "ASR", name)
