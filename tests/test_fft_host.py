"""The in-register FFT building blocks (csrc/fft512.cuh, csrc/fft400.cuh) are __host__ __device__: the index maps and
butterflies the kernels use are compiled for the host and checked against a float64 DFT (no GPU needed)."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.parametrize("name", ["fft512_host_check", "fft400_host_check", "fft320_host_check", "resample_host_check"])
def test_fft_maps_on_host(name, tmp_path):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    exe = tmp_path / name
    subprocess.run([NVCC, "-O1", "-std=c++17", "-o", str(exe), os.path.join(HERE, "host", name + ".cu")],
                   check=True, capture_output=True, timeout=300)
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stdout + res.stderr


def test_sweep_program_replayed_on_host(tmp_path):
    """The mel sweep PROGRAM of the headline kernel (per-bin weights, retire counts and masks, cost-balanced warp ranges,
    plane rows, two-row combine table; fbank512.cu build_bins / build_v3_program) is replayed on the host the way
    sweep_v3 and phase C execute it, and must reproduce the dense filterbank product -- for the conformer's Kaldi bank,
    an HTK bank and a Slaney-normalised bank; a bank with three filters on one bin must be refused."""
    import numpy as np
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    from mindaudio_b200 import _tables as T
    from mindaudio_b200.build import build
    build()
    libdir = os.path.join(os.path.dirname(HERE), "mindaudio_b200")
    exe = tmp_path / "sweep_program_check"
    subprocess.run([NVCC, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(exe),
                    os.path.join(HERE, "host", "sweep_program_check.cu"), "-L", libdir, "-lmafe",
                    "-Xlinker", "-rpath", "-Xlinker", libdir], check=True, capture_output=True, timeout=600)
    banks = {
        "kaldi_20_8000": T.kaldi_triangle_bank(80, 512, 16000, 20.0, 8000.0),
        "kaldi_0_7600": T.kaldi_triangle_bank(80, 512, 16000, 0.0, 7600.0),
        "htk": T.hz_triangle_bank(257, 80, 16000, 0.0, 8000.0),
        "slaney": T.hz_triangle_bank(257, 80, 16000, 60.0, 7800.0, norm="slaney", mel_type="slaney"),
    }
    for name, fb in banks.items():
        fb = np.ascontiguousarray(fb, dtype=np.float32)
        assert fb.shape == (80, 257), name
        path = tmp_path / (name + ".f32")
        fb.tofile(path)
        res = subprocess.run([str(exe), str(path)], capture_output=True, text=True, timeout=60)
        assert res.returncode == 0, name + ": " + res.stdout + res.stderr
    # the one-pass program of fbank400_kernel (n_fft 400: features.fbank / mfcc / melspectrogram defaults)
    for name, nm, fb in (("f400_htk80", 80, T.hz_triangle_bank(201, 80, 16000, 0.0, 8000.0)),
                         ("f400_htk40", 40, T.hz_triangle_bank(201, 40, 16000, 0.0, 8000.0)),
                         ("f400_htk23", 23, T.hz_triangle_bank(201, 23, 16000, 0.0, 8000.0)),
                         ("f400_slaney128", 128, T.hz_triangle_bank(201, 128, 16000, 0.0, 8000.0, norm="slaney", mel_type="slaney"))):
        fb = np.ascontiguousarray(fb, dtype=np.float32)
        assert fb.shape == (nm, 201), name
        path = tmp_path / (name + ".f32")
        fb.tofile(path)
        res = subprocess.run([str(exe), str(path), str(nm)], capture_output=True, text=True, timeout=60)
        assert res.returncode in (0, 3), name + ": " + res.stdout + res.stderr      # 3 = refused -> generic kernel, never wrong
        if name in ("f400_htk80", "f400_htk40", "f400_htk23"):
            assert res.returncode == 0, name + ": " + res.stdout + res.stderr
    wide = np.ascontiguousarray(banks["htk"], dtype=np.float32).copy()
    wide[10, 100] = wide[11, 100] = wide[12, 100] = 0.3                      # three filters on one bin
    path = tmp_path / "wide.f32"
    wide.tofile(path)
    res = subprocess.run([str(exe), str(path)], capture_output=True, text=True, timeout=60)
    assert res.returncode == 3, res.stdout + res.stderr
