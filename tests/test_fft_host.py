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
