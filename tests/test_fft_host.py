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


def test_front2048_decomposition_and_layouts():
    """front2048.cuh on the host (numpy): the 2048 = 16 x 16 x 8 decomposition with the twiddles the kernel keeps in TMEM
    reproduces the DFT; the three shared-memory layouts are injective inside the 2176-slot scratch and conflict free for the
    8-byte accesses of a half-warp; the compile-time store offsets (k + k / 16 = k1 + 17 k2 + 272 k3) and the pair separation
    hold; the staged part of a frame covers the window's support."""
    import numpy as np
    N = 2048
    rng = np.random.default_rng(0)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    W = lambda n, e: np.exp(-2j * np.pi * (e % n) / n)
    # stage A: thread t = 8 n2 + n3 transforms x[128 n1 + t] over n1, times W2048^(t k1)
    t = np.arange(128)
    A = np.fft.fft(x.reshape(16, 128), axis=0) * W(N, np.outer(np.arange(16), t))            # [k1][t]
    # stage B: thread (k1, n3) transforms A[k1][8 n2 + n3] over n2, times W128^(n3 k2)
    B = np.fft.fft(A.reshape(16, 16, 8), axis=1) * W(128, np.outer(np.arange(16), np.arange(8)))[None]   # [k1][k2][n3]
    # stage C: 8-point DFT over n3 -> Z[k1 + 16 k2 + 256 k3]
    Cc = np.fft.fft(B, axis=2)                                                                # [k1][k2][k3]
    Z = np.empty(N, dtype=complex)
    k1, k2, k3 = np.meshgrid(np.arange(16), np.arange(16), np.arange(8), indexing="ij")
    Z[k1 + 16 * k2 + 256 * k3] = Cc
    assert np.max(np.abs(Z - np.fft.fft(x))) <= 1e-9 * np.max(np.abs(Z))
    # pair separation: a + i b -> X_a = (Z[k] + conj Z[N-k]) / 2, X_b = -i (Z[k] - conj Z[N-k]) / 2
    a, b = x.real, x.imag
    k = np.arange(1025)
    zn = np.conj(Z[(N - k) % N])
    assert np.allclose((Z[k] + zn) / 2, np.fft.rfft(a)) and np.allclose(-1j * (Z[k] - zn) / 2, np.fft.rfft(b))

    def conflict_free(addr):   # 8-byte slots: a half-warp (16 consecutive threads) must hit 16 different bank pairs
        addr = np.asarray(addr).reshape(-1, 16)
        return all(len(set(row % 16)) == 16 for row in addr)
    # layout 1 (A -> B): store [k1][t] at k1 * 136 + t, load thread (k1 = t >> 3, n3 = t & 7) at k1 * 136 + 8 n2 + n3
    lay1 = (np.arange(16)[:, None] * 136 + t[None]).ravel()
    assert len(set(lay1)) == lay1.size and lay1.max() < 2176
    for kk in range(16):
        assert conflict_free(kk * 136 + t)
    for n2 in range(16):
        assert conflict_free((t >> 3) * 136 + 8 * n2 + (t & 7))
    # layout 2 (B -> C): store thread (k1, n3) at n3 * 257 + 16 k1 + (k2 ^ 8 (k1 & 1)) through the two bases dlo / dhi
    lay2 = set()
    for kk2 in range(16):
        sw = ((t >> 3) & 1) << 3
        dlo, dhi = (t & 7) * 257 + (t >> 3) * 16 + sw, (t & 7) * 257 + (t >> 3) * 16 + (sw ^ 8)
        addr = dlo + kk2 if kk2 < 8 else dhi + (kk2 - 8)
        assert np.array_equal(addr, (t & 7) * 257 + (t >> 3) * 16 + (kk2 ^ sw))
        assert conflict_free(addr)
        lay2 |= set(addr)
    assert len(lay2) == 2048 and max(lay2) < 2176
    for n3 in range(8):                          # stage C loads: combos (k1a = t >> 4, k2 = t & 15) and k1b = k1a + 8
        for k1x in ((t >> 4), (t >> 4) + 8):
            assert conflict_free(n3 * 257 + k1x * 16 + ((t & 15) ^ ((k1x & 1) << 3)))
    # layout 3 (Z, skewed k + k / 16): stores from one base with compile-time offsets
    kk = np.arange(N)
    skew = kk + (kk >> 4)
    assert len(set(skew)) == N and skew.max() < 2176
    for kk3 in range(8):
        for plus8 in (0, 8):
            ka = (t >> 4) + plus8 + 16 * (t & 15) + 256 * kk3
            addr = (t >> 4) + 17 * (t & 15) + 272 * kk3 + plus8
            assert np.array_equal(addr, ka + (ka >> 4)) and conflict_free(addr)
    for i in range(8):                           # pair-separation loads: bin t + 128 i and its partner
        kb = t + 128 * i
        assert conflict_free(kb + (kb >> 4))
    # staged part of a frame: rows [j0, j1) of 128 samples must cover the centred window's support
    for win in (1200, 2048, 1000, 900, 1):
        lo_w = (N - win) // 2
        j0, j1 = lo_w // 128, (lo_w + win + 127) // 128
        assert 128 * j0 <= lo_w and lo_w + win <= 128 * j1 and 0 <= j0 < j1 <= 16
    assert ((N - 1200) // 2 // 128, ((N - 1200) // 2 + 1200 + 127) // 128) == (3, 13)      # the <3, 13> instantiation
