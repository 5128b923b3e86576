"""Stub ``mindspore`` package -- TEST INFRASTRUCTURE (oracle/).

Just enough surface for ``/root/reference/mindaudio/data/{io,spectrum,features}.py``
to import and run UNCHANGED in a container without MindSpore.  The dataset audio
ops delegate to ``oracle.restated`` (float64 numpy restatement of the
mindspore==2.3.0 C++ ops, parity unpinned vs the binary -- see oracle/__init__.py).
"""
import numpy as np

from . import nn  # noqa: F401
from . import dataset  # noqa: F401

float32 = np.float32
float64 = np.float64
int32 = np.int32


class Tensor:
    """Minimal stand-in for ``mindspore.Tensor`` (numpy-backed)."""

    def __init__(self, data, dtype=None):
        if isinstance(data, Tensor):
            data = data._a
        self._a = np.asarray(data, dtype=dtype)

    @property
    def shape(self):
        return self._a.shape

    def asnumpy(self):
        return self._a

    def reshape(self, shape):
        return Tensor(self._a.reshape(shape))

    def transpose(self, axes):
        return Tensor(self._a.transpose(axes))

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)
