"""Stub ``mindspore.nn``: only the grouped ``Conv1d`` that ``features.context_window`` builds
(``mindaudio/data/features.py:134-144``), evaluated literally with ``torch.conv1d``."""
import numpy as np


class Cell:
    def __call__(self, *a, **k):
        return self.construct(*a, **k)


class Conv1d(Cell):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, pad_mode="same",
                 padding=0, dilation=1, group=1, has_bias=False, weight_init=None, **_):
        assert pad_mode == "pad" and stride == 1 and dilation == 1 and not has_bias
        self.padding = padding
        self.group = group
        self.weight = np.asarray(weight_init.asnumpy(), dtype=np.float32)

    def construct(self, x):
        import torch
        from . import Tensor
        xt = torch.from_numpy(np.ascontiguousarray(x.asnumpy(), dtype=np.float32))
        wt = torch.from_numpy(np.ascontiguousarray(self.weight))
        y = torch.conv1d(xt, wt, padding=self.padding, groups=self.group)
        return Tensor(y.numpy())
