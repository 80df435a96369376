"""``BicubicDownSample`` (stage glue): drop-in counterpart of utils/bicubic.py:6-78 -- the separable
bicubic decimation HairFast applies to every 1024^2 image (Embedding.py:36-37,66-67; Blending.py:33,64).

Same constructor and ``forward`` signature as the reference class, but the work is ONE fused CUDA kernel
(``hf_bicubic_downsample_f32``: reflect padding by index arithmetic, vertical pass into shared memory, horizontal
pass out of it) instead of two ``F.pad`` + grouped ``conv2d`` round trips.  CUDA tensors only.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib

__all__ = ["BicubicDownSample"]


def _keys_weight(t: torch.Tensor, a: float) -> torch.Tensor:
    """Keys cubic-convolution weight w(t) for |t| < 2, vectorised (bicubic.py:7-18 evaluates it one tap at a time)."""
    t = t.abs()
    inner = (a + 2.) * t ** 3 - (a + 3.) * t ** 2 + 1
    outer = a * t ** 3 - 5. * a * t ** 2 + 8. * a * t - 4. * a
    return torch.where(t <= 1., inner, torch.where(t < 2., outer, torch.zeros_like(t)))


class BicubicDownSample(nn.Module):
    def __init__(self, factor=4, cuda=True, padding='reflect'):
        super().__init__()
        if padding != 'reflect':
            raise NotImplementedError("BicubicDownSample: only padding='reflect' (the HairFast setting) is implemented")
        self.factor = factor
        self.padding = padding
        n_taps = 4 * factor
        centres = (torch.arange(n_taps, dtype=torch.float32) - float(n_taps // 2) + 0.5) / factor
        taps = _keys_weight(centres, -0.5)
        self.k = taps / taps.sum()                       # bit-identical to the reference's k1 / k2 (tests/test_glue.py)
        self._k_dev = {}

    def bicubic_kernel(self, x, a=-0.50):
        return _keys_weight(torch.as_tensor(x, dtype=torch.float32), a)

    def forward(self, x, nhwc=False, clip_round=False, byte_output=False):
        if not x.is_cuda:
            raise RuntimeError("BicubicDownSample: input must be a CUDA tensor (no CPU fallback)")
        img = x.permute(0, 3, 1, 2) if nhwc else x
        img = img.detach().float().contiguous()
        b, c, h, w = img.shape
        f = self.factor
        taps = self._k_dev.get(x.device)
        if taps is None:
            taps = self._k_dev[x.device] = self.k.to(x.device)
        # output size of a `4f`-tap, stride-`f` filter over the image padded by `3f` in total
        ho, wo = (h - f) // f + 1, (w - f) // f + 1
        y = torch.empty(b, c, ho, wo, device=x.device, dtype=torch.float32)
        _lib.use_device(x.device.index)
        _lib.check(_lib.lib().hf_bicubic_downsample_f32(img.data_ptr(), taps.data_ptr(), y.data_ptr(), b * c, h, w, f,
                                                        int(bool(clip_round)), _lib.stream_ptr()),
                   "hf_bicubic_downsample_f32")
        out = y.permute(0, 2, 3, 1) if nhwc else y
        # the reference's `x.type('torch.ByteTensor')` (bicubic.py:74) yields a CPU uint8 tensor
        return out.to("cpu", torch.uint8) if byte_output else out
