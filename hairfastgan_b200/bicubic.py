"""``BicubicDownSample`` (stage glue) -- drop-in counterpart of utils/bicubic.py:6-78, the separable
bicubic decimation HairFast applies to every 1024^2 image (Embedding.py:36-37,66-67; Blending.py:33,64).  Same constructor
and ``forward`` signature; one fused CUDA kernel (both 1-D passes through shared memory) instead of two padded grouped
``conv2d`` calls.  CUDA tensors only."""
from __future__ import annotations

import torch
from torch import nn

from . import _lib

__all__ = ["BicubicDownSample"]


class BicubicDownSample(nn.Module):
    def bicubic_kernel(self, x, a=-0.50):
        """Keys cubic convolution kernel (utils/bicubic.py:7-18)."""
        abs_x = torch.abs(x)
        if abs_x <= 1.:
            return (a + 2.) * torch.pow(abs_x, 3.) - (a + 3.) * torch.pow(abs_x, 2.) + 1
        if 1. < abs_x < 2.:
            return a * torch.pow(abs_x, 3) - 5. * a * torch.pow(abs_x, 2.) + 8. * a * abs_x - 4. * a
        return 0.0

    def __init__(self, factor=4, cuda=True, padding='reflect'):
        super().__init__()
        if padding != 'reflect':
            raise NotImplementedError("BicubicDownSample: only padding='reflect' (the HairFast setting) is implemented")
        self.factor = factor
        size = factor * 4
        k = torch.tensor([self.bicubic_kernel((i - torch.floor(torch.tensor(size / 2)) + 0.5) / factor)
                          for i in range(size)], dtype=torch.float32)
        self.k = k / torch.sum(k)                       # the 1-D taps behind the reference's k1 / k2 (bicubic.py:24-31)
        self.cuda = '.cuda' if cuda else ''
        self.padding = padding
        self._k_dev = {}

    def forward(self, x, nhwc=False, clip_round=False, byte_output=False):
        if not x.is_cuda:
            raise RuntimeError("BicubicDownSample: input must be a CUDA tensor (no CPU fallback)")
        if nhwc:
            x = x.permute(0, 3, 1, 2)
        xf = x.detach().float().contiguous()
        b, c, h, w = xf.shape
        f = self.factor
        k = self._k_dev.get(x.device)
        if k is None:
            k = self._k_dev[x.device] = self.k.to(x.device)
        pad = 4 * f - f
        ho, wo = (h + pad - 4 * f) // f + 1, (w + pad - 4 * f) // f + 1
        y = torch.empty(b, c, ho, wo, device=x.device, dtype=torch.float32)
        _lib.use_device(x.device.index)
        _lib.check(_lib.lib().hf_bicubic_downsample_f32(xf.data_ptr(), k.data_ptr(), y.data_ptr(), b * c, h, w, f,
                                                        1 if clip_round else 0, _lib.stream_ptr()),
                   "hf_bicubic_downsample_f32")
        if nhwc:
            y = y.permute(0, 2, 3, 1)
        if byte_output:
            return y.type('torch.ByteTensor')           # like the reference (bicubic.py:74): a CPU uint8 tensor
        return y
