"""fused bias + LeakyReLU -- mirrors models/stylegan2/op/fused_act.py (FusedLeakyReLU :73-82,
fused_leaky_relu :85-96) on top of hf_bias_act_f32 (include/hairfast_b200.h).

Differences from the reference, by design: forward only (swap() runs under torch.inference_mode,
models/Embedding.py:44) and CUDA tensors only -- a CPU tensor raises instead of taking a slow path.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import _lib


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (hairfastgan_b200 has no CPU fallback)")


def fused_bias_act(input: torch.Tensor, bias, refer=None, act: int = 3, grad: int = 0,
                   alpha: float = 0.2, scale: float = 2 ** 0.5) -> torch.Tensor:
    """Same argument order as the reference pybind op ``fused.fused_bias_act``
    (op/fused_bias_act.cpp:11-21); only grad == 0 (forward) is implemented."""
    if grad != 0:
        raise RuntimeError("fused_bias_act: only the forward (grad=0) is implemented")
    _require_cuda(input, "input")
    x = input.detach().contiguous().float()
    y = torch.empty_like(x)
    n = x.numel()
    size_b = 0
    bptr = None
    if bias is not None and bias.numel() > 0:
        _require_cuda(bias, "bias")
        b = bias.detach().contiguous().float()
        size_b = b.numel()
        bptr = b.data_ptr()
    step_b = 1
    for i in range(2, x.dim()):
        step_b *= x.size(i)
    _lib.use_device(x.device.index)
    _lib.check(_lib.lib().hf_bias_act_f32(x.data_ptr(), bptr, y.data_ptr(), n, size_b, step_b, act, float(alpha),
                                          float(scale), _lib.stream_ptr()), "hf_bias_act_f32")
    return y.to(input.dtype)


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel))
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    if input.device.type == "cpu" and input.dim() == 2:
        # The one CPU use the reference makes of this op at inference: Net.build_PCA_model (models/Net.py:50-55) moves
        # the mapping MLP to the CPU and pushes 1e6 z vectors through EqualLinear(activation='fused_lrelu')
        # (models/stylegan2/model.py:153-157) when <ckpt>_PCA.npz is missing -- an init-time one-off on [N, 512]
        # vectors, restated from the reference's own CPU branch (op/fused_act.py:86-93).  Feature maps (the hot path)
        # stay CUDA-only: a CPU image tensor still raises below.
        return torch.nn.functional.leaky_relu(input + bias.view(1, -1), negative_slope=negative_slope) * scale
    return fused_bias_act(input, bias, None, 3, 0, negative_slope, scale)
