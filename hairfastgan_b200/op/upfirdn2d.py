"""upfirdn2d -- mirrors models/stylegan2/op/upfirdn2d.py:145-156 on top of hf_upfirdn2d_f32.
Forward only, CUDA tensors only (see fused_act.py)."""
from __future__ import annotations

import torch

from .. import _lib


def upfirdn2d_op(input: torch.Tensor, kernel: torch.Tensor, up_x: int, up_y: int, down_x: int, down_y: int,
                 pad_x0: int, pad_x1: int, pad_y0: int, pad_y1: int) -> torch.Tensor:
    """Same signature as the reference pybind op ``upfirdn2d_op.upfirdn2d`` (op/upfirdn2d.cpp:12-22):
    input [major, H, W, minor] with minor == 1, returns [major, H', W', 1]."""
    if not input.is_cuda or not kernel.is_cuda:
        raise RuntimeError("upfirdn2d: input and kernel must be CUDA tensors (no CPU fallback)")
    if input.dim() != 4 or input.size(3) != 1:
        raise RuntimeError("upfirdn2d: expected input of shape [major, H, W, 1]")
    x = input.detach().contiguous().float()
    k = kernel.detach().contiguous().float()
    major, in_h, in_w, _ = x.shape
    kh, kw = k.shape
    out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) // down_y + 1
    out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) // down_x + 1
    y = torch.empty(major, max(out_h, 0), max(out_w, 0), 1, device=x.device, dtype=torch.float32)
    _lib.use_device(x.device.index)
    _lib.check(_lib.lib().hf_upfirdn2d_f32(x.data_ptr(), y.data_ptr(), k.data_ptr(), major, in_h, in_w, kh, kw,
                                           up_x, up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1,
                                           _lib.stream_ptr()), "hf_upfirdn2d_f32")
    return y.to(input.dtype)


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    batch, channel, in_h, in_w = input.shape
    out = upfirdn2d_op(input.reshape(-1, in_h, in_w, 1), kernel, up, up, down, down, pad[0], pad[1], pad[0], pad[1])
    return out.view(-1, channel, out.shape[1], out.shape[2])
