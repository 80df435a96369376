"""F-space glue of the Alignment / Embedding stages: the mask algebra and the chained latent blends that
the reference writes inline as ~15 small torch ops on [1,512,32,32] tensors, as two kernels.

Reference lines and the call that replaces them (INTEGRATION.md section 3 shows the two-line edit):

* models/Alignment.py:139-144   ``masks = torch.cat([1 - (1-hm1)*(1-hmx), hmx, hm2*hmx])``  ->  ``align_masks(hm1, hm2, hmx)``
* models/Alignment.py:153-159   ``interpolation_low = 1 - F.interpolate(free_mask.float(), (32,32), 'bicubic')`` and the
  three lerps                                                                 ->  ``align_f_space(...)``
* models/Embedding.py:86-92     ``latent_F + mixing * bicubic32(hair_mask) * (latent_F_from_W - latent_F)``
                                                                              ->  ``mix_f_space(...)``
fp32, CUDA tensors only, forward only (the stages run under torch.inference_mode).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

__all__ = ["align_masks", "align_f_space", "mix_f_space", "fspace_blend"]


def _f32(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError("fspace: tensors must be CUDA tensors (no CPU fallback)")
    return t.detach().float().contiguous()


def align_masks(hair_mask1: torch.Tensor, hair_mask2: torch.Tensor, hair_mask_target: torch.Tensor) -> torch.Tensor:
    """[1,1,H,W] x3 (0/1 masks) -> [3,1,H,W]: Alignment.py:139-144."""
    a, b, x = _f32(hair_mask1), _f32(hair_mask2), _f32(hair_mask_target)
    if not (a.shape == b.shape == x.shape) or a.dim() != 4 or a.shape[0] != 1 or a.shape[1] != 1:
        raise RuntimeError(f"align_masks: expected three [1,1,H,W] masks, got {list(a.shape)}, {list(b.shape)}, "
                           f"{list(x.shape)}")
    out = torch.empty(3, 1, a.shape[2], a.shape[3], device=a.device, dtype=torch.float32)
    _lib.use_device(a.device.index)
    _lib.check(_lib.lib().hf_align_masks_f32(a.data_ptr(), b.data_ptr(), x.data_ptr(), out.data_ptr(),
                                             a.numel(), _lib.stream_ptr()), "hf_align_masks_f32")
    return out


def fspace_blend(first: torch.Tensor, stages, size_check: bool = True) -> torch.Tensor:
    """``stages`` = [(src [1,C,h,w], mask [Hm,Wm]-like, scale_a, scale_b), ...]:
    F = first; for each stage: w = scale_a + scale_b * bicubic(mask -> (h,w)); F = src + w * (F - src)."""
    f = _f32(first)
    if f.dim() != 4 or f.shape[0] != 1:
        raise RuntimeError(f"fspace_blend: expected [1,C,h,w], got {list(f.shape)}")
    _, ch, h, w = f.shape
    n = len(stages)
    if not 1 <= n <= 4:
        raise RuntimeError("fspace_blend: 1..4 stages")
    keep, srcs, masks = [], (C.c_void_p * n)(), (C.c_void_p * n)()
    sa, sb = (C.c_float * n)(), (C.c_float * n)()
    hm = wm = None
    for i, (src, mask, a, b) in enumerate(stages):
        s, m = _f32(src), _f32(mask)
        if s.shape != f.shape:
            raise RuntimeError(f"fspace_blend: stage {i} source {list(s.shape)} != {list(f.shape)}")
        m = m.reshape(m.shape[-2], m.shape[-1]) if m.numel() == m.shape[-2] * m.shape[-1] else None
        if m is None or (hm is not None and (m.shape[0], m.shape[1]) != (hm, wm)):
            raise RuntimeError("fspace_blend: every mask must be one [Hm,Wm] plane of the same size")
        hm, wm = m.shape
        keep += [s, m]
        srcs[i], masks[i], sa[i], sb[i] = s.data_ptr(), m.data_ptr(), float(a), float(b)
    out = torch.empty_like(f)
    _lib.use_device(f.device.index)
    _lib.check(_lib.lib().hf_fspace_blend_f32(f.data_ptr(), srcs, masks, sa, sb, out.data_ptr(), n, ch, h, w, hm, wm,
                                              _lib.stream_ptr()), "hf_fspace_blend_f32")
    return out


def align_f_space(intermediate_align, latent_F_1, latent_F_out_new, latent_F_2, free_mask) -> torch.Tensor:
    """Alignment.py:153-159.  ``free_mask`` = the stacked [3,1,256,256] (dilate[0], erosion[1], erosion[2])."""
    if free_mask.dim() != 4 or free_mask.shape[0] != 3 or free_mask.shape[1] != 1:
        raise RuntimeError(f"align_f_space: free_mask must be [3,1,H,W], got {list(free_mask.shape)}")
    fm = _f32(free_mask)
    # :157  intermediate_align + low[0] * (latent_F_1 - intermediate_align)  ==  src + w * (F - src) with F = latent_F_1
    return fspace_blend(latent_F_1, [(intermediate_align, fm[0], 1.0, -1.0), (latent_F_out_new, fm[1], 1.0, -1.0),
                                     (latent_F_2, fm[2], 1.0, -1.0)])


def mix_f_space(latent_F, latent_F_from_W, hair_mask, mixing: float) -> torch.Tensor:
    """Embedding.py:86-92 for ONE image: latent_F + mixing * bicubic32(hair_mask) * (latent_F_from_W - latent_F)
    = src + w * (first - src) with src = latent_F, first = latent_F_from_W, w = mixing * m."""
    return fspace_blend(latent_F_from_W, [(latent_F, hair_mask, 0.0, float(mixing))])
