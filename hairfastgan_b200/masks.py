"""``DilateErosion`` (stage glue): counterpart of utils/image_utils.py:27-55, the mask growing / shrinking
that Alignment and Blending apply to the 256^2 hair masks (Alignment.py:40,147; Blending.py:32,41,52).

The reference repeats {conv2d with a 3x3 cross, threshold} ``dilate_erosion`` times on a doubled batch; here each
round is one launch of ``hf_dilate_erode_f32`` that advances the dilated and the eroded mask together.  For masks
with values in {0, 1} (what HairFast feeds it) the result is bit-identical.  CUDA tensors only.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import _lib

__all__ = ["DilateErosion"]


class DilateErosion:
    def __init__(self, dilate_erosion=5, device='cuda'):
        self.dilate_erosion = dilate_erosion
        cross = torch.zeros(1, 1, 3, 3)
        cross[0, 0, 1, :] = 1.
        cross[0, 0, :, 1] = 1.
        self.weight = cross.to(device)          # kept for callers that read it; the kernel hard-codes the cross

    def hair_from_mask(self, mask):
        """Label map -> (dilated, eroded) 256^2 hair masks (label 13 = hair, image_utils.py:36-40)."""
        hair = torch.where(mask == 13, torch.ones_like(mask), torch.zeros_like(mask))
        hair = F.interpolate(hair, size=(256, 256), mode='nearest')
        return self.mask(hair)

    def mask(self, mask):
        if not mask.is_cuda:
            raise RuntimeError("DilateErosion: input must be a CUDA tensor (no CPU fallback)")
        if mask.dim() != 4 or mask.shape[1] != 1:
            raise RuntimeError("DilateErosion.mask: expected a [N, 1, H, W] mask")
        m = mask.detach().float().contiguous()
        n, _, h, w = m.shape
        grown, shrunk = torch.empty_like(m), torch.empty_like(m)
        rounds = int(self.dilate_erosion)
        scratch = torch.empty(4 * m.numel() if rounds > 1 else 1, device=m.device, dtype=torch.float32)
        _lib.use_device(m.device.index)
        _lib.check(_lib.lib().hf_dilate_erode_f32(m.data_ptr(), grown.data_ptr(), shrunk.data_ptr(), scratch.data_ptr(),
                                                  n, h, w, rounds, _lib.stream_ptr()), "hf_dilate_erode_f32")
        return grown, shrunk
