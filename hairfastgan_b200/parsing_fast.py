"""Label-only face parsing.  ``FaceParsing_tensor.parsing_img`` of the reference
(models/CtrlHair/external_code/face_parsing/my_parsing_util.py:70-89) runs BiSeNet, keeps the first of its three
outputs and immediately reduces the [1,19,H,W] fp32 logits to labels (``out.squeeze(0).argmax(0)``); the caller
(`get_segmentation`, models/Net.py:108-115) never sees the logits.  With this package's BiSeNet in place the same labels
come from ``BiSeNet.parse_labels``: the two auxiliary heads are not evaluated and the align_corners bilinear upsampling
is fused with the arg-max, so 76 B per pixel of logits (80 MB per 1024^2 image, written once and read once) are
replaced by 8 B of label.  The labels are bit-identical to ``forward(img)[0].squeeze(0).argmax(0)`` (tested), so the
patch is on by default with ``install(segmentation=True)``; ``install(fuse_face_parsing=False)`` leaves the reference
function untouched.  Anything other than a single CUDA image falls through to the original."""
from __future__ import annotations

import torch


def make_fast_parsing(module, original):
    def parsing_img(img=None):
        net = module.FaceParsing.bise_net
        if (img is None or net is None or not hasattr(net, "parse_labels") or not torch.is_tensor(img)
                or img.dim() != 4 or img.shape[0] != 1 or not img.is_cuda):
            return original(img)
        return net.parse_labels(img)[0], None

    parsing_img.__wrapped__ = original
    return parsing_img


def patch_parsing_module(module) -> bool:
    cls = getattr(module, "FaceParsing_tensor", None)
    if cls is None or "parsing_img" not in cls.__dict__:
        return False
    current = cls.__dict__["parsing_img"]
    fn = current.__func__ if isinstance(current, staticmethod) else current
    if getattr(fn, "__wrapped__", None) is not None:
        return False
    cls.parsing_img = staticmethod(make_fast_parsing(module, fn))
    return True


def unpatch_parsing_module(module) -> None:
    cls = getattr(module, "FaceParsing_tensor", None)
    if cls is None:
        return
    current = cls.__dict__.get("parsing_img")
    fn = current.__func__ if isinstance(current, staticmethod) else current
    orig = getattr(fn, "__wrapped__", None)
    if orig is not None:
        cls.parsing_img = staticmethod(orig)
