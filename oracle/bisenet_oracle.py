"""CPU oracle for the BiSeNet face-parsing network.  TEST INFRASTRUCTURE ONLY -- same rules as
stylegan2_oracle.py (imported only by tests/, smoke() and bench.py's CPU legs).

Functional fp32 restatement (torch CPU, no nn.Module, no reference imports) of ``BiSeNet.forward`` --
models/CtrlHair/external_code/face_parsing/model.py:227-244 -- with ``ContextPath`` (:98-132), ``AttentionRefinementModule``
(:70-95), ``FeatureFusionModule`` (:178-224), ``BiSeNetOutput`` (:38-48), ``ConvBNReLU`` (:11-30) and the ``Resnet18``
backbone / ``BasicBlock`` (face_parsing/resnet.py:20-84).  BatchNorm in eval mode, as FaceParsing runs it
(my_parsing_util.py: ``bise_net.eval()``).

Pinned by tests/golden/bisenet.npz, produced by oracle/gen_golden_bisenet.py from the unmodified reference classes
with seeded synthetic parameters (no pretrained weights exist here; the reference's ResNet-18 download is stubbed).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _bn(x, p, pre, eps=1e-5):
    return F.batch_norm(x, p[pre + "running_mean"], p[pre + "running_var"], p[pre + "weight"], p[pre + "bias"],
                        False, 0.0, eps)


def conv_bn_relu_ref(x, p, pre, stride=1, padding=1):
    """model.py:11-24."""
    return F.relu(_bn(F.conv2d(x, p[pre + "conv.weight"], stride=stride, padding=padding), p, pre + "bn."))


def basic_block_ref(x, p, pre, stride):
    """resnet.py:20-47: relu(shortcut + bn2(conv2(relu(bn1(conv1(x))))))."""
    r = F.relu(_bn(F.conv2d(x, p[pre + "conv1.weight"], stride=stride, padding=1), p, pre + "bn1."))
    r = _bn(F.conv2d(r, p[pre + "conv2.weight"], padding=1), p, pre + "bn2.")
    s = x
    if pre + "downsample.0.weight" in p:
        s = _bn(F.conv2d(x, p[pre + "downsample.0.weight"], stride=stride), p, pre + "downsample.1.")
    return F.relu(s + r)


def resnet18_ref(x, p, pre="cp.resnet."):
    """resnet.py:57-84 -> (feat8, feat16, feat32)."""
    x = F.relu(_bn(F.conv2d(x, p[pre + "conv1.weight"], stride=2, padding=3), p, pre + "bn1."))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
        x = basic_block_ref(x, p, f"{pre}layer{li}.0.", stride)
        x = basic_block_ref(x, p, f"{pre}layer{li}.1.", 1)
        if li >= 2:
            feats.append(x)
    return feats


def arm_ref(x, p, pre):
    """model.py:70-88."""
    feat = conv_bn_relu_ref(x, p, pre + "conv.")
    atten = feat.mean((2, 3), keepdim=True)
    atten = torch.sigmoid(_bn(F.conv2d(atten, p[pre + "conv_atten.weight"]), p, pre + "bn_atten."))
    return feat * atten


def context_path_ref(x, p, pre="cp."):
    """model.py:108-132 -> (feat8, feat16_up [1/8], feat32_up [1/16])."""
    feat8, feat16, feat32 = resnet18_ref(x, p, pre + "resnet.")
    avg = conv_bn_relu_ref(feat32.mean((2, 3), keepdim=True), p, pre + "conv_avg.", padding=0)
    feat32_sum = arm_ref(feat32, p, pre + "arm32.") + avg                      # nearest-upsampled 1x1 -> broadcast
    feat32_up = conv_bn_relu_ref(F.interpolate(feat32_sum, feat16.shape[2:], mode="nearest"), p, pre + "conv_head32.")
    feat16_sum = arm_ref(feat16, p, pre + "arm16.") + feat32_up
    feat16_up = conv_bn_relu_ref(F.interpolate(feat16_sum, feat8.shape[2:], mode="nearest"), p, pre + "conv_head16.")
    return feat8, feat16_up, feat32_up


def ffm_ref(fsp, fcp, p, pre="ffm."):
    """model.py:200-211."""
    feat = conv_bn_relu_ref(torch.cat([fsp, fcp], dim=1), p, pre + "convblk.", padding=0)
    atten = feat.mean((2, 3), keepdim=True)
    atten = torch.sigmoid(F.conv2d(F.relu(F.conv2d(atten, p[pre + "conv1.weight"])), p[pre + "conv2.weight"]))
    return feat * atten + feat


def output_ref(x, p, pre):
    """model.py:38-48."""
    return F.conv2d(conv_bn_relu_ref(x, p, pre + "conv."), p[pre + "conv_out.weight"])


def bisenet_ref(p: Dict[str, Tensor], x: Tensor, return_lowres: bool = False):
    """model.py:227-244 -> (feat_out, feat_out16, feat_out32), each [B, n_classes, H, W]."""
    H, W = x.shape[2:]
    feat_res8, feat_cp8, feat_cp16 = context_path_ref(x, p)
    fuse = ffm_ref(feat_res8, feat_cp8, p)
    low = [output_ref(fuse, p, "conv_out."), output_ref(feat_cp8, p, "conv_out16."), output_ref(feat_cp16, p, "conv_out32.")]
    if return_lowres:
        return low
    return tuple(F.interpolate(t, (H, W), mode="bilinear", align_corners=True) for t in low)
