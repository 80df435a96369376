"""CPU oracle for the inversion encoders.  TEST INFRASTRUCTURE ONLY -- same
rules as stylegan2_oracle.py (imported only by tests/, smoke() and bench.py's CPU legs).

Functional fp32 restatements (torch CPU, no nn.Module, no reference imports) of
* e4e ``Encoder4Editing.forward`` -- models/encoder4editing/models/encoders/psp_encoders.py:173-200 with
  ``bottleneck_IR_SE`` / ``SEModule`` / ``_upsample_add`` (helpers.py:57-140) and ``GradualStyleBlock``
  (psp_encoders.py:34-55);
* FSE ``fs_encoder_v2.forward`` -- models/FeatureStyleEncoder/nets/feature_style_encoder.py:47-65 with
  ``IBasicBlock`` (arcface/iresnet.py:28-57).
* PostProcess conv stack: ``FeatureEncoderMult(fs_layers=[9]).forward`` -- models/Net.py:396-477 --
  and ``FeatureiResnet.forward`` -- models/Encoders.py:35-57 -- with ``IBasicBlock`` (models/Net.py:162-190).
BatchNorm is evaluated in eval mode (running statistics), as on the swap() path.

Pinned by tests/golden/encoders.npz and tests/golden/postprocess.npz, produced by oracle/gen_golden_encoders.py /
gen_golden_postprocess.py from the unmodified reference classes with the same seeded synthetic parameters
(``synth_params_like``; no pretrained weights exist).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


def _bn(x, p, pre, eps=1e-5):
    return F.batch_norm(x, p[pre + "running_mean"], p[pre + "running_var"], p[pre + "weight"], p[pre + "bias"],
                        False, 0.0, eps)


def _prelu(x, w):
    return torch.where(x > 0, x, x * w.view(1, -1, 1, 1))


# ------------------------------------------------------------------------------------------------ e4e
E4E_UNITS = [(64, 64, 3), (64, 128, 4), (128, 256, 14), (256, 512, 3)]       # get_blocks(50), helpers.py:31-37


def e4e_block_specs():
    specs = []
    for cin, depth, n in E4E_UNITS:
        specs.append((cin, depth, 2))
        specs += [(depth, depth, 1)] * (n - 1)
    return specs


def bottleneck_ir_se_ref(x: Tensor, p: Dict[str, Tensor], pre: str, cin: int, depth: int, stride: int) -> Tensor:
    """helpers.py:98-120."""
    if cin == depth:
        shortcut = x[:, :, ::stride, ::stride]                                    # MaxPool2d(1, stride)
    else:
        shortcut = _bn(F.conv2d(x, p[pre + "shortcut_layer.0.weight"], stride=stride), p, pre + "shortcut_layer.1.")
    r = _bn(x, p, pre + "res_layer.0.")
    r = F.conv2d(r, p[pre + "res_layer.1.weight"], padding=1)
    r = _prelu(r, p[pre + "res_layer.2.weight"])
    r = F.conv2d(r, p[pre + "res_layer.3.weight"], stride=stride, padding=1)
    r = _bn(r, p, pre + "res_layer.4.")
    s = r.mean((2, 3), keepdim=True)                                              # SEModule, helpers.py:57-75
    s = F.relu(F.conv2d(s, p[pre + "res_layer.5.fc1.weight"]))
    s = torch.sigmoid(F.conv2d(s, p[pre + "res_layer.5.fc2.weight"]))
    return r * s + shortcut


def gradual_style_block_ref(x: Tensor, p: Dict[str, Tensor], pre: str, spatial: int) -> Tensor:
    """psp_encoders.py:34-55: log2(spatial) x [Conv3x3 s2 + LeakyReLU(0.01)] + EqualLinear(lr_mul=1)."""
    for i in range(int(math.log2(spatial))):
        x = F.leaky_relu(F.conv2d(x, p[pre + f"convs.{2 * i}.weight"], p[pre + f"convs.{2 * i}.bias"], 2, 1), 0.01)
    x = x.view(-1, x.shape[1])
    w = p[pre + "linear.weight"]
    return F.linear(x, w * (1.0 / math.sqrt(w.shape[1])), p[pre + "linear.bias"])


def e4e_ref(p: Dict[str, Tensor], x: Tensor, style_count: int = 18, return_taps: bool = False):
    """psp_encoders.py:173-200 (progressive_stage = Inference)."""
    x = F.conv2d(x, p["input_layer.0.weight"], padding=1)
    x = _prelu(_bn(x, p, "input_layer.1."), p["input_layer.2.weight"])
    taps = {}
    for i, (cin, depth, stride) in enumerate(e4e_block_specs()):
        x = bottleneck_ir_se_ref(x, p, f"body.{i}.", cin, depth, stride)
        if i in (6, 20, 23):
            taps[i] = x
    c1, c2, c3 = taps[6], taps[20], taps[23]
    w0 = gradual_style_block_ref(c3, p, "styles.0.", 16)
    w = w0.unsqueeze(1).repeat(1, style_count, 1)
    feats = c3
    for i in range(1, style_count):
        if i == 3:
            l = F.conv2d(c2, p["latlayer1.weight"], p["latlayer1.bias"])
            p2 = F.interpolate(c3, size=l.shape[2:], mode="bilinear", align_corners=True) + l
            feats = p2
        elif i == 7:
            l = F.conv2d(c1, p["latlayer2.weight"], p["latlayer2.bias"])
            p1 = F.interpolate(p2, size=l.shape[2:], mode="bilinear", align_corners=True) + l
            feats = p1
        spatial = 16 if i < 3 else (32 if i < 7 else 64)
        w[:, i] += gradual_style_block_ref(feats, p, f"styles.{i}.", spatial)
    return (w, taps) if return_taps else w


# ------------------------------------------------------------------------------------------------ FSE
FSE_STAGES = [("block_1", 64, 64, 3), ("block_2", 64, 128, 4), ("block_3", 128, 256, 14), ("block_4", 256, 512, 3)]


def ibasic_block_ref(x: Tensor, p: Dict[str, Tensor], pre: str, stride: int, has_ds: bool) -> Tensor:
    """arcface/iresnet.py:45-57."""
    out = _bn(x, p, pre + "bn1.")
    out = F.conv2d(out, p[pre + "conv1.weight"], padding=1)
    out = _prelu(_bn(out, p, pre + "bn2."), p[pre + "prelu.weight"])
    out = F.conv2d(out, p[pre + "conv2.weight"], stride=stride, padding=1)
    out = _bn(out, p, pre + "bn3.")
    idt = x
    if has_ds:
        idt = _bn(F.conv2d(x, p[pre + "downsample.0.weight"], stride=stride), p, pre + "downsample.1.")
    return out + idt


def fse_ref(p: Dict[str, Tensor], x: Tensor, content_stride: int = 2, n_styles: int = 18,
            content_stage: str = "block_3", content_prefix: str = "content_layer."):
    """nets/feature_style_encoder.py:47-65.  With ``content_stage="block_2", content_prefix="content_layer.0.",
    content_stride=1`` this is ``FeatureEncoderMult(fs_layers=[9]).forward`` after its resize (models/Net.py:446-477:
    max(fs_layers) > 7 taps the content branch after block_2, kernel (3,3), stride (1,1))."""
    x = F.conv2d(x, p["conv.0.weight"], padding=1)
    x = _prelu(_bn(x, p, "conv.1."), p["conv.2.weight"])
    feats, content = [], None
    for name, cin, planes, n in FSE_STAGES:
        for j in range(n):
            x = ibasic_block_ref(x, p, f"{name}.{j}.", 2 if j == 0 else 1, j == 0)
        if name == content_stage:
            c = _bn(x, p, content_prefix + "0.")
            c = F.conv2d(c, p[content_prefix + "1.weight"], padding=1)
            c = _prelu(_bn(c, p, content_prefix + "2."), p[content_prefix + "3.weight"])
            c = F.conv2d(c, p[content_prefix + "4.weight"], stride=content_stride, padding=1)
            content = _bn(c, p, content_prefix + "5.")
        feats.append(F.adaptive_avg_pool2d(x, (3, 3)))
    f = torch.cat(feats, dim=1).view(x.size(0), -1)
    out = torch.stack([F.linear(f, p[f"styles.{i}.weight"], p[f"styles.{i}.bias"]) for i in range(n_styles)], dim=1)
    return out, content


# ------------------------------------------------------------------------------------------------ PostProcess (8f-1)
def transform_to_256_ref(x: Tensor) -> Tensor:
    """models/Net.py:12-14 ``transforms.Resize((256, 256))`` on a tensor: bilinear, align_corners=False, and no
    antialias under the reference's pinned torchvision (0.14: ``antialias=None`` -> False for tensors)."""
    if x.shape[-2:] == (256, 256):
        return x
    return F.interpolate(x, size=(256, 256), mode="bilinear", align_corners=False)


def feature_encoder_mult_ref(p: Dict[str, Tensor], x: Tensor):
    """``FeatureEncoderMult(fs_layers=[9]).forward`` (models/Net.py:446-477) -> (latents [B,18,512], [content])."""
    lat, content = fse_ref(p, transform_to_256_ref(x), content_stride=1, content_stage="block_2",
                           content_prefix="content_layer.0.")
    return lat, [content]


PP_BLOCKS = [[1024, 2], [768, 2], [512, 2]]          # PostProcessModel.to_feature, models/Encoders.py:113


def feature_iresnet_ref(p: Dict[str, Tensor], x: Tensor, blocks=PP_BLOCKS, inplanes: int = 1024) -> Tensor:
    """``FeatureiResnet.forward`` (models/Encoders.py:35-57): IBasicBlocks at stride 1, a conv1x1 + BN shortcut
    where the width changes."""
    for n, (planes, num) in enumerate(blocks, start=1):
        for k in range(1, num + 1):
            x = ibasic_block_ref(x, p, f"res_blocks.res_block_{n}_{k}.", 1, inplanes != planes)
            inplanes = planes
    return x


# ------------------------------------------------------------------------------------------------ synthetic parameters
def _fill(sd: Dict[str, Tensor], seed: int) -> Dict[str, Tensor]:
    """Seeded values for every tensor of a freshly constructed module's state_dict: fan-in scaled conv / linear
    weights, non-trivial BatchNorm affine + running statistics and PReLU slopes (activations stay O(1))."""
    import zlib
    out = {}
    for k, v in sd.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))   # order independent
        if k.endswith("num_batches_tracked"):
            out[k] = torch.zeros_like(v)
        elif k.endswith("running_var"):
            out[k] = torch.rand(v.shape, generator=g) + 0.5
        elif k.endswith("running_mean"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.ndim == 4:                                            # conv weight
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            out[k] = torch.randn(v.shape, generator=g) * math.sqrt(1.5 / fan_in)
        elif v.ndim == 2 and "linear.weight" in k:                   # EqualLinear (scaled at run time)
            out[k] = torch.randn(v.shape, generator=g)
        elif v.ndim == 2:                                            # nn.Linear
            out[k] = torch.randn(v.shape, generator=g) / math.sqrt(v.shape[1])
        elif k.endswith("bias"):
            out[k] = torch.randn(v.shape, generator=g) * 0.1
        elif v.ndim == 1 and ("prelu" in k or k.endswith("res_layer.2.weight") or k.endswith("input_layer.2.weight")
                              or k.endswith("conv.2.weight") or k.endswith("content_layer.3.weight")
                              or k.endswith("content_layer.0.3.weight")):
            out[k] = torch.rand(v.shape, generator=g) * 0.3 + 0.1    # PReLU slopes
        elif k.endswith("res_layer.4.weight") or k.endswith("bn3.weight") or k.endswith("downsample.1.weight") \
                or k.endswith("shortcut_layer.1.weight"):
            # last BatchNorm of a residual branch / shortcut: keep the 24 stacked blocks from blowing up
            out[k] = torch.rand(v.shape, generator=g) * 0.3 + 0.2
        else:                                                        # BatchNorm weight
            out[k] = torch.rand(v.shape, generator=g) + 0.5
    return out


def synth_params_like(module, seed: int) -> Dict[str, Tensor]:
    """Seeded parameters with the key/shape layout of ``module.state_dict()`` (the layout itself is pinned against
    the reference classes in gen_golden_encoders.py via strict load_state_dict)."""
    return _fill({k: v for k, v in module.state_dict().items()}, seed)
