"""CPU oracle for the HairFastGAN hot path (StyleGAN2 generator forward).

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product path
(``hairfastgan_b200``) never imports anything from ``oracle/`` and has no CPU
fallback.

It is an independent fp32 restatement (torch CPU tensors; functional style, no
nn.Module, no reference imports) of the reference algorithm.  Every function
cites the reference lines it follows (paths relative to the reference repo
AIRI-Institute/HairFastGAN @ 49e98019).

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against the reference's own
Python implementation imported in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
re-checks the oracle against those committed vectors on every run).

Two families of functions live here:

* ``*_ref``   -- literal restatements of the reference formulation
  (materialised per-sample weights, grouped conv, conv_transpose + blur).
* ``*_fused`` -- the algebraically equivalent formulations the CUDA kernels use
  (shared weights + pre-scaled activations + demod in the epilogue; the
  polyphase form of the upsampling conv).  They exist so tests can check the
  algebra separately from the kernels, and so kernel intermediates can be
  compared one stage at a time.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SQRT2 = 2.0 ** 0.5


# --------------------------------------------------------------------------- #
# op boundary: upfirdn2d / fused_leaky_relu
# --------------------------------------------------------------------------- #

def make_kernel(k: Sequence[float]) -> Tensor:
    """models/stylegan2/model.py:24-32 -- normalised outer-product FIR."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


def upfirdn2d_ref(x: Tensor, kernel: Tensor, up: int = 1, down: int = 1,
                  pad: Tuple[int, int] = (0, 0)) -> Tensor:
    """models/stylegan2/op/upfirdn2d.py:145-200 (semantics of
    op/upfirdn2d_kernel.cu:107-207): zero-stuff, pad/crop, correlate with the
    flipped kernel (true convolution), decimate.  Same up/down/pad on both axes.
    x: [N,C,H,W] -> [N,C,H',W'] with H' = (H*up + p0 + p1 - kh)//down + 1.
    """
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    p0, p1 = pad
    # 1. zero-stuff: zeros FOLLOW each sample (upfirdn2d.py:168-170)
    u = x.new_zeros(n, c, h * up, w * up)
    u[:, :, ::up, ::up] = x
    # 2. pad (negative = crop) (upfirdn2d.py:172-180)
    u = F.pad(u, [max(p0, 0), max(p1, 0), max(p0, 0), max(p1, 0)])
    u = u[:, :, max(-p0, 0): u.shape[2] - max(-p1, 0), max(-p0, 0): u.shape[3] - max(-p1, 0)]
    # 3. valid correlation with the flipped kernel (upfirdn2d.py:186-187)
    kf = torch.flip(kernel, [0, 1]).reshape(1, 1, kh, kw).to(x.dtype)
    hh, ww = u.shape[2], u.shape[3]
    f = F.conv2d(u.reshape(n * c, 1, hh, ww), kf)
    # 4. decimate (upfirdn2d.py:195)
    f = f[:, :, ::down, ::down]
    return f.reshape(n, c, f.shape[2], f.shape[3])


def fused_leaky_relu_ref(x: Tensor, bias: Optional[Tensor], negative_slope: float = 0.2,
                         scale: float = SQRT2) -> Tensor:
    """models/stylegan2/op/fused_act.py:85-96 and op/fused_bias_act_kernel.cu:27-47:
    y = lrelu(x + b[c]) * scale, channel = dim 1.  (The CUDA kernel honours
    ``negative_slope``; the reference CPU branch hard-codes 0.2 -- identical for
    every in-tree caller.  We follow the CUDA kernel.)"""
    if bias is not None:
        x = x + bias.reshape(1, -1, *([1] * (x.ndim - 2)))
    return torch.where(x > 0, x, x * negative_slope) * scale


# --------------------------------------------------------------------------- #
# module level, reference formulation
# --------------------------------------------------------------------------- #

def equal_linear_ref(x: Tensor, weight: Tensor, bias: Optional[Tensor], lr_mul: float = 1.0,
                     activation: bool = False) -> Tensor:
    """models/stylegan2/model.py:134-163 EqualLinear."""
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        out = F.linear(x, weight * scale)
        return fused_leaky_relu_ref(out, bias * lr_mul)
    return F.linear(x, weight * scale, bias=None if bias is None else bias * lr_mul)


def modulated_conv2d_ref(x: Tensor, style: Tensor, weight: Tensor, mod_weight: Tensor,
                         mod_bias: Tensor, demodulate: bool = True, upsample: bool = False,
                         blur_kernel: Optional[Tensor] = None) -> Tensor:
    """models/stylegan2/model.py:238-279 (plain and upsample branches).
    weight: [1,Cout,Cin,k,k]; style: [B,style_dim]; x: [B,Cin,H,W]."""
    b, cin, h, w = x.shape
    _, cout, _, k, _ = weight.shape
    s = equal_linear_ref(style, mod_weight, mod_bias).view(b, 1, cin, 1, 1)      # :241
    scale = 1.0 / math.sqrt(cin * k * k)                                          # :220-221
    wgt = scale * weight * s                                                      # :242
    if demodulate:
        demod = torch.rsqrt(wgt.pow(2).sum([2, 3, 4]) + 1e-8)                     # :245
        wgt = wgt * demod.view(b, cout, 1, 1, 1)
    if upsample:
        xin = x.reshape(1, b * cin, h, w)                                         # :253
        wt = wgt.transpose(1, 2).reshape(b * cin, cout, k, k)                     # :257-259
        out = F.conv_transpose2d(xin, wt, padding=0, stride=2, groups=b)          # :260
        out = out.view(b, cout, out.shape[2], out.shape[3])
        # Blur(pad=(pad0,pad1)) from :197-204: p = (len(k)-2)-(ksize-1)
        assert blur_kernel is not None
        p = (blur_kernel.shape[0] - 2) - (k - 1)
        pad0, pad1 = (p + 1) // 2 + 2 - 1, p // 2 + 1
        return upfirdn2d_ref(out, blur_kernel, pad=(pad0, pad1))                  # :263
    xin = x.reshape(1, b * cin, h, w)
    out = F.conv2d(xin, wgt.view(b * cout, cin, k, k), padding=k // 2, groups=b)  # :275
    return out.view(b, cout, out.shape[2], out.shape[3])


def styled_conv_ref(x: Tensor, style: Tensor, p: Dict[str, Tensor], prefix: str,
                    noise: Tensor, upsample: bool) -> Tensor:
    """models/stylegan2/model.py:337-343 StyledConv = ModulatedConv2d -> NoiseInjection
    (:288-293) -> FusedLeakyReLU (op/fused_act.py:73-82).  ``noise`` must be given
    (the oracle never draws random numbers)."""
    out = modulated_conv2d_ref(
        x, style, p[prefix + "conv.weight"], p[prefix + "conv.modulation.weight"],
        p[prefix + "conv.modulation.bias"], True, upsample,
        p.get(prefix + "conv.blur.kernel"))
    out = out + p[prefix + "noise.weight"] * noise
    return fused_leaky_relu_ref(out, p[prefix + "activate.bias"])


def to_rgb_ref(x: Tensor, style: Tensor, p: Dict[str, Tensor], prefix: str,
               skip: Optional[Tensor]) -> Tensor:
    """models/stylegan2/model.py:346-365 ToRGB: modulated 1x1 (no demod) + bias +
    Upsample(skip) (model.py:35-53: up=2, kernel*4, pad=(2,1))."""
    out = modulated_conv2d_ref(
        x, style, p[prefix + "conv.weight"], p[prefix + "conv.modulation.weight"],
        p[prefix + "conv.modulation.bias"], demodulate=False)
    out = out + p[prefix + "bias"]
    if skip is not None:
        k = p[prefix + "upsample.kernel"]
        pp = k.shape[0] - 2
        out = out + upfirdn2d_ref(skip, k, up=2, down=1, pad=((pp + 1) // 2 + 1, pp // 2))
    return out


def generator_ref(p: Dict[str, Tensor], latent: Tensor, noise: List[Tensor],
                  start_layer: int = 0, end_layer: int = 8, layer_in: Optional[Tensor] = None,
                  skip: Optional[Tensor] = None):
    """models/stylegan2/model.py:477-565 Generator.forward with
    ``input_is_latent=True`` and a [B,n_latent,512] latent (the only form swap()
    uses, model.py:521-522).  ``noise`` = 17 explicit tensors.  Returns
    ``(image, None)`` or, on the early exit (:537-538, :550-551), ``(out, skip)``."""
    b = latent.shape[0]
    out = p["input.input"].repeat(b, 1, 1, 1)                                     # :532
    if start_layer == 0:
        out = styled_conv_ref(out, latent[:, 0], p, "conv1.", noise[0], False)   # :535
        skip = to_rgb_ref(out, latent[:, 1], p, "to_rgb1.", None)                 # :536
    if end_layer == 0:
        return out, skip
    n_layers = (len(noise) - 1) // 2
    i = 1
    for layer in range(1, n_layers + 1):                                          # :541
        c1, c2, rgb = f"convs.{2 * layer - 2}.", f"convs.{2 * layer - 1}.", f"to_rgbs.{layer - 1}."
        n1, n2 = noise[2 * layer - 1], noise[2 * layer]
        if layer < start_layer:
            pass
        elif layer > end_layer:
            return out, skip                                                      # :550-551
        else:
            src = layer_in if layer == start_layer else out                       # :546-547
            out = styled_conv_ref(src, latent[:, i], p, c1, n1, True)
            out = styled_conv_ref(out, latent[:, i + 1], p, c2, n2, False)
            skip = to_rgb_ref(out, latent[:, i + 2], p, rgb, skip)
        i += 2
    return skip, None


def generator_fse_ref(p: Dict[str, Tensor], latent: Tensor, noise: List[Tensor],
                      features_in: Optional[List[Optional[Tensor]]] = None, feature_scale: float = 1.0):
    """FeatureStyleEncoder's generator copy
    (models/FeatureStyleEncoder/pixel2style2pixel/models/stylegan2/model.py:527-560): full forward with
    ``insert_feature`` (x = (1-a)*x + a*features_in[idx]) before every styled conv of the layer loop and
    ``return_features`` (list of the ConstantInput output and every StyledConv output).
    Returns (image, outs)."""
    def insert(x, idx):
        if features_in is not None and features_in[idx] is not None:
            x = (1 - feature_scale) * x + feature_scale * features_in[idx]
        return x
    b = latent.shape[0]
    outs = []
    out = p["input.input"].repeat(b, 1, 1, 1)
    outs.append(out)
    out = styled_conv_ref(out, latent[:, 0], p, "conv1.", noise[0], False)
    outs.append(out)
    skip = to_rgb_ref(out, latent[:, 1], p, "to_rgb1.", None)
    n_layers = (len(noise) - 1) // 2
    i = 1
    for layer in range(1, n_layers + 1):
        c1, c2, rgb = f"convs.{2 * layer - 2}.", f"convs.{2 * layer - 1}.", f"to_rgbs.{layer - 1}."
        out = insert(out, i)
        out = styled_conv_ref(out, latent[:, i], p, c1, noise[2 * layer - 1], True)
        outs.append(out)
        out = insert(out, i + 1)
        out = styled_conv_ref(out, latent[:, i + 1], p, c2, noise[2 * layer], False)
        outs.append(out)
        skip = to_rgb_ref(out, latent[:, i + 2], p, rgb, skip)
        i += 2
    return skip, outs


def mapping_ref(p: Dict[str, Tensor], z: Tensor, n_mlp: int = 8, lr_mlp: float = 0.01) -> Tensor:
    """models/stylegan2/model.py:16-21 PixelNorm + :383-393 eight EqualLinear
    (lr_mul=0.01, fused_lrelu)."""
    x = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(1, n_mlp + 1):
        x = equal_linear_ref(x, p[f"style.{i}.weight"], p[f"style.{i}.bias"], lr_mlp, True)
    return x


# --------------------------------------------------------------------------- #
# the algebraic forms the CUDA kernels implement
# --------------------------------------------------------------------------- #

def modulation_tables(style: Tensor, weight: Tensor, mod_weight: Tensor, mod_bias: Tensor,
                      demodulate: bool = True):
    """s[b,i] (model.py:241) and d[b,o] = rsqrt(sum_i s^2 * Wsq[o,i] + 1e-8) with
    Wsq[o,i] = sum_taps (scale*W)^2 -- same value as model.py:244-245."""
    _, cout, cin, k, _ = weight.shape
    s = equal_linear_ref(style, mod_weight, mod_bias)                              # [B,Cin]
    wt = weight[0] * (1.0 / math.sqrt(cin * k * k))
    if not demodulate:
        return s, torch.ones(style.shape[0], cout), wt
    wsq = wt.pow(2).sum([2, 3])                                                   # [Cout,Cin]
    d = torch.rsqrt((s * s) @ wsq.t() + 1e-8)
    return s, d, wt


def modulated_conv2d_fused(x, style, weight, mod_weight, mod_bias, demodulate=True):
    """Shared-weight form: y = d[b,o] * conv(x * s[b,i], W~) (Appendix C-1)."""
    s, d, wt = modulation_tables(style, weight, mod_weight, mod_bias, demodulate)
    k = weight.shape[-1]
    y = F.conv2d(x * s[:, :, None, None], wt, padding=k // 2)
    return y * d[:, :, None, None]


def polyphase_weights(wt: Tensor, blur_kernel: Tensor) -> Tensor:
    """Compose the stride-2 transposed 3x3 conv (model.py:260) with the 4x4 blur
    (pad (1,1), model.py:263) into four 3x3 correlation kernels, one per output
    parity (Appendix C-2).

    wt: [Cout,Cin,3,3] (already scaled).  Returns Kp[py,px,Cout,Cin,3,3] such that
      out[o, 2m+py, 2n+px] = sum_{i,dy,dx} Kp[py,px,o,i,dy,dx] * x[i, m+dy-1, n+dx-1].
    Derivation: out[y,x] = sum_{i,j} x[i,j] K6[y-2i, x-2j],
      K6[u,v] = sum_{a,b} W[a,b] k[2-a+u, 2-b+v]   (k = blur kernel, indices in 0..3),
    and y=2m+py, i=m+dy-1  =>  u = py + 2 - 2*dy.
    """
    cout, cin, kh, kw = wt.shape
    assert kh == 3 and kw == 3 and tuple(blur_kernel.shape) == (4, 4)
    kp = wt.new_zeros(2, 2, cout, cin, 3, 3)
    for py in range(2):
        for px in range(2):
            for dy in range(3):
                for dx in range(3):
                    u, v = py + 2 - 2 * dy, px + 2 - 2 * dx
                    acc = wt.new_zeros(cout, cin)
                    for a in range(3):
                        for b in range(3):
                            iy, ix = 2 - a + u, 2 - b + v
                            if 0 <= iy < 4 and 0 <= ix < 4:
                                acc = acc + wt[:, :, a, b] * blur_kernel[iy, ix]
                    kp[py, px, :, :, dy, dx] = acc
    return kp


def modulated_conv2d_up_fused(x, style, weight, mod_weight, mod_bias, blur_kernel):
    """Polyphase + shared-weight form of the upsampling ModulatedConv2d."""
    s, d, wt = modulation_tables(style, weight, mod_weight, mod_bias, True)
    kp = polyphase_weights(wt, blur_kernel)
    b, cin, h, w = x.shape
    cout = wt.shape[0]
    xs = x * s[:, :, None, None]
    y = x.new_zeros(b, cout, 2 * h, 2 * w)
    for py in range(2):
        for px in range(2):
            y[:, :, py::2, px::2] = F.conv2d(xs, kp[py, px], padding=1)
    return y * d[:, :, None, None]


def upsample2_polyphase(skip: Tensor, kernel: Tensor) -> Tensor:
    """upfirdn2d(skip, k, up=2, pad=(2,1)) (model.py:35-53) written per output
    parity: out[2m+py] = sum_t k[3-(py+2t... )] -- used to check the fused
    rgb_combine kernel's tap table."""
    n, c, h, w = skip.shape
    out = skip.new_zeros(n, c, 2 * h, 2 * w)
    sp = F.pad(skip, [1, 1, 1, 1])
    # out[y] = sum_p up_pad[y+p] kflip[p], up_pad[u] = stuffed[u-2], stuffed[2i]=x[i]
    # y+p-2 = 2i -> for y=2m+py: p = 2(i-m) + 2 - py, kflip[p] = k[3-p]
    for py in range(2):
        for px in range(2):
            acc = 0
            for di in (-1, 0, 1):
                for dj in (-1, 0, 1):
                    p_, q_ = 2 * di + 2 - py, 2 * dj + 2 - px
                    if 0 <= p_ < 4 and 0 <= q_ < 4:
                        acc = acc + kernel[3 - p_, 3 - q_] * sp[:, :, 1 + di:1 + di + h, 1 + dj:1 + dj + w]
            out[:, :, py::2, px::2] = acc
    return out


# --------------------------------------------------------------------------- #
# synthetic parameters (no pretrained weights exist anywhere)
# --------------------------------------------------------------------------- #

CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512}


def channels_for(size: int, channel_multiplier: int = 2) -> Dict[int, int]:
    """models/stylegan2/model.py:395-405."""
    return {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * channel_multiplier,
            128: 128 * channel_multiplier, 256: 64 * channel_multiplier,
            512: 32 * channel_multiplier, 1024: 16 * channel_multiplier}


def synth_generator_params(size: int = 1024, style_dim: int = 512, n_mlp: int = 8,
                           channel_multiplier: int = 2, seed: int = 0,
                           small: Optional[Dict[int, int]] = None) -> Dict[str, Tensor]:
    """Seeded random parameters with exactly the reference Generator's state_dict
    keys and shapes (models/stylegan2/model.py:369-451), with
    non-trivial noise.weight / activate.bias / to_rgb bias so every term of the
    path is exercised.  ``small`` optionally overrides the channel table (tests)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    ch = dict(channels_for(size, channel_multiplier))
    if small:
        ch.update(small)
    log_size = int(math.log2(size))
    p: Dict[str, Tensor] = {}
    for i in range(1, n_mlp + 1):
        p[f"style.{i}.weight"] = rn(style_dim, style_dim) / 0.01
        p[f"style.{i}.bias"] = rn(style_dim) * 0.1
    p["input.input"] = rn(1, ch[4], 4, 4)
    blur = make_kernel([1, 3, 3, 1]) * 4

    def styled(prefix, cin, cout, up):
        p[prefix + "conv.weight"] = rn(1, cout, cin, 3, 3)
        p[prefix + "conv.modulation.weight"] = rn(cin, style_dim)
        p[prefix + "conv.modulation.bias"] = torch.ones(cin) + 0.1 * rn(cin)
        if up:
            p[prefix + "conv.blur.kernel"] = blur.clone()
        p[prefix + "noise.weight"] = rn(1) * 0.1
        p[prefix + "activate.bias"] = rn(cout) * 0.1

    def torgb(prefix, cin, up):
        p[prefix + "bias"] = rn(1, 3, 1, 1) * 0.1
        if up:
            p[prefix + "upsample.kernel"] = blur.clone()
        p[prefix + "conv.weight"] = rn(1, 3, cin, 1, 1)
        p[prefix + "conv.modulation.weight"] = rn(cin, style_dim)
        p[prefix + "conv.modulation.bias"] = torch.ones(cin) + 0.1 * rn(cin)

    styled("conv1.", ch[4], ch[4], False)
    torgb("to_rgb1.", ch[4], False)
    cin = ch[4]
    for i in range(3, log_size + 1):
        cout = ch[2 ** i]
        styled(f"convs.{2 * (i - 3)}.", cin, cout, True)
        styled(f"convs.{2 * (i - 3) + 1}.", cout, cout, False)
        torgb(f"to_rgbs.{i - 3}.", cout, True)
        cin = cout
    for layer_idx in range((log_size - 2) * 2 + 1):
        res = (layer_idx + 5) // 2
        p[f"noises.noise_{layer_idx}"] = rn(1, 1, 2 ** res, 2 ** res)
    return p


def synth_noise(size: int, batch: int = 1, seed: int = 1) -> List[Tensor]:
    """Explicit noise list in the layout of Generator.make_noise (model.py:455-464)."""
    g = torch.Generator().manual_seed(seed)
    log_size = int(math.log2(size))
    out = [torch.randn(batch, 1, 4, 4, generator=g)]
    for i in range(3, log_size + 1):
        for _ in range(2):
            out.append(torch.randn(batch, 1, 2 ** i, 2 ** i, generator=g))
    return out
