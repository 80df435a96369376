"""Seeded synthetic `pretrained_models/` tree for the unmodified reference: there are no pretrained
weights on disk and no network, so every checkpoint `HairFast.__init__` loads is written here with the state-dict
schema its loader expects, by instantiating the STOCK reference class (its own default init under a fixed seed) and
saving `state_dict()`.

    python baseline/synth_checkpoints.py <workdir> [--seed 0]

Loader of each file (reference file:line):
  StyleGAN/ffhq.pt {'g_ema','latent_avg'}, ffhq_PCA.npz                      models/Net.py:37-42,65-76
  encoder4editing/e4e_ffhq_encode.pt {'opts','state_dict','latent_avg'}      encoder4editing/utils/model_utils.py:17-28, models/psp.py:41-47,94-104
  FeatureStyleEncoder/{143_enc.pth,psp_ffhq_encode.pt,backbone.pth,79999_iter.pth}   FSencoder.py:27-40, trainer.py:188-201
  BiSeNet/face_parsing_79999_iter.pth                                        face_parsing/my_parsing_util.py:78-79
  sean_checkpoints/CelebA-HQ_pretrained/latest_net_G.pth                     sean_codes/util/util.py:204-209
  ShapeAdaptor/mask_generator.pth                                            models/Alignment.py:32-34
  Rotate/rotate_best.pth, Blending/checkpoint.pth, PostProcess/pp_model.pth  models/Alignment.py:36-37, models/Blending.py:24-30
  PostProcess/latent_avg.pt, ArcFace/backbone_ir50.pth                       models/Encoders.py:109-112, models/Net.py:340-341

The generator gets non-zero `noise.weight`, `activate.bias` and ToRGB biases (the stock init leaves them 0, which would
hide those code paths) and ToRGB weights scaled so the synthetic RGB stays within about [-1, 1] like a trained model's.
Harness code (tests / bench legs only); the product never imports it.
"""
from __future__ import annotations

import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import refenv  # noqa: E402

RGB_WEIGHT_SCALE = 0.12


def _save(obj, path):
    import torch
    os.makedirs(os.path.dirname(path), exist_ok=True)
    torch.save(obj, path)


def _cpu_sd(module):
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}


_RESIDUAL_TAILS = ("bn3.weight", "downsample.1.weight", "res_layer.4.weight", "shortcut_layer.1.weight")


def _tame(sd, seed: int, gain: float = 1.5):
    """The stock constructors' init (e.g. conv std 0.1 in arcface/iresnet.py:104-105, models/Net.py:214-215) is meant
    to be overwritten by a trained checkpoint: run as is, 24 un-normalised residual blocks reach 1e25 and the swap ends
    in NaN.  Re-draw the conv weights fan-in scaled, damp the last BatchNorm of every residual branch and give the
    BatchNorm statistics / PReLU slopes non-trivial values, so activations stay O(1) through the whole pipeline."""
    import math
    import zlib
    import torch
    for k, v in sd.items():
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(k.encode())) % (2 ** 31))
        if v.ndim == 4 and v.shape[2] == v.shape[3] and v.shape[1] > 1:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            v.copy_(torch.randn(v.shape, generator=g) * math.sqrt(gain / fan_in))
        elif k.endswith(_RESIDUAL_TAILS):
            v.copy_(torch.rand(v.shape, generator=g) * 0.3 + 0.2)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    return sd


def generator_state(seed: int):
    import torch
    from models.stylegan2.model import Generator
    torch.manual_seed(seed)
    g = Generator(1024, 512, 8, channel_multiplier=2)
    sd = _cpu_sd(g)
    gen = torch.Generator().manual_seed(seed + 1)
    for k, v in sd.items():
        if k.endswith("noise.weight"):
            v.copy_(0.1 * torch.randn(v.shape, generator=gen) + 0.05)
        elif k.endswith("activate.bias") or (k.startswith("to_rgb") and k.endswith(".bias") and v.ndim == 4):
            v.copy_(0.1 * torch.randn(v.shape, generator=gen))
        elif k.startswith("to_rgb") and k.endswith("conv.weight"):
            v.mul_(RGB_WEIGHT_SCALE)
    return sd


def write_all(workdir: str, seed: int = 0, verbose: bool = True) -> str:
    """Writes <workdir>/pretrained_models/...; the reference must already be importable (refenv.activate) and the cwd
    is switched to `workdir` (several constructors read other checkpoints by relative path)."""
    import numpy as np
    import torch
    pm = os.path.join(workdir, "pretrained_models")
    marker = os.path.join(pm, f".complete_seed{seed}")
    if os.path.exists(marker):
        return pm
    os.makedirs(pm, exist_ok=True)
    os.chdir(workdir)
    say = (lambda *a: print("[synth]", *a, flush=True)) if verbose else (lambda *a: None)
    gen = torch.Generator().manual_seed(seed + 100)

    # --- StyleGAN2 -----------------------------------------------------------------------------------------------
    g_sd = generator_state(seed)
    latent_avg = 0.5 * torch.randn(512, generator=gen)
    _save({"g_ema": g_sd, "latent_avg": latent_avg}, os.path.join(pm, "StyleGAN/ffhq.pt"))
    q, _ = torch.linalg.qr(torch.randn(512, 512, generator=gen))
    np.savez(os.path.join(pm, "StyleGAN/ffhq_PCA.npz"), X_mean=torch.randn(512, generator=gen).numpy(),
             X_comp=q.numpy(), X_stdev=(torch.rand(512, generator=gen) + 0.5).numpy(),
             X_var_ratio=(torch.ones(512) / 512).numpy())
    say("StyleGAN/ffhq.pt", len(g_sd), "entries")
    lat18 = latent_avg.unsqueeze(0).repeat(18, 1).contiguous()

    # --- ArcFace trunks (iresnet50 of models/Net.py and of FeatureStyleEncoder/arcface) ---------------------------
    from models.Net import iresnet50
    torch.manual_seed(seed + 2)
    ir_sd = _tame(_cpu_sd(iresnet50()), seed + 2)
    _save(ir_sd, os.path.join(pm, "ArcFace/backbone_ir50.pth"))
    torch.manual_seed(seed + 3)
    _save(_tame(_cpu_sd(iresnet50()), seed + 3), os.path.join(pm, "FeatureStyleEncoder/backbone.pth"))
    say("ArcFace trunks", len(ir_sd), "entries")

    # --- e4e -------------------------------------------------------------------------------------------------------
    from models.encoder4editing.models.encoders.psp_encoders import Encoder4Editing
    opts = {"encoder_type": "Encoder4Editing", "stylegan_size": 1024, "start_from_latent_avg": True,
            "input_nc": 3, "output_size": 1024, "device": "cuda", "checkpoint_path": None}
    torch.manual_seed(seed + 4)
    enc = Encoder4Editing(50, "ir_se", argparse.Namespace(**opts))
    sd = {"encoder." + k: v for k, v in _tame(_cpu_sd(enc), seed + 4).items()}
    sd.update({"decoder." + k: v for k, v in g_sd.items()})
    _save({"opts": opts, "state_dict": sd, "latent_avg": lat18}, os.path.join(pm, "encoder4editing/e4e_ffhq_encode.pt"))
    say("encoder4editing/e4e_ffhq_encode.pt", len(sd), "entries")
    del enc, sd

    # --- FeatureStyleEncoder -----------------------------------------------------------------------------------------
    _save({"state_dict": {"decoder." + k: v for k, v in g_sd.items()}, "latent_avg": lat18},
          os.path.join(pm, "FeatureStyleEncoder/psp_ffhq_encode.pt"))
    fse_dir = os.path.join(refenv.ref_root(), "models", "FeatureStyleEncoder")
    if fse_dir not in sys.path:
        sys.path.insert(0, fse_dir)                           # what FSencoder.py:12-13 does
    from nets.feature_style_encoder import fs_encoder_v2
    torch.manual_seed(seed + 5)
    fse = fs_encoder_v2(n_styles=18, opts=argparse.Namespace(
        arcface_model_path=os.path.join(pm, "FeatureStyleEncoder/backbone.pth")), residual=False, use_coeff=False,
        resnet_layer=[4, 5, 6], stride=(2, 2))
    fse_sd = _tame(_cpu_sd(fse), seed + 5)
    _save(fse_sd, os.path.join(pm, "FeatureStyleEncoder/143_enc.pth"))
    say("FeatureStyleEncoder/143_enc.pth", len(fse_sd), "entries")
    del fse, fse_sd

    # --- BiSeNet (HairFast's and FSE's copies share the architecture) --------------------------------------------------
    # Resnet18.init_weight (face_parsing/resnet.py:82-88) pulls torchvision's resnet18 from torch.hub: no network here,
    # so a seeded stand-in with torchvision's key layout is placed in the hub cache ($TORCH_HOME, set by refenv)
    import torchvision
    hub = os.path.join(os.environ["TORCH_HOME"], "hub", "checkpoints")
    torch.manual_seed(seed + 13)
    _save(_cpu_sd(torchvision.models.resnet18()), os.path.join(hub, "resnet18-5c106cde.pth"))
    from models.CtrlHair.external_code.face_parsing.model import BiSeNet
    torch.manual_seed(seed + 6)
    bs = _cpu_sd(BiSeNet(n_classes=19))
    _save(bs, os.path.join(pm, "BiSeNet/face_parsing_79999_iter.pth"))
    from face_parsing.model import BiSeNet as FseBiSeNet      # FeatureStyleEncoder/face_parsing (trainer.py:22)
    torch.manual_seed(seed + 7)
    _save(_cpu_sd(FseBiSeNet(n_classes=19)), os.path.join(pm, "FeatureStyleEncoder/79999_iter.pth"))
    say("BiSeNet", len(bs), "entries")

    # --- SEAN ------------------------------------------------------------------------------------------------------
    from models.sean_codes.models import networks
    from models.sean_codes.models.pix2pix_model import SEAN_OPT
    torch.manual_seed(seed + 8)
    netG = networks.define_G(SEAN_OPT)
    _save(_cpu_sd(netG), os.path.join(pm, "sean_checkpoints/CelebA-HQ_pretrained/latest_net_G.pth"))
    say("sean latest_net_G.pth", len(netG.state_dict()), "entries")
    del netG

    # --- CtrlHair shape adaptor ----------------------------------------------------------------------------------------
    from models.CtrlHair.shape_branch.config import cfg as cfg_mask
    from models.CtrlHair.shape_branch.model import Generator as MaskGenerator
    torch.manual_seed(seed + 9)
    mg = MaskGenerator(cfg_mask)
    _save(_cpu_sd(mg), os.path.join(pm, "ShapeAdaptor/mask_generator.pth"))
    say("ShapeAdaptor/mask_generator.pth", len(mg.state_dict()), "entries")

    # --- Rotate / Blending / PostProcess heads ---------------------------------------------------------------------------
    from models.Encoders import RotateModel, ClipBlendingModel, PostProcessModel
    torch.manual_seed(seed + 10)
    _save({"model_state_dict": _cpu_sd(RotateModel())}, os.path.join(pm, "Rotate/rotate_best.pth"))
    torch.manual_seed(seed + 11)
    cb = {k: v for k, v in _cpu_sd(ClipBlendingModel()).items() if not k.startswith("clip_model.")}
    _save({"model_state_dict": cb}, os.path.join(pm, "Blending/checkpoint.pth"))
    _save(lat18.clone(), os.path.join(pm, "PostProcess/latent_avg.pt"))
    torch.manual_seed(seed + 12)
    pp = PostProcessModel()
    _save({"model_state_dict": _tame(_cpu_sd(pp), seed + 12)}, os.path.join(pm, "PostProcess/pp_model.pth"))
    say("Rotate / Blending / PostProcess", len(pp.state_dict()), "entries")

    open(marker, "w").write("ok\n")
    return pm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workdir")
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args()
    refenv.activate(overlay=False, workdir=os.path.abspath(a.workdir))
    import torch
    if not torch.cuda.is_available():
        refenv.cpu_dryrun_patches()
    print(write_all(os.path.abspath(a.workdir), a.seed))


if __name__ == "__main__":
    main()
