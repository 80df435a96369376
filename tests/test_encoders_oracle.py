"""CPU: pins the encoder oracle (oracle/encoders_oracle.py) against the golden vectors that
oracle/gen_golden_encoders.py produced from the unmodified reference classes, and pins the state_dict layout
of our drop-in encoder modules to the reference's (same keys / shapes, strict load)."""
import os
import types

import numpy as np
import torch

from oracle import encoders_oracle as EO

torch.set_grad_enabled(False)


def test_e4e_oracle_and_layout(golden_dir):
    import hairfastgan_b200.encoders as E
    g = np.load(os.path.join(golden_dir, "encoders.npz"))
    enc = E.Encoder4Editing(50, "ir_se", types.SimpleNamespace(stylegan_size=1024)).eval()
    params = EO.synth_params_like(enc, seed=11)
    assert len(params) == int(g["e4e_n_keys"])
    enc.load_state_dict(params, strict=True)
    x = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(12)) * 2 - 1
    w, taps = EO.e4e_ref(params, x, return_taps=True)
    assert w.shape == (2, 18, 512)
    assert float((w - torch.from_numpy(g["e4e_w"])).abs().max()) < 1e-4
    assert float((taps[23][:, ::32, ::2, ::2] - torch.from_numpy(g["e4e_c3_sub"])).abs().max()) < 1e-4


def test_fse_oracle_and_layout(golden_dir):
    import hairfastgan_b200.encoders as E
    g = np.load(os.path.join(golden_dir, "encoders.npz"))
    enc = E.fs_encoder_v2(n_styles=18, opts=None, stride=(2, 2)).eval()
    params = EO.synth_params_like(enc, seed=21)
    assert len(params) == int(g["fse_n_keys"])
    enc.load_state_dict(params, strict=True)
    x = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(22)) * 2 - 1
    lat, content = EO.fse_ref(params, x, content_stride=2)
    assert lat.shape == (2, 18, 512) and content.shape == (2, 512, 16, 16)
    assert float((lat - torch.from_numpy(g["fse_latent"])).abs().max()) < 1e-4
    assert float((content[:, ::16] - torch.from_numpy(g["fse_content_sub"])).abs().max()) < 1e-4


def test_bisenet_oracle_and_layout(golden_dir):
    """BiSeNet: oracle vs the reference golden, state_dict layout of the drop-in module."""
    import hairfastgan_b200.bisenet as B
    from oracle import bisenet_oracle as BO
    g = np.load(os.path.join(golden_dir, "bisenet.npz"))
    net = B.BiSeNet(n_classes=19).eval()
    params = EO.synth_params_like(net, seed=51)
    assert len(params) == int(g["n_keys"])
    net.load_state_dict(params, strict=True)
    x = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(52)) * 2 - 1
    out, out16, out32 = BO.bisenet_ref(params, x)
    assert out.shape == (2, 19, 256, 256)
    assert float((out[:, :, ::4, ::4] - torch.from_numpy(g["out_sub"])).abs().max()) < 2e-4
    assert float((out16[:, :, ::8, ::8] - torch.from_numpy(g["out16_sub"])).abs().max()) < 2e-4
    assert float((out32[:, :, ::8, ::8] - torch.from_numpy(g["out32_sub"])).abs().max()) < 2e-4
    low = BO.bisenet_ref(params, x, return_lowres=True)
    assert float((low[0] - torch.from_numpy(g["low_out"])).abs().max()) < 2e-4


def test_postprocess_oracle_and_layout(golden_dir):
    """PostProcess conv stack: FeatureEncoderMult(fs_layers=[9]) and FeatureiResnet."""
    import hairfastgan_b200.postprocess as P
    g = np.load(os.path.join(golden_dir, "postprocess.npz"))
    enc = P.FeatureEncoderMult(fs_layers=[9], opts=None).eval()
    params = EO.synth_params_like(enc, seed=31)
    assert len(params) == int(g["mult_n_keys"])
    enc.load_state_dict(params, strict=True)
    x = torch.rand(2, 3, 256, 256, generator=torch.Generator().manual_seed(32)) * 2 - 1
    lat, (content,) = EO.feature_encoder_mult_ref(params, x)
    assert lat.shape == (2, 18, 512) and content.shape == (2, 512, 64, 64)
    assert float((lat - torch.from_numpy(g["mult_latent"])).abs().max()) < 1e-4
    assert float((content[:, ::16, ::2, ::2] - torch.from_numpy(g["mult_content_sub"])).abs().max()) < 1e-4

    fr = P.FeatureiResnet([[1024, 2], [768, 2], [512, 2]]).eval()
    fparams = EO.synth_params_like(fr, seed=41)
    assert len(fparams) == int(g["fres_n_keys"])
    fr.load_state_dict(fparams, strict=True)
    xf = torch.randn(2, 1024, 16, 16, generator=torch.Generator().manual_seed(42))
    y = EO.feature_iresnet_ref(fparams, xf)
    assert y.shape == (2, 512, 16, 16)
    assert float((y[:, ::4] - torch.from_numpy(g["fres_out_sub"])).abs().max()) < 1e-4
    with __import__("pytest").raises(NotImplementedError):
        P.FeatureEncoderMult(fs_layers=[3], opts=None)          # 6x6 / stride-4 content conv: not on the path
