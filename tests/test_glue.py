"""Stage glue: BicubicDownSample.  CPU: oracle vs the reference golden, taps of the drop-in module;
GPU: the fused kernel vs golden and oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import glue_oracle as GO

torch.set_grad_enabled(False)


def test_bicubic_oracle_and_taps(golden_dir):
    import hairfastgan_b200.bicubic as B
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    x = torch.from_numpy(g["x"])
    for f in (2, 4):
        assert float((GO.bicubic_downsample_ref(x, f) - torch.from_numpy(g[f"y_f{f}"])).abs().max()) < 1e-6
        assert torch.equal(B.BicubicDownSample(factor=f).k, torch.from_numpy(g[f"k_f{f}"]))   # same taps, bit for bit
        assert torch.equal(GO.bicubic_taps(f), torch.from_numpy(g[f"k_f{f}"]))
    y = GO.bicubic_downsample_ref((x + 1) * 127.5, 4, clip_round=True)
    assert float((y - torch.from_numpy(g["y_f4_clip_round"])).abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        B.BicubicDownSample(factor=4)(x)
    with pytest.raises(NotImplementedError):
        B.BicubicDownSample(factor=4, padding="zeros")


def test_dilate_erode_oracle(golden_dir):
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    m = torch.from_numpy(g["mask_in"]).float()
    for it in (1, 5):
        d, e = GO.dilate_erode_ref(m, it)
        assert torch.equal(d, torch.from_numpy(g[f"dilate_it{it}"]).float())      # integer work: bit-exact
        assert torch.equal(e, torch.from_numpy(g[f"erode_it{it}"]).float())
    d0, e0 = GO.dilate_erode_ref(m, 0)
    assert torch.equal(d0, m) and torch.equal(e0, m)


@pytest.mark.gpu
def test_dilate_erosion_gpu(golden_dir):
    import hairfastgan_b200.masks as MK
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    m = torch.from_numpy(g["mask_in"]).float()
    for it in (1, 5):
        d, e = MK.DilateErosion(dilate_erosion=it).mask(m.cuda())
        assert torch.equal(d.cpu(), torch.from_numpy(g[f"dilate_it{it}"]).float())
        assert torch.equal(e.cpu(), torch.from_numpy(g[f"erode_it{it}"]).float())
    d0, e0 = MK.DilateErosion(dilate_erosion=0).mask(m.cuda())
    assert torch.equal(d0.cpu(), m) and torch.equal(e0.cpu(), m)
    labels = torch.from_numpy(g["labels"]).float()
    d, e = MK.DilateErosion(dilate_erosion=2).hair_from_mask(labels.cuda())
    assert torch.equal(d.cpu(), torch.from_numpy(g["hair_dilate"]).float())
    assert torch.equal(e.cpu(), torch.from_numpy(g["hair_erode"]).float())
    big = (torch.rand(8, 1, 256, 256, generator=torch.Generator().manual_seed(64)) > 0.7).float()   # swap() size
    d, e = MK.DilateErosion(dilate_erosion=5).mask(big.cuda())
    do, eo = GO.dilate_erode_ref(big, 5)
    assert torch.equal(d.cpu(), do) and torch.equal(e.cpu(), eo)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MK.DilateErosion(dilate_erosion=5, device="cpu").mask(big)


@pytest.mark.gpu
def test_bicubic_downsample_gpu(golden_dir):
    import hairfastgan_b200.bicubic as B
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    x = torch.from_numpy(g["x"])
    for f in (2, 4):
        y = B.BicubicDownSample(factor=f)(x.cuda()).cpu()
        assert y.shape == g[f"y_f{f}"].shape
        assert float((y - torch.from_numpy(g[f"y_f{f}"])).abs().max()) < 2e-6
    y = B.BicubicDownSample(factor=4)((x.cuda() + 1) * 127.5, clip_round=True).cpu()
    assert float((y - torch.from_numpy(g["y_f4_clip_round"])).abs().max()) <= 1.0      # a .5 tie may round either way
    assert float((y != torch.from_numpy(g["y_f4_clip_round"])).float().mean()) < 1e-3
    yb = B.BicubicDownSample(factor=4)((x.cuda() + 1) * 127.5, nhwc=False, clip_round=True, byte_output=True)
    assert yb.dtype == torch.uint8 and not yb.is_cuda
    # the swap() shapes: 1024^2 -> 512^2 / 256^2, three images, against the oracle
    big = torch.rand(3, 3, 1024, 1024, generator=torch.Generator().manual_seed(62)) * 2 - 1
    for f in (2, 4):
        y = B.BicubicDownSample(factor=f)(big.cuda()).cpu()
        assert float((y - GO.bicubic_downsample_ref(big, f)).abs().max()) < 2e-6
    xn = big[:1, :, :64, :96].permute(0, 2, 3, 1).contiguous()
    yn = B.BicubicDownSample(factor=2)(xn.cuda(), nhwc=True).cpu()
    assert yn.shape == (1, 32, 48, 3)
    assert float((yn.permute(0, 3, 1, 2) - GO.bicubic_downsample_ref(big[:1, :, :64, :96], 2)).abs().max()) < 2e-6


# ---- F-space alignment / Embedding mixing (models/Alignment.py:139-159, models/Embedding.py:86-92) --------------------
def test_fspace_oracle_matches_reference_lines(golden_dir):
    """The oracle restatement against the golden produced by EXECUTING the reference's own source lines."""
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    t = {k: torch.from_numpy(g[k]) for k in g.files if k.startswith(("fs_", "mix_"))}
    masks = GO.align_masks_ref(t["fs_hair_mask1"], t["fs_hair_mask2"], t["fs_hair_mask_target"])
    assert torch.equal(masks, t["fs_masks"].float())
    f = GO.align_f_space_ref(t["fs_intermediate_align"], t["fs_latent_F_1"], t["fs_latent_F_out_new"], t["fs_latent_F_2"],
                             t["fs_free_mask"].float())
    assert torch.equal(f, t["fs_latent_F_align"])
    m = GO.mix_f_space_ref(t["mix_latent_F"], t["mix_latent_F_from_W"], t["mix_labels"].long(), 0.95)
    assert torch.equal(m, t["mix_out"])


@pytest.mark.gpu
def test_fspace_kernels_gpu(golden_dir):
    """hf_align_masks_f32 (bit-exact: 0/1 arithmetic) and hf_fspace_blend_f32 (<= 1e-6: fp32, fma contraction only)
    against the reference-generated golden; plus a ragged size against the oracle."""
    import hairfastgan_b200.fspace as FS
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    t = {k: torch.from_numpy(g[k]).cuda() for k in g.files if k.startswith(("fs_", "mix_"))}
    masks = FS.align_masks(t["fs_hair_mask1"], t["fs_hair_mask2"], t["fs_hair_mask_target"])
    assert torch.equal(masks.cpu(), torch.from_numpy(g["fs_masks"]).float())
    f = FS.align_f_space(t["fs_intermediate_align"], t["fs_latent_F_1"], t["fs_latent_F_out_new"], t["fs_latent_F_2"],
                         t["fs_free_mask"].float())
    err = float((f.cpu() - torch.from_numpy(g["fs_latent_F_align"])).abs().max())
    assert f.shape == (1, 64, 32, 32) and err <= 2e-6, err
    hair = (t["mix_labels"] == 13).float()
    for b in range(2):
        m = FS.mix_f_space(t["mix_latent_F"][b:b + 1], t["mix_latent_F_from_W"][b:b + 1], hair[b], 0.95)
        e = float((m.cpu()[0] - torch.from_numpy(g["mix_out"])[b]).abs().max())
        assert e <= 2e-6, e
    # ragged: 3 channels, 20x12 output from a 50x70 mask (clamped taps at the borders, scalar tail path)
    gen = torch.Generator().manual_seed(5)
    a, b2 = torch.randn(1, 3, 20, 12, generator=gen), torch.randn(1, 3, 20, 12, generator=gen)
    mk = torch.rand(1, 1, 50, 70, generator=gen)
    want = b2 + (1 - torch.nn.functional.interpolate(mk, size=(20, 12), mode="bicubic")) * (a - b2)
    got = FS.fspace_blend(a.cuda(), [(b2.cuda(), mk.cuda(), 1.0, -1.0)])
    assert float((got.cpu() - want).abs().max()) <= 2e-6
    with pytest.raises(RuntimeError):
        FS.align_masks(t["fs_hair_mask1"].cpu(), t["fs_hair_mask2"].cpu(), t["fs_hair_mask_target"].cpu())
