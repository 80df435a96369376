"""CPU oracle for the stage glue.  TEST INFRASTRUCTURE ONLY -- same rules as stylegan2_oracle.py.

``bicubic_downsample_ref`` restates ``BicubicDownSample.forward`` (utils/bicubic.py:36-78) in matrix form: with
D[o, i] = sum of the taps k[t] whose reflect-padded position o*f + t - pad_top lands on input index i, the two padded
strided 1-D convolutions are  y = D_h @ x @ D_w^T  (vertical pass first, optional clip/round after each pass).
Pinned by tests/golden/glue.npz (oracle/gen_golden_glue.py runs the unmodified reference class)."""
from __future__ import annotations

import torch


def _cross_sum(m: torch.Tensor) -> torch.Tensor:
    """Sum over the 5-point cross with zero padding (the reference's conv2d with a 3x3 cross weight, padding='same')."""
    p = torch.nn.functional.pad(m, (1, 1, 1, 1))
    return p[..., 1:-1, 1:-1] + p[..., :-2, 1:-1] + p[..., 2:, 1:-1] + p[..., 1:-1, :-2] + p[..., 1:-1, 2:]


def dilate_erode_ref(mask: torch.Tensor, iterations: int):
    """``DilateErosion.mask`` (utils/image_utils.py:42-55): `iterations` rounds of cross sum + threshold (> 0 for the
    dilated copy, == 5 for the eroded copy), both starting from ``mask`` [N,1,H,W]."""
    grown, shrunk = mask.float().clone(), mask.float().clone()
    for _ in range(iterations):
        grown = (_cross_sum(grown) > 0).float()
        shrunk = (_cross_sum(shrunk) == 5.0).float()
    return grown, shrunk


def bicubic_taps(factor: int, a: float = -0.5) -> torch.Tensor:
    """utils/bicubic.py:7-25: Keys kernel sampled at (i - floor(2f) + 0.5) / f, i < 4f, normalised to sum 1."""
    size = 4 * factor
    xs = (torch.arange(size, dtype=torch.float32) - float(size // 2) + 0.5) / factor
    ax = xs.abs()
    k = torch.where(ax <= 1., (a + 2.) * ax ** 3 - (a + 3.) * ax ** 2 + 1,
                    torch.where(ax < 2., a * ax ** 3 - 5. * a * ax ** 2 + 8. * a * ax - 4. * a, torch.zeros_like(ax)))
    return k / k.sum()


def _reflect(i: int, n: int) -> int:
    if i < 0:
        i = -i
    if i >= n:
        i = 2 * (n - 1) - i
    return i


def decimation_matrix(n: int, factor: int, k: torch.Tensor) -> torch.Tensor:
    taps = 4 * factor
    pad = taps - factor
    p0 = pad // 2
    n_out = (n + pad - taps) // factor + 1
    d = torch.zeros(n_out, n, dtype=torch.float64)
    for o in range(n_out):
        for t in range(taps):
            d[o, _reflect(o * factor + t - p0, n)] += float(k[t])
    return d


def bicubic_downsample_ref(x: torch.Tensor, factor: int, clip_round: bool = False) -> torch.Tensor:
    k = bicubic_taps(factor)
    dh = decimation_matrix(x.shape[2], factor, k)
    dw = decimation_matrix(x.shape[3], factor, k)
    v = torch.einsum("oh,bchw->bcow", dh, x.double())
    if clip_round:
        v = torch.clamp(torch.round(v.float()), 0.0, 255.0).double()
    y = torch.einsum("pw,bcow->bcop", dw, v)
    if clip_round:
        y = torch.clamp(torch.round(y.float()), 0.0, 255.0).double()
    return y.float()


# ---------------------------------------------------------------------------------------------------------------------
# F-space alignment: models/Alignment.py:139-159 and the Embedding mixing models/Embedding.py:86-92.  These are inline
# torch expressions in the reference; restated literally.  Pinned by tests/golden/glue.npz: gen_golden_glue.py EXECUTES
# the reference's own source lines (read from the checkout at generation time) on the same inputs.
# ---------------------------------------------------------------------------------------------------------------------
def align_masks_ref(hair_mask1: torch.Tensor, hair_mask2: torch.Tensor, hair_mask_target: torch.Tensor) -> torch.Tensor:
    """models/Alignment.py:139-144."""
    return torch.cat([1 - (1 - hair_mask1) * (1 - hair_mask_target), hair_mask_target, hair_mask2 * hair_mask_target],
                     dim=0)


def align_f_space_ref(intermediate_align, latent_F_1, latent_F_out_new, latent_F_2, free_mask):
    """models/Alignment.py:153-159 (free_mask = stack(dilate[0], erosion[1], erosion[2]), :147-152)."""
    low = 1 - torch.nn.functional.interpolate(free_mask.float(), size=(32, 32), mode='bicubic')
    f = intermediate_align + low[0] * (latent_F_1 - intermediate_align)
    f = latent_F_out_new + low[1] * (f - latent_F_out_new)
    return latent_F_2 + low[2] * (f - latent_F_2)


def mix_f_space_ref(latent_F, latent_F_from_W, masks, mixing: float):
    """models/Embedding.py:86-92 (masks = the [B,1,256,256] label maps)."""
    hair = torch.where(masks == 13, torch.ones_like(masks), torch.zeros_like(masks))
    hair = torch.nn.functional.interpolate(hair.float(), size=(32, 32), mode='bicubic')
    return latent_F + mixing * hair * (latent_F_from_W - latent_F)
