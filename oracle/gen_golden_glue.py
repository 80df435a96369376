"""Golden vectors for the stage glue: the UNMODIFIED reference ``utils/bicubic.py::BicubicDownSample``
(factor 2 and 4, as Embedding / Blending build it) on CPU -> tests/golden/glue.npz.  Build-container only; test
infrastructure (see oracle/README.md)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("HAIRFAST_REFERENCE", "/root/reference")


def main():
    torch.set_grad_enabled(False)
    sys.path.insert(0, REF)
    from utils.bicubic import BicubicDownSample
    from oracle import glue_oracle as GO
    out = {}
    x = torch.rand(2, 3, 96, 160, generator=torch.Generator().manual_seed(61)) * 2 - 1
    out["x"] = x.numpy()
    for f in (2, 4):
        ref = BicubicDownSample(factor=f, cuda=False)
        y = ref(x)
        yo = GO.bicubic_downsample_ref(x, f)
        k_ref = ref.k1[0, 0, :, 0]
        print(f"bicubic f={f}: ref vs oracle max abs", float((y - yo).abs().max()), "taps diff",
              float((k_ref - GO.bicubic_taps(f)).abs().max()), tuple(y.shape))
        out[f"y_f{f}"] = y.numpy()
        out[f"k_f{f}"] = k_ref.numpy()
    x255 = (x + 1) * 127.5
    y = BicubicDownSample(factor=4, cuda=False)(x255, clip_round=True)
    yo = GO.bicubic_downsample_ref(x255, 4, clip_round=True)
    print("bicubic f=4 clip_round: ref vs oracle max abs", float((y - yo).abs().max()))
    out["y_f4_clip_round"] = y.numpy()
    # ---- DilateErosion (utils/image_utils.py:27-55).  The module imports models.Net (-> gdown, clip, the JIT ops):
    # stub what this image lacks, nothing of it is executed by the class.
    import types
    import torch.utils.cpp_extension as ce
    ce.load = lambda *a, **k: None
    for name in ("gdown", "clip"):
        sys.modules.setdefault(name, types.ModuleType(name))
    from utils.image_utils import DilateErosion
    g = torch.Generator().manual_seed(63)
    blobs = (torch.nn.functional.avg_pool2d(torch.rand(3, 1, 64, 96, generator=g), 9, 1, 4) > 0.52).float()
    blobs[0, 0, :3, :] = 1.0                                   # touch the border: zero padding erodes it
    out["mask_in"] = blobs.numpy().astype(np.uint8)
    for it in (1, 5):
        d, e = DilateErosion(dilate_erosion=it, device="cpu").mask(blobs)
        do, eo = GO.dilate_erode_ref(blobs, it)
        print(f"dilate/erode it={it}: ref vs oracle mismatches", int((d != do).sum()), int((e != eo).sum()),
              "ones", int(blobs.sum()), int(d.sum()), int(e.sum()))
        out[f"dilate_it{it}"] = d.numpy().astype(np.uint8)
        out[f"erode_it{it}"] = e.numpy().astype(np.uint8)
    labels = torch.randint(0, 19, (2, 1, 128, 128), generator=g).float()
    d, e = DilateErosion(dilate_erosion=2, device="cpu").hair_from_mask(labels)
    out["labels"] = labels.numpy().astype(np.uint8)
    out["hair_dilate"] = d.numpy().astype(np.uint8)
    out["hair_erode"] = e.numpy().astype(np.uint8)
    np.savez_compressed(os.path.join(GOLD, "glue.npz"), **out)
    print("glue.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
