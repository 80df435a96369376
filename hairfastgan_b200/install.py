"""Make an unmodified HairFastGAN checkout use this implementation.

    import hairfastgan_b200.install as hfi
    hfi.install()                       # before `import hair_swap`
    from hair_swap import HairFast, get_parser

After ``install()`` the import statements of the reference
(``from models.stylegan2.op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d`` in
models/stylegan2/model.py:11, models/encoder4editing/models/stylegan2/model.py:7 and
models/FeatureStyleEncoder/pixel2style2pixel/models/stylegan2/model.py:7; ``from models.stylegan2.model
import Generator`` in models/Net.py:9; ``PixelNorm`` in models/Encoders.py:10) resolve to
``hairfastgan_b200.op`` / ``hairfastgan_b200.model``.  Nothing in the reference tree is edited and its
JIT build of the two 2019 CUDA extensions (op/fused_act.py:10-16, op/upfirdn2d.py:10-16) never runs.
"""
from __future__ import annotations

import importlib
import sys

_TARGETS = {
    "models.stylegan2.op": "hairfastgan_b200.op",
    "models.stylegan2.op.fused_act": "hairfastgan_b200.op.fused_act",
    "models.stylegan2.op.upfirdn2d": "hairfastgan_b200.op.upfirdn2d",
    "models.stylegan2.model": "hairfastgan_b200.model",
    # FeatureStyleEncoder puts its own directory on sys.path (FSencoder.py:12-13) and imports the generator
    # copy as `pixel2style2pixel.models.stylegan2.model` (trainer.py:18)
    "pixel2style2pixel.models.stylegan2.model": "hairfastgan_b200.fse_model",
    "models.FeatureStyleEncoder.pixel2style2pixel.models.stylegan2.model": "hairfastgan_b200.fse_model",
}


def install(generator: bool = True) -> None:
    """Register the overlay.  ``generator=False`` swaps only the operator package (L1 boundary) and
    leaves the reference's own ``models/stylegan2/model.py`` classes in place on top of our ops."""
    for ref_name, ours in _TARGETS.items():
        if ref_name.endswith(".model") and not generator:
            continue
        if ref_name.startswith("pixel2style2pixel"):
            # parents of this name only exist once FSencoder.py has extended sys.path; register stubs so the
            # absolute import resolves without executing the reference copy
            import types
            parts = ref_name.split(".")
            for i in range(1, len(parts)):
                sys.modules.setdefault(".".join(parts[:i]), types.ModuleType(".".join(parts[:i])))
        sys.modules[ref_name] = importlib.import_module(ours)


def uninstall() -> None:
    for ref_name in _TARGETS:
        sys.modules.pop(ref_name, None)
