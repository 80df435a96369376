"""Make an unmodified HairFastGAN checkout use this implementation.

    import hairfastgan_b200.install as hfi
    hfi.install()                       # before `import hair_swap`
    from hair_swap import HairFast, get_parser

After ``install()`` the import statements of the reference resolve to this package:

* ``from models.stylegan2.op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d`` (models/stylegan2/model.py:11,
  models/encoder4editing/models/stylegan2/model.py:7,
  models/FeatureStyleEncoder/pixel2style2pixel/models/stylegan2/model.py:7)      -> ``hairfastgan_b200.op``
* ``from models.stylegan2.model import Generator`` (models/Net.py:9), ``PixelNorm`` (models/Encoders.py:10)
                                                                                    -> ``hairfastgan_b200.model``
* ``from pixel2style2pixel.models.stylegan2.model import Generator, get_keys`` (FeatureStyleEncoder/trainer.py:18)
                                                                                    -> ``hairfastgan_b200.fse_model``
* ``from models.encoder4editing.models.encoders import psp_encoders`` (encoder4editing/models/psp.py:6;
  ``psp_encoders.Encoder4Editing(50, 'ir_se', opts)`` :34) and ``from nets.feature_style_encoder import *``
  (FeatureStyleEncoder/trainer.py:20; ``fs_encoder_v2(...)`` :168)                  -> ``hairfastgan_b200.encoders``

Nothing in the reference tree is edited and its JIT build of the two 2019 CUDA extensions
(op/fused_act.py:10-16, op/upfirdn2d.py:10-16) never runs.
"""
from __future__ import annotations

import importlib
import sys
import types

_OPS = {
    "models.stylegan2.op": "hairfastgan_b200.op",
    "models.stylegan2.op.fused_act": "hairfastgan_b200.op.fused_act",
    "models.stylegan2.op.upfirdn2d": "hairfastgan_b200.op.upfirdn2d",
}
_GENERATORS = {
    "models.stylegan2.model": "hairfastgan_b200.model",
    # FeatureStyleEncoder puts its own directory on sys.path (FSencoder.py:12-13) and imports by these names
    "pixel2style2pixel.models.stylegan2.model": "hairfastgan_b200.fse_model",
    "models.FeatureStyleEncoder.pixel2style2pixel.models.stylegan2.model": "hairfastgan_b200.fse_model",
}
_ENCODERS = {
    "models.encoder4editing.models.encoders.psp_encoders": "hairfastgan_b200.encoders",
    "nets.feature_style_encoder": "hairfastgan_b200.encoders",
}
_created_stubs = []


def _register(ref_name: str, ours: str) -> None:
    top = ref_name.split(".")[0]
    if top in ("pixel2style2pixel", "nets"):
        # parents of these names only exist once FSencoder.py has extended sys.path; register stub packages
        # so the absolute import resolves without executing the reference copy
        parts = ref_name.split(".")
        for i in range(1, len(parts)):
            name = ".".join(parts[:i])
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
                _created_stubs.append(name)
    sys.modules[ref_name] = importlib.import_module(ours)


def install(generator: bool = True, encoders: bool = True) -> None:
    """Register the overlay.  ``generator=False`` swaps only the operator package (L1 boundary) and leaves the
    reference's own ``models/stylegan2/model.py`` classes in place on top of our ops; ``encoders=False`` keeps
    the reference's PyTorch encoders."""
    for ref_name, ours in _OPS.items():
        _register(ref_name, ours)
    if generator:
        for ref_name, ours in _GENERATORS.items():
            _register(ref_name, ours)
    if encoders:
        for ref_name, ours in _ENCODERS.items():
            _register(ref_name, ours)


def uninstall() -> None:
    for ref_name in list(_OPS) + list(_GENERATORS) + list(_ENCODERS) + _created_stubs:
        sys.modules.pop(ref_name, None)
    _created_stubs.clear()
