"""Drop-in for the FeatureStyleEncoder copy of the generator,
``models/FeatureStyleEncoder/pixel2style2pixel/models/stylegan2/model.py`` (Generator.forward :474-560,
get_keys :9-13): same parameters / state_dict as the main Generator plus ``features_in``,
``feature_scale`` and ``return_features``.  ``Trainer.get_image`` calls it as
``StyleGAN([w], input_is_latent=True, return_features=True, features_in=[None]*5+[fea]+[None]*12,
feature_scale=1.0)`` (trainer.py:295).  Everything else it imports from that module is re-exported from
``hairfastgan_b200.model``."""
from __future__ import annotations

from .model import *  # noqa: F401,F403  (PixelNorm, EqualLinear, StyledConv, ToRGB, ... same surface)
from . import model as _m


def get_keys(d, name):
    if "state_dict" in d:
        d = d["state_dict"]
    return {k[len(name) + 1:]: v for k, v in d.items() if k[:len(name)] == name}


class Generator(_m.Generator):
    def forward(self, styles, return_latents=False, return_features=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True, features_in=None,
                feature_scale=1.0):
        latent = self._build_latent(styles, inject_index, truncation, truncation_latent, input_is_latent)
        _, _, image, outs = self._run(latent, noise, randomize_noise, start_layer=0, end_layer=self.log_size - 2,
                                      features_in=features_in, feature_scale=feature_scale,
                                      return_features=return_features and not return_latents)
        if return_latents:
            return image, latent
        if return_features:
            return image, outs
        return image, None
