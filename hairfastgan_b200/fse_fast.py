"""Opt-in fast path for ``Trainer.test(img=..., return_latent=True)`` of the FeatureStyleEncoder.

HairFast's Embedding stage calls ``self.encoder.test(img=..., return_latent=True)`` and keeps only the last two
entries of the result, ``w_recon`` and ``fea`` (models/Embedding.py:74-76).  The reference computes them in
``Trainer.get_image`` (FeatureStyleEncoder/trainer.py:268-297) and then runs a full 1024^2 StyleGAN forward for a
reconstruction image that swap() never reads: 445.6 of the 1057.6 generator GFLOP of a triple.

``fast_test`` returns the same ``w_recon`` / ``fea`` (they do not depend on the reconstruction), ``None`` in place of the
unused image, and draws the noise tensors the skipped forward would have drawn, so every later random number of the
swap -- and therefore the final image -- is unchanged.  It only takes the short cut when the generator is this
package's (``consume_noise``), the FS-encoder configuration is active and no explicit noise / latent was passed;
otherwise it defers to the original method.  Enabled by ``hairfastgan_b200.install.install(skip_fse_reconstruction=True)``.
"""
from __future__ import annotations

import torch.nn.functional as F


def _downscale(x, times, mode):
    for _ in range(times):
        x = F.interpolate(x, scale_factor=0.5, mode=mode)          # trainer.py:61-64
    return x


def make_fast_test(original_test):
    def fast_test(self, w=None, img=None, noise=None, zero_noise_input=True, return_latent=False, training_mode=False):
        config = getattr(self, "config", {}) or {}
        gen = getattr(self, "StyleGAN", None)
        if not (return_latent and img is not None and w is None and noise is None and config.get("use_fs_encoder")
                and hasattr(gen, "consume_noise")):
            return original_test(self, w=w, img=img, noise=noise, zero_noise_input=zero_noise_input,
                                 return_latent=return_latent, training_mode=training_mode)
        if "n_iter" not in self.__dict__:
            self.n_iter = 1e5                                        # trainer.py:358-359
        w_recon, fea = self.enc(_downscale(img, self.scale, self.scale_mode))      # trainer.py:290-291
        w_recon = w_recon + self.dlatent_avg
        gen.consume_noise(img.shape[0], device=img.device)           # the RNG draws of the skipped trainer.py:295
        return [img[:, :3, :, :], None, w_recon, fea]

    fast_test.__wrapped__ = original_test
    return fast_test


def patch_trainer_module(module) -> bool:
    """Rebind ``Trainer.test`` inside the imported FeatureStyleEncoder ``trainer`` module."""
    cls = getattr(module, "Trainer", None)
    if cls is None or getattr(cls.test, "__wrapped__", None) is not None:
        return False
    cls.test = make_fast_test(cls.test)
    return True


def unpatch_trainer_module(module) -> None:
    cls = getattr(module, "Trainer", None)
    orig = getattr(getattr(cls, "test", None), "__wrapped__", None)
    if orig is not None:
        cls.test = orig
