"""Stand-in for OpenAI CLIP (requirements.txt:6; call sites models/Encoders.py:78,92,143), which is neither vendored in
the reference nor installed here, and is OUT of the hot-path scope.  `load()` returns a small deterministic image
encoder with CLIP's interface -- `.encode_image([B,3,224,224]) -> [B,512]`, `.parameters()` -- so that the unmodified
`ClipBlendingModel` runs.  Same seed in every process: the reference arm and the overlay arm see the same function."""
import torch
from torch import nn


class _ImageEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.pool = nn.AdaptiveAvgPool2d((16, 16))
        self.proj = nn.Linear(3 * 16 * 16, 512, bias=False)
        g = torch.Generator().manual_seed(224)
        with torch.no_grad():
            self.proj.weight.copy_(torch.randn(512, 768, generator=g) / 768 ** 0.5)

    def encode_image(self, image):
        return self.proj(self.pool(image.float()).flatten(1))


def load(name="ViT-B/32", device="cpu", jit=False, download_root=None):
    model = _ImageEncoder().to(device).eval()
    return model, (lambda img: img)


def available_models():
    return ["ViT-B/32"]


def tokenize(texts, context_length=77, truncate=False):
    raise RuntimeError("clip stub: text path is not part of HairFast inference")
