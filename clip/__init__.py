"""Stand-in for OpenAI's ``clip`` package, which the reference imports (31_…py:18, ``clip.load``
at :26, ``model.encode_image`` at :35) but which is not installable offline and has no weights
here.

``clip.load("ViT-B/32", device)`` returns a deterministic RANDOM-INIT ViT-B/32 image tower
(same architecture and output width, 512) and the standard 224x224 preprocessing, so
31_clip_embedding_and_save_vector.py runs unchanged end to end.  The embeddings are
shape/dtype-faithful but semantically meaningless; the encoder stays in PyTorch exactly as the
north star says -- it is not part of the accelerated hot path.  If the real ``clip`` package is
installed ahead of this directory on sys.path, it is used instead.
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

_MODELS = {
    # name: (width, layers, heads, patch, embed_dim, image_size)
    "ViT-B/32": (768, 12, 12, 32, 512, 224),
    "ViT-B/16": (768, 12, 12, 16, 512, 224),
    "ViT-L/14": (1024, 24, 16, 14, 768, 224),
}
_MEAN = (0.48145466, 0.4578275, 0.40821073)
_STD = (0.26862954, 0.26130258, 0.27577711)


def available_models() -> List[str]:
    return list(_MODELS)


class _Block(nn.Module):
    def __init__(self, width: int, heads: int):
        super().__init__()
        self.ln_1 = nn.LayerNorm(width)
        self.attn = nn.MultiheadAttention(width, heads, batch_first=True)
        self.ln_2 = nn.LayerNorm(width)
        self.mlp = nn.Sequential(nn.Linear(width, 4 * width), nn.GELU(), nn.Linear(4 * width, width))

    def forward(self, x):
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class _VisionTower(nn.Module):
    def __init__(self, width, layers, heads, patch, embed_dim, image_size):
        super().__init__()
        self.input_resolution = image_size
        self.conv1 = nn.Conv2d(3, width, patch, patch, bias=False)
        n_tokens = (image_size // patch) ** 2 + 1
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(n_tokens, width))
        self.ln_pre = nn.LayerNorm(width)
        self.blocks = nn.ModuleList([_Block(width, heads) for _ in range(layers)])
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, embed_dim))

    def forward(self, x):
        x = self.conv1(x).flatten(2).transpose(1, 2)
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        for blk in self.blocks:
            x = blk(x)
        return self.ln_post(x[:, 0]) @ self.proj.to(x.dtype)


class CLIP(nn.Module):
    def __init__(self, name: str):
        super().__init__()
        self.visual = _VisionTower(*_MODELS[name])

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.to(self.dtype))


class _Preprocess:
    """Resize(bicubic, shorter side) -> CenterCrop -> RGB -> tensor -> Normalize, as upstream."""

    def __init__(self, size: int):
        self.size = size

    def __call__(self, image):
        from PIL import Image
        import numpy as np

        image = image.convert("RGB")
        w, h = image.size
        s = self.size / min(w, h)
        nw, nh = max(self.size, round(w * s)), max(self.size, round(h * s))
        image = image.resize((nw, nh), Image.BICUBIC)
        left, top = (nw - self.size) // 2, (nh - self.size) // 2
        image = image.crop((left, top, left + self.size, top + self.size))
        x = torch.from_numpy(np.asarray(image, dtype=np.float32) / 255.0).permute(2, 0, 1)
        mean = torch.tensor(_MEAN).view(3, 1, 1)
        std = torch.tensor(_STD).view(3, 1, 1)
        return (x - mean) / std


def load(name: str = "ViT-B/32", device="cpu", jit: bool = False, download_root=None) -> Tuple[CLIP, _Preprocess]:
    if name not in _MODELS:
        raise RuntimeError(f"Model {name} not found; available models = {available_models()}")
    gen_state = torch.random.get_rng_state()
    torch.manual_seed(0xC11B)
    try:
        model = CLIP(name)
    finally:
        torch.random.set_rng_state(gen_state)
    model.eval()
    device = torch.device(device)
    if device.type == "cuda":
        model = model.half()   # upstream keeps fp16 weights on CUDA, fp32 on CPU
    model = model.to(device)
    for p in model.parameters():
        p.requires_grad_(False)
    return model, _Preprocess(_MODELS[name][5])


def tokenize(texts, context_length: int = 77, truncate: bool = False):
    raise NotImplementedError("the stand-in provides the image tower only (the reference never tokenizes)")
