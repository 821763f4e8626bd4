"""Stand-in for OpenAI's ``clip`` package, which the reference imports (31_…py:18, ``clip.load``
at :26, ``model.encode_image`` at :35) but which is not installable offline and has no weights
here.

``clip.load("ViT-B/32", device)`` returns a deterministic RANDOM-INIT ViT-B/32 image tower
(same architecture and output width, 512) and the standard 224x224 preprocessing, so
31_clip_embedding_and_save_vector.py runs unchanged end to end.  The embeddings are
shape/dtype-faithful but semantically meaningless; the encoder stays in PyTorch exactly as the
north star says -- it is not part of the accelerated hot path.

This directory usually sits FIRST on sys.path (the repo root), so it would shadow a real ``clip``
installation.  ``load`` therefore looks for a real one on the rest of sys.path first and hands the
call over to it; only when there is none does it build the random tower, and it says so loudly
(``UserWarning`` + a line on stderr) unless ``RBOD_FAKE_CLIP=1`` acknowledges the stand-in.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import warnings
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


_HERE = os.path.dirname(os.path.abspath(__file__))
_real = None


def _find_real_clip():
    """The ``clip`` package of a real installation (OpenAI's: a ``clip/clip.py`` next to its ``__init__``), looked up
    on every sys.path entry except the one this stand-in lives in; None if there is none."""
    global _real
    if _real is not None:
        return _real or None
    _real = False
    for entry in sys.path:
        base = os.path.abspath(entry or os.getcwd())
        cand = os.path.join(base, "clip")
        if os.path.abspath(cand) == _HERE or not os.path.isfile(os.path.join(cand, "__init__.py")):
            continue
        if not os.path.isfile(os.path.join(cand, "clip.py")):
            continue
        spec = importlib.util.spec_from_file_location("_rbod_real_clip", os.path.join(cand, "__init__.py"),
                                                      submodule_search_locations=[cand])
        try:
            mod = importlib.util.module_from_spec(spec)
            sys.modules["_rbod_real_clip"] = mod
            spec.loader.exec_module(mod)
            _real = mod
            break
        except Exception as exc:  # noqa: BLE001 -- a broken installation must not take the scripts down
            sys.modules.pop("_rbod_real_clip", None)
            warnings.warn(f"found a clip installation at {cand} but could not import it ({exc!r}); using the stand-in")
    return _real or None


def load(name: str = "ViT-B/32", device="cpu", jit: bool = False, download_root=None) -> Tuple[CLIP, _Preprocess]:
    real = _find_real_clip()
    if real is not None:
        return real.load(name, device=device, jit=jit, download_root=download_root)
    if os.environ.get("RBOD_FAKE_CLIP", "0") != "1":
        msg = (f"clip.load({name!r}): no real `clip` installation found -- returning a RANDOM-INIT stand-in image tower "
               f"({__file__}). Embeddings have the right shape and dtype but carry NO meaning. "
               "Set RBOD_FAKE_CLIP=1 to acknowledge this and silence the warning.")
        warnings.warn(msg, UserWarning, stacklevel=2)
        print("WARNING: " + msg, file=sys.stderr)
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
