"""Batched CLIP ingest: decode + preprocess on host threads, ``encode_image`` on batches, and the
embeddings go straight from the encoder's CUDA output into K1 (``rbod_upsert`` with a device pointer) -- no
``.cpu().numpy().tolist()`` round trip and no per-image RPC.

The reference does this one image at a time: ``embed_image_with_clip`` (31_clip_embedding_and_save_vector.py:30-39,
batch of 1, result copied to host as a Python list) followed by one ``client.upsert`` per image (:161-179).  The CLIP
encoder itself stays plain PyTorch, as BASELINE.json's north star says; ids and payloads are built exactly as the
script builds them (:42-43, :166-175), so a collection filled by this module is indistinguishable from one filled
by the script.
"""
from __future__ import annotations

import hashlib
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
from typing import Callable, Dict, Iterable, List, Optional, Sequence


def reference_point_id(img_path) -> str:
    """generate_id_from_path (31:42-43): md5 of the resolved path."""
    return hashlib.md5(str(Path(img_path).resolve()).encode()).hexdigest()


def reference_payload(img_path, class_name: str, img_type: str, is_segmented: bool, is_augmented: bool) -> dict:
    """The 8-key payload of 31:166-175."""
    return {"data_type": f"{img_type}_images", "is_cropped": True, "is_segmented": bool(is_segmented),
            "is_augmented": bool(is_augmented), "class_name": class_name, "is_delegate": False, "delegate_type": None,
            "img_path": str(img_path)}


def embed_images(model, preprocess, paths: Sequence, device="cuda", batch_size: int = 64, workers: int = 8,
                 on_error: Optional[Callable] = None):
    """Yields (kept_paths, embeddings[b, D] on ``device``, model dtype) per batch.  Images that fail to decode
    are skipped like the reference skips them (31:31-39 returns None)."""
    import torch
    from PIL import Image

    def load(p):
        try:
            return p, preprocess(Image.open(p).convert("RGB"))
        except Exception as exc:  # unreadable image: skipped, as upstream
            if on_error is not None:
                on_error(p, exc)
            return p, None

    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        for a in range(0, len(paths), batch_size):
            loaded = list(pool.map(load, paths[a:a + batch_size]))
            keep = [(p, t) for p, t in loaded if t is not None]
            if not keep:
                continue
            batch = torch.stack([t for _, t in keep]).to(device, non_blocking=True)
            with torch.no_grad():
                emb = model.encode_image(batch)
            yield [p for p, _ in keep], emb


def ingest_directory(client, collection_name: str, model, preprocess, class_dirs: Dict[str, Path], img_type: str,
                     is_segmented: bool = False, is_augmented: bool = False, device="cuda", batch_size: int = 64,
                     workers: int = 8) -> Dict[str, int]:
    """The embedding loop of 31:161-179 for a ``{class_name: directory}`` mapping, batched.  Returns the per-class
    counts the script prints (:182-185)."""
    counts: Dict[str, int] = {}
    for cls_name, cls_path in class_dirs.items():
        files = [f for f in Path(cls_path).iterdir() if f.suffix.lower() in (".png", ".jpg", ".jpeg")]
        n = 0
        for kept, emb in embed_images(model, preprocess, files, device=device, batch_size=batch_size, workers=workers):
            ids = [reference_point_id(p) for p in kept]
            payloads = [reference_payload(p, cls_name, img_type, is_segmented, is_augmented) for p in kept]
            client.upsert_embeddings(collection_name, ids, emb, payloads)
            n += len(kept)
        counts[cls_name] = n
    return counts
