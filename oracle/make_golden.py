"""Generates tests/golden/* by running the REFERENCE's own functions (builder container only).

    python oracle/make_golden.py

Inputs are seeded; outputs come from /root/reference code imported as is
(32_create_delegate_vector.py:9-31, 31_clip_embedding_and_save_vector.py:42-43) or AST-extracted
(33_run_all_experiments.py:76-77), plus known answers read from the reference's committed run
results/2025-06-20-1.  The fixtures travel with the repo; /root/reference does not.
"""
from __future__ import annotations

import json
import os
import sys
from pathlib import Path

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import oracle_np as O  # noqa: E402
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main() -> None:
    if not R.available():
        raise SystemExit("reference tree not found; fixtures can only be regenerated in the builder container")
    os.makedirs(OUT, exist_ok=True)
    ref = R.delegate_module()
    ref_cos = R.cosine_similarity()

    # ---- delegate vectors: the reference's four functions on stored (normalised fp32 -> f64) rows
    cases = {}
    for seed, (n, dim) in enumerate([(40, 512), (7, 512), (1, 512), (93, 768), (200, 64)]):
        x, _, _ = O.synthetic_clustered(n, dim, 1, seed=100 + seed)
        stored, _ = O.l2_normalize_store(x, "f32")
        v64 = np.array([row.tolist() for row in stored])          # exactly np.array([r.vector ...]) of 32:137
        assert v64.dtype == np.float64
        cases[f"in_{seed}"] = stored
        cases[f"average_{seed}"] = ref.compute_average(v64)
        cases[f"centroid_{seed}"] = ref.compute_centroid(v64)
        cases[f"weighted_{seed}"] = ref.compute_weighted_average(v64)
        cases[f"medoid_{seed}"] = ref.compute_medoid(v64)
    np.savez_compressed(os.path.join(OUT, "delegates.npz"), **cases)

    # ---- cosine_similarity on pairs, incl. the self-match known answer
    rng = np.random.default_rng(7)
    a = O.l2_normalize_store(rng.standard_normal((64, 512)).astype(np.float32), "f32")[0].astype(np.float64)
    b = O.l2_normalize_store(rng.standard_normal((64, 512)).astype(np.float32), "f32")[0].astype(np.float64)
    b[:8] = a[:8]                                   # self matches
    b[8:16] = 3.5 * b[8:16]                         # un-normalised operand
    cos = np.array([ref_cos(a[i], b[i]) for i in range(64)])
    np.savez_compressed(os.path.join(OUT, "cosine_pairs.npz"), a=a, b=b, cos=cos)

    # ---- deterministic ids
    emb = R.embed_module()
    payloads = [
        {"class_name": "cup", "data_type": "original_images", "is_segmented": False, "is_augmented": False},
        {"class_name": "remote control", "data_type": "natural_images", "is_segmented": True, "is_augmented": False},
        {"class_name": "컵", "data_type": "original_images", "is_segmented": False, "is_augmented": True},
    ]
    ids = {"delegate": [], "path": []}
    for p in payloads:
        for t in ("average", "centroid", "weighted", "medoid"):
            ids["delegate"].append({"payload": p, "type": t, "id": ref.generate_delegate_id(p, t)})
    for s in ("/data/dataset_cropped/original_images/cup/IMG_0001.png", "/tmp/x y/컵/a.jpg"):
        ids["path"].append({"path": s, "id": emb.generate_id_from_path(Path(s))})
    with open(os.path.join(OUT, "ids.json"), "w", encoding="utf-8") as f:
        json.dump(ids, f, ensure_ascii=False, indent=1)

    # ---- known answers from the reference's committed run
    sd = os.path.join(R.REFERENCE_ROOT, "results", "2025-06-20-1", "score_distribution")
    kat = {}
    arrays = {}
    for fn in sorted(os.listdir(sd)):
        arr = np.load(os.path.join(sd, fn))
        key = fn.replace("_scores.npy", "")
        arrays[key] = arr
        kat[key] = {"n": int(arr.shape[0]), "dtype": str(arr.dtype), "min": float(arr.min()), "max": float(arr.max()),
                    "mean": float(arr.mean()), "max_hex": float(arr.max()).hex()}
    kat["centroid_equals_medoid_pre_a"] = bool(np.array_equal(arrays["pre_a_centroid"], arrays["pre_a_medoid"]))
    kat["centroid_equals_medoid_pre_b"] = bool(np.array_equal(arrays["pre_b_centroid"], arrays["pre_b_medoid"]))
    with open(os.path.join(R.REFERENCE_ROOT, "results", "2025-06-20-1", "result_2025-06-20-1.csv"), encoding="utf-8") as f:
        kat["csv_header"] = f.readline().strip()
    with open(os.path.join(OUT, "reference_run_kat.json"), "w", encoding="utf-8") as f:
        json.dump(kat, f, indent=1)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
