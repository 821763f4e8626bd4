"""Runs the reference's four hot-path scripts UNMODIFIED against the drop-in ``qdrant_client`` / ``clip`` packages,
driven through their interactive prompts by piped stdin, on a synthetic image tree.

    util/qdrant_manager.py -> 31_clip_embedding_and_save_vector.py -> 32_create_delegate_vector.py
    -> 33_run_all_experiments.py

Builder-container test (``-m "not gpu"``): the scripts live in /root/reference, which does not exist on the GPU
box, and this container has no GPU, so vector arithmetic is answered by the Gallery test double
(tests/fakes/fake_gallery.py = the oracle) installed by tests/fakes/run_script.py.  What is proven here is the
drop-in boundary: every call the scripts make, with their argument spellings and access patterns, and the chain of
separate processes sharing one on-disk store.  tests/test_gpu_shim.py replays the same call sequences against
librbod.so on a B200.
"""
import csv
import os
import subprocess
import sys
import uuid

import numpy as np
import pytest

from oracle import oracle_np as O
from oracle import ref_loader as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
RUNNER = os.path.join(ROOT, "tests", "fakes", "run_script.py")

pytestmark = pytest.mark.skipif(not R.available(), reason="reference tree not present (builder container only)")

CLASSES = ("cup", "dog")
PER_CLASS = 3


def _make_images(base):
    from PIL import Image

    rng = np.random.default_rng(0)
    for root in ("dataset_cropped", "dataset_segmented"):
        for kind in ("original_images", "natural_images"):
            for cls in CLASSES:
                d = base / root / kind / cls
                d.mkdir(parents=True)
                for i in range(PER_CLASS):
                    arr = rng.integers(0, 255, (48, 64, 3), dtype=np.uint8)
                    Image.fromarray(arr).save(d / f"{cls}_{i}.png")


def _run(script, stdin, cwd, store):
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env["RBOD_STORE_DIR"] = str(store)
    env["CUDA_VISIBLE_DEVICES"] = ""
    env["OMP_NUM_THREADS"] = "4"
    p = subprocess.run([sys.executable, RUNNER, os.path.join(REF, script)], input=stdin, text=True, cwd=cwd, env=env,
                       capture_output=True, timeout=600)
    assert p.returncode == 0, f"{script} failed:\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    return p.stdout


def test_manager_31_32_33_run_unchanged(tmp_path):
    work, store = tmp_path / "work", tmp_path / "store"
    work.mkdir()
    _make_images(work)

    # ---- util/qdrant_manager.py: create "thesis" (dim/distance defaults = 512 / COSINE), a scratch collection that
    # is renamed and deleted, list, quit
    out = _run("util/qdrant_manager.py",
               "\n\n" "2\nthesis\n\n\n" "2\nscratch\n16\n3\n" "3\nscratch\nscratch2\n" "1\n" "4\n1\n" "1\n" "q\n",
               work, store)
    assert "'thesis' collection" in out and "'scratch' → 'scratch2'" in out and "- scratch2 (0)" in out
    assert "'scratch2' collection" in out                        # deleted by number 1 (sorted: scratch2 < thesis)
    assert out.count("- thesis (0)") == 2

    # ---- 31: embed + upsert, three passes (cropped/natural, segmented/original, segmented/natural), all classes,
    # collection 1.  (cropped/original is left out so that the pre_a delegates of 32 always inherit
    # data_type=natural_images -- 32 copies it from the first scrolled point, whose md5 id depends on the tmp path --
    # and 33 is guaranteed to find them.)
    passes = [("1", "2"), ("2", "1"), ("2", "2")]
    stdin = "\n\n" + "".join(f"{ds}\n{kind}\ny\n1\n" + ("y\n" if i < len(passes) - 1 else "n\n")
                              for i, (ds, kind) in enumerate(passes))
    out = _run("31_clip_embedding_and_save_vector.py", stdin, work, store)
    assert out.count(f"- cup: {PER_CLASS}") == 3 and out.count(f"- dog: {PER_CLASS}") == 3
    n_points = 3 * len(CLASSES) * PER_CLASS

    # ---- 32: delegates for class 1 (cup) then class 2 (dog)
    out = _run("32_create_delegate_vector.py", "\n\n" "1\n1\ny\n" "1\n2\nn\n", work, store)
    assert f"1) thesis ({n_points}" in out
    assert out.count("대표 벡터 저장 완료") == 4                  # pre_a and pre_b for both classes; pre_c has no data
    assert out.count("조건에 해당하는 벡터가 없습니다") == 2

    # ---- 33: TestGroup2 (dataset_cropped), collection 1
    out = _run("33_run_all_experiments.py", "2\n\n\n1\n", work, store)
    assert "실험 결과 저장 완료" in out
    result_csvs = list((work / "results").glob("*/result_*.csv"))
    assert len(result_csvs) == 1
    rows = list(csv.DictReader(open(result_csvs[0])))
    assert list(rows[0].keys()) == ["experiment_id", "case", "delegate_type", "image_path", "true_class",
                                   "predicted_class", "similarity_score"]          # 33:173-175
    assert len(rows) >= len(CLASSES) * PER_CLASS * 4                       # pre_a: every natural image x 4 delegate types
    assert all(r["true_class"] == r["predicted_class"] for r in rows)
    npys = sorted((result_csvs[0].parent / "score_distribution").glob("*.npy"))
    assert npys and all(np.load(p).dtype == np.float64 for p in npys)

    # ---- check the stored state and the scripts' numbers against the oracle, from a fresh "process"
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from fakes import fake_gallery

    import retrieval_based_object_detection_b200.gallery as gallery

    saved = gallery.Gallery
    gallery.Gallery = fake_gallery.FakeGallery
    os.environ["RBOD_STORE_DIR"] = str(store)
    try:
        import qdrant_client as qc
        from qdrant_client.models import FieldCondition, Filter, MatchValue

        c = qc.QdrantClient(host="localhost", port=6333)
        assert c.count("thesis", exact=True).count == n_points + 2 * 2 * 4        # + 4 delegates x 2 cases x 2 classes
        ref32 = R.delegate_module()
        for cls in CLASSES:
            members, _ = c.scroll("thesis", limit=10000, with_vectors=True, scroll_filter=Filter(must=[
                FieldCondition(key="class_name", match=MatchValue(value=cls)),
                FieldCondition(key="is_delegate", match=MatchValue(value=False)),
                FieldCondition(key="is_cropped", match=MatchValue(value=True)),
                FieldCondition(key="is_segmented", match=MatchValue(value=False)),
                FieldCondition(key="is_augmented", match=MatchValue(value=False))]))
            assert len(members) == PER_CLASS
            v = np.array([r.vector for r in members])
            assert np.allclose(np.linalg.norm(v, axis=1), 1.0, atol=1e-6)          # stored vectors are normalised
            for dtype, fn in (("average", ref32.compute_average), ("centroid", ref32.compute_centroid),
                              ("weighted", ref32.compute_weighted_average), ("medoid", ref32.compute_medoid)):
                got, _ = c.scroll("thesis", limit=10, with_vectors=True, scroll_filter=Filter(must=[
                    FieldCondition(key="delegate_type", match=MatchValue(value=dtype)),
                    FieldCondition(key="is_delegate", match=MatchValue(value=True)),
                    FieldCondition(key="class_name", match=MatchValue(value=cls)),
                    FieldCondition(key="is_segmented", match=MatchValue(value=False))]))
                assert len(got) == 1
                want = O.l2_normalize_store(fn(v).astype(np.float32)[None], "f32")[0][0]
                assert np.array_equal(np.array(got[0].vector, dtype=np.float32), want), (cls, dtype)
                expect_id = str(uuid.UUID(ref32.generate_delegate_id(got[0].payload, dtype)))
                assert got[0].id == expect_id
        # every CSV score is the float64 cosine (33:76-77) of the two stored vectors it names
        cos = R.cosine_similarity()
        for r in rows[:40]:
            test, _ = c.scroll("thesis", with_vectors=True, scroll_filter=Filter(must=[
                FieldCondition(key="img_path", match=MatchValue(value=r["image_path"])),
                FieldCondition(key="is_delegate", match=MatchValue(value=False))]))
            must = [FieldCondition(key="delegate_type", match=MatchValue(value=r["delegate_type"])),
                    FieldCondition(key="is_delegate", match=MatchValue(value=True)),
                    FieldCondition(key="class_name", match=MatchValue(value=r["true_class"])),
                    FieldCondition(key="data_type", match=MatchValue(value=test[0].payload["data_type"])),
                    FieldCondition(key="is_augmented", match=MatchValue(value=False)),
                    FieldCondition(key="is_segmented", match=MatchValue(value=(r["case"] == "pre_b")))]
            dele, _ = c.scroll("thesis", with_vectors=True, limit=1, scroll_filter=Filter(must=must))
            assert float(r["similarity_score"]) == float(cos(np.array(test[0].vector), np.array(dele[0].vector)))
        qc._close_all()
    finally:
        gallery.Gallery = saved
        os.environ.pop("RBOD_STORE_DIR", None)
