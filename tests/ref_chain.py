"""The chain of the reference's four hot-path scripts, run UNMODIFIED by piped stdin on a synthetic image tree, and the
checks of what they leave behind.  Shared by tests/test_reference_scripts.py (builder container: scripts from
/root/reference, vector arithmetic answered by the Gallery test double) and tests/test_gpu_reference_scripts.py (B200:
the same scripts shipped as job inputs, the real librbod.so underneath).

    util/qdrant_manager.py -> 31_clip_embedding_and_save_vector.py -> 32_create_delegate_vector.py
    -> 33_run_all_experiments.py
"""
import csv
import os
import subprocess
import sys
import uuid

import numpy as np

from oracle import oracle_np as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLASSES = ("cup", "dog")
PER_CLASS = 3


def make_images(base):
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


def run_script(ref_dir, script, stdin, cwd, store, runner=None, cpu_only=True):
    """One script = one process.  runner: a wrapper that installs the Gallery test double (CPU), or None to run the
    script itself against librbod.so."""
    env = dict(os.environ)
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    env["RBOD_STORE_DIR"] = str(store)
    env["RBOD_FAKE_CLIP"] = "1"            # the random-init clip stand-in is what these runs mean to use
    env["OMP_NUM_THREADS"] = "4"
    if cpu_only:
        env["CUDA_VISIBLE_DEVICES"] = ""
    cmd = [sys.executable] + ([runner] if runner else []) + [os.path.join(ref_dir, script)]
    p = subprocess.run(cmd, input=stdin, text=True, cwd=cwd, env=env, capture_output=True, timeout=900)
    assert p.returncode == 0, f"{script} failed:\n{p.stdout[-3000:]}\n{p.stderr[-3000:]}"
    return p.stdout


def run_chain(ref_dir, work, store, runner=None, cpu_only=True, log=None):
    """-> (rows of the result CSV, number of image points)."""
    def run(script, stdin):
        out = run_script(ref_dir, script, stdin, work, store, runner, cpu_only)
        if log is not None:
            log.write(f"\n===== {script} =====\n{out}")
        return out

    # ---- util/qdrant_manager.py: create "thesis" (dim/distance defaults = 512 / COSINE), a scratch collection that
    # is renamed and deleted, list, quit
    out = run("util/qdrant_manager.py",
              "\n\n" "2\nthesis\n\n\n" "2\nscratch\n16\n3\n" "3\nscratch\nscratch2\n" "1\n" "4\n1\n" "1\n" "q\n")
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
    out = run("31_clip_embedding_and_save_vector.py", stdin)
    assert out.count(f"- cup: {PER_CLASS}") == 3 and out.count(f"- dog: {PER_CLASS}") == 3
    n_points = 3 * len(CLASSES) * PER_CLASS

    # ---- 32: delegates for class 1 (cup) then class 2 (dog)
    out = run("32_create_delegate_vector.py", "\n\n" "1\n1\ny\n" "1\n2\nn\n")
    assert f"1) thesis ({n_points}" in out
    assert out.count("대표 벡터 저장 완료") == 4                  # pre_a and pre_b for both classes; pre_c has no data
    assert out.count("조건에 해당하는 벡터가 없습니다") == 2

    # ---- 33: TestGroup2 (dataset_cropped), collection 1
    out = run("33_run_all_experiments.py", "2\n\n\n1\n")
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
    return rows, n_points


def check_store(client, rows, n_points, delegate_fns, delegate_id_fn, cos, ulp_tol=0):
    """The stored state and the scripts' numbers against the reference's own functions (or their oracle restatements),
    through a fresh client: stored vectors are unit fp32, every delegate is the stored form of what the reference
    function gives on the scrolled members, every CSV score is the float64 cosine of the two stored vectors it names."""
    from qdrant_client.models import FieldCondition, Filter, MatchValue

    c = client
    assert c.count("thesis", exact=True).count == n_points + 2 * 2 * 4        # + 4 delegates x 2 cases x 2 classes
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
        for dtype, fn in delegate_fns:
            got, _ = c.scroll("thesis", limit=10, with_vectors=True, scroll_filter=Filter(must=[
                FieldCondition(key="delegate_type", match=MatchValue(value=dtype)),
                FieldCondition(key="is_delegate", match=MatchValue(value=True)),
                FieldCondition(key="class_name", match=MatchValue(value=cls)),
                FieldCondition(key="is_segmented", match=MatchValue(value=False))]))
            assert len(got) == 1
            want = O.l2_normalize_store(fn(v).astype(np.float32)[None], "f32")[0][0]
            have = np.array(got[0].vector, dtype=np.float32)
            ulp = np.abs(have.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64)).max()
            assert ulp <= ulp_tol, (cls, dtype, int(ulp))         # K1 on the device: <= 1 ulp from the oracle's rounding
            if delegate_id_fn is not None:
                assert got[0].id == str(uuid.UUID(delegate_id_fn(got[0].payload, dtype)))
    # every CSV score is the float64 cosine (33:76-77) of the two stored vectors it names
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
