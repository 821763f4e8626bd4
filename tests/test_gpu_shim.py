"""The drop-in ``qdrant_client`` end to end on a B200: the call sequences of the reference's scripts
(util/qdrant_manager.py, 31, 32, 33 -- replayed here because /root/reference does not exist on the GPU box;
the unmodified scripts themselves run in tests/test_reference_scripts.py) with every vector operation going
through librbod.so, checked against the oracle."""
import hashlib
import uuid

import numpy as np
import pytest

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


def _payload(cls, i, data_type, seg=False, aug=False):
    return {"data_type": data_type, "is_cropped": True, "is_segmented": seg, "is_augmented": aug, "class_name": cls,
            "is_delegate": False, "delegate_type": None, "img_path": f"dataset_cropped/{data_type}/{cls}/{i}.png"}


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_script_chain_on_device(store_dir, monkeypatch, dtype):
    monkeypatch.setenv("RBOD_GALLERY_DTYPE", dtype)
    import qdrant_client as qc
    from qdrant_client.models import Distance, FieldCondition, Filter, MatchValue, PointStruct, VectorParams

    c = qc.QdrantClient(host="localhost", port=6333)
    c.recreate_collection(collection_name="thesis", vectors_config=VectorParams(size=512, distance=Distance.COSINE))
    classes = [f"class{j:02d}" for j in range(32)]                     # config C1: 1k gallery, 32 classes
    x, labels, _ = O.synthetic_clustered(1000, 512, 32, seed=0)
    x = x * 7.0                                                         # CLIP embeddings are not unit norm
    for i in range(1000):
        pl = _payload(classes[labels[i]], i, "natural_images" if i % 2 else "original_images")
        pid = hashlib.md5(pl["img_path"].encode()).hexdigest()          # 31:42-43
        c.upsert(collection_name="thesis", points=[PointStruct(id=pid, vector=x[i].tolist(), payload=pl)])   # 31:178-179
    assert c.get_collection("thesis").points_count == 1000
    stored = O.l2_normalize_store(x, dtype)[0]

    # 32: scroll class members with vectors, numpy delegates, upsert them
    cls = classes[3]
    flt = Filter(must=[FieldCondition(key="class_name", match=MatchValue(value=cls)),
                       FieldCondition(key="is_delegate", match=MatchValue(value=False)),
                       FieldCondition(key="is_cropped", match=MatchValue(value=True)),
                       FieldCondition(key="is_segmented", match=MatchValue(value=False)),
                       FieldCondition(key="is_augmented", match=MatchValue(value=False))])
    results, _ = c.scroll(collection_name="thesis", scroll_filter=flt, with_vectors=True, with_payload=True, limit=10000)
    members = np.flatnonzero(labels == 3)
    assert len(results) == len(members)
    vectors_np = np.array([r.vector for r in results])
    by_path = {r.payload["img_path"]: np.array(r.vector, dtype=np.float32) for r in results}
    for i in members:
        p = _payload(cls, i, "natural_images" if i % 2 else "original_images")["img_path"]
        ulp = np.abs(by_path[p].view(np.int32).astype(np.int64) - stored[i].view(np.int32).astype(np.int64)).max()
        assert ulp <= 1                                                 # K1 vs oracle (fp64 normalise, RNE)
    for name, fn in (("average", O.compute_average), ("centroid", O.compute_centroid),
                     ("weighted", O.compute_weighted_average), ("medoid", O.compute_medoid)):
        payload = {"class_name": cls, "data_type": "original_images", "is_segmented": False, "is_augmented": False,
                   "is_delegate": True, "delegate_type": name}
        key = f"{cls}::{name}::original_images::False::False"
        c.upsert(collection_name="thesis", points=[PointStruct(id=hashlib.md5(key.encode()).hexdigest(),
                                                               vector=fn(vectors_np).tolist(), payload=payload)])
    # 33: fetch the delegate (limit=1) and compare in float64
    test_vec = np.array(results[0].vector)
    for name, fn in (("average", O.compute_average), ("medoid", O.compute_medoid)):
        got, _ = c.scroll(collection_name="thesis", with_vectors=True, with_payload=True, limit=1, scroll_filter=Filter(must=[
            FieldCondition(key="delegate_type", match=MatchValue(value=name)),
            FieldCondition(key="is_delegate", match=MatchValue(value=True)),
            FieldCondition(key="class_name", match=MatchValue(value=cls))]))
        ref_vec = np.array(got[0].vector)
        want = O.l2_normalize_store(fn(vectors_np).astype(np.float32)[None], dtype)[0][0]
        assert np.abs(ref_vec.astype(np.float32).view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64)).max() <= 1
        s = O.cosine_similarity(test_vec, ref_vec)
        assert 0.5 < s <= 1.0000000000000002

    # K2 through the batched entry point == compute_average per class + renormalise
    names, cents = c.build_delegates("thesis", group_key="class_name",
                                     scroll_filter=Filter(must=[FieldCondition(key="is_delegate", match=MatchValue(value=False))]))
    assert names == classes
    col = c._root.get("thesis")
    stored_dev = col.stored_vectors(range(1000))
    for j in (0, 3, 31):
        rows = [col.slot_of[str(uuid.UUID(hashlib.md5(_payload(classes[j], i, "natural_images" if i % 2 else "original_images")["img_path"].encode()).hexdigest()))]
                for i in np.flatnonzero(labels == j)]
        want = O.segment_mean_renorm(stored_dev, np.array(rows), np.array([0, len(rows)]))[0]
        assert np.abs(cents[j] - want).max() <= 2e-7

    for kind, fn in (("centroid", O.compute_centroid), ("weighted", O.compute_weighted_average), ("medoid", O.compute_medoid)):
        _, vk = c.build_delegates("thesis", group_key="class_name", kind=kind,
                                  scroll_filter=Filter(must=[FieldCondition(key="is_delegate", match=MatchValue(value=False))]))
        rows = [col.slot_of[str(uuid.UUID(hashlib.md5(_payload(classes[3], i, "natural_images" if i % 2 else "original_images")["img_path"].encode()).hexdigest()))]
                for i in np.flatnonzero(labels == 3)]
        want = O.segment_mean_renorm(stored_dev, np.array(rows), np.array([0, len(rows)]), average_fn=fn)[0]
        assert np.abs(vk[3] - want).max() <= 2e-7, kind

    # north-star search API: delegate-vector top-5 (config C1) with a payload filter evaluated on the device
    only_delegates = Filter(must=[FieldCondition(key="is_delegate", match=MatchValue(value=True))])
    hits = c.search("thesis", query_vector=test_vec.tolist(), query_filter=only_delegates, limit=5)
    assert len(hits) == 4 and all(h.payload["is_delegate"] for h in hits)
    assert [h.score for h in hits] == sorted((h.score for h in hits), reverse=True)
    allowed = np.zeros(len(col), dtype=bool)
    allowed[[col.slot_of[h.id] for h in hits]] = True
    full = col.stored_vectors(range(len(col)))
    # centroid and medoid are usually the same member: identical rows must tie and resolve to the smaller slot
    ws, wi = O.cosine_topk(test_vec[None].astype(np.float32), full, 5, row_mask=allowed, rowwise=True)
    assert [col.slot_of[h.id] for h in hits] == list(wi[0][:4])
    assert np.allclose([h.score for h in hits], ws[0][:4], rtol=1e-5)
    # batched tensor entry point against the brute force over members only
    members_only = Filter(must=[FieldCondition(key="is_delegate", match=MatchValue(value=False))])
    scores, idlists = c.search_batch("thesis", queries=x[:200], k=5, row_filter=members_only)
    mask = np.array([not p.get("is_delegate") for p in col.payloads])
    ws, wi = O.cosine_topk(x[:200], full, 5, row_mask=mask)
    assert [[col.slot_of[i] for i in row] for row in idlists] == wi.tolist()
    assert np.allclose(scores, ws, rtol=1e-5)

    # a second process sees everything (snapshot + WAL), then the collection is dropped (util/qdrant_manager.py:121)
    qc._close_all()
    c2 = qc.QdrantClient(host="localhost", port=6333)
    assert c2.count("thesis", exact=True).count == 1004
    again, _ = c2.scroll(collection_name="thesis", scroll_filter=flt, with_vectors=True, limit=10000)
    assert [(r.id, r.vector) for r in again] == [(r.id, r.vector) for r in results]
    assert c2.delete_collection("thesis") and c2.get_collections().collections == []


def test_batched_clip_ingest_feeds_k1_on_device(store_dir, tmp_path):
    """§8 f3: images -> batched encode_image on the GPU -> rbod_upsert by device pointer; ids, payloads and stored
    vectors are what the per-image loop of 31_…py:161-179 would have produced."""
    import torch
    from PIL import Image

    import clip
    import qdrant_client as qc
    from qdrant_client.models import Distance, PointStruct, VectorParams
    from retrieval_based_object_detection_b200 import ingest

    rng = np.random.default_rng(0)
    dirs = {}
    for cls in ("cup", "dog"):
        d = tmp_path / "dataset_cropped" / "natural_images" / cls
        d.mkdir(parents=True)
        for i in range(5):
            Image.fromarray(rng.integers(0, 255, (40, 56, 3), dtype=np.uint8)).save(d / f"{cls}_{i}.png")
        (d / "broken.png").write_bytes(b"not an image")                 # skipped, like 31:31-39
        dirs[cls] = d
    model, preprocess = clip.load("ViT-B/32", device="cuda")
    c = qc.QdrantClient(host="localhost", port=6333)
    c.recreate_collection(collection_name="batched", vectors_config=VectorParams(size=512, distance=Distance.COSINE))
    counts = ingest.ingest_directory(c, "batched", model, preprocess, dirs, "natural", batch_size=4, workers=2)
    assert counts == {"cup": 5, "dog": 5} and c.count("batched").count == 10
    # the reference's loop, one image at a time, into a second collection
    c.recreate_collection(collection_name="single", vectors_config=VectorParams(size=512, distance=Distance.COSINE))
    for cls, d in dirs.items():
        for f in sorted(d.glob("c*_?.png")) + sorted(d.glob("d*_?.png")):
            x = preprocess(Image.open(f).convert("RGB")).unsqueeze(0).to("cuda")
            with torch.no_grad():
                v = model.encode_image(x).squeeze().cpu().numpy().tolist()
            c.upsert(collection_name="single", points=[PointStruct(id=ingest.reference_point_id(f), vector=v,
                     payload=ingest.reference_payload(f, cls, "natural", False, False))])
    a, _ = c.scroll("batched", limit=100, with_vectors=True)
    b, _ = c.scroll("single", limit=100, with_vectors=True)
    assert [(r.id, r.payload) for r in a] == [(r.id, r.payload) for r in b]
    va, vb = np.array([r.vector for r in a]), np.array([r.vector for r in b])
    assert np.abs(va - vb).max() < 2e-3            # fp16 encoder: batch-of-4 and batch-of-1 GEMMs round differently
    assert np.allclose(np.linalg.norm(va, axis=1), 1.0, atol=1e-6)
    qc._close_all()                                # device-resident upserts are snapshotted on close
    c2 = qc.QdrantClient(host="localhost", port=6333)
    a2, _ = c2.scroll("batched", limit=100, with_vectors=True)
    assert [(r.id, r.vector) for r in a2] == [(r.id, r.vector) for r in a]
