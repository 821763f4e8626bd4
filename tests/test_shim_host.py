"""Host side of the drop-in ``qdrant_client`` (ids, payload filters, scroll paging, staging, persistence, error
behaviour) on CPU.  Vector arithmetic is answered by the Gallery test double (tests/fakes/fake_gallery.py = the
oracle); the same scenarios run against librbod.so in tests/test_gpu_shim.py.

Behaviours and the reference call sites that rely on them are listed in SURVEY.md §8(b)/(c).
"""
import hashlib
import uuid

import numpy as np
import pytest

from oracle import oracle_np as O
from fakes import fake_gallery


@pytest.fixture()
def client(store_dir, monkeypatch):
    import retrieval_based_object_detection_b200 as pkg
    import retrieval_based_object_detection_b200.gallery as gallery

    monkeypatch.setattr(gallery, "Gallery", fake_gallery.FakeGallery)
    monkeypatch.setattr(pkg, "Gallery", fake_gallery.FakeGallery)
    from qdrant_client import QdrantClient

    return QdrantClient(host="localhost", port=6333)


def _models():
    from qdrant_client import models

    return models


def _payload(cls, i, data_type="original_images", seg=False, aug=False):
    # the 8 keys of 31_clip_embedding_and_save_vector.py:166-175
    return {"data_type": data_type, "is_cropped": True, "is_segmented": seg, "is_augmented": aug, "class_name": cls,
            "is_delegate": False, "delegate_type": None, "img_path": f"dataset_cropped/{data_type}/{cls}/{i}.png"}


def _fill(client, name="thesis", n=60, dim=512, classes=("cup", "dog", "tree")):
    m = _models()
    client.recreate_collection(collection_name=name, vectors_config=m.VectorParams(size=dim, distance=m.Distance.COSINE))
    rng = np.random.default_rng(0)
    vecs = rng.standard_normal((n, dim)).astype(np.float32) * 3.0
    ids = []
    for i in range(n):
        cls = classes[i % len(classes)]
        pl = _payload(cls, i, data_type="natural_images" if i % 2 else "original_images", seg=(i % 4 == 0))
        pid = hashlib.md5(pl["img_path"].encode()).hexdigest()            # 31:42-43
        ids.append(pid)
        client.upsert(collection_name=name, points=[m.PointStruct(id=pid, vector=vecs[i].tolist(), payload=pl)])
    return vecs, ids


def test_collection_admin_matches_qdrant_manager_calls(client):
    m = _models()
    assert client.get_collections().collections == []
    assert client.recreate_collection(collection_name="a", vectors_config=m.VectorParams(size=512, distance=m.Distance.COSINE))
    client.recreate_collection(collection_name="b", vectors_config=m.VectorParams(size=8, distance=m.Distance.DOT))
    assert [c.name for c in client.get_collections().collections] == ["a", "b"]
    assert client.get_collection("a").points_count == 0                   # util/qdrant_manager.py:46-47
    assert client.count("a", exact=True).count == 0                        # 32:67
    client.rename_collection(old_collection_name="a", new_collection_name="c")   # util/qdrant_manager.py:99
    assert [c.name for c in client.get_collections().collections] == ["b", "c"]
    with pytest.raises(Exception):
        client.get_collection("a")                                          # unknown collection raises (404)
    with pytest.raises(Exception):
        client.create_collection("b", vectors_config=m.VectorParams(size=8, distance=m.Distance.DOT))
    client.recreate_collection(collection_name="b", vectors_config=m.VectorParams(size=4, distance=m.Distance.COSINE))
    assert client.get_collection("b").config.params.vectors.size == 4      # recreate = drop + create
    assert client.delete_collection("b") and not client.delete_collection("b")
    for bad in ("", "x/y", ".."):
        with pytest.raises(ValueError):
            client.create_collection(bad, vectors_config=m.VectorParams(size=4, distance=m.Distance.COSINE))


def test_ids_are_canonical_and_upsert_overwrites(client):
    m = _models()
    client.recreate_collection(collection_name="t", vectors_config=m.VectorParams(size=4, distance=m.Distance.COSINE))
    hex_id = hashlib.md5(b"some/path.png").hexdigest()
    client.upsert("t", points=[m.PointStruct(id=hex_id, vector=[1, 0, 0, 0], payload={"v": 1})])
    client.upsert("t", points=[m.PointStruct(id=str(uuid.UUID(hex_id)), vector=[0, 2, 0, 0], payload={"v": 2})])
    client.upsert("t", points=[m.PointStruct(id=7, vector=[0, 0, 3, 0], payload=None)])
    assert client.count("t").count == 2                                    # same UUID in two spellings = one point
    recs, nxt = client.scroll("t", with_vectors=True)
    assert nxt is None and [r.id for r in recs] == [7, str(uuid.UUID(hex_id))]    # ints first, then UUIDs; hyphenated
    assert recs[1].payload == {"v": 2} and recs[0].payload == {}
    assert np.allclose(recs[1].vector, [0, 1, 0, 0]) and np.allclose(recs[0].vector, [0, 0, 1, 0])   # normalised
    for bad in ("not-a-uuid", -1, 1 << 64, 1.5, True):
        with pytest.raises(ValueError):
            client.upsert("t", points=[m.PointStruct(id=bad, vector=[1, 0, 0, 0], payload={})])
    with pytest.raises(ValueError):
        client.upsert("t", points=[m.PointStruct(id=1, vector=[1, 0, 0], payload={})])          # wrong dimension


def test_scroll_filters_paging_and_access_patterns(client):
    m = _models()
    vecs, ids = _fill(client)
    # 32:78-82 -- no filter, limit 9999, [0] index, payload.get
    results = client.scroll(collection_name="thesis", limit=9999, with_payload=True)[0]
    assert len(results) == 60
    assert sorted({p.payload.get("class_name") for p in results if p.payload.get("is_delegate") != True}) == \
        ["cup", "dog", "tree"]
    assert results[0].vector is None
    assert [r.id for r in results] == sorted((str(uuid.UUID(i)) for i in ids), key=lambda s: uuid.UUID(s).int)
    # default limit is 10 (33:96-106 relies on it returning at least the one match)
    page, nxt = client.scroll(collection_name="thesis")
    assert len(page) == 10 and nxt == results[10].id
    seen = [r.id for r in page]
    while nxt is not None:
        page, nxt = client.scroll(collection_name="thesis", offset=nxt, limit=7)
        seen += [r.id for r in page]
    assert seen == [r.id for r in results]
    # 32:123-131 -- AND of equalities, bools and strings
    flt = m.Filter(must=[m.FieldCondition(key="class_name", match=m.MatchValue(value="dog")),
                         m.FieldCondition(key="is_delegate", match=m.MatchValue(value=False)),
                         m.FieldCondition(key="is_cropped", match=m.MatchValue(value=True)),
                         m.FieldCondition(key="is_segmented", match=m.MatchValue(value=False)),
                         m.FieldCondition(key="is_augmented", match=m.MatchValue(value=False))])
    recs, _ = client.scroll(collection_name="thesis", scroll_filter=flt, with_vectors=True, with_payload=True, limit=10000)
    want = [i for i in range(60) if i % 3 == 1 and i % 4 != 0]
    assert sorted(r.payload["img_path"] for r in recs) == sorted(_payload("dog", i, "natural_images" if i % 2 else "original_images")["img_path"] for i in want)
    stored = O.l2_normalize_store(vecs, "f32")[0]
    by_path = {r.payload["img_path"]: np.array(r.vector) for r in recs}
    for i in want:
        p = _payload("dog", i, "natural_images" if i % 2 else "original_images")["img_path"]
        assert np.array_equal(by_path[p].astype(np.float32), stored[i])    # scroll returns the stored (normalised) row
        assert isinstance(recs[0].vector, list) and isinstance(recs[0].vector[0], float)
    # a None payload value never matches; True does not match 1; unknown key matches nothing
    none = m.Filter(must=[m.FieldCondition(key="delegate_type", match=m.MatchValue(value=None))])
    assert client.scroll("thesis", scroll_filter=none)[0] == []
    one = m.Filter(must=[m.FieldCondition(key="is_cropped", match=m.MatchValue(value=1))])
    assert client.scroll("thesis", scroll_filter=one)[0] == []
    assert client.scroll("thesis", scroll_filter=m.Filter(must=[m.FieldCondition(key="nope", match=m.MatchValue(value="x"))]))[0] == []
    assert client.count("thesis", count_filter=flt).count == len(want)
    with pytest.raises(ValueError):
        client.scroll("thesis", limit=0)


def test_delegate_round_trip_like_script_32_and_33(client):
    """The arithmetic chain of 32 (scroll -> numpy delegate -> upsert) and 33 (scroll limit=1 -> cosine)."""
    from oracle import ref_loader as R

    m = _models()
    vecs, ids = _fill(client)
    flt = m.Filter(must=[m.FieldCondition(key="class_name", match=m.MatchValue(value="cup")),
                         m.FieldCondition(key="is_delegate", match=m.MatchValue(value=False))])
    results, _ = client.scroll(collection_name="thesis", scroll_filter=flt, with_vectors=True, with_payload=True, limit=10000)
    vectors_np = np.array([r.vector for r in results])                     # 32:137
    assert vectors_np.dtype == np.float64 and vectors_np.shape == (20, 512)
    mean = O.compute_average(vectors_np)
    if R.available():
        assert np.array_equal(mean, R.delegate_module().compute_average(vectors_np))
    payload = {"class_name": "cup", "is_delegate": True, "delegate_type": "average"}
    pid = hashlib.md5("cup::average::x::False::False".encode()).hexdigest()          # 32:29-31
    client.upsert(collection_name="thesis", points=[m.PointStruct(id=pid, vector=mean.tolist(), payload=payload)])
    got, _ = client.scroll(collection_name="thesis", with_vectors=True, limit=1, scroll_filter=m.Filter(must=[
        m.FieldCondition(key="delegate_type", match=m.MatchValue(value="average")),
        m.FieldCondition(key="is_delegate", match=m.MatchValue(value=True)),
        m.FieldCondition(key="class_name", match=m.MatchValue(value="cup"))]))
    assert len(got) == 1
    want = O.l2_normalize_store(mean.astype(np.float32)[None], "f32")[0][0]
    assert np.array_equal(np.array(got[0].vector, dtype=np.float32), want)
    # the batched K2 entry point gives the same delegate for every class in one call
    names, cents = client.build_delegates("thesis", group_key="class_name",
                                          scroll_filter=m.Filter(must=[m.FieldCondition(key="is_delegate", match=m.MatchValue(value=False))]))
    assert names == ["cup", "dog", "tree"] and np.array_equal(cents[0], want)
    for kind, fn in (("centroid", O.compute_centroid), ("weighted", O.compute_weighted_average), ("medoid", O.compute_medoid)):
        _, vecs_k = client.build_delegates("thesis", kind=kind, scroll_filter=m.Filter(must=[
            m.FieldCondition(key="is_delegate", match=m.MatchValue(value=False))]))
        assert np.array_equal(vecs_k[0], O.l2_normalize_store(fn(vectors_np).astype(np.float32)[None], "f32")[0][0])
    # 33:151 on stored vectors: a member compared with itself gives the reference's known answer
    v = np.array(results[0].vector)
    assert O.cosine_similarity(v, v) in (1.0, 1.0000000000000002, 0.9999999999999999)


def test_search_api_filters_thresholds_and_ties(client):
    m = _models()
    vecs, ids = _fill(client)
    stored = O.l2_normalize_store(vecs, "f32")[0]
    hits = client.search("thesis", query_vector=vecs[5].tolist(), limit=3, with_vectors=True)
    assert hits[0].id == str(uuid.UUID(ids[5])) and abs(hits[0].score - 1.0) < 1e-6
    assert [h.score for h in hits] == sorted((h.score for h in hits), reverse=True)
    ws, wi = O.cosine_topk(vecs[5:6], stored, 3)
    # slots are insertion order here, so oracle row i <-> ids[i]
    assert [h.id for h in hits] == [str(uuid.UUID(ids[i])) for i in wi[0]]
    flt = m.Filter(must=[m.FieldCondition(key="class_name", match=m.MatchValue(value="tree"))])
    hits = client.query_points("thesis", query=vecs[5].tolist(), query_filter=flt, limit=50).points
    assert len(hits) == 20 and all(h.payload["class_name"] == "tree" for h in hits)
    assert client.search("thesis", vecs[5].tolist(), limit=5, score_threshold=0.999)[0].id == str(uuid.UUID(ids[5]))
    assert len(client.search("thesis", vecs[5].tolist(), limit=5, score_threshold=0.999)) == 1
    assert [h.id for h in client.search("thesis", vecs[5].tolist(), limit=2, offset=1)] == \
        [h.id for h in client.search("thesis", vecs[5].tolist(), limit=3)][1:]
    scores, idlists = client.search_batch("thesis", queries=vecs[:4], k=2)
    assert scores.shape == (4, 2) and [l[0] for l in idlists] == [str(uuid.UUID(i)) for i in ids[:4]]
    with pytest.raises(ValueError):
        client.search("thesis", [1.0, 2.0], limit=1)


def test_state_survives_processes_and_deletes(client, store_dir):
    """The scripts are separate processes talking to one server: WAL + snapshot must carry everything."""
    import qdrant_client as qc

    m = _models()
    vecs, ids = _fill(client, n=12)
    recs_before, _ = client.scroll("thesis", limit=100, with_vectors=True)
    qc._close_all()                                                        # "process exit" with a device copy: snapshot
    c2 = qc.QdrantClient(host="localhost", port=6333)
    assert c2.count("thesis").count == 12
    # a second "process" that only upserts (31_…py) never touches the device: its points live in the WAL
    c2.upsert("thesis", points=[m.PointStruct(id=3, vector=(np.arange(512) + 1.0).tolist(), payload={"class_name": "new"})])
    qc._ROOTS.clear()                                                      # drop without close(): crash-like
    c3 = qc.QdrantClient(host="localhost", port=6333)
    assert c3.count("thesis").count == 13
    recs_after, _ = c3.scroll("thesis", limit=100, with_vectors=True)
    assert recs_after[0].id == 3 and recs_after[0].payload == {"class_name": "new"}
    assert [(r.id, r.payload, r.vector) for r in recs_after[1:]] == [(r.id, r.payload, r.vector) for r in recs_before]
    c3.delete("thesis", points_selector=[3, ids[0]])
    assert c3.count("thesis").count == 11
    assert str(uuid.UUID(ids[0])) not in [r.id for r in c3.scroll("thesis", limit=100)[0]]
    qc._close_all()
    c4 = qc.QdrantClient(host="localhost", port=6333)
    left, _ = c4.scroll("thesis", limit=100, with_vectors=True)
    assert len(left) == 11 and {r.id for r in left} == {r.id for r in recs_before} - {str(uuid.UUID(ids[0]))}
    want = {r.id: r.vector for r in recs_before}
    assert all(r.vector == want[r.id] for r in left)
    # another (host, port) is another server
    assert qc.QdrantClient(host="localhost", port=7000).get_collections().collections == []


def test_single_point_upserts_are_batched_into_one_device_flush(client):
    m = _models()
    _fill(client, n=30)
    col = client._root.get("thesis")
    assert col.gallery is None and len(col.pending) == 30                  # 30 RPC-style upserts, no device work yet
    client.scroll("thesis", with_vectors=True, limit=1)                    # first read that needs vectors
    assert [c for c in col.gallery.calls if c[0] == "upsert"] == [("upsert", 30)]
    assert not col.pending


def test_batched_ingest_builds_the_scripts_ids_and_payloads(client, tmp_path):
    """§8 f3 on CPU: the batched ingest loop produces the ids (31:42-43) and payloads (31:166-175) of the script."""
    import clip
    from PIL import Image

    from retrieval_based_object_detection_b200 import ingest

    m = _models()
    rng = np.random.default_rng(0)
    d = tmp_path / "dataset_segmented" / "original_images" / "cup"
    d.mkdir(parents=True)
    for i in range(3):
        Image.fromarray(rng.integers(0, 255, (32, 32, 3), dtype=np.uint8)).save(d / f"cup_{i}.png")
    (d / "bad.jpg").write_bytes(b"garbage")
    model, preprocess = clip.load("ViT-B/32", device="cpu")
    client.recreate_collection(collection_name="t", vectors_config=m.VectorParams(size=512, distance=m.Distance.COSINE))
    counts = ingest.ingest_directory(client, "t", model, preprocess, {"cup": d}, "original", is_segmented=True,
                                     device="cpu", batch_size=2, workers=2)
    assert counts == {"cup": 3}
    recs, _ = client.scroll("t", limit=10, with_vectors=True)
    assert sorted(r.id for r in recs) == sorted(str(uuid.UUID(hashlib.md5(str((d / f"cup_{i}.png").resolve()).encode()).hexdigest()))
                                                for i in range(3))
    assert all(r.payload == {"data_type": "original_images", "is_cropped": True, "is_segmented": True, "is_augmented": False,
                             "class_name": "cup", "is_delegate": False, "delegate_type": None,
                             "img_path": r.payload["img_path"]} and r.payload["img_path"].endswith(".png") for r in recs)
    assert np.allclose(np.linalg.norm(np.array([r.vector for r in recs]), axis=1), 1.0, atol=1e-6)


def test_filter_mask_compiles_equalities_like_the_general_evaluator(client):
    """§8 f1: the vectorised column compile and the set-based evaluator give the same row bitmask."""
    m = _models()
    _fill(client, n=200)
    col = client._root.get("thesis")
    F, C, V = m.Filter, m.FieldCondition, m.MatchValue
    filters = [
        F(must=[C(key="class_name", match=V(value="dog"))]),
        F(must=[C(key="class_name", match=V(value="dog")), C(key="is_segmented", match=V(value=True))]),
        F(must=[C(key="is_cropped", match=V(value=1))]),                     # 1 is not True
        F(must=[C(key="delegate_type", match=V(value=None))]),               # None never matches
        F(must=[C(key="missing_key", match=V(value="x"))]),
        F(must=[C(key="class_name", match=V(value="unknown"))]),
        F(must=[C(key="class_name", match=m.MatchAny(any=["dog", "cup"]))]),  # general path
        F(must_not=[C(key="class_name", match=V(value="dog"))]),             # general path
    ]
    for flt in filters:
        got = col.filter_mask(flt)
        want = col.row_mask(col.filter_slots(flt))
        assert np.array_equal(got, want), flt
    assert col.filter_mask(None) is None
    client.upsert("thesis", points=[m.PointStruct(id=5, vector=np.ones(512).tolist(), payload={"class_name": "dog"})])
    got = col.filter_mask(filters[0])                                       # column cache follows mutations
    assert np.array_equal(got, col.row_mask(col.filter_slots(filters[0]))) and O.unpack_row_mask(got, len(col))[-1]


@pytest.mark.parametrize("distance,metric", [("EUCLID", "euclid"), ("MANHATTAN", "manhattan")])
def test_distance_menu_collections_rank_ascending_and_store_vectors_as_given(client, distance, metric):
    """util/qdrant_manager.py:61-79 lets the operator pick EUCLID or MANHATTAN: such a collection stores vectors as
    given, search scores are distances (ascending) and score_threshold is an upper bound."""
    m = _models()
    dim, n = 32, 40
    client.recreate_collection(collection_name="dist", vectors_config=m.VectorParams(size=dim,
                                                                                    distance=getattr(m.Distance, distance)))
    rng = np.random.default_rng(2)
    vecs = (rng.standard_normal((n, dim)) * 2.5).astype(np.float32)
    client.upsert("dist", points=[m.PointStruct(id=i, vector=vecs[i].tolist(), payload={"class_name": "a" if i % 2 else "b"})
                                  for i in range(n)])
    assert client.get_collection("dist").config.params.vectors.distance == getattr(m.Distance, distance)
    rec = client.retrieve("dist", ids=[7], with_vectors=True)[0]
    assert np.array_equal(np.asarray(rec.vector, dtype=np.float32), vecs[7])         # not normalised
    hits = client.search("dist", query_vector=vecs[7].tolist(), limit=5)
    want_d, want_i, _ = O.distance_topk(vecs[7:8], vecs, 5, metric)
    assert [h.id for h in hits] == list(want_i[0]) and hits[0].id == 7 and hits[0].score == 0.0
    assert [h.score for h in hits] == sorted(h.score for h in hits)
    assert np.allclose([h.score for h in hits], want_d[0], rtol=1e-6)
    near = client.search("dist", query_vector=vecs[7].tolist(), limit=40, score_threshold=float(want_d[0][2]) * 1.0000001)
    assert [h.id for h in near] == list(want_i[0][:3])
    flt = m.Filter(must=[m.FieldCondition(key="class_name", match=m.MatchValue(value="b"))])
    only_b = client.query_points("dist", query=vecs[7].tolist(), query_filter=flt, limit=3).points
    assert all(h.payload["class_name"] == "b" for h in only_b) and 7 not in [h.id for h in only_b]
    names, means = client.build_delegates("dist", group_key="class_name")
    assert names == ["a", "b"]
    assert np.allclose(means[0], vecs[1::2].astype(np.float64).mean(axis=0), rtol=1e-6, atol=1e-7)   # not renormalised


def test_group_rows_matches_the_per_class_loop(client):
    """Collection.group_rows (vectorised CSR for build_delegates) against the obvious loop over payloads, with a
    filter, missing / None values and mixed value types."""
    m = _models()
    client.recreate_collection(collection_name="grp", vectors_config=m.VectorParams(size=8, distance=m.Distance.COSINE))
    rng = np.random.default_rng(9)
    values = ["cup", "dog", "tree", None, 3, 7, True]
    pts, labels = [], []
    for i in range(400):
        v = values[int(rng.integers(0, len(values)))]
        pl = {"class_name": v, "is_augmented": bool(i % 3 == 0)}
        if i % 17 == 0:
            pl.pop("class_name")
            v = None
        labels.append(v)
        pid = int(rng.integers(0, 1 << 40)) if i % 2 else hashlib.md5(str(i).encode()).hexdigest()
        pts.append(m.PointStruct(id=pid, vector=rng.standard_normal(8).tolist(), payload=pl))
    client.upsert("grp", points=pts)
    col = client._root.get("grp")
    for flt in (None, m.Filter(must=[m.FieldCondition(key="is_augmented", match=m.MatchValue(value=False))])):
        names, row_idx, offsets = col.group_rows("class_name", flt)
        allowed = col.filter_slots(flt)
        groups = {}
        for s in col.ordered_slots():
            if allowed is not None and s not in allowed:
                continue
            v = col.payloads[s].get("class_name")
            if v is not None:
                groups.setdefault((type(v).__name__, v), []).append(s)
        want_names = sorted(groups)
        assert [(type(x).__name__, x) for x in names] == want_names
        assert list(row_idx) == [s for n in want_names for s in groups[n]]
        assert list(np.diff(offsets)) == [len(groups[n]) for n in want_names]
    names, means = client.build_delegates("grp", group_key="class_name")
    assert len(names) == means.shape[0] == 6 and means.shape[1] == 8
    scores, ids = client.search_batch("grp", queries=rng.standard_normal((3, 8)).astype(np.float32), k=500)
    assert len(ids) == 3 and len(ids[0]) == 500 and ids[0][399] is not None and ids[0][400] is None


def test_distance_of_a_collection_survives_reopening(client, store_dir):
    """A collection created from the distance menu keeps its distance across processes: after a snapshot and a reopen
    (memory-mapped vectors.npy streamed back in raw form) an EUCLID collection still stores vectors as given, ranks
    by ascending distance and pages its scroll by id."""
    import os

    import qdrant_client as qc

    m = _models()
    dim, n = 24, 30
    client.recreate_collection(collection_name="eu", vectors_config=m.VectorParams(size=dim, distance=m.Distance.EUCLID))
    rng = np.random.default_rng(4)
    vecs = (rng.standard_normal((n, dim)) * 3.0).astype(np.float32)
    client.upsert("eu", points=[m.PointStruct(id=100 + i, vector=vecs[i].tolist(), payload={"i": i}) for i in range(n)])
    assert client.search("eu", vecs[4].tolist(), limit=1)[0].id == 104            # materialises the device copy
    qc._close_all()                                                               # snapshot: vectors.npy + points.json
    assert os.path.exists(os.path.join(store_dir, "localhost_6333", "eu", "snap-1", "vectors.npy"))
    assert open(os.path.join(store_dir, "localhost_6333", "eu", "CURRENT")).read() == "1"
    c2 = qc.QdrantClient(host="localhost", port=6333)
    assert c2.get_collection("eu").config.params.vectors.distance == m.Distance.EUCLID
    hits = c2.search("eu", vecs[4].tolist(), limit=4, with_vectors=True)
    wd, wi, _ = O.distance_topk(vecs[4:5], vecs, 4, "euclid")
    assert [h.id for h in hits] == [100 + int(i) for i in wi[0]] and hits[0].score == 0.0
    assert np.allclose([h.score for h in hits], wd[0], rtol=1e-6)
    assert np.array_equal(np.asarray(hits[0].vector, dtype=np.float32), vecs[4])  # still not normalised
    page1, nxt = c2.scroll("eu", limit=7)
    page2, nxt2 = c2.scroll("eu", limit=7, offset=nxt)
    assert [r.id for r in page1] == list(range(100, 107)) and nxt == 107
    assert [r.id for r in page2] == list(range(107, 114)) and nxt2 == 114


def test_wal_replays_in_log_order_without_a_snapshot(client, store_dir):
    """upsert X, delete X, upsert X(v2) in processes that never save(): the reopened collection holds X(v2) -- the WAL
    is replayed strictly in order (deletes used to be queued behind every upsert, which lost X) -- and the host
    metadata is right before any vector is read."""
    import qdrant_client as qc

    m = _models()
    client.recreate_collection(collection_name="w", vectors_config=m.VectorParams(size=8, distance=m.Distance.COSINE))
    v1, v2, v3 = (np.arange(8) + 1.0), (np.arange(8)[::-1] + 1.0), np.ones(8)
    client.upsert("w", points=[m.PointStruct(id=1, vector=v1.tolist(), payload={"v": 1}),
                               m.PointStruct(id=2, vector=v3.tolist(), payload={"v": 3})])
    qc._ROOTS.clear()                                                      # crash-like: no save(), the WAL holds it all
    c2 = qc.QdrantClient(host="localhost", port=6333)
    c2.delete("w", points_selector=[1])
    c2.upsert("w", points=[m.PointStruct(id=1, vector=v2.tolist(), payload={"v": 2})])
    qc._ROOTS.clear()
    c3 = qc.QdrantClient(host="localhost", port=6333)
    col = c3._root.get("w")
    assert col.gallery is None                                             # nothing below needs the device yet
    assert c3.count("w").count == 2 and c3.get_collection("w").points_count == 2
    assert [(r.id, r.payload) for r in c3.scroll("w", limit=10)[0]] == [(1, {"v": 2}), (2, {"v": 3})]
    recs, _ = c3.scroll("w", limit=10, with_vectors=True)
    want = {1: O.l2_normalize_store(v2[None, :].astype(np.float32), "f32")[0][0], 2: O.l2_normalize_store(v3[None, :].astype(np.float32), "f32")[0][0]}
    for r in recs:
        assert np.array_equal(np.asarray(r.vector, dtype=np.float32), want[r.id])
    # ... and a plain delete survives a reopen without save() too
    c3.delete("w", points_selector=[2])
    qc._ROOTS.clear()
    c4 = qc.QdrantClient(host="localhost", port=6333)
    assert [r.id for r in c4.scroll("w", limit=10)[0]] == [1]


def test_wal_deletes_of_snapshot_rows_replay_on_the_host(client, store_dir):
    """Deletes and re-upserts journalled AFTER a snapshot: the replay permutes the slot -> snapshot-row map on the
    host, and the rows that reach the device are the right ones in the right slots."""
    import qdrant_client as qc

    m = _models()
    vecs, ids = _fill(client, n=10)
    before = {r.id: r.vector for r in client.scroll("thesis", limit=100, with_vectors=True)[0]}
    qc._close_all()                                                        # snapshot of 10 rows
    c2 = qc.QdrantClient(host="localhost", port=6333)
    c2.delete("thesis", points_selector=[ids[2], ids[9]])                  # journalled: a middle row and the last row
    newv = (np.arange(512) % 7 + 1.0).astype(np.float32)
    c2.upsert("thesis", points=[m.PointStruct(id=ids[5], vector=newv.tolist(), payload={"class_name": "re"})])   # overwrite
    c2.upsert("thesis", points=[m.PointStruct(id=77, vector=(newv * 2 + 1).tolist(), payload={"class_name": "new"})])
    c2.delete("thesis", points_selector=[ids[0]])
    qc._ROOTS.clear()                                                      # crash-like
    c3 = qc.QdrantClient(host="localhost", port=6333)
    col = c3._root.get("thesis")
    assert col.gallery is None and c3.count("thesis").count == 8
    got = {r.id: (r.vector, r.payload) for r in c3.scroll("thesis", limit=100, with_vectors=True)[0]}
    gone = {str(uuid.UUID(ids[i])) for i in (0, 2, 9)}
    assert set(got) == (set(before) - gone) | {77}
    for pid, (vec, payload) in got.items():
        if pid == str(uuid.UUID(ids[5])):
            assert payload == {"class_name": "re"}
            assert np.array_equal(np.asarray(vec, np.float32), O.l2_normalize_store(newv[None, :], "f32")[0][0])
        elif pid == 77:
            assert np.array_equal(np.asarray(vec, np.float32), O.l2_normalize_store((newv * 2 + 1)[None, :], "f32")[0][0])
        else:
            assert vec == before[pid]


def test_torn_wal_tail_is_truncated_and_later_records_survive(client, store_dir):
    """A process dies in the middle of a WAL append.  The next open drops the torn record AND truncates the file, so
    what the next process appends starts on a fresh line and is still there after yet another reopen."""
    import os

    import qdrant_client as qc

    m = _models()
    client.recreate_collection(collection_name="t", vectors_config=m.VectorParams(size=4, distance=m.Distance.COSINE))
    client.upsert("t", points=[m.PointStruct(id=1, vector=[1, 0, 0, 0], payload={})])
    qc._ROOTS.clear()
    wal = os.path.join(store_dir, "localhost_6333", "t", "wal.jsonl")
    good = os.path.getsize(wal)
    with open(wal, "ab") as f:
        f.write(b'{"op":"upsert","id":2,"payload":{},"vec":"AAAA')          # torn: no closing quote, no newline
    c2 = qc.QdrantClient(host="localhost", port=6333)
    assert c2.count("t").count == 1 and os.path.getsize(wal) == good
    c2.upsert("t", points=[m.PointStruct(id=3, vector=[0, 1, 0, 0], payload={})])
    qc._ROOTS.clear()
    c3 = qc.QdrantClient(host="localhost", port=6333)
    assert [r.id for r in c3.scroll("t", limit=10)[0]] == [1, 3]


def test_snapshot_is_one_atomic_unit(client, store_dir):
    """save() writes a fresh snap-<N> directory and switches CURRENT with one rename: a crash before the switch
    leaves the previous snapshot in force (a half-written snap directory is ignored and later overwritten)."""
    import os

    import qdrant_client as qc

    m = _models()
    _fill(client, n=6)
    client.scroll("thesis", with_vectors=True, limit=1)                     # materialise, so close() snapshots
    qc._close_all()
    base = os.path.join(store_dir, "localhost_6333", "thesis")
    assert open(os.path.join(base, "CURRENT")).read() == "1" and os.path.isdir(os.path.join(base, "snap-1"))
    # a crashed second save: snap-2 exists with a vectors file of the wrong shape, CURRENT still says 1
    os.makedirs(os.path.join(base, "snap-2"))
    np.save(os.path.join(base, "snap-2", "vectors.npy"), np.zeros((99, 512), np.float32))
    c2 = qc.QdrantClient(host="localhost", port=6333)
    assert c2.count("thesis").count == 6
    c2.upsert("thesis", points=[m.PointStruct(id=5, vector=np.ones(512).tolist(), payload={})])
    c2.scroll("thesis", with_vectors=True, limit=1)                         # materialise, so close() snapshots
    qc._close_all()
    assert open(os.path.join(base, "CURRENT")).read() == "2" and not os.path.isdir(os.path.join(base, "snap-1"))
    assert np.load(os.path.join(base, "snap-2", "vectors.npy")).shape == (7, 512)
    assert qc.QdrantClient(host="localhost", port=6333).count("thesis").count == 7


def test_list_valued_payloads_match_per_element_in_search_too(client):
    """MatchValue on a list-valued payload field matches any element (Qdrant semantics).  scroll / count / delete
    always did; the bitmask fast path of search / query_points / build_delegates must agree with them."""
    m = _models()
    client.recreate_collection(collection_name="tags", vectors_config=m.VectorParams(size=8, distance=m.Distance.COSINE))
    rng = np.random.default_rng(3)
    vecs = rng.standard_normal((6, 8)).astype(np.float32)
    tags = [["red", "round"], ["blue"], "red", ["green", "red"], None, ["blue", "round"]]
    client.upsert("tags", points=[m.PointStruct(id=i, vector=vecs[i].tolist(), payload={"tag": tags[i], "n": i})
                                  for i in range(6)])
    flt = m.Filter(must=[m.FieldCondition(key="tag", match=m.MatchValue(value="red"))])
    by_scroll = sorted(r.id for r in client.scroll("tags", scroll_filter=flt, limit=10)[0])
    assert by_scroll == [0, 2, 3] and client.count("tags", count_filter=flt).count == 3
    hits = client.search("tags", query_vector=vecs[3].tolist(), query_filter=flt, limit=10)
    assert sorted(h.id for h in hits) == by_scroll and hits[0].id == 3
    col = client._root.get("tags")
    assert np.array_equal(col.filter_mask(flt), col.row_mask(col.filter_slots(flt)))
    # a key without lists still takes the vectorised column path and agrees with the general evaluator
    flt_n = m.Filter(must=[m.FieldCondition(key="n", match=m.MatchValue(value=4))])
    assert np.array_equal(col.filter_mask(flt_n), col.row_mask(col.filter_slots(flt_n)))


def test_clip_stand_in_is_loud_and_defers_to_a_real_installation(tmp_path, monkeypatch):
    """The top-level ``clip`` directory is a random-init stand-in that sits first on sys.path.  It must say so
    (UserWarning) unless RBOD_FAKE_CLIP=1 acknowledges it, and hand ``load`` over to a real clip package found
    anywhere else on sys.path."""
    import sys
    import warnings

    import clip

    monkeypatch.setattr(clip, "_real", None)
    monkeypatch.delenv("RBOD_FAKE_CLIP", raising=False)
    with pytest.warns(UserWarning, match="RANDOM-INIT stand-in"):
        model, _ = clip.load("ViT-B/32", device="cpu")
    assert model.visual.proj.shape == (768, 512)
    monkeypatch.setenv("RBOD_FAKE_CLIP", "1")
    with warnings.catch_warnings():
        warnings.simplefilter("error")
        clip.load("ViT-B/32", device="cpu")
    # a "real" installation elsewhere on sys.path (OpenAI's layout: clip/__init__.py + clip/clip.py) wins
    real = tmp_path / "site" / "clip"
    real.mkdir(parents=True)
    (real / "clip.py").write_text("def load(name, device='cpu', jit=False, download_root=None):\n    return ('REAL', name, device)\n")
    (real / "__init__.py").write_text("from .clip import *\nfrom .clip import load\n")
    monkeypatch.setattr(clip, "_real", None)
    monkeypatch.syspath_prepend(str(tmp_path / "site"))
    assert clip.load("ViT-B/32", device="cpu") == ("REAL", "ViT-B/32", "cpu")
    sys.modules.pop("_rbod_real_clip", None)
