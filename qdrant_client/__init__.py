"""Drop-in ``qdrant_client`` for dmweapon/Retrieval_based_object_detection, backed by librbod.so.

The reference scripts do ``from qdrant_client import QdrantClient`` and talk to a Qdrant server
(util/qdrant_manager.py:1, 31_…py:19, 32_…py:5, 33_…py:12).  With this package first on
``sys.path`` the same scripts run unchanged: the "server" becomes a directory on disk
(``$RBOD_STORE_DIR/<host>_<port>``, default ``~/.rbod/store``) plus a B200-resident gallery, and
every vector operation (normalise-on-upsert, stored-vector reads, delegate means, cosine top-k)
is a call through the C ABI in include/rbod.h.  There is no CPU implementation of those
operations: without the built extension and a B200 they raise.

Method subset = what the four scripts call + ``search`` / ``query_points`` /
``search_batch`` as BASELINE.json's north star specifies.  Configuration is by environment only,
because the scripts take no flags: ``RBOD_STORE_DIR``, ``RBOD_GALLERY_DTYPE`` (f32 | bf16 | f16),
``RBOD_DEVICE`` (CUDA ordinal).
"""
from __future__ import annotations

import atexit
import os
from typing import Any, Dict, Iterable, List, Optional, Sequence

import numpy as np

from retrieval_based_object_detection_b200.store import Collection, CollectionNotFound, StoreRoot, canonical_id

from . import models
from .models import (Batch, CollectionConfig, CollectionDescription, CollectionInfo, CollectionParams,
                     CollectionsResponse, CountResult, Distance, PointStruct, QueryResponse, Record, ScoredPoint,
                     UpdateResult, UpdateStatus, VectorParams)

__all__ = ["QdrantClient", "models", "CollectionNotFound"]
__version__ = "1.9.0+rbod"

_ROOTS: Dict[str, StoreRoot] = {}


def _close_all() -> None:
    for root in list(_ROOTS.values()):
        try:
            root.close()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass
    _ROOTS.clear()


atexit.register(_close_all)


def _distance_name(d) -> str:
    if isinstance(d, Distance):
        return d.value
    for cand in Distance:
        if str(d).lower() in (cand.value.lower(), cand.name.lower()):
            return cand.value
    raise ValueError(f"unknown distance {d!r}")


class QdrantClient:
    """The client class the reference scripts instantiate as ``QdrantClient(host=..., port=...)``."""

    def __init__(self, location: Optional[str] = None, url: Optional[str] = None, port: Optional[int] = 6333,
                 host: Optional[str] = None, path: Optional[str] = None, **_ignored: Any):
        dtype = os.environ.get("RBOD_GALLERY_DTYPE", "f32")
        device = int(os.environ.get("RBOD_DEVICE", "0"))
        if location == ":memory:":
            self._root = StoreRoot(None, dtype, device)
            self._key = None
            return
        if path is not None:
            directory = os.path.abspath(path)
        else:
            if url is not None:
                hostport = url.split("://", 1)[-1].rstrip("/")
                host, _, p = hostport.partition(":")
                port = int(p) if p else port
            host = host or location or "localhost"
            if not isinstance(port, int):
                raise ValueError(f"port must be an int, got {port!r}")
            base = os.environ.get("RBOD_STORE_DIR") or os.path.join(os.path.expanduser("~"), ".rbod", "store")
            directory = os.path.join(os.path.abspath(base), f"{host}_{port}")
        self._key = directory
        root = _ROOTS.get(directory)
        if root is None:
            root = StoreRoot(directory, dtype, device)
            _ROOTS[directory] = root
        self._root = root

    # ------------------------------------------------------------------ collections
    def get_collections(self) -> CollectionsResponse:
        return CollectionsResponse(collections=[CollectionDescription(name=n) for n in self._root.names()])

    def collection_exists(self, collection_name: str) -> bool:
        return self._root.exists(collection_name)

    def get_collection(self, collection_name: str) -> CollectionInfo:
        col = self._root.get(collection_name)
        params = CollectionParams(vectors=VectorParams(size=col.dim, distance=Distance(col.distance)))
        return CollectionInfo(points_count=len(col), vectors_count=len(col), indexed_vectors_count=len(col),
                              config=CollectionConfig(params=params))

    def _vector_params(self, vectors_config) -> VectorParams:
        if isinstance(vectors_config, dict):
            if set(vectors_config) == {"size", "distance"}:
                return VectorParams(**vectors_config)
            raise ValueError("named vectors are not supported by this drop-in")
        if not hasattr(vectors_config, "size") or not hasattr(vectors_config, "distance"):
            raise ValueError("vectors_config must be VectorParams(size=..., distance=...)")
        return vectors_config

    def create_collection(self, collection_name: str, vectors_config=None, **_ignored: Any) -> bool:
        vp = self._vector_params(vectors_config)
        size = int(vp.size)
        if size < 1 or size > 65536:
            raise ValueError(f"vector size {size} out of range")
        self._root.create(collection_name, size, _distance_name(vp.distance))
        return True

    def recreate_collection(self, collection_name: str, vectors_config=None, **kwargs: Any) -> bool:
        # util/qdrant_manager.py:82-85 -- drop if present, then create empty
        self._vector_params(vectors_config)
        self._root.delete(collection_name)
        return self.create_collection(collection_name, vectors_config, **kwargs)

    def delete_collection(self, collection_name: str, **_ignored: Any) -> bool:
        return self._root.delete(collection_name)

    def rename_collection(self, old_collection_name: str, new_collection_name: str) -> bool:
        # util/qdrant_manager.py:99 calls this; the upstream client has no such method, the drop-in does
        self._root.rename(old_collection_name, new_collection_name)
        return True

    # ------------------------------------------------------------------ points
    def count(self, collection_name: str, count_filter=None, exact: bool = True, **_ignored: Any) -> CountResult:
        col = self._root.get(collection_name)
        if count_filter is None:
            return CountResult(count=len(col))
        return CountResult(count=len(col.filter_slots(count_filter)))

    def upsert(self, collection_name: str, points, wait: bool = True, **_ignored: Any) -> UpdateResult:
        col = self._root.get(collection_name)
        if isinstance(points, Batch):
            vectors = np.asarray(points.vectors, dtype=np.float32)
            col.upsert_many(points.ids, vectors, points.payloads)
        else:
            for p in points:
                vec = p.vector
                if isinstance(vec, dict):
                    raise ValueError("named vectors are not supported by this drop-in")
                col.upsert(p.id, vec, p.payload)
        return UpdateResult(operation_id=0, status=UpdateStatus.COMPLETED)

    def upsert_embeddings(self, collection_name: str, ids, embeddings, payloads=None) -> UpdateResult:
        """Batched upsert of embeddings held in a torch CUDA tensor [n, dim] (e.g. ``model.encode_image(batch)``):
        the device-to-device counterpart of the per-image ``client.upsert`` at 31_…py:178-179 (§8 f3)."""
        col = self._root.get(collection_name)
        if hasattr(embeddings, "is_cuda") and embeddings.is_cuda:
            col.upsert_device(list(ids), embeddings, payloads)
        else:
            col.upsert_many(list(ids), np.asarray(embeddings, dtype=np.float32), payloads)
        return UpdateResult(operation_id=0, status=UpdateStatus.COMPLETED)

    def upload_collection(self, collection_name: str, vectors, payload=None, ids=None, **_ignored: Any) -> None:
        """Bulk path: vectors [n, dim] (numpy), ids default to 0..n-1 appended after existing ints."""
        col = self._root.get(collection_name)
        vectors = np.asarray(vectors, dtype=np.float32)
        n = vectors.shape[0]
        if ids is None:
            start = len(col)
            ids = range(start, start + n)
        col.upsert_many(list(ids), vectors, payload)

    def delete(self, collection_name: str, points_selector, wait: bool = True, **_ignored: Any) -> UpdateResult:
        col = self._root.get(collection_name)
        if hasattr(points_selector, "points"):
            ids = list(points_selector.points)
        elif hasattr(points_selector, "filter"):
            ids = [col.ids[s] for s in col.filter_slots(points_selector.filter) or []]
        elif hasattr(points_selector, "must") or hasattr(points_selector, "should"):
            ids = [col.ids[s] for s in col.filter_slots(points_selector) or []]
        else:
            ids = list(points_selector)
        col.delete(ids)
        return UpdateResult(operation_id=0, status=UpdateStatus.COMPLETED)

    def _records(self, col: Collection, slots: Sequence[int], with_payload, with_vectors) -> List[Record]:
        vecs = col.stored_vectors(slots) if with_vectors and len(slots) else None
        out = []
        for i, s in enumerate(slots):
            out.append(Record(id=col.ids[s], payload=dict(col.payloads[s]) if with_payload else None,
                              vector=vecs[i].tolist() if vecs is not None else None))
        return out

    def scroll(self, collection_name: str, scroll_filter=None, limit: int = 10, offset=None,
               with_payload=True, with_vectors=False, **_ignored: Any):
        """-> (records, next_page_offset); call sites 32_…py:78-82,123-131 and 33_…py:96-106,139-145."""
        col = self._root.get(collection_name)
        if limit is None or int(limit) < 1:
            raise ValueError("limit must be >= 1")
        slots, nxt = col.scroll(scroll_filter, int(limit), offset)
        return self._records(col, slots, bool(with_payload), bool(with_vectors)), nxt

    def retrieve(self, collection_name: str, ids: Iterable, with_payload=True, with_vectors=False,
                 **_ignored: Any) -> List[Record]:
        col = self._root.get(collection_name)
        slots = [col.slot_of[c] for c in (canonical_id(i) for i in ids) if c in col.slot_of]
        return self._records(col, slots, bool(with_payload), bool(with_vectors))

    # ------------------------------------------------------------------ search
    def _scored(self, col: Collection, scores, slots, with_payload, with_vectors, score_threshold) -> List[ScoredPoint]:
        # EUCLID / MANHATTAN scores are distances (ascending, smaller = closer): the threshold is an upper bound
        is_distance = col.distance in ("Euclid", "Manhattan")
        keep = [(float(sc), int(s)) for sc, s in zip(scores, slots)
                if s >= 0 and (score_threshold is None
                               or (sc <= score_threshold if is_distance else sc >= score_threshold))]
        vecs = col.stored_vectors([s for _, s in keep]) if with_vectors and keep else None
        return [ScoredPoint(id=col.ids[s], version=0, score=sc,
                            payload=dict(col.payloads[s]) if with_payload else None,
                            vector=vecs[i].tolist() if vecs is not None else None)
                for i, (sc, s) in enumerate(keep)]

    def search(self, collection_name: str, query_vector, query_filter=None, limit: int = 10, offset: int = 0,
               with_payload=True, with_vectors=False, score_threshold: Optional[float] = None,
               **_ignored: Any) -> List[ScoredPoint]:
        col = self._root.get(collection_name)
        q = np.asarray(query_vector, dtype=np.float32).reshape(1, -1)
        if q.shape[1] != col.dim:
            raise ValueError(f"Wrong input: Vector dimension error: expected dim: {col.dim}, got {q.shape[1]}")
        k = int(limit) + int(offset or 0)
        if k < 1:
            return []
        scores, slots = col.search(q, k, query_filter)
        return self._scored(col, scores[0], slots[0], bool(with_payload), bool(with_vectors), score_threshold)[int(offset or 0):]

    def query_points(self, collection_name: str, query=None, query_filter=None, limit: int = 10, offset: int = 0,
                     with_payload=True, with_vectors=False, score_threshold: Optional[float] = None,
                     **_ignored: Any) -> QueryResponse:
        return QueryResponse(points=self.search(collection_name, query, query_filter, limit, offset, with_payload,
                                                with_vectors, score_threshold))

    def search_batch(self, collection_name: str, requests=None, queries=None, k: Optional[int] = None,
                     row_filter=None, **_ignored: Any):
        """Two forms: ``requests=[SearchRequest]`` -> list of ScoredPoint lists (upstream API), or the
        tensor form ``queries=[Q, dim], k=`` -> (scores [Q,k] float32, ids: list of Q lists)."""
        col = self._root.get(collection_name)
        if requests is not None:
            return [self.search(collection_name, r.vector, r.filter, r.limit, r.offset or 0,
                                bool(r.with_payload), bool(r.with_vector), r.score_threshold) for r in requests]
        if queries is None or k is None:
            raise ValueError("search_batch needs either requests= or queries= and k=")
        scores, slots = col.search(np.asarray(queries, dtype=np.float32), int(k), row_filter)
        id_of = np.asarray(col.ids + [None], dtype=object)          # slot -1 -> None through the last entry
        ids = id_of[np.asarray(slots)].tolist()
        return scores, ids

    # ------------------------------------------------------------------ delegates (K2)
    def build_delegates(self, collection_name: str, group_key: str = "class_name", scroll_filter=None,
                        kind: str = "average", alpha: float = 2.0):
        """Per-group delegate vectors of stored vectors in ONE segmented launch (K2 / K2b).

        The batched counterpart of the per-class loop in 32_create_delegate_vector.py:119-156: ``kind`` is
        "average" (compute_average :9-10), "centroid" (:12-15), "weighted" (:17-21, ``alpha``) or "medoid"
        (:23-26); the result is the stored (renormalised) form the script's upsert would leave.
        -> (group values, [G, dim] float32)."""
        col = self._root.get(collection_name)
        names, row_idx, offsets = col.group_rows(group_key, scroll_filter)
        col.flush()
        if kind == "average":
            return names, col.gallery.segment_mean(offsets, row_idx=row_idx)
        return names, col.gallery.segment_delegates(kind, offsets, row_idx=row_idx, alpha=alpha)[0]

    def close(self, **_ignored: Any) -> None:
        if self._key is None:
            self._root.close()
        else:
            for col in list(self._root.open_collections.values()):
                col.save()
