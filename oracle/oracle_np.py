"""CPU oracle for the retrieval hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the product (``retrieval_based_object_detection_b200``,
``qdrant_client``) never does and has no CPU path.

It restates, in float64 numpy, the arithmetic of the reference scripts for this path
(paths relative to the reference repository):

* ``compute_average``          32_create_delegate_vector.py:9-10
* ``compute_centroid``         32_create_delegate_vector.py:12-15
* ``compute_weighted_average`` 32_create_delegate_vector.py:17-21
* ``compute_medoid``           32_create_delegate_vector.py:23-26
* ``cosine_similarity``        33_run_all_experiments.py:76-77

Pinning: these five restatements are checked bit-for-bit against the reference's own functions
(imported / AST-extracted from /root/reference by ``oracle/make_golden.py``; vectors committed under
``tests/golden/``), and against the known answers in the reference's committed run
``results/2025-06-20-1`` (self-match cosine == 1.0000000000000002, centroid == medoid arrays).

PARITY UNPINNED for the third-party piece: the normalise-on-upsert of a ``Distance.COSINE``
collection and the ordering/top-k of a Qdrant ``search`` live in ``qdrant-client`` / the
``qdrant/qdrant`` server (both unpinned in the reference: 31_clip_embedding_and_save_vector.py:1,
02_qdrant_environment_setting.txt:2-10) and are not installable offline.  ``l2_normalize_store`` and
``cosine_topk`` restate their documented behaviour (store v/||v||_2 as float32; score = cosine,
descending) with the tie rule this project defines (smaller id first).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------------------------------------
# 16-bit storage formats
# ---------------------------------------------------------------------------------------------

def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest bfloat16 (ties to even), returned as float32 values."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    b = x.view(np.uint32).astype(np.uint64)
    nan = np.isnan(x)
    rounded = ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    out = rounded.view(np.float32).copy()
    out[nan] = np.nan
    return out


def round_to_f16(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest float16 (ties to even), returned as float32 values."""
    with np.errstate(over="ignore"):
        return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def round_store(x: np.ndarray, dtype: str) -> np.ndarray:
    if dtype == "f32":
        return np.asarray(x, dtype=np.float32)
    if dtype == "bf16":
        return round_to_bf16(x)
    if dtype == "f16":
        return round_to_f16(x)
    raise ValueError(f"unknown store dtype {dtype!r}")


# ---------------------------------------------------------------------------------------------
# K1: normalise on upsert (Qdrant COSINE collection behaviour; third party, see module docstring)
# ---------------------------------------------------------------------------------------------

def l2_normalize_store(x: np.ndarray, dtype: str = "f32"):
    """Rows of ``x`` (float32) as a COSINE collection stores them.

    s = sum x_i^2 in float64; r = 1/sqrt(s) (0 for a zero row, which stays zero);
    y_i = float32(float64(x_i) * r); 16-bit stores keep RNE(y_i).
    Returns (stored rows widened to float32, float32 input norms).
    """
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    x64 = x.astype(np.float64)
    s = np.einsum("ij,ij->i", x64, x64)
    with np.errstate(divide="ignore"):
        r = np.where(s > 0.0, 1.0 / np.sqrt(s), 0.0)
    y = (x64 * r[:, None]).astype(np.float32)
    return round_store(y, dtype), np.sqrt(s).astype(np.float32)


# ---------------------------------------------------------------------------------------------
# delegate vectors (32_create_delegate_vector.py:9-26)
# ---------------------------------------------------------------------------------------------

def compute_average(vectors):
    return np.mean(vectors, axis=0)


def compute_centroid(vectors):
    avg = compute_average(vectors)
    distances = np.linalg.norm(vectors - avg, axis=1)
    return vectors[np.argmin(distances)]


def compute_weighted_average(vectors, alpha=2.0):
    mean_vec = compute_average(vectors)
    weights = np.exp(-alpha * np.linalg.norm(vectors - mean_vec, axis=1))
    weights /= np.sum(weights)
    return np.sum(vectors * weights[:, np.newaxis], axis=0)


def compute_medoid(vectors):
    distances = np.linalg.norm(vectors[:, np.newaxis] - vectors, axis=2)
    total_distances = np.sum(distances, axis=1)
    return vectors[np.argmin(total_distances)]


def segment_mean_renorm(stored: np.ndarray, row_idx, offsets, average_fn=compute_average,
                        normalize: bool = True) -> np.ndarray:
    """K2: per class c, the stored form of compute_average(rows of class c).

    Exactly what the reference does per class: scroll the class's stored vectors into a float64
    array (32:137), compute_average (32:9-10), upsert the mean (32:41-42) -- which the COSINE
    collection stores L2-normalised (float32).  Empty classes give zero rows.  ``normalize=False``: the
    collection is not COSINE, the mean is stored as float32 as given.
    """
    stored = np.asarray(stored, dtype=np.float32)
    offsets = np.asarray(offsets, dtype=np.int64)
    C = len(offsets) - 1
    out = np.zeros((C, stored.shape[1]), dtype=np.float32)
    for c in range(C):
        a, b = int(offsets[c]), int(offsets[c + 1])
        if b <= a:
            continue
        rows = np.arange(a, b) if row_idx is None else np.asarray(row_idx[a:b], dtype=np.int64)
        vectors_np = stored[rows].astype(np.float64)          # np.array([r.vector ...]) is float64
        mean = average_fn(vectors_np)
        m32 = mean.astype(np.float32)
        out[c] = l2_normalize_store(m32[None, :], "f32")[0][0] if normalize else m32
    return out


# ---------------------------------------------------------------------------------------------
# cosine (33_run_all_experiments.py:76-77) and its Q x N top-k generalisation
# ---------------------------------------------------------------------------------------------

def cosine_similarity(a, b):
    return np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b))


def cosine_matrix(queries: np.ndarray, stored: np.ndarray, rowwise: bool = False, _g64=None) -> np.ndarray:
    """cosine_similarity for every (query, stored row) pair, float64.

    ``rowwise=True`` evaluates each row with the same operation sequence, so bitwise-identical
    rows get bitwise-identical scores (needed when a test plants duplicates); the default uses one
    float64 GEMM.  A zero-norm operand scores 0 (the reference formula would give nan).
    ``_g64`` = (float64 copy of ``stored``, its row norms), so a chunked caller widens the gallery once.
    """
    q = np.atleast_2d(np.asarray(queries, dtype=np.float32)).astype(np.float64)
    qn = np.sqrt(np.einsum("ij,ij->i", q, q))
    if _g64 is not None:
        g, gn = _g64
    else:
        g = np.atleast_2d(np.asarray(stored, dtype=np.float32)).astype(np.float64)
        gn = np.sqrt(np.einsum("ij,ij->i", g, g))
    if rowwise:
        dots = np.stack([(g * q[i][None, :]).sum(axis=1) for i in range(q.shape[0])])
    else:
        dots = q @ g.T
    den = qn[:, None] * gn[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(den > 0.0, dots / den, 0.0)


def topk_from_scores(scores: np.ndarray, k: int, ids=None, row_mask=None):
    """Top-k per row of ``scores`` ordered by (score desc, id asc).  Short rows pad with (-inf, -1)."""
    Q, N = scores.shape
    ids = np.arange(N, dtype=np.int64) if ids is None else np.asarray(ids, dtype=np.int64)
    out_s = np.full((Q, k), -np.inf, dtype=np.float64)
    out_i = np.full((Q, k), -1, dtype=np.int64)
    allowed = np.ones(N, dtype=bool) if row_mask is None else np.asarray(row_mask, dtype=bool)
    cols = np.nonzero(allowed)[0]
    col_ids = ids[cols]
    for qi in range(Q):
        s = scores[qi, cols]
        if len(s) > 8 * k:
            # same result as sorting everything: keep the rows tied with or above the k-th largest score
            kth = np.partition(s, len(s) - k)[len(s) - k]
            cand = np.nonzero(s >= kth)[0]
            order = cand[np.lexsort((col_ids[cand], -s[cand]))][:k]
        else:
            order = np.lexsort((col_ids, -s))[:k]
        out_s[qi, : len(order)] = s[order]
        out_i[qi, : len(order)] = col_ids[order]
    return out_s, out_i


def cosine_topk(queries, stored, k: int, row_mask=None, rowwise: bool = False, chunk: int = 256):
    """Exact float64 brute-force cosine top-k on the stored values.  Returns (scores f64, rows i64)."""
    queries = np.atleast_2d(np.asarray(queries, dtype=np.float32))
    Q = queries.shape[0]
    out_s = np.empty((Q, k), dtype=np.float64)
    out_i = np.empty((Q, k), dtype=np.int64)
    g64 = np.atleast_2d(np.asarray(stored, dtype=np.float32)).astype(np.float64)
    g64 = (g64, np.sqrt(np.einsum("ij,ij->i", g64, g64)))
    for a in range(0, Q, chunk):
        sc = cosine_matrix(queries[a : a + chunk], stored, rowwise=rowwise, _g64=g64)
        out_s[a : a + chunk], out_i[a : a + chunk] = topk_from_scores(sc, k, row_mask=row_mask)
    return out_s, out_i


def distance_topk(queries, stored, k: int, metric: str, row_mask=None):
    """Exact float64 brute force for the two distances of the collection menu that are not inner products
    (util/qdrant_manager.py:61-66): metric "euclid" -> sqrt(sum (q-g)^2), "manhattan" -> sum |q-g|, on the
    stored values (these collections store vectors as given).  Ordered by (ordering key desc, row asc) with
    key = -(squared L2) / -(L1), i.e. distance ascending, ties to the smaller row.  Qdrant-side semantics
    (score = distance, smaller is closer) are third-party behaviour: parity unpinned, see the module header.
    Returns (distances f64 [Q,k] padded with +inf, rows i64 [Q,k] padded with -1, keys f64 [Q,k])."""
    if metric not in ("euclid", "manhattan"):
        raise ValueError(metric)
    q = np.atleast_2d(np.asarray(queries, dtype=np.float32)).astype(np.float64)
    g = np.atleast_2d(np.asarray(stored, dtype=np.float32)).astype(np.float64)
    Q = q.shape[0]
    dist = np.full((Q, k), np.inf)
    rows = np.full((Q, k), -1, dtype=np.int64)
    keys = np.full((Q, k), -np.inf)
    for i in range(Q):
        d = g - q[i][None, :]
        key = -(d * d).sum(axis=1) if metric == "euclid" else -np.abs(d).sum(axis=1)
        s, r = topk_from_scores(key[None, :], k, row_mask=row_mask)
        keys[i], rows[i] = s[0], r[0]
        ok = r[0] >= 0
        dist[i, ok] = np.sqrt(-s[0][ok]) if metric == "euclid" else -s[0][ok]
    return dist, rows, keys


def unpack_row_mask(words: np.ndarray, n_rows: int) -> np.ndarray:
    """uint32 bitmask words (bit r%32 of word r//32) -> bool[n_rows]."""
    words = np.asarray(words, dtype=np.uint32)
    bits = np.unpackbits(words.view(np.uint8), bitorder="little")
    return bits[:n_rows].astype(bool)


def pack_row_mask(allowed: np.ndarray) -> np.ndarray:
    allowed = np.asarray(allowed, dtype=bool)
    n = len(allowed)
    padded = np.zeros((n + 31) // 32 * 32, dtype=np.uint8)
    padded[:n] = allowed
    return np.packbits(padded, bitorder="little").view(np.uint32).copy()


def merge_topk(scores: np.ndarray, ids: np.ndarray, k: int):
    """K4: [G, Q, k] per-shard lists (ids global, -1 = empty) -> global (score desc, id asc) top-k."""
    G, Q, kk = scores.shape
    out_s = np.full((Q, k), -np.inf, dtype=np.float64)
    out_i = np.full((Q, k), -1, dtype=np.int64)
    for qi in range(Q):
        s = scores[:, qi, :].reshape(-1)
        i = ids[:, qi, :].reshape(-1)
        keep = i >= 0
        s, i = s[keep], i[keep]
        order = np.lexsort((i, -s))[:k]
        out_s[qi, : len(order)] = s[order]
        out_i[qi, : len(order)] = i[order]
    return out_s, out_i


# ---------------------------------------------------------------------------------------------
# synthetic inputs shared by tests, smoke() and bench.py's CPU leg
# ---------------------------------------------------------------------------------------------

def synthetic_unit_rows(n: int, dim: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, dim), dtype=np.float32)
    return x


def synthetic_clustered(n: int, dim: int, n_classes: int, seed: int, noise: float = 0.40):
    """Rows near class centres (cos(row, centre) ~ 0.93), like the reference's CLIP scores."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_classes, dim)).astype(np.float32)
    centres /= np.linalg.norm(centres, axis=1, keepdims=True)
    labels = rng.permutation(n) % n_classes
    x = centres[labels] + (noise / np.sqrt(dim)) * rng.standard_normal((n, dim)).astype(np.float32)
    return x.astype(np.float32), labels.astype(np.int64), centres
