"""Seeded sweep over small shapes, every storage dtype and every distance of the collection menu
(util/qdrant_manager.py:61-66), with and without a row mask: ids and scores through the C ABI against the float64
oracle.  Catches edge cases of the planners (padding columns, one-row galleries, k beyond the gallery, tiny dims)."""
import itertools

import numpy as np
import pytest

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


def _want(metric, q, stored, k, mask):
    if metric == "cosine":
        s, i = O.cosine_topk(q, stored, k, row_mask=mask, rowwise=True)
        return s, i, s
    if metric == "dot":
        sc = q.astype(np.float64) @ stored.astype(np.float64).T
        s, i = O.topk_from_scores(sc, k, row_mask=mask)
        return s, i, s
    d, i, keys = O.distance_topk(q, stored, k, metric, row_mask=mask)
    return d, i, keys


CASES = []
_rng = np.random.default_rng(2024)
for metric, dtype in itertools.product(("cosine", "dot", "euclid", "manhattan"), ("f32", "bf16", "f16")):
    for _ in range(6):
        CASES.append((metric, dtype, int(_rng.choice([1, 2, 37, 128, 129, 1000, 4500])),
                      int(_rng.choice([3, 31, 64, 100, 320, 512, 768])), int(_rng.choice([1, 2, 9, 130])),
                      int(_rng.choice([1, 5, 33])), bool(_rng.integers(0, 2)), int(_rng.integers(0, 1 << 30))))


# vectors wider than 768 columns stream the query tile through shared memory (K3 variant 1; MANHATTAN still takes the
# fp64 sweep), and k above the tensor-core candidate lists takes the exact fp64 sweep (K5) for every distance
for metric in ("cosine", "dot", "euclid", "manhattan"):
    CASES.append((metric, "f32", 3000, 1024, 9, 5, False, 77))
    CASES.append((metric, "bf16", 2500, 1280, 40, 10, True, 78))
    CASES.append((metric, "f16", 6000, 256, 3, 300, False, 79))
    CASES.append((metric, "f32", 9000, 768, 33, 129, True, 80))


@pytest.mark.parametrize("metric,dtype,n,dim,Q,k,masked,seed", CASES)
def test_small_shapes_all_distances(metric, dtype, n, dim, Q, k, masked, seed):
    from retrieval_based_object_detection_b200 import Gallery

    rng = np.random.default_rng(seed)
    scale = 1.0 if metric == "cosine" else 0.5
    x = (rng.standard_normal((n, dim)) * scale * rng.uniform(0.3, 1.5, (n, 1))).astype(np.float32)
    if n > 20:
        x[11] = x[3]                                        # an exact duplicate: ties go to the smaller row
    g = Gallery(dim, dtype=dtype, metric=metric, capacity=max(1, n // 3))
    g.upsert(x[: n // 2])
    g.upsert(x[n // 2:])
    stored = g.get_rows(np.arange(n))
    q = (rng.standard_normal((Q, dim)) * scale).astype(np.float32)
    q[0] = stored[min(3, n - 1)]
    mask = None
    if masked:
        mask = rng.random(n) < 0.6
        mask[min(3, n - 1)] = True
    res = g.search(q, k, row_mask=None if mask is None else O.pack_row_mask(mask), want_scores64=True)
    ws, wi, wk = _want(metric, q, stored, k, mask)
    assert np.array_equal(res.rows, wi), (metric, dtype, n, dim, Q, k, masked, int((res.rows != wi).any(axis=1).sum()))
    fin = wi >= 0
    assert np.allclose(res.scores64[fin], wk[fin], rtol=1e-9, atol=1e-12)
    assert np.allclose(res.scores[fin], ws[fin].astype(np.float32), rtol=2e-6, atol=1e-6)
    if metric in ("euclid", "manhattan"):
        assert np.all(np.isinf(res.scores[~fin])) and np.all(res.scores[~fin] > 0)
    else:
        assert np.all(np.isneginf(res.scores[~fin]))
    g.close()
