"""The oracle against the reference's own outputs (tests/golden, made by oracle/make_golden.py) and
against the known answers in the reference's committed run.  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import oracle_np as O
from oracle import ref_loader as R


@pytest.fixture(scope="module")
def delegates(golden_dir):
    return np.load(os.path.join(golden_dir, "delegates.npz"))


@pytest.mark.parametrize("case", range(5))
def test_delegate_restatements_match_reference_bitwise(delegates, case):
    stored = delegates[f"in_{case}"]
    v64 = stored.astype(np.float64)
    for name, fn in (("average", O.compute_average), ("centroid", O.compute_centroid),
                     ("weighted", O.compute_weighted_average), ("medoid", O.compute_medoid)):
        got = fn(v64)
        want = delegates[f"{name}_{case}"]
        assert got.dtype == np.float64 and got.shape == want.shape
        assert np.array_equal(got, want), f"{name} case {case}"


def test_cosine_restatement_matches_reference_bitwise(golden_dir):
    z = np.load(os.path.join(golden_dir, "cosine_pairs.npz"))
    got = np.array([O.cosine_similarity(z["a"][i], z["b"][i]) for i in range(len(z["cos"]))])
    assert np.array_equal(got, z["cos"])
    # C oracle: same formula, sequential summation -> within a few ulp of numpy's pairwise/BLAS dot
    got_c = np.array([OC.cosine_pair(z["a"][i], z["b"][i]) for i in range(len(z["cos"]))])
    assert np.allclose(got_c, z["cos"], rtol=0, atol=4e-16)


def test_reference_run_known_answers(golden_dir):
    kat = json.load(open(os.path.join(golden_dir, "reference_run_kat.json")))
    # (i) member-type delegates self-match at exactly 1.0000000000000002 in the reference's run
    assert kat["pre_a_centroid"]["max_hex"] == "0x1.0000000000001p+0"
    assert kat["centroid_equals_medoid_pre_a"] and kat["centroid_equals_medoid_pre_b"]
    assert kat["csv_header"] == "experiment_id,case,delegate_type,image_path,true_class,predicted_class,similarity_score"
    # ... which the oracle chain reproduces: stored fp32 unit vectors widened to f64 can self-match above 1
    x = O.synthetic_unit_rows(4000, 512, seed=3)
    stored, _ = O.l2_normalize_store(x, "f32")
    v = stored.astype(np.float64)
    self_cos = np.array([O.cosine_similarity(r, r) for r in v[:2000]])
    assert self_cos.max() <= 1.0000000000000002 and self_cos.min() >= 0.9999999999999998
    assert (self_cos == 1.0000000000000002).any()
    for key in ("pre_a_average", "pre_b_weighted"):
        assert kat[key]["n"] == 93 and kat[key]["dtype"] == "float64" and 0.85 < kat[key]["min"] <= kat[key]["max"] < 1.0


def test_golden_ids(golden_dir):
    import hashlib

    ids = json.load(open(os.path.join(golden_dir, "ids.json"), encoding="utf-8"))
    for e in ids["delegate"]:
        p, t = e["payload"], e["type"]
        key = f"{p.get('class_name')}::{t}::{p.get('data_type')}::{p.get('is_segmented')}::{p.get('is_augmented')}"
        assert hashlib.md5(key.encode()).hexdigest() == e["id"]
    from retrieval_based_object_detection_b200.store import canonical_id

    for e in ids["delegate"] + ids["path"]:
        c = canonical_id(e["id"])           # 32-hex md5 is a valid UUID; canonical form is hyphenated
        assert c.replace("-", "") == e["id"] and len(c) == 36


@pytest.mark.skipif(not R.available(), reason="reference tree only exists in the builder container")
def test_oracle_against_live_reference():
    ref = R.delegate_module()
    cos = R.cosine_similarity()
    x, _, _ = O.synthetic_clustered(33, 512, 1, seed=11)
    v = O.l2_normalize_store(x, "f32")[0].astype(np.float64)
    assert np.array_equal(ref.compute_average(v), O.compute_average(v))
    assert np.array_equal(ref.compute_medoid(v), O.compute_medoid(v))
    assert cos(v[0], v[1]) == O.cosine_similarity(v[0], v[1])


def test_bf16_f16_rounding():
    import torch

    rng = np.random.default_rng(0)
    x = np.concatenate([rng.standard_normal(20000).astype(np.float32) * 10 ** rng.uniform(-6, 3, 20000).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, 1.00390625, 1.01171875, 3.3895314e38, 65504.0, 6e-8], np.float32)])
    t = torch.from_numpy(x)
    assert np.array_equal(O.round_to_bf16(x), t.to(torch.bfloat16).to(torch.float32).numpy())
    assert np.array_equal(O.round_to_f16(x), t.to(torch.float16).to(torch.float32).numpy())


def test_normalize_c_vs_numpy():
    x = O.synthetic_unit_rows(300, 513, seed=5) * 7.0
    x[17] = 0.0
    a, na = O.l2_normalize_store(x, "f32")
    b, nb = OC.l2_normalize(x)
    ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max()
    assert ulp <= 1 and np.all(a[17] == 0) and np.allclose(na, nb, rtol=1e-6)
    n = np.linalg.norm(a.astype(np.float64), axis=1)
    assert np.all(np.abs(np.delete(n, 17) - 1.0) < 1e-6)


def test_topk_c_vs_numpy_with_ties_and_mask():
    rng = np.random.default_rng(9)
    g = O.l2_normalize_store(rng.standard_normal((700, 96)).astype(np.float32), "bf16")[0]
    g[100] = g[5]; g[650] = g[5]; g[651] = g[5]        # exact duplicates -> ties broken by row id
    q = rng.standard_normal((9, 96)).astype(np.float32)
    q[0] = g[5] * 2.0
    allowed = rng.random(700) < 0.8
    allowed[[5, 100, 650]] = True
    for mask in (None, allowed):
        s1, i1 = O.cosine_topk(q, g, 12, row_mask=mask, rowwise=True)
        s2, i2 = OC.cosine_topk(q, g, 12, row_allowed=mask)
        assert np.array_equal(i1, i2)
        assert np.allclose(s1, s2, rtol=0, atol=1e-14)
    s, i = O.cosine_topk(q, g, 12, rowwise=True)
    assert list(i[0][:4]) == [5, 100, 650, 651]
    # fewer rows than k pads with (-inf, -1)
    s, i = O.cosine_topk(q[:2], g[:5], 8)
    assert np.all(i[:, 5:] == -1) and np.all(np.isneginf(s[:, 5:]))
    w = O.pack_row_mask(allowed)
    assert np.array_equal(O.unpack_row_mask(w, 700), allowed)


def test_segment_mean_oracles_agree(delegates):
    stored = delegates["in_3"]                      # 93 x 768
    offsets = np.array([0, 10, 10, 60, 93], dtype=np.int64)
    perm = np.random.default_rng(1).permutation(93).astype(np.int64)
    a = O.segment_mean_renorm(stored, perm, offsets)
    b = OC.segment_mean(stored, perm, offsets)
    assert np.abs(a - b).max() <= 1.2e-7 and np.all(a[1] == 0)
    # class 0 of the identity layout == the reference's compute_average, stored (normalised) form
    full = O.segment_mean_renorm(stored, None, np.array([0, 93]))
    want = O.l2_normalize_store(delegates["average_3"].astype(np.float32)[None], "f32")[0][0]
    assert np.array_equal(full[0], want)


def test_merge_topk_oracle():
    rng = np.random.default_rng(2)
    G, Q, k = 4, 6, 5
    full = rng.standard_normal((Q, 40))
    full[:, 7] = full[:, 33]                       # a cross-shard tie
    per_s, per_i = [], []
    for gidx in range(G):
        s, i = O.topk_from_scores(full[:, gidx * 10:(gidx + 1) * 10], k, ids=np.arange(gidx * 10, gidx * 10 + 10))
        per_s.append(s); per_i.append(i)
    ms, mi = O.merge_topk(np.stack(per_s), np.stack(per_i), k)
    ws, wi = O.topk_from_scores(full, k)
    assert np.array_equal(mi, wi) and np.array_equal(ms, ws)


@pytest.mark.parametrize("metric", ["euclid", "manhattan"])
def test_numpy_and_c_restatements_of_the_distance_topk_agree(metric):
    """Two independent restatements (vectorised numpy, strictly sequential C) of the distance top-k give the same
    rows and keys, with duplicates (ties to the smaller row), a row mask and k beyond the allowed rows."""
    from oracle import oracle_c as OC

    rng = np.random.default_rng(12)
    n, dim, Q, k = 700, 48, 9, 12
    g = (rng.standard_normal((n, dim)) * 1.3).astype(np.float32)
    g[300:305] = g[17]
    q = rng.standard_normal((Q, dim)).astype(np.float32)
    q[0] = g[17]
    allowed = rng.random(n) < 0.5
    allowed[[17, 300, 301]] = True
    for mask in (None, allowed):
        d, rows, keys = O.distance_topk(q, g, k, metric, row_mask=mask)
        ck, crows = OC.distance_topk(q, g, k, metric, row_allowed=mask)
        assert np.array_equal(rows, crows)
        assert np.allclose(keys, ck, rtol=1e-12, atol=0)
        assert np.all(np.diff(d, axis=1) >= 0) and d[0, 0] == 0.0 and rows[0, 0] == 17
    few = np.zeros(n, dtype=bool)
    few[[5, 9]] = True
    d, rows, keys = O.distance_topk(q[:2], g, 4, metric, row_mask=few)
    ck, crows = OC.distance_topk(q[:2], g, 4, metric, row_allowed=few)
    assert np.array_equal(rows, crows) and np.all(rows[:, 2:] == -1) and np.all(np.isinf(d[:, 2:]))
