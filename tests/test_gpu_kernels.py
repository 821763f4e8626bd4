"""GPU parity tests proper: every kernel through the C ABI against the oracle on seeded inputs."""
import numpy as np
import pytest

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


def _ulp(a, b):
    return np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32).astype(np.int64)).max()


@pytest.fixture(scope="module")
def G():
    from retrieval_based_object_detection_b200 import Gallery

    return Gallery


# ------------------------------------------------------------------ K1
@pytest.mark.parametrize("dtype", ["f32", "bf16", "f16"])
@pytest.mark.parametrize("n,dim", [(1, 512), (1000, 512), (777, 768), (300, 100), (64, 1024), (50, 3)])
def test_k1_normalize_pack(G, dtype, n, dim):
    x = O.synthetic_unit_rows(n, dim, seed=n + dim) * np.float32(3.7)
    if n > 10:
        x[3] = 0.0                                    # zero vector stays zero
    g = G(dim, dtype=dtype, capacity=n)
    norms = g.upsert(x, return_norms=True)
    assert len(g) == n
    want, want_norms = O.l2_normalize_store(x, dtype)
    got = g.get_rows(np.arange(n))
    assert got.shape == (n, dim)
    assert _ulp(got, want) <= (1 if dtype == "f32" else 0) or np.mean(got != want) < 1e-4
    assert np.mean(got != want) < 1e-5
    assert np.allclose(norms, want_norms, rtol=1e-6, atol=0)
    g.close()


def test_k1_overwrite_append_and_errors(G):
    from retrieval_based_object_detection_b200._native import RbodError

    dim = 512
    g = G(dim, dtype="f32", capacity=4)              # forces growth
    a = O.synthetic_unit_rows(3000, dim, seed=1)
    g.upsert(a[:2000])
    g.upsert(a[2000:])                               # append after growth
    b = O.synthetic_unit_rows(3, dim, seed=2)
    g.upsert(b, slots=np.array([5, 2999, 3000]))     # overwrite two, append one
    assert len(g) == 3001
    ref = np.concatenate([a, b[2:3]])
    ref[5], ref[2999] = b[0], b[1]
    want, _ = O.l2_normalize_store(ref, "f32")
    assert np.mean(g.get_rows(np.arange(3001)) != want) < 1e-5
    with pytest.raises(RbodError):
        g.upsert(b, slots=np.array([0, 1, 4000]))    # would leave holes
    with pytest.raises(RbodError):
        g.get_rows(np.array([3001]))
    with pytest.raises(ValueError):
        g.upsert(np.zeros((2, 100), np.float32))
    raw = g.get_rows(np.array([7, 8]))
    g.upsert(raw * 1.0, slots=np.array([7, 8]), raw=True)     # RAW round trip is bit exact
    assert np.array_equal(g.get_rows(np.array([7, 8])), raw)
    g.truncate(10)
    assert len(g) == 10
    g.close()


def test_k1_device_pointers(G):
    import torch

    from retrieval_based_object_detection_b200 import l2norm_pack

    x = torch.randn(5000, 768, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    for dt in ("f32", "bf16", "f16"):
        y = l2norm_pack(x, dt)
        want, _ = O.l2_normalize_store(x.cpu().numpy(), dt)
        assert np.mean(y.float().cpu().numpy() != want) < 1e-5
    g = G(768, dtype="bf16", capacity=5000)
    g.upsert(x)
    rows = g.get_rows(torch.arange(5000, device="cuda"))
    assert rows.is_cuda and np.mean(rows.cpu().numpy() != O.l2_normalize_store(x.cpu().numpy(), "bf16")[0]) < 1e-5
    g.close()


# ------------------------------------------------------------------ K2
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n,dim,ncls", [(2000, 512, 32), (5000, 768, 100), (700, 100, 7)])
def test_k2_segment_mean(G, dtype, n, dim, ncls):
    x, labels, _ = O.synthetic_clustered(n, dim, ncls, seed=n)
    g = G(dim, dtype=dtype, capacity=n)
    g.upsert(x)
    stored = g.get_rows(np.arange(n))
    order = np.argsort(labels, kind="stable").astype(np.int64)
    offsets = np.zeros(ncls + 1, np.int64)
    np.cumsum(np.bincount(labels, minlength=ncls), out=offsets[1:])
    got = g.segment_mean(offsets, row_idx=order)
    want = O.segment_mean_renorm(stored, order, offsets)
    assert got.shape == (ncls, dim)
    assert _ulp(got, want) <= 2 and np.mean(got != want) < 1e-3
    g.close()


def test_k2_skewed_empty_and_identity_layout(G):
    n, dim = 9000, 768
    x, _, _ = O.synthetic_clustered(n, dim, 5, seed=4)
    g = G(dim, dtype="f32", capacity=n)
    g.upsert(x)
    stored = g.get_rows(np.arange(n))
    # one class of 5000 rows (multi-chunk path), an empty class, a single-row class, the rest
    offsets = np.array([0, 5000, 5000, 5001, 9000], np.int64)
    got = g.segment_mean(offsets)                       # rows already "label sorted": identity row_idx
    want = O.segment_mean_renorm(stored, None, offsets)
    assert _ulp(got, want) <= 2 and np.all(got[1] == 0)
    assert _ulp(got[2:3], stored[5000:5001]) <= 1       # mean of one unit row is that row (renormalised)
    import torch
    got_dev = g.segment_mean(torch.from_numpy(offsets).cuda())
    assert got_dev.is_cuda and np.array_equal(got_dev.cpu().numpy(), got)
    from retrieval_based_object_detection_b200._native import RbodError
    with pytest.raises(RbodError):
        g.segment_mean(np.array([0, 5], np.int64), row_idx=np.array([0, 1, 2, 3, 99999], np.int64))
    g.close()


def test_k2_matches_reference_golden(G, golden_dir):
    """The golden 'average' delegates came out of the reference's compute_average itself."""
    import os

    z = np.load(os.path.join(golden_dir, "delegates.npz"))
    for case in range(5):
        stored = z[f"in_{case}"]
        g = G(stored.shape[1], dtype="f32", capacity=len(stored))
        g.upsert(stored, raw=True)
        got = g.segment_mean(np.array([0, len(stored)], np.int64))[0]
        want = O.l2_normalize_store(z[f"average_{case}"].astype(np.float32)[None], "f32")[0][0]
        assert _ulp(got[None], want[None]) <= 2
        g.close()


@pytest.mark.parametrize("dtype", ["f32", "bf16"])
def test_k2_shard_sums_then_finish_equals_one_gallery(G, dtype):
    """The sharded delegate build on one device: two galleries hold the two halves of the rows, each gives its
    fp64 column sums (rbod_segment_sums), the sums and counts are added (what the NCCL all-reduce does) and
    rbod_segment_finish produces the delegates -- same vectors as K2 on a single gallery holding every row."""
    import torch
    from retrieval_based_object_detection_b200.gallery import segment_finish

    n, dim, ncls = 6001, 768, 37
    x, labels, _ = O.synthetic_clustered(n, dim, ncls, seed=11)
    labels = labels.copy()
    labels[labels == 4] = 3                                   # class 4 has no rows anywhere
    labels[:3000][labels[:3000] == 9] = 8                     # class 9 lives on the second shard only
    whole = G(dim, dtype=dtype, capacity=n)
    whole.upsert(x)
    stored = whole.get_rows(np.arange(n))
    order = np.argsort(labels, kind="stable").astype(np.int64)
    offsets = np.zeros(ncls + 1, np.int64)
    np.cumsum(np.bincount(labels, minlength=ncls), out=offsets[1:])
    want_k2 = whole.segment_mean(offsets, row_idx=order)
    want = O.segment_mean_renorm(stored, order, offsets)
    sums, counts = None, None
    for a, b in ((0, 3000), (3000, n)):
        g = G(dim, dtype=dtype, capacity=b - a)
        g.upsert(x[a:b])
        lo = np.argsort(labels[a:b], kind="stable").astype(np.int64)
        off = np.zeros(ncls + 1, np.int64)
        np.cumsum(np.bincount(labels[a:b], minlength=ncls), out=off[1:])
        s = g.segment_sums(off, row_idx=lo)
        assert s.is_cuda and s.dtype == torch.float64 and tuple(s.shape) == (ncls, dim)
        ref = np.stack([stored[a:b][lo[off[c]:off[c + 1]]].astype(np.float64).sum(axis=0) for c in range(ncls)])
        assert np.allclose(s.cpu().numpy(), ref, rtol=0, atol=1e-12)
        c = torch.from_numpy(off[1:] - off[:-1]).cuda()
        sums = s.clone() if sums is None else sums + s
        counts = c if counts is None else counts + c
        g.close()
    got = segment_finish(sums, counts, normalize=True).cpu().numpy()
    assert np.all(got[4] == 0) and _ulp(got, want) <= 2 and _ulp(got, want_k2) <= 1
    raw = segment_finish(sums, counts, normalize=False).cpu().numpy()
    assert np.allclose(raw[8], (sums[8].cpu().numpy() / float(counts[8])).astype(np.float32), rtol=0, atol=0)
    whole.close()


# ------------------------------------------------------------------ K2b: centroid / weighted / medoid
@pytest.mark.parametrize("dtype", ["f32", "bf16"])
@pytest.mark.parametrize("n,dim,ncls", [(1200, 512, 12), (900, 768, 5), (300, 100, 7)])
def test_k2b_delegate_types_match_the_reference_functions(G, dtype, n, dim, ncls):
    """compute_centroid / compute_weighted_average / compute_medoid (32:12-26) per class, in stored form."""
    x, labels, _ = O.synthetic_clustered(n, dim, ncls, seed=n + 1)
    g = G(dim, dtype=dtype, capacity=n)
    g.upsert(x)
    stored = g.get_rows(np.arange(n))
    order = np.argsort(labels, kind="stable").astype(np.int64)
    offsets = np.zeros(ncls + 1, np.int64)
    np.cumsum(np.bincount(labels, minlength=ncls), out=offsets[1:])
    for kind, fn in (("centroid", O.compute_centroid), ("weighted", O.compute_weighted_average),
                     ("medoid", O.compute_medoid), ("average", O.compute_average)):
        got, members = g.segment_delegates(kind, offsets, row_idx=order)
        want = O.segment_mean_renorm(stored, order, offsets, average_fn=fn)
        assert got.shape == (ncls, dim) and _ulp(got, want) <= 2, kind
        if kind in ("centroid", "medoid"):
            for c in range(ncls):
                rows = order[offsets[c]:offsets[c + 1]]
                v = stored[rows].astype(np.float64)
                assert members[c] in rows and np.array_equal(stored[members[c]].astype(np.float64), fn(v)), (kind, c)
        else:
            assert np.all(members == -1)
    g.close()


def test_k2b_golden_empty_duplicates_and_device_io(G, golden_dir):
    import os

    import torch

    z = np.load(os.path.join(golden_dir, "delegates.npz"))
    for case in range(5):                                    # outputs of the reference's own functions
        stored = z[f"in_{case}"]
        g = G(stored.shape[1], dtype="f32", capacity=len(stored))
        g.upsert(stored, raw=True)
        for kind in ("centroid", "weighted", "medoid"):
            got, _ = g.segment_delegates(kind, np.array([0, len(stored)], np.int64))
            want = O.l2_normalize_store(z[f"{kind}_{case}"].astype(np.float32)[None], "f32")[0][0]
            assert _ulp(got, want[None]) <= 2, (case, kind)
        g.close()
    n, dim = 400, 256
    x = O.synthetic_unit_rows(n, dim, seed=9)
    x[40] = x[10]
    x[70] = x[10]                                            # duplicates: argmin ties resolve to the first member
    x[10] = x[:100].mean(axis=0)                             # ... and make that member the one nearest the mean
    x[40] = x[10]
    x[70] = x[10]
    g = G(dim, dtype="f32", capacity=n)
    g.upsert(x)
    stored = g.get_rows(np.arange(n))
    offsets = np.array([0, 100, 100, 101, 400], np.int64)    # a class, an empty class, a single member, the rest
    for kind, fn in (("centroid", O.compute_centroid), ("medoid", O.compute_medoid), ("weighted", O.compute_weighted_average)):
        got, members = g.segment_delegates(kind, offsets)
        want = O.segment_mean_renorm(stored, None, offsets, average_fn=fn)
        assert _ulp(got, want) <= 2 and np.all(got[1] == 0)
        if kind != "weighted":
            assert list(members[:3]) == [10, -1, 100]
        d_got, d_mem = g.segment_delegates(kind, torch.from_numpy(offsets).cuda())
        assert d_got.is_cuda and np.array_equal(d_got.cpu().numpy(), got) and np.array_equal(d_mem.cpu().numpy(), members)
    from retrieval_based_object_detection_b200._native import RbodError
    with pytest.raises(RbodError):
        g.segment_delegates("medoid", np.array([0, 3], np.int64), row_idx=np.array([0, 1, 99999], np.int64))
    with pytest.raises(ValueError):
        g.segment_delegates("mode", offsets)
    g.close()


# ------------------------------------------------------------------ K4 merge
def test_k4_merge_topk(G):
    import torch

    from retrieval_based_object_detection_b200 import merge_topk

    rng = np.random.default_rng(5)
    for Gn, Q, k in [(2, 50, 10), (8, 333, 100), (1, 7, 5), (4, 20, 1)]:
        full = rng.standard_normal((Q, Gn * 150))
        full[:, 3] = full[:, 151 % full.shape[1]]       # cross-shard exact tie
        ss, ii = [], []
        for r in range(Gn):
            s, i = O.topk_from_scores(full[:, r * 150:(r + 1) * 150], k, ids=np.arange(r * 150, (r + 1) * 150))
            if r == Gn - 1:
                s[:, k // 2:], i[:, k // 2:] = -np.inf, -1     # a short shard
            ss.append(s); ii.append(i)
        ss, ii = np.stack(ss), np.stack(ii)
        ws, wi = O.merge_topk(ss, ii, k)
        s32, ids, s64 = merge_topk(torch.from_numpy(ss).cuda(), torch.from_numpy(ii).cuda(), k)
        assert np.array_equal(ids.cpu().numpy(), wi)
        assert np.array_equal(s64.cpu().numpy(), ws)
        assert np.array_equal(s32.cpu().numpy(), ws.astype(np.float32))
