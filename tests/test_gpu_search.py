"""K3 (tcgen05 cosine top-k) + exact rescoring through rbod_search, against the fp64 brute force."""
import numpy as np
import pytest

from oracle import oracle_c as OC
from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    from retrieval_based_object_detection_b200 import Gallery

    return Gallery


def _mk(G, n, dim, dtype, seed, clustered=False):
    if clustered:
        x, _, _ = O.synthetic_clustered(n, dim, max(2, n // 100), seed=seed)
    else:
        x = O.synthetic_unit_rows(n, dim, seed=seed)
    g = G(dim, dtype=dtype, capacity=n)
    g.upsert(x)
    return g, g.get_rows(np.arange(n)), x


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("dtype,n,dim,Q", [("bf16", 1000, 768, 130), ("f16", 200, 512, 5), ("f32", 513, 512, 128),
                                           ("bf16", 300, 64, 7), ("bf16", 4097, 320, 260)])
def test_raw_tensor_core_scores(G, variant, dtype, n, dim, Q):
    """The tcgen05 pass itself: q16 . g16 with fp32 accumulation, every (query, row) cell."""
    g, stored, _ = _mk(G, n, dim, dtype, seed=n)
    g.set_option("k3_variant", variant)
    q = O.synthetic_unit_rows(Q, dim, seed=99)
    got = g.debug_scores(q)
    kind = "bf16" if dtype == "bf16" else "f16"        # fp32 galleries search an fp16 shadow
    qn = O.l2_normalize_store(q, kind)[0].astype(np.float64)
    g16 = O.round_store(stored, kind).astype(np.float64)
    want = qn @ g16.T
    assert got.shape == (Q, n) and not np.isnan(got).any()
    assert np.abs(got - want).max() < 2e-5, np.abs(got - want).max()
    g.close()


CASES = [
    # dtype, n, dim, Q, k, clustered
    ("f32", 1000, 512, 1000, 5, True),        # config C1 shape: 1k gallery, top-5
    ("bf16", 20000, 768, 300, 10, False),
    ("bf16", 20000, 768, 300, 100, False),
    ("f16", 5000, 768, 64, 10, True),
    ("f32", 30000, 512, 257, 10, True),
    ("bf16", 3, 512, 4, 10, False),           # fewer rows than k
    ("bf16", 64, 512, 1, 1, False),
    ("f32", 65, 100, 3, 40, False),           # dim not a multiple of 64 (padded operand)
    ("bf16", 150000, 768, 129, 10, False),    # many slices
]


@pytest.mark.parametrize("variant", [0, 2])
@pytest.mark.parametrize("dtype,n,dim,Q,k,clustered", CASES)
def test_search_matches_fp64_brute_force(G, variant, dtype, n, dim, Q, k, clustered):
    g, stored, x = _mk(G, n, dim, dtype, seed=n + k, clustered=clustered)
    g.set_option("k3_variant", variant)
    rng = np.random.default_rng(k)
    q = rng.standard_normal((Q, dim)).astype(np.float32)
    near = rng.integers(0, n, Q // 2)
    q[: Q // 2] = x[near] + 0.3 * q[: Q // 2] / np.sqrt(dim)       # queries close to stored rows
    res = g.search(q, k, want_scores64=True)
    ws, wi = O.cosine_topk(q, stored, k)
    assert np.array_equal(res.rows, wi), f"ids differ in {(res.rows != wi).any(axis=1).sum()} of {Q} queries"
    finite = np.isfinite(ws)
    assert np.allclose(res.scores64[finite], ws[finite], rtol=1e-5, atol=1e-9)      # tolerance of the north star
    assert np.array_equal(res.scores, res.scores64.astype(np.float32))
    assert np.all(np.isneginf(res.scores64[~finite]))
    assert res.stats["queries"] == Q and res.stats["k3_launches"] in (1, 2)
    g.close()


def test_duplicates_ties_and_fallback(G):
    """Exact duplicates tie; ties resolve to the smaller row slot; the certification fallback fires."""
    n, dim, k = 6000, 512, 10
    g, stored, x = _mk(G, n, dim, "bf16", seed=3)
    x2 = x.copy()
    x2[1000:1060] = x2[7]                    # 61 identical rows > kc: forces the exact fallback
    x2[2000] = x2[11]
    g.upsert(x2, slots=np.arange(n))
    stored = g.get_rows(np.arange(n))
    q = np.stack([x2[7] * 5.0, x2[11], x2[4000], -x2[7]]).astype(np.float32)
    res = g.search(q, k, want_scores64=True)
    ws, wi = OC.cosine_topk(q, stored, k)    # sequential C oracle: identical rows -> identical scores
    assert np.array_equal(res.rows, wi)
    assert list(res.rows[0]) == [7] + list(range(1000, 1009))
    assert list(res.rows[1][:2]) == [11, 2000]
    assert res.stats["fallback_queries"] >= 1
    assert np.allclose(res.scores64, ws, rtol=1e-5, atol=1e-9)
    g.close()


def test_wide_tie_cluster_goes_to_exact_sweep(G):
    """More identical rows than the collecting second pass records (1024): the fp64 sweep answers."""
    n, dim, k = 9000, 512, 10
    g, stored, x = _mk(G, n, dim, "bf16", seed=5)
    x2 = x.copy()
    x2[3000:4200] = x2[17]                   # 1201 copies of row 17
    g.upsert(x2, slots=np.arange(n))
    stored = g.get_rows(np.arange(n))
    q = np.stack([x2[17], x2[5000]]).astype(np.float32)
    res = g.search(q, k, want_scores64=True)
    ws, wi = OC.cosine_topk(q, stored, k)
    assert np.array_equal(res.rows, wi)
    assert list(res.rows[0]) == [17] + list(range(3000, 3009))
    assert res.stats["fallback_queries"] >= 1 and res.stats["sweep_queries"] >= 1
    assert res.stats["k3_launches"] == 2
    g.set_option("collect_pass", 0)          # the sweep alone gives the same answer
    res2 = g.search(q, k, want_scores64=True)
    assert np.array_equal(res2.rows, wi) and res2.stats["k3_launches"] == 1
    g.close()


@pytest.mark.parametrize("tau_share", [0, 1])
def test_k100_bf16_uncertified_queries_take_the_collect_pass(G, tau_share):
    """bf16 rounding leaves little slack at k=100 (kc=128): a good share of the queries is not certified by
    the first pass; the collecting pass must return exactly the brute-force ids for them."""
    n, dim, Q, k = 120000, 768, 200, 100
    g, stored, x = _mk(G, n, dim, "bf16", seed=21)
    g.set_option("tau_share", tau_share)
    q = O.synthetic_unit_rows(Q, dim, seed=77)
    res = g.search(q, k, want_scores64=True)
    ws, wi = O.cosine_topk(q, stored, k)
    assert np.array_equal(res.rows, wi), f"ids differ in {(res.rows != wi).any(axis=1).sum()} of {Q} queries"
    assert np.allclose(res.scores64, ws, rtol=1e-5, atol=1e-9)
    assert res.stats["sweep_queries"] == 0
    g.close()


@pytest.mark.parametrize("dtype,ordered", [("bf16", False), ("f32", True), ("f16", True)])
def test_threshold_prepass_keeps_the_answer_exact(G, dtype, ordered):
    """Galleries large enough for the sampled starting thresholds (>= 320 tiles), random and stored in class
    order (the adversarial layout for a sample): ids identical to the brute force, with and without it."""
    n, dim, Q, k = 60000, 256, 300, 10
    x, labels, _ = O.synthetic_clustered(n, dim, 40, seed=4)
    if ordered:
        x = x[np.argsort(labels, kind="stable")]
    g = G(dim, dtype=dtype, capacity=n)
    g.upsert(x)
    stored = g.get_rows(np.arange(n))
    rng = np.random.default_rng(1)
    q = x[rng.integers(0, n, Q)] + 0.2 * rng.standard_normal((Q, dim)).astype(np.float32) / np.sqrt(dim)
    mask = rng.random(n) < 0.5
    ws, wi = O.cosine_topk(q, stored, k)
    wms, wmi = O.cosine_topk(q, stored, k, row_mask=mask)
    for presample in (1, 0):
        g.set_option("presample", presample)
        res = g.search(q, k, want_scores64=True)
        assert np.array_equal(res.rows, wi), (presample, (res.rows != wi).any(axis=1).sum())
        assert np.allclose(res.scores64, ws, rtol=1e-5, atol=1e-9)
        assert res.stats["total_launches"] >= (7 if presample else 5)      # pre-pass = K3 sample launch + tau_init
        resm = g.search(q, k, row_mask=O.pack_row_mask(mask), want_scores64=True)
        assert np.array_equal(resm.rows, wmi)
    g.close()


def test_prepass_threshold_too_high_is_retried(G):
    """Every sampled tile holds a copy of the query: the starting threshold equals the best score, fewer than k
    rows beat it, and the call must be redone without the pre-pass -- same exact answer."""
    n, dim, k = 60000, 256, 10
    x = O.synthetic_unit_rows(n, dim, seed=12)
    tiles = (n + 127) // 128
    for g_ in range(4):                                   # K3_SAMPLE_GROUPS one-tile groups, spread over the gallery
        x[(g_ * tiles // 4) * 128 + 5] = x[77]
    g = G(dim, dtype="bf16", capacity=n)
    g.upsert(x)
    stored = g.get_rows(np.arange(n))
    q = np.stack([x[77], x[1234]]).astype(np.float32)
    res = g.search(q, k, want_scores64=True)
    ws, wi = OC.cosine_topk(q, stored, k)
    assert np.array_equal(res.rows, wi)
    assert res.stats["presample_retries"] >= 1
    g.close()


def test_row_mask_and_zero_vectors(G):
    n, dim, k, Q = 5000, 768, 10, 40
    g, stored, x = _mk(G, n, dim, "bf16", seed=8)
    rng = np.random.default_rng(0)
    allowed = rng.random(n) < 0.3
    q = rng.standard_normal((Q, dim)).astype(np.float32)
    q[0] = 0.0                                # zero query: every score is 0 -> ids 0..k-1 of the allowed rows
    res = g.search(q, k, row_mask=O.pack_row_mask(allowed), want_scores64=True)
    ws, wi = O.cosine_topk(q, stored, k, row_mask=allowed)
    assert np.array_equal(res.rows, wi)
    assert np.all(allowed[res.rows])
    none = g.search(q[:3], k, row_mask=O.pack_row_mask(np.zeros(n, bool)))
    assert np.all(none.rows == -1)
    g.close()


def test_empty_gallery_bad_args_and_device_io(G):
    import torch

    from retrieval_based_object_detection_b200._native import RbodError

    g = G(512, dtype="bf16")
    r = g.search(np.ones((2, 512), np.float32), 3)
    assert np.all(r.rows == -1) and np.all(np.isneginf(r.scores))
    x = O.synthetic_unit_rows(3000, 512, seed=1)
    g.upsert(x)
    with pytest.raises(RbodError):
        g.search(x[:2], 0)
    with pytest.raises(RbodError):
        g.search(x[:2], 500)                 # more candidates than the kernel keeps
    with pytest.raises(ValueError):
        g.search(np.ones((2, 100), np.float32), 3)
    stored = g.get_rows(np.arange(3000))
    qd = torch.from_numpy(x[:50]).cuda()
    r = g.search(qd, 10, want_scores64=True)  # device in -> device out
    assert r.rows.is_cuda and r.scores.is_cuda
    ws, wi = O.cosine_topk(x[:50], stored, 10)
    assert np.array_equal(r.rows.cpu().numpy(), wi)
    assert np.array_equal(r.rows.cpu().numpy()[:, 0], np.arange(50))       # each row finds itself first
    g.close()


def test_full_size_properties(G):
    """BASELINE-scale shapes, checked through size-independent properties (no brute force)."""
    import torch

    n, dim, Q, k = 2_000_000, 768, 512, 10
    g = G(dim, dtype="bf16", capacity=n)
    gen = torch.Generator("cuda").manual_seed(0)
    for a in range(0, n, 500_000):
        g.upsert(torch.randn(500_000, dim, device="cuda", generator=gen))
    assert len(g) == n
    probe = torch.randint(0, n, (Q,), device="cuda", generator=gen)
    q = g.get_rows(probe)                                  # stored rows as queries
    r = g.search(q, k, want_scores64=True)
    rows, s64 = r.rows.cpu().numpy(), r.scores64.cpu().numpy()
    assert np.array_equal(rows[:, 0], probe.cpu().numpy())               # self match is rank 0 ...
    assert np.all(np.abs(s64[:, 0] - 1.0) < 1e-12)                        # ... with cosine 1
    assert np.all(np.diff(s64, axis=1) <= 0)                              # sorted descending
    assert np.all((rows >= 0) & (rows < n)) and all(len(set(rw)) == k for rw in rows)
    # idempotence / linearity in the query norm: scaling a query changes nothing
    r2 = g.search(q * 7.5, k, want_scores64=True)
    assert np.array_equal(r2.rows.cpu().numpy(), rows)
    # the reported scores are the exact cosines of the reported rows
    chk = g.get_rows(torch.from_numpy(rows[:8].reshape(-1)).cuda()).double().reshape(8, k, dim)
    qq = q[:8].double()
    cos = torch.einsum("qd,qkd->qk", qq, chk) / (qq.norm(dim=1, keepdim=True) * chk.norm(dim=2))
    assert np.allclose(cos.cpu().numpy(), s64[:8], rtol=1e-9, atol=1e-12)
    # spot-check one query against the brute force over the whole gallery (torch fp64 on device)
    allrows = torch.cat([g.get_rows(torch.arange(a, a + 250_000, device="cuda")).double() @ qq[0]
                         / (qq[0].norm() * g.get_rows(torch.arange(a, a + 250_000, device="cuda")).double().norm(dim=1))
                         for a in range(0, n, 250_000)])
    top = torch.topk(allrows, k)
    assert np.array_equal(np.sort(top.indices.cpu().numpy()), np.sort(rows[0]))
    g.close()
