"""Split search (rbod_search_begin / rbod_global_cut / rbod_search_end / rbod_merge_topk_certified): the row-sharded
flow in which the shards exchange the global k-th best APPROXIMATE score before the exact rescoring, so a shard rescoring
only what can still reach the global answer.  Several shards are emulated on one GPU (one Gallery per shard, torch.stack
in place of the all-gathers); the NCCL version of the same flow is ShardedGallery._search_split, checked on 2 GPUs by
tools/check_sharded_nccl.py and by the bench line's parity check at N > 1.  Oracle: the float64 brute force."""
import numpy as np
import pytest

from oracle import oracle_np as O

pytestmark = pytest.mark.gpu


def _split_search(shards, offsets, q, k, m, fallback=True):
    """-> (scores32, ids, scores64, n_flagged) as numpy; the flow of ShardedGallery._search_split on one device."""
    import torch

    from retrieval_based_object_detection_b200 import merge_topk_packed
    from retrieval_based_object_detection_b200.gallery import global_cut, merge_topk_certified

    dev = torch.device("cuda", 0)
    qd = torch.as_tensor(q, device=dev).float().contiguous()
    Q = qd.shape[0]
    approx = [torch.empty((Q, m + 1), dtype=torch.float32, device=dev) for _ in shards]
    # the handles keep per-search state, so every begin precedes every end, as on separate GPUs
    for g, a in zip(shards, approx):
        g.search_begin(qd, k, m, a)
    cut = global_cut(torch.stack(approx, 0), k)
    packed = [torch.empty((2 * Q * k + Q,), dtype=torch.int64, device=dev) for _ in shards]
    for g, p in zip(shards, packed):
        g.search_end(cut, k, p)
    s32, ids, s64, flag_q, n_flag = merge_topk_certified(torch.stack(packed, 0), offsets, Q, k)
    n = int(n_flag.item())
    if n and fallback:
        idx = torch.sort(flag_q[:n].long()).values
        loc = []
        for g in shards:
            buf = torch.empty((2, n, k), dtype=torch.int64, device=dev)
            g.search(qd[idx].contiguous(), k, out=(torch.empty((n, k), device=dev), buf[1], buf[0].view(torch.float64)))
            loc.append(buf)
        f32, fids, f64 = merge_topk_packed(torch.stack(loc, 0), offsets, k)
        s32[idx], ids[idx], s64[idx] = f32, fids, f64
    rescored = sum(int((p[Q * k:2 * Q * k] >= 0).sum().item()) for p in packed)
    return s32.cpu().numpy(), ids.cpu().numpy(), s64.cpu().numpy(), n, rescored


def _shards(n, dim, dtype, metric, cuts, seed, clustered=False):
    from retrieval_based_object_detection_b200 import Gallery

    if clustered:
        x, _, _ = O.synthetic_clustered(n, dim, max(2, n // 200), seed=seed)
    else:
        x = O.synthetic_unit_rows(n, dim, seed=seed)
    if metric != "cosine":
        x = x * np.linspace(0.5, 2.0, n, dtype=np.float32)[:, None]
    bounds = [0] + list(cuts) + [n]
    shards, stored = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        g = Gallery(dim, dtype=dtype, metric=metric, capacity=b - a)
        g.upsert(x[a:b])
        stored.append(g.get_rows(np.arange(b - a)))
        shards.append(g)
    return shards, bounds[:-1], np.concatenate(stored, 0), x


@pytest.mark.parametrize("dtype,metric,k,m", [("bf16", "cosine", 100, 42), ("f16", "cosine", 10, 10), ("f32", "dot", 33, 20),
                                              ("f16", "euclid", 50, 25), ("bf16", "cosine", 128, 128)])
def test_split_search_equals_the_brute_force(dtype, metric, k, m):
    n, dim, Q = 90_000, 256, 70
    shards, offs, stored, x = _shards(n, dim, dtype, metric, (20_000, 65_000), seed=k)
    q = O.synthetic_unit_rows(Q, dim, seed=3)
    q[:30] = x[100:130] + 0.3 * O.synthetic_unit_rows(30, dim, seed=4)       # queries with real neighbours
    s32, ids, s64, flagged, rescored = _split_search(shards, offs, q, k, m)
    if metric == "cosine":
        ws, wi = O.cosine_topk(q, stored, k)
    elif metric == "dot":
        ws, wi = O.topk_from_scores(q.astype(np.float64) @ stored.astype(np.float64).T, k)
    else:
        _, wi, ws = O.distance_topk(q, stored, k, "euclid")          # keys = -squared distance, what the lists carry
    assert np.array_equal(ids, wi), (flagged, np.argwhere(ids != wi)[:5])
    np.testing.assert_allclose(s64, ws, rtol=1e-9, atol=1e-9)
    # the point of the exchange: far fewer exact scores than 3 shards x Q x k
    assert rescored < 0.75 * 3 * Q * k, (rescored, 3 * Q * k)
    for g in shards:
        g.close()


def test_split_search_flags_what_it_cannot_certify_and_the_fallback_answers_it():
    """bf16 operand without its fp16 shadow, k = 100 on tightly clustered rows: the approximate scores cannot separate
    the k-th from the (k + slack)-th row, the merge flags those queries, and the plain search answers them."""
    n, dim, Q, k = 60_000, 128, 40, 100
    shards, offs, stored, x = _shards(n, dim, "bf16", "cosine", (25_000,), seed=11, clustered=True)
    for g in shards:
        g.set_option("auto_shadow", 0)
    q = x[::1500][:Q] + 0.01 * O.synthetic_unit_rows(Q, dim, seed=12)
    _, ids0, _, flagged, _ = _split_search(shards, offs, q, k, 60, fallback=False)
    s32, ids, s64, flagged2, _ = _split_search(shards, offs, q, k, 60)
    ws, wi = O.cosine_topk(q, stored, k)
    assert flagged == flagged2
    assert np.array_equal(ids, wi)
    if flagged == 0:
        assert np.array_equal(ids0, wi)
    else:
        # unflagged queries were already exact before the fallback
        bad = np.flatnonzero((ids0 != wi).any(axis=1))
        assert len(bad) <= flagged
    for g in shards:
        g.close()


def test_split_search_small_shards_masks_and_errors():
    from retrieval_based_object_detection_b200 import Gallery
    import torch

    # shards smaller than k: lists are padded, the cut is -inf, everything is rescored
    n, dim, Q, k = 150, 64, 9, 64
    shards, offs, stored, x = _shards(n, dim, "f32", "cosine", (40, 41), seed=2)
    q = O.synthetic_unit_rows(Q, dim, seed=8)
    s32, ids, s64, flagged, _ = _split_search(shards, offs, q, k, 64)
    ws, wi = O.cosine_topk(q, stored, k)
    assert np.array_equal(ids, wi)
    g = shards[0]
    dev = torch.device("cuda", 0)
    with pytest.raises(RuntimeError, match="no matching rbod_search_begin"):
        g.search_end(torch.zeros((Q, 2), device=dev), k, torch.empty((2 * Q * k + Q,), dtype=torch.int64, device=dev))
    a = torch.empty((Q, 11), dtype=torch.float32, device=dev)
    g.search_begin(torch.as_tensor(q, device=dev), 10, 10, a)
    g.search(q, 5)                                           # a full search in between invalidates the pending half
    with pytest.raises(RuntimeError, match="no matching rbod_search_begin"):
        g.search_end(torch.zeros((Q, 2), device=dev), 10, torch.empty((2 * Q * 10 + Q,), dtype=torch.int64, device=dev))
    for s in shards:
        s.close()
    gm = Gallery(dim, dtype="f32", metric="manhattan", capacity=10)
    gm.upsert(x[:10])
    with pytest.raises(RuntimeError, match="exact sweep"):
        gm.search_begin(torch.as_tensor(q, device=dev), 5, 5, torch.empty((Q, 6), dtype=torch.float32, device=dev))
    gm.close()


def test_split_search_with_an_empty_shard():
    """A rank that holds no rows yet still takes part in both exchanges: -inf scores, empty lists, no bound."""
    from retrieval_based_object_detection_b200 import Gallery

    n, dim, Q, k = 500, 64, 11, 10
    x = O.synthetic_unit_rows(n, dim, seed=21)
    full = Gallery(dim, dtype="f16", capacity=n)
    full.upsert(x)
    empty = Gallery(dim, dtype="f16", capacity=16)
    stored = full.get_rows(np.arange(n))
    q = O.synthetic_unit_rows(Q, dim, seed=22)
    for shards, offs in (([empty, full], [0, 0]), ([full, empty], [0, n])):
        s32, ids, s64, flagged, _ = _split_search(shards, offs, q, k, 10)
        ws, wi = O.cosine_topk(q, stored, k)
        assert np.array_equal(ids, wi) and flagged == 0
        np.testing.assert_allclose(s64, ws, rtol=1e-9, atol=1e-12)
    full.close()
    empty.close()


def test_sharded_gallery_split_path_on_one_rank():
    """ShardedGallery's own orchestration of the split form and of the split query upload, on a one-rank NCCL group
    (every line of the multi-GPU path except the peers; tools/check_sharded_nccl.py runs it on 2 GPUs)."""
    import socket

    import torch
    import torch.distributed as dist

    from retrieval_based_object_detection_b200 import ShardedGallery

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        n, dim, Q, k = 30_000, 128, 50, 40
        x = O.synthetic_unit_rows(n, dim, seed=5)
        sg = ShardedGallery(dim, n, dtype="f16", device=0)
        sg.upsert_local(x)
        sg.local.set_option("time_k3", 1)
        stored = sg.local.get_rows(np.arange(n))
        q = O.synthetic_unit_rows(Q, dim, seed=6)
        ws, wi = O.cosine_topk(q, stored, k)
        qd = sg._upload_split(q)                                  # host batch -> device through the gather
        assert torch.equal(qd.cpu(), torch.from_numpy(q))
        s32, ids, s64 = sg._search_split(qd, k)
        assert np.array_equal(ids.cpu().numpy(), wi)
        np.testing.assert_allclose(s64.cpu().numpy(), ws, rtol=1e-9, atol=1e-12)
        assert sg.last_split == {"flagged": 0, "approx_m": k} and sg.last_stats["k3_ms"] > 0
        host = (torch.empty(Q, k).pin_memory(), torch.empty(Q, k, dtype=torch.int64).pin_memory(),
                torch.empty(Q, k, dtype=torch.float64).pin_memory())
        h32, hids, h64 = sg.search(q, k, out_host=host)           # world 1: the plain path, delivered to host buffers
        assert np.array_equal(hids.numpy(), wi)
        sg.local.close()
    finally:
        dist.destroy_process_group()


def test_global_cut_and_certified_merge_against_numpy():
    """The two shard-independent kernels of the split form on synthetic buffers: the k-th largest of G*m gathered
    scores (short lists padded with -inf, duplicates, negative values) with the largest bound, and K4's certification
    flags (k-th merged score against the largest per-shard bound; lists that dropped nothing never flag)."""
    import torch

    from retrieval_based_object_detection_b200.gallery import global_cut, merge_topk_certified

    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(3)
    G, Q, m, k = 5, 300, 37, 50
    a = rng.standard_normal((G, Q, m + 1)).astype(np.float32)
    a[:, :, :m] = -np.sort(-a[:, :, :m], axis=2)
    a[1, :, 20:m] = -np.inf                                   # a shard with a short list
    a[:, 7, :m] = 0.25                                        # every score equal
    a[:, :, m] = (rng.random((G, Q)) * 1e-3).astype(np.float32)
    cut = global_cut(torch.as_tensor(a, device=dev), k).cpu().numpy()
    vals = np.sort(a[:, :, :m].transpose(1, 0, 2).reshape(Q, -1), axis=1)[:, ::-1]
    assert np.array_equal(cut[:, 0], vals[:, k - 1])
    assert np.array_equal(cut[:, 1], a[:, :, m].max(0))

    G, Q, k = 3, 40, 6
    s = -np.sort(-rng.standard_normal((G, Q, k)), axis=2)
    ids = rng.permutation(10 * G * Q * k)[: G * Q * k].reshape(G, Q, k).astype(np.int64)
    s[0, 5, 3:], ids[0, 5, 3:] = -np.inf, -1                  # a short list
    s[:, 9, 1:], ids[:, 9, 1:] = -np.inf, -1                  # fewer than k results in total
    ub = rng.standard_normal((G, Q)) - 0.5
    ub[:, :8] = -np.inf                                       # nothing was dropped anywhere: exact by construction
    offs = [0, 1_000_000, 2_000_000]
    packed = np.stack([np.concatenate([s[g].reshape(-1).view(np.int64), ids[g].reshape(-1), ub[g].view(np.int64)])
                       for g in range(G)])
    s32, mi, ms, flag_q, n_flag = merge_topk_certified(torch.as_tensor(packed, device=dev), offs, Q, k)
    gid = np.where(ids >= 0, ids + np.asarray(offs)[:, None, None], -1)
    ws, wi = O.merge_topk(s, gid, k)
    assert np.array_equal(mi.cpu().numpy(), wi) and np.array_equal(ms.cpu().numpy(), ws)
    kth = np.where(wi[:, k - 1] >= 0, ws[:, k - 1], -np.inf)
    ubm = ub.max(0)
    want = set(np.flatnonzero(np.isfinite(ubm) & ~(ubm < kth)).tolist())
    n = int(n_flag.item())
    assert set(flag_q[:n].cpu().tolist()) == want and n == len(want) and 9 in want and not (want & set(range(8)))
