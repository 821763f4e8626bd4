"""Multi-GPU search path (row shards -> local top-k -> all-gather -> K4 merge), host logic on CPU:
two gloo ranks, the local search and the merge answered by the oracle (injected), results compared with
a single brute force over the whole gallery.  The NCCL/K4 version runs in bench.py --gpus N on B200s."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle_np as O
from retrieval_based_object_detection_b200.sharded import ShardedGallery, shard_range


def test_shard_range_partitions_exactly():
    for n, w in ((10, 3), (7, 8), (10_000_000, 8), (0, 2), (1, 1)):
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, dim, Q, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = O.synthetic_unit_rows(n, dim, seed=0)
        x[n // 2 + 3] = x[5]                                  # an exact tie across the two shards
        stored = O.l2_normalize_store(x, "bf16")[0]
        q = O.synthetic_unit_rows(Q, dim, seed=1)
        q[0] = x[5]

        def local_search(queries, kk):                        # what Gallery.search returns for this rank's rows
            a, b = shard_range(n, rank, world)
            s, i = O.cosine_topk(np.asarray(queries), stored[a:b], kk)
            return torch.from_numpy(s), torch.from_numpy(i)

        def merge(g_s, g_i, kk):                              # what rbod_merge_topk does
            s, i = O.merge_topk(g_s.numpy(), g_i.numpy(), kk)
            return torch.from_numpy(s.astype(np.float32)), torch.from_numpy(i), torch.from_numpy(s)

        sg = ShardedGallery(dim, n, dtype="bf16", local_search=local_search, merge=merge, create_local=False)
        assert (sg.row_start, sg.row_end) == shard_range(n, rank, world)
        s32, ids, s64 = sg.search(q, k)
        np.save(os.path.join(out_dir, f"ids_{rank}.npy"), ids.numpy())
        np.save(os.path.join(out_dir, f"s64_{rank}.npy"), s64.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k", [(301, 10), (7, 5)])
def test_two_rank_search_equals_single_brute_force(tmp_path, n, k):
    dim, Q, world = 64, 9, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, dim, Q, k, str(tmp_path)), nprocs=world, join=True)
    x = O.synthetic_unit_rows(n, dim, seed=0)
    x[n // 2 + 3] = x[5]
    stored = O.l2_normalize_store(x, "bf16")[0]
    q = O.synthetic_unit_rows(Q, dim, seed=1)
    q[0] = x[5]
    ws, wi = O.cosine_topk(q, stored, k)
    for r in range(world):
        ids, s64 = np.load(tmp_path / f"ids_{r}.npy"), np.load(tmp_path / f"s64_{r}.npy")
        assert np.array_equal(ids, wi)                        # every rank holds the same global answer
        fin = np.isfinite(ws)
        assert np.allclose(s64[fin], ws[fin], rtol=0, atol=1e-15)
    assert list(wi[0][:2]) == [5, n // 2 + 3]                 # the cross-shard tie resolves to the smaller global id


# ---------------------------------------------------------------------------------------------
# K2 over row shards: per-rank fp64 column sums -> all-reduce -> finish
# ---------------------------------------------------------------------------------------------
def _k2_inputs(n, dim, n_cls):
    x, labels, _ = O.synthetic_clustered(n, dim, n_cls, seed=3)
    labels = labels.copy()
    labels[labels == 2] = 1                                   # class 2 is empty on every rank
    stored = O.l2_normalize_store(x, "f32")[0]
    return stored, labels


def _worker_k2(rank, world, port, n, dim, n_cls, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stored, labels = _k2_inputs(n, dim, n_cls)
        a, b = shard_range(n, rank, world)
        loc_labels = labels[a:b]
        order = np.argsort(loc_labels, kind="stable")          # local slots in class order
        offsets = np.zeros(n_cls + 1, dtype=np.int64)
        np.cumsum(np.bincount(loc_labels, minlength=n_cls), out=offsets[1:])

        def local_sums(off, idx):                             # what Gallery.segment_sums returns for this shard
            s = np.zeros((n_cls, dim), dtype=np.float64)
            for c in range(n_cls):
                rows = np.asarray(idx[off[c]:off[c + 1]], dtype=np.int64)
                if len(rows):
                    s[c] = stored[a:b][rows].astype(np.float64).sum(axis=0)
            return torch.from_numpy(s)

        def finish(sums, counts, normalize):                  # what rbod_segment_finish does
            sums, counts = sums.numpy(), counts.numpy()
            out = np.zeros((n_cls, dim), dtype=np.float32)
            for c in range(n_cls):
                if counts[c] > 0:
                    m = (sums[c] / float(counts[c])).astype(np.float32)
                    out[c] = O.l2_normalize_store(m[None, :], "f32")[0][0] if normalize else m
            return torch.from_numpy(out)

        sg = ShardedGallery(dim, n, dtype="f32", create_local=False, local_sums=local_sums, finish=finish)
        cent = sg.segment_mean(offsets, order)
        np.save(os.path.join(out_dir, f"cent_{rank}.npy"), cent.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_delegate_means_equal_single_gallery(tmp_path):
    n, dim, n_cls, world = 403, 64, 7, 2
    port = _free_port()
    mp.spawn(_worker_k2, args=(world, port, n, dim, n_cls, str(tmp_path)), nprocs=world, join=True)
    stored, labels = _k2_inputs(n, dim, n_cls)
    order = np.argsort(labels, kind="stable")
    offsets = np.zeros(n_cls + 1, dtype=np.int64)
    np.cumsum(np.bincount(labels, minlength=n_cls), out=offsets[1:])
    want = O.segment_mean_renorm(stored, order, offsets)
    assert np.all(want[2] == 0)
    for r in range(world):
        got = np.load(tmp_path / f"cent_{r}.npy")
        ulp = np.abs(got.view(np.int32).astype(np.int64) - want.view(np.int32).astype(np.int64)).max()
        assert ulp <= 1, ulp                                  # fp64 sums added in a different order: <= 1 fp32 ulp


# ---------------------------------------------------------------------------------------------
# EUCLID collection over row shards: the lists travel and merge as ordering keys (-d^2), distances come back
# ---------------------------------------------------------------------------------------------
def _worker_euclid(rank, world, port, n, dim, Q, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = O.synthetic_unit_rows(n, dim, seed=5) * np.float32(1.7)
        q = O.synthetic_unit_rows(Q, dim, seed=6)
        a, b = shard_range(n, rank, world)

        def local_search(queries, kk):                        # what Gallery.search gives for an EUCLID shard
            _, rows, keys = O.distance_topk(np.asarray(queries), x[a:b], kk, "euclid")
            return torch.from_numpy(keys), torch.from_numpy(rows)

        def merge(g_s, g_i, kk):
            s, i = O.merge_topk(g_s.numpy(), g_i.numpy(), kk)
            return torch.from_numpy(s.astype(np.float32)), torch.from_numpy(i), torch.from_numpy(s)

        sg = ShardedGallery(dim, n, dtype="f32", metric="euclid", local_search=local_search, merge=merge,
                            create_local=False)
        d32, ids, keys = sg.search(q, k)
        np.save(os.path.join(out_dir, f"e_ids_{rank}.npy"), ids.numpy())
        np.save(os.path.join(out_dir, f"e_d_{rank}.npy"), d32.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_euclid_search_returns_distances(tmp_path):
    n, dim, Q, k, world = 205, 32, 6, 7, 2
    port = _free_port()
    mp.spawn(_worker_euclid, args=(world, port, n, dim, Q, k, str(tmp_path)), nprocs=world, join=True)
    x = O.synthetic_unit_rows(n, dim, seed=5) * np.float32(1.7)
    q = O.synthetic_unit_rows(Q, dim, seed=6)
    wd, wi, _ = O.distance_topk(q, x, k, "euclid")
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"e_ids_{r}.npy"), wi)
        d = np.load(tmp_path / f"e_d_{r}.npy")
        assert np.allclose(d, wd, rtol=1e-6) and np.all(np.diff(d, axis=1) >= 0)


# ---------------------------------------------------------------------------------------------
# The split form (k >= split_min_k): begin -> all-gather -> global cut -> end -> all-gather -> merge + certification ->
# plain path for what was flagged.  The orchestration is ShardedGallery._search_split itself; the six arithmetic steps
# librbod.so performs on a GPU are answered here by the oracle (numpy float64), with a deliberately noisy
# "approximate" score so that the cut, the bound and the certification all have work to do.
class _OracleSplitOps:
    def __init__(self, stored_local, eps, slack):
        self.g = stored_local.astype(np.float64)
        self.gn = np.sqrt((self.g * self.g).sum(1))
        self.eps, self.slack = float(eps), int(slack)
        self.rescored = 0

    def _exact(self, q):
        q = np.asarray(q, dtype=np.float64)
        return (q @ self.g.T) / (np.sqrt((q * q).sum(1))[:, None] * self.gn[None, :])

    def begin(self, queries, k, m, approx):
        q = queries.numpy()
        ex = self._exact(q)
        rng = np.random.default_rng(1234)                     # same noise on every call: |approx - exact| <= 0.9 eps
        self.approx = (ex + (rng.random(ex.shape) * 1.8 - 0.9) * self.eps).astype(np.float32)
        self.q = q
        n = ex.shape[1]
        kc = min(n, k + self.slack)
        order = np.argsort(-self.approx, axis=1, kind="stable")[:, :kc]
        self.cand = order
        self.tau = np.where(n > kc, np.take_along_axis(self.approx, order[:, -1:], 1)[:, 0], -np.inf)
        top = np.full((q.shape[0], m + 1), -np.inf, dtype=np.float32)
        mm = min(m, kc)
        top[:, :mm] = np.take_along_axis(self.approx, order[:, :mm], 1)
        top[:, m] = self.eps
        approx.copy_(torch.from_numpy(top))
        return {"total_launches": 3, "candidates": kc, "fallback_queries": 0}

    def global_cut(self, g_approx, k):
        a = g_approx.numpy()
        G, Q, m1 = a.shape
        vals = np.sort(a[:, :, : m1 - 1].transpose(1, 0, 2).reshape(Q, -1), axis=1)[:, ::-1]
        return torch.from_numpy(np.stack([vals[:, k - 1], a[:, :, m1 - 1].max(0)], 1).astype(np.float32))

    def end(self, cut, k, packed):
        c = cut.numpy()
        Q = c.shape[0]
        ex = self._exact(self.q)
        s64 = np.full((Q, k), -np.inf)
        rows = np.full((Q, k), -1, dtype=np.int64)
        for i in range(Q):
            keep = self.cand[i][self.approx[i, self.cand[i]] >= c[i, 0] - 2.0 * max(self.eps, c[i, 1])]
            self.rescored += len(keep)
            sc = ex[i, keep]
            o = np.lexsort((keep, -sc))[:k]
            s64[i, : len(o)], rows[i, : len(o)] = sc[o], keep[o]
        ub = np.where(np.isfinite(self.tau), self.tau.astype(np.float64) + self.eps, -np.inf)
        buf = np.concatenate([s64.reshape(-1).view(np.int64), rows.reshape(-1), ub.view(np.int64)])
        packed.copy_(torch.from_numpy(buf))
        return {"total_launches": 4}

    def merge_certified(self, g_packed, offsets, Q, k):
        w = g_packed.numpy()
        G = w.shape[0]
        s = w[:, : Q * k].copy().view(np.float64).reshape(G, Q, k)
        r = w[:, Q * k:2 * Q * k].reshape(G, Q, k)
        ub = w[:, 2 * Q * k:].copy().view(np.float64).reshape(G, Q).max(0)
        ids = np.where(r >= 0, r + np.asarray(offsets, dtype=np.int64)[:, None, None], -1)
        ms, mi = O.merge_topk(s, ids, k)
        kth = np.where(mi[:, k - 1] >= 0, ms[:, k - 1], -np.inf)
        flagged = np.flatnonzero(np.isfinite(ub) & ~(ub < kth)).astype(np.int32)
        fq = np.zeros(Q, dtype=np.int32)
        fq[: len(flagged)] = flagged[::-1]                    # unordered on the device: the host must sort
        return (torch.from_numpy(ms.astype(np.float32)), torch.from_numpy(mi), torch.from_numpy(ms),
                torch.from_numpy(fq), torch.tensor([len(flagged)], dtype=torch.int32))

    def k3_ms(self):
        return None

    def plain(self, sub, k, loc):
        ex = self._exact(sub.numpy())
        s, i = O.topk_from_scores(ex, k)
        loc[0].copy_(torch.from_numpy(s.view(np.int64)))
        loc[1].copy_(torch.from_numpy(i))
        return {"total_launches": 5}

    def merge_packed(self, g_loc, offsets, k):
        w = g_loc.numpy()
        s = w[:, 0].copy().view(np.float64)
        ids = np.where(w[:, 1] >= 0, w[:, 1] + np.asarray(offsets, dtype=np.int64)[:, None, None], -1)
        ms, mi = O.merge_topk(s, ids, k)
        return torch.from_numpy(ms.astype(np.float32)), torch.from_numpy(mi), torch.from_numpy(ms)


def _split_worker(rank, world, port, n, dim, Q, k, eps, slack, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        stored = O.l2_normalize_store(O.synthetic_unit_rows(n, dim, seed=0), "bf16")[0]
        q = O.synthetic_unit_rows(Q, dim, seed=1)
        a, b = shard_range(n, rank, world)
        ops = _OracleSplitOps(stored[a:b], eps, slack)
        sg = ShardedGallery(dim, n, dtype="bf16", create_local=False, split_ops=ops)
        sg.split_min_k = 4
        s32, ids, s64 = sg.search(torch.from_numpy(q), k)
        np.save(os.path.join(out_dir, f"ids_{rank}.npy"), ids.numpy())
        np.save(os.path.join(out_dir, f"s64_{rank}.npy"), s64.numpy())
        np.save(os.path.join(out_dir, f"info_{rank}.npy"),
                np.array([sg.last_split["flagged"], ops.rescored, sg.last_stats["total_launches"]]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,k,eps,slack", [(400, 10, 2e-3, 22), (400, 10, 6e-2, 6), (37, 12, 2e-3, 4)])
def test_two_rank_split_search_equals_single_brute_force(tmp_path, n, k, eps, slack):
    """eps = 2e-3: nearly everything is certified and each rank rescoring only what clears the global cut does about
    half the work; eps = 6e-2: the bound is useless, most queries are flagged and the plain path answers them;
    n = 37: shards smaller than the candidate lists (nothing is ever dropped, nothing can be flagged)."""
    dim, Q, world = 48, 13, 2
    mp.spawn(_split_worker, args=(world, _free_port(), n, dim, Q, k, eps, slack, str(tmp_path)), nprocs=world, join=True)
    stored = O.l2_normalize_store(O.synthetic_unit_rows(n, dim, seed=0), "bf16")[0]
    q = O.synthetic_unit_rows(Q, dim, seed=1)
    ws, wi = O.cosine_topk(q, stored, k)
    infos = [np.load(tmp_path / f"info_{r}.npy") for r in range(world)]
    for r in range(world):
        ids, s64 = np.load(tmp_path / f"ids_{r}.npy"), np.load(tmp_path / f"s64_{r}.npy")
        assert np.array_equal(ids, wi), (r, infos)
        assert np.allclose(s64, ws, rtol=0, atol=1e-14)
    assert infos[0][0] == infos[1][0]                         # every rank saw the same flag list
    if eps > 1e-2:
        assert infos[0][0] > 0
    if n == 37:
        assert infos[0][0] == 0
    if (n, eps) == (400, 2e-3):
        assert infos[0][1] + infos[1][1] < 0.8 * world * Q * (k + slack)   # the cut spared exact scores
