"""Row-sharded multi-GPU search: one process per GPU, gallery rows block-partitioned over ranks.

Plain form (k < 32): each rank searches its own rows (K3 + exact rescoring -> local top-k with float64 scores) straight
into one packed [2, Q, k] buffer (scores, then local row slots), the buffers are exchanged with ONE all-gather over
NCCL/NVLink, and every rank merges the G lists with K4 (``rbod_merge_topk_packed``, which also turns local slots into
global ids).  Split form (k >= 32, ``_search_split``): the shards first exchange their best APPROXIMATE scores, derive
the global k-th best, and rescore only what can still reach the global answer; a second all-gather carries the exact
lists and a per-query bound, K4 merges and certifies.  Scores travel as float64 so the merged order is exactly the
(score desc, id asc) order a single GPU would produce.  A query batch that arrives in host memory is uploaded 1/G per
rank and completed by an all-gather (``_upload_split``).

The reference has no multi-GPU path at all; this implements BASELINE.json's
"shard the gallery by rows ... merge the global top-k with an NCCL allgather".
"""
from __future__ import annotations

import sys
from typing import Callable, Optional, Tuple


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block partition: rank r owns global rows [start, end).  Sizes differ by <= 1."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"shard_range: rank {rank} of world {world}")
    base, rem = divmod(int(n_rows), world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_stack(t, group=None):
    """[...]-tensor per rank -> [G, ...] on every rank (NCCL: one all_gather_into_tensor)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = t.contiguous()
    if t.is_cuda:
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t, group=group)
        return out
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return torch.stack(parts, 0)


class ShardedGallery:
    """A gallery whose rows are split over the ranks of a torch.distributed process group.

    ``local_search`` / ``merge`` are injectable so the host-side logic (partitioning, id offsets,
    gather layout) is testable on CPU with the gloo backend; the defaults call librbod.so.
    """

    def __init__(self, dim: int, n_rows_total: int, dtype: str = "bf16", metric: str = "cosine", group=None,
                 device: Optional[int] = None, local_search: Optional[Callable] = None,
                 merge: Optional[Callable] = None, create_local: bool = True,
                 local_sums: Optional[Callable] = None, finish: Optional[Callable] = None, split_ops=None):
        import torch.distributed as dist

        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.dim = dim
        self.n_rows_total = int(n_rows_total)
        self.row_start, self.row_end = shard_range(self.n_rows_total, self.rank, self.world)
        self._local_search = local_search
        self._merge = merge
        self._local_sums = local_sums
        self._finish = finish
        self._split_ops = split_ops    # test double for the split search's arithmetic (see _LibSplitOps)
        self.metric = metric
        self.local = None
        self._buf = None       # packed [2, Q, k] result buffer + fp32 scores, reused across searches of one shape
        self._qbuf = None      # device staging of the query batch when it arrives in host memory
        self._sbuf = None      # buffers of the split search, reused across searches of one shape
        self.split_min_k = 32  # from this k on a multi-GPU search exchanges a global cut before rescoring (_search_split)
        self.last_split = None # {"flagged": n} of the last split search
        self.last_stats = None
        if create_local:
            from .gallery import Gallery

            self.local = Gallery(dim, dtype=dtype, metric=metric, capacity=self.row_end - self.row_start,
                                 device=self.rank if device is None else device)

    @classmethod
    def wrap(cls, local_gallery, n_rows_total: int, metric: str = "cosine", group=None):
        """A sharded view over an existing per-rank ``Gallery`` that already holds rows shard_range(n, rank, G)."""
        sg = cls(local_gallery.dim, n_rows_total, metric=metric, group=group, create_local=False)
        sg.local = local_gallery
        return sg

    @property
    def local_rows(self) -> int:
        return self.row_end - self.row_start

    def upsert_local(self, rows):
        """Appends this rank's rows (global ids row_start + local slot)."""
        return self.local.upsert(rows)

    def shard_offsets(self):
        """Global row offset of every rank's shard (what K4 adds to the local row slots)."""
        return [shard_range(self.n_rows_total, r, self.world)[0] for r in range(self.world)]

    def _upload_split(self, queries):
        """Host query batch -> device, each rank paying for 1/G of the PCIe traffic.

        Every rank holds the same [Q, dim] float32 batch in host memory (the caller's contract for a sharded search);
        rank r copies rows [r*per, (r+1)*per) over its own PCIe link and one all-gather over NVLink completes the batch
        on every GPU, instead of G full copies competing for host memory bandwidth."""
        import numpy as np
        import torch
        import torch.distributed as dist

        q = torch.from_numpy(np.ascontiguousarray(queries, dtype=np.float32)) if isinstance(queries, np.ndarray) \
            else queries.to(torch.float32).contiguous()
        if q.ndim == 1:
            q = q.unsqueeze(0)
        Q, dim = int(q.shape[0]), int(q.shape[1])
        per = -(-Q // self.world)
        dev = torch.device("cuda", self.local.device)
        if self._qbuf is None or tuple(self._qbuf[0].shape) != (self.world * per, dim):
            self._qbuf = (torch.empty((self.world * per, dim), dtype=torch.float32, device=dev),
                          torch.zeros((per, dim), dtype=torch.float32, device=dev))
        full, mine = self._qbuf
        lo, hi = min(Q, self.rank * per), min(Q, (self.rank + 1) * per)
        if hi > lo:
            mine[: hi - lo].copy_(q[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(full, mine, group=self.group)
        return full[:Q]

    def _split_applies(self, k: int) -> bool:
        if self.world < 2 or k < self.split_min_k or k > 128:
            return False
        if self._split_ops is not None:
            return True
        return (self._local_search is None and self._merge is None and self.metric in ("cosine", "dot", "euclid")
                and self.dim <= 2048 and self.n_rows_total >= self.world)

    def _gather_into(self, out, t):
        """out [G, ...] <- every rank's t: one all_gather_into_tensor over NCCL, a list all-gather on CPU (gloo)."""
        import torch.distributed as dist

        if t.is_cuda:
            dist.all_gather_into_tensor(out, t, group=self.group)
        else:
            parts = [out[g] for g in range(self.world)]
            dist.all_gather(parts, t.contiguous(), group=self.group)
        return out

    def _search_split(self, queries, k: int):
        """Global top-k with ONE small exchange before the exact rescoring (include/rbod.h "Split search").

        A plain sharded search makes every rank rescore its own k best candidates per query in fp64 although the merge
        keeps k of the G*k.  Here each rank first publishes its best APPROXIMATE scores (all-gather of [Q, m+1] floats),
        every rank derives the global k-th best approximate score, and a rank rescoring only its candidates within two
        error bounds of that cut does 1/G of the work.  The second all-gather carries the exact lists plus, per query,
        a bound on what the shard never listed; K4 merges and certifies; the rare uncertified queries are answered by
        the plain path and patched in.

        The six arithmetic steps are the methods of ``_LibSplitOps`` (librbod.so) or of an injected object with the
        same methods (the gloo tests answer them with the oracle); everything else -- buffer shapes, the two
        all-gathers, the patching of uncertified queries -- is this function."""
        import torch

        ops = self._split_ops if self._split_ops is not None else _LibSplitOps(self.local)
        Q = int(queries.shape[0])
        G = self.world
        m = min(k, max(8, -(-2 * k // G) + 8))
        dev = queries.device
        words = 2 * Q * k + Q
        if self._sbuf is None or self._sbuf[0] != (Q, k, m, dev):
            self._sbuf = ((Q, k, m, dev), torch.empty((Q, m + 1), dtype=torch.float32, device=dev),
                          torch.empty((G, Q, m + 1), dtype=torch.float32, device=dev),
                          torch.empty((words,), dtype=torch.int64, device=dev),
                          torch.empty((G, words), dtype=torch.int64, device=dev))
        _, approx, g_approx, packed, g_packed = self._sbuf
        st0 = ops.begin(queries, k, m, approx)
        self._gather_into(g_approx, approx)
        cut = ops.global_cut(g_approx, k)
        st1 = ops.end(cut, k, packed)
        self._gather_into(g_packed, packed)
        s32, ids, s64, flag_q, n_flag = ops.merge_certified(g_packed, self.shard_offsets(), Q, k)
        n = int(n_flag.item())                           # the one synchronisation of the call
        self.last_stats = dict(st0, total_launches=st1["total_launches"] + 2)   # + global_cut, K4
        self.last_split = {"flagged": n, "approx_m": m}
        k3_ms = ops.k3_ms()
        if k3_ms is not None:
            self.last_stats["k3_ms"] = k3_ms
        if n > 0:
            # every rank sees the same gathered data, hence the same list: answer those queries the plain way
            idx = torch.sort(flag_q[:n].to(torch.int64)).values
            sub = queries[idx].contiguous()
            loc = torch.empty((2, n, k), dtype=torch.int64, device=dev)
            st2 = ops.plain(sub, k, loc)
            g_loc = torch.empty((G, 2, n, k), dtype=torch.int64, device=dev)
            self._gather_into(g_loc, loc)
            f32, fids, f64 = ops.merge_packed(g_loc, self.shard_offsets(), k)
            s32[idx], ids[idx], s64[idx] = f32, fids, f64
            self.last_stats["total_launches"] += st2["total_launches"] + 1
            self.last_stats["fallback_queries"] = n
        return s32, ids, s64

    def search(self, queries, k: int, out_host=None):
        """Global top-k on every rank: (scores f32 [Q,k], global ids i64 [Q,k], scores f64 [Q,k]).

        The local search writes its float64 scores and local row slots into ONE [2, Q, k] buffer of 8-byte words;
        a single all-gather moves it; K4 merges the gathered [G, 2, Q, k] buffer and adds each shard's row offset.
        ``queries`` in host memory must be the same batch on every rank (each rank uploads 1/G of it, see
        ``_upload_split``); ``out_host`` = three host tensors (pinned for an asynchronous copy) that receive the
        merged (scores f32, ids, scores f64) and are returned after one synchronisation."""
        import torch

        if (self._local_search is None and self._split_ops is None and self.world > 1
                and not getattr(queries, "is_cuda", False)):
            queries = self._upload_split(queries)
        if self._split_applies(k):
            if not hasattr(queries, "is_cuda"):
                queries = torch.as_tensor(queries)
            if not queries.is_cuda and self._split_ops is None:
                queries = queries.to(torch.device("cuda", self.local.device), torch.float32)
            if queries.ndim == 1:
                queries = queries.unsqueeze(0)
            out = self._search_split(queries.to(torch.float32).contiguous(), k)
            return self._deliver(out, out_host)
        if self._local_search is not None:
            s64, rows = self._local_search(queries, k)
            packed = torch.stack([s64.to(torch.float64).contiguous().view(torch.int64), rows.to(torch.int64)], 0)
        else:
            Q = int(queries.shape[0]) if getattr(queries, "ndim", 2) == 2 else 1
            dev = torch.device("cuda", self.local.device)
            if self._buf is None or tuple(self._buf[0].shape) != (2, Q, k):
                self._buf = (torch.empty((2, Q, k), dtype=torch.int64, device=dev),
                             torch.empty((Q, k), dtype=torch.float32, device=dev))
            packed, s32 = self._buf
            self.last_stats = dict(self.local.search(queries, k, out=(s32, packed[1], packed[0].view(torch.float64))).stats)
            self.last_stats["total_launches"] += 1       # K4 below
            self.last_split = None
        if self.world == 1:
            gathered = packed.unsqueeze(0)
        else:
            gathered = all_gather_stack(packed, self.group)
        if self._merge is not None:
            off = torch.tensor(self.shard_offsets(), dtype=torch.int64).view(-1, 1, 1)
            g_i = gathered[:, 1]
            g_i = torch.where(g_i >= 0, g_i + off.to(g_i.device), g_i)
            out = self._merge(gathered[:, 0].contiguous().view(torch.float64), g_i, k)
        else:
            from .gallery import merge_topk_packed

            out = merge_topk_packed(gathered, self.shard_offsets(), k)
        return self._deliver(out, out_host)

    def _deliver(self, out, out_host):
        import torch

        if self.metric in ("euclid", "manhattan"):
            # the lists travel and merge as ordering keys (-d^2 / -d, larger = closer); hand back distances
            s32, ids, keys = out
            dist = torch.sqrt(-keys) if self.metric == "euclid" else -keys
            dist = torch.where(ids >= 0, dist, torch.full_like(dist, float("inf")))
            out = (dist.to(torch.float32), ids, keys)
        if out_host is not None:
            for h, m in zip(out_host, out):
                h.copy_(m, non_blocking=True)
            if out[0].is_cuda:
                torch.cuda.synchronize(out[0].device)
            return tuple(out_host)
        return out

    def segment_mean(self, offsets, row_idx=None):
        """"average" delegates of classes whose rows are spread over the shards.

        Every rank passes the SAME class list: ``offsets`` [C+1] is the CSR over this rank's own rows
        (``row_idx`` = local slots, or None when the local rows are stored in class order); a class with no
        local rows is an empty range.  Local K2 launch -> float64 column sums [C, dim]; ONE all-reduce of the
        sums and one of the counts; the finish kernel divides, rounds to fp32 and renormalises, so every rank
        ends with the same [C, dim] float32 delegates a single GPU holding all the rows would produce."""
        import torch
        import torch.distributed as dist

        if self._local_sums is not None:
            sums = self._local_sums(offsets, row_idx)
        else:
            sums = self.local.segment_sums(offsets, row_idx)
        off = torch.as_tensor(offsets).to(torch.int64)
        counts = (off[1:] - off[:-1]).to(sums.device)
        if self.world > 1:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
        if self._finish is not None:
            return self._finish(sums, counts, self.metric == "cosine")
        from .gallery import segment_finish

        return segment_finish(sums, counts, normalize=self.metric == "cosine")


class _LibSplitOps:
    """The arithmetic steps of ShardedGallery._search_split, answered by librbod.so on this rank's Gallery."""

    def __init__(self, local):
        self.local = local

    def begin(self, queries, k, m, approx):
        return self.local.search_begin(queries, k, m, approx)

    def global_cut(self, g_approx, k):
        from .gallery import global_cut

        return global_cut(g_approx, k)

    def end(self, cut, k, packed):
        return self.local.search_end(cut, k, packed)

    def merge_certified(self, g_packed, offsets, Q, k):
        from .gallery import merge_topk_certified

        return merge_topk_certified(g_packed, offsets, Q, k)

    def k3_ms(self):
        return self.local.last_k3_ms() if self.local.options.get("time_k3") else None

    def plain(self, sub, k, loc):
        import torch

        l32 = torch.empty((sub.shape[0], k), dtype=torch.float32, device=sub.device)
        return self.local.search(sub, k, out=(l32, loc[1], loc[0].view(torch.float64))).stats

    def merge_packed(self, g_loc, offsets, k):
        from .gallery import merge_topk_packed

        return merge_topk_packed(g_loc, offsets, k)
