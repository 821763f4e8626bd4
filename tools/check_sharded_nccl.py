#!/usr/bin/env python
"""Multi-GPU check of the two sharded paths over real NCCL (run under torchrun, one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded_nccl.py

  * ShardedGallery.search        -- row shards, all-gather, K4 merge: ids identical to a torch fp64 brute force over the
                                    whole gallery (COSINE bf16 and EUCLID fp32 collections)
  * ShardedGallery.segment_mean  -- per-shard K2 sums, all-reduce, finish: delegates within 1 fp32 ulp of K2 on one
                                    gallery holding every row
Rank 0 prints one JSON line; every rank exits non-zero on a mismatch.  The CPU oracle is not involved: the reference
answer is computed with torch in float64 on the device (test plumbing, not product code).
"""
from __future__ import annotations

import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from retrieval_based_object_detection_b200 import Gallery, ShardedGallery, shard_range

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    out = {"world": world}

    n, dim, Q, k, n_cls = 200_003, 768, 300, 10, 211
    gen = torch.Generator(dev).manual_seed(5)                   # same stream on every rank -> same global data
    x = torch.randn(n, dim, device=dev, generator=gen)
    q = torch.randn(Q, dim, device=dev, generator=gen)
    labels = torch.randint(0, n_cls, (n,), device=dev, generator=gen)
    labels[labels == 7] = 8                                     # an empty class
    a, b = shard_range(n, rank, world)

    # ---- search, COSINE bf16
    sg = ShardedGallery(dim, n, dtype="bf16", device=local)
    sg.upsert_local(x[a:b])
    s32, ids, s64 = sg.search(q, k)
    whole = Gallery(dim, dtype="bf16", capacity=n, device=local)
    whole.upsert(x)
    stored = whole.get_rows(torch.arange(n, device=dev)).double()
    sc = (q.double() @ stored.T) / (q.double().norm(dim=1, keepdim=True) * stored.norm(dim=1)[None, :])
    top = torch.topk(sc, k, dim=1)
    out["cosine_ids_equal"] = bool(torch.equal(top.indices, ids))
    out["cosine_max_score_err"] = float((top.values - s64).abs().max())
    # the same batch from HOST memory (Q = 300 is not a multiple of the world size): every rank uploads 1/G of it and an
    # all-gather completes it; results land in caller-owned host tensors
    host = (torch.empty(Q, k).pin_memory(), torch.empty(Q, k, dtype=torch.int64).pin_memory(),
            torch.empty(Q, k, dtype=torch.float64).pin_memory())
    h32, hids, h64 = sg.search(q.cpu().numpy(), k, out_host=host)
    out["host_queries_ids_equal"] = bool(torch.equal(hids, ids.cpu()) and torch.equal(h64, s64.cpu()))

    # top-100 takes the split form (global cut exchanged before the exact rescoring), device and host queries
    s32c, idsc, s64c = sg.search(q, 100)
    top100 = torch.topk(sc, 100, dim=1)
    out["split_k100_ids_equal"] = bool(torch.equal(top100.indices, idsc))
    out["split_k100_max_score_err"] = float((top100.values - s64c).abs().max())
    out["split_k100"] = dict(sg.last_split or {})
    _, idsh, _ = sg.search(q.cpu().numpy(), 100)
    out["split_k100_host_ids_equal"] = bool(torch.equal(idsh, idsc))
    sg.split_min_k = 1                                  # the same form at k = 10
    _, idsd, s64d = sg.search(q, k)
    out["split_k10_ids_equal"] = bool(torch.equal(idsd, ids) and torch.equal(s64d, s64))
    sg.split_min_k = 32

    # ---- delegates over shards vs one gallery
    loc = labels[a:b]
    order = torch.argsort(loc, stable=True)
    off = torch.zeros(n_cls + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(torch.bincount(loc, minlength=n_cls), 0)
    cent = sg.segment_mean(off, order)
    order_w = torch.argsort(labels, stable=True)
    off_w = torch.zeros(n_cls + 1, dtype=torch.int64, device=dev)
    off_w[1:] = torch.cumsum(torch.bincount(labels, minlength=n_cls), 0)
    want = whole.segment_mean(off_w, row_idx=order_w)
    ulp = (cent.view(torch.int32).long() - want.view(torch.int32).long()).abs().max()
    out["delegate_max_ulp"] = int(ulp)
    out["delegate_empty_class_zero"] = bool((cent[7] == 0).all())
    whole.close()

    # ---- search, EUCLID fp32 (K3 with the row-bias epilogue + key merge)
    n2, dim2 = 60_001, 256
    x2, q2 = x[:n2, :dim2].contiguous(), q[:64, :dim2].contiguous()
    a2, b2 = shard_range(n2, rank, world)
    se = ShardedGallery(dim2, n2, dtype="f32", metric="euclid", device=local)
    se.upsert_local(x2[a2:b2])
    d32, ids2, keys = se.search(q2, k)
    top2 = torch.topk(torch.cdist(q2.double(), x2.double()), k, dim=1, largest=False)
    out["euclid_ids_equal"] = bool(torch.equal(top2.indices, ids2))
    out["euclid_max_rel_err"] = float(((top2.values - d32.double()).abs() / top2.values.clamp_min(1e-30)).max())
    d40, ids40, _ = se.search(q2, 40)                   # split form on a EUCLID collection
    top40 = torch.topk(torch.cdist(q2.double(), x2.double()), 40, dim=1, largest=False)
    out["euclid_split_k40_ids_equal"] = bool(torch.equal(top40.indices, ids40)) and se.last_split is not None

    ok = (out["cosine_ids_equal"] and out["host_queries_ids_equal"] and out["cosine_max_score_err"] < 1e-9 and out["delegate_max_ulp"] <= 1
          and out["delegate_empty_class_zero"] and out["euclid_ids_equal"] and out["euclid_max_rel_err"] < 1e-6
          and out["split_k100_ids_equal"] and out["split_k100_max_score_err"] < 1e-9 and out["split_k100_host_ids_equal"]
          and out["split_k10_ids_equal"] and out["euclid_split_k40_ids_equal"])
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    out["ok_all_ranks"] = int(flag.item()) == 0
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 0 else 1)


if __name__ == "__main__":
    main()
