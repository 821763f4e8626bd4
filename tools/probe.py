#!/usr/bin/env python
"""GPU probe: times one kernel family at a given shape and prints one JSON line per measurement
(flushed immediately, so a run that is cut off still leaves what it measured).

    python tools/probe.py search --rows 10000000 --dim 768 --dtype bf16 --queries 1,16,128,1024 --k 10
    python tools/probe.py k1 --rows 10000000 --dim 768 --dtype bf16
    python tools/probe.py k2 --rows 1000000 --dim 768 --classes 10000 [--zipf]
    python tools/probe.py merge --shards 8 --queries 10000 --k 100

Not a benchmark of record (bench.py is); used to steer kernel work and for the ncu captures.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def time_ms(fn, iters, warmup=2):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def build_gallery(a, dev):
    import torch

    from retrieval_based_object_detection_b200 import Gallery

    g = Gallery(a.dim, dtype=a.dtype, capacity=a.rows, device=0)
    for kv in a.opt:
        key, _, val = kv.partition("=")
        g.set_option(key, int(val))
    gen = torch.Generator(dev).manual_seed(1234)
    chunk = 500_000
    for s in range(0, a.rows, chunk):
        g.upsert(torch.randn(min(chunk, a.rows - s), a.dim, device=dev, generator=gen))
    torch.cuda.synchronize()
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("op", choices=["search", "k1", "k2", "merge", "copy", "dist"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--queries", default="1024")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--classes", type=int, default=10_000)
    ap.add_argument("--zipf", action="store_true")
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--mask-frac", type=float, default=0.0, help="search with a random row mask keeping this fraction")
    ap.add_argument("--check", action="store_true", help="verify ids against a torch fp64 brute force (small shapes)")
    ap.add_argument("--opt", action="append", default=[])
    a = ap.parse_args()

    import torch

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    hbm = 6555.8
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        hbm = json.load(open(peaks)).get("hbm_gbs", hbm)

    if a.op == "copy":
        # this box's copy bandwidth, measured the way MEASURED_PEAKS.json was (read + write bytes of b.copy_(a)):
        # boxes differ by several percent, so kernel bandwidths are best compared with this number
        n = 1 << 30
        src = torch.empty(n, dtype=torch.bfloat16, device=dev).normal_()
        dst = torch.empty_like(src)
        best = min(time_ms(lambda: dst.copy_(src), 3, warmup=1) for _ in range(5))
        rd = min(time_ms(lambda: src.sum(), 3, warmup=1) for _ in range(3))
        # same traffic mix as K1 with a 16-bit gallery: read fp32, write 16 bit (torch's own elementwise cast)
        m = 1 << 29
        s32 = torch.empty(m, dtype=torch.float32, device=dev).normal_()
        d16 = torch.empty(m, dtype=torch.bfloat16, device=dev)
        cast = min(time_ms(lambda: d16.copy_(s32), 3, warmup=1) for _ in range(5))
        emit(op="copy", bytes=4 * n, ms=round(best, 4), copy_gbs=round(4 * n / best / 1e6, 1),
             read_only_gbs=round(2 * n / rd / 1e6, 1), cast_f32_to_bf16_gbs=round(6 * m / cast / 1e6, 1),
             measured_peaks_gbs=hbm)
        return
    if a.op == "dist":
        from retrieval_based_object_detection_b200 import Gallery

        for metric in ("euclid", "manhattan"):
            g = Gallery(a.dim, dtype=a.dtype, metric=metric, capacity=a.rows, device=0)
            gen = torch.Generator(dev).manual_seed(3)
            for s0 in range(0, a.rows, 500_000):
                g.upsert(torch.randn(min(500_000, a.rows - s0), a.dim, device=dev, generator=gen))
            for Q in [int(x) for x in a.queries.split(",")]:
                q = torch.randn(Q, a.dim, device=dev, generator=torch.Generator(dev).manual_seed(7))
                stats = {}

                def run():
                    stats.update(g.search(q, a.k).stats)

                ms = time_ms(run, max(1, a.iters // 2), warmup=1)
                pair_elems = float(Q) * a.rows * a.dim
                emit(op="dist", metric=metric, rows=a.rows, dim=a.dim, dtype=a.dtype, Q=Q, k=a.k, ms=round(ms, 3),
                     qps=round(Q / ms * 1e3, 1), fp64_tflops=round(2 * pair_elems / ms / 1e9, 2),
                     launches=stats.get("total_launches"), resweeps=stats.get("fallback_queries"))
            g.close()
        return
    if a.op == "search":
        g = build_gallery(a, dev)
        g.set_option("time_k3", 1)
        mask = None
        if a.mask_frac > 0:
            bits = torch.rand(a.rows, device=dev) < a.mask_frac
            pad = (-a.rows) % 32
            b = torch.cat([bits, torch.zeros(pad, dtype=torch.bool, device=dev)]).view(-1, 32).to(torch.int64)
            w = (b << torch.arange(32, device=dev, dtype=torch.int64)).sum(1)
            mask = (w & 0xFFFFFFFF).to(torch.int64)
            mask = torch.where(mask >= 2 ** 31, mask - 2 ** 32, mask).to(torch.int32)
        esz = 4 if a.dtype in ("f32", "fp32") else 2
        for Q in [int(x) for x in a.queries.split(",")]:
            q = torch.randn(Q, a.dim, device=dev, generator=torch.Generator(dev).manual_seed(7))
            out = (torch.empty((Q, a.k), dtype=torch.float32, device=dev),
                   torch.empty((Q, a.k), dtype=torch.int64, device=dev),
                   torch.empty((Q, a.k), dtype=torch.float64, device=dev))
            stats = {}

            def run():
                stats.update(g.search(q, a.k, row_mask=mask, out=out).stats)

            ms = time_ms(run, a.iters)
            flop = 2.0 * Q * a.rows * a.dim
            bytes_ = a.rows * a.dim * 2 + Q * a.dim * 2 + Q * a.k * 12
            rec = dict(op="search", rows=a.rows, dim=a.dim, dtype=a.dtype, Q=Q, k=a.k, ms=round(ms, 4),
                       k3_ms=round(stats.get("k3_ms", 0.0), 4), qps=round(Q / ms * 1e3, 1),
                       tflops=round(flop / ms / 1e9, 1), gbs=round(bytes_ / ms / 1e6, 1),
                       hbm_frac=round(bytes_ / ms / 1e6 / hbm, 3), slices=stats.get("slices"),
                       kc=stats.get("candidates"), fallback=stats.get("fallback_queries"),
                       launches=stats.get("total_launches"), max_eps=stats.get("max_eps"), opts=a.opt)
            if a.check:
                stored = g.get_rows(torch.arange(a.rows, device=dev)).double()
                qn = q.double()
                sc = (qn @ stored.T) / (qn.norm(dim=1, keepdim=True) * stored.norm(dim=1)[None, :])
                if mask is not None:
                    sc = torch.where(bits[None, :], sc, torch.full_like(sc, -float("inf")))
                top = torch.topk(sc, min(a.k, a.rows), dim=1)
                rec["ids_equal"] = bool((top.indices == out[1][:, : top.indices.shape[1]]).all().item())
                rec["max_score_err"] = float((top.values - out[2][:, : top.values.shape[1]]).abs().max().item())
            emit(**rec)
    elif a.op == "k1":
        from retrieval_based_object_detection_b200 import l2norm_pack

        n = a.rows
        x = torch.randn(n, a.dim, device=dev)
        for od in [a.dtype]:
            ms = time_ms(lambda: l2norm_pack(x, od), a.iters)
            osz = 4 if od == "f32" else 2
            b = n * a.dim * (4 + osz)
            emit(op="k1", rows=n, dim=a.dim, out=od, ms=round(ms, 4), gbs=round(b / ms / 1e6, 1),
                 hbm_frac=round(b / ms / 1e6 / hbm, 3), alg_bytes=b)
        from retrieval_based_object_detection_b200 import Gallery

        g = Gallery(a.dim, dtype=a.dtype, capacity=n, device=0)

        def ups():
            g.truncate(0)
            g.upsert(x)

        ms = time_ms(ups, a.iters)
        osz = 2 + (4 if a.dtype == "f32" else 0)
        b = n * a.dim * (4 + osz)
        emit(op="k1_upsert", rows=n, dim=a.dim, gallery=a.dtype, ms=round(ms, 4), gbs=round(b / ms / 1e6, 1),
             hbm_frac=round(b / ms / 1e6 / hbm, 3), alg_bytes=b)
    elif a.op == "k2":
        from retrieval_based_object_detection_b200 import Gallery

        n, C = a.rows, a.classes
        g = Gallery(a.dim, dtype=a.dtype, capacity=n, device=0)
        gen = torch.Generator(dev).manual_seed(5)
        for s in range(0, n, 500_000):
            g.upsert(torch.randn(min(500_000, n - s), a.dim, device=dev, generator=gen))
        if a.zipf:
            w = 1.0 / torch.arange(1, C + 1, device=dev, dtype=torch.float64)
            labels = torch.multinomial(w / w.sum(), n, replacement=True, generator=gen)
        else:
            labels = torch.arange(n, device=dev) % C
            labels = labels[torch.randperm(n, device=dev, generator=gen)]
        order = torch.argsort(labels, stable=True)
        offsets = torch.zeros(C + 1, dtype=torch.int64, device=dev)
        offsets[1:] = torch.cumsum(torch.bincount(labels, minlength=C), 0)
        esz = 4 if a.dtype == "f32" else 2
        for name, idx in (("gathered", order), ("contiguous", None)):
            ms = time_ms(lambda: g.segment_mean(offsets, row_idx=idx), a.iters)
            b = n * a.dim * esz + (n * 8 if idx is not None else 0) + C * a.dim * 4
            emit(op="k2", mode=name, rows=n, dim=a.dim, classes=C, zipf=a.zipf, dtype=a.dtype, ms=round(ms, 4),
                 gbs=round(b / ms / 1e6, 1), hbm_frac=round(b / ms / 1e6 / hbm, 3), alg_bytes=b,
                 max_class=int((offsets[1:] - offsets[:-1]).max().item()))
    elif a.op == "merge":
        from retrieval_based_object_detection_b200 import merge_topk

        for Q in [int(x) for x in a.queries.split(",")]:
            s = torch.rand(a.shards, Q, a.k, device=dev, dtype=torch.float64).sort(dim=2, descending=True).values
            ids = torch.randint(0, 1 << 40, (a.shards, Q, a.k), device=dev)
            ms = time_ms(lambda: merge_topk(s, ids, a.k), a.iters)
            b = a.shards * Q * a.k * 16 + Q * a.k * 20
            emit(op="merge", shards=a.shards, Q=Q, k=a.k, ms=round(ms, 4), gbs=round(b / ms / 1e6, 1))


if __name__ == "__main__":
    main()
