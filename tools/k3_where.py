#!/usr/bin/env python
"""Where does K3's time go?  Builds one gallery, then for each option set runs the headline search with the
wait-cycle counters on (option k3_prof, rbod_debug_profile) and prints, per warp role, the share of the CTA's
cycles it spent waiting -- the MMA issuer's shares say whether the tensor pipe starves on gallery data (TMA / L2 /
pipeline depth) or on accumulators (epilogue).

    python tools/k3_where.py [--rows N] [--queries Q] [--k K] [--dtype bf16] [--sets "a=1,b=2;c=3"]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--sets", default=";k3_kbs=2;k3_variant=2;hybrid=0;l2_sync=0",
                    help="';'-separated option sets, each 'key=value,key=value' (empty = defaults)")
    a = ap.parse_args()

    import torch

    from retrieval_based_object_detection_b200 import Gallery

    dev = torch.device("cuda", 0)
    g = Gallery(a.dim, dtype=a.dtype, capacity=a.rows, device=0)
    gen = torch.Generator(dev).manual_seed(1234)
    for s in range(0, a.rows, 500_000):
        g.upsert(torch.randn(min(500_000, a.rows - s), a.dim, device=dev, generator=gen))
    q = torch.randn(a.queries, a.dim, device=dev, generator=torch.Generator(dev).manual_seed(99))
    out = (torch.empty((a.queries, a.k), dtype=torch.float32, device=dev), torch.empty((a.queries, a.k), dtype=torch.int64, device=dev),
           torch.empty((a.queries, a.k), dtype=torch.float64, device=dev))
    g.set_option("time_k3", 1)
    defaults = {"k3_kbs": 0, "k3_variant": -1, "hybrid": 1, "l2_sync": 1, "presample": 1, "tau_share": 1, "debug_epi": 0,
                "sync_window": 16, "sync_lead": 4}
    ref_rows = None
    for spec in a.sets.split(";"):
        opts = dict(defaults)
        for kv in [x for x in spec.replace("+", ",").split(",") if x]:
            key, _, val = kv.partition("=")
            opts[key] = int(val)
        for key, val in opts.items():
            g.set_option(key, val)
        g.set_option("k3_prof", 0)
        for _ in range(2):
            g.search(q, a.k, out=out)
        g.set_option("k3_prof", 1)
        g.debug_profile()
        ms, k3 = [], []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            st = g.search(q, a.k, out=out).stats
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
            k3.append(st["k3_ms"])
        p = g.debug_profile()
        same = None
        if opts.get("debug_epi", 0) == 0:
            if ref_rows is None:
                ref_rows = out[1].clone()
            same = bool((ref_rows == out[1]).all().item())
        cyc, epi = max(p["cta_cycles"], 1), max(p["epi_warps"], 1)
        ctas = max(p["ctas"], 1)
        rec = {"opts": spec or "defaults", "ms_p50": round(sorted(ms)[len(ms) // 2], 3), "ms": round(min(ms), 3), "k3_ms": round(min(k3), 3),
               "tflops": round(2.0 * a.queries * a.rows * a.dim / (min(k3) / 1e3) / 1e12, 1), "slices": st["slices"],
               "kc": st["candidates"], "same_ids_as_first": same,
               "mma_wait_data": round(p["mma_wait_data"] / cyc, 4), "mma_wait_accumulator": round(p["mma_wait_accumulator"] / cyc, 4),
               "mma_wait_query_tile": round(p["mma_wait_query_tile"] / cyc, 4),
               "prod_wait_empty": round(p["prod_wait_empty"] / cyc, 4), "prod_wait_empty_follower": round(p["prod_wait_empty_follower"] / cyc, 4),
               "prod_issue": round(p["prod_issue"] / cyc, 4), "mma_issue": round(p["mma_issue"] / cyc, 4), "prod_wait_throttle": round(p["prod_wait_throttle"] / cyc, 4),
               "epi_wait_accumulator": round(p["epi_wait_accumulator"] / (cyc * epi / ctas), 4),
               "epi_prune": round(p["epi_prune"] / (cyc * epi / ctas), 4), "prunes": p["prunes"], "ctas": p["ctas"],
               "mcycles_per_cta": round(cyc / ctas / 1e6 / max(a.iters, 1), 3)}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
