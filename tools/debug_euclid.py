"""Bring-up helper: the EUCLID fp32 scenario of tests/test_gpu_search.py with the library's switches toggled one at a
time, printing which queries differ from torch.cdist and through which path they were answered."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from retrieval_based_object_detection_b200 import Gallery  # noqa: E402


def main():
    dtype = sys.argv[1] if len(sys.argv) > 1 else "f32"
    n, dim, Q, k = 300_000, 512, 1500, 10
    gen = torch.Generator("cuda").manual_seed(41)
    x = torch.randn(n, dim, device="cuda", generator=gen) * (0.05 + 2.0 * torch.rand(n, 1, device="cuda", generator=gen))
    g = Gallery(dim, dtype=dtype, metric="euclid", capacity=1000)
    for a in range(0, n, 100_000):
        g.upsert(x[a:a + 100_000])
    stored = g.get_rows(torch.arange(n, device="cuda"))
    q = torch.randn(Q, dim, device="cuda", generator=gen) * (0.1 + torch.rand(Q, 1, device="cuda", generator=gen))
    q[:50] = stored[1000:1050] + 0.01 * torch.randn(50, dim, device="cuda", generator=gen)
    d = torch.cdist(q.double(), stored.double())
    top = torch.topk(d, k, dim=1, largest=False)
    for opts in ({}, {"presample": 0}, {"tau_share": 0}, {"collect_pass": 0}, {"presample": 0, "tau_share": 0, "collect_pass": 0}):
        for key in ("presample", "tau_share", "collect_pass"):
            g.set_option(key, 1)
        for key, val in opts.items():
            g.set_option(key, val)
        r = g.search(q, k, want_scores64=True)
        bad = (top.indices != r.rows).any(dim=1)
        nb = int(bad.sum())
        print(opts, "bad", nb, {s: r.stats[s] for s in ("fallback_queries", "sweep_queries", "presample_retries", "k3_launches",
                                                         "candidates", "slices", "max_eps")})
        if nb:
            i = int(torch.nonzero(bad)[0])
            print("  first bad query", i, "want", top.indices[i].tolist(), "got", r.rows[i].tolist())
            print("  want d", [round(v, 5) for v in top.values[i].tolist()], "got d", [round(v, 5) for v in r.scores[i].tolist()])
            print("  bad query indices (first 20):", torch.nonzero(bad)[:20, 0].tolist())


if __name__ == "__main__":
    main()
