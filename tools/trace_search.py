#!/usr/bin/env python
"""Host-side anatomy of one rbod_search call (RBOD_TRACE=1): builds a gallery, runs a few searches and lets the library
print where the wall-clock time of the call goes, next to the CUDA-event time of the whole call."""
import os
import sys

os.environ["RBOD_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from retrieval_based_object_detection_b200 import Gallery  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_250_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda", 0)
g = Gallery(768, dtype="bf16", capacity=rows, device=0)
gen = torch.Generator(dev).manual_seed(1)
for s in range(0, rows, 500_000):
    g.upsert(torch.randn(min(500_000, rows - s), 768, device=dev, generator=gen))
g.set_option("time_k3", 1)
q = torch.randn(Q, 768, device=dev, generator=gen)
out = (torch.empty((Q, k), dtype=torch.float32, device=dev), torch.empty((Q, k), dtype=torch.int64, device=dev),
       torch.empty((Q, k), dtype=torch.float64, device=dev))
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    st = g.search(q, k, out=out).stats
    e1.record()
    torch.cuda.synchronize()
    print(f"call {i}: events {e0.elapsed_time(e1):.3f} ms, k3 {st['k3_ms']:.3f} ms, launches {st['total_launches']}", file=sys.stderr)
