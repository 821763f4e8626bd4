"""Bring-up aid: raw tcgen05 scores vs an fp64 matmul of the rounded operands, one variant per process.

    python tools/k3_debug.py <variant> [dtype] [n] [dim] [Q]
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_np as O  # noqa: E402  (debug tool, not product)
from retrieval_based_object_detection_b200 import Gallery  # noqa: E402


def main():
    variant = int(sys.argv[1])
    dtype = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    dim = int(sys.argv[4]) if len(sys.argv) > 4 else 768
    Q = int(sys.argv[5]) if len(sys.argv) > 5 else 130
    x = O.synthetic_unit_rows(n, dim, seed=1)
    g = Gallery(dim, dtype=dtype, capacity=n)
    g.upsert(x)
    stored = g.get_rows(np.arange(n))
    print("K1 mismatch fraction:", float(np.mean(stored != O.l2_normalize_store(x, dtype)[0])), g.info())
    g.set_option("k3_variant", variant)
    q = O.synthetic_unit_rows(Q, dim, seed=2)
    t0 = time.time()
    got = g.debug_scores(q)
    kind = "bf16" if dtype == "bf16" else "f16"
    qn = O.l2_normalize_store(q, kind)[0].astype(np.float64)
    want = qn @ O.round_store(stored, kind).astype(np.float64).T
    err = np.abs(got - want)
    print(f"variant {variant} {dtype} n={n} dim={dim} Q={Q}: max err {np.nanmax(err):.3e}, nan cells {int(np.isnan(got).sum())}, "
          f"{time.time() - t0:.2f}s")
    if np.nanmax(err) > 1e-4 or np.isnan(got).any():
        bad = np.argwhere(~(err < 1e-4))
        print("first bad cells (q,row):", bad[:10].tolist())
        print("got :", got[:2, :8])
        print("want:", want[:2, :8])
        rows_bad = np.unique(bad[:, 0])
        cols_bad = np.unique(bad[:, 1])
        print("bad query rows:", rows_bad[:20], "... count", len(rows_bad), " bad gallery cols:", cols_bad[:20], "... count", len(cols_bad))
        sys.exit(1)
    res = g.search(q, 10, want_scores64=True)
    ws, wi = O.cosine_topk(q, stored, 10)
    print("search ids identical:", bool(np.array_equal(res.rows, wi)), "stats", res.stats)


if __name__ == "__main__":
    main()
