#!/usr/bin/env python
"""Prints the few numbers of a bench.py JSON line that steer kernel work (used by tools/gpu_job.sh)."""
import json
import sys


def main():
    try:
        d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print("no bench line:", e)
        return
    if "sweep" in d:
        for r in d["sweep"]:
            print({k: r[k] for k in ("Q", "p50_ms", "min_ms", "qps", "bound", "frac_of_bound", "slices", "fallback") if k in r})
        return
    r, c = d.get("roofline", {}), d.get("config", {})
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 3), "k3_ms",
          round(r.get("kernel_ms", 0), 3), "TF/s", round(r.get("achieved", 0), 1), "frac", round(r.get("frac", 0), 4),
          "share", round(r.get("kernel_share_of_step") or 0, 4), "k", c.get("k"), "kc", c.get("candidates_per_query"),
          "slices", c.get("slices"), "fb", c.get("fallback_queries_per_step"), "clk", d.get("clocks", {}).get("sm_mhz"),
          d.get("clocks", {}).get("reasons"), "parity", (d.get("parity") or {}).get("ids_identical"))
    for o in d.get("other_kernels", []) or []:
        print("  other:", o.get("kernel"), round(o.get("achieved", 0), 1), o.get("unit"), "frac", round(o.get("frac", 0), 3),
              {k: v for k, v in o.items() if k in ("ok", "ms")})
    for o in d.get("configs", []) or []:
        print("  config:", json.dumps(o)[:700])


if __name__ == "__main__":
    main()
