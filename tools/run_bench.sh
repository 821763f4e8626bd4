#!/bin/bash
# usage: run_bench.sh tag args...
tag=$1; shift
timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
tail -2 gpurun_out/bench_$tag.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$tag.json"))
    print("$tag", round(d["value"]), round(d["e2e"]["value"]), round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "k3_ms", round(d["roofline"]["kernel_ms"],2), "slices", d["config"]["slices"], "fb", d["config"]["fallback_queries_per_step"])
except Exception as e:
    print("$tag FAILED", e)
PY
