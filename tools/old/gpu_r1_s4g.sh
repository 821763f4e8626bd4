#!/bin/bash
# session-3 GPU pass 4g (1 GPU): full suite on the round's final code, bench line with the side measurements of K1 / K2
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300 | tee $O/s4g_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 500 python bench.py > $O/s4g_bench.json 2> $O/s4g_bench.err; tail -2 $O/s4g_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s4g_bench.json").read())
print({k:d[k] for k in ("value","ms_per_step","gpu_launches","parity","clocks")})
print(d["e2e"]); print(d["roofline"]); print(d["other_kernels"]); print(d["cpu_baseline"])
PY
