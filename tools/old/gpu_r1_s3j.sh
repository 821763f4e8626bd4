#!/bin/bash
# session-3 GPU pass J (1 GPU): bench line with the exactness check, full-set ncu capture of the MAIN K3 launch
# (each search now launches the pre-pass first: skip an odd number of K3 launches)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 500 python bench.py > $O/s3j_bench.json 2> $O/s3j_bench.err; tail -2 $O/s3j_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/s3j_bench.json").read())
print({k:d[k] for k in ("value","e2e","gpu_launches","parity","clocks")}, d["roofline"]["frac"])
PY
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $B1 > $O/s3j_plain_bench1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_cosine -s 7 -c 1 -f -o $O/r01_k3bench_full $B1 > $O/s3j_ncu_k3bench.log 2>&1
mkdir -p $O/prof
python tools/make_profiles.py r01 && cp profiles/r01_k3bench_ncu.txt profiles/k3_traffic.json $O/prof/
cat profiles/k3_traffic.json
