#!/bin/bash
# session-3 GPU pass Z2 (1 GPU): launch list of the bench command without the exactness checker's torch kernels
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity"
timeout 300 $B > $O/s3z2_plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches_bench.csv $B > $O/s3z2_ncu_bench.log 2>&1
mkdir -p $O/prof; python tools/make_profiles.py r01 && cp profiles/r01_launches_bench.csv profiles/r01_launches_bench_summary.txt $O/prof/
cat profiles/r01_launches_bench_summary.txt | cut -c1-150
