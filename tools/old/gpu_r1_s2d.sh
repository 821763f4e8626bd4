#!/bin/bash
# session-2 GPU pass D: collect pass parity, k=100, bench, ncu launch list + full captures of K3/K1/K2
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/s2d_pytest.log
( timeout 200 python tools/probe.py search --rows 1000000 --dim 768 --k 100 --queries 1024 --iters 2 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 4000000 --dim 768 --k 100 --queries 10000 --iters 2 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 1250000 --dim 768 --k 100 --queries 10000 --iters 2 2>&1 | tail -1 ) | tee $O/s2d_probe.jsonl
timeout 400 python bench.py > $O/s2d_bench.json 2> $O/s2d_bench.err; tail -2 $O/s2d_bench.err; cat $O/s2d_bench.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/s2d_plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches_bench.csv $B > $O/s2d_ncu_bench.log 2>&1
S="python tools/probe.py search --rows 2000000 --dim 768 --k 10 --queries 8192 --iters 1"
timeout 300 $S > $O/s2d_plain_k3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_cosine -s 2 -c 1 -f -o $O/r01_k3_full $S > $O/s2d_ncu_k3.log 2>&1
K1="python tools/probe.py k1 --rows 4000000 --dim 768 --dtype bf16 --iters 1"
timeout 300 $K1 > $O/s2d_plain_k1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2norm_pack -s 2 -c 1 -f -o $O/r01_k1_full $K1 > $O/s2d_ncu_k1.log 2>&1
K2="python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --iters 1"
timeout 300 $K2 > $O/s2d_plain_k2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:seg_mean -s 2 -c 1 -f -o $O/r01_k2_full $K2 > $O/s2d_ncu_k2.log 2>&1
ls -la $O | grep r01_
