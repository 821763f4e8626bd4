#!/bin/bash
# session-2 GPU pass K (1 GPU): full parity suite, final bench line, launch list and full-set captures of the
# final K1 / K2 / K3 (bench shape) -> profiles/
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/s2k_pytest.log
timeout 400 python bench.py > $O/s2k_bench.json 2> $O/s2k_bench.err; tail -2 $O/s2k_bench.err; cat $O/s2k_bench.json | cut -c1-400
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/s2k_plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches_bench.csv $B > $O/s2k_ncu_bench.log 2>&1
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $B1 > $O/s2k_plain_bench1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_cosine -s 7 -c 1 -f -o $O/r01_k3bench_full $B1 > $O/s2k_ncu_k3bench.log 2>&1
K1="python tools/probe.py k1 --rows 4000000 --dim 768 --dtype bf16 --iters 1"
timeout 300 $K1 > $O/s2k_plain_k1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2norm_pack -s 2 -c 1 -f -o $O/r01_k1_full $K1 > $O/s2k_ncu_k1.log 2>&1
K2="python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --iters 1"
timeout 300 $K2 > $O/s2k_plain_k2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:seg_mean -s 2 -c 1 -f -o $O/r01_k2_full $K2 > $O/s2k_ncu_k2.log 2>&1
tail -2 $O/s2k_plain_k1.log $O/s2k_plain_k2.log
ls -la $O | grep r01_
