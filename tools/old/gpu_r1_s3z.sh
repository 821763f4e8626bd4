#!/bin/bash
# session-3 GPU pass Z (1 GPU): evidence of record with the round's final code -- bench line, launch list, full-set ncu
# captures of the main K3 launch (bench shape), K1 and K2, summarised on the box
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 500 python bench.py > $O/s3z_bench.json 2> $O/s3z_bench.err; tail -2 $O/s3z_bench.err; cut -c1-300 $O/s3z_bench.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/s3z_plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches_bench.csv $B > $O/s3z_ncu_bench.log 2>&1
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $B1 > $O/s3z_plain_bench1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_cosine -s 7 -c 1 -f -o $O/r01_k3bench_full $B1 > $O/s3z_ncu_k3bench.log 2>&1
K1="python tools/probe.py k1 --rows 4000000 --dim 768 --dtype bf16 --iters 1"
timeout 300 $K1 > $O/s3z_plain_k1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2norm_pack -s 2 -c 1 -f -o $O/r01_k1_full $K1 > $O/s3z_ncu_k1.log 2>&1
K2="python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --iters 1"
timeout 300 $K2 > $O/s3z_plain_k2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:seg_mean -s 2 -c 1 -f -o $O/r01_k2_full $K2 > $O/s3z_ncu_k2.log 2>&1
mkdir -p $O/prof
python tools/make_profiles.py r01 && cp profiles/r01_k3bench_ncu.txt profiles/r01_k1_ncu.txt profiles/r01_k2_ncu.txt profiles/r01_launches_bench.csv profiles/r01_launches_bench_summary.txt profiles/k3_traffic.json $O/prof/
cat profiles/k3_traffic.json; head -12 profiles/r01_k1_ncu.txt | cut -c1-150
rm -f $O/r01_k1_full.ncu-rep $O/r01_k2_full.ncu-rep $O/r01_k3bench_full.ncu-rep
