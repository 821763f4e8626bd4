#!/bin/bash
# session-3 GPU pass I (1 GPU): evidence of record for this round -- full parity suite, smoke, the bench line (with CPU
# baseline), reference arm, launch list and full-set ncu captures of K3 (bench shape) / K1 / K2 / K5, summarised on the box
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 | tee $O/s3i_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee $O/s3i_smoke.log
timeout 500 python bench.py > $O/s3i_bench.json 2> $O/s3i_bench.err; tail -2 $O/s3i_bench.err; cut -c1-600 $O/s3i_bench.json
timeout 500 python bench.py --impl reference > $O/s3i_bench_ref.json 2>> $O/s3i_bench.err; cut -c1-300 $O/s3i_bench_ref.json
timeout 300 python bench.py --k 100 --no-cpu-baseline > $O/s3i_bench_k100.json 2>> $O/s3i_bench.err; cut -c1-300 $O/s3i_bench_k100.json
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/s3i_plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r01_launches_bench.csv $B > $O/s3i_ncu_bench.log 2>&1
B1="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $B1 > $O/s3i_plain_bench1.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_cosine -s 4 -c 1 -f -o $O/r01_k3bench_full $B1 > $O/s3i_ncu_k3bench.log 2>&1
K1="python tools/probe.py k1 --rows 4000000 --dim 768 --dtype bf16 --iters 1"
timeout 300 $K1 > $O/s3i_plain_k1.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2norm_pack -s 2 -c 1 -f -o $O/r01_k1_full $K1 > $O/s3i_ncu_k1.log 2>&1
K2="python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --iters 1"
timeout 300 $K2 > $O/s3i_plain_k2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:seg_mean -s 2 -c 1 -f -o $O/r01_k2_full $K2 > $O/s3i_ncu_k2.log 2>&1
K5="python tools/probe.py dist --rows 1000000 --dim 512 --dtype f32 --queries 32 --k 10 --iters 2"
timeout 300 $K5 > $O/s3i_plain_k5.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_collect -s 3 -c 1 -f -o $O/r01_k5_full $K5 > $O/s3i_ncu_k5.log 2>&1
tail -2 $O/s3i_plain_k1.log $O/s3i_plain_k2.log $O/s3i_plain_k5.log
mkdir -p $O/prof
python tools/make_profiles.py r01 && cp profiles/r01_k3bench_ncu.txt profiles/r01_k1_ncu.txt profiles/r01_k2_ncu.txt profiles/r01_k5_ncu.txt profiles/r01_launches_bench.csv profiles/r01_launches_bench_summary.txt profiles/k3_traffic.json $O/prof/ 2>&1 | tail -3
ls -la $O | grep ncu-rep
rm -f $O/r01_k1_full.ncu-rep $O/r01_k2_full.ncu-rep $O/r01_k5_full.ncu-rep
