#!/bin/bash
# session-3 GPU pass R (1 GPU): K5 with one row per lane -- tests, probe, ncu
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "distance" 2>&1 | tail -8 | cut -c1-300 | tee $O/s3r_pytest.log
P=$O/s3r_probe.jsonl; : > $P
timeout 300 python tools/probe.py dist --rows 1000000 --dim 512 --dtype f32 --queries 1,8,32,256,1024 --k 10 --iters 4 >> $P 2>>$O/s3r.err
timeout 300 python tools/probe.py dist --rows 1000000 --dim 768 --dtype bf16 --queries 256 --k 100 --iters 2 >> $P 2>>$O/s3r.err
cat $P
K5="python tools/probe.py dist --rows 1000000 --dim 512 --dtype f32 --queries 32 --k 10 --iters 2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dist_collect -s 3 -c 1 -f -o $O/r01_k5_full $K5 > $O/s3r_ncu_k5.log 2>&1
mkdir -p $O/prof; python tools/make_profiles.py r01 > /dev/null 2>&1; cp profiles/r01_k5_ncu.txt $O/prof/; rm -f $O/r01_k5_full.ncu-rep
head -32 profiles/r01_k5_ncu.txt | cut -c1-150
tail -3 $O/s3r.err
