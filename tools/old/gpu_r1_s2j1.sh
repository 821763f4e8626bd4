#!/bin/bash
# session-2 GPU pass J1 (1 GPU): K2b delegate kernels, 8-group pre-pass, C5 query sweep on one 12.5M x 768 fp16 shard
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/s2j_pytest.log
timeout 900 python bench.py --rows 12500000 --dtype f16 --k 10 --sweep 1,2,4,8,16,32,64,128,256,512,1024,2048,4096,8192,16384,32768,65536 > $O/s2j_sweep_c5.json 2> $O/s2j_sweep_c5.err; tail -2 $O/s2j_sweep_c5.err; cat $O/s2j_sweep_c5.json
timeout 400 python bench.py --no-cpu-baseline > $O/s2j_bench.json 2> $O/s2j_bench.err; cat $O/s2j_bench.json | cut -c1-300
