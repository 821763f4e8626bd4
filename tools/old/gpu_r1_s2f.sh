#!/bin/bash
# session-2 GPU pass F: threshold pre-pass, K2 tree reduction / unroll
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/s2f_pytest.log
( timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 2>&1 | tail -2
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --zipf 2>&1 | tail -2
timeout 200 python tools/probe.py k2 --rows 4000000 --dim 768 --dtype bf16 --classes 10000 2>&1 | tail -2
P="timeout 200 python tools/probe.py search --dim 768 --k 10 --queries 10000 --iters 3"
$P --rows 4000000 2>&1 | tail -1
$P --rows 4000000 --opt presample=0 2>&1 | tail -1
$P --rows 1250000 2>&1 | tail -1
$P --rows 1250000 --opt presample=0 2>&1 | tail -1
$P --rows 4000000 --dim 512 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 1250000 --dim 768 --k 100 --queries 10000 --iters 2 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 1250000 --dim 768 --k 100 --queries 10000 --iters 2 --opt presample=0 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 1000000 --dim 512 --dtype f32 --k 10 --queries 10000 --iters 3 2>&1 | tail -1 ) | tee $O/s2f_probe.jsonl
timeout 400 python bench.py --no-cpu-baseline > $O/s2f_bench.json 2> $O/s2f_bench.err; tail -2 $O/s2f_bench.err; cat $O/s2f_bench.json
