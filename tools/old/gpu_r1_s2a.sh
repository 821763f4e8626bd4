#!/bin/bash
# session-2 GPU pass A: parity, k=100 diagnosis, small-Q latency, K1/K2 bandwidth, default bench
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/s2a_pytest.log
timeout 200 python tools/probe.py search --rows 1000000 --dim 768 --k 100 --queries 1024 --iters 2 > $O/s2a_k100_1m.jsonl 2>&1; tail -3 $O/s2a_k100_1m.jsonl
timeout 200 python tools/probe.py search --rows 1000000 --dim 768 --k 10 --queries 1024,10000 --iters 3 > $O/s2a_k10_1m.jsonl 2>&1; tail -3 $O/s2a_k10_1m.jsonl
timeout 300 python tools/probe.py search --rows 10000000 --dim 768 --k 10 --queries 1,16,128,512,2048 --iters 3 > $O/s2a_smallq.jsonl 2>&1; tail -6 $O/s2a_smallq.jsonl
timeout 200 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 > $O/s2a_k1.jsonl 2>&1; tail -3 $O/s2a_k1.jsonl
timeout 200 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 >> $O/s2a_k1.jsonl 2>&1; tail -2 $O/s2a_k1.jsonl
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 > $O/s2a_k2.jsonl 2>&1; tail -3 $O/s2a_k2.jsonl
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --zipf >> $O/s2a_k2.jsonl 2>&1; tail -2 $O/s2a_k2.jsonl
timeout 400 python bench.py > $O/s2a_bench.json 2> $O/s2a_bench.err; tail -2 $O/s2a_bench.err; cat $O/s2a_bench.json
