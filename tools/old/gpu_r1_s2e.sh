#!/bin/bash
# session-2 GPU pass E: new tests (shim on device), K1 prefetch, warp-per-item K2, cooperative K3 launch,
# full-set ncu capture of K3 on the bench command (-> profiles/k3_traffic.json)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/s2e_pytest.log
( timeout 200 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 2>&1 | tail -2
timeout 200 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 2>&1 | tail -2
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 2>&1 | tail -2
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --zipf 2>&1 | tail -2
timeout 200 python tools/probe.py k2 --rows 4000000 --dim 768 --dtype bf16 --classes 10000 2>&1 | tail -2
timeout 200 python tools/probe.py search --rows 4000000 --dim 768 --k 10 --queries 10000 --iters 3 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 4000000 --dim 768 --k 10 --queries 10000 --iters 3 --opt k3_variant=2 2>&1 | tail -1
timeout 200 python tools/probe.py merge --shards 8 --queries 10000 --k 100 2>&1 | tail -1
timeout 200 python tools/probe.py merge --shards 8 --queries 10000 --k 10 2>&1 | tail -1 ) | tee $O/s2e_probe.jsonl
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 300 $B > $O/s2e_plain_bench.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k3_cosine -s 3 -c 1 -f -o $O/r01_k3bench_full $B > $O/s2e_ncu_k3bench.log 2>&1
tail -3 $O/s2e_ncu_k3bench.log
