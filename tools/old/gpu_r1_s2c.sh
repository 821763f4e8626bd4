#!/bin/bash
# session-2 GPU pass C: shared threshold across slices + more slices for small Q
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee $O/s2c_pytest.log
P="timeout 200 python tools/probe.py search --rows 4000000 --dim 768 --k 10 --queries 10000 --iters 3"
( $P --opt tau_share=1 2>&1 | tail -1
$P --opt tau_share=0 2>&1 | tail -1
$P --opt tau_share=1 --opt k3_variant=2 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 4000000 --dim 512 --k 10 --queries 10000 --iters 3 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 1000000 --dim 768 --k 100 --queries 1024 --iters 2 2>&1 | tail -1
timeout 200 python tools/probe.py search --rows 4000000 --dim 768 --k 100 --queries 10000 --iters 2 --dtype f16 2>&1 | tail -1
timeout 300 python tools/probe.py search --rows 10000000 --dim 768 --k 10 --queries 1,16,128,512 --iters 3 2>&1 | tail -4 ) | tee $O/s2c_probe.jsonl
