#!/bin/bash
# session-2 GPU pass J2 (8 GPUs): parity (1 GPU), k=100 sharded bench with per-rank diagnostics, with/without shadow16
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee $O/s2j2_pytest.log
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29800+RANDOM%100)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@"; }
run 8 --k 100 2> $O/s2j2_n8_k100.err | tail -1 > $O/s2j2_n8_k100.json; cat $O/s2j2_n8_k100.json
run 8 --k 100 --opt shadow16=1 2> $O/s2j2_n8_k100_shadow.err | tail -1 > $O/s2j2_n8_k100_shadow.json; cat $O/s2j2_n8_k100_shadow.json
run 8 2> $O/s2j2_n8.err | tail -1 > $O/s2j2_n8.json; cat $O/s2j2_n8.json
