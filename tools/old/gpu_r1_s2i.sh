#!/bin/bash
# session-2 GPU pass I (2 GPUs): new tests (C2/C3 full size, shadow16, pre-pass retry), k=100 with and without the
# fp16 shadow, per-rank diagnostics of the k=100 sharded bench, smoke
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/s2i_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee $O/s2i_smoke.log
( P="timeout 200 python tools/probe.py search --dim 768 --k 100 --queries 10000 --iters 2"
$P --rows 1250000 2>&1 | tail -1
$P --rows 1250000 --opt shadow16=1 2>&1 | tail -1
$P --rows 4000000 --opt shadow16=1 2>&1 | tail -1
timeout 200 python tools/probe.py search --dim 768 --k 10 --queries 10000 --iters 3 --rows 4000000 --opt shadow16=1 2>&1 | tail -1 ) | tee $O/s2i_probe.jsonl
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+RANDOM%100)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline "$@"; }
run 2 --rows 2500000 --k 100 2> $O/s2i_n2_k100.err | tail -1 > $O/s2i_n2_k100.json; cat $O/s2i_n2_k100.json
run 2 --rows 2500000 --k 100 --opt shadow16=1 2> $O/s2i_n2_k100_shadow.err | tail -1 > $O/s2i_n2_k100_shadow.json; cat $O/s2i_n2_k100_shadow.json
