#!/bin/bash
# session-2 GPU pass H (8 GPUs): strong-scaling bench at N=8 and N=4, k=100 at N=8 (config C4)
cd "$(dirname "$0")/.."
O=gpurun_out
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@"; }
run 8 > $O/s2h_bench_n8.json 2> $O/s2h_bench_n8.err; tail -2 $O/s2h_bench_n8.err; cat $O/s2h_bench_n8.json
run 4 > $O/s2h_bench_n4.json 2> $O/s2h_bench_n4.err; cat $O/s2h_bench_n4.json
run 8 --k 100 > $O/s2h_bench_n8_k100.json 2> $O/s2h_bench_n8_k100.err; cat $O/s2h_bench_n8_k100.json
run 8 --k 100 --queries 1024 > $O/s2h_bench_n8_k100_q1024.json 2> $O/s2h_bench_n8_k100_q1024.err; cat $O/s2h_bench_n8_k100_q1024.json
