#!/bin/bash
# session-2 GPU pass G (2 GPUs): parity on one GPU, then the 2-rank bench and the single-rank bench
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee $O/s2g_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $O/s2g_bench_n2.json 2> $O/s2g_bench_n2.err; tail -3 $O/s2g_bench_n2.err; cat $O/s2g_bench_n2.json
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/s2g_bench_n1.json 2> $O/s2g_bench_n1.err; cat $O/s2g_bench_n1.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --k 100 > $O/s2g_bench_n2_k100.json 2> $O/s2g_bench_n2_k100.err; cat $O/s2g_bench_n2_k100.json
