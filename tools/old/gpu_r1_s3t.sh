#!/bin/bash
# session-3 GPU pass T (1 GPU): K1 critical-path experiment (4 fp64 chains + Newton rsqrt, block size), same box A/B
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "k1 or shadow or smoke" 2>&1 | tail -4 | cut -c1-300
timeout 200 python tools/probe.py copy 2>>$O/s3t.err
for CFG in "RBOD_K1_ILP=0" "RBOD_K1_ILP=1" "RBOD_K1_ILP=1 RBOD_K1_WARPS=4" "RBOD_K1_ILP=1 RBOD_K1_WARPS=2" "RBOD_K1_ILP=0 RBOD_K1_WARPS=4" "RBOD_K1_ILP=1 RBOD_K1_WARPS=4 RBOD_K1_CTAS=32" "RBOD_K1_ILP=0" "RBOD_K1_ILP=1"; do
  echo "cfg $CFG"
  env $CFG timeout 100 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 --iters 10 2>>$O/s3t.err | head -1 | cut -c1-120
  env $CFG timeout 100 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 --iters 10 2>>$O/s3t.err | head -1 | cut -c1-120
done
