#!/bin/bash
# session-3 GPU pass V (1 GPU): config C3's second half -- 10k queries against the 10k delegate vectors, top-5
cd "$(dirname "$0")/.."
O=gpurun_out
for ARGS in "--rows 10000 --queries 10000 --k 5 --dim 768 --dtype f32" "--rows 10000 --queries 10000 --k 5 --dim 768 --dtype bf16" "--rows 10000 --queries 1000 --k 5 --dim 768 --dtype f32" "--rows 100000 --queries 10000 --k 5 --dim 768 --dtype f32"; do
  timeout 200 python bench.py --no-cpu-baseline --steps 20 $ARGS 2>>$O/s3v.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$ARGS', 'ms/step', round(d['ms_per_step'],3), 'k3_ms', round(d['roofline']['kernel_ms'],3), 'TF', round(d['roofline']['achieved'],1), 'slices', d['config']['slices'], 'launches', d['gpu_launches'], d['parity']['ids_identical'])"
done
python tools/probe.py search --rows 10000 --dim 768 --dtype f32 --queries 10000 --k 5 --iters 20 2>>$O/s3v.err | cut -c1-300
tail -2 $O/s3v.err
