#!/bin/bash
# session-3 GPU pass 4b (1 GPU): full suite with EUCLID on the tensor-core path, EUCLID / MANHATTAN throughput
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 | cut -c1-300
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 300 python tools/probe.py dist --rows 1000000 --dim 512 --dtype f32 --queries 1,32,1024,10000 --k 10 --iters 4 2>>$O/s4b.err | cut -c1-230
timeout 200 python bench.py --no-cpu-baseline --steps 3 2>>$O/s4b.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('headline', round(d['value']), round(d['roofline']['achieved'],1), d['parity']['ids_identical'])"
tail -2 $O/s4b.err
