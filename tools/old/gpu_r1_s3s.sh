#!/bin/bash
# session-3 GPU pass S (1 GPU): distance tests with the 8192-row sample, cast bandwidth reference for K1's traffic mix
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "distance or k1" 2>&1 | tail -4 | cut -c1-300
timeout 200 python tools/probe.py copy 2>>$O/s3s.err
timeout 100 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 --iters 10 2>>$O/s3s.err | head -1
timeout 300 python tools/probe.py dist --rows 1000000 --dim 512 --dtype f32 --queries 32,256 --k 10 --iters 4 2>>$O/s3s.err
timeout 300 python tools/probe.py dist --rows 1000000 --dim 768 --dtype bf16 --queries 256 --k 100 --iters 2 2>>$O/s3s.err
