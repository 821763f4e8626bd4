#!/bin/bash
# session-3 GPU pass U (1 GPU): memcheck of the kernels added this session (K5, shard sums/finish, K1 shape), then the full suite
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "distance_collection_device_io or k2_shard_sums or (k1_normalize and 777) or (distance_collections_exact and 3000)" 2>&1 | tail -12 | cut -c1-250 | tee $O/s3u_memcheck.log
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-300 | tee $O/s3u_pytest.log
timeout 100 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 --iters 10 2>>$O/s3u.err | head -2 | cut -c1-140
timeout 100 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 --iters 10 2>>$O/s3u.err | head -2 | cut -c1-140
