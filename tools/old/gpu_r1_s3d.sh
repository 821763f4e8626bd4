#!/bin/bash
# session-3 GPU pass D (1 GPU): K5 transposed kernel (tests + probe), K1 fp64 vs error-free-fp32 on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "distance or k1 or shadow or smoke" 2>&1 | tail -8 | cut -c1-300 | tee $O/s3d_pytest.log
P=$O/s3d_probe.jsonl; : > $P
timeout 200 python tools/probe.py copy >> $P 2>$O/s3d.err
for CFG in "RBOD_K1_IMPL=f64" "RBOD_K1_IMPL=f64 RBOD_K1_CTAS=2" "RBOD_K1_IMPL=f64 RBOD_K1_CTAS=4" "RBOD_K1_IMPL=f64 RBOD_K1_CTAS=16" "RBOD_K1_IMPL=f64 RBOD_K1_PF=0" "RBOD_K1_IMPL=f64 RBOD_K1_PF=4" "RBOD_K1_IMPL=f64 RBOD_K1_CS=1" "RBOD_K1_IMPL=ff RBOD_K1_CTAS=12" "RBOD_K1_IMPL=f64"; do
  echo "{\"k1_cfg\": \"$CFG\"}" >> $P
  env $CFG timeout 100 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 --iters 10 2>>$O/s3d.err | head -1 >> $P
  env $CFG timeout 100 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 --iters 10 2>>$O/s3d.err | head -1 >> $P
done
timeout 300 python tools/probe.py dist --rows 1000000 --dim 512 --dtype f32 --queries 1,32,256,1024 --k 10 --iters 4 >> $P 2>>$O/s3d.err
timeout 300 python tools/probe.py dist --rows 1000000 --dim 768 --dtype bf16 --queries 256 --k 100 --iters 2 >> $P 2>>$O/s3d.err
cat $P
tail -3 $O/s3d.err
