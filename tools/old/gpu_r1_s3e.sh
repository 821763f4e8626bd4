#!/bin/bash
# session-3 GPU pass E (1 GPU): K5 with cp.async staging (tests + probe), K1 grid size probe
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "distance or k1 or shadow" 2>&1 | tail -8 | cut -c1-300 | tee $O/s3e_pytest.log
P=$O/s3e_probe.jsonl; : > $P
timeout 200 python tools/probe.py copy >> $P 2>$O/s3e.err
for CFG in "RBOD_K1_CTAS=8" "RBOD_K1_CTAS=16" "RBOD_K1_CTAS=32" "RBOD_K1_CTAS=64" "RBOD_K1_CTAS=100000" "RBOD_K1_CTAS=16"; do
  echo "{\"k1_cfg\": \"$CFG\"}" >> $P
  env $CFG timeout 100 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 --iters 10 2>>$O/s3e.err | head -1 >> $P
  env $CFG timeout 100 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 --iters 10 2>>$O/s3e.err | head -1 >> $P
done
timeout 300 python tools/probe.py dist --rows 1000000 --dim 512 --dtype f32 --queries 1,32,256,1024 --k 10 --iters 4 >> $P 2>>$O/s3e.err
timeout 300 python tools/probe.py dist --rows 1000000 --dim 768 --dtype bf16 --queries 256 --k 100 --iters 2 >> $P 2>>$O/s3e.err
cat $P
tail -3 $O/s3e.err
