#!/bin/bash
# session-3 GPU pass A (1 GPU): parity suite on the error-free-fp32 K1 / TwoSum K2, K1/K2 probes (old vs new
# accumulators), CTA-pair variant against the single-CTA kernel at mid-size batches (C5 shard shape)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/s3a_pytest.log
P=$O/s3a_probe.jsonl; : > $P
timeout 200 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 >> $P 2>$O/s3a_probe.err
timeout 200 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 >> $P 2>>$O/s3a_probe.err
for ACC in f64 ff; do
  export RBOD_K2_ACC=$ACC
  echo "{\"k2_acc\": \"$ACC\"}" >> $P
  timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 >> $P 2>>$O/s3a_probe.err
  timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --zipf >> $P 2>>$O/s3a_probe.err
  timeout 200 python tools/probe.py k2 --rows 4000000 --dim 768 --dtype bf16 --classes 10000 >> $P 2>>$O/s3a_probe.err
done
unset RBOD_K2_ACC
cat $P
for V in 0 2; do
  timeout 400 python bench.py --rows 12500000 --dtype f16 --k 10 --variant $V --sweep 64,128,192,256,384,512,768,1024,2048,4096 > $O/s3a_sweep_v$V.json 2> $O/s3a_sweep_v$V.err
  tail -2 $O/s3a_sweep_v$V.err
  python - <<PY
import json
d=json.load(open("$O/s3a_sweep_v$V.json"))
for r in d["sweep"]: print("variant $V", r["Q"], r["p50_ms"], r["qps"], r["frac_of_bound"], r["slices"])
PY
done
