#!/bin/bash
# session-3 GPU pass H (1 GPU): full parity suite, C5 query sweep with the pre-pass skipped for small batches, K1/K2 defaults
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | cut -c1-300 | tee $O/s3h_pytest.log
timeout 900 python bench.py --rows 12500000 --dtype f16 --k 10 --sweep 1,2,4,8,16,32,64,128,256,512,1024,2048,4096,8192,16384,32768,65536 > $O/s3h_sweep_c5.json 2> $O/s3h.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/s3h_sweep_c5.json"))
for r in d["sweep"]: print(r["Q"], r["p50_ms"], r["qps"], r["bound"], r["frac_of_bound"], r["slices"])
PY
P=$O/s3h_probe.jsonl; : > $P
timeout 200 python tools/probe.py copy >> $P 2>>$O/s3h.err
timeout 100 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 --iters 10 >> $P 2>>$O/s3h.err
timeout 100 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 --iters 10 >> $P 2>>$O/s3h.err
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 >> $P 2>>$O/s3h.err
timeout 200 python tools/probe.py k2 --rows 1000000 --dim 768 --dtype f32 --classes 10000 --zipf >> $P 2>>$O/s3h.err
timeout 200 python tools/probe.py k2 --rows 4000000 --dim 768 --dtype bf16 --classes 10000 >> $P 2>>$O/s3h.err
cat $P
tail -3 $O/s3h.err
