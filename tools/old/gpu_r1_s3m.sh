#!/bin/bash
# session-3 GPU pass M (1 GPU): 4 k-blocks per stage as the default -- parity suite, sweep and other shapes against the
# 2-k-block build on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 | tee $O/s3m_pytest.log
for LIB in librbod.so librbod_kbs2.so; do
  export RBOD_LIBRARY=$PWD/retrieval_based_object_detection_b200/$LIB
  timeout 300 python bench.py --rows 12500000 --dtype f16 --k 10 --sweep 1,16,128,256,512,1024,4096,16384 2>>$O/s3m.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$LIB C5', ' '.join(f\"Q{r['Q']}:{r['p50_ms']}ms\" for r in d['sweep']))"
  timeout 300 python bench.py --no-cpu-baseline --k 100 2>>$O/s3m.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$LIB k100', round(d['value']), round(d['roofline']['achieved'],1), d['config']['fallback_queries_per_step'], d['parity']['ids_identical'])"
  timeout 300 python bench.py --no-cpu-baseline --rows 1000000 --dim 512 --dtype f32 2>>$O/s3m.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$LIB C2 1Mx512 f32', round(d['value']), round(d['roofline']['achieved'],1), d['parity']['ids_identical'])"
  timeout 300 python bench.py --no-cpu-baseline --rows 10000000 --dim 512 --dtype bf16 2>>$O/s3m.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$LIB 10Mx512 bf16', round(d['value']), round(d['roofline']['achieved'],1), d['parity']['ids_identical'])"
done
unset RBOD_LIBRARY
tail -3 $O/s3m.err
