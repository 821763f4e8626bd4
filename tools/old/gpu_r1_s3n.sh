#!/bin/bash
# session-3 GPU pass N (1 GPU): k-blocks per stage chosen by the planner -- parity suite, then 2 vs 4 vs auto, three
# alternating repetitions per shape on the same box
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-300 | tee $O/s3n_pytest.log
run() { # name, args...
  local name=$1; shift
  for rep in 1 2 3; do for K in 2 4; do
    timeout 300 python bench.py --no-cpu-baseline --opt k3_kbs=$K "$@" 2>>$O/s3n.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$name kbs=$K', round(d['value']), round(d['roofline']['achieved'],1), d['parity']['ids_identical'], d['clocks']['sm_mhz'])"
  done; done
}
run "C2 1Mx512 f32 k10" --rows 1000000 --dim 512 --dtype f32
run "10Mx768 bf16 k10" 
run "10Mx768 bf16 Q2048" --queries 2048
run "10Mx768 bf16 Q1024" --queries 1024
timeout 300 python bench.py --no-cpu-baseline --k 100 2>>$O/s3n.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('k100 auto', round(d['value']), round(d['roofline']['achieved'],1), d['config']['fallback_queries_per_step'], d['parity']['ids_identical'])"
timeout 300 python bench.py --rows 12500000 --dtype f16 --k 10 --sweep 1,128,256,512,1024,2048,4096,16384 2>>$O/s3n.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C5 auto', ' '.join(f\"Q{r['Q']}:{r['p50_ms']}ms\" for r in d['sweep']))"
tail -3 $O/s3n.err
