#!/bin/bash
# session-3 GPU pass L (1 GPU): k-blocks per pipeline stage of the single-CTA K3 kernel (1 / 2 / 4), same box, headline shape
cd "$(dirname "$0")/.."
O=gpurun_out
for LIB in librbod.so librbod_kbs1.so librbod_kbs4.so librbod.so; do
  RBOD_LIBRARY=$PWD/retrieval_based_object_detection_b200/$LIB timeout 300 python bench.py --no-cpu-baseline --steps 5 2>>$O/s3l.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$LIB', round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['parity']['ids_identical'], d['clocks']['sm_mhz'], d['clocks'].get('power_w'))"
done
for LIB in librbod.so librbod_kbs1.so; do
  RBOD_LIBRARY=$PWD/retrieval_based_object_detection_b200/$LIB timeout 300 python bench.py --rows 12500000 --dtype f16 --k 10 --sweep 1,128,256,512,1024,4096 2>>$O/s3l.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$LIB', ' '.join(f\"Q{r['Q']}:{r['p50_ms']}ms\" for r in d['sweep']))"
done
tail -3 $O/s3l.err
