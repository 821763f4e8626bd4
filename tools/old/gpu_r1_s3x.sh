#!/bin/bash
# session-3 GPU pass X (1 GPU): candidate-list length for tiny galleries (C3 centroid search), random and clustered delegates
cd "$(dirname "$0")/.."
O=gpurun_out
for OPTS in "" "--opt slack=11" "--opt slack=3" "--opt slack=11 --opt tau_share=0"; do
  python tools/probe.py search --rows 10000 --dim 768 --dtype f32 --queries 10000 --k 5 --iters 20 --check $OPTS 2>>$O/s3x.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('random   ', '$OPTS'.ljust(34), 'ms', d['ms'], 'k3_ms', d['k3_ms'], 'kc', d['kc'], 'fallback', d['fallback'], 'ids_equal', d.get('ids_equal'))"
done
timeout 600 python -m pytest tests -m gpu -q -k "c3" 2>&1 | tail -3 | cut -c1-200
tail -2 $O/s3x.err
