#!/bin/bash
# session-3 GPU pass W (1 GPU): where does the time go for a tiny gallery (10k queries x 10k delegate vectors)?
cd "$(dirname "$0")/.."
O=gpurun_out
for OPTS in "" "--opt l2_sync=0" "--opt tau_share=0" "--opt debug_epi=1" "--opt debug_epi=2" "--opt hybrid=0" "--opt k3_variant=1" "--opt k3_variant=2" "--opt l2_sync=0 --opt debug_epi=2"; do
  python tools/probe.py search --rows 10000 --dim 768 --dtype f32 --queries 10000 --k 5 --iters 20 $OPTS 2>>$O/s3w.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$OPTS'.ljust(44), 'ms', d['ms'], 'k3_ms', d['k3_ms'], 'slices', d['slices'])"
done
python tools/probe.py search --rows 10000 --dim 768 --dtype f32 --queries 1000 --k 5 --iters 20 --opt debug_epi=2 2>>$O/s3w.err | cut -c1-200
python tools/probe.py search --rows 10000 --dim 768 --dtype f32 --queries 128 --k 5 --iters 20 2>>$O/s3w.err | cut -c1-200
tail -2 $O/s3w.err
