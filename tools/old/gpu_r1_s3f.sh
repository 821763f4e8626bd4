#!/bin/bash
# session-3 GPU pass F (1 GPU): staggered start of the query tiles that share a slice (mid-size batches), headline check
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "search" 2>&1 | tail -4 | cut -c1-300 | tee $O/s3f_pytest.log
S=$O/s3f_stagger.txt; : > $S
for OPTS in "" "--opt stagger=4" "--opt stagger=4 --opt sync_window=2 --opt sync_lead=1" "--opt stagger=8 --opt sync_window=4 --opt sync_lead=1" "--opt stagger=16 --opt sync_window=4 --opt sync_lead=2" "--opt stagger=2 --opt sync_window=1 --opt sync_lead=1" "--opt stagger=32 --opt sync_window=8 --opt sync_lead=2" "--opt stagger=4 --opt l2_sync=0"; do
  timeout 300 python bench.py --rows 12500000 --dtype f16 --k 10 $OPTS --sweep 128,192,256,384,512,1024,2048 > $O/s3f_tmp.json 2>> $O/s3f.err
  python - "$OPTS" >> $S <<'PY'
import json,sys
d=json.load(open("gpurun_out/s3f_tmp.json"))
print(sys.argv[1] or "default", " ".join(f"Q{r['Q']}:{r['p50_ms']}ms/{r['frac_of_bound']}" for r in d["sweep"]))
PY
done
cat $S
for OPTS in "" "--opt stagger=4" "--opt stagger=16"; do
  timeout 300 python bench.py --no-cpu-baseline --steps 5 $OPTS 2>>$O/s3f.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('headline $OPTS', round(d['value']), d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"
done
tail -3 $O/s3f.err
