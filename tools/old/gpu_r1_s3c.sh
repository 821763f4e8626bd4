#!/bin/bash
# session-3 GPU pass C (1 GPU): full parity suite (K5 distances, shard sums), same-box copy bandwidth, K1 old/new
# A/B, K3 producer L2-prefetch sweep at small and mid batches + headline, K5 probe
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 | cut -c1-300 | tee $O/s3c_pytest.log
P=$O/s3c_probe.jsonl; : > $P
timeout 200 python tools/probe.py copy >> $P 2>$O/s3c.err
for CFG in "RBOD_K1_IMPL=f64" "RBOD_K1_CTAS=4" "RBOD_K1_CTAS=6" "RBOD_K1_CTAS=8" "RBOD_K1_CTAS=12" "RBOD_K1_CTAS=8 RBOD_K1_PF=1" "RBOD_K1_CTAS=8 RBOD_K1_PF=3"; do
  echo "{\"k1_cfg\": \"$CFG\"}" >> $P
  env $CFG timeout 100 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 --iters 10 2>>$O/s3c.err | head -1 >> $P
  env $CFG timeout 100 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 --iters 10 2>>$O/s3c.err | head -1 >> $P
done
timeout 300 python tools/probe.py dist --rows 1000000 --dim 512 --dtype f32 --queries 1,32,256 --k 10 --iters 4 >> $P 2>>$O/s3c.err
cat $P
S=$O/s3c_pf.txt; : > $S
for OPTS in "" "--opt l2_prefetch=2" "--opt l2_prefetch=4" "--opt l2_prefetch=8" "--opt l2_prefetch=16" "--opt hybrid=0" "--opt hybrid=0 --opt l2_prefetch=4"; do
  timeout 300 python bench.py --rows 12500000 --dtype f16 --k 10 $OPTS --sweep 1,16,128,192,256,384,512,1024,4096 > $O/s3c_tmp.json 2>> $O/s3c.err
  python - "$OPTS" >> $S <<'PY'
import json,sys
d=json.load(open("gpurun_out/s3c_tmp.json"))
print(sys.argv[1] or "default", " ".join(f"Q{r['Q']}:{r['p50_ms']}ms/{r['frac_of_bound']}" for r in d["sweep"]))
PY
done
cat $S
for OPTS in "" "--opt l2_prefetch=4" "--opt l2_prefetch=8"; do
  timeout 300 python bench.py --no-cpu-baseline --steps 5 $OPTS 2>>$O/s3c.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('headline $OPTS', round(d['value']), d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])"
done
