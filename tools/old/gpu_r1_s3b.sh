#!/bin/bash
# session-3 GPU pass B (1 GPU): K1 knob sweep (CTAs per SM, L2 prefetch distance, streaming stores), L2-sharing
# window/lead sweep at mid-size batches (C5 shard shape), shadow test re-run
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "shadow or k1 or k2" 2>&1 | tail -5 | tee $O/s3b_pytest.log
P=$O/s3b_k1.jsonl; : > $P
for CT in 2 3 4 6; do for PF in 0 2 4; do for CS in 0 1; do
  export RBOD_K1_CTAS=$CT RBOD_K1_PF=$PF RBOD_K1_CS=$CS
  echo "{\"ctas\": $CT, \"pf\": $PF, \"cs\": $CS}" >> $P
  timeout 100 python tools/probe.py k1 --rows 8000000 --dim 768 --dtype bf16 --iters 10 2>>$O/s3b.err | head -1 >> $P
  timeout 100 python tools/probe.py k1 --rows 4000000 --dim 512 --dtype f32 --iters 10 2>>$O/s3b.err | head -1 >> $P
done; done; done
unset RBOD_K1_CTAS RBOD_K1_PF RBOD_K1_CS
python - <<'PY'
import json
cfg=None
for l in open("gpurun_out/s3b_k1.jsonl"):
    d=json.loads(l)
    if "ctas" in d: cfg=d
    else: print(cfg, d["out"], d["ms"], d["hbm_frac"])
PY
S=$O/s3b_sync.txt; : > $S
for OPTS in "" "--opt l2_sync=0" "--opt sync_window=1 --opt sync_lead=2" "--opt sync_window=2 --opt sync_lead=2" "--opt sync_window=4 --opt sync_lead=2" "--opt sync_window=8 --opt sync_lead=2" "--opt sync_window=4 --opt sync_lead=4"; do
  timeout 300 python bench.py --rows 12500000 --dtype f16 --k 10 $OPTS --sweep 128,192,256,384,512,1024,2048 > $O/s3b_tmp.json 2>> $O/s3b.err
  python - "$OPTS" >> $S <<'PY'
import json,sys
d=json.load(open("gpurun_out/s3b_tmp.json"))
print(sys.argv[1] or "default", " ".join(f"Q{r['Q']}:{r['p50_ms']}ms/{r['frac_of_bound']}" for r in d["sweep"]))
PY
done
cat $S
