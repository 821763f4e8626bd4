#!/bin/bash
# session-3 GPU pass G (1 GPU): what bounds mid-size batches?  DRAM traffic and L2 behaviour of K3 at Q = 128 / 256 / 512
# on the C5 shard shape, single-CTA kernel and CTA-pair kernel
cd "$(dirname "$0")/.."
O=gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum,sm__cycles_active.avg,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
for V in 0 2; do
  for Q in 128 256 512; do
    timeout 300 ncu --metrics $M --clock-control none -k regex:k3_cosine -c 6 --csv --log-file $O/s3g_v${V}_q$Q.csv \
      python tools/probe.py search --rows 12500000 --dim 768 --dtype f16 --queries $Q --k 10 --iters 1 --opt k3_variant=$V > $O/s3g_v${V}_q$Q.log 2>&1
    python - $O/s3g_v${V}_q$Q.csv $V $Q <<'PY'
import csv,sys
rows=[r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h=rows[0]; iid=h.index("ID"); mn=h.index("Metric Name"); mv=h.index("Metric Value"); mu=h.index("Metric Unit")
by={}
for r in rows[1:]:
    by.setdefault(r[iid],{})[r[mn]]=(r[mv],r[mu])
for k,v in by.items():
    print("variant",sys.argv[2],"Q",sys.argv[3],"launch",k," ".join(f"{m.split('__')[-1][:28]}={a}{u}" for m,(a,u) in v.items()))
PY
  done
done
