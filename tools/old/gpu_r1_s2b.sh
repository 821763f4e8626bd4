#!/bin/bash
# session-2 GPU pass B: where does K3 lose time?  (epilogue off / loads only / variants)
cd "$(dirname "$0")/.."
O=gpurun_out
P="timeout 200 python tools/probe.py search --rows 4000000 --dim 768 --k 10 --queries 10000 --iters 3"
for v in 0 2; do
  for e in 0 1 2; do
    $P --opt k3_variant=$v --opt debug_epi=$e 2>&1 | tail -1
  done
done | tee $O/s2b_epi.jsonl
$P --opt k3_variant=1 2>&1 | tail -1 | tee -a $O/s2b_epi.jsonl
$P --opt k3_variant=0 --opt hybrid=0 2>&1 | tail -1 | tee -a $O/s2b_epi.jsonl
$P --opt k3_variant=0 --opt hybrid=0 --opt debug_epi=2 2>&1 | tail -1 | tee -a $O/s2b_epi.jsonl
$P --opt k3_variant=0 --opt l2_sync=0 --opt debug_epi=2 2>&1 | tail -1 | tee -a $O/s2b_epi.jsonl
timeout 200 python tools/probe.py search --rows 4000000 --dim 512 --k 10 --queries 10000 --iters 3 --opt debug_epi=0 2>&1 | tail -1 | tee -a $O/s2b_epi.jsonl
timeout 200 python tools/probe.py search --rows 4000000 --dim 512 --k 10 --queries 10000 --iters 3 --opt debug_epi=2 2>&1 | tail -1 | tee -a $O/s2b_epi.jsonl
