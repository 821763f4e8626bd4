#!/bin/bash
# One parameterised GPU job script (replaces the per-pass scripts of round 1, kept under tools/old/).
#   gpurun --timeout 900 -- 'bash tools/gpu_job.sh <tag> <step> [<step> ...]'
# Every step runs under its own `timeout`, appends to gpurun_out/<tag>_*.log and prints a short tail, so a step that
# fails or hangs costs its own limit and nothing else.  Steps:
#   tests[:expr]     pytest -m gpu (optionally -k expr)
#   smoke            __graft_entry__.smoke()
#   bench[:args]     python bench.py --no-cpu-baseline <args>   (args: comma separated, e.g. bench:--k,100,--steps,3)
#   fullbench[:args] python bench.py <args>                     (the driver's command line)
#   mbench:<N>[:args] torchrun --nproc-per-node N bench.py --gpus N <args>   (run under gpurun --gpus N)
#   sweep:<rows>:<dtype>:<Qlist>   bench.py --sweep on one shard (Qlist with '+' for ',')
#   probe:<args>     python tools/probe.py <args>  (comma separated)
#   copy             this box's copy bandwidth (tools/probe.py copy)
#   launches[:args]  ncu launch list of bench.py (after a plain run of the same command)
#   ncu:<kernel-regex>[:args]   ncu --set full of the named kernel in bench.py
#   py:<file>        python <file>
#   nccl:<N>         torchrun --nproc-per-node N tools/check_sharded_nccl.py   (run under gpurun --gpus N)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
tag=$1; shift
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -8
for step in "$@"; do
  name=${step%%:*}
  rest=""
  [[ "$step" == *:* ]] && rest=${step#*:}
  args=${rest//,/ }
  echo "=== $step"
  case $name in
    tests)
      if [ -n "$rest" ]; then
        timeout 900 python -m pytest tests -m gpu -x -q -k "$rest" > $O/${tag}_pytest.log 2>&1
      else
        timeout 1200 python -m pytest tests -m gpu -q > $O/${tag}_pytest.log 2>&1
      fi
      echo "rc=$?"; tail -12 $O/${tag}_pytest.log | cut -c1-400 ;;
    smoke)
      timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | cut -c1-300 ;;
    bench)
      n=$(ls $O/${tag}_bench*.json 2>/dev/null | wc -l)
      timeout 600 python bench.py --no-cpu-baseline $args > $O/${tag}_bench$n.json 2> $O/${tag}_bench$n.err
      echo "rc=$?"; tail -3 $O/${tag}_bench$n.err | cut -c1-300
      python tools/bench_brief.py $O/${tag}_bench$n.json ;;
    fullbench)
      n=$(ls $O/${tag}_full*.json 2>/dev/null | wc -l)
      timeout 900 python bench.py $args > $O/${tag}_full$n.json 2> $O/${tag}_full$n.err
      echo "rc=$?"; tail -3 $O/${tag}_full$n.err | cut -c1-300
      python tools/bench_brief.py $O/${tag}_full$n.json ;;
    mbench)
      IFS=: read -r ngpu margs <<< "$rest"
      n=$(ls $O/${tag}_mbench*.json 2>/dev/null | wc -l)
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $ngpu --master-addr 127.0.0.1 --master-port 29517 \
        bench.py --gpus $ngpu ${margs//,/ } > $O/${tag}_mbench$n.json 2> $O/${tag}_mbench$n.err
      echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM_THREADS" $O/${tag}_mbench$n.err | tail -5 | cut -c1-300
      python tools/bench_brief.py $O/${tag}_mbench$n.json ;;
    sweep)
      IFS=: read -r rows dtype qs <<< "$rest"
      timeout 600 python bench.py --rows $rows --dtype $dtype --sweep ${qs//+/,} > $O/${tag}_sweep_${rows}_${dtype}.json 2> $O/${tag}_sweep.err
      echo "rc=$?"; tail -2 $O/${tag}_sweep.err | cut -c1-300
      python tools/bench_brief.py $O/${tag}_sweep_${rows}_${dtype}.json ;;
    probe)
      timeout 600 python tools/probe.py $args 2>> $O/${tag}_probe.err | tee -a $O/${tag}_probe.jsonl | cut -c1-600
      tail -2 $O/${tag}_probe.err 2>/dev/null | cut -c1-300 ;;
    copy)
      timeout 200 python tools/probe.py copy 2>> $O/${tag}_probe.err | tee -a $O/${tag}_probe.jsonl | cut -c1-300 ;;
    launches)
      cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-configs $args"
      timeout 600 $cmd > $O/${tag}_plain.log 2>&1 &&
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${tag}_launches.csv $cmd > $O/${tag}_ncu_l.log 2>&1
      echo "rc=$?"; python tools/ncu_summary.py $O/${tag}_launches.csv 2>&1 | head -20 ;;
    ncu)
      IFS=: read -r kre nargs <<< "$rest"
      cmd="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --no-configs ${nargs//,/ }"
      timeout 600 $cmd > $O/${tag}_plain2.log 2>&1 &&
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$kre -s 4 -c 2 -f -o $O/${tag}_prof $cmd > $O/${tag}_ncu_f.log 2>&1
      echo "rc=$?"; tail -3 $O/${tag}_ncu_f.log | cut -c1-300 ;;
    nccl)
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $rest --master-addr 127.0.0.1 --master-port 29511 \
        tools/check_sharded_nccl.py > $O/${tag}_nccl.json 2> $O/${tag}_nccl.err
      echo "rc=$?"; grep -v "^W\|^\*\*\*\|OMP_NUM_THREADS" $O/${tag}_nccl.err | tail -15 | cut -c1-300; cut -c1-1500 $O/${tag}_nccl.json ;;
    py)
      timeout 900 python $args 2>&1 | tail -40 | cut -c1-400 ;;
    *) echo "unknown step $step" ;;
  esac
done
