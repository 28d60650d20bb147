#!/bin/bash
# Multi-GPU Barnes-Hut, replicated against partitioned build, at the BASELINE size (strong scaling)
# and at a larger N (weak scaling).  Usage: scripts/scale_bh_build.sh <world> "<N list>" [tag]
mkdir -p gpurun_out
w=$1; tag=${3:-bhbuild}
port=29700
for n in $2; do
  for b in replicated partitioned; do
    port=$((port+1))
    log=gpurun_out/${tag}_${b}_${n}_x${w}.log
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $w --workload barneshut --particles $n --bh-build $b --no-extra \
      --steps 5 --warmup 3 > $log 2>&1
    echo "== $b N=$n x$w rc=$?"
    grep -h '^{' $log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: d[k] for k in ('value', 'ms_per_step', 'comm_ms', 'build_ms', 'traverse_ms')}, 'e2e', d['e2e']['ms_per_step'])
"
  done
done
