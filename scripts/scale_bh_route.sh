#!/bin/bash
# Multi-GPU Barnes-Hut, partitioned build: accelerations routed by all-to-all against all-gather.
# Usage: scripts/scale_bh_route.sh <world> "<N list>" "<routes>" [tag]
mkdir -p gpurun_out
w=$1; tag=${4:-bhroute}
port=29800
for n in $2; do
  for r in $3; do
    port=$((port+1))
    log=gpurun_out/${tag}_partitioned_${r}_${n}_x${w}.log
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 \
      --master-port $port bench.py --gpus $w --workload barneshut --particles $n --bh-build partitioned \
      --bh-route $r --no-extra --steps 5 --warmup 3 > $log 2>&1
    echo "== partitioned/$r N=$n x$w rc=$?"
    grep -h '^{' $log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print({k: d[k] for k in ('value', 'ms_per_step', 'comm_ms', 'build_ms', 'traverse_ms')}, 'e2e', d['e2e']['ms_per_step'])
"
  done
done
