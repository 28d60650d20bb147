#!/bin/bash
# Scaling evidence on one multi-GPU box: bench.py for the three sharded workloads at the given
# world sizes (each run bounded by its own timeout).  Usage: scripts/scale_run.sh "8 4" [tag]
mkdir -p gpurun_out
tag=${2:-scale}
nproc > gpurun_out/${tag}_env.txt; nvidia-smi -L >> gpurun_out/${tag}_env.txt
port=29600
for n in $1; do
  for w in bruteforce split barneshut; do
    port=$((port+1))
    if [ "$n" = "1" ]; then
      timeout 150 python bench.py --gpus 1 --workload $w --no-extra > gpurun_out/${tag}_${w}_${n}.log 2>&1
    else
      timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 \
        --master-port $port bench.py --gpus $n --workload $w --no-extra > gpurun_out/${tag}_${w}_${n}.log 2>&1
    fi
    echo "== $w x$n rc=$?"; grep -h '^{' gpurun_out/${tag}_${w}_${n}.log | cut -c1-330
  done
done
