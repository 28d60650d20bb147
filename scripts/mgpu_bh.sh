#!/bin/bash
# Multi-GPU Barnes-Hut builds side by side (device-resident + e2e ms per step), no parity legs.
# Usage (GPU box): bash scripts/mgpu_bh.sh N "let partitioned replicated" [trace]
N=${1:-2}
mkdir -p gpurun_out
for b in ${2:-let partitioned replicated}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload barneshut --steps 8 --warmup 3 --bh-build $b --no-parity --no-extra \
    2>gpurun_out/mgb_${N}_$b.err | tail -1 > gpurun_out/mgb_${N}_$b.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/mgb_${N}_$b.json"))
    print("$b", "N=$N", {k: round(d[k], 3) for k in ("ms_per_step", "comm_ms", "build_ms", "traverse_ms")}, "e2e ms", round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("$b failed", e); print(open("gpurun_out/mgb_${N}_$b.err").read()[-1500:])
PY
done
if [ -n "$3" ]; then
  PCUDA_DEBUG=bh_let_trace=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus $N --workload barneshut --steps 1 --warmup 3 --bh-build let --no-parity --no-extra 2>&1 | grep "let rank" | tail -$N
fi
