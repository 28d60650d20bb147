#!/bin/bash
# Locally essential trees at N GPUs: one walk at the end against the two-phase walk with R reserved SMs.
# Usage (GPU box): bash scripts/mgpu_overlap.sh N "0:16 1:16 1:8 1:32"   (overlap:reserve pairs)
N=${1:-8}
mkdir -p gpurun_out
for cfg in ${2:-0:16 1:16}; do
  ov=${cfg%%:*}; rs=${cfg##*:}
  PCUDA_DEBUG=bh_let_overlap=$ov,bh_let_reserve=$rs python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload barneshut --steps 10 --warmup 3 \
    --bh-build let --no-parity --no-extra 2>gpurun_out/ov_${N}_${ov}_${rs}.err | tail -1 > gpurun_out/ov_${N}_${ov}_${rs}.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ov_${N}_${ov}_${rs}.json"))
    print("overlap=$ov reserve=$rs N=$N", {k: round(d[k], 3) for k in ("ms_per_step", "comm_ms", "build_ms", "traverse_ms")}, "e2e ms", round(d["e2e"]["ms_per_step"], 3), "spread", d.get("step_spread"))
except Exception as e:
    print("$cfg failed", e); print(open("gpurun_out/ov_${N}_${ov}_${rs}.err").read()[-1500:])
PY
done
