#!/bin/bash
# Multi-GPU check on one box: parity tests at this world size, then the Barnes-Hut builds side by side and
# the default bench line.  Usage (GPU box): bash scripts/mgpu_run.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_multigpu.py -x -q -m gpu 2>&1 | tail -5
for b in let replicated partitioned; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --workload barneshut --steps 5 --warmup 3 --bh-build $b --no-parity --no-extra \
    2>gpurun_out/mg_${N}_$b.err | tail -1 > gpurun_out/mg_${N}_$b.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/mg_${N}_$b.json"))
    print("$b", "N=$N", {k: round(d[k], 3) for k in ("ms_per_step", "comm_ms", "build_ms", "traverse_ms")}, "e2e ms", round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("$b failed", e); print(open("gpurun_out/mg_${N}_$b.err").read()[-1500:])
PY
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/mg_${N}_default.err | tail -1 > gpurun_out/mg_${N}_default.json
tail -c 1800 gpurun_out/mg_${N}_default.json; echo; tail -3 gpurun_out/mg_${N}_default.err
