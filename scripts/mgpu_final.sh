#!/bin/bash
# Round-end check on an 8-GPU box: the driver's own command at N = 8 and N = 4 (default line with parity and
# ride-alongs), then Barnes-Hut at N = 80M, locally essential trees against the partitioned build.
mkdir -p gpurun_out
for N in 8 4; do
  S=$(date +%s)
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/final_${N}.err | tail -1 > gpurun_out/final_${N}.json
  echo "N=$N: $(( $(date +%s) - S )) s"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/final_${N}.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "frac", round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"], 1))
    print("parity", d.get("parity", {}).get("ok"), "bh", d.get("bh"), "split", d.get("split"))
except Exception as e:
    print("failed", e); print(open("gpurun_out/final_${N}.err").read()[-2000:])
PY
done
for b in let partitioned; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 8 --workload barneshut --particles 80000000 --steps 3 --warmup 3 --bh-build $b --no-parity --no-extra \
    2>gpurun_out/final_80M_$b.err | tail -1 > gpurun_out/final_80M_$b.json
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/final_80M_$b.json"))
    print("80M $b", {k: round(d[k], 3) for k in ("ms_per_step", "comm_ms", "build_ms", "traverse_ms")}, "e2e ms", round(d["e2e"]["ms_per_step"], 3))
except Exception as e:
    print("80M $b failed", e); print(open("gpurun_out/final_80M_$b.err").read()[-1500:])
PY
done
