#!/bin/bash
# Scaling run on N GPUs: DP parity + hash and wide-SIREN workloads. Results in gpurun_out/scale_w<N>_*.json
N=$1
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 scripts/dp_parity.py 2>&1 | grep -E "^\{" | tail -1
for W in ankle_hash synthetic_hash siren_wide; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus $N --workload $W --steps 60 --warmup 5 --no-cpu-baseline 2> gpurun_out/scale_w${N}_${W}.err | tail -1 > gpurun_out/scale_w${N}_${W}.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_w${N}_${W}.json"))
    print("$W N=$N train %.1f Mcoord/s (%.3f ms/step) e2e %.1f  infer %.1f Mvox/s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["infer"]["value"]/1e6))
except Exception as e:
    print("$W no result:", e); print(open("gpurun_out/scale_w${N}_${W}.err").read()[-800:])
PY
done
