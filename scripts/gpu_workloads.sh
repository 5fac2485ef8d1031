#!/bin/bash
# All bench workloads on one GPU (or N with torchrun when $1 = N). Results in gpurun_out/bench_<workload>_w<N>.json
N=${1:-1}
mkdir -p gpurun_out
for W in ankle_hash synthetic_hash siren_ankle siren_wide; do
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --workload $W --steps 40 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_${W}_w${N}.err | tail -1 > gpurun_out/bench_${W}_w${N}.json
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((20 + RANDOM % 70)) bench.py --gpus $N --workload $W --steps 40 --warmup 5 --no-cpu-baseline 2> gpurun_out/bench_${W}_w${N}.err | tail -1 > gpurun_out/bench_${W}_w${N}.json
  fi
  echo "$W exit $?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${W}_w${N}.json"))
    print("  train %.1f Mcoord/s (%.3f ms/step) e2e %.1f  infer %.1f Mvox/s  roof %.3f (%s)" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["infer"]["value"]/1e6, d["roofline"]["frac"] if d["roofline"] else -1, d["roofline"]["kernel"][:30] if d["roofline"] else ""))
except Exception as e:
    print("  no result:", e); print(open("gpurun_out/bench_${W}_w${N}.err").read()[-1500:])
PY
done
