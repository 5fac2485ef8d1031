#!/bin/bash
# N-GPU check of the headline workload: DP parity (fused forward/backward kernels + sharded Adam) and the ankle_hash bench.
N=$1
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 scripts/dp_parity.py 2>&1 | grep -E "^\{" | tail -1 | tee gpurun_out/dp_parity_w${N}.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline 2> gpurun_out/scale_w${N}_ankle_hash.err | tail -1 > gpurun_out/scale_w${N}_ankle_hash.json
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_w${N}_ankle_hash.json"))
    print("ankle_hash N=$N train %.1f Mcoord/s (%.3f ms/step) e2e %.1f  infer %.1f Mvox/s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["infer"]["value"]/1e6))
except Exception as e:
    print("no result:", e); print(open("gpurun_out/scale_w${N}_ankle_hash.err").read()[-800:])
PY
