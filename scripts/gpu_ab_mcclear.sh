#!/bin/bash
# A/B of the gradient clear in multimem mode (N GPUs): multicast store by the slice owner vs local memset after the barrier.
N=$1
mkdir -p gpurun_out
for mc in 0 1; do
  MRI_DP_MULTIMEM=1 MRI_DP_MULTICAST_CLEAR=$mc MRI_BENCH_NO_CLOCKS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2960$mc bench.py --gpus $N --steps 150 --warmup 5 --no-cpu-baseline --no-infer 2> gpurun_out/mcclear${mc}_w$N.err | tail -1 > gpurun_out/mcclear${mc}_w$N.json
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/mcclear${mc}_w$N.json"))
    print("multicast_clear=$mc N=$N %.1f Mcoord/s (%.4f ms/step)" % (d["value"]/1e6, d["ms_per_step"]))
except Exception as e:
    print("no result:", e); print(open("gpurun_out/mcclear${mc}_w$N.err").read()[-800:])
PY
done
