#!/bin/bash
# A/B of the fused gradient-exchange kernel under torchrun (usage: scripts/gpu_ab_exchange.sh <ngpus>); prints ms/step and
# the device-time breakdown of the optimiser step (opening barrier, exchange kernel, closing barrier, gradient clear).
N=${1:-8}
mkdir -p gpurun_out
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
    bench.py --gpus $N --steps 100 --warmup 5 --no-workloads --no-infer --no-e2e --no-cpu-baseline > gpurun_out/abx_$tag.json 2> gpurun_out/abx_$tag.err
  python - <<PY
import json
for line in open("gpurun_out/abx_$tag.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("$tag: step %.4f ms  opt %.4f  exposed %.4f  parts %s" % (d["ms_per_step"], d["optimizer_step_ms"], d.get("exchange_exposed_ms") or 0, d.get("exchange_parts")))
PY
}
run sync MRI_DP_INKERNEL_SYNC=1
run barriers MRI_DP_INKERNEL_SYNC=0
if [ "$2" == "all" ]; then
run u1 MRI_DP_UNROLL=1
run u4 MRI_DP_UNROLL=4
run p2p MRI_DP_MULTIMEM=0
run mm MRI_DP_MULTIMEM=1
fi
