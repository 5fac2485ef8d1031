#!/bin/bash
# ncu --set full capture of the two fused HashMLP kernels (encoder+decoder forward, decoder-backward+scatter), one B200.
# usage: scripts/gpu_ncu_fused.sh <tag>   (MRI_BATCH_ORDER etc. are taken from the environment)
TAG=${1:-fused}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-infer --no-e2e"
timeout 600 $CMD > gpurun_out/plain_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hashdecoder_mma' -s 8 -c 2 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | grep prof_$TAG
