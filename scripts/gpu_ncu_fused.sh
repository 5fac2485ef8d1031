#!/bin/bash
# ncu --set full capture of the two fused HashMLP kernels (encoder+decoder forward, decoder-backward+scatter), one B200.
mkdir -p gpurun_out
export MRI_FUSED_FORWARD=1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain_fused.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_fused.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hashdecoder_mma' -s 8 -c 4 -f -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_fused.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | grep prof_fused
